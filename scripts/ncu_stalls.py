"""Developer aid: per-instruction stall breakdown of one kernel (ncu source page): top instructions per stall reason."""
import csv, io, subprocess, sys
rep, idx = sys.argv[1], int(sys.argv[2])
out = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"], text=True)
lines = out.split('\n')
rows = list(csv.reader(io.StringIO('\n'.join(lines[1:]))))
hdr = rows[0]
n = len(rows[1:]) // 2
body = [r for r in rows[1:1 + n] if len(r) == len(hdr)]   # the listing is printed twice
reasons = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = {h: 0 for h in reasons}
for r in body:
    for h in reasons:
        v = r[hdr.index(h)]
        tot[h] += int(v) if v.isdigit() else 0
allsum = sum(tot.values())
print('samples by reason:', {h: tot[h] for h in sorted(tot, key=tot.get, reverse=True) if tot[h]})
iS = hdr.index('Source')
for h in sorted(tot, key=tot.get, reverse=True)[:int(sys.argv[3]) if len(sys.argv) > 3 else 4]:
    print(f'--- {h} ({tot[h]} = {100*tot[h]/allsum:.1f}%)')
    top = sorted(body, key=lambda r: -(int(r[hdr.index(h)]) if r[hdr.index(h)].isdigit() else 0))[:12]
    for r in top:
        print(f'   {r[hdr.index(h)]:>6s}  x{r[hdr.index("Instructions Executed")]:>8s}  {r[iS].strip()}')
