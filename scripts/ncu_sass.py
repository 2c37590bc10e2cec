"""Developer aid: dynamic SASS mix of one kernel from an .ncu-rep (source page): executed counts per opcode,
and optionally the listing with per-instruction counts / stall samples."""
import collections, csv, io, subprocess, sys
rep, idx = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"], text=True)
lines = out.split('\n')
print(lines[0][:150])
rows = list(csv.reader(io.StringIO('\n'.join(lines[1:]))))
hdr = rows[0]
iS, iE, iSmp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
ops, smp = collections.Counter(), collections.Counter()
tot = 0
listing = []
for r in rows[1:]:
    if len(r) <= iE:
        continue
    s = r[iS].strip()
    parts = s.split()
    op = parts[1] if parts[0].startswith('@') else parts[0]
    op = op.split('.')[0]
    if not (r[iE] or "0").isdigit():
        continue
    n = int(r[iE] or 0)
    ops[op] += n
    smp[op] += int(r[iSmp] or 0)
    tot += n
    listing.append((r[0], n, int(r[iSmp] or 0), s))
print('total executed warp-insts', tot)
for op, n in ops.most_common(30):
    print(f'  {op:10s} {n:12d} {100*n/tot:5.1f}%   samples {smp[op]}')
if len(sys.argv) > 3:
    with open(sys.argv[3], 'w') as f:
        for a, n, sm_, s in listing:
            f.write(f'{n:10d} {sm_:6d}  {s}\n')
