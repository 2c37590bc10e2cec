"""Developer aid: SASS instruction mix per kernel of an object file (cuobjdump -sass)."""
import collections, re, subprocess, sys
txt = subprocess.check_output(["cuobjdump", "-sass", sys.argv[1]], text=True)
for f in re.split(r'\n\s*Function : ', txt)[1:]:
    name = f.split('\n')[0]
    ops = collections.Counter()
    for line in f.split('\n'):
        m = re.search(r'/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m:
            ops[m.group(1).split('.')[0]] += 1
    print(name[:100], sum(ops.values()))
    print('   ', ops.most_common(28))
