"""SASS size per kernel of an object file / library (cuobjdump -sass, 16 B per instruction), largest first."""
import re, subprocess, sys
path = sys.argv[1] if len(sys.argv) > 1 else "pressio-demoapps_b200/build/engine.o"
pat = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
sizes, cur = {}, None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); sizes[cur] = 0; continue
    if cur and re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", line):
        sizes[cur] += 16
names = list(sizes)
dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
for n, d in sorted(zip(names, dem), key=lambda x: -sizes[x[0]]):
    if pat in d:
        print("%9.3f KB  %s" % (sizes[n] / 1024.0, d[:150]))
