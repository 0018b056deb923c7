"""Register-bank view of a kernel's FFMA loop from `cuobjdump -sass`: counts FFMAs whose three source registers share
bank parity (a conflict unless a .reuse flag serves one of them), per region between the first and the last LDS.
python tools/sass_bank_check.py <object> <mangled function name>"""
import re, subprocess, sys
from collections import Counter
obj, fun = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True).stdout
ops = [m.group(1) for m in (re.search(r'^\s+/\*[0-9a-f]+\*/\s+(.*?)\s*;', l) for l in txt.split('\n')) if m]
lds = [k for k, o in enumerate(ops) if o.startswith('LDS')]
first, k = lds[0], lds[-1]
while k < len(ops) and not ops[k].startswith('SYNCS'):
    k += 1
seg = ops[first:k]
ff = [o for o in seg if o.startswith('FFMA')]
conf = noreuse = 0
for o in ff:
    regs = re.findall(r'(-?\|?)R(\d+)(\.reuse)?', o)
    if len(regs) < 4:
        continue
    srcs = regs[1:4]
    par = {int(r[1]) % 2 for r in srcs}
    if len(par) == 1:
        conf += 1
        if not any(r[2] for r in srcs):
            noreuse += 1
print("%d instructions between the first LDS and the slot release; FFMA %d, three sources of one parity %d (without any .reuse %d)" % (len(seg), len(ff), conf, noreuse))
print(Counter(o.split()[0] for o in seg).most_common(10))
