#!/bin/bash
# Static look at K2's flat hot loop before spending GPU time: compiles ac_hash.cu for k = 27 only, prints registers /
# spills and the opcode histogram of the longest straight-line block of k2_hash<27, FLAT> (the 8-window block).
# usage: tools/sass_k2.sh [extra nvcc flags]
set -e
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
cd $root/auriclass_b200/csrc
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -DAC_ONLY_K27 "$@" -Xptxas -v -c ac_hash.cu -o $tmp/ac_hash.o 2> $tmp/ptxas.log || { cat $tmp/ptxas.log; exit 1; }
grep -A2 "k2_hashILi27ELb1" $tmp/ptxas.log | grep -E "Used|spill"
cuobjdump -sass $tmp/ac_hash.o | awk '/Function : /{p=0} /Function : .*k2_hashILi27ELb1/{p=1} p{print}' > $tmp/k2flat.sass
python3 - $tmp/k2flat.sass <<'PY'
import re, sys
from collections import Counter
ins = []
for l in open(sys.argv[1]):
    m = re.match(r'\s+/\*([0-9a-f]{4})\*/\s+(.*?);', l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
print(len(ins), "instructions in the kernel")
ctrl = [i for i, (a, t) in enumerate(ins) if re.search(r'\b(BRA|CALL|BSSY|BSYNC|EXIT|RET|WARPSYNC)\b', t)]
best, prev = (0, 0, 0), -1
for c in ctrl + [len(ins)]:
    if c - prev - 1 > best[0]: best = (c - prev - 1, prev + 1, c)
    prev = c
n, a, b = best
def cls(t):
    t = re.sub(r'^@!?U?P\d\s+', '', t); op = t.split()[0]
    if op.startswith('IMAD.WIDE') or op.startswith('IMAD.HI'): return 'IMAD.WIDE'
    for k in ('IMAD.MOV', 'IMAD.X', 'IMAD.IADD', 'IMAD.SHL'):
        if op.startswith(k): return k
    return 'IMAD' if op.startswith('IMAD') else op.split('.')[0]
c = Counter(cls(t) for _, t in ins[a:b])
print("longest straight block:", n, "instructions =", round(n / 8, 1), "per window (8 windows + the group's loads)")
print(sorted(c.items(), key=lambda x: -x[1]))
fma = c['IMAD.WIDE'] * 4 + sum(c[k] for k in ('IMAD', 'IMAD.X', 'IMAD.MOV', 'IMAD.IADD', 'IMAD.SHL')) * 2
alu = sum(v for k, v in c.items() if k in ('SHF', 'LOP3', 'IADD3', 'SEL', 'ISETP', 'PRMT', 'LEA', 'VIADD', 'PLOP3')) * 2
print("pipe cycles per warp and window: fmaheavy", fma / 8, " alu", alu / 8)
PY
