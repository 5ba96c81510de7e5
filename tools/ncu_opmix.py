"""Dynamic opcode mix + stall samples from an ncu source page:
   python tools/ncu_opmix.py file.ncu-rep kernel-regex"""
import csv
import subprocess
import sys
from collections import Counter

rep, kre = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# first kernel only
start = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
end = next((i for i in range(start + 1, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))
hdr = rows[start]
ie, ss, so = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
ops, samples = Counter(), Counter()
tot = 0
for r in rows[start + 1:end]:
    if len(r) <= ie:
        continue
    src = r[so].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    n = int(r[ie] or 0)
    ops[op] += n
    samples[op] += int(r[ss] or 0)
    tot += n
print("total warp-instructions", tot)
fma_pipe = sum(v for k, v in ops.items() if k.startswith(("IMAD", "HFMA2", "FFMA", "FMUL", "FADD")))
alu_pipe = sum(v for k, v in ops.items() if k.startswith(("IADD3", "LOP3", "SHF", "SEL", "ISETP", "MOV", "VIADD", "LEA", "PRMT", "IABS", "PLOP3")))
print(f"fma-pipe {fma_pipe} ({100*fma_pipe/tot:.1f}%)  alu-pipe {alu_pipe} ({100*alu_pipe/tot:.1f}%)")
for k, v in ops.most_common(28):
    print(f"{k:28s} {v:12d} {100*v/tot:6.2f}%  samples {samples[k]}")
