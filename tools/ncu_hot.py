"""Top stall lines of a kernel: python tools/ncu_hot.py rep kernel-regex [min_frac]"""
import csv, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
minf = float(sys.argv[3]) if len(sys.argv) > 3 else 0.004
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
start = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
end = next((i for i in range(start + 1, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))
hdr = rows[start]; H = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[start + 1:end] if len(r) > H['stall_wait']]
reasons = [h for h in hdr if h.startswith('stall_') and '(' not in h]
tot = sum(int(r[H['# Samples']] or 0) for r in body)
print('total samples', tot, 'lines', len(body))
agg = {k: sum(int(r[H[k]] or 0) for r in body) for k in reasons}
print({k: v for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v})
# bucket by 5% of the instruction stream to see phases
nb = 24
per = max(1, len(body) // nb)
for b in range(0, len(body), per):
    seg = body[b:b + per]
    s = sum(int(r[H['# Samples']] or 0) for r in seg)
    ex = sum(int(r[H['Instructions Executed']] or 0) for r in seg)
    top = sorted(reasons, key=lambda k: -sum(int(r[H[k]] or 0) for r in seg))[:3]
    ops = {}
    for r in seg:
        t = r[H['Source']].split()
        op = t[1] if t and t[0].startswith('@') and len(t) > 1 else (t[0] if t else '?')
        if op.split('.')[0] in ('LDG', 'LDS', 'STS', 'STG', 'BAR', 'SYNCS', 'UTMALDG', 'LDL', 'STL', 'CALL', 'B2R'):
            ops[op.split('.')[0]] = ops.get(op.split('.')[0], 0) + 1
    print(f"[{b:5d}] samples {s:6d} ({100*s/tot:5.1f}%) exec {ex:10d} top {[(k[6:], sum(int(r[H[k]] or 0) for r in seg)) for k in top]} {ops}")
for n, r in enumerate(body):
    s = int(r[H['# Samples']] or 0)
    if s > tot * minf:
        top = max(reasons, key=lambda k: int(r[H[k]] or 0))
        print(n, r[H['Source']].strip()[:64], s, top, r[H[top]])
