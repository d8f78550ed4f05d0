"""Register-file read cycles of the FP64 instructions of a kernel, from an ncu source-page CSV
(ncu -i rep --page source --csv --kernel-name regex:<kernel>).  Model measured with micro/fp64_operands.cu on
B200: an FP64 instruction occupies its scheduler for max(2, number of distinct vector-register source operands)
cycles (DFMA r,r,r: 3; DMUL/DADD r,r: 2; an immediate / constant-bank / uniform-register operand is free)."""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
nsamp = float(sys.argv[2])
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
inst = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = []; inst.append(cur); continue
    if cur is not None: cur.append(r)
d = [r for r in inst[0][1:] if len(r) >= len(hdr)]
W = nsamp / 32
tot = collections.Counter(); cyc = 0.0; n_all = 0.0; n_fp = 0.0
for r in d:
    try: n = int(r[ix['Instructions Executed']])
    except ValueError: continue
    src = re.sub(r'^@!?U?P\d+\s+', '', r[ix['Source']].strip())
    op = src.split()[0].split('.')[0]
    n_all += n
    if op not in ('DFMA', 'DMUL', 'DADD', 'DSETP'): continue
    n_fp += n
    ops = src.split(None, 1)[1].split(',')
    srcs = ops[1:] if op != 'DSETP' else ops[2:]
    regs = set()
    for o in srcs:
        m = re.search(r'\bR(\d+)', o)
        if m and 'UR' not in o.split('R' + m.group(1))[0][-1:]: regs.add(m.group(1))
    k = max(2, len(regs))
    tot[(op, len(regs))] += n
    cyc += n * k
print(f"instructions per sample {n_all / W:.0f}, FP64 {n_fp / W:.0f}, other {(n_all - n_fp) / W:.0f}")
for (op, k), n in sorted(tot.items()): print(f"  {op} with {k} register sources: {n / W:7.1f} per sample")
print(f"FP64 register-read cycles per sample: {cyc / W:.0f}")
