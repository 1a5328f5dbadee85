"""Executed SASS opcode mix of one kernel: python tools/sass_ops.py report.ncu-rep kernel-regex [launch-skip]"""
import csv, subprocess, sys, io, collections
rep, kre = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kre,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
print(rows[0][1][:80])
hdr = rows[1]
si, ii, ti, sm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
ops = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
n_static = 0
for r in rows[2:]:
    if len(r) <= ti or r[0] in ("Address", "Kernel Name"): continue
    src = r[si].strip()
    parts = src.split()
    if not parts: continue
    op = parts[1] if parts[0].startswith('@') and len(parts) > 1 else parts[0]
    op = op.split('.')[0]
    n_static += 1
    o = ops[op]; o[0] += 1; o[1] += float(r[ii]); o[2] += float(r[ti]); o[3] += float(r[sm])
tot = sum(o[1] for o in ops.values()); tots = sum(o[3] for o in ops.values())
print(f"static SASS instructions {n_static}, executed warp-instructions {tot:.4g}, samples {tots:.0f}")
for op, o in sorted(ops.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"  {op:10s} static {o[0]:5d}  exec {100*o[1]/tot:5.1f}%  thr/inst {o[2]/max(o[1],1):5.1f}  samples {100*o[3]/max(tots,1):5.1f}%")
