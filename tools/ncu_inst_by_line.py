"""Executed instructions per source line for one kernel: python tools/ncu_inst_by_line.py report.ncu-rep kernel-substring [top]"""
import csv, subprocess, sys, io
rep, kf = sys.argv[1], sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
kname = ""; hdr = None; out = []
for r in csv.reader(io.StringIO(txt)):
    if not r: continue
    if r[0] in ("Kernel Name", "Function Name"): kname = r[1]; hdr = None; continue
    if kf not in kname: continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] == "" or len(r) < 10: continue
    try:
        g = lambda n: float(r[hdr.index(n)])
        out.append((g("Instructions Executed"), g("Thread Instructions Executed"), g("# Samples"), r[0], r[1].strip()[:110]))
    except Exception: pass
ti = sum(o[0] for o in out); ts = sum(o[2] for o in out)
print(f"{kf}: executed warp-instructions {ti:.4g}, samples {ts:.0f}")
acc = 0
for o in sorted(out, key=lambda x: -x[0])[:top]:
    acc += o[0]
    print(f"  {100*o[0]/ti:5.2f}% (cum {100*acc/ti:5.1f})  thr {o[1]/max(o[0],1):4.1f}  smp {100*o[2]/max(ts,1):4.1f}%  | {o[3]:>4s} {o[4]}")
