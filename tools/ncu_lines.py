"""Per-source-line hot spots from an ncu report: python tools/ncu_lines.py report.ncu-rep [top]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
kfilter = sys.argv[3] if len(sys.argv) > 3 else None; kname = ""
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fname = "?"; hdr = None; out = []
for r in csv.reader(io.StringIO(txt)):
    if not r: continue
    if r[0] in ("Kernel Name", "Function Name"): kname = r[1]; continue
    if kfilter and kfilter not in kname: continue
    if r[0] == "File Name": fname = (r[1] or "?").split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] == "" or len(r) < 10: continue
    try:
        g = lambda n: float(r[hdr.index(n)])
        out.append((g("# Samples"), g("Instructions Executed"), g("Thread Instructions Executed"), fname, r[0], r[1].strip()[:90],
                    g("stall_wait"), g("stall_no_inst"), g("stall_short_sb"), g("stall_branch_resolving"), g("stall_long_sb"), g("stall_mio"), g("stall_lg")))
    except Exception: pass
ts = sum(o[0] for o in out); ti = sum(o[1] for o in out)
print(f"total samples {ts:.0f} inst {ti:.3g} thr/inst {sum(o[2] for o in out)/ti:.1f}")
byfile = {}
for o in out: byfile.setdefault(o[3], [0,0,0]); byfile[o[3]][0]+=o[0]; byfile[o[3]][1]+=o[1]; byfile[o[3]][2]+=o[2]
for f,v in sorted(byfile.items(), key=lambda x:-x[1][0]): print(f"  {f:22s} samples {100*v[0]/ts:5.1f}%  inst {100*v[1]/ti:5.1f}%  thr/inst {v[2]/max(v[1],1):.1f}")
print("  smp%  inst%  thr/i  wait noinst shortsb branch longsb mio lg | file:line source")
for o in sorted(out, key=lambda x: -x[0])[:top]:
    s=max(o[0],1)
    print(f"  {100*o[0]/ts:4.1f}  {100*o[1]/ti:4.1f}  {o[2]/max(o[1],1):5.1f}  {100*o[6]/s:3.0f} {100*o[7]/s:3.0f} {100*o[8]/s:3.0f} {100*o[9]/s:3.0f} {100*o[10]/s:3.0f} {100*o[11]/s:3.0f} {100*o[12]/s:3.0f} | {o[3]}:{o[4]} {o[5]}")
