"""Second tuning sweep (after the certificate / sampler changes): python tools/variant_sweep2.py [events] [reps]"""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = sys.argv[1] if len(sys.argv) > 1 else "3000000"
reps = sys.argv[2] if len(sys.argv) > 2 else "2"
cases = [("default", {})]
for ns in ("3", "6", "8"):
    cases.append((f"bulk_nstep={ns}", {"ART_BULK_NSTEP": ns}))
for mg in ("8", "12", "20", "24"):
    cases.append((f"min_go={mg}", {"ART_BULK_MIN_GO": mg}))
cases += [("nstep=6 min_go=12", {"ART_BULK_NSTEP": "6", "ART_BULK_MIN_GO": "12"}),
          ("nstep=8 min_go=8", {"ART_BULK_NSTEP": "8", "ART_BULK_MIN_GO": "8"}),
          ("pool=2^19", {"ART_POOL_SLOTS": "524288"}),
          ("pool=2^21", {"ART_POOL_SLOTS": "2097152"}),
          ("sample_block=2^17", {"ART_SAMPLE_BLOCK": "131072", "ART_POOL_SLOTS": "1048576"}),
          ("scan_grid=2", {"ART_SCAN_GRID_PER_SM": "2"}),
          ("scan_grid=6", {"ART_SCAN_GRID_PER_SM": "6"}),
          ("stage_grid=24", {"ART_STAGE_GRID_PER_SM": "24"}),
          ("stage_grid=6", {"ART_STAGE_GRID_PER_SM": "6"})]
for name, env in cases:
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "profile_step.py"), n, reps], env=e, capture_output=True, text=True, timeout=600)
    last = [l for l in r.stdout.splitlines() if l.startswith("pass") or l.startswith("{")]
    print(f"{name:28s} {last[-1] if last else 'FAILED: ' + r.stderr[-300:]}", flush=True)
