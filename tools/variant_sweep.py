"""Times one pipeline pass per build variant / tuning environment (python tools/variant_sweep.py [events] [reps]).
Each case runs in its own process (ART_SO selects the library; env knobs are read at art_create)."""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = sys.argv[1] if len(sys.argv) > 1 else "3000000"
reps = sys.argv[2] if len(sys.argv) > 2 else "2"
V = os.path.join(root, "adiabatic_raytracer_b200", "variants")
cases = [("default", {}),
         ("default stage_grid=12", {"ART_STAGE_GRID_PER_SM": "12"}),
         ("default bulk_nstep=2", {"ART_BULK_NSTEP": "2"}),
         ("default bulk_nstep=6", {"ART_BULK_NSTEP": "6"}),
         ("default min_go=8", {"ART_BULK_MIN_GO": "8"}),
         ("default min_go=24", {"ART_BULK_MIN_GO": "24"}),
         ("default pool=2^19", {"ART_POOL_SLOTS": "524288"}),
         ("default pool=2^21", {"ART_POOL_SLOTS": "2097152"}),
         ("default refill_min=pool/4", {"ART_REFILL_MIN": "262144"}),
         ("default refill_min=pool/64", {"ART_REFILL_MIN": "16384"}),
         ("mb10", {"ART_SO": f"{V}/libart_mb10.so", "ART_STAGE_GRID_PER_SM": "10"}),
         ("mb12 grid 8", {"ART_SO": f"{V}/libart_mb12.so"}),
         ("mb12 grid 12", {"ART_SO": f"{V}/libart_mb12.so", "ART_STAGE_GRID_PER_SM": "12"}),
         ("roll", {"ART_SO": f"{V}/libart_roll.so"}),
         ("roll_mb10 grid 10", {"ART_SO": f"{V}/libart_roll_mb10.so", "ART_STAGE_GRID_PER_SM": "10"}),
         ("roll_mb12 grid 12", {"ART_SO": f"{V}/libart_roll_mb12.so", "ART_STAGE_GRID_PER_SM": "12"}),
         ("roll_mb12 grid 24", {"ART_SO": f"{V}/libart_roll_mb12.so", "ART_STAGE_GRID_PER_SM": "24"})]
only = os.environ.get("SWEEP_ONLY")
for name, env in cases:
    if only and only not in name:
        continue
    if "ART_SO" in env and not os.path.exists(env["ART_SO"]):
        print(f"{name:28s} (variant not built)"); continue
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "profile_step.py"), n, reps], env=e, capture_output=True, text=True, timeout=600)
    last = [l for l in r.stdout.splitlines() if l.startswith("pass") or l.startswith("{")]
    print(f"{name:28s} {last[-1] if last else 'FAILED: ' + r.stderr[-300:]}", flush=True)
