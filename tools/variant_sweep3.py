"""Sampler skip-pass sweep: python tools/variant_sweep3.py [events] [reps]"""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = sys.argv[1] if len(sys.argv) > 1 else "3000000"
reps = sys.argv[2] if len(sys.argv) > 2 else "2"
V = os.path.join(root, "adiabatic_raytracer_b200", "variants")
cases = [("default (8 passes)", {})] + [(f"skip passes {k}", {"ART_SO": f"{V}/libart_skip{k}.so"}) for k in (3, 5, 12, 20)]
for name, env in cases:
    if "ART_SO" in env and not os.path.exists(env["ART_SO"]):
        print(f"{name:28s} (variant not built)"); continue
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "profile_step.py"), n, reps], env=e, capture_output=True, text=True, timeout=600)
    last = [l for l in r.stdout.splitlines() if l.startswith("pass") or l.startswith("{")]
    print(f"{name:28s} {last[-1] if last else 'FAILED: ' + r.stderr[-300:]}", flush=True)
