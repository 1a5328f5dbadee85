import os, subprocess, sys
root = "/root/repo"
V = os.path.join(root, "adiabatic_raytracer_b200", "variants")
for rep in range(2):
  for name, env in (("new sampler", {}), ("old sampler (six fixed passes)", {"ART_SO": f"{V}/libart_oldsampler.so"})):
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "profile_step.py"), "3000000", "3"], env=e, capture_output=True, text=True, timeout=600)
    last = [l for l in r.stdout.splitlines() if l.startswith("pass")]
    print(f"{name:32s} {last[-1] if last else r.stderr[-300:]}", flush=True)
