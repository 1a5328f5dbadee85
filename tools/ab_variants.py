"""A/B of build variants: python tools/ab_variants.py name=path.so[,ENV=VAL...] ... [--events N] (default library: name=default)"""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = [a for a in sys.argv[1:] if not a.startswith("--")]
n = "3000000"
for a in sys.argv[1:]:
    if a.startswith("--events="):
        n = a.split("=", 1)[1]
for rep in range(2):
    for a in args:
        name, _, rest = a.partition("=")
        parts = rest.split(",")
        e = dict(os.environ)
        if parts[0] != "default":
            e["ART_SO"] = os.path.join(root, parts[0])
        for kv in parts[1:]:
            k, _, v = kv.partition("=")
            e[k] = v
        r = subprocess.run([sys.executable, os.path.join(root, "tools", "profile_step.py"), n, "3"], env=e, capture_output=True, text=True, timeout=600)
        last = [l for l in r.stdout.splitlines() if l.startswith("pass")]
        print(f"{name:28s} {last[-1] if last else r.stderr[-300:]}", flush=True)
