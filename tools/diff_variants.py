"""Runs the same seeded job under two builds (ART_SO) and lists the events whose rows differ
(python tools/diff_variants.py <so_a> <so_b> [events]).  Each build runs in its own process."""
import os, subprocess, sys, json
import numpy as np
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)

def child(n, out):
    from adiabatic_raytracer_b200 import lib
    from adiabatic_raytracer_b200.params import ArtParams
    h = lib.Handle(ArtParams.default(theta_m=0.2), 0)
    r = h.run(1769, 1, n)
    rows = r["rows"]
    st = r["stats"]
    np.save(out, rows)
    smp = h.sample(1769, 1, n)
    np.save(out + ".x.npy", smp["x"])
    np.save(out + ".nf.npy", smp["n_failed"])
    print(json.dumps({k: int(st[k]) for k in ("n_rays", "n_final_photon", "n_acc", "n_rej", "n_rhs", "n_attempts")} | {"f_inx": int(r["f_inx"]), "rows": len(rows)}))
    h.close()

if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(int(sys.argv[2]), sys.argv[3])
        sys.exit(0)
    a, b = sys.argv[1], sys.argv[2]
    n = sys.argv[3] if len(sys.argv) > 3 else "1000000"
    outs = []
    for i, so in enumerate((a, b)):
        e = dict(os.environ)
        so, _, envs = so.partition("@")   # <library>[@KEY=VAL,KEY=VAL]: tuning environment of that run
        for kv in filter(None, envs.split(",")):
            k, _, v = kv.partition("=")
            e[k] = v
        if so != "default":
            e["ART_SO"] = so
        out = f"/tmp/diff_rows_{i}.npy"
        r = subprocess.run([sys.executable, __file__, "--child", n, out], env=e, capture_output=True, text=True, timeout=900)
        print(so, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-500:])
        outs.append(np.load(out))
    xa, xb = np.load("/tmp/diff_rows_0.npy.x.npy"), np.load("/tmp/diff_rows_1.npy.x.npy")
    fa, fb = np.load("/tmp/diff_rows_0.npy.nf.npy"), np.load("/tmp/diff_rows_1.npy.nf.npy")
    print("art_sample (ordered): positions equal", np.array_equal(xa, xb), "n_failed equal", np.array_equal(fa, fb), "sum n_failed", fa.sum(), fb.sum())
    A, B = outs
    for tag, R, x in (("a", A, xa), ("b", B, xb)):   # the sample position a pipeline row carries (columns 9..11) vs art_sample's
        ev = R[:, 0].astype(np.int64)
        same = np.all(R[:, 9:12] == x[:, ev - 1].T, axis=1)
        print(tag, "pipeline sample positions == art_sample bit for bit:", bool(same.all()), "rows off:", int((~same).sum()))
    ea, eb = set(A[:, 0].astype(np.int64)), set(B[:, 0].astype(np.int64))
    bad = []
    for e in sorted(ea | eb):
        pass
    # rows are ordered by (event, tree position): compare event by event
    def groups(R):
        ev = R[:, 0].astype(np.int64)
        idx = np.flatnonzero(np.diff(ev)) + 1
        return dict(zip(ev[np.r_[0, idx]], np.split(R, idx)))
    ga, gb = groups(A), groups(B)
    for e in sorted(set(ga) | set(gb)):
        ra, rb = ga.get(e), gb.get(e)
        if ra is None or rb is None or ra.shape != rb.shape or not np.array_equal(ra, rb):
            bad.append(int(e))
    print("events with different rows:", len(bad), bad[:50])
    for e in bad[:6]:
        for tag, g in (("a", ga), ("b", gb)):
            r = g.get(e)
            print(tag, e, None if r is None else r[:, :14].tolist())
        print("  art_sample x", xa[:, e - 1].tolist(), xb[:, e - 1].tolist(), "n_failed", fa[e - 1], fb[e - 1])
