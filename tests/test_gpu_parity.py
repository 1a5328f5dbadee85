"""GPU parity tests (run on the B200 box: `pytest -m gpu`).  Every call goes through the C ABI
(include/art.h → libart_b200.so) and is compared with the CPU oracle on the same seeded inputs,
and with the committed golden fixtures (tests/golden).

Tolerances (FP64 everywhere):
  * pointwise kernels (RHS, weights, conversion probability): 1e-11 relative — closed-form
    gradients vs dual numbers, different operation order;
  * sampler: identical accept flags / crossing counts, positions 1e-10 relative;
  * rays: same scheme on both sides.  DP5 (integrator=1): round-off-level agreement for rays whose accept/reject
    sequence is identical (required for ≥ 97 % of rays, 1e-7 relative on final position), and
    truncation-level agreement (the integrator's own reltol 1e-7 / abstol 1e-6 error, measured
    2e-4 against the converged solution) for the few whose step sequence flips at the |B_z|
    kink of the GJ model.  Vern6 (integrator=0, the reference's scheme and the default): its error estimator has
    a round-off floor (b̃7 = −176.5, b̃8 = +173.0 on the nearly equal k7, k8 ⇒ ≈ 1e-9 absolute noise in EEst), so two
    implementations choose step sizes that differ at 1e-10 and a third of the rays end up with different reject counts
    at the kink; final states agree to round-off in the median, ≥ 70 % within 1e-6, the tail within the scheme's
    truncation error (5e-5).  At tight tolerance every ray agrees (test_propagate_tight_tolerance_matches_oracle).
"""
import json
import math
import os

import numpy as np
import pytest

from adiabatic_raytracer_b200.params import ArtParams

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def lib():
    from adiabatic_raytracer_b200 import build, lib as L
    build.build()
    L.load()
    return L


def handle(lib, **kw):
    return lib.Handle(ArtParams.default(**kw), 0)


def rel(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


def test_rhs_matches_oracle(lib, oracle):
    rng = np.random.default_rng(3)
    for kw in (dict(theta_m=0.2), dict(theta_m=0.9, omega_pul=2 * math.pi, flat=1), dict(theta_m=0.3, isotropic=1),
               dict(theta_m=0.2, melrose=0), dict(theta_m=0.9, omega_pul=2 * math.pi, flat=1, melrose=0)):
        p = ArtParams.default(**kw)
        h = lib.Handle(p, 0)
        n = 4096
        u = np.zeros((n, 7))
        u[:, 0] = 10.2 + 60 * rng.random(n)
        u[:, 1] = 0.05 + 3.0 * rng.random(n)
        u[:, 2] = -3 + 6 * rng.random(n)
        u[:, 3] = rng.normal(size=n) * 0.5
        u[:, 4] = rng.normal(size=n) * 0.5 * u[:, 0]
        u[:, 5] = rng.normal(size=n) * 0.5 * u[:, 0] * np.sin(u[:, 1])
        erg = np.full(n, 1.0000003e-5)
        u[:, 6] = -erg * (1 + 1e-6 * rng.normal(size=n))
        lnt = rng.uniform(-25, -4, n)
        for sp in (1, 0):
            if sp == 0:
                u[: n // 4, 0] = 1.0 + 8.9 * rng.random(n // 4)  # axions inside the star (interior metric)
            species = np.full(n, sp, np.int32)
            d_gpu = h.rhs(species, u, erg, lnt)
            d_ref = oracle.rhs(p, species, u, erg, lnt)
            scale = np.abs(d_ref).max(axis=1, keepdims=True) + 1e-300
            # (the non-Melrose form divides by Wm K² − ωp² κ, which can get small: 3e-11 there)
            assert (np.abs(d_gpu - d_ref) / scale).max() < (1e-11 if p.melrose or p.isotropic else 3e-11)
        h.close()


def test_sampler_from_uniforms_matches_oracle(lib, oracle):
    rng = np.random.default_rng(11)
    for kw in (dict(theta_m=0.2), dict(theta_m=1.1, Mass_a=2e-5, B0=2e14, omega_pul=3.0), dict(theta_m=0.2, thick_surface=0),
               dict(theta_m=0.2, isotropic=1)):
        p = ArtParams.default(**kw)
        h = lib.Handle(p, 0)
        U = rng.random((6000, 10))
        a = h.sample_from_uniforms(U)
        b = oracle.sample_from_uniforms(p, U)
        same = (a["accepted"] == b["accepted"]) & (a["n_cross"] == b["n_cross"])
        assert same.mean() >= 0.9995  # round-off-level grazing crossings, see test_philox_event_sampler_matches_oracle
        m = (b["accepted"] == 1) & same
        assert 0.1 < m.mean() < 0.5
        assert rel(a["x"][:, m], b["x"][:, m]).max() < 1e-10
        assert rel(a["v_loc"][:, m], b["v_loc"][:, m]).max() < 1e-12
        assert rel(a["v_inf"][:, m], b["v_inf"][:, m]).max() < 1e-14
        h.close()


def test_sampler_edge_cases(lib):
    """Lines that miss the surface, lines through the star, degenerate uniforms."""
    p = ArtParams.default(theta_m=0.2)
    h = lib.Handle(p, 0)
    U = np.full((4, 10), 0.5)
    U[0, 5] = 1.0 - 1e-16   # impact parameter = maxR: outside every crossing
    U[1, 5] = 0.0           # straight through the centre of the star
    U[2, 9] = 0.999999      # randInx = n_max
    U[3, 0] = 0.0           # θi = 0 (line along the rotation axis)
    out = h.sample_from_uniforms(U)
    assert out["accepted"][0] == 0 and out["n_cross"][0] == 0
    assert np.all(np.isfinite(out["x"]))
    h.close()


def test_philox_event_sampler_matches_oracle(lib, oracle):
    p = ArtParams.default(theta_m=0.2)
    h = lib.Handle(p, 0)
    a = h.sample(1769, 1, 3000)
    b = oracle.sample(p, 1769, 1, 3000)
    # a grazing crossing whose condition is at round-off level at a grid point can be seen on one side and
    # missed on the other (FMA contraction, libm): at most 1 event in 1000
    same = (a["n_failed"] == b["n_failed"]) & (a["n_cross"] == b["n_cross"])
    assert same.mean() >= 0.999
    assert rel(a["x"][:, same], b["x"][:, same]).max() < 1e-10
    # per-event streams: a later window reproduces the same events
    c = h.sample(1769, 1001, 500)
    assert np.array_equal(c["x"], a["x"][:, 1000:1500])
    h.close()


def test_weights_and_conversion_probability(lib, oracle):
    for kw in (dict(theta_m=0.2), dict(theta_m=0.2, flat=1), dict(theta_m=0.7, omega_pul=2 * math.pi, Mass_NS=1.4, rNS=12.0)):
        p = ArtParams.default(**kw)
        h = lib.Handle(p, 0)
        s = oracle.sample(p, 7, 1, 2000)
        a = h.weights(s["x"], s["v_loc"], s["v_inf"])
        b = oracle.weights(p, s["x"], s["v_loc"], s["v_inf"])
        for k in b:
            assert rel(a[k], b[k]).max() < 1e-11, k
        for bs, kk in ((1, b["k_init"]), (-1, -b["k_init"])):
            Pa = h.conv_prob(s["x"], kk, b["erg_inf"], bs)
            Pb = oracle.conv_prob(p, s["x"], kk, b["erg_inf"], bs)
            assert rel(Pa, Pb).max() < 1e-9
        h.close()


def test_golden_conversion_probability(lib):
    for c in json.load(open(os.path.join(G, "conv_prob.json"))):
        par = c["par"]
        p = ArtParams.default(theta_m=par["thm"], omega_pul=par["w"], B0=abs(par["B0"]), rNS=par["rNS"],
                              Mass_NS=par["MNS"], Mass_a=par["ma"], Ax_g=par["Ax_g"], rho_DM=par["rho_DM"],
                              n_maxSample=par["n_max"])
        h = lib.Handle(p, 0)
        x = np.array(c["x"]).reshape(3, 1)
        W = h.weights(x, np.array(c["v_loc"]).reshape(3, 1), np.array(c["v_inf"]).reshape(3, 1))
        rc, maxR = lib.max_radius(p)
        assert rel(W["k_init"][:, 0], np.array(c["k_init"])).max() < 1e-12
        assert rel(W["cos_w"][0], c["cos_w"]) < 1e-10
        assert rel(W["jac_gr"][0], c["jac_gr"]) < 1e-12
        assert rel(W["sln_prob"][0], c["sln_prob_over_maxR2"] * maxR**2) < 1e-10
        P = h.conv_prob(x, np.array(c["k_init"]).reshape(3, 1), [c["erg_inf"]], 1 if par["B0"] > 0 else -1)[0]
        assert rel(P, c["P_nonAD"]) < 1e-9
        h.close()
        h = lib.Handle(p.copy(isotropic=1), 0)  # isotropic switch (MR.jl:82-85, RT.jl:1420-1422)
        P = h.conv_prob(x, np.array(c["k_init"]).reshape(3, 1), [c["erg_inf"]], 1 if par["B0"] > 0 else -1)[0]
        assert rel(P, c["P_nonAD_iso"]) < 1e-9
        h.close()


def _ray_inputs(oracle, p, n, seed=5):
    s = oracle.sample(p, seed, 1, n)
    w = oracle.weights(p, s["x"], s["v_loc"], s["v_inf"])
    return s["x"], w["k_init"], w["erg_inf"]


@pytest.mark.parametrize("integrator", [0, 1], ids=["vern6", "dp5"])
@pytest.mark.parametrize("kw", [dict(theta_m=0.2), dict(theta_m=0.2, flat=1), dict(theta_m=0.6, omega_pul=2 * math.pi),
                                dict(theta_m=0.2, melrose=0)])
def test_propagate_matches_oracle(lib, oracle, kw, integrator):
    p = ArtParams.default(integrator=integrator, **kw)
    h = lib.Handle(p, 0)
    n = 1500
    x, k, erg = _ray_inputs(oracle, p, n)
    for sp, mc, bs in ((1, -1, 1), (0, 100000, -1)):
        k0 = k if bs == 1 else -k
        a = h.propagate(np.full(n, sp, np.int32), x, k0, erg, np.zeros(n), -np.ones(n), bs, mc)
        b = oracle.propagate(p, np.full(n, sp, np.int32), x, k0, erg, np.zeros(n), -np.ones(n), bs, mc)
        same = (a["status"] == b["status"]) & (a["n_cross"] == b["n_cross"])
        assert same.mean() > 0.99
        steps = same & (a["counters"][:, 0] == b["counters"][:, 0]) & (a["counters"][:, 1] == b["counters"][:, 1])
        if integrator == 1:
            assert steps.mean() > 0.97
        else:
            steps = same  # Vern6: see the module doc
            # the device skips the three dense-output stages of steps with a certified sign: fewer RHS than the literal oracle
            # (at least the 8 new stages of every attempt against the oracle's 8 + 3 per accepted step: ratio > 8/11)
            assert 0.73 * b["counters"][:, 2].mean() < a["counters"][:, 2].mean() < b["counters"][:, 2].mean()
        dx = np.linalg.norm(a["x_f"] - b["x_f"], axis=0) / np.linalg.norm(b["x_f"], axis=0)
        dk = np.linalg.norm(a["k_f"] - b["k_f"], axis=0) / np.linalg.norm(b["k_f"], axis=0)
        if sp == 1:
            # round-off (FMA contraction, libm) is amplified where a step lands next to the |B_z|
            # kink of the GJ model; bounded by the integrator's own truncation error (≈2e-4)
            assert np.percentile(dx[steps], 99.5) < 2e-3 and np.percentile(dk[steps], 99.5) < 2e-3
            assert np.median(dx[steps]) < 1e-9
            assert (dx[steps] < 1e-6).mean() > (0.9 if integrator == 1 else 0.7)
            esc = steps & (b["status"] == 0)
            assert np.percentile(rel(a["u7_f"][esc], b["u7_f"][esc]), 99.5) < 1e-6
        else:
            # back-traced axions: what matters are the crossings near the star
            assert np.median(dx[steps]) < 1e-9
            nc = b["n_cross"]
            worst = []
            for i in np.where(steps & (nc > 0) & (nc <= 16))[0][:200]:
                ra, rb = a["cross_rec"][i, : nc[i]], b["cross_rec"][i, : nc[i]]
                worst.append(np.abs(ra[:, :3] - rb[:, :3]).max())
            # (km; rays whose step sequence diverged — a fifth with Vern6 — differ by their truncation error, which the slow
            # axion's later crossings amplify: measured 5e-4 … 1e-3 km there)
            assert np.percentile(worst, 99) < (1e-4 if integrator == 1 else 2e-2) and np.median(worst) < 1e-8
    h.close()


@pytest.mark.parametrize("fixture", ["rays.json", "rays_nonmelrose.json", "rays_flat.json"])
def test_rays_against_dop853_golden(lib, fixture):
    """GPU integrator at tight tolerance vs scipy DOP853(rtol 1e-12) fixtures (second fixture: melrose = false; third:
    flat = true)."""
    worst = 0.0
    for ray in json.load(open(os.path.join(G, fixture))):
        par = ray["par"]
        p = ArtParams.default(theta_m=par["thm"], omega_pul=par["w"], B0=par["B0"], rNS=par["rNS"], Mass_NS=par["MNS"],
                              Mass_a=par["ma"], reltol=1e-12, ode_err=1e-14, melrose=0 if "nonmelrose" in fixture else 1,
                              flat=1 if "flat" in fixture else 0)
        h = lib.Handle(p, 0)
        R = h.propagate([ray["species"]], np.array(ray["x0"]).reshape(3, 1), np.array(ray["khat"]).reshape(3, 1),
                        [ray["erg"]], [0.0], [ray["dOmega"]], 1, 100000, 16)
        uf = np.array(ray["uf"])
        xg = uf[0] * np.array([math.sin(uf[1]) * math.cos(uf[2]), math.sin(uf[1]) * math.sin(uf[2]), math.cos(uf[1])])
        worst = max(worst, np.linalg.norm(R["x_f"][:, 0] - xg) / np.linalg.norm(xg))
        assert abs(R["u7_f"][0] - uf[6]) <= 1e-9 * abs(uf[6])
        gold = [c for c in ray["crossings"] if c["u"][0] >= 1.01 * par["rNS"] and c["lnt"] > -25.0]
        assert R["n_cross"][0] == len(gold)
        for gc, c in zip(R["cross_rec"][0], gold):
            assert abs(math.log(gc[6]) - c["lnt"]) < 1e-7
        h.close()
    assert worst < 1e-8


@pytest.mark.parametrize("kw", [dict(theta_m=0.2, save_mode=1), dict(theta_m=0.2, save_mode=0, Ax_g=1e-10, MC_nodes=2, max_nodes=12),
                                dict(theta_m=0.5, save_mode=1, omega_pul=2 * math.pi, Mass_a=2e-5),
                                dict(theta_m=0.2, save_mode=1, melrose=0), dict(theta_m=0.2, thick_surface=0),
                                dict(theta_m=0.2, flat=1, omega_pul=2 * math.pi),   # BASELINE configs[0] (flat, P = 1 s)
                                dict(theta_m=0.2, isotropic=1), dict(theta_m=0.2, flat=1, isotropic=1, melrose=0, thick_surface=0)])
def test_fused_pipeline_matches_oracle(lib, oracle, kw):
    p = ArtParams.default(**kw)
    h = lib.Handle(p, 0)
    n = 1500
    a = h.run(1769, 1, n)
    b = oracle.run(p, 1769, 1, n)
    ra, rb = a["rows"], b["rows"]
    # events whose trees have the same shape on both sides: same sequence of final species.
    # (a ray whose accept/reject sequence flips at a kink can gain or lose a crossing — see module doc)
    def signature(rows):
        sig = {}
        for ev_id, sp in zip(rows[:, 0].astype(int), rows[:, 1].astype(int)):
            sig.setdefault(ev_id, []).append(sp)
        return sig
    sa, sb = signature(ra), signature(rb)
    good = np.zeros(n + 2, bool)
    for e in range(1, n + 1):
        good[e] = sa.get(e) == sb.get(e)
    assert good[1 : n + 1].mean() > 0.98
    assert abs(a["f_inx"] - b["f_inx"]) <= 0.01 * b["f_inx"] + 2
    assert abs(a["tot_count"] - b["tot_count"]) <= 0.01 * b["tot_count"] + 2
    A = ra[good[ra[:, 0].astype(int)]]
    B = rb[good[rb[:, 0].astype(int)]]
    assert A.shape == B.shape
    assert np.array_equal(A[:, :2], B[:, :2])  # event id, species, in the reference's order
    ph = B[:, 1] == 1
    for col, tol in ((7, 1e-9), (9, 1e-10), (10, 1e-10), (11, 1e-10)):  # sln_prob, x0
        assert rel(A[:, col], B[:, col]).max() < tol
    dth = np.abs(A[ph, 2] - B[ph, 2])
    dph = np.abs((A[ph, 3] - B[ph, 3] + math.pi) % (2 * math.pi) - math.pi)
    # The sample points of the two sides differ at 1e-10 (sampler root finder), which is enough to decorrelate the step
    # sizes Vern6 picks in its noise-dominated regime (module doc): every ray ends within the difference of two truncation
    # errors of the other side's, not within round-off.  Measured on B200: median 6e-7, p90 1e-5.
    assert np.median(dth) < 3e-6 and np.median(dph) < 3e-6
    # measured on B200 (r01): percentiles 50/90/99/max = 2e-9 / 1e-5 / 2e-4 / 5e-4 — round-off amplified at
    # the |B_z| kink, bounded by the scheme's truncation error at reltol 1e-7 (7e-5 in angle, tests/golden)
    assert np.percentile(dth, 90) < 1e-4
    if p.flat and p.omega_pul > 2.0:
        # BASELINE configs[0] (flat, P = 1 s): the conversion surface sits at 50–100 km and a few per cent of the photons
        # are refracted strongly enough that round-off-level differences grow to 0.01–0.3 rad — the host-compiled device
        # code (no FMA contraction) against the oracle shows the same: 44 of 1293 photons above 1e-3, max 0.23
        assert (dth > 1e-3).mean() < 0.06
    else:
        assert dth.max() < 2e-2
    w_ok = rel(A[:, 8], B[:, 8]) < 1e-5
    # weights are products of conversion probabilities at the crossing points, which move with the truncation error of
    # the ray that found them (Vern6 on B200: 93 % within 1e-5 in GR, 77 % in the strongly refracting flat P = 1 s
    # configuration; 98 % within 1e-3)
    assert w_ok.mean() > (0.65 if p.flat and p.omega_pul > 2.0 else 0.82)
    assert (rel(A[:, 8], B[:, 8]) < 1e-3).mean() > 0.94
    if p.save_mode > 0:
        assert (A[:, 20] == B[:, 20]).mean() > 0.99 and (A[:, 21] == B[:, 21]).mean() > 0.99
    h.close()


@pytest.mark.parametrize("kw", [dict(theta_m=0.2, integrator=1), dict(theta_m=0.2, integrator=0, reltol=1e-11, ode_err=1e-13)],
                         ids=["dp5-production-tolerance", "vern6-tight-tolerance"])
def test_fused_pipeline_agrees_per_ray_where_the_scheme_is_not_noise_limited(lib, oracle, kw):
    """The other two corners of the parity statement: with DP5 (no cancellation in its error estimator) the pipeline agrees
    with the oracle to round-off in the median at the reference tolerances; with Vern6 at reltol 1e-11 both sides sit on the
    converged solution and EVERY photon agrees to 1e-6 in its final angles."""
    p = ArtParams.default(**kw)
    n = 600
    h = lib.Handle(p, 0)
    a = h.run(1769, 1, n)
    h.close()
    b = oracle.run(p, 1769, 1, n)
    ca = np.bincount(a["rows"][:, 0].astype(int), minlength=n + 1)
    cb = np.bincount(b["rows"][:, 0].astype(int), minlength=n + 1)
    good = ca == cb
    assert good[1:].mean() > 0.99
    A = a["rows"][good[a["rows"][:, 0].astype(int)]]
    B = b["rows"][good[b["rows"][:, 0].astype(int)]]
    assert np.array_equal(A[:, :2], B[:, :2])
    ph = B[:, 1] == 1
    dth = np.abs(A[ph, 2] - B[ph, 2])
    dph = np.abs((A[ph, 3] - B[ph, 3] + math.pi) % (2 * math.pi) - math.pi)
    assert np.median(dth) < 1e-7 and np.median(dph) < 1e-7
    if p.integrator == 0:
        assert dth.max() < 1e-6 and dph.max() < 1e-6
        assert rel(A[ph, 8], B[ph, 8]).max() < 1e-5      # weights (conversion probabilities at the crossing points)
        assert np.abs(A[ph, 12] - B[ph, 12]).max() < 1e-9  # Δω
    else:
        assert np.percentile(dth, 90) < 1e-4


def test_pipeline_is_deterministic_and_chunk_invariant(lib):
    p = ArtParams.default(theta_m=0.2)
    h = lib.Handle(p, 0)
    a = h.run(42, 1, 4000)
    b = h.run(42, 1, 4000)
    assert np.array_equal(a["rows"], b["rows"]) and a["f_inx"] == b["f_inx"]
    c1 = h.run(42, 1, 1500)
    c2 = h.run(42, 1501, 2500)
    assert np.array_equal(np.vstack([c1["rows"], c2["rows"]]), a["rows"])
    assert c1["f_inx"] + c2["f_inx"] == a["f_inx"] and c1["tot_count"] + c2["tot_count"] == a["tot_count"]
    h.close()


def test_histogram_matches_numpy(lib):
    p = ArtParams.default(theta_m=0.2)
    h = lib.Handle(p, 0)
    h.hist_config(8, 12)
    out = h.run(3, 1, 5000)
    r = out["rows"]
    H = h.hist_get()
    bins = H[:-8].reshape(2, 8, 12, 4)
    w = r[:, 8] * r[:, 7]
    for sp in (0, 1):
        m = r[:, 1] == sp
        ic = np.clip(np.floor((np.cos(r[m, 2]) + 1) * 0.5 * 8).astype(int), 0, 7)
        ip = np.clip(np.floor((r[m, 3] + math.pi) / (2 * math.pi) * 12).astype(int), 0, 11)
        ref = np.zeros((8, 12))
        np.add.at(ref, (ic, ip), w[m])
        cnt = np.zeros((8, 12))
        np.add.at(cnt, (ic, ip), 1.0)
        assert np.allclose(bins[sp, :, :, 0], ref, rtol=1e-9, atol=0)
        assert np.array_equal(bins[sp, :, :, 3], cnt)
    assert H[-8] == out["f_inx"] and H[-7] == len(r) and H[-6] == 5000
    h.close()


def test_full_size_properties(lib):
    """BASELINE config 2 (1e6 GR events): size-independent properties."""
    p = ArtParams.default(theta_m=0.2)
    h = lib.Handle(p, 0)
    n = 1_000_000
    out = h.run_device(1769, 1, n)
    st = out["stats"]
    assert st["n_events"] == n and st["n_rays"] >= 2 * n
    assert out["f_inx"] == (st["n_attempts"] - n) + st["n_final_photon"]
    assert sum(st["n_status"]) == st["n_rays"] and st["n_status"][3] == 0  # no maxiters
    rows = h.fetch_rows(out["n_rows"])
    assert rows.shape[0] == st["n_final_photon"] + st["n_final_axion"]
    assert np.all(np.diff(rows[:, 0]) >= 0) and rows[0, 0] >= 1 and rows[-1, 0] <= n
    assert np.all(np.isfinite(rows))
    ph = rows[:, 1] == 1
    assert np.all(rows[ph, 6] > 2.9e5)              # photons reach r ≈ c/ωPul: the observer sphere
    assert np.all((rows[:, 8] >= 0) & (rows[:, 8] <= 1.0))  # 0: a back-traced crossing with P = 1
    # per event the final weights cannot exceed the back-traced weight (≤ P0 ≤ 1)
    wsum = np.bincount(rows[:, 0].astype(np.int64), weights=rows[:, 8])
    assert wsum.max() <= 1.0 + 1e-12
    assert np.all((rows[ph, 12] > -1.5) & (rows[ph, 12] < -0.5))  # Δω stays O(−1) (MainRunner.jl:168)
    # acceptance of the sampler ≈ 0.23-0.24
    assert 0.2 < n / st["n_attempts"] < 0.28
    h.close()


def test_empty_and_error_paths(lib):
    p = ArtParams.default(theta_m=0.2)
    h = lib.Handle(p, 0)
    out = h.run(1, 1, 0)
    assert out["rows"].shape[0] == 0 and out["f_inx"] == 0
    h.close()
    with pytest.raises(lib.ArtError) as e:
        lib.Handle(ArtParams.default(theta_m=0.2, Mass_a=1e-4), 0)
    assert e.value.code == -3


def test_main_runner_tree_drop_in(lib, tmp_path):
    from adiabatic_raytracer_b200 import main_runner_tree
    tot, rows, fname, out = main_runner_tree(1e-5, 1e-12, 0.2, 1.0, 1e14, 10.0, 1.0, "Simple", 501, [1.0, 1.0],
                                             flat=False, isotropic=False, melrose=True, ode_err=1e-6, file_tag="t0",
                                             ntimes=3, saveMode=0, dir_tag=str(tmp_path / "results"), iseed=1769,
                                             return_rows=True)
    assert os.path.basename(fname) == ("tree_MassAx_1.0e-5_AxionG_1.0e-12_ThetaM_0.2_rotPulsar_1.0_B0_1.0e14_Ax_trajs_501"
                                       "_N_Times_3_num_cutoff_5_MC_nodes_5_max_nodes_50_t0.npy")
    back = np.load(fname)
    assert back.shape == rows.shape and back.shape[1] == 13 and np.array_equal(back, rows)
    assert back[:, 0].max() <= 500 and tot >= 500            # Ntajs − 1 events (MainRunner.jl:399,450)
    assert np.allclose(back[:, 7] * out["f_inx"], out["rows"][:, 7] * out["f_inx"])


def test_multi_gpu_entry_point_with_one_rank_equals_main_runner_tree(lib, tmp_path):
    """main_runner_tree_multi (the product's multi-GPU entry point) degenerates to main_runner_tree for world = 1; the
    world > 1 checks on real NCCL are tools/check_multi_gpu.py (profiles/r02_check_multi_gpu_*.json)."""
    from adiabatic_raytracer_b200 import main_runner_tree, main_runner_tree_multi
    args = (1e-5, 1e-12, 0.2, 1.0, 1e14, 10.0, 1.0, "Simple", 801, [1.0, 1.0])
    kw = dict(flat=False, isotropic=False, melrose=True, ode_err=1e-6, ntimes=3, iseed=1769)
    tot, rows, fname, out = main_runner_tree(*args, file_tag="a", dir_tag=str(tmp_path / "one"), return_rows=True, **kw)
    r = main_runner_tree_multi(*args, rank=0, world=1, device=0, file_tag="a", dir_tag=str(tmp_path / "multi"), **kw)
    assert os.path.basename(r["file"]) == os.path.basename(fname).replace("_a.npy", "_a0.npy")
    assert np.array_equal(np.load(r["file"]), np.load(fname))
    assert r["tot_count"] == tot and r["f_inx_total"] == out["f_inx"] and r["n_events_total"] == 800
    H = r["hist"]
    assert H[-8] == out["f_inx"] and H[-7] == rows.shape[0] and np.array_equal(H, r["hist_local"])


def test_propagate_tight_tolerance_matches_oracle(lib, oracle):
    """Route (b) of SURVEY §7: at reltol 1e-11 / abstol 1e-13 both sides sit on the converged
    solution, so step-sequence sensitivity disappears and EVERY ray must agree."""
    p = ArtParams.default(theta_m=0.2, reltol=1e-11, ode_err=1e-13)
    h = lib.Handle(p, 0)
    n = 400
    x, k, erg = _ray_inputs(oracle, p, n, seed=9)
    for sp, mc, bs in ((1, -1, 1), (0, 100000, -1)):
        k0 = k if bs == 1 else -k
        a = h.propagate(np.full(n, sp, np.int32), x, k0, erg, np.zeros(n), -np.ones(n), bs, mc)
        b = oracle.propagate(p, np.full(n, sp, np.int32), x, k0, erg, np.zeros(n), -np.ones(n), bs, mc)
        assert np.array_equal(a["status"], b["status"]) and np.array_equal(a["n_cross"], b["n_cross"])
        dx = np.linalg.norm(a["x_f"] - b["x_f"], axis=0) / np.linalg.norm(b["x_f"], axis=0)
        dk = np.linalg.norm(a["k_f"] - b["k_f"], axis=0) / np.linalg.norm(b["k_f"], axis=0)
        tol = 1e-6 if sp == 1 else 1e-4   # slow axions: δv∞/v∞ amplifies by (v_loc/v∞)² ≈ 5e5
        assert dx.max() < tol and dk.max() < tol, (sp, dx.max(), dk.max())
        assert np.median(dx) < 1e-9
    h.close()


def _binned(rows, n_ct=16):
    """Σ weight·sln_prob of the final photons in n_ct equal-cosθf viewing-angle bins (plot/flux.py:38-52)."""
    ph = rows[:, 1] == 1
    ic = np.clip(np.floor((np.cos(rows[ph, 2]) + 1) * 0.5 * n_ct).astype(int), 0, n_ct - 1)
    out = np.zeros(n_ct)
    np.add.at(out, ic, rows[ph, 8] * rows[ph, 7])
    return out


@pytest.mark.parametrize("ma", [2e-6, 1e-5, 3e-5])
def test_mass_sweep_binned_flux_matches_oracle(lib, oracle, ma):
    """BASELINE config 3 at test size: axion-mass sweep, 16 viewing-angle bins of the final photon direction.
    Same seeded events on both sides ⇒ the binned flux agrees far inside its Monte-Carlo error."""
    p = ArtParams.default(theta_m=0.2, Mass_a=ma)
    h = lib.Handle(p, 0)
    h.hist_config(16, 1)
    n = 1200
    a = h.run(1769, 1, n)
    b = oracle.run(p, 1769, 1, n)
    fa, fb = _binned(a["rows"]), _binned(b["rows"])
    H = h.hist_get()[:-8].reshape(2, 16, 1, 4)
    assert np.allclose(H[1, :, 0, 0], fa, rtol=1e-9)             # device histogram = binning of the device rows
    assert abs(fa.sum() - fb.sum()) <= 0.02 * fb.sum()
    # per bin: within 3 sigma of the Monte-Carlo error of that bin (sigma² = Σ w², taken from the oracle rows)
    ph = b["rows"][:, 1] == 1
    ic = np.clip(np.floor((np.cos(b["rows"][ph, 2]) + 1) * 0.5 * 16).astype(int), 0, 15)
    var = np.zeros(16)
    np.add.at(var, ic, (b["rows"][ph, 8] * b["rows"][ph, 7]) ** 2)
    assert np.all(np.abs(fa - fb) <= 3.0 * np.sqrt(var) + 1e-3 * fb.sum())
    assert abs(a["f_inx"] - b["f_inx"]) <= 0.01 * b["f_inx"] + 2
    h.close()


def test_mass_above_the_surface_limit_has_no_rays(lib):
    """Top of the 1e-6…1e-4 eV sweep: maxR < rNS ⇒ "Too small Max R", no output (MainRunner.jl:389-396)."""
    with pytest.raises(lib.ArtError) as e:
        lib.Handle(ArtParams.default(theta_m=0.2, Mass_a=6e-5), 0)
    assert e.value.code == -3


@pytest.mark.parametrize("phase", [0, 17, 40])
def test_rotation_phase_resolved_run_matches_oracle(lib, oracle, phase):
    """BASELINE config 4 at test size: field time t0 = 2π·phase/(64 Ω) in sampler, weights and conversion
    probability (generalises the reference's t0 / fix_time knobs, RT.jl:1480, GS.jl:162)."""
    w = 1.0
    p = ArtParams.default(theta_m=0.2, omega_pul=w, t0_sample=2 * math.pi * phase / (64 * w))
    h = lib.Handle(p, 0)
    n = 600
    sa, sb = h.sample(7, 1, n), oracle.sample(p, 7, 1, n)
    assert np.array_equal(sa["n_failed"], sb["n_failed"]) and np.array_equal(sa["n_cross"], sb["n_cross"])
    assert rel(sa["x"], sb["x"]).max() < 1e-9
    a, b = h.run(7, 1, n), oracle.run(p, 7, 1, n)
    ida, idb = a["rows"][:, 0].astype(int), b["rows"][:, 0].astype(int)
    good = np.bincount(ida, minlength=n + 2) == np.bincount(idb, minlength=n + 2)
    assert good[1:n + 1].mean() > 0.98
    A, B = a["rows"][good[ida]], b["rows"][good[idb]]
    assert np.array_equal(A[:, :2], B[:, :2])
    assert rel(A[:, 7], B[:, 7]).max() < 1e-9                    # sln_prob (cos_w, g_det at the phase's field)
    assert np.median(np.abs(A[:, 2] - B[:, 2])) < 3e-6   # (Vern6 at the reference tolerances: see test_fused_pipeline_matches_oracle)
    if phase:
        # a different phase is a different magnetosphere: the sampled points must differ from phase 0
        s0 = lib.Handle(ArtParams.default(theta_m=0.2, omega_pul=w), 0)
        x0 = s0.sample(7, 1, n)["x"]
        s0.close()
        assert np.abs(x0 - sa["x"]).max() > 1e-3
    h.close()


def test_pinned_row_buffer_and_evaluation_counters(lib):
    """art_host_alloc / art_host_free: art_run into a page-locked buffer gives the same rows as a pageable one;
    n_cond_eval counts the conditions actually evaluated (≤ the reference's count, the rest is the far-field
    shortcut)."""
    p = ArtParams.default(theta_m=0.2)
    h = lib.Handle(p, 0)
    n = 3000
    a = h.run(11, 1, n)
    with lib.PinnedRows(n * 3 + 1024, h.ncols) as pin:
        b = h.run(11, 1, n, rows_out=pin.array)
        assert np.array_equal(a["rows"], b["rows"])
    st = a["stats"]
    assert 0 < st["n_cond_eval"] < st["n_cond"]
    assert st["n_cond_eval"] > 0.1 * st["n_cond"]  # (18 % with the dark / lit / in-star / far-field certificates)
    h.close()


def test_binned_spectra_histogram_matches_numpy(lib):
    """BASELINE config 3 "binned radio spectra": 16 viewing-angle bins × energy bins in Δω (row column 12)."""
    p = ArtParams.default(theta_m=0.2)
    h = lib.Handle(p, 0)
    n_ct, n_dw, lo, hi = 16, 12, -1.0 - 3e-6, -1.0 + 3e-6
    h.hist_config(n_ct, 1, n_dw, lo, hi)
    out = h.run(5, 1, 4000)
    r = out["rows"]
    H = h.hist_get()
    assert H.size == 2 * n_ct * n_dw * 4 + 8
    bins = H[:-8].reshape(2, n_ct, 1, n_dw, 4)
    w = r[:, 8] * r[:, 7]
    for sp in (0, 1):
        m = r[:, 1] == sp
        ic = np.clip(np.floor((np.cos(r[m, 2]) + 1) * 0.5 * n_ct).astype(int), 0, n_ct - 1)
        iw = np.clip(np.floor((r[m, 12] - lo) * (n_dw / (hi - lo))).astype(int), 0, n_dw - 1)
        ref = np.zeros((n_ct, n_dw))
        np.add.at(ref, (ic, iw), w[m])
        cnt = np.zeros((n_ct, n_dw))
        np.add.at(cnt, (ic, iw), 1.0)
        assert np.array_equal(bins[sp, :, 0, :, 3], cnt)
        assert np.allclose(bins[sp, :, 0, :, 0], ref, rtol=1e-9, atol=0)
    assert bins[1, :, 0, :, 3].sum() == (r[:, 1] == 1).sum()
    ph = r[:, 1] == 1
    assert (np.abs(r[ph, 12] + 1.0) < 3e-6).mean() > 0.5  # the chosen window holds most of the photon line
    h.close()


def test_save_mode_2_rows_event_table_and_text_files(lib, oracle, tmp_path):
    """saveMode 2: 31-column rows, event table and the clear-text final_* / event_* files
    (MainRunner.jl:435-444, 593-610, 689-701) against the oracle."""
    p = ArtParams.default(theta_m=0.2, save_mode=2, Ax_g=1e-10)
    n = 600
    h = lib.Handle(p, 0)
    a = h.run(1769, 1, n)
    ev = h.fetch_events(n)
    h.close()
    b = oracle.run(p, 1769, 1, n)
    assert a["rows"].shape[1] == 31 and ev.shape == (n, 12)
    assert np.array_equal(ev[:, 0], b["events"][:, 0])
    assert rel(ev[:, 1:11], b["events"][:, 1:11]).max() < 1e-9          # v_inf, sln_prob, x, k_init
    assert (ev[:, 11] == b["events"][:, 11]).mean() > 0.99              # count (tree shape, see module doc)
    ca = np.bincount(a["rows"][:, 0].astype(int), minlength=n + 1)
    cb = np.bincount(b["rows"][:, 0].astype(int), minlength=n + 1)
    good = ca == cb
    assert good[1:].mean() > 0.99
    A = a["rows"][good[a["rows"][:, 0].astype(int)]]
    B = b["rows"][good[b["rows"][:, 0].astype(int)]]
    dk = rel(A[:, 29], B[:, 29])
    assert np.median(dk) < 1e-6 and dk.max() < 2e-3   # Vern6: agreement to the difference of two truncation errors (module doc)
    assert (A[:, 30] == 0).sum() == (B[:, 30] == 0).sum() and np.any(B[:, 30] > 0)
    m = B[:, 30] > 0
    assert np.median(rel(A[m, 30], B[m, 30])) < 1e-6
    # the drop-in writes the files the reference writes
    from adiabatic_raytracer_b200 import main_runner_tree
    tot, rows, fname, out = main_runner_tree(1e-5, 1e-10, 0.2, 1.0, 1e14, 10.0, 1.0, "Simple", n + 1, [1.0, 1.0], flat=False,
                                             melrose=True, file_tag="s2", saveMode=2, dir_tag=str(tmp_path / "res"),
                                             iseed=1769, return_rows=True)
    assert np.load(fname).shape == (a["rows"].shape[0], 29)
    fin = open(tmp_path / "res" / "event" / "final_s2").read().splitlines()
    evl = open(tmp_path / "res" / "event" / "event_s2").read().splitlines()
    assert len(fin) == a["rows"].shape[0] and len(evl) == n
    f0 = fin[0].split()
    assert len(f0) == 10 and f0[0] == "1" and f0[2] in ("0", "1") and float(f0[1]) == a["rows"][0, 13]
    e0 = evl[0].split()
    assert len(e0) == 19 and e0[0] == "1" and float(e0[5]) == ev[0, 5] and float(e0[8]) == -ev[0, 8] and e0[-1] == str(int(ev[0, 11]))


def _load_tree_file(path):
    """Parser of saveNode's layout (MainRunner.jl:17-65): header; xc / yc / zc / tc or three "-" lines; x / y / z / times."""
    nodes = []
    lines = open(path).read().split("\n")
    i = 0
    while i < len(lines) and lines[i].strip():
        sp, w, pr, pw = lines[i].split()
        nd = dict(species=sp, weight=float(w), prob=float(pr), parent_weight=float(pw))
        if lines[i + 1].strip() == "-":
            assert lines[i + 2].strip() == "-" and lines[i + 3].strip() == "-"
            nd["cross"] = np.zeros((0, 4))
            i += 4
        else:
            nd["cross"] = np.array([[float(v) for v in lines[i + 1 + c].split()] for c in range(4)]).T
            i += 5
        nd["traj"] = np.array([[float(v) for v in lines[i + c].split()] for c in range(4)]).T
        i += 4
        nodes.append(nd)
    return nodes


def test_save_mode_3_tree_files(lib, tmp_path):
    """saveMode 3: one tree_<tag><event> file per event (MainRunner.jl:574-578) in saveNode's layout, trajectories at the
    `ntimes` saveat points.  Checked against the rows of the same run and against the host-compiled sampler."""
    from adiabatic_raytracer_b200 import main_runner_tree
    n, nt = 120, 25
    d = str(tmp_path / "res")
    tot, rows, fname, out = main_runner_tree(1e-5, 1e-10, 0.2, 1.0, 1e14, 10.0, 1.0, "Simple", n + 1, [1.0, 1.0], flat=False,
                                             melrose=True, file_tag="tr", saveMode=3, ntimes=nt, dir_tag=d, iseed=1769,
                                             return_rows=True)
    assert "_N_Times_25_" in fname and np.load(fname).shape[1] == 29
    files = sorted(os.listdir(os.path.join(d, "tree")), key=lambda s: int(s[len("tree_tr"):]))
    assert files == [f"tree_tr{e}" for e in range(1, n + 1)]
    full = out["rows"]                    # 31 columns, sln_prob not normalised
    n_nodes = 0
    for e in (1, 2, 3, 17, 60, n):
        nodes = _load_tree_file(os.path.join(d, "tree", f"tree_tr{e}"))
        n_nodes += len(nodes)
        R = full[full[:, 0] == e]
        # node 0: the back-traced axion, weight = Π (1 − P_j), prob = P0 (rows: column 25 = P0·weight, column 28 = P0)
        b = nodes[0]
        assert b["species"] == "axion" and b["parent_weight"] == -1.0
        assert abs(b["weight"] * b["prob"] - R[0, 25]) <= 1e-12 * R[0, 25] and abs(b["prob"] - R[0, 28]) <= 1e-12 * R[0, 28]
        assert np.abs(b["traj"][0, :3] - R[0, 9:12]).max() < 1e-9 and b["traj"][0, 3] == -30.0
        # node 1: the first photon; every forward node starts where its parent crossed
        f1 = nodes[1]
        assert f1["species"] == "photon" and f1["weight"] == 1.0 and f1["parent_weight"] == -1.0
        assert np.abs(f1["traj"][0, :3] - R[0, 9:12]).max() < 1e-9
        for nd in nodes[1:]:
            assert np.all(np.diff(nd["traj"][:, 3]) > 0) and len(nd["traj"]) <= nt + 2
            if len(nd["cross"]):   # stopped at its first crossing: the trajectory ends on it
                assert len(nd["cross"]) == 1 and np.abs(nd["traj"][-1, :3] - nd["cross"][0, :3]).max() < 1e-9
        # final nodes (no crossing) that left the star are the rows of the event: same end point
        finals = [nd for nd in nodes[1:] if len(nd["cross"]) == 0 and np.linalg.norm(nd["traj"][-1, :3]) > 11.0]
        assert len(finals) == len(R)
        for nd, r in zip(finals, R):
            x = nd["traj"][-1, :3]
            assert abs(np.linalg.norm(x) - r[6]) <= 1e-9 * r[6] and abs(math.acos(x[2] / np.linalg.norm(x)) - r[4]) < 1e-9
            assert (nd["species"] == "photon") == (r[1] == 1) and abs(nd["weight"] * R[0, 25] - r[8]) <= 1e-9 * r[8]
    assert n_nodes >= 12


def test_baseline_config0_flat_run_full_size_matches_oracle(lib, oracle):
    """BASELINE configs[0] at full size: flat-space run, m_a=1e-5 eV, B0=1e14 G, P=1 s (rotW=2π), θm=0.2, M=1, R=10 km,
    1e4 conversion-surface samples (Ntajs = 1e4 ⇒ 9999 events).  Per-event agreement of the tree shapes, and the
    sky-binned photon flux (16 cosθf bins) within Monte-Carlo error."""
    p = ArtParams.default(theta_m=0.2, flat=1, omega_pul=2 * math.pi)
    n = 9999
    h = lib.Handle(p, 0)
    a = h.run(1769, 1, n)
    h.close()
    b = oracle.run(p, 1769, 1, n)
    ca = np.bincount(a["rows"][:, 0].astype(int), minlength=n + 1)
    cb = np.bincount(b["rows"][:, 0].astype(int), minlength=n + 1)
    assert (ca == cb)[1:].mean() > 0.99
    assert abs(a["f_inx"] - b["f_inx"]) <= 0.005 * b["f_inx"] and abs(a["tot_count"] - b["tot_count"]) <= 0.01 * b["tot_count"]
    fa, fb = _binned(a["rows"]), _binned(b["rows"])
    ph = b["rows"][:, 1] == 1
    ic = np.clip(np.floor((np.cos(b["rows"][ph, 2]) + 1) * 0.5 * 16).astype(int), 0, 15)
    var = np.zeros(16)
    np.add.at(var, ic, (b["rows"][ph, 8] * b["rows"][ph, 7]) ** 2)
    assert abs(fa.sum() - fb.sum()) <= 0.01 * fb.sum()
    assert np.all(np.abs(fa - fb) <= 3.0 * np.sqrt(var) + 1e-3 * fb.sum())


def test_refilled_pool_runs_equal_the_single_block_run(lib, monkeypatch):
    """The pipeline keeps a pool of event slots full: slots handed back by finished events are refilled with freshly
    sampled ones while the loop runs (BASELINE configs 3-5 run 1e7+ events per GPU through a 2^20-slot pool).  Rows,
    counters and the event table must not depend on the pool size or on when blocks were injected."""
    p = ArtParams.default(theta_m=0.2, save_mode=2)
    n = 5000
    h = lib.Handle(p, 0)
    a = h.run(11, 1, n)
    ea = h.fetch_events(n)
    h.close()
    for env in (dict(ART_POOL_SLOTS="256"), dict(ART_POOL_SLOTS="1000", ART_REFILL_MIN="10", ART_SAMPLE_BLOCK="512"),
                dict(ART_POOL_SLOTS="4096", ART_SIDE_STREAM="0", ART_SAMPLE_BLOCK="1000"), dict(ART_POOL_SLOTS="64", ART_SAMPLE_BLOCK="100")):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        h = lib.Handle(p, 0)
        b = h.run(11, 1, n)
        eb = h.fetch_events(n)
        h.close()
        for k in env:
            monkeypatch.delenv(k)
        assert np.array_equal(a["rows"], b["rows"]) and np.array_equal(ea, eb), env
        assert a["f_inx"] == b["f_inx"] and a["tot_count"] == b["tot_count"]
        for k in ("n_rays", "n_attempts", "n_final_photon", "n_final_axion", "n_acc", "n_rej", "n_rhs"):
            assert a["stats"][k] == b["stats"][k], k
        if "ART_POOL_SLOTS" in env and int(env["ART_POOL_SLOTS"]) < n:
            assert b["stats"]["n_launches"] > a["stats"]["n_launches"]  # it was refilled


def test_carried_events_draw_the_same_samples_as_art_sample(lib, monkeypatch):
    """The sampler does not wait for its unluckiest events: once a launch's queue is empty, an event whose attempt fails is
    carried to the next launch and samples are appended in completion order (k_sample_events).  Attempt j of event e draws
    from the Philox stream (seed, id(e), j) wherever it runs: the conversion point a pipeline row carries (columns 9..11)
    and the failure count f_inx sums must be those of art_sample, bit for bit, also with blocks so small that almost every
    event is carried at least once."""
    p = ArtParams.default(theta_m=0.2)
    n = 20000
    for env in (dict(), dict(ART_SAMPLE_BLOCK="300", ART_POOL_SLOTS="2048"), dict(ART_SAMPLE_BLOCK="5000", ART_POOL_SLOTS="4096")):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        h = lib.Handle(p, 0)
        a = h.run(23, 1, n)
        smp = h.sample(23, 1, n)
        h.close()
        for k in env:
            monkeypatch.delenv(k)
        ev = a["rows"][:, 0].astype(np.int64)
        assert np.array_equal(a["rows"][:, 9:12], smp["x"][:, ev - 1].T), env
        # (f_inx = failed attempts + photons that reached the observer, MainRunner.jl:463-494, 725)
        assert a["f_inx"] == int(smp["n_failed"].sum()) + a["stats"]["n_final_photon"], env
        assert a["stats"]["n_attempts"] == n + int(smp["n_failed"].sum())


def test_batches_as_one_pipelined_job_equal_separate_calls(lib):
    """art_run_device_batches: contiguous and strided batches (the multi-GPU bench gives rank r the event ids
    1 + (step·world + r)·E + [0, E)) against one art_run per batch."""
    p = ArtParams.default(theta_m=0.2)
    h = lib.Handle(p, 0)
    E, nb = 700, 4
    for stride in (E, 3 * E):
        out = h.run_device_batches(21, 1 + E, E, nb, stride)
        rows = h.fetch_rows(out["n_rows"])
        parts = [h.run(21, 1 + E + b * stride, E) for b in range(nb)]
        ref = np.vstack([q["rows"] for q in parts])
        assert np.array_equal(rows, ref), stride
        assert out["f_inx"] == sum(q["f_inx"] for q in parts) and out["tot_count"] == sum(q["tot_count"] for q in parts)
        assert out["stats"]["n_events"] == E * nb
    h.close()


def test_row_buffers_grow_instead_of_failing(lib, oracle, monkeypatch):
    """The reference appends rows without bound (MainRunner.jl:729).  Device side: a row buffer that turns out too small is
    re-allocated at the exact size and the call re-run (identical, per-event random streams); host side: ART_ERR_CAPACITY
    reports the number of rows and art_fetch_rows is called again with a larger buffer.  Parameters of the reference's own
    convergence sweep (jonas_test_analyses/runner_convergence.sh: MassA 2e-5, Axg 5e-10, MCNodes 50, maxNodes 100,
    numCutoff 1000, probCutoff 1e-10, ThetaM 0); with them an event still ends in about one row, so the device buffer is
    made to overflow by starting it at a fifth of an event per row (ART_ROWS_INIT_FRAC, a test hook)."""
    p = ArtParams.default(theta_m=0.0, Mass_a=2e-5, Ax_g=5e-10, MC_nodes=50, max_nodes=100, num_cutoff=1000, prob_cutoff=1e-10,
                          save_mode=1)
    n = 400
    h = lib.Handle(p, 0)
    a = h.run(1769, 1, n)
    small = np.empty((10, h.ncols))
    c = h.run(1769, 1, n, rows_out=small)   # host buffer too small: re-fetched into a new one
    h.close()
    assert np.array_equal(a["rows"], c["rows"]) and a["rows"].shape[0] > 10
    monkeypatch.setenv("ART_ROWS_INIT_FRAC", "0.2")
    h = lib.Handle(p, 0)
    b = h.run(1769, 1, n)
    h.close()
    monkeypatch.delenv("ART_ROWS_INIT_FRAC")
    assert b["rows"].shape[0] > 0.2 * n and np.array_equal(a["rows"], b["rows"])
    assert a["f_inx"] == b["f_inx"] and a["tot_count"] == b["tot_count"]
    assert b["stats"]["n_launches"] > 0
    o = oracle.run(p, 1769, 1, n, rows_cap=200 * n)
    assert abs(a["rows"].shape[0] - o["rows"].shape[0]) <= 0.03 * o["rows"].shape[0] + 2
    assert abs(a["tot_count"] - o["tot_count"]) <= 0.03 * o["tot_count"] + 2


def _switch_cases():
    from tests.test_hostcheck_device_math import _OTHERS, _SWITCHES
    return _SWITCHES + _OTHERS


@pytest.mark.parametrize("kw", _switch_cases(), ids=lambda kw: ",".join(f"{k}={v}" for k, v in kw.items() if k != "theta_m" or len(kw) == 1))
def test_switch_matrix_gpu_pipeline_matches_oracle(lib, oracle, kw):
    """flat × isotropic × melrose × thick_surface and a spread of the other knobs through art_run, event by event."""
    p = ArtParams.default(**kw)
    n = 300
    h = lib.Handle(p, 0)
    a = h.run(1769, 1, n)
    h.close()
    b = oracle.run(p, 1769, 1, n)
    from tests.test_hostcheck_device_math import _same_shape_rows
    A, B, good = _same_shape_rows(a["rows"], b["rows"], n)  # events whose trees end in the same sequence of species
    assert good.mean() >= 0.97
    assert abs(a["f_inx"] - b["f_inx"]) <= 0.02 * b["f_inx"] + 2
    assert abs(a["tot_count"] - b["tot_count"]) <= 0.03 * b["tot_count"] + 2
    assert np.array_equal(A[:, :2], B[:, :2])
    assert rel(A[:, 7], B[:, 7]).max() < 1e-9 and np.abs(A[:, 9:12] - B[:, 9:12]).max() < 1e-9
    ph = B[:, 1] == 1
    assert np.median(np.abs(A[ph, 2] - B[ph, 2])) < 3e-6   # (Vern6 at the reference tolerances: see test_fused_pipeline_matches_oracle)


@pytest.mark.parametrize("thick", [False, True], ids=["thin", "thick"])
def test_sampler_against_independent_root_solve(lib, thick):
    """art_sample_from_uniforms, both surface conditions, against the numpy/scipy fixtures (tests/golden/sampler_*.json)."""
    from tests.test_oracle_golden import _sampler_check
    handles = {}

    def sampler(p, U):
        key = bytes(p)
        if key not in handles:
            handles[key] = lib.Handle(p, 0)
        return handles[key].sample_from_uniforms(U)

    _sampler_check(sampler, thick)
    for h in handles.values():
        h.close()
