"""Pin the CPU oracle against the golden vectors of tests/golden (sympy/mpmath at 50 digits,
scipy DOP853 at rtol 1e-12; generator: tests/golden/make_golden.py).  CPU only."""
import json
import math
import os

import numpy as np
import pytest

from adiabatic_raytracer_b200.params import ArtParams, c_km

G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    with open(os.path.join(G, name)) as f:
        return json.load(f)


def params_of(a):
    # args = r th ph kr kth kph t om thm w B0 rNS MNS ma
    return ArtParams.default(theta_m=a[8], omega_pul=a[9], B0=a[10], rNS=a[11], Mass_NS=a[12], Mass_a=a[13])


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def test_philox_known_answers(oracle):
    # Random123 kat_vectors for philox4x32-10
    assert list(oracle.philox([0, 0, 0, 0], [0, 0])) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert list(oracle.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2)) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert list(oracle.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0])) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    U = oracle.sampler_uniforms(1769, 3, 0)
    assert np.all((U >= 0) & (U < 1)) and len(set(U)) == 10


def test_pointwise_values(oracle):
    for case in load("pointwise.json"):
        a, o = case["args"], case["out"]
        p = params_of(a)
        wp = oracle.pointwise(p, 0, [[a[0], a[1], a[2], a[6]]])[0]
        assert rel(wp, o["omega_p"]) < 1e-13
        g = oracle.pointwise(p, 1, [[a[0], a[1]]], 4)[0]
        assert rel(g[0], o["gtt"]) < 1e-14 and rel(g[1], o["grr"]) < 1e-14
        gi = oracle.pointwise(p, 1, [[o["r_int_used"], a[1]]], 4)[0]
        assert rel(gi[0], o["gtt_int"]) < 1e-13 and rel(gi[1], o["grr_int"]) < 1e-13
        H = oracle.pointwise(p, 2, [[a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]]])[0]
        # H is a difference of O(E^2) terms: compare on that scale
        assert abs(H - o["H"]) < 1e-13 * a[7] ** 2
        if a[0] < 1e4:  # thick condition is only evaluated near the star
            tc = oracle.pointwise(p, 3, [[a[0], a[1], a[2], a[3], a[4], a[5], a[7], math.log(a[6]) if a[6] > 0 else -800.0]])[0]
            assert abs(tc - o["thick_cond"]) < 1e-12 * max(1.0, abs(o["thick_cond"]))
        of = oracle.pointwise(p, 4, [[a[0], a[1], a[2], a[3], a[4], a[5], a[6]]])[0]
        assert rel(of, o["omega_fn"]) < 1e-13


def test_rhs_against_symbolic_gradients(oracle):
    """func!/func_axion! (Dual-number gradients) vs sympy's closed-form partials."""
    n_checked = 0
    for case in load("pointwise.json"):
        a, o = case["args"], case["out"]
        if a[6] <= 0:
            continue
        p = params_of(a)
        erg = 1.0e-5
        t, om = a[6], a[7]
        u = np.array([[a[0], a[1], a[2], a[3] / erg, a[4] / erg, a[5] / erg, -om]])
        du = oracle.rhs(p, [1], u, [erg], [math.log(t)])[0]
        grr = o["grr"]
        f = c_km * t * grr / om
        exp = [o["dH_dkr"] * f, o["dH_dkth"] * f, o["dH_dkph"] * f, -o["dH_dr"] * f / erg, -o["dH_dth"] * f / erg,
               -o["dH_dph"] * f / erg, o["dH_dt"] * t * grr / om]
        scale_x = max(abs(e) for e in exp[:3])
        for i in range(7):
            assert abs(du[i] - exp[i]) <= 1e-11 * max(abs(exp[i]), 1e-3 * scale_x if i < 3 else 0.0) + 1e-300, (i, du[i], exp[i])
        # axion, exterior
        ua = u.copy()
        ua[0, 6] = -om
        dua = oracle.rhs(p, [0], ua, [om], [math.log(t)])[0]
        ka = np.array([a[3], a[4], a[5]]) / erg * om  # hamiltonian_axion sees u[4:6]*erg with erg=om here
        # golden partials were taken at k = a[3:6]; rescale: H_ax is quadratic in k
        s = om / erg
        fa = c_km * t * grr / om
        expa = [o["Hax_dkr"] * s * fa, o["Hax_dkth"] * s * fa, o["Hax_dkph"] * s * fa]
        for i in range(3):
            assert abs(dua[i] - expa[i]) <= 1e-11 * abs(expa[i]) + 1e-300
        n_checked += 1
    assert n_checked >= 6


def test_nonmelrose_hamiltonian_against_symbolic_gradients(oracle):
    """melrose=false (`Ctheta_B_sphere` form, RT.jl:546-548): H and func!'s seven components vs sympy."""
    n_checked = 0
    for case in load("pointwise_nonmelrose.json"):
        a, o = case["args"], case["out"]
        p = params_of(a).copy(melrose=0)
        H = oracle.pointwise(p, 2, [[a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]]])[0]
        assert abs(H - o["H"]) < 1e-13 * max(a[7] ** 2, abs(o["H"]))
        if a[6] <= 0:
            continue
        erg = 1.0e-5
        t, om = a[6], a[7]
        u = np.array([[a[0], a[1], a[2], a[3] / erg, a[4] / erg, a[5] / erg, -om]])
        du = oracle.rhs(p, [1], u, [erg], [math.log(t)])[0]
        f = c_km * t * o["grr"] / om
        exp = [o["dH_dkr"] * f, o["dH_dkth"] * f, o["dH_dkph"] * f, -o["dH_dr"] * f / erg, -o["dH_dth"] * f / erg,
               -o["dH_dph"] * f / erg, o["dH_dt"] * t * o["grr"] / om]
        scale_x = max(abs(e) for e in exp[:3])
        for i in range(7):
            assert abs(du[i] - exp[i]) <= 1e-10 * max(abs(exp[i]), 1e-3 * scale_x if i < 3 else 0.0) + 1e-300, (i, du[i], exp[i])
        n_checked += 1
    assert n_checked >= 6


@pytest.mark.parametrize("form", ["melrose", "nonmelrose"])
def test_flat_mode_rhs_against_symbolic_gradients(oracle, form):
    """flat = true: func! evaluates the Hamiltonian with Mass_NS = 0 (RT.jl:77-79) — BASELINE configs[0]'s ray equations."""
    from tests.hostcheck import loader as hc
    n_checked = 0
    for case in load("pointwise_flat.json"):
        a, o = case["args"], case[form]
        if a[6] <= 0:
            continue
        p = ArtParams.default(theta_m=a[8], omega_pul=a[9], B0=a[10], rNS=a[11], Mass_NS=1.0, Mass_a=a[13], flat=1,
                              melrose=1 if form == "melrose" else 0)
        erg = 1.0e-5
        t, om = a[6], a[7]
        u = np.array([[a[0], a[1], a[2], a[3] / erg, a[4] / erg, a[5] / erg, -om]])
        f = c_km * t / om  # g^rr = 1
        exp = [o["dH_dkr"] * f, o["dH_dkth"] * f, o["dH_dkph"] * f, -o["dH_dr"] * f / erg, -o["dH_dth"] * f / erg,
               -o["dH_dph"] * f / erg, o["dH_dt"] * t / om]
        scale_x = max(abs(e) for e in exp[:3])
        for which, du in (("oracle", oracle.rhs(p, [1], u, [erg], [math.log(t)])[0]),
                          ("device code on the host", hc.rhs(p, np.array([1], np.int32), u, [erg], [math.log(t)])[0])):
            for i in range(7):
                assert abs(du[i] - exp[i]) <= 1e-10 * max(abs(exp[i]), 1e-3 * scale_x if i < 3 else 0.0) + 1e-300, (which, i, du[i], exp[i])
        n_checked += 1
    assert n_checked >= 6


def test_invariant_u7_tracks_u6(oracle):
    """∂H/∂t = -Ω ∂H/∂φ (the field depends on t only through ψ = φ - Ωt, RT.jl:1142), hence
    du7 = (Ω·erg/c_km)·du6 — the identity the CUDA kernel exploits."""
    for case in load("pointwise.json"):
        a = case["args"]
        if a[6] <= 0:
            continue
        p = params_of(a)
        erg = 1.0e-5
        u = np.array([[a[0], a[1], a[2], a[3] / erg, a[4] / erg, a[5] / erg, -a[7]]])
        du = oracle.rhs(p, [1], u, [erg], [math.log(a[6])])[0]
        if a[0] <= 1.01 * a[11]:
            continue
        assert abs(du[6] - p.omega_pul * erg / c_km * du[5]) <= 1e-10 * abs(du[6]) + 1e-30


def test_weights_and_conversion_probability(oracle):
    for c in load("conv_prob.json"):
        par = c["par"]
        p = ArtParams.default(theta_m=par["thm"], omega_pul=par["w"], B0=abs(par["B0"]), rNS=par["rNS"],
                              Mass_NS=par["MNS"], Mass_a=par["ma"], Ax_g=par["Ax_g"], rho_DM=par["rho_DM"],
                              n_maxSample=par["n_max"])
        b_sign = 1 if par["B0"] > 0 else -1
        x = np.array(c["x"]).reshape(3, 1)
        vl = np.array(c["v_loc"]).reshape(3, 1)
        vi = np.array(c["v_inf"]).reshape(3, 1)
        W = oracle.weights(p, x, vl, vi)
        rc, maxR = oracle.max_radius(p)
        assert rc == 0
        for i in range(3):
            assert rel(W["k_init"][i, 0], c["k_init"][i]) < 1e-12
        assert rel(W["erg_inf"][0], c["erg_inf"]) < 1e-15
        assert rel(W["vel_eng"][0], c["vel_eng"]) < 1e-14
        assert rel(W["cos_w"][0], c["cos_w"]) < 1e-10
        assert rel(W["jac_gr"][0], c["jac_gr"]) < 1e-12
        assert rel(W["sln_prob"][0], c["sln_prob_over_maxR2"] * maxR**2) < 1e-10
        P = oracle.conv_prob(p, x, np.array(c["k_init"]).reshape(3, 1), [c["erg_inf"]], b_sign)[0]
        assert rel(P, c["P_nonAD"]) < 1e-9, (P, c["P_nonAD"])
        # isotropic switch (MR.jl:82-85, RT.jl:1420-1422)
        Pi = oracle.conv_prob(p.copy(isotropic=1), x, np.array(c["k_init"]).reshape(3, 1), [c["erg_inf"]], b_sign)[0]
        assert rel(Pi, c["P_nonAD_iso"]) < 1e-9, (Pi, c["P_nonAD_iso"])
        Wi = oracle.weights(p.copy(isotropic=1), x, vl, vi)  # cos_w from omega_function's iso branch, RT.jl:578-579
        assert rel(Wi["cos_w"][0], c["cos_w_iso"]) < 1e-10


def test_max_radius_matches_survey(oracle):
    p = ArtParams.default(theta_m=0.2)
    rc, maxR = oracle.max_radius(p)
    assert rc == 0 and abs(maxR - 25.17) < 0.01  # SURVEY.md §6 sizing
    rc, _ = oracle.max_radius(ArtParams.default(theta_m=0.2, Mass_a=1e-4))
    assert rc == -3  # "Too small Max R", MainRunner.jl:389-396


@pytest.mark.parametrize("tight", [True, False])
@pytest.mark.parametrize("fixture", ["rays.json", "rays_nonmelrose.json", "rays_flat.json"])
def test_rays_against_dop853(oracle, tight, fixture):
    """Oracle DP5 vs scipy DOP853(rtol 1e-12) on the same ICs.  tight: oracle at reltol 1e-12 /
    abstol 1e-14 (agreement = accuracy of the restated RHS + integrator);  not tight: production
    tolerances reltol 1e-7 / abstol 1e-6 (agreement = the integrator's own truncation error,
    the level at which ANY two schemes — incl. the reference's Vern6 — differ)."""
    worst = 0.0
    for ray in load(fixture):
        par = ray["par"]
        kw = dict(theta_m=par["thm"], omega_pul=par["w"], B0=par["B0"], rNS=par["rNS"], Mass_NS=par["MNS"], Mass_a=par["ma"],
                  melrose=0 if "nonmelrose" in fixture else 1,  # second fixture: photons of the melrose = false Hamiltonian
                  flat=1 if "flat" in fixture else 0)           # third: propagate's mixed "flat" mode (BASELINE configs[0])
        if tight:
            kw.update(reltol=1e-12, ode_err=1e-14)
        p = ArtParams.default(**kw)
        x0 = np.array(ray["x0"]).reshape(3, 1)
        kh = np.array(ray["khat"]).reshape(3, 1)
        R = oracle.propagate(p, [ray["species"]], x0, kh, [ray["erg"]], [0.0], [ray["dOmega"]], 1, 100000, 16)
        uf = np.array(ray["uf"])
        # compare final Cartesian position direction and |x|
        xf = R["x_f"][:, 0]
        xg = uf[0] * np.array([math.sin(uf[1]) * math.cos(uf[2]), math.sin(uf[1]) * math.sin(uf[2]), math.cos(uf[1])])
        err = np.linalg.norm(xf - xg) / np.linalg.norm(xg)
        worst = max(worst, err)
        assert abs(R["u7_f"][0] - uf[6]) <= (1e-9 if tight else 1e-5) * abs(uf[6])
        # crossings found by scipy's event locator that survive affect!'s filters
        gold = [c for c in ray["crossings"] if c["u"][0] >= 1.01 * par["rNS"]]
        if tight and R["status"][0] == 0:
            got = R["cross_rec"][0, : R["n_cross"][0]]
            # the oracle ignores the start-point crossing; scipy may or may not report it
            gold = [c for c in gold if c["lnt"] > -25.0]
            assert len(got) == len(gold), (len(got), len(gold))
            for gc, c in zip(got, gold):
                assert abs(math.log(gc[6]) - c["lnt"]) < 1e-7
    assert worst < (1e-8 if tight else 1e-3), worst  # measured: 2e-4 (radial), 7e-5 (angles)


def _sampler_check(sampler, thick):
    """Every crossing of every fixture line, selected through the tenth uniform (randInx = 1 + floor(u·n_max), RT.jl:1623)."""
    cases = load("sampler_thick.json" if thick else "sampler_thin.json")
    par = cases[0]["par"]
    p = ArtParams.default(theta_m=par["thm"], omega_pul=par["w"], B0=par["B0"], rNS=par["rNS"], Mass_NS=par["MNS"],
                          Mass_a=par["ma"], n_maxSample=par["n_max"], thick_surface=1 if thick else 0)
    n_checked = 0
    for c in cases:
        nc = len(c["crossings"])
        for j in range(par["n_max"]):
            U = np.array(c["U"])
            U[9] = (j + 0.5) / par["n_max"]
            o = sampler(p, U.reshape(1, 10))
            assert o["n_cross"][0] == nc
            assert o["accepted"][0] == (1 if j < nc else 0)     # RT.jl:1628-1632
            if j < nc:
                assert np.abs(o["x"][:, 0] - np.array(c["crossings"][j])).max() < 1e-11
                n_checked += 1
    assert n_checked >= 28
    return p, cases


@pytest.mark.parametrize("thick", [False, True], ids=["thin", "thick"])
def test_sampler_against_independent_root_solve(oracle, thick):
    """find_samples_new, both surface conditions: line construction, crossing roots, affect!'s filter and the index draw
    against a numpy/scipy restatement that shares no code with oracle/ (tests/golden/make_golden.py: sampler_*_cases)."""
    p, cases = _sampler_check(oracle.sample_from_uniforms, thick)
    assert rel(oracle.max_radius(p)[1], cases[0]["maxR"]) < 1e-14


@pytest.mark.parametrize("thick", [False, True], ids=["thin", "thick"])
def test_sampler_device_code_against_independent_root_solve(thick):
    from tests.hostcheck import loader as hc
    _sampler_check(hc.sample_from_uniforms, thick)
