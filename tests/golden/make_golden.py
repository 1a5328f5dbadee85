#!/usr/bin/env python
"""Generate the golden vectors that pin the CPU oracle (oracle/) — run in THIS container.

The reference (Julia + OrdinaryDiffEq + ForwardDiff) cannot run here and ships no tests or
golden vectors, so the pins are produced from the reference's FORMULAS by an
independent route:

  * the Hamiltonians / plasma frequency / metric / thick-surface condition / ω(x,k) /
    conversion probability are typed into sympy exactly as written in
    /root/reference/src/RayTracer.jl (line numbers in the comments) and
    /root/reference/src/MainRunner.jl:67-124, differentiated SYMBOLICALLY and evaluated with
    mpmath at 50 digits  → pointwise.json
  * the ray equations func!/func_axion! (RT.jl:71-123) built from those symbolic gradients
    are integrated by scipy's DOP853 at rtol 1e-12 (a different scheme from the oracle's
    DP5) with scipy's own event location → rays.json

Nothing here imports oracle/ or the product.  Usage:  python tests/golden/make_golden.py
"""
import json
import math
import os

import mpmath as mp
import numpy as np
import sympy as sp
from scipy.integrate import solve_ivp

HERE = os.path.dirname(os.path.abspath(__file__))
mp.mp.dps = 50

# Constants.jl:3-5
C_KM, HBAR, GNEW = 2.99792e5, 6.582119e-16, 1.32712e11

r, th, ph, kr, kth, kph, t, om = sp.symbols("r theta phi k_r k_theta k_phi t omega", real=True)
thm, w, B0, rNS, MNS, ma = sp.symbols("theta_m w B0 rNS M_NS m_a", real=True)
c_km, hbar, GN = sp.symbols("c_km hbar G_N", positive=True)


def metric_ext():
    """g_schwartz exterior branch, RT.jl:490-493."""
    rs = 2 * GN * MNS / c_km**2
    return -1 / (1 - rs / r), (1 - rs / r), 1 / r**2, 1 / (r**2 * sp.sin(th) ** 2)


def metric_int(R=sp.Integer(10)):
    """g_schwartz interior branch, RT.jl:463,496-497 (rs already rescaled)."""
    rs = 2 * GN * MNS / c_km**2 * (r / R) ** 3
    gtt = -4 / (3 * sp.sqrt(1 - rs / R) - sp.sqrt(1 - r**2 * rs / R**3)) ** 2
    grr = 1 - r**2 * rs / R**3
    return gtt, grr, 1 / r**2, 1 / (r**2 * sp.sin(th) ** 2)


def dipole(time):
    """RT.jl:1142-1147."""
    psi = ph - w * time
    Bn = B0 * (rNS / r) ** 3 / 2
    Br = 2 * Bn * (sp.cos(thm) * sp.cos(th) + sp.sin(thm) * sp.sin(th) * sp.cos(psi))
    Bth = Bn * (sp.cos(thm) * sp.sin(th) - sp.sin(thm) * sp.cos(th) * sp.cos(psi))
    Bph = Bn * sp.sin(thm) * sp.sin(psi)
    return Br, Bth, Bph


def omega_p(time):
    """GJ_Model_ωp_vecSPH without boundary layer / zeroIn, RT.jl:1151-1154."""
    Br, Bth, _ = dipole(time)
    Bz = Br * sp.cos(th) - Bth * sp.sin(th)
    nelec = sp.Abs((2 * w * Bz) / sp.sqrt(4 * sp.pi / 137) * sp.Rational(195, 10000) * hbar)
    return sp.sqrt(4 * sp.pi * nelec / 137 / 500000)


def b_cov(time, g):
    """GJ_Model_Sphereical return_comp=-1, RT.jl:1298."""
    Br, Bth, Bph = dipole(time)
    return Br / sp.sqrt(g[1]), Bth / sp.sqrt(g[2]), Bph / sp.sqrt(g[3])


def k_par(time, g, k):
    """K_par, RT.jl:1044-1058."""
    B = b_cov(time, g)
    Bmag = sp.sqrt(g[1] * B[0] ** 2 + g[2] * B[1] ** 2 + g[3] * B[2] ** 2)
    return (g[1] * k[0] * B[0] + g[2] * k[1] * B[1] + g[3] * k[2] * B[2]) / Bmag


def hamiltonian(g):
    """hamiltonian melrose branch, RT.jl:537,550-551."""
    k = (kr, kth, kph)
    ksqr = g[0] * om**2 + g[1] * kr**2 + g[2] * kth**2 + g[3] * kph**2
    kp = k_par(t, g, k)
    wp = omega_p(t)
    return sp.Rational(1, 2) * (ksqr + wp**2 * (om**2 / g[1] - kp**2) / (om**2 / g[1]))


def ctheta_B_sphere(time, g, k):
    """Ctheta_B_sphere, RT.jl:957-971."""
    Br, Bth, Bph = dipole(time)
    Br, Bth, Bph = Br * sp.sqrt(g[1]), Bth * sp.sqrt(g[2]), Bph * sp.sqrt(g[3])
    Bnorm = sp.sqrt(Br**2 / g[1] + Bth**2 / g[2] + Bph**2 / g[3])
    knorm = sp.sqrt(g[1] * k[0] ** 2 + g[2] * k[1] ** 2 + g[3] * k[2] ** 2)
    return (k[0] * Br + k[1] * Bth + k[2] * Bph) / knorm / Bnorm


def hamiltonian_nonmelrose(g):
    """hamiltonian with melrose=false ("original form"), RT.jl:537,546-548."""
    k = (kr, kth, kph)
    ksqr = g[0] * om**2 + g[1] * kr**2 + g[2] * kth**2 + g[3] * kph**2
    ct = ctheta_B_sphere(t, g, k)
    wp = omega_p(t)
    return sp.Rational(1, 2) * (ksqr - wp**2 * (1 - ct**2) / (wp**2 * ct**2 - om**2 / g[1]) * om**2 / g[1])


def hamiltonian_axion(g):
    """RT.jl:632-640."""
    return sp.Rational(1, 2) * (g[0] * om**2 + g[1] * kr**2 + g[2] * kth**2 + g[3] * kph**2)


def omega_function(g):
    """RT.jl:558-589, anisotropic branch incl. the `/sqrt(2)`."""
    k = (kr, kth, kph)
    ksqr = g[1] * kr**2 + g[2] * kth**2 + g[3] * kph**2
    kp = k_par(t, g, k)
    wp = omega_p(t)
    Ham = (ksqr + wp**2 + sp.sqrt(ksqr**2 + 2 * ksqr * wp**2 - 4 * kp**2 * wp**2 + wp**4)) / sp.sqrt(2)
    return sp.sqrt(Ham)


def thick_condition(g):
    """propagate's condition, RT.jl:262-296: (w0 = k direction, erg_inf = om)."""
    wv = (kr, kth, kph)
    NrmSq = (-om**2 * g[0] - ma**2) / (wv[0] ** 2 * g[1] + wv[1] ** 2 * g[2] + wv[2] ** 2 * g[3])
    wn = tuple(x * sp.sqrt(NrmSq) for x in wv)
    kp = k_par(t, g, wn)
    wp = omega_p(t)
    ksqr = g[0] * om**2 + g[1] * wn[0] ** 2 + g[2] * wn[1] ** 2 + g[3] * wn[2] ** 2
    return sp.Rational(1, 2) * (ksqr + wp**2 * (om**2 / g[1] - kp**2) / (om**2 / g[1])) / om**2


CONSTS = {c_km: C_KM, hbar: HBAR, GN: GNEW}
ARGS = (r, th, ph, kr, kth, kph, t, om, thm, w, B0, rNS, MNS, ma)


def lam(expr, module):
    e = expr.subs({c_km: sp.Float(C_KM, 60), hbar: sp.Float(HBAR, 60), GN: sp.Float(GNEW, 60)})
    return sp.lambdify(ARGS, e, modules=module)


def fnum(x):
    return float(x)


def mpf_args(a):
    return [mp.mpf(float(x)) for x in a]


# ------------------------------------------------------------------------------------------
def pointwise():
    g = metric_ext()
    gi = metric_int()
    H = hamiltonian(g)
    Hax = hamiltonian_axion(g)
    Hax_i = hamiltonian_axion(gi)
    OM = omega_function(g)
    TC = thick_condition(g)
    WP = omega_p(t)
    exprs = {
        "omega_p": WP,
        "H": H,
        "dH_dr": sp.diff(H, r), "dH_dth": sp.diff(H, th), "dH_dph": sp.diff(H, ph),
        "dH_dkr": sp.diff(H, kr), "dH_dkth": sp.diff(H, kth), "dH_dkph": sp.diff(H, kph),
        "dH_dt": sp.diff(H, t),
        "Hax_dr": sp.diff(Hax, r), "Hax_dth": sp.diff(Hax, th), "Hax_dkr": sp.diff(Hax, kr),
        "Hax_dkth": sp.diff(Hax, kth), "Hax_dkph": sp.diff(Hax, kph),
        "Haxi_dr": sp.diff(Hax_i, r), "Haxi_dth": sp.diff(Hax_i, th), "Haxi_dkr": sp.diff(Hax_i, kr),
        "omega_fn": OM, "domega_dr": sp.diff(OM, r), "domega_dth": sp.diff(OM, th), "domega_dph": sp.diff(OM, ph),
        "thick_cond": TC,
        "gtt": g[0], "grr": g[1], "gtt_int": gi[0], "grr_int": gi[1],
    }
    fns = {k: lam(v, "mpmath") for k, v in exprs.items()}
    rng = np.random.default_rng(20261018)
    pts = []
    base = dict(thm=0.2, w=1.0, B0=1e14, rNS=10.0, MNS=1.0, ma=1e-5)
    special = [  # (r, th, ph) incl. near-pole, near GJ null surface region, large r, r just above 10
        (25.0, 1.0, 0.3), (11.0, 0.05, 2.0), (15.0, math.pi - 0.03, -1.0), (40.0, 0.9553166, 0.1),
        (200.0, 2.2, 3.0), (10.2, 1.3, -2.5), (12.0, 0.7, 6.0), (3.0e5, 1.1, 0.4),
    ]
    for i in range(16):
        if i < len(special):
            rr, tt, pp = special[i]
        else:
            rr = float(10.5 + 40 * rng.random())
            tt = float(0.1 + (math.pi - 0.2) * rng.random())
            pp = float(-math.pi + 2 * math.pi * rng.random())
        kk = rng.normal(size=3)
        E = 1e-5 * (1 + 2.7e-7)
        # covariant components of roughly axion-momentum size
        kv = [float(kk[0] * 5e-6), float(kk[1] * 5e-6 * rr), float(kk[2] * 5e-6 * rr * math.sin(tt))]
        time = float(10 ** rng.uniform(-9, -3)) if i % 2 else 0.0
        par = dict(base)
        if i % 5 == 4:
            par.update(thm=0.7, w=2 * math.pi, B0=-1e14)
        a = [rr, tt, pp, kv[0], kv[1], kv[2], time, E, par["thm"], par["w"], par["B0"], par["rNS"], par["MNS"], par["ma"]]
        out = {}
        for k, f in fns.items():
            aa = mpf_args(a)
            if k.startswith("Haxi") or k.endswith("_int"):
                aa[0] = mp.mpf(float(2.0 + 7.5 * (i / 16.0)))  # interior radius
                out["r_int_used"] = float(aa[0])
            out[k] = fnum(f(*aa))
        pts.append(dict(args=a, out=out))
    return pts


def pointwise_nonmelrose():
    """H and its 7 partials for melrose=false at the argument points of pointwise.json."""
    g = metric_ext()
    H = hamiltonian_nonmelrose(g)
    exprs = {"H": H, "dH_dr": sp.diff(H, r), "dH_dth": sp.diff(H, th), "dH_dph": sp.diff(H, ph),
             "dH_dkr": sp.diff(H, kr), "dH_dkth": sp.diff(H, kth), "dH_dkph": sp.diff(H, kph),
             "dH_dt": sp.diff(H, t), "grr": g[1]}
    fns = {k: lam(v, "mpmath") for k, v in exprs.items()}
    pts = []
    for case in json.load(open(os.path.join(HERE, "pointwise.json"))):
        a = case["args"]
        pts.append(dict(args=a, out={k: fnum(f(*mpf_args(a))) for k, f in fns.items()}))
    return pts


def pointwise_flat():
    """func!'s Hamiltonian in "flat" mode (RT.jl:77-79: Mass_NS = 0 inside the right-hand side): H and its 7 partials,
    Melrose and non-Melrose forms, at the argument points of pointwise.json with M_NS = 0."""
    g = metric_ext()
    out = []
    forms = {"melrose": hamiltonian(g), "nonmelrose": hamiltonian_nonmelrose(g)}
    fns = {name: {k: lam(v, "mpmath") for k, v in
                  {"H": H, "dH_dr": sp.diff(H, r), "dH_dth": sp.diff(H, th), "dH_dph": sp.diff(H, ph), "dH_dkr": sp.diff(H, kr),
                   "dH_dkth": sp.diff(H, kth), "dH_dkph": sp.diff(H, kph), "dH_dt": sp.diff(H, t)}.items()}
           for name, H in forms.items()}
    for case in json.load(open(os.path.join(HERE, "pointwise.json"))):
        a = list(case["args"])
        a[12] = 0.0  # M_NS
        out.append(dict(args=a, **{name: {k: fnum(f(*mpf_args(a))) for k, f in fd.items()} for name, fd in fns.items()}))
    return out


# ------------------------------------------------------------------------------------------
def conversion_prob_cases():
    """get_Prob_nonAD + conversion_prob (MR.jl:67-124, RT.jl:1405-1473) and the per-sample
    scalars of MR.jl:498-557, at 50 digits, non-flat, anisotropic, bndry_lyr off."""
    g = metric_ext()
    WP = omega_p(t)
    Br, Bth, Bph = dipole(t)
    Bcov = b_cov(t, g)
    comp0 = sp.sqrt(Br**2 + Bth**2 + Bph**2) * sp.Rational(195, 10000)  # RT.jl:1300
    comps = [Bcov[i] * g[i + 1] * sp.Rational(195, 10000) for i in range(3)]  # RT.jl:1302-1306
    X = (r, th, ph)
    d_wp = [sp.diff(WP, v) for v in X]
    d_B = [sp.diff(comp0, v) for v in X]
    d_c = [[sp.diff(comps[i], v) for v in X] for i in range(3)]
    OM = omega_function(g)
    d_om = [sp.diff(OM, v) for v in X]
    # omega_function's iso branch, RT.jl:578-579: ω = √(g^ij k_i k_j + ωp²)
    OM_iso = sp.sqrt(g[1] * kr**2 + g[2] * kth**2 + g[3] * kph**2 + WP**2)
    d_om_iso = [lam(sp.diff(OM_iso, v), "mpmath") for v in X]
    f = dict(
        g=[lam(x, "mpmath") for x in g], wp=lam(WP, "mpmath"), Bcov=[lam(x, "mpmath") for x in Bcov],
        d_wp=[lam(x, "mpmath") for x in d_wp], d_B=[lam(x, "mpmath") for x in d_B],
        d_c=[[lam(x, "mpmath") for x in row] for row in d_c], d_om=[lam(x, "mpmath") for x in d_om],
    )
    rng = np.random.default_rng(7)
    cases = []
    cK, hB, Gn = mp.mpf(C_KM), mp.mpf(HBAR), mp.mpf(GNEW)
    for i in range(10):
        par = dict(thm=0.2, w=1.0, B0=1e14 if i % 3 else -1e14, rNS=10.0, MNS=1.0, ma=1e-5, Ax_g=1e-12, rho_DM=0.45, n_max=6)
        # a point on (near) the conversion surface: solve ωp(r) = ma along a random direction
        tt = float(0.15 + (math.pi - 0.3) * rng.random())
        pp = float(-math.pi + 2 * math.pi * rng.random())
        a0 = [20.0, tt, pp, 0, 0, 0, 0.0, 1e-5, par["thm"], par["w"], par["B0"], par["rNS"], par["MNS"], par["ma"]]
        wp20 = float(f["wp"](*mpf_args(a0)))
        rr = 20.0 * (wp20 / par["ma"]) ** (2.0 / 3.0)
        if rr < 10.5:
            rr = 10.5 + rng.random()
        x = np.array([rr * math.sin(tt) * math.cos(pp), rr * math.sin(tt) * math.sin(pp), rr * math.cos(tt)])
        vdir = rng.normal(size=3)
        vdir /= np.linalg.norm(vdir)
        vinf = (220.0 + rng.random(3) * 1e-5) / math.sqrt(3.0) / C_KM
        vinf_mag = float(np.linalg.norm(vinf))
        vmag_loc = math.sqrt((vinf_mag * C_KM) ** 2 + 2 * GNEW * par["MNS"] / rr) / C_KM
        vloc = vdir * vmag_loc
        # ---- high-precision restatement (all mp) ------------------------------------------
        X3 = [mp.mpf(float(v)) for v in x]
        V3 = [mp.mpf(float(v)) for v in vloc]
        rmag = mp.sqrt(sum(v * v for v in X3))
        thS = mp.acos(X3[2] / rmag)
        phS = mp.atan2(X3[1], X3[0])
        vI = [mp.mpf(float(v)) for v in vinf]
        vI_mag = mp.sqrt(sum(v * v for v in vI))
        vel_eng = sum((v / cK) ** 2 for v in vI) / 2  # MR.jl:522
        gammaA = 1 / mp.sqrt(1 - vI_mag**2)
        maM = mp.mpf(par["ma"])
        erg_inf = maM * mp.sqrt(1 + (vI_mag * gammaA) ** 2)
        M = mp.mpf(par["MNS"])
        rs = 2 * M * Gn / cK**2
        AA = 1 - rs / rmag
        args = [rmag, thS, phS, 0, 0, 0, mp.mpf(0), erg_inf] + [mp.mpf(par[k]) for k in ("thm", "w", "B0", "rNS", "MNS", "ma")]
        gv = [fn(*args) for fn in f["g"]]

        def celer(kvec):
            dr = sum(a * b for a, b in zip(X3, kvec)) / rmag
            st = mp.sin(thS)
            v0 = dr
            v1 = (X3[2] * dr - rmag * kvec[2]) / (rmag * st)
            v2 = (-X3[1] * kvec[0] + X3[0] * kvec[1]) / (rmag * st)
            return [v0 / mp.sqrt(AA) / AA, v1 * rmag / AA, v2 * (rmag * st) / AA]

        wv = celer(V3)
        NrmSq = (-erg_inf**2 * gv[0] - maM**2) / (wv[0] ** 2 * gv[1] + wv[1] ** 2 * gv[2] + wv[2] ** 2 * gv[3])
        k_init = [mp.sqrt(NrmSq) * v for v in V3]  # RT.jl:683
        ks = celer(k_init)  # k_sphere
        sdot = lambda a, b: gv[1] * a[0] * b[0] + gv[2] * a[1] * b[1] + gv[3] * a[2] * b[2]
        kmag = mp.sqrt(sdot(ks, ks))
        argsk = list(args)
        argsk[3:6] = ks
        # cos_w, RT.jl:1370-1375
        go = [fn(*argsk) for fn in f["d_om"]]
        nrm = mp.sqrt(sdot(go, go))
        cos_w = abs(sdot([k / kmag for k in ks], [x_ / nrm for x_ in go]))
        go_i = [fn(*argsk) for fn in d_om_iso]
        cos_w_iso = abs(sdot([k / kmag for k in ks], [x_ / mp.sqrt(sdot(go_i, go_i)) for x_ in go_i]))
        # g_det, RT.jl:743-753
        dwp = [fn(*args) for fn in f["d_wp"]]
        dr_th, dr_p = dwp[1] / dwp[0], dwp[2] / dwp[0]
        st = mp.sin(thS)
        jac = mp.sqrt(st**2 * (gv[1] * rmag**2 + dr_th**2) + dr_p**2) / mp.sqrt(st**2 * (rmag**2 + dr_th**2) + dr_p**2)
        maxR = None  # filled by the test from the oracle's own maxR (host scalar) – stored separately
        dense_extra = 2 / mp.sqrt(mp.pi) * (1 / (mp.mpf(220) / cK)) * mp.sqrt(2 * M * Gn / cK**2 / rmag)
        redshift = mp.sqrt(1 - 2 * Gn * M / rmag / cK**2)
        # sln_prob / maxR^2 (MR.jl:555-556) so that the golden value does not depend on maxR
        sln_over_maxR2 = cos_w * redshift * dense_extra * (2 * mp.pi) * (mp.mpf(par["rho_DM"]) * mp.mpf(10) ** 9) / maxM(par) * jac * mp.mpf(10) ** 10 * cK * mp.mpf(10) ** 5 * par["n_max"]
        # ---- conversion probability --------------------------------------------------------
        Bs = [fn(*args) for fn in f["Bcov"]]
        c195 = mp.mpf(195) / 10000
        Bmag = mp.sqrt(sdot(Bs, Bs)) * c195
        ctB = sdot(Bs, ks) * c195 / (kmag * Bmag)
        stB = mp.sin(mp.acos(ctB))
        wE = erg_inf / mp.sqrt(1 - 2 * Gn * M / rmag / cK**2)
        wp = f["wp"](*args)
        GM = Gn * M / cK**2
        G_rrr = -GM / (rmag * (rmag - 2 * GM))
        G_rtt = -(rmag - 2 * GM)
        G_rpp = -(rmag - 2 * GM) * st**2
        G_trt = G_prp = G_ttr = G_ppr = 1 / rmag
        G_tpp = -st * mp.cos(thS)
        G_ptp = G_ppt = mp.cos(thS) / st
        dB = [fn(*args) for fn in f["d_B"]]
        term1 = [sum(ks[i] * f["d_c"][i][j](*args) for i in range(3)) for j in range(3)]
        t2r = ks[0] * (gv[1] * Bs[0] * c195) * G_rrr + ks[1] * G_trt * (Bs[1] * gv[2] * c195) + ks[2] * G_prp * (Bs[2] * gv[3] * c195)
        t2t = ks[0] * (gv[2] * Bs[1] * c195) * G_rtt + ks[2] * G_ptp * (Bs[2] * gv[3] * c195) + ks[1] * (gv[1] * Bs[0] * c195) * G_ttr
        t2p = (ks[0] * (gv[3] * Bs[2] * c195) * G_rpp + ks[1] * G_tpp * (Bs[2] * gv[3] * c195)
               + ks[2] * G_ppr * (Bs[0] * gv[1] * c195) + ks[2] * G_ppt * (Bs[1] * gv[2] * c195))
        t2 = [t2r, t2t, t2p]
        dct = [(term1[j] + t2[j]) / (kmag * Bmag) - ctB * dB[j] / Bmag for j in range(3)]
        preF = wp / abs(wE**5 + ctB**2 * wE * (wp**4 - 2 * wp**2 * wE**2))
        dE = [preF * (wE**4 * stB**2 * dwp[j] - wE**2 * ctB * wp * (wE**2 - wp**2) * dct[j]) for j in range(3)]
        vhat = sdot([k / kmag for k in ks], dE)
        vloc_ = mp.sqrt(wE**2 - maM**2) / wE
        pref = wE**4 * stB**2 / (ctB**2 * wp**2 * (wp**2 - 2 * wE**2) + wE**4)
        Prob = mp.pi / 2 * pref * (mp.mpf(par["Ax_g"]) * mp.mpf(10) ** -9 * Bmag) ** 2 / (abs(vhat) * vloc_ * cK * hB)
        # isotropic switch (MR.jl:82-85: cosθ_B := 0, sinθ_B := 1; RT.jl:1420-1422: ∇E = ∇ω with ω = √(kmag² + ωp²),
        # omega_function's iso branch RT.jl:578-579)
        om_iso = mp.sqrt(kmag**2 + wp**2)
        dE_iso = [wp * dwp[j] / om_iso for j in range(3)]
        vhat_iso = sdot([k / kmag for k in ks], dE_iso)
        Prob_iso = mp.pi / 2 * (mp.mpf(par["Ax_g"]) * mp.mpf(10) ** -9 * Bmag) ** 2 / (abs(vhat_iso) * vloc_ * cK * hB)
        cases.append(dict(P_nonAD_iso=float(Prob_iso), cos_w_iso=float(cos_w_iso),
                          par=par, x=[float(v) for v in x], v_loc=[float(v) for v in vloc], v_inf=[float(v) for v in vinf],
                          k_init=[float(v) for v in k_init], erg_inf=float(erg_inf), vel_eng=float(vel_eng),
                          cos_w=float(cos_w), jac_gr=float(jac), sln_prob_over_maxR2=float(sln_over_maxR2),
                          P_nonAD=float(Prob)))
    return cases


def maxM(par):
    return mp.mpf(par["ma"])


# ------------------------------------------------------------------------------------------
def rays(ham=None, want_ph=6, want_ax=4, seed=99, flat=False):
    """DOP853 (rtol 1e-12) integration of func!/func_axion! from symbolic gradients.
    ham: photon Hamiltonian builder (default: the Melrose form).
    flat: propagate's mixed "flat" mode — k_norm_Cart keeps the star's mass (RT.jl:181,185 are called before
    `if flat Mass_NS = 0`, RT.jl:187-189), the initial celerity, the right-hand side (RT.jl:77-79) and the crossing
    condition (closure over the zeroed Mass_NS) use M = 0."""
    g = metric_ext()
    gi = metric_int()
    H = (ham or hamiltonian)(g)
    Hax = hamiltonian_axion(g)
    Haxi = hamiltonian_axion(gi)
    TC = thick_condition(g)
    num = lambda e: lam(e, "numpy")
    dH = [num(sp.diff(H, v)) for v in (r, th, ph, kr, kth, kph, t)]
    dHa = [num(sp.diff(Hax, v)) for v in (r, th, ph, kr, kth, kph)]
    dHai = [num(sp.diff(Haxi, v)) for v in (r, th, ph, kr, kth, kph)]
    grr_e, grr_i = num(g[1]), num(gi[1])
    gtt_e = num(g[0])
    tc = num(TC)

    def make_rhs(par, erg, axion):
        P = (par["thm"], par["w"], par["B0"], par["rNS"], 0.0 if flat else par["MNS"], par["ma"])

        def f(lnt, u):
            tt = math.exp(lnt)
            if axion:
                a = (u[0], u[1], u[2], u[3] * erg, u[4] * erg, u[5] * erg, tt, erg) + P
                inside = u[0] <= 10.0
                d = dHai if inside else dHa
                grr = (grr_i if inside else grr_e)(*a)
                du = np.zeros(7)
                for i in range(3):
                    du[i] = d[3 + i](*a) * C_KM * tt * grr / erg
                    du[3 + i] = -d[i](*a) * C_KM * tt * grr / erg / erg
                return du
            omg = -u[6]
            a = (u[0], u[1], u[2], u[3] * erg, u[4] * erg, u[5] * erg, tt, omg) + P
            grr = grr_e(*a)
            du = np.zeros(7)
            for i in range(3):
                du[i] = dH[3 + i](*a) * C_KM * tt * grr / omg
                du[3 + i] = -dH[i](*a) * C_KM * tt * grr / omg / erg
            if u[0] <= par["rNS"] * 1.01:
                du[:] = 0.0
            du[6] = dH[6](*a) * tt * grr / omg
            return du

        def cond(lnt, u):
            a = (u[0], u[1], u[2], u[3], u[4], u[5], math.exp(lnt), u[6]) + P
            if u[0] <= par["rNS"]:
                return -0.5 * par["ma"] ** 2 / u[6] ** 2  # zeroIn: ωp = 0 inside (RT.jl:284, 1165)
            return float(tc(*a))

        return f, cond

    def initial_state(par, x0, khat, erg, dOmega, flat=False):
        """propagate's IC transform, RT.jl:179-216 (k_norm_Cart + celerity)."""
        x0 = np.asarray(x0, float)
        rr = np.linalg.norm(x0)
        tt_, pp_ = math.acos(x0[2] / rr), math.atan2(x0[1], x0[0])
        a = (rr, tt_, pp_, 0, 0, 0, 0.0, erg, par["thm"], par["w"], par["B0"], par["rNS"], par["MNS"], par["ma"])
        rs = 2 * par["MNS"] * GNEW / C_KM**2
        AA = 1 - rs / rr

        def cel(kv, AA=AA):
            dr = float(np.dot(x0, kv)) / rr
            st = math.sin(tt_)
            v1 = (x0[2] * dr - rr * kv[2]) / (rr * st)
            v2 = (-x0[1] * kv[0] + x0[0] * kv[1]) / (rr * st)
            return np.array([dr / math.sqrt(AA) / AA, v1 * rr / AA, v2 * rr * st / AA])

        gtt, grr = gtt_e(*a), grr_e(*a)
        wv = cel(khat)
        NrmSq = (-erg**2 * gtt - par["ma"] ** 2) / (wv[0] ** 2 * grr + wv[1] ** 2 / rr**2 + wv[2] ** 2 / (rr * math.sin(tt_)) ** 2)
        k0 = math.sqrt(NrmSq) * np.asarray(khat)
        w0 = (cel(k0, 1.0) if flat else cel(k0)) / erg
        return np.array([rr, tt_, pp_, w0[0], w0[1], w0[2], erg * dOmega]), k0

    out = []
    rng = np.random.default_rng(seed)
    par = dict(thm=0.2, w=1.0, B0=1e14, rNS=10.0, MNS=1.0, ma=1e-5)
    erg = 1e-5 * math.sqrt(1 + (220.0 / C_KM) ** 2)
    ln0, ln1 = -30.0, math.log(1.0 / par["w"])
    n_ph, n_ax = 0, 0
    tries = 0
    while (n_ph < want_ph or n_ax < want_ax) and tries < 200:
        tries += 1
        axion = n_ph >= want_ph
        # start on/near the conversion surface, direction roughly outward (photon) / any (axion)
        tt_ = float(0.2 + (math.pi - 0.4) * rng.random())
        pp_ = float(-math.pi + 2 * math.pi * rng.random())
        rr = float(14 + 12 * rng.random())
        x0 = rr * np.array([math.sin(tt_) * math.cos(pp_), math.sin(tt_) * math.sin(pp_), math.cos(tt_)])
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        if not axion and np.dot(d, x0) < 0:
            d = -d
        u0, k0 = initial_state(par, x0, d * 0.3, erg, -1.0, flat)
        f, cond = make_rhs(par, erg, axion)
        ev = lambda s, u: cond(s, u)
        sol = solve_ivp(f, (ln0, ln1), u0, method="DOP853", rtol=1e-12, atol=1e-16, events=ev, dense_output=False)
        if sol.status != 0:
            continue
        uf = sol.y[:, -1]
        if not axion and uf[0] < 50:
            continue
        # crossings (all sign changes); positions/times
        te = sol.t_events[0]
        ue = sol.y_events[0]
        cross = []
        for s_, u_ in zip(te, ue):
            cross.append(dict(lnt=float(s_), u=[float(v) for v in u_]))
        rec = dict(species=0 if axion else 1, par=par, erg=erg, x0=[float(v) for v in x0], khat=[float(v) for v in d * 0.3],
                   dOmega=-1.0, u0=[float(v) for v in u0], uf=[float(v) for v in uf], crossings=cross,
                   nfev=int(sol.nfev))
        out.append(rec)
        if axion:
            n_ax += 1
        else:
            n_ph += 1
    return out


def sampler_cases(thick, n_lines=14, seed=2024):
    """The same for the default thick_surface = true condition (RT.jl:1552-1581: celerity of the local velocity direction,
    axion mass shell, K_par with the covariant dipole field, the AA = 1 quirk inside the star, g_schwartz's interior
    branch below 10 km)."""
    return sampler_thin_cases(n_lines, seed, thick)


def sampler_thin_cases(n_lines=14, seed=2024, thick_mode=False):
    """find_samples_new with thick_surface = false (RT.jl:1480-1653), restated independently of oracle/: the line of an
    attempt from its uniforms (1486-1497, 1521-1529, 1542), Find_Conversion_Surface's radius (1250-1263), the thin
    condition log ωp − log m_a (1548-1551) solved for ALL its roots on τ ∈ (0, 2.2 maxR) by a dense scan + brentq at
    1e-15, and affect!'s filter (1585-1597).  Lines whose roots are closer than 0.1 km to each other or to the end
    points are skipped: there the 20-point sampling of the 0.5-km Euler steps decides which roots are seen."""
    from scipy.optimize import brentq
    K2 = 4.0 * math.pi / 137.0 / 5.0e5 * (1.95e-2 * HBAR / math.sqrt(4.0 * math.pi / 137.0))  # RT.jl:1153-1154

    def omega_p_cart(x, par):  # GJ_Model_ωp_vec, RT.jl:1066-1103 at t = 0
        rr = math.sqrt(x[0] ** 2 + x[1] ** 2 + x[2] ** 2)
        phi_, th_ = math.atan2(x[1], x[0]), math.acos(x[2] / rr)
        Bn = par["B0"] * (par["rNS"] / rr) ** 3 / 2
        Br = 2 * Bn * (math.cos(par["thm"]) * math.cos(th_) + math.sin(par["thm"]) * math.sin(th_) * math.cos(phi_))
        Bt = Bn * (math.cos(par["thm"]) * math.sin(th_) - math.sin(par["thm"]) * math.cos(th_) * math.cos(phi_))
        Bz = Br * math.cos(th_) - Bt * math.sin(th_)
        return math.sqrt(K2 * abs(2.0 * par["w"] * Bz))  # nelec = |2ΩB_z|·…, ωp = √(4π nelec/137/5e5), RT.jl:1089-1090

    rng = np.random.default_rng(seed)
    out = []
    par = dict(thm=0.2, w=1.0, B0=1e14, rNS=10.0, MNS=1.0, ma=1e-5, n_max=6)
    ths = par["thm"] / 2.0  # Find_Conversion_Surface, RT.jl:1250-1263
    maxR = 1.01 * par["rNS"] * (omega_p_cart([par["rNS"] * math.sin(ths), 0.0, par["rNS"] * math.cos(ths)], par) / par["ma"]) ** (2.0 / 3.0)
    while len(out) < n_lines:
        U = rng.random(10)
        thi, phi = math.acos(1 - 2 * U[0]), 2 * math.pi * U[1]
        phR, rR = 2 * math.pi * U[4], math.sqrt(U[5]) * maxR
        va = np.array([math.sin(thi) * math.cos(phi), math.sin(thi) * math.sin(phi), math.cos(thi)])
        x1, x2 = rR * math.cos(phR), rR * math.sin(phR)
        x0 = np.array([x1 * math.cos(-phi) * math.cos(-thi) + x2 * math.sin(-phi), x2 * math.cos(-phi) - x1 * math.sin(-phi) * math.cos(-thi),
                       x1 * math.sin(-thi)]) - 1.1 * maxR * va
        vI = (220.0 + U[6:9] * 1e-5) / math.sqrt(3.0)
        vmag = float(np.linalg.norm(vI))
        erg_inf = par["ma"] * math.sqrt(1 + (vmag / C_KM / math.sqrt(1 - (vmag / C_KM) ** 2)) ** 2)
        thl, phl = math.acos(1 - 2 * U[2]), 2 * math.pi * U[3]
        vl = np.array([math.sin(thl) * math.cos(phl), math.sin(thl) * math.sin(phl), math.cos(thl)])  # vvec_loc, RT.jl:1524

        def g_schwartz(rq, thq, R=10.0):  # RT.jl:455-501 (interior branch below the hard-wired 10 km)
            rs = 2 * GNEW * par["MNS"] / C_KM**2
            if rq <= R:
                rs *= (rq / R) ** 3
                return (-4 / (3 * math.sqrt(1 - rs / R) - math.sqrt(1 - rq**2 * rs / R**3)) ** 2, 1 - rq**2 * rs / R**3,
                        1 / rq**2, 1 / (rq * math.sin(thq)) ** 2)
            return -1 / (1 - rs / rq), 1 - rs / rq, 1 / rq**2, 1 / (rq * math.sin(thq)) ** 2

        def thick(tau):  # condition, RT.jl:1552-1581
            u = x0 + tau * va
            rq = float(np.linalg.norm(u))
            thq, phq = math.acos(u[2] / rq), math.atan2(u[1], u[0])
            AA = 1.0 - 2 * GNEW * par["MNS"] / C_KM**2 / rq
            if rq < par["rNS"]:
                AA = 1.0
            dr = float(np.dot(u, vl)) / rq
            st = math.sin(thq)
            v0 = [dr, (u[2] * dr - rq * vl[2]) / (rq * st), (-u[1] * vl[0] + u[0] * vl[1]) / (rq * st)]
            w = np.array([v0[0] / math.sqrt(AA), v0[1] / rq * rq**2, v0[2] / (rq * st) * (rq * st) ** 2]) / AA
            gtt, grr, gth, gpp = g_schwartz(rq, thq)
            NrmSq = (-erg_inf**2 * gtt - par["ma"] ** 2) / (w[0] ** 2 * grr + w[1] ** 2 * gth + w[2] ** 2 * gpp)
            w = w * math.sqrt(NrmSq)
            omP = omega_p_cart(u, par)
            Bn = par["B0"] * (par["rNS"] / rq) ** 3 / 2  # K_par, RT.jl:1044-1058 with GJ_Model_Sphereical's covariant B (1298)
            Bs = [2 * Bn * (math.cos(par["thm"]) * math.cos(thq) + math.sin(par["thm"]) * st * math.cos(phq)) / math.sqrt(grr),
                  Bn * (math.cos(par["thm"]) * st - math.sin(par["thm"]) * math.cos(thq) * math.cos(phq)) / math.sqrt(gth),
                  Bn * math.sin(par["thm"]) * math.sin(phq) / math.sqrt(gpp)]
            Bmag = math.sqrt(grr * Bs[0] ** 2 + gth * Bs[1] ** 2 + gpp * Bs[2] ** 2)
            kpar = (grr * w[0] * Bs[0] + gth * w[1] * Bs[1] + gpp * w[2] * Bs[2]) / Bmag
            ksqr = gtt * erg_inf**2 + grr * w[0] ** 2 + gth * w[1] ** 2 + gpp * w[2] ** 2
            return 0.5 * (ksqr + omP**2 * (erg_inf**2 / grr - kpar**2) / (erg_inf**2 / grr)) / erg_inf**2

        f = thick if thick_mode else (lambda tau: math.log(omega_p_cart(x0 + tau * va, par)) - math.log(par["ma"]))
        T = 2.2 * maxR
        taus = np.arange(0.0, T, 0.004)
        vals = np.array([f(tq) for tq in taus])
        roots = [brentq(f, taus[i], taus[i + 1], xtol=1e-15, rtol=8.9e-16) for i in range(len(taus) - 1) if vals[i] * vals[i + 1] < 0]
        if len(roots) == 0 or min(np.diff([0.0] + roots + [T])) < 0.1:
            continue
        kept = []
        for tq in roots:  # affect!, RT.jl:1585-1597
            xq = x0 + tq * va
            rq = float(np.linalg.norm(xq))
            grr = 1.0 - 2.0 * GNEW * par["MNS"] / C_KM**2 / rq
            if rq > par["rNS"] and erg_inf / math.sqrt(grr) > omega_p_cart(xq, par):
                kept.append([float(v) for v in xq])
        if not kept:
            continue
        out.append(dict(par=par, maxR=maxR, U=[float(u) for u in U], crossings=kept))
    return out


def main():
    st = sampler_thin_cases()
    with open(os.path.join(HERE, "sampler_thin.json"), "w") as fjs:
        json.dump(st, fjs, indent=1)
    sk = sampler_cases(True, seed=2025)
    with open(os.path.join(HERE, "sampler_thick.json"), "w") as fjs:
        json.dump(sk, fjs, indent=1)
    pw = pointwise()
    with open(os.path.join(HERE, "pointwise.json"), "w") as fjs:
        json.dump(pw, fjs, indent=1)
    nm = pointwise_nonmelrose()
    with open(os.path.join(HERE, "pointwise_nonmelrose.json"), "w") as fjs:
        json.dump(nm, fjs, indent=1)
    pf = pointwise_flat()
    with open(os.path.join(HERE, "pointwise_flat.json"), "w") as fjs:
        json.dump(pf, fjs, indent=1)
    cp = conversion_prob_cases()
    with open(os.path.join(HERE, "conv_prob.json"), "w") as fjs:
        json.dump(cp, fjs, indent=1)
    ry = rays()
    with open(os.path.join(HERE, "rays.json"), "w") as fjs:
        json.dump(ry, fjs, indent=1)
    rf = rays(want_ph=5, want_ax=0, seed=11, flat=True)  # flat = true photons (BASELINE configs[0])
    with open(os.path.join(HERE, "rays_flat.json"), "w") as fjs:
        json.dump(rf, fjs, indent=1)
    rn = rays(ham=hamiltonian_nonmelrose, want_ph=5, want_ax=0, seed=7)  # melrose = false photons
    with open(os.path.join(HERE, "rays_nonmelrose.json"), "w") as fjs:
        json.dump(rn, fjs, indent=1)
    print("pointwise", len(pw), "conv_prob", len(cp), "rays", len(ry))


if __name__ == "__main__":
    main()
