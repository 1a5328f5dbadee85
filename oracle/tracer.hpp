// oracle/tracer.hpp — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the ray right-hand sides func!/func_axion! (RayTracer.jl:71-123) and of
// `propagate` (RayTracer.jl:171-452): initial-condition transform, adaptive integration in
// ln t with the thick-surface crossing event and the NS-hit termination, back-transform.
//
// Third-party arithmetic restated here (NOT in /root/reference; un-vendored, unpinned):
//   OrdinaryDiffEq.jl `solve(prob, Vern6(), reltol=1e-7, abstol=ode_err, dtmin=1e-13,
//   force_dtmin=true, maxiters=1e5)` (RT.jl:383-384) and DiffEqBase `ContinuousCallback(...,
//   rootfind=true, interp_points=50)` / `DiscreteCallback` (RT.jl:357-359).
// Verner's 9-stage 6(5) tableau is not available offline (no network, not on disk), so this
// project fixes its OWN integrator specification and uses it on both sides of the parity
// tests (see DESIGN.md §"Integrator specification"):
//   * Dormand–Prince 5(4) with FSAL and Hairer's 4th-order continuous extension
//     (Hairer, Nørsett, Wanner, "Solving ODEs I", II.5/II.6; public-domain dopri5.f),
//   * OrdinaryDiffEq's error norm (RMS of err_i/(abstol+reltol·max(|u0_i|,|u1_i|))), its
//     Hairer-style initial-step heuristic and its PI controller with the DP5 constants
//     (β1 = 0.17, β2 = 0.04, γ = 0.9, qmin = 0.2, qmax = 10, steady band [1, 1.2]),
//   * DiffEqBase's event semantics: sign scan at `interp_points` equally spaced points of the
//     dense output per accepted step, first sign change is located by the bracketing root finder
//     below (find_root: Illinois + Brent's minimum step + bisection safeguards, 1e-13 of the step),
//     the step is truncated there, the callback runs, integration resumes from the root with
//     a repeat-nudge of dt/100 for the next step's reference sign.
// Accuracy against the converged solution is pinned separately by tests/golden (scipy DOP853,
// rtol 1e-12).
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

#include "physics.hpp"

namespace orc {

enum { RAY_END = 0, RAY_HIT_NS = 1, RAY_CROSSING = 2, RAY_MAXITERS = 3, RAY_CROSS_OVERFLOW = 4 };

struct RayParams {
  NSParams ns;
  double erg;      // erg_inf_ini (Mvars `erg`, MR.jl:178,186)
  double B0;       // signed (back-trace: -B0)
  double reltol, abstol, dtmin;
  int maxiters, interp_points;
  bool is_axion;
};

// func! (RT.jl:71-91) / func_axion! (RT.jl:95-123).  u = [r θ φ k_r/E k_θ/E k_φ/E u7].
inline void rhs(const RayParams& P, const double u[7], double lnt, double du[7]) {
  const NSParams& N = P.ns;
  double t = std::exp(lnt);
  double Mass_NS = N.flat ? 0.0 : N.Mass_NS;  // RT.jl:77-79, 103-105
  double time = 0.0 + t;                      // time0 = zeros(batchsize), MR.jl:177
  Metric<double> g = g_schwartz<double>(u[0], u[1], Mass_NS);
  const double erg = P.erg;
  if (!P.is_axion) {
    using D3 = Dual<3>;
    using D1 = Dual<1>;
    double om = -u[6];
    {  // RT.jl:84
      D3 k[3] = {D3(u[3] * erg), D3(u[4] * erg), D3(u[5] * erg)};
      D3 H = hamiltonian<D3>(D3::seed(u[0], 0), D3::seed(u[1], 1), D3::seed(u[2], 2), k, D3(time),
                             D3(om), N.thm, N.w, P.B0, N.rNS, Mass_NS, N.Mass_a, N.isotropic, false, -1.0, N.melrose);
      for (int i = 0; i < 3; ++i) du[3 + i] = -H.d[i] * c_km * t * (g.grr / om) / erg;
    }
    {  // RT.jl:85
      D3 k[3] = {D3::seed(u[3] * erg, 0), D3::seed(u[4] * erg, 1), D3::seed(u[5] * erg, 2)};
      D3 H = hamiltonian<D3>(D3(u[0]), D3(u[1]), D3(u[2]), k, D3(time), D3(om), N.thm, N.w, P.B0,
                             N.rNS, Mass_NS, N.Mass_a, N.isotropic, false, -1.0, N.melrose);
      for (int i = 0; i < 3; ++i) du[i] = H.d[i] * c_km * t * (g.grr / om);
    }
    if (u[0] <= N.rNS * 1.01)  // RT.jl:86
      for (int i = 0; i < 7; ++i) du[i] = 0.0;
    {  // RT.jl:88 (the only term that sees bndry_lyr)
      D1 k[3] = {D1(u[3] * erg), D1(u[4] * erg), D1(u[5] * erg)};
      D1 H = hamiltonian<D1>(D1(u[0]), D1(u[1]), D1(u[2]), k, D1::seed(time, 0), D1(om), N.thm, N.w,
                             P.B0, N.rNS, Mass_NS, N.Mass_a, N.isotropic, false, N.bndry_lyr, N.melrose);
      du[6] = H.d[0] * t * (g.grr / om);
    }
  } else {
    using D3 = Dual<3>;
    {  // RT.jl:112-115
      D3 k[3] = {D3(u[3] * erg), D3(u[4] * erg), D3(u[5] * erg)};
      D3 H = hamiltonian_axion<D3>(D3::seed(u[0], 0), D3::seed(u[1], 1), k, D3(erg), Mass_NS);
      for (int i = 0; i < 3; ++i) du[3 + i] = -H.d[i] * c_km * t * (g.grr / erg) / erg;
    }
    {  // RT.jl:116-119
      D3 k[3] = {D3::seed(u[3] * erg, 0), D3::seed(u[4] * erg, 1), D3::seed(u[5] * erg, 2)};
      D3 H = hamiltonian_axion<D3>(D3(u[0]), D3(u[1]), k, D3(erg), Mass_NS);
      for (int i = 0; i < 3; ++i) du[i] = H.d[i] * c_km * t * (g.grr / erg);
    }
    du[6] = 0.0;  // RT.jl:120
  }
}

// `condition` inside propagate, RT.jl:254-298 (thick_surface hard-wired true at RT.jl:255).
inline double ray_condition(const RayParams& P, const double u[7], double lnt) {
  const NSParams& N = P.ns;
  double Mass_NS = N.flat ? 0.0 : N.Mass_NS;  // RT.jl:187-189 (closure sees the reassigned value)
  double sph[3] = {u[0], u[1], u[2]};
  double w[3] = {u[3], u[4], u[5]};
  double erg_inf = u[6];
  double t0 = std::exp(lnt);
  // RT.jl:284 — zeroIn defaults to true here
  double omP = omega_p_sph<double>(u[0], u[1], u[2], t0, N.thm, N.w, P.B0, N.rNS, true, N.bndry_lyr, N.Mass_a);
  return thick_condition(sph, w, erg_inf, omP, t0, Mass_NS, N, P.B0);
}

struct Crossing {
  double x, y, z, kx, ky, kz, t, dw;
};

// false (default): start-box rule on.  true: locate and then discard start-point events exactly like
// the reference does (used by tests/test_oracle_invariants.py to bound the effect of the rule).
static bool g_literal_start_events = false;

struct StepTrace {
  double s, dt, EEst;
  int accepted, event;
};

struct RayResult {
  double xf[3], kf[3], u7f, lnt_end;
  std::vector<StepTrace>* trace = nullptr;  // optional step log (tests / diagnostics)
  int status;
  std::vector<Crossing> cr;
  int n_acc = 0, n_rej = 0, n_rhs = 0, n_cond = 0;
};

// Dormand–Prince 5(4) coefficients (Hairer's dopri5).
namespace dp {
constexpr double c2 = 1.0 / 5, c3 = 3.0 / 10, c4 = 4.0 / 5, c5 = 8.0 / 9;
constexpr double a21 = 1.0 / 5;
constexpr double a31 = 3.0 / 40, a32 = 9.0 / 40;
constexpr double a41 = 44.0 / 45, a42 = -56.0 / 15, a43 = 32.0 / 9;
constexpr double a51 = 19372.0 / 6561, a52 = -25360.0 / 2187, a53 = 64448.0 / 6561, a54 = -212.0 / 729;
constexpr double a61 = 9017.0 / 3168, a62 = -355.0 / 33, a63 = 46732.0 / 5247, a64 = 49.0 / 176,
                 a65 = -5103.0 / 18656;
constexpr double a71 = 35.0 / 384, a73 = 500.0 / 1113, a74 = 125.0 / 192, a75 = -2187.0 / 6784,
                 a76 = 11.0 / 84;
constexpr double e1 = 71.0 / 57600, e3 = -71.0 / 16695, e4 = 71.0 / 1920, e5 = -17253.0 / 339200,
                 e6 = 22.0 / 525, e7 = -1.0 / 40;
constexpr double d1 = -12715105075.0 / 11282082432.0, d3 = 87487479700.0 / 32700410799.0,
                 d4 = -10690763975.0 / 1880347072.0, d5 = 701980252875.0 / 199316789632.0,
                 d6 = -1453857185.0 / 822651844.0, d7 = 69997945.0 / 29380423.0;
constexpr double beta1 = 0.17, beta2 = 0.04, gamma = 0.9, qmin = 0.2, qmax = 10.0;
constexpr double qsteady_min = 1.0, qsteady_max = 1.2, qoldinit = 1e-4;
}  // namespace dp

// Bracketing root finder of the event location (stand-in for DiffEqBase's find_callback_time):
// Illinois-modified regula falsi with Brent's minimum step (an iterate that has converged moves by the
// tolerance towards the other end, which collapses the bracket); bisection whenever the bracket has not
// halved over two iterations, and for good once three iterations have not brought |f| down (a jump has
// no zero to converge to).  Tolerance 1e-13 of the step.  Returns the LEFT end (in time) of the final
// bracket, like DiffEqBase's left-biased root.
constexpr double ROOT_RTOL = 1e-13;
template <class F>
inline double find_root(F&& f, double a, double fa, double b, double fb, int* n_eval) {
  // invariant: fa*fb < 0 ; a is the "old" end, b the most recent iterate
  int side = 0;
  const double f0 = std::min(std::fabs(fa), std::fabs(fb));
  double w1 = 2.0, w2 = 2.0;
  bool pure = false;
  for (int it = 0; it < 200; ++it) {
    double width = std::fabs(b - a);
    double tol = ROOT_RTOL * std::max(std::max(std::fabs(a), std::fabs(b)), 0.01);  // 1e-13 of the step, floor 1e-15
    if (width <= tol) break;
    if (it == 3 && std::min(std::fabs(fa), std::fabs(fb)) > 0.25 * f0) pure = true;
    double c = (a * fb - b * fa) / (fb - fa);
    double lo = std::min(a, b), hi = std::max(a, b);
    if (pure || (it >= 3 && width > 0.5 * w2) || !(c > lo && c < hi) || width <= 2.0 * tol) c = 0.5 * (a + b);
    else if (std::fabs(c - b) < tol) c = b + std::copysign(tol, a - b);
    w2 = w1;
    w1 = width;
    double fc = f(c);
    ++*n_eval;
    if (fc == 0.0) return c;
    if ((fc < 0) != (fb < 0)) {  // root between b and c: old b becomes a
      a = b;
      fa = fb;
      side = 0;
    } else {  // same side as b: keep a, Illinois halving on repeated retention
      if (side == 1) fa *= 0.5;
      side = 1;
    }
    b = c;
    fb = fc;
  }
  return std::min(a, b);
}

inline double rms7(const double v[7]) {
  double s = 0;
  for (int i = 0; i < 7; ++i) s += v[i] * v[i];
  return std::sqrt(s / 7.0);
}

// propagate, RT.jl:171-452, for one node.  x0,k0 Cartesian.
inline RayResult propagate(const double x0[3], const double k0_in[3], const RayParams& P,
                           double ln_tstart, double ln_tend, int max_crossings, double dOmega,
                           int cross_cap = 1 << 20, std::vector<StepTrace>* trace = nullptr) {
  const NSParams& N = P.ns;
  RayResult R;
  R.trace = trace;
  // RT.jl:179-186: renormalise |k| to the axion mass shell (real Mass_NS, even if flat)
  double k0[3];
  k_norm_cart(x0, k0_in, P.erg, N.Mass_NS, N.Mass_a, k0);
  double Mass_NS = N.flat ? 0.0 : N.Mass_NS;  // RT.jl:187-189
  double r_s0 = 2.0 * Mass_NS * GNew / (c_km * c_km);
  double s0[3];
  cart_to_sph(x0, s0);  // RT.jl:197-199
  double AA = 1.0 - r_s0 / s0[0];
  double w0[3];
  celerity(x0, s0, k0, AA, w0);  // RT.jl:205-211
  double u[7] = {s0[0], s0[1], s0[2], w0[0] / P.erg, w0[1] / P.erg, w0[2] / P.erg, P.erg * dOmega};  // RT.jl:212-216

  const double rtol = P.reltol, atol = P.abstol;
  const int M = P.interp_points;
  double s = ln_tstart, s1 = ln_tend;
  double k1[7], k2[7], k3[7], k4[7], k5[7], k6[7], k7[7], un[7], tmp[7], err[7];
  rhs(P, u, s, k1);
  R.n_rhs++;

  // ---- initial step (OrdinaryDiffEq ode_determine_initdt, Hairer II.4) -------------------
  double dt;
  {
    double sk[7], a[7];
    for (int i = 0; i < 7; ++i) sk[i] = atol + std::fabs(u[i]) * rtol;
    for (int i = 0; i < 7; ++i) a[i] = u[i] / sk[i];
    double d0 = rms7(a);
    for (int i = 0; i < 7; ++i) a[i] = k1[i] / sk[i];
    double d1 = rms7(a);
    double dtmax = s1 - s;
    double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
    dt0 = std::min(dt0, dtmax);
    for (int i = 0; i < 7; ++i) tmp[i] = u[i] + dt0 * k1[i];
    double f1[7];
    rhs(P, tmp, s + dt0, f1);
    R.n_rhs++;
    for (int i = 0; i < 7; ++i) a[i] = (f1[i] - k1[i]) / sk[i];
    double d2 = rms7(a) / dt0;
    double dm = std::max(d1, d2);
    double dt1 = (dm <= 1e-15) ? std::max(1e-6, dt0 * 1e-3) : std::pow(10.0, -(2.0 + std::log10(dm)) / 5.0);
    dt = std::min(std::min(100.0 * dt0, dt1), dtmax);
  }

  double qold = dp::qoldinit;
  bool event_last = false;
  double g_prev = ray_condition(P, u, s);
  R.n_cond++;
  int callback_count = 0;
  int iters = 0;
  R.status = RAY_END;
  bool first_filter_active = true;  // callback_count == 0 branch of affect!, RT.jl:303-314

  while (s < s1) {
    if (++iters > P.maxiters) {
      R.status = RAY_MAXITERS;
      break;
    }
    bool last_step = false;
    if (s + dt >= s1) {
      dt = s1 - s;
      last_step = true;
    }
    // ---- stages --------------------------------------------------------------------------
    for (int i = 0; i < 7; ++i) tmp[i] = u[i] + dt * (dp::a21 * k1[i]);
    rhs(P, tmp, s + dp::c2 * dt, k2);
    for (int i = 0; i < 7; ++i) tmp[i] = u[i] + dt * (dp::a31 * k1[i] + dp::a32 * k2[i]);
    rhs(P, tmp, s + dp::c3 * dt, k3);
    for (int i = 0; i < 7; ++i) tmp[i] = u[i] + dt * (dp::a41 * k1[i] + dp::a42 * k2[i] + dp::a43 * k3[i]);
    rhs(P, tmp, s + dp::c4 * dt, k4);
    for (int i = 0; i < 7; ++i)
      tmp[i] = u[i] + dt * (dp::a51 * k1[i] + dp::a52 * k2[i] + dp::a53 * k3[i] + dp::a54 * k4[i]);
    rhs(P, tmp, s + dp::c5 * dt, k5);
    for (int i = 0; i < 7; ++i)
      tmp[i] = u[i] + dt * (dp::a61 * k1[i] + dp::a62 * k2[i] + dp::a63 * k3[i] + dp::a64 * k4[i] + dp::a65 * k5[i]);
    rhs(P, tmp, s + dt, k6);
    for (int i = 0; i < 7; ++i)
      un[i] = u[i] + dt * (dp::a71 * k1[i] + dp::a73 * k3[i] + dp::a74 * k4[i] + dp::a75 * k5[i] + dp::a76 * k6[i]);
    double s_new = last_step ? s1 : s + dt;
    rhs(P, un, s_new, k7);
    R.n_rhs += 6;
    for (int i = 0; i < 7; ++i) {
      double e = dt * (dp::e1 * k1[i] + dp::e3 * k3[i] + dp::e4 * k4[i] + dp::e5 * k5[i] + dp::e6 * k6[i] + dp::e7 * k7[i]);
      err[i] = e / (atol + std::max(std::fabs(u[i]), std::fabs(un[i])) * rtol);
    }
    double EEst = rms7(err);
    // ---- PI controller (OrdinaryDiffEq PIController, DP5 constants) ----------------------
    double q11 = std::pow(EEst, dp::beta1);
    double q = q11 / std::pow(qold, dp::beta2);
    q = std::max(1.0 / dp::qmax, std::min(1.0 / dp::qmin, q / dp::gamma));
    if (EEst == 0.0) q = 1.0 / dp::qmax;
    bool accept = (EEst <= 1.0) || (dt <= P.dtmin);  // force_dtmin
    if (!(EEst == EEst)) {  // NaN error estimate: shrink like a rejection
      accept = false;
      q11 = 1e30;
    }
    if (R.trace) R.trace->push_back({s, dt, EEst, accept ? 1 : 0, 0});
    if (!accept) {
      R.n_rej++;
      dt = dt / std::min(1.0 / dp::qmin, q11 / dp::gamma);
      if (dt < P.dtmin) dt = P.dtmin;
      continue;
    }
    R.n_acc++;
    if (q <= dp::qsteady_max && q >= dp::qsteady_min) q = 1.0;
    qold = std::max(EEst, dp::qoldinit);
    double dt_new = dt / q;

    // ---- continuous callback: thick-surface crossings ------------------------------------
    bool terminated = false, had_event = false;
    double rc[5][7];
    for (int i = 0; i < 7; ++i) {  // Hairer contd5
      rc[0][i] = u[i];
      rc[1][i] = un[i] - u[i];
      rc[2][i] = dt * k1[i] - rc[1][i];
      rc[3][i] = rc[1][i] - dt * k7[i] - rc[2][i];
      rc[4][i] = dt * (dp::d1 * k1[i] + dp::d3 * k3[i] + dp::d4 * k4[i] + dp::d5 * k5[i] + dp::d6 * k6[i] + dp::d7 * k7[i]);
    }
    auto dense = [&](double th, double out[7]) {
      double th1 = 1.0 - th;
      for (int i = 0; i < 7; ++i)
        out[i] = rc[0][i] + th * (rc[1][i] + th1 * (rc[2][i] + th * (rc[3][i] + th1 * rc[4][i])));
    };
    auto cond_at = [&](double th) {
      double v[7];
      if (th >= 1.0) {
        return ray_condition(P, un, s_new);
      }
      dense(th, v);
      return ray_condition(P, v, s + th * dt);
    };
    double g_end = 0.0;
    {
      double prev = g_prev;
      double th_left = 0.0;
      if (event_last) {  // DiffEqBase repeat_nudge = 1//100
        th_left = 0.01;
        prev = cond_at(th_left);
        R.n_cond++;
      }
      double th_a = th_left, g_a = prev;
      int sgn_prev = (prev > 0) - (prev < 0);
      double th_ev = -1.0;
      for (int i = 1; i < M; ++i) {
        double th = (i == M - 1) ? 1.0 : (double)i * (1.0 / (double)(M - 1));
        double gv = cond_at(th);
        R.n_cond++;
        if (i == M - 1) g_end = gv;
        if (th <= th_left) continue;
        int sg = (gv > 0) - (gv < 0);
        if (sgn_prev != 0 && sgn_prev * sg < 0 && callback_count == 0 && !g_literal_start_events) {
          // Start-box rule (this project's integrator specification): a ray starts ON
          // the conversion surface, so while it has not moved the condition is round-off noise and
          // flips sign at random.  The reference locates every such flip with the root finder and then
          // discards it in affect! (RT.jl:303-314, the ±0.01 % box around the start point).  Here a sign
          // change whose grid point still lies inside that box is discarded WITHOUT locating it: the
          // reference sign flips and the scan continues.  Recorded crossings are the same; only the
          // (noise-driven, irreproducible) step truncations at discarded events disappear.
          double v[7];
          if (th >= 1.0) for (int c = 0; c < 7; ++c) v[c] = un[c]; else dense(th, v);
          double pos[3] = {v[0] * std::sin(v[1]) * std::cos(v[2]), v[0] * std::sin(v[1]) * std::sin(v[2]), v[0] * std::cos(v[1])};
          const double sb = 1.0001;
          bool in = true;
          for (int c = 0; c < 3; ++c)
            in = in && (std::fabs(pos[c]) < std::fabs(x0[c]) * sb) && (std::fabs(pos[c]) > std::fabs(x0[c]) / sb);
          if (in) {
            sgn_prev = sg;
            th_a = th;
            g_a = gv;
            continue;
          }
        }
        if (sgn_prev != 0 && sgn_prev * sg < 0) {
          int ne = 0;
          th_ev = find_root(cond_at, th_a, g_a, th, gv, &ne);
          R.n_cond += ne;
          break;
        }
        // DiffEqBase compares every point against the step's prev_sign, but brackets the
        // root with the last same-sign point — keep the tightest left bracket.
        if (sg == sgn_prev) {
          th_a = th;
          g_a = gv;
        }
      }
      if (th_ev < 0 && sgn_prev != 0 && g_end == 0.0) th_ev = 1.0;  // prev_sign*next_sign <= 0
      if (th_ev >= 0) {
        had_event = true;
        if (R.trace) R.trace->back().event = 1;
        double ue[7];
        double se;
        if (th_ev >= 1.0) {
          for (int i = 0; i < 7; ++i) ue[i] = un[i];
          se = s_new;
        } else {
          dense(th_ev, ue);
          se = s + th_ev * dt;
        }
        // ---- affect!, RT.jl:301-350 ------------------------------------------------------
        bool skip = false;
        double xpos = ue[0] * std::sin(ue[1]) * std::cos(ue[2]);
        double ypos = ue[0] * std::sin(ue[1]) * std::sin(ue[2]);
        double zpos = ue[0] * std::cos(ue[1]);
        if (callback_count == 0 && first_filter_active) {
          const double sb = 1.0001;
          double pos[3] = {xpos, ypos, zpos};
          bool in = true;
          for (int c = 0; c < 3; ++c)
            in = in && (std::fabs(pos[c]) < std::fabs(x0[c]) * sb) && (std::fabs(pos[c]) > std::fabs(x0[c]) / sb);
          if (in) skip = true;
        }
        if (!skip && std::sqrt(xpos * xpos + ypos * ypos + zpos * zpos) < N.rNS * 1.01) skip = true;  // RT.jl:322
        if (!skip) {
          Crossing c;
          c.x = xpos; c.y = ypos; c.z = zpos;
          c.t = std::exp(se);          // RT.jl:328
          c.dw = ue[6] / P.erg;        // RT.jl:329
          double r_s = 2.0 * Mass_NS * GNew / (c_km * c_km);  // RT.jl:335
          double om = 1.0 - r_s / ue[0];
          double v0 = ue[3] * std::sqrt(om), v1 = ue[4] / ue[0], v2 = ue[5] / (ue[0] * std::sin(ue[1]));
          v0 *= P.erg * om; v1 *= P.erg * om; v2 *= P.erg * om;  // RT.jl:337-338
          double vt = std::sin(ue[1]) * v0 + std::cos(ue[1]) * v1;
          c.kx = std::cos(ue[2]) * vt - std::sin(ue[2]) * v2;
          c.ky = std::sin(ue[2]) * vt + std::cos(ue[2]) * v2;
          c.kz = std::cos(ue[1]) * v0 - std::sin(ue[1]) * v1;
          if ((int)R.cr.size() < cross_cap) R.cr.push_back(c);
          else R.status = RAY_CROSS_OVERFLOW;
          callback_count++;
          if (callback_count >= max_crossings) {  // RT.jl:345-349
            terminated = true;
            if (R.status != RAY_CROSS_OVERFLOW) R.status = RAY_CROSSING;
          }
        }
        // step truncated at the event; resume from there
        for (int i = 0; i < 7; ++i) u[i] = ue[i];
        s = se;
        if (!terminated) {
          rhs(P, u, s, k1);
          R.n_rhs++;
        }
        event_last = true;
      }
    }
    if (!had_event) {
      for (int i = 0; i < 7; ++i) {
        u[i] = un[i];
        k1[i] = k7[i];
      }
      s = s_new;
      g_prev = g_end;
      event_last = false;
    }
    dt = dt_new;
    if (terminated) break;
    // ---- discrete callback, RT.jl:352-359 (photons only) ---------------------------------
    if (!P.is_axion && u[0] < N.rNS * 1.01) {
      R.status = RAY_HIT_NS;
      break;
    }
  }

  // ---- back-transform of the last state, RT.jl:393-416 -----------------------------------
  R.lnt_end = s;
  R.u7f = u[6];
  double kk[3] = {u[3] * P.erg, u[4] * P.erg, u[5] * P.erg};
  double M_eff = Mass_NS;
  if (u[0] < N.rNS) M_eff *= u[0] * u[0] * u[0] / (N.rNS * N.rNS * N.rNS);  // RT.jl:398-402
  double r_s = 2.0 * M_eff * GNew / (c_km * c_km);
  double om = 1.0 - r_s / u[0];
  double st = std::sin(u[1]), ct = std::cos(u[1]), sp = std::sin(u[2]), cp = std::cos(u[2]);
  double v0 = kk[0] * std::sqrt(om) * om, v1 = kk[1] / u[0] * om, v2 = kk[2] / (u[0] * st) * om;  // RT.jl:411
  R.xf[0] = u[0] * st * cp;
  R.xf[1] = u[0] * st * sp;
  R.xf[2] = u[0] * ct;
  R.kf[0] = cp * (st * v0 + ct * v1) - st * sp * v2 / st;  // RT.jl:416
  R.kf[1] = sp * (st * v0 + ct * v1) + st * cp * v2 / st;
  R.kf[2] = ct * v0 - st * v1;
  return R;
}

}  // namespace orc
