// art_tracer.cuh — one-ray-per-thread adaptive integrator with level-crossing events.
//
// Replaces `propagate` (reference RayTracer.jl:171-452): OrdinaryDiffEq's adaptive solve in
// ln t with a ContinuousCallback (thick-surface crossings, root-found) and a DiscreteCallback
// (photon hits the star).  Integrator specification: Dormand–Prince 5(4), FSAL,
// Hairer's continuous extension, OrdinaryDiffEq's error norm / initial step / PI controller,
// DiffEqBase's event scan at `interp_points` dense-output points per accepted step.
//
// State lives in registers: six integrated components u = [r θ φ k_r/E k_θ/E k_φ/E]; the 7th
// component of the reference's state (u7 = −ω) is carried by the exact linear invariant
// u7 = u7₀ + (Ω·erg/c_km)(u6 − u6₀) of the photon equations (0 slope for axions), which every
// Runge–Kutta method preserves, and still contributes to the error norm like in the reference.
#pragma once
#include "art_physics.cuh"

namespace art {

enum { RAY_END = 0, RAY_HIT_NS = 1, RAY_CROSSING = 2, RAY_MAXITERS = 3, RAY_CROSS_OVERFLOW = 4 };

// Dormand–Prince 5(4) (Hairer, Nørsett, Wanner, Solving ODEs I, Table II.5.2; dense output II.6)
namespace dp {
constexpr double c2 = 1.0 / 5, c3 = 3.0 / 10, c4 = 4.0 / 5, c5 = 8.0 / 9;
constexpr double a21 = 1.0 / 5;
constexpr double a31 = 3.0 / 40, a32 = 9.0 / 40;
constexpr double a41 = 44.0 / 45, a42 = -56.0 / 15, a43 = 32.0 / 9;
constexpr double a51 = 19372.0 / 6561, a52 = -25360.0 / 2187, a53 = 64448.0 / 6561, a54 = -212.0 / 729;
constexpr double a61 = 9017.0 / 3168, a62 = -355.0 / 33, a63 = 46732.0 / 5247, a64 = 49.0 / 176,
                 a65 = -5103.0 / 18656;
constexpr double b1 = 35.0 / 384, b3 = 500.0 / 1113, b4 = 125.0 / 192, b5 = -2187.0 / 6784, b6 = 11.0 / 84;
constexpr double e1 = 71.0 / 57600, e3 = -71.0 / 16695, e4 = 71.0 / 1920, e5 = -17253.0 / 339200,
                 e6 = 22.0 / 525, e7 = -1.0 / 40;
constexpr double d1 = -12715105075.0 / 11282082432.0, d3 = 87487479700.0 / 32700410799.0,
                 d4 = -10690763975.0 / 1880347072.0, d5 = 701980252875.0 / 199316789632.0,
                 d6 = -1453857185.0 / 822651844.0, d7 = 69997945.0 / 29380423.0;
// OrdinaryDiffEq PIController constants for DP5
constexpr double beta1 = 0.17, beta2 = 0.04, gamma = 0.9, qmin = 0.2, qmax = 10.0;
constexpr double qsteady_min = 1.0, qsteady_max = 1.2, qoldinit = 1e-4;
}  // namespace dp

struct RayIn {
  double x0[3], k0[3];  // Cartesian start position [km] and momentum direction·magnitude [eV]
  double erg;           // erg_inf_ini
  double t_start;       // node.t [s]
  double dOmega;        // node.Δω
  int max_cross;        // get_tree's splittings_cutoff as passed to propagate
};

struct RayOut {
  double xf[3], kf[3], u7f;
  int status, n_cross;
  int n_acc, n_rej, n_rhs, n_cond;
};

template <bool AXION>
struct Ray {
  double u[6], k1[6];
  double s, dt, qold, g_prev;
  double erg, fac7, u7_0, u6_0;
  double x0[3];
  double* rec;  // crossing records, ART_CROSS_REC doubles each
  int rec_cap, max_cross;
  int n_cross, iters, status;
  int n_acc, n_rej, n_rhs, n_cond;
  bool event_last, done;

  ART_HD double u7_of(double u6) const { return AXION ? u7_0 : u7_0 + fac7 * (u6 - u6_0); }

  ART_HD void f(const DevParams& P, const double v[6], double lnt, double dv[6]) const {
    const double t = exp(lnt);
    if (AXION)
      rhs_axion(P, v[0], v[1], v[3] * erg, v[4] * erg, v[5] * erg, t, erg, dv);
    else
      rhs_photon(P, v[0], v[1], v[2], v[3] * erg, v[4] * erg, v[5] * erg, -u7_of(v[5]), t, erg, dv);
  }

  // `condition` of propagate, RT.jl:254-298
  ART_HD double cond(const DevParams& P, const double v[6], double lnt) const {
#ifdef ART_COUNT_FAR
    
    g_cond_total++;
    if (v[0] > P.Rcut) g_cond_far++;
    if (v[0] <= P.rNS) g_cond_in++;
    if (lnt < -14.0) g_cond_early++;
#endif
    const double t = exp(lnt);
    return thick_condition(P, v[0], v[1], v[2] - P.Omega * t, v[3], v[4], v[5], u7_of(v[5]), P.rs_ray, true);
  }

  ART_HD static double rms7(const double a[6], double a7) {
    double q = a7 * a7;
    for (int i = 0; i < 6; ++i) q += a[i] * a[i];
    return sqrt(q / 7.0);
  }

  // ---- set-up: IC transform (RT.jl:179-216), first RHS, initial step --------------------
  ART_HD void init(const DevParams& P, const RayIn& in, double* rec_, int rec_cap_) {
    rec = rec_;
    rec_cap = rec_cap_;
    max_cross = in.max_cross;
    erg = in.erg;
    for (int i = 0; i < 3; ++i) x0[i] = in.x0[i];
    ray_initial_state(P, in.x0, in.k0, erg, in.dOmega, u, u7_0);
    u6_0 = u[5];
    fac7 = AXION ? 0.0 : P.Omega * erg / C_KM;
    s = log(fmax(in.t_start, P.t_floor));  // MR.jl:166
    n_cross = iters = n_acc = n_rej = n_rhs = n_cond = 0;
    status = RAY_END;
    event_last = false;
    done = false;
    qold = dp::qoldinit;
    f(P, u, s, k1);
    n_rhs = 1;
    // OrdinaryDiffEq ode_determine_initdt (Hairer II.4), order 5, 7-component RMS norm
    {
      const double rtol = P.rtol, atol = P.atol;
      double a[6], sk[6];
      for (int i = 0; i < 6; ++i) sk[i] = atol + fabs(u[i]) * rtol;
      const double sk7 = atol + fabs(u7_0) * rtol;
      for (int i = 0; i < 6; ++i) a[i] = u[i] / sk[i];
      const double d0 = rms7(a, u7_0 / sk7);
      for (int i = 0; i < 6; ++i) a[i] = k1[i] / sk[i];
      const double d1 = rms7(a, fac7 * k1[5] / sk7);
      const double dtmax = P.ln_tend - s;
      double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
      dt0 = fmin(dt0, dtmax);
      double v[6], f1[6];
      for (int i = 0; i < 6; ++i) v[i] = u[i] + dt0 * k1[i];
      f(P, v, s + dt0, f1);
      n_rhs++;
      for (int i = 0; i < 6; ++i) a[i] = (f1[i] - k1[i]) / sk[i];
      const double d2 = rms7(a, fac7 * (f1[5] - k1[5]) / sk7) / dt0;
      const double dm = fmax(d1, d2);
      const double dt1 = (dm <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(dm)) / 5.0);
      dt = fmin(fmin(100.0 * dt0, dt1), dtmax);
    }
    g_prev = cond(P, u, s);
    n_cond = 1;
    if (!(s < P.ln_tend)) done = true;
  }

  // ---- affect! of the crossing callback, RT.jl:301-350.  Returns true to terminate. -------
  ART_HD bool on_crossing(const DevParams& P, const double ue[6], double se) {
    double x[3], k[3];
    state_to_cartesian(P, ue, erg, false, x, k);
    if (n_cross == 0) {  // "callback_count == 0": ignore the crossing the ray starts on
      const double sb = 1.0001;
      bool in = true;
      for (int c = 0; c < 3; ++c)
        in = in && (fabs(x[c]) < fabs(x0[c]) * sb) && (fabs(x[c]) > fabs(x0[c]) / sb);
      if (in) return false;
    }
    if (sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]) < P.rNS101) return false;  // RT.jl:322
    if (n_cross < rec_cap) {
      double* q = rec + (size_t)n_cross * 8;
      q[0] = x[0]; q[1] = x[1]; q[2] = x[2];
      q[3] = k[0]; q[4] = k[1]; q[5] = k[2];
      q[6] = exp(se);                // RT.jl:328
      q[7] = u7_of(ue[5]) / erg;     // RT.jl:329
    } else {
      status = RAY_CROSS_OVERFLOW;
    }
    n_cross++;
    if (n_cross >= max_cross) {      // RT.jl:345-349
      if (status != RAY_CROSS_OVERFLOW) status = RAY_CROSSING;
      return true;
    }
    return false;
  }

  // ---- one attempted step (stages, error control, event scan, callbacks) -----------------
  ART_HD void step(const DevParams& P) {
    if (++iters > P.maxiters) {
      status = RAY_MAXITERS;
      done = true;
      return;
    }
    const double s1 = P.ln_tend;
    double h = dt;
    bool last = false;
    if (s + h >= s1) {
      h = s1 - s;
      last = true;
    }
    double k2[6], k3[6], k4[6], k5[6], k6[6], k7[6], un[6], v[6];
    for (int i = 0; i < 6; ++i) v[i] = u[i] + h * (dp::a21 * k1[i]);
    f(P, v, s + dp::c2 * h, k2);
    for (int i = 0; i < 6; ++i) v[i] = u[i] + h * (dp::a31 * k1[i] + dp::a32 * k2[i]);
    f(P, v, s + dp::c3 * h, k3);
    for (int i = 0; i < 6; ++i) v[i] = u[i] + h * (dp::a41 * k1[i] + dp::a42 * k2[i] + dp::a43 * k3[i]);
    f(P, v, s + dp::c4 * h, k4);
    for (int i = 0; i < 6; ++i)
      v[i] = u[i] + h * (dp::a51 * k1[i] + dp::a52 * k2[i] + dp::a53 * k3[i] + dp::a54 * k4[i]);
    f(P, v, s + dp::c5 * h, k5);
    for (int i = 0; i < 6; ++i)
      v[i] = u[i] + h * (dp::a61 * k1[i] + dp::a62 * k2[i] + dp::a63 * k3[i] + dp::a64 * k4[i] + dp::a65 * k5[i]);
    f(P, v, s + h, k6);
    for (int i = 0; i < 6; ++i)
      un[i] = u[i] + h * (dp::b1 * k1[i] + dp::b3 * k3[i] + dp::b4 * k4[i] + dp::b5 * k5[i] + dp::b6 * k6[i]);
    const double s_new = last ? s1 : s + h;
    f(P, un, s_new, k7);
    n_rhs += 6;
    // error estimate, RMS over the reference's 7 components
    double EEst;
    {
      double q = 0.0, e6 = 0.0;
      for (int i = 0; i < 6; ++i) {
        const double e = h * (dp::e1 * k1[i] + dp::e3 * k3[i] + dp::e4 * k4[i] + dp::e5 * k5[i] + dp::e6 * k6[i] +
                              dp::e7 * k7[i]);
        if (i == 5) e6 = e;
        const double r_ = e / (P.atol + fmax(fabs(u[i]), fabs(un[i])) * P.rtol);
        q += r_ * r_;
      }
      const double r7 = fac7 * e6 / (P.atol + fmax(fabs(u7_of(u[5])), fabs(u7_of(un[5]))) * P.rtol);
      q += r7 * r7;
      EEst = sqrt(q / 7.0);
    }
    // PI controller
    double q11 = pow(EEst, dp::beta1);
    double q = q11 / pow(qold, dp::beta2);
    q = fmax(1.0 / dp::qmax, fmin(1.0 / dp::qmin, q / dp::gamma));
    if (EEst == 0.0) q = 1.0 / dp::qmax;
    bool accept = (EEst <= 1.0) || (h <= P.dtmin);  // force_dtmin
    if (!(EEst == EEst)) {
      accept = false;
      q11 = 1e30;
    }
    if (!accept) {
      n_rej++;
      dt = h / fmin(1.0 / dp::qmin, q11 / dp::gamma);
      if (dt < P.dtmin) dt = P.dtmin;
      return;
    }
    n_acc++;
    if (q <= dp::qsteady_max && q >= dp::qsteady_min) q = 1.0;
    qold = fmax(EEst, dp::qoldinit);
    const double dt_new = h / q;

    // ---- crossing scan on the dense output ------------------------------------------------
    double r1[6], r2[6], r3[6], r4[6];  // Hairer contd5 (rcont2..rcont5); rcont1 = u
    for (int i = 0; i < 6; ++i) {
      r1[i] = un[i] - u[i];
      r2[i] = h * k1[i] - r1[i];
      r3[i] = r1[i] - h * k7[i] - r2[i];
      r4[i] = h * (dp::d1 * k1[i] + dp::d3 * k3[i] + dp::d4 * k4[i] + dp::d5 * k5[i] + dp::d6 * k6[i] + dp::d7 * k7[i]);
    }
    auto dense = [&](double th, double out[6]) {
      const double th1 = 1.0 - th;
      for (int i = 0; i < 6; ++i) out[i] = u[i] + th * (r1[i] + th1 * (r2[i] + th * (r3[i] + th1 * r4[i])));
    };
    auto cond_at = [&](double th) {
      if (th >= 1.0) return cond(P, un, s_new);
      double w_[6];
      dense(th, w_);
      return cond(P, w_, s + th * h);
    };
    const int M = P.M_ray;
    double prev = g_prev, th_left = 0.0;
    if (event_last) {  // DiffEqBase repeat_nudge = 1/100
      th_left = 0.01;
      prev = cond_at(th_left);
      n_cond++;
    }
    const int sgn_prev = (prev > 0.0) - (prev < 0.0);
    double th_a = th_left, g_a = prev, th_ev = -1.0, g_end = 0.0;
    double th_b = 0.0, g_b = 0.0;
    bool bracket = false;
    for (int i = 1; i < M; ++i) {
      const double th = (i == M - 1) ? 1.0 : (double)i / (double)(M - 1);
      const double gv = cond_at(th);
      n_cond++;
      if (i == M - 1) g_end = gv;
      if (th <= th_left) continue;
      const int sg = (gv > 0.0) - (gv < 0.0);
      if (sgn_prev != 0 && sgn_prev * sg < 0) {
        th_b = th;
        g_b = gv;
        bracket = true;
        break;
      }
      if (sg == sgn_prev) {
        th_a = th;
        g_a = gv;
      }
    }
    if (bracket) {
      // Illinois regula falsi with bisection safeguard, to floating precision; left-biased
      double a = th_a, fa = g_a, b = th_b, fb = g_b;
      int side = 0;
      bool exact = false;
      for (int it = 0; it < 200; ++it) {
        const double width = fabs(b - a);
        const double tol = 4.0 * 2.220446049250313e-16 * fmax(fabs(a), fabs(b)) + 1e-300;
        if (width <= tol) break;
        double c = (a * fb - b * fa) / (fb - fa);
        const double lo = fmin(a, b), hi = fmax(a, b);
        if (!(c > lo && c < hi) || it >= 60) c = 0.5 * (a + b);
        const double fc = cond_at(c);
        n_cond++;
        if (fc == 0.0) {
          th_ev = c;
          exact = true;
          break;
        }
        if ((fc < 0.0) != (fb < 0.0)) {
          a = b;
          fa = fb;
          side = 0;
        } else {
          if (side == 1) fa *= 0.5;
          side = 1;
        }
        b = c;
        fb = fc;
      }
      if (!exact) th_ev = fmin(a, b);
    } else if (sgn_prev != 0 && g_end == 0.0) {
      th_ev = 1.0;
    }

    bool terminated = false;
    if (th_ev >= 0.0) {
      double ue[6];
      double se;
      if (th_ev >= 1.0) {
        for (int i = 0; i < 6; ++i) ue[i] = un[i];
        se = s_new;
      } else {
        dense(th_ev, ue);
        se = s + th_ev * h;
      }
      terminated = on_crossing(P, ue, se);
      for (int i = 0; i < 6; ++i) u[i] = ue[i];
      s = se;
      if (!terminated) {
        f(P, u, s, k1);
        n_rhs++;
      }
      event_last = true;
    } else {
      for (int i = 0; i < 6; ++i) {
        u[i] = un[i];
        k1[i] = k7[i];
      }
      s = s_new;
      g_prev = g_end;
      event_last = false;
    }
    dt = dt_new;
    if (terminated) {
      done = true;
      return;
    }
    if (!AXION && u[0] < P.rNS101) {  // DiscreteCallback, RT.jl:352-359
      status = RAY_HIT_NS;
      done = true;
      return;
    }
    if (!(s < s1)) done = true;
  }

  ART_HD void finish(const DevParams& P, RayOut& o) const {
    state_to_cartesian(P, u, erg, true, o.xf, o.kf);  // RT.jl:393-416
    o.u7f = u7_of(u[5]);
    o.status = status;
    o.n_cross = n_cross;
    o.n_acc = n_acc;
    o.n_rej = n_rej;
    o.n_rhs = n_rhs;
    o.n_cond = n_cond;
  }
};

}  // namespace art
