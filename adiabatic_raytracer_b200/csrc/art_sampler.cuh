// art_sampler.cuh — Monte-Carlo sampler of the conversion surface, one attempt per thread.
//
// Replaces find_samples_new (reference RayTracer.jl:1480-1653): a random straight line through
// a disk of radius maxR is marched with OrdinaryDiffEq's Euler (dt = 0.5 km) under a
// ContinuousCallback (interp_points = 20) that root-finds every sign change of the thick-surface
// condition; one of the kept crossings is selected with probability 1/n_max.
//
// B200 formulation: the attempt is a small state machine whose `advance()` performs exactly ONE
// condition evaluation (scan point, repeat-nudge point or root-finder iterate), so that the 32
// lanes of a warp — each on its own attempt, at its own march position — execute the same
// instruction stream.  Two exact shortcuts relative to the literal march:
//   * the condition is strictly negative wherever max_angles ωp(r) < m_a, i.e. outside the
//     sphere r = Rcut = maxR/1.01 (θ = θm/2 maximises |B_z|, which is how the reference derives
//     maxR, RT.jl:1250-1263): grid steps entirely outside that sphere cannot contain a sign
//     change and are skipped — the evaluation grid inside is unchanged;
//   * cosθ, sinθ, cosψ, sinψ come from the Cartesian point directly (no acos/atan2/sincos).
#pragma once
#include "art_philox.cuh"
#include "art_physics.cuh"

namespace art {

constexpr int SAMP_KEEP = 8;  // crossings remembered per line (n_maxSample ≤ 8)

struct SampleResult {
  double x[3], vloc[3], vinf[3];
  int accepted, n_cross, n_cond;
};

struct Sampler {
  double xs[3], va[3], vl[3];  // line origin, line direction, local velocity direction
  double erg_inf, vI[3], vI_mag;
  double cpsi0, spsi0;         // cos/sin of Ω t0
  double mh[3], mv;            // magnetic axis m̂ at the field time t0, and m̂·v
  double tau, tau_new, h, T, tau_out;
  double g_prev, th_a, g_a, a, fa, b, fb;
  double taus[SAMP_KEEP];
  int i, sgn_prev, side, it, phase;  // phase: 0 scan, 1 nudge, 2 root
  int n_cross, n_cond;
  bool event_last, done;

  ART_HD void point(double t_, double u[3]) const {
    for (int c = 0; c < 3; ++c) u[c] = xs[c] + t_ * va[c];
  }

  // condition of find_samples_new, RT.jl:1554-1581, as a division-light equivalent.
  //
  // The reference builds covariant components of the unit vector v (vvec_loc), normalises them to
  // the axion mass shell and forms Ham = ½(−m_a² + ωp²(1 − g^rr k∥²/E²))/E².  With physical
  // components, k∥²·g^rr/E² = g^rr·NrmSq·N²/(D·E²), NrmSq = (−E² g^tt − m_a²)/ww, and for a point
  // outside the star (AA = g^rr = A, g^tt = −1/A):
  //     ww = v_r² + v_θ² + v_φ² = 1,   N = v·b,   b = 3(m̂·r̂) r̂ − m̂   (dipole direction, |b|² = D),
  //     b_z = 3(m̂·r̂)(ẑ·r̂) − cosθm,   m̂ = (sinθm cosΩt0, sinθm sinΩt0, cosθm)
  // so everything is a dot product of the Cartesian point with two fixed vectors:
  //     F = Q|b_z| r⁻³ (D E² − N² (E² − m_a² A)) − m_a² D E²   has the sign (and the roots) of Ham.
  // One sqrt, one reciprocal, no trigonometry (the literal form costs 14 divisions and 5 square
  // roots per evaluation — profiles/r01_sampler_v1_hotlines.txt: 60 % stall_wait on those chains).
  // Inside the star (r < rNS: AA = 1 quirk RT.jl:1558-1560; r ≤ r_int: interior metric) the same
  // algebra is kept with the general g^tt, g^rr:  ww = 1 − (1 − g^rr/AA') v_r² …, see below.
  ART_HD double cond(const DevParams& P, double t_) {
    n_cond++;
    const double x = xs[0] + t_ * va[0], y = xs[1] + t_ * va[1], z = xs[2] + t_ * va[2];
    const double r2 = x * x + y * y + z * z;
    const double r = sqrt(r2);
    const double rinv = 1.0 / r;
    const double mx = (mh[0] * x + mh[1] * y + mh[2] * z) * rinv;  // m̂·r̂
    const double vx = (vl[0] * x + vl[1] * y + vl[2] * z) * rinv;  // v·r̂  (= v_r)
    const double cz = z * rinv;                                    // ẑ·r̂
    const double D = 3.0 * mx * mx + 1.0;
    const double bz = 3.0 * mx * cz - P.cm;
    const double Pq = P.Q * fabs(bz) * rinv * rinv * rinv;         // ωp²
    const double E2 = erg_inf * erg_inf;
    if (r >= P.rNS && r > P.r_int) {
      const double N = 3.0 * mx * vx - mv;                         // v·b
      const double A = 1.0 - P.rs_real * rinv;
      return Pq * (D * E2 - N * N * (E2 - P.ma2 * A)) - P.ma2 * D * E2;
    }
    // general form: covariant w = (v_r/(√AA·AA), v_θ r/AA, v_φ r s/AA); the common 1/AA drops out
    const Met m = metric(r, P.rs_real, P.r_int);
    const double AA = (r < P.rNS) ? 1.0 : 1.0 - P.rs_real * rinv;
    const double br = 2.0 * mx;
    const double vt2 = 1.0 - vx * vx;                              // v_θ² + v_φ²
    const double wr2 = vx * vx / AA;                               // (v_r/√AA)²
    const double ww = m.grr * wr2 + vt2;
    // N = √g^rr w_r b_r + v_θ b_θ + v_φ b_φ,  v_θ b_θ + v_φ b_φ = v·b − v_r b_r
    const double N = sqrt(m.grr / AA) * vx * br + ((3.0 * mx * vx - mv) - vx * br);
    const double nrm = -E2 * m.gtt - P.ma2;                        // NrmSq·ww
    if (nrm < 0.0) return NAN;
    return Pq * (ww * D * E2 - m.grr * N * N * nrm) - P.ma2 * ww * D * E2;
  }

  // affect!, RT.jl:1585-1597
  ART_HD void affect(const DevParams& P, double t_) {
    double u[3];
    point(t_, u);
    const double rho2 = u[0] * u[0] + u[1] * u[1];
    const double r = sqrt(rho2 + u[2] * u[2]);
    const double rho = sqrt(rho2);
    const double cph = u[0] / rho, sph = u[1] / rho;
    const Fld f = field_cs<false>(P, u[2] / r, rho / r, cph * cpsi0 + sph * spsi0, sph * cpsi0 - cph * spsi0);
    const double wp = sqrt(P.Q / (r * r * r) * fabs(f.bz));
    const Met m = metric(r, P.rs_real, P.r_int);
    if (r > P.rNS && erg_inf / sqrt(m.grr) > wp) {
      if (n_cross < SAMP_KEEP) taus[n_cross] = t_;
      n_cross++;
    }
  }

  ART_HD void next_step(const DevParams& P) {
    if (!(tau < T) || !(tau < tau_out)) {
      done = true;
      return;
    }
    h = P.samp_dt;
    tau_new = tau + h;
    if (tau_new >= T) {
      h = T - tau;
      tau_new = T;
    }
    i = 1;
    if (event_last) {
      phase = 1;
    } else {
      phase = 0;
      sgn_prev = (g_prev > 0.0) - (g_prev < 0.0);
      th_a = 0.0;
      g_a = g_prev;
    }
  }

  // set up one attempt from its 10 uniforms (RT.jl:1486-1542)
  ART_HD void begin(const DevParams& P, const double U[10]) {
    double sti, cti, spi, cpi, stl, ctl, spl, cpl, sR, cR;
    cti = 1.0 - 2.0 * U[0];                 // θi = acos(1 − 2u)
    sti = sqrt(fmax(0.0, 1.0 - cti * cti));
    sincos(U[1] * 2.0 * PI, &spi, &cpi);
    ctl = 1.0 - 2.0 * U[2];
    stl = sqrt(fmax(0.0, 1.0 - ctl * ctl));
    sincos(U[3] * 2.0 * PI, &spl, &cpl);
    sincos(U[4] * 2.0 * PI, &sR, &cR);
    const double rR = sqrt(U[5]) * P.maxR;
    va[0] = sti * cpi; va[1] = sti * spi; va[2] = cti;
    vl[0] = stl * cpl; vl[1] = stl * spl; vl[2] = ctl;
    const double x1 = rR * cR, x2 = rR * sR;
    // Inv[EulerMatrix(ϕi, θi, 0)]·(x1, x2, 0), RT.jl:1529 (cos(−a) = cos a, sin(−a) = −sin a)
    xs[0] = x1 * cpi * cti - x2 * spi;
    xs[1] = x2 * cpi + x1 * spi * cti;
    xs[2] = -x1 * sti;
    double v2 = 0.0;
    for (int c = 0; c < 3; ++c) {
      vI[c] = (220.0 + U[6 + c] * 1.0e-5) / sqrt(3.0);  // RT.jl:1531
      v2 += vI[c] * vI[c];
    }
    vI_mag = sqrt(v2);
    const double bc = vI_mag / C_KM;
    erg_inf = P.ma * sqrt(1.0 + bc * bc / (1.0 - bc * bc));  // RT.jl:1537-1538
    for (int c = 0; c < 3; ++c) xs[c] += va[c] * (-P.maxR * 1.1);  // RT.jl:1542
    sincos(P.Omega * P.t0, &spsi0, &cpsi0);
    mh[0] = P.sm * cpsi0; mh[1] = P.sm * spsi0; mh[2] = P.cm;
    mv = mh[0] * vl[0] + mh[1] * vl[1] + mh[2] * vl[2];
    n_cross = 0;
    n_cond = 0;
    event_last = false;
    done = false;
    T = 2.2 * P.maxR;
    // chord of the line inside the sphere r = Rcut
    const double bq = xs[0] * va[0] + xs[1] * va[1] + xs[2] * va[2];
    const double cq = xs[0] * xs[0] + xs[1] * xs[1] + xs[2] * xs[2] - P.Rcut * P.Rcut;
    const double disc = bq * bq - cq;
    if (!(disc > 0.0)) {
      done = true;
      return;
    }
    const double sd = sqrt(disc);
    const double tau_in = -bq - sd;
    tau_out = -bq + sd;
    const double n0 = floor(fmax(tau_in, 0.0) / P.samp_dt);
    tau = n0 * P.samp_dt;
    g_prev = cond(P, tau);
    next_step(P);
  }

  ART_HD void event(const DevParams& P, double th) {
    const double te = th >= 1.0 ? tau_new : tau + th * h;
    affect(P, te);
    tau = te;
    event_last = true;
    next_step(P);
  }

  // exactly one condition evaluation
  ART_HD void advance(const DevParams& P) {
    if (phase == 1) {  // repeat-nudge reference sign after an event (DiffEqBase repeat_nudge = 1/100)
      const double g = cond(P, tau + 0.01 * h);
      sgn_prev = (g > 0.0) - (g < 0.0);
      th_a = 0.01;
      g_a = g;
      phase = 0;
      return;
    }
    if (phase == 0) {
      const int M = P.M_samp;
      const bool end = (i == M - 1);
      const double th = end ? 1.0 : (double)i * P.inv_Msm1;
      const double g = cond(P, end ? tau_new : tau + th * h);
      const int sg = (g > 0.0) - (g < 0.0);
      if (sgn_prev != 0 && sgn_prev * sg < 0) {
        a = th_a; fa = g_a; b = th; fb = g;
        side = 0; it = 0; phase = 2;
        return;
      }
      if (sg == sgn_prev) {
        th_a = th;
        g_a = g;
      }
      if (end) {
        if (sgn_prev != 0 && g == 0.0) {
          event(P, 1.0);
        } else {
          tau = tau_new;
          g_prev = g;
          event_last = false;
          next_step(P);
        }
      } else {
        i++;
      }
      return;
    }
    // phase 2: Illinois regula falsi with bisection safeguard (same rule as the ray tracer)
    {
      const double tol = 4.0 * 2.220446049250313e-16 * fmax(fabs(a), fabs(b)) + 1e-300;
      if (fabs(b - a) <= tol || it >= 200) {
        event(P, fmin(a, b));
        return;
      }
      double c = (a * fb - b * fa) / (fb - fa);
      const double lo = fmin(a, b), hi = fmax(a, b);
      if (!(c > lo && c < hi) || it >= 60) c = 0.5 * (a + b);
      const double fc = cond(P, tau + c * h);
      it++;
      if (fc == 0.0) {
        event(P, c);
        return;
      }
      if ((fc < 0.0) != (fb < 0.0)) {
        a = b; fa = fb; side = 0;
      } else {
        if (side == 1) fa *= 0.5;
        side = 1;
      }
      b = c; fb = fc;
    }
  }

  // selection of one crossing, RT.jl:1618-1640
  ART_HD void finish(const DevParams& P, double u_idx, SampleResult& R) const {
    R.accepted = 0;
    R.n_cross = n_cross;
    R.n_cond = n_cond;
    for (int c = 0; c < 3; ++c) R.x[c] = R.vloc[c] = R.vinf[c] = 0.0;
    if (n_cross == 0) return;
    int randInx = 1 + (int)floor(u_idx * P.n_max);
    if (randInx > P.n_max) randInx = P.n_max;
    if (n_cross < randInx) return;
    point(taus[randInx - 1], R.x);
    const double rmag = sqrt(R.x[0] * R.x[0] + R.x[1] * R.x[1] + R.x[2] * R.x[2]);
    const double vmag_loc = sqrt(vI_mag * vI_mag + P.rs_real * C_KM * C_KM / rmag) / C_KM;  // 2GM/r
    R.accepted = 1;
    for (int c = 0; c < 3; ++c) {
      R.vloc[c] = vl[c] * vmag_loc;
      R.vinf[c] = vI[c] / C_KM;
    }
  }
};

// whole attempt on one thread (fine-grained API, host check)
ART_HD void sample_attempt(const DevParams& P, const double U[10], SampleResult& R) {
  Sampler S;
  S.begin(P, U);
  while (!S.done) S.advance(P);
  S.finish(P, U[9], R);
}

}  // namespace art
