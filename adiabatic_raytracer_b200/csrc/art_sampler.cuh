// art_sampler.cuh — Monte-Carlo sampler of the conversion surface, one attempt per thread.
//
// Replaces find_samples_new (reference RayTracer.jl:1480-1653): a random straight line through
// a disk of radius maxR is marched with OrdinaryDiffEq's Euler (dt = 0.5 km) under a
// ContinuousCallback (interp_points = 20) that root-finds every sign change of the thick-surface
// condition; one of the kept crossings is selected with probability 1/n_max.
//
// B200 formulation: the attempt is a small state machine advanced in TRIPS of M−1 condition evaluations
// (Sampler::trip): the 32 lanes of a warp — each on its own attempt, at its own march position, scanning a
// step or iterating on a root — evaluate the condition together at one call site.  Two exact shortcuts
// relative to the literal march:
//   * the condition is strictly negative wherever max_angles ωp(r) < m_a, i.e. outside the
//     sphere r = Rcut = maxR/1.01 (θ = θm/2 maximises |B_z|, which is how the reference derives
//     maxR, RT.jl:1250-1263): grid steps entirely outside that sphere cannot contain a sign
//     change and are skipped — the evaluation grid inside is unchanged;
//   * cosθ, sinθ, cosψ, sinψ come from the Cartesian point directly (no acos/atan2/sincos).
#pragma once
#include "art_philox.cuh"
#include "art_physics.cuh"

namespace art {

#ifndef ART_ROOT_RTOL_DEFINED
#define ART_ROOT_RTOL_DEFINED
constexpr double ROOT_RTOL = 1e-13;  // event tolerance of the root finders, relative to the step
#endif

constexpr int SAMP_KEEP = 8;  // crossings remembered per line (n_maxSample ≤ 8)

struct SampleResult {
  double x[3], vloc[3], vinf[3];
  int accepted, n_cross, n_cond;
};

// Attempt state.  Kept small enough to live in registers: the field orientation at t0 (m̂, cos/sin Ωt0)
// and the march length are run constants (DevParams), the crossing list lives outside the struct (device:
// a shared-memory column; a dynamically indexed member would push the whole struct to local memory —
// profiles/r01_sampler_v2_hotlines.txt: 1.2e10 local sectors per launch, 10 active threads per
// instruction).
struct Sampler {
  double xs[3], va[3], vl[3];  // line origin, line direction, local velocity direction
  double E2, vI[3];            // E∞², v∞ components [km/s]
  double mv;                   // m̂·v
  double mva, zva;             // |m̂·â|, |ẑ·â| for the line direction â (dark-step bound)
  double tau, tau_new, h, tau_out;
  double a, fa, b, fb, f0, w1, w2;  // root finder
  double* taus;    // crossing list, SAMP_KEEP entries `taus_stride` apart (device: shared memory column)
  int taus_stride;
  // mode: what the attempt evaluates next — SCAN: the M−1 grid points of one Euler step (a whole trip());
  // ROOT: one root-finder iterate; NUDGE: the reference sign 1 % into the step after an event; INIT: the sign at
  // the first point of the march
  // SKIP: the attempt is passing over certified steps and has not come to one it must scan yet (skip_certified_steps
  // does a bounded number of tests per trip; the lane idles through the evaluation slots of the trip)
  enum { SCAN = 0, ROOT = 1, NUDGE = 2, INIT = 3, SKIP = 4 };
  int sgn_prev, side, it, mode, rinit;
  int sblk;  // skip_certified_steps: current block size (16, 4 or 1 steps) and, in bits 8.., successes in a row
  int n_cross, n_cond;
  bool event_last, done, pure;

  ART_HD void point(double t_, double u[3]) const {
    for (int c = 0; c < 3; ++c) u[c] = xs[c] + t_ * va[c];
  }

  // condition of find_samples_new, RT.jl:1554-1581, as a division-light equivalent.
  //
  // The reference builds covariant components of the unit vector v (vvec_loc), normalises them to
  // the axion mass shell and forms Ham = ½(−m_a² + ωp²(1 − g^rr k∥²/E²))/E².  With physical
  // components, k∥²·g^rr/E² = g^rr·NrmSq·N²/(D·E²), NrmSq = (−E² g^tt − m_a²)/ww, and
  //     b = 3(m̂·r̂) r̂ − m̂   (dipole direction, |b|² = D),   b_z = 3(m̂·r̂)(ẑ·r̂) − cosθm,
  //     covariant w = (v_r/(√AA·AA), v_θ r/AA, v_φ r s/AA)  (the common 1/AA drops out)
  //     ww = g^rr v_r²/AA + v_θ² + v_φ²,   N = √(g^rr/AA) v_r b_r + (v·b − v_r b_r)
  //     F = Q|b_z| r⁻³ (ww D E² − g^rr NrmSq·ww N²) − m_a² ww D E²   has the sign (and the roots) of Ham.
  // Outside the star AA = g^rr = A and g^tt = −1/A, so ww = 1, N = v·b, g^rr NrmSq·ww = E² − m_a² A:
  // one rsqrt, no division, no trigonometry.  Inside (r < rNS: AA = 1 quirk RT.jl:1558-1560; r ≤ r_int:
  // interior metric) only the three metric scalars differ; the formula is shared, so lanes inside and
  // outside the star stay converged except for the few instructions of the interior metric.
  ART_HD double cond(const DevParams& P, double t_) {
    n_cond++;
    const double x = xs[0] + t_ * va[0], y = xs[1] + t_ * va[1], z = xs[2] + t_ * va[2];
    const double r2 = x * x + y * y + z * z;
    const double rinv = frsqrt(r2);
    const double r = r2 * rinv;
    const double mx = (P.mh[0] * x + P.mh[1] * y + P.mh[2] * z) * rinv;  // m̂·r̂
    const double vx = (vl[0] * x + vl[1] * y + vl[2] * z) * rinv;        // v·r̂  (= v_r)
    const double cz = z * rinv;                                          // ẑ·r̂
    const double D = 3.0 * mx * mx + 1.0;
    const double bz = 3.0 * mx * cz - P.cm;
    double Pq = P.Q * fabs(bz) * rinv * rinv * rinv;                     // ωp²
    if (P.bl > 0.0 && r >= P.rNS) Pq = bl_wp2(P, Pq, r, P.bl_rmax);     // boundary layer (GJ_Model_ωp_vec, RT.jl:1092-1098)
    // thick_surface = false (RT.jl:1548-1551): log ωp − log m_a, i.e. the sign and zeros of ωp² − m_a²
    // (GJ_Model_ωp_vec: no zeroIn, the plasma frequency continues inside the star)
    if (!P.thick) return Pq - P.ma2;
    const double vb = 3.0 * mx * vx - mv;                                // v·b
    double ww = 1.0, N = vb, gn = E2 - P.ma2 * (1.0 - P.rs_real * rinv);
    if (!(r >= P.rNS && r > P.r_int)) {
      double grr, sq, gtt;  // g^rr, √g^rr, g^tt
      if (r > P.r_int) {
        grr = 1.0 - P.rs_real * rinv;
        const double y = frsqrt(grr);
        sq = grr * y;
        gtt = -(y * y);
      } else {  // g_schwartz interior branch (RT.jl:463, 496-497), see metric()
        const double q = r * P.inv_r_int;
        const double q2 = q * q;
        const double aa = P.rs_real * P.inv_r_int * q2 * q;
        const double bb = aa * q2;
        grr = 1.0 - bb;
        sq = fsqrt(grr);
        const double den = 3.0 * fsqrt(1.0 - aa) - sq;
        gtt = -4.0 * frcp(den * den);
      }
      // AA = 1 inside the star (RT.jl:1558-1560 quirk), 1 − rs/r outside
      const bool in = r < P.rNS;
      const double A = 1.0 - P.rs_real * rinv;
      const double yA = in ? 1.0 : frsqrt(A);  // 1/√AA
      const double br = 2.0 * mx;
      ww = grr * vx * vx * yA * yA + (1.0 - vx * vx);
      N = sq * yA * vx * br + (vb - vx * br);
      const double nrm = -E2 * gtt - P.ma2;  // NrmSq·ww
      gn = grr * nrm;
      if (nrm < 0.0) gn = NAN;
    }
    // NrmSq < 0 (E below the local mass shell): the reference would throw; NaN = "no crossing"
    if (gn < 0.0) gn = NAN;
    const double base = ww * D * E2;
    if (P.isotropic) N = 0.0;  // `if iso kpar = 0.0`, RT.jl:1573-1574 (gn stays in the product: its NaN must survive)
    return Pq * (base - gn * N * N) - P.ma2 * base;
  }

  // affect!, RT.jl:1585-1597: keep the crossing iff r > rNS and E/√g^rr > ωp, i.e. E² > g^rr ωp² —
  // evaluated like cond(): dot products with m̂, one rsqrt, no square root of the two sides
  ART_HD void affect(const DevParams& P, double t_) {
    double u[3];
    point(t_, u);
    const double r2 = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
    const double rinv = frsqrt(r2);
    const double r = r2 * rinv;
    const double mx = (P.mh[0] * u[0] + P.mh[1] * u[1] + P.mh[2] * u[2]) * rinv;
    const double bz = 3.0 * mx * (u[2] * rinv) - P.cm;
    double wp2 = P.Q * fabs(bz) * rinv * rinv * rinv;
    if (P.bl > 0.0 && r >= P.rNS) wp2 = bl_wp2(P, wp2, r, P.bl_rmax);
    double grr = 1.0 - P.rs_real * rinv;
    if (!(r > P.r_int)) grr = metric(r, P.rs_real, P.r_int).grr;
    if (r > P.rNS && E2 > grr * wp2) {
      if (n_cross < SAMP_KEEP) taus[n_cross * taus_stride] = t_;
      n_cross++;
    }
  }

  // start of the next Euler step from tau (OrdinaryDiffEq Euler, dt = 0.5 km, last step clipped to the span)
  ART_HD void next_step(const DevParams& P) {
    if (!(tau < P.T_samp) || !(tau < tau_out)) {
      done = true;
      return;
    }
    h = P.samp_dt;
    tau_new = tau + h;
    if (tau_new >= P.T_samp) {
      h = P.T_samp - tau;
      tau_new = P.T_samp;
    }
    mode = event_last ? NUDGE : SCAN;
  }

  // Certified sign of the condition on the WHOLE segment [t0, t0 + L] of the line, when the momentum is not
  // needed to know it (cond(): F = base·(ωp² − m_a²) − ωp²·g^rr NrmSq·ww N², 0 ≤ N² ≤ ww D):
  //   ωp² < m_a² everywhere on the segment          ⇒ F < 0  ("dark"), return −1
  //   ωp² (1 − rs/r) > E² everywhere, outside the star ⇒ F > 0  ("lit"),  return +1  (inside the star: see below)
  // Rigorous interval bound: on the segment |x − x_mid| ≤ L/2, so r ∈ [r_mid − L/2, r_mid + L/2] and, for a
  // unit vector n, n·x ∈ [n·x_mid ± (L/2)|n·â|]; hence intervals for m̂·r̂ and ẑ·r̂, for their product, and for
  // b_z = 3(m̂·r̂)(ẑ·r̂) − cosθm.  A direction-aware refinement of the Rcut sphere: only the steps that straddle
  // the conversion surface (and those inside the star) are left to evaluate.
  ART_HD int seg_sign(const DevParams& P, double t0, double L) const {
    const double hm = 0.5 * L, tm = t0 + hm;
    const double x = xs[0] + tm * va[0], y = xs[1] + tm * va[1], z = xs[2] + tm * va[2];
    const double r2 = x * x + y * y + z * z;
    const double rm = r2 * frsqrt(r2);
    const double rlo = rm - hm;
    if (!(rlo > 0.25 * rm)) return 0;
    const double ilo = frcp(rlo), ihi = frcp(rm + hm);  // 1/r ∈ [ihi, ilo]
    const double am = P.mh[0] * x + P.mh[1] * y + P.mh[2] * z, dm = hm * mva, dz = hm * zva;
    const double m0 = am - dm, m1 = am + dm, z0 = z - dz, z1 = z + dz;
    const double mlo = m0 * (m0 >= 0.0 ? ihi : ilo), mhi = m1 * (m1 >= 0.0 ? ilo : ihi);
    const double zlo = z0 * (z0 >= 0.0 ? ihi : ilo), zhi = z1 * (z1 >= 0.0 ? ilo : ihi);
    const double p1 = mlo * zlo, p2 = mlo * zhi, p3 = mhi * zlo, p4 = mhi * zhi;
    const double plo = fmin(fmin(p1, p2), fmin(p3, p4)), phi = fmax(fmax(p1, p2), fmax(p3, p4));
    const double b0 = 3.0 * plo - P.cm, b1 = 3.0 * phi - P.cm;  // b_z ∈ [b0, b1]
    const double bzmax = fmax(fabs(b0), fabs(b1));
    if (P.bl > 0.0) {
      // boundary layer (cond(): for r ≥ rNS ωp = √(Q|b_z|/r³) + B(r), B ≥ 0 falling with r): ωp ≤ √(Q b_z,max/r_lo³) + B(r_lo)
      // on the whole segment (points inside the star carry no B: smaller still); the lit bounds below stay sufficient as
      // they are — the term only raises ωp, and F grows with ωp²
      const double wmax = sqrt(P.Q * bzmax * ilo * ilo * ilo) + bl_w(P, rlo, P.bl_rmax);
      if (wmax * wmax < P.ma2 * (1.0 - 1e-9)) return -1;
    } else if (P.Q * bzmax * ilo * ilo * ilo < P.ma2 * (1.0 - 1e-9)) {
      return -1;
    }
    const double bzmin = (b0 > 0.0) ? b0 : ((b1 < 0.0) ? -b1 : 0.0);
    if (!P.thick) return (P.Q * bzmin * ihi * ihi * ihi > P.ma2 * (1.0 + 1e-9)) ? 1 : 0;  // ωp > m_a on the whole segment
    if (P.isotropic) return 0;
    if (rlo > P.rNS && rlo > P.r_int) return (P.Q * bzmin * ihi * ihi * ihi * (1.0 - P.rs_real * ilo) > E2 * (1.0 + 1e-9)) ? 1 : 0;
    // wholly inside the star (a line through it spends up to 40 steps there): the plasma frequency continues inwards
    // (no zeroIn in the sampler) and dwarfs m_a except near b_z = 0.  cond(): N² ≤ ww D (Cauchy–Schwarz on the covariant
    // components, any metric branch), so F ≥ ww D [ωp² (E² − g^rr NrmSq·ww) − m_a² E²]; in the interior metric g^rr ≤ 1 and
    // −g^tt grows with r up to 1/(1 − rs/r_int) at the switch radius (metric(): den falls with r), hence
    // g^rr NrmSq·ww ≤ E²/(1 − rs/r_int) − m_a², and NrmSq > 0 because −g^tt ≥ 1.
    const double rhi = rm + hm;
    if (rhi < P.rNS && rhi < P.r_int) {
      const double gn_max = E2 * frcp(1.0 - P.rs_real * P.inv_r_int) - P.ma2;
      const double room = E2 - gn_max;
      return (room > 0.0 && P.Q * bzmin * ihi * ihi * ihi * room > P.ma2 * E2 * (1.0 + 1e-9)) ? 1 : 0;
    }
    return 0;
  }

  // Steps whose sign is certified and agrees with the reference sign hold no event: all their grid points
  // have that sign.  They are passed over in blocks of 16, 4 or 1 steps, with the reference's own accumulation of tau:
  // the block shrinks on failure and grows back after four successes in a row; a failed single step ends the skipping
  // (that step is scanned).  At most SKIP_PASSES tests per call — the lanes of a warp run this together at the end of a
  // trip, and a lane in a long dark stretch must not keep the others waiting: it stays in mode SKIP, sits out the next
  // trip's evaluation slots and goes on skipping at its end.  Called where a step ends without an event.
  // (Tuning history, profiles/r02_*: six fixed passes 16-16-4-4-1-1 scanned every third or eleventh step of a stretch
  // whose long blocks could not be certified — a third of all scan trips; an unbounded greedy loop evaluated 44 % fewer
  // conditions but ran slower, its tail of one or two lanes holding up the warp.)
#ifndef ART_SKIP_PASSES
#define ART_SKIP_PASSES 8
#endif
  static constexpr int SKIP_PASSES = ART_SKIP_PASSES;
  ART_HD void skip_certified_steps(const DevParams& P) {
#if defined(ART_NO_DARK_SKIP)
    return;
#endif
    if (done || mode != SKIP) return;
    const double lim = fmin(P.T_samp, tau_out);
    int blk = sblk & 0xff, streak = sblk >> 8;
    bool stopped = false;  // a single step could not be certified
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int pass = 0; pass < SKIP_PASSES; ++pass) {
      const double L = (double)blk * P.samp_dt;
      int sg = 0;
      if (tau + L < lim) sg = seg_sign(P, tau, L);
      if (sg != 0 && sgn_prev * sg >= 0) {
        // (the reference's own accumulation: blk additions do not round like one multiply-add — 5e-4 of random cases differ
        // in the last bit even for dt = 0.5; straight-line per block size: a loop of blk trips, or sixteen predicated
        // additions, ran for two or three lanes at a time and made up 4–11 % of the kernel's instructions)
        const double dt_ = P.samp_dt;
        tau += dt_;
        if (blk >= 4) {
          tau += dt_; tau += dt_; tau += dt_;
        }
        if (blk == 16) {
          tau += dt_; tau += dt_; tau += dt_; tau += dt_; tau += dt_; tau += dt_;
          tau += dt_; tau += dt_; tau += dt_; tau += dt_; tau += dt_; tau += dt_;
        }
        sgn_prev = sg;
        if (++streak >= 4 && blk < 16) {
          blk <<= 2;
          streak = 0;
        }
      } else if (blk > 1) {
        blk >>= 2;
        streak = 0;
      } else {
        stopped = true;
        break;
      }
    }
    next_step(P);  // (event_last is false here: mode = SCAN unless the march is over)
    if (stopped) {
      sblk = 4;  // the next skipping phase starts with a modest block
    } else {
      sblk = blk | (streak << 8);
      if (!done) mode = SKIP;
    }
  }

  // set up one attempt from its 10 uniforms (RT.jl:1486-1542)
  ART_HD void begin(const DevParams& P, const double U[10]) {
    double sti, cti, spi, cpi, stl, ctl, spl, cpl, sR, cR;
    cti = 1.0 - 2.0 * U[0];                 // θi = acos(1 − 2u)
    sti = sqrt(fmax(0.0, 1.0 - cti * cti));
    fsincos(U[1] * 2.0 * PI, &spi, &cpi);
    ctl = 1.0 - 2.0 * U[2];
    stl = sqrt(fmax(0.0, 1.0 - ctl * ctl));
    fsincos(U[3] * 2.0 * PI, &spl, &cpl);
    fsincos(U[4] * 2.0 * PI, &sR, &cR);
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
    const double bc = sqrt(v2) / C_KM;
    const double erg_inf = P.ma * sqrt(1.0 + bc * bc / (1.0 - bc * bc));  // RT.jl:1537-1538
    E2 = erg_inf * erg_inf;
    for (int c = 0; c < 3; ++c) xs[c] += va[c] * (-P.maxR * 1.1);  // RT.jl:1542
    mv = P.mh[0] * vl[0] + P.mh[1] * vl[1] + P.mh[2] * vl[2];
    mva = fabs(P.mh[0] * va[0] + P.mh[1] * va[1] + P.mh[2] * va[2]);
    zva = fabs(va[2]);
    n_cross = 0;
    n_cond = 0;
    event_last = false;
    done = false;
    pure = false;
    mode = INIT;
    sgn_prev = 0; side = 0; it = 0; rinit = 0;
    sblk = 16;
    tau_new = h = a = fa = b = fb = f0 = w1 = w2 = 0.0;
    // chord of the line inside the sphere r = Rcut
    const double bq = xs[0] * va[0] + xs[1] * va[1] + xs[2] * va[2];
    const double cq = xs[0] * xs[0] + xs[1] * xs[1] + xs[2] * xs[2] - P.Rcut * P.Rcut;
    const double disc = bq * bq - cq;
    tau = 0.0;
    tau_out = 0.0;
    if (!(disc > 0.0)) {
      done = true;
      return;
    }
    const double sd = sqrt(disc);
    const double tau_in = -bq - sd;
    tau_out = -bq + sd;
    const double n0 = floor(fmax(tau_in, 0.0) / P.samp_dt);
    tau = n0 * P.samp_dt;  // the first trip evaluates the condition here (mode INIT)
  }

  ART_HD void event(const DevParams& P, double th) {
    const double te = th >= 1.0 ? tau_new : tau + th * h;
    affect(P, te);
    tau = te;
    event_last = true;
    next_step(P);
  }

  // ---- the attempt as three kinds of work -------------------------------------------------------------------------------
  // An attempt alternates between
  //   SCAN   the M−1 grid points of one Euler step (scan_trip): nothing is kept but the signs (three bit masks); the first
  //          sign change — with the reference's rule for the left end of the bracket: the last point that still had the old
  //          sign — starts the root finder, otherwise the step is done and the attempt goes on to SKIP;
  //   ROOT / NUDGE / INIT   one condition evaluation at a time (point_slot): the root finder's iterates (its two bracket
  //          values are re-evaluated first — same numbers as the scan saw), the reference sign 1 % into the step after an
  //          event, the sign at the first point of the march;
  //   SKIP   certified steps are passed over without evaluating anything (skip_certified_steps).
  // The sampler kernel runs each kind for the lanes of a warp that are in it TOGETHER and only when enough of them are
  // (k_sample_events): as one loop over "slots" in which every lane did whatever its mode asked for, the root finder's
  // iterate ran for 2–3 lanes, the interior-metric branch for 6, the skip tests for 4–9, and the kernel averaged 10 of 32
  // lanes per instruction (profiles/r02_ncu_summary.txt).  The host check calls them one after the other (trip()).
  ART_HD bool in_point_mode() const { return mode == ROOT || mode == NUDGE || mode == INIT; }

  // the step has ended without an event: reference sign for the next one, then on to the certified-step tests
  ART_HD void step_done(const DevParams& P, int sg_end) {
    sgn_prev = sg_end;
    tau = tau_new;
    event_last = false;
    next_step(P);
    if (!done) mode = SKIP;
  }

  ART_HD void scan_trip(const DevParams& P) {
    const int M = P.M_samp;
    unsigned mpos = 0u, mneg = 0u, mzero = 0u;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int k = 1; k <= M - 1; ++k) {
      const double tq = (k == M - 1) ? tau_new : tau + ((double)k * P.inv_Msm1) * h;
      const double g = cond(P, tq);
      const unsigned bit = 1u << k;
      mpos |= (g > 0.0) ? bit : 0u;
      mneg |= (g < 0.0) ? bit : 0u;
      mzero |= (g == 0.0) ? bit : 0u;
    }
    // first point whose sign is opposite to the reference sign; the left end of the bracket is the last
    // point before it that still had the reference sign (exact zeros and NaNs neither trigger nor move it)
    const unsigned opp = sgn_prev > 0 ? mneg : (sgn_prev < 0 ? mpos : 0u);
    if (opp) {
#if defined(__CUDA_ARCH__)
      const int ib = __ffs((int)opp) - 1;
#else
      const int ib = __builtin_ffs((int)opp) - 1;
#endif
      const unsigned same = (sgn_prev > 0 ? mpos : mneg) & ((1u << ib) - 1u);
      int ia = 0;
      if (same) {
#if defined(__CUDA_ARCH__)
        ia = 31 - __clz((int)same);
#else
        ia = 31 - __builtin_clz(same);
#endif
      }
      a = (ia == 0) ? (event_last ? 0.01 : 0.0) : (double)ia * P.inv_Msm1;
      b = (ib == M - 1) ? 1.0 : (double)ib * P.inv_Msm1;
      w1 = w2 = 2.0;
      pure = false;
      side = 0;
      it = 0;
      rinit = 2;
      mode = ROOT;
    } else if (sgn_prev != 0 && ((mzero >> (M - 1)) & 1u)) {
      event(P, 1.0);  // the end point lies exactly on the surface
    } else {
      const unsigned endbit = 1u << (M - 1);
      step_done(P, (mpos & endbit) ? 1 : ((mneg & endbit) ? -1 : 0));
    }
  }

  // one evaluation of an attempt in ROOT, NUDGE or INIT mode
  ART_HD void point_slot(const DevParams& P) {
    double tq = tau, th = 0.0;
    bool ev = true;
    if (mode == ROOT) {
      if (rinit > 0) {  // bracket values first: fa, then fb
        th = (rinit == 2) ? a : b;
      } else {
        const double tol = ROOT_RTOL * fmax(fmax(fabs(a), fabs(b)), 0.01);
        const double width = fabs(b - a);
        if (width <= tol || it >= 200) {
          event(P, fmin(a, b));
          ev = false;
        } else {
          if (it == 3 && fmin(fabs(fa), fabs(fb)) > 0.25 * f0) pure = true;
          double c = (a * fb - b * fa) * frcp(fb - fa);
          const double lo = fmin(a, b), hi = fmax(a, b);
          if (pure || (it >= 3 && width > 0.5 * w2) || !(c > lo && c < hi) || width <= 2.0 * tol) c = 0.5 * (a + b);
          else if (fabs(c - b) < tol) c = b + copysign(tol, a - b);
          w2 = w1;
          w1 = width;
          th = c;
        }
      }
      tq = tau + th * h;
      if (!ev) return;  // (the root was accepted: event() has moved the attempt on)
    } else if (mode == NUDGE) {  // DiffEqBase repeat_nudge = 1/100
      tq = tau + 0.01 * h;
    }
    const double g = cond(P, tq);
    if (mode == ROOT) {
      if (rinit == 2) {
        fa = g;
        rinit = 1;
      } else if (rinit == 1) {
        fb = g;
        rinit = 0;
        f0 = fmin(fabs(fa), fabs(fb));
      } else {
        it++;
        if (g == 0.0) {
          event(P, th);
        } else {
          if ((g < 0.0) != (fb < 0.0)) {
            a = b; fa = fb; side = 0;
          } else {
            if (side == 1) fa *= 0.5;
            side = 1;
          }
          b = th; fb = g;
        }
      }
    } else if (mode == NUDGE) {
      sgn_prev = (g > 0.0) - (g < 0.0);
      mode = SCAN;
    } else {  // INIT
      sgn_prev = (g > 0.0) - (g < 0.0);
      next_step(P);
      if (!done) mode = SKIP;
    }
  }

  // sequential form (host check, fine-grained API): one call = one piece of work of whatever kind the attempt is in
  ART_HD void trip(const DevParams& P, bool active) {
    if (!active || done) return;
    if (mode == SCAN) {
      scan_trip(P);
    } else if (mode == SKIP) {
      skip_certified_steps(P);
    } else {
      for (int k = 1; k <= P.M_samp - 1 && !done && in_point_mode(); ++k) point_slot(P);
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
    point(taus[(randInx - 1) * taus_stride], R.x);
    const double rmag = sqrt(R.x[0] * R.x[0] + R.x[1] * R.x[1] + R.x[2] * R.x[2]);
    const double vI_mag = sqrt(vI[0] * vI[0] + vI[1] * vI[1] + vI[2] * vI[2]);
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
  double taus[SAMP_KEEP];
  S.taus = taus;
  S.taus_stride = 1;
  S.begin(P, U);
  while (!S.done) S.trip(P, true);
  S.finish(P, U[9], R);
}

}  // namespace art
