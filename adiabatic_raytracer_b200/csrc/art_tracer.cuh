// art_tracer.cuh — one-ray-per-thread adaptive integrator with level-crossing events.
//
// Replaces `propagate` (reference RayTracer.jl:171-452): OrdinaryDiffEq's adaptive solve in
// ln t with a ContinuousCallback (thick-surface crossings, root-found) and a DiscreteCallback
// (photon hits the star).  Integrator specification: Dormand–Prince 5(4), FSAL,
// Hairer's continuous extension, OrdinaryDiffEq's error norm / initial step / PI controller,
// DiffEqBase's event scan at `interp_points` dense-output points per accepted step.
//
// Layout (B200): the integrated state u = [r θ φ k_r/E k_θ/E k_φ/E] and the controller state
// live in registers; the seven stage derivatives k1..k7 live in shared memory (one column per
// thread, conflict-free) so that the stage loop can be ROLLED: the kernel contains exactly one
// copy of the right-hand side and one copy of the crossing condition, and the hot loop fits the
// instruction cache (the first, fully unrolled version was 28-35k SASS instructions and spent
// 30-43 % of its issue slots in stall_no_inst — profiles/r01_trace_v1_*.txt).
// The 7th component of the reference's state (u7 = −ω) is carried by the exact linear invariant
// u7 = u7₀ + (Ω·erg/c_km)(u6 − u6₀) of the photon equations (slope 0 for axions), which every
// Runge–Kutta method and its dense output preserve; it still enters the error norm.
#pragma once
#include "art_physics.cuh"

namespace art {

enum { RAY_END = 0, RAY_HIT_NS = 1, RAY_CROSSING = 2, RAY_MAXITERS = 3, RAY_CROSS_OVERFLOW = 4 };
enum { SPECIES_AXION = 0, SPECIES_PHOTON = 1 };

// Dormand–Prince 5(4) (Hairer, Nørsett, Wanner, Solving ODEs I, Table II.5.2; dense output II.6)
// as flat tables for the rolled stage loop: row st (1..6) holds a_{st+1, 1..st}; row 6 = b.
#if defined(__CUDACC__)
#define ART_TABLE __device__ __constant__
#else
#define ART_TABLE static const
#endif
ART_TABLE double DP_A[7][6] = {
    {0, 0, 0, 0, 0, 0},
    {1.0 / 5, 0, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656, 0},
    {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84}};
ART_TABLE double DP_C[7] = {0, 1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1.0, 1.0};
ART_TABLE double DP_E[7] = {71.0 / 57600, 0, -71.0 / 16695, 71.0 / 1920, -17253.0 / 339200, 22.0 / 525, -1.0 / 40};
ART_TABLE double DP_D[7] = {-12715105075.0 / 11282082432.0, 0, 87487479700.0 / 32700410799.0,
                            -10690763975.0 / 1880347072.0, 701980252875.0 / 199316789632.0,
                            -1453857185.0 / 822651844.0, 69997945.0 / 29380423.0};

namespace dp {  // OrdinaryDiffEq PIController constants for DP5
constexpr double beta1 = 0.17, beta2 = 0.04, gamma = 0.9, qmin = 0.2, qmax = 10.0;
constexpr double qsteady_min = 1.0, qsteady_max = 1.2, qoldinit = 1e-4;
}  // namespace dp

struct RayIn {
  double x0[3], k0[3];  // Cartesian start position [km] and momentum [eV]
  double erg;           // erg_inf_ini
  double t_start;       // node.t [s]
  double dOmega;        // node.Δω
  int max_cross;        // get_tree's splittings_cutoff as passed to propagate
  int species;          // SPECIES_*
};

struct RayOut {
  double xf[3], kf[3], u7f;
  int status, n_cross;
  int n_acc, n_rej, n_rhs, n_cond;
};

// Stage-derivative store: K(j, i) = component i of k_{j+1}.  Device: shared memory, one column
// per thread (stride = block size ⇒ bank-conflict free); host check: a plain local array.
struct KStore {
  double* base;
  int stride;
  ART_HD double& operator()(int j, int i) const { return base[(j * 6 + i) * stride]; }
  static constexpr bool PARKS = false;  // see KStoreT
  ART_HD double& park(int) const { return base[0]; }
};
// same with the stride known at compile time (device: STRIDE = CTA size ⇒ addresses fold into immediates)
// Rows 42.. of the device store hold t = e^{ln t} of the six stage times of the attempted step.
constexpr int K_ROWS = 42, K_PARK = 6;
template <int STRIDE>
struct KStoreT {
  double* base;
  ART_HD double& operator()(int j, int i) const { return base[(j * 6 + i) * STRIDE]; }
#ifndef ART_PRE_T
#define ART_PRE_T 1
#endif
  static constexpr bool PARKS = ART_PRE_T != 0;
  ART_HD double& park(int k) const { return base[(K_ROWS + k) * STRIDE]; }
};

// Dense-output segment of an accepted step, published (device: in shared memory, one column
// per thread) so that ANY lane of the warp can evaluate the crossing condition on it.
enum { SEG_U = 0, SEG_R1 = 6, SEG_R2 = 12, SEG_R3 = 18, SEG_R4 = 24, SEG_S = 30, SEG_H = 31, SEG_U70 = 32,
       SEG_FAC7 = 33, SEG_U60 = 34, SEG_N = 35 };
struct SegStore {
  double* base;
  int stride;
  ART_HD double& operator()(int f) const { return base[f * stride]; }
};
struct SegAoS {  // contiguous record (stride 1)
  double* base;
  ART_HD double& operator()(int f) const { return base[f]; }
};
ART_HD int ffs64(unsigned long long v) {  // 1-based index of the lowest set bit
#if defined(__CUDA_ARCH__)
  return __ffsll((long long)v);
#else
  return __builtin_ffsll((long long)v);
#endif
}
ART_HD int clz64(unsigned long long v) {
#if defined(__CUDA_ARCH__)
  return __clzll((long long)v);
#else
  return __builtin_clzll(v);
#endif
}

#ifndef ART_ROOT_RTOL_DEFINED
#define ART_ROOT_RTOL_DEFINED
constexpr double ROOT_RTOL = 1e-13;  // event tolerance of the root finders, relative to the step
#endif

struct ScanMasks {
  unsigned long long pos, neg, zero;  // bit i: sign of the condition at grid point i (bit 0: nudge point)
};

// dense-output grid of the crossing scan: θ_i = i/(M−1), i = 1..M−1; index 0 = the repeat-nudge point 0.01
ART_HD double seg_theta(const DevParams& P, int i) { return i == 0 ? 0.01 : ((i == P.M_ray - 1) ? 1.0 : (double)i * P.inv_Mrm1); }

template <class DS>
ART_HD double seg_comp(const DS D, int i, double th, double th1) {
  return D(SEG_U + i) + th * (D(SEG_R1 + i) + th1 * (D(SEG_R2 + i) + th * (D(SEG_R3 + i) + th1 * D(SEG_R4 + i))));
}
template <class DS>
ART_HD void seg_state(const DS D, double th, double w[6]) {
  const double th1 = 1.0 - th;
  for (int i = 0; i < 6; ++i) w[i] = seg_comp(D, i, th, th1);
}

// Far-field classification of a dense-output point: −1 the condition is provably negative
// (r > Rcut ⇒ ωp < m_a in every direction, and NrmSq ≥ 0 ⇔ E² r ≥ m_a² (r − rs)), 2 it is NaN
// (NrmSq < 0, only reachable in flat mode), 0 it has to be evaluated.
template <class DS>
ART_HD int seg_quick(const DevParams& P, const DS D, double th) {
  const double th1 = 1.0 - th;
  const double r = seg_comp(D, 0, th, th1);
  if (!(r > P.Rcut && r > P.r_int)) return 0;
  const double E = D(SEG_U70) + D(SEG_FAC7) * (seg_comp(D, 5, th, th1) - D(SEG_U60));
  return (E * E * r >= P.ma2 * (r - P.rs_ray)) ? -1 : 2;
}

// The same test with the two polynomials it needs (r and u6) held in registers — used where all
// M−1 points of a segment are classified in a row.
struct SegQuick {
  double a0[5], a5[5], u70, fac7, u60;
  template <class DS>
  ART_HD explicit SegQuick(const DS D) {
    a0[0] = D(SEG_U + 0); a0[1] = D(SEG_R1 + 0); a0[2] = D(SEG_R2 + 0); a0[3] = D(SEG_R3 + 0); a0[4] = D(SEG_R4 + 0);
    a5[0] = D(SEG_U + 5); a5[1] = D(SEG_R1 + 5); a5[2] = D(SEG_R2 + 5); a5[3] = D(SEG_R3 + 5); a5[4] = D(SEG_R4 + 5);
    u70 = D(SEG_U70); fac7 = D(SEG_FAC7); u60 = D(SEG_U60);
  }
  // Rigorous whole-segment test: θ(1−θ) ≤ ¼ on [0,1] bounds the deviation of the dense-output polynomial
  // from its chord, so min r over the step ≥ min(r₀, r₁) − ¼(|r2| + |r3| + ¼|r4|); if that clears Rcut
  // (with a 1e-9 margin against Horner round-off) and E² ≥ m_a² over the step, every scan point is
  // in the provably negative far field and the per-point classification is skipped.
  ART_HD bool all_far(const DevParams& P) const {
    const double lo = fmin(a0[0], a0[0] + a0[1]) - 0.25 * (fabs(a0[2]) + fabs(a0[3]) + 0.25 * fabs(a0[4]));
    if (!(lo > fmax(P.Rcut, P.r_int) * (1.0 + 1e-9))) return false;
    const double e0 = u70 + fac7 * (a5[0] - u60), e1 = e0 + fac7 * a5[1];
    const double de = fabs(fac7) * 0.25 * (fabs(a5[2]) + fabs(a5[3]) + 0.25 * fabs(a5[4]));
    const double emin = fmin(fabs(e0), fabs(e1)) - de;
    return emin > 0.0 && emin * emin >= P.ma2 * (1.0 + 1e-12);
  }
  ART_HD int classify(const DevParams& P, double th) const {
    const double th1 = 1.0 - th;
    const double r = a0[0] + th * (a0[1] + th1 * (a0[2] + th * (a0[3] + th1 * a0[4])));
    if (!(r > P.Rcut && r > P.r_int)) return 0;
    const double u6 = a5[0] + th * (a5[1] + th1 * (a5[2] + th * (a5[3] + th1 * a5[4])));
    const double E = u70 + fac7 * (u6 - u60);
    return (E * E * r >= P.ma2 * (r - P.rs_ray)) ? -1 : 2;
  }
};

// `condition` of propagate (RT.jl:254-298) at dense-output parameter th of a published segment, as the
// sign-equivalent division-free thick_F (art_physics.cuh); the scan keeps only its sign, the root finder
// its zero
template <bool SIGN_ONLY = false, class DS>
ART_HD double seg_cond_full(const DevParams& P, const DS D, double th) {
  double w[6];
  seg_state(D, th, w);
  const double E = D(SEG_U70) + D(SEG_FAC7) * (w[5] - D(SEG_U60));
  const double t = fexp(D(SEG_S) + th * D(SEG_H));
  return thick_F<SIGN_ONLY>(P, w[0], w[1], w[2] - P.Omega * t, w[3], w[4], w[5], E, P.rs_ray, true);
}
template <class DS>
ART_HD double seg_cond(const DevParams& P, const DS D, double th) {
  const int qk = seg_quick(P, D, th);
  if (qk == -1) return -1.0;
  if (qk == 2) return NAN;
  return seg_cond_full<true>(P, D, th);  // the scan keeps the sign only
}

// Ray state, split by access pattern: the HOT record is what an attempted step reads
// and writes (one 192-byte, 16-byte-aligned record per ray: 12 vector loads); the COLD record is only
// touched by callbacks (start-box test, crossing records, root brackets).
struct alignas(16) RayHot {
  double u[6];    // state at the start of the next step
  double k1[6];   // f(u, s): FSAL derivative, always valid between steps
  double s, dt, qold;
  double erg, fac7, u7_0, u6_0;
  int iters, n_cross, status;
  int sgn_end;    // sign of the condition at the end of the last accepted step (= start of the next)
  int n_acc, n_rej, n_rhs, n_cond;
  bool axion, starting, event_last, done;
  int left_box;   // the ray has (possibly) left the ±0.01 % box around its start point: no more in-box shortcuts
};
struct alignas(16) RayCold {
  double x0[3];
  double* rec;  // crossing records, ART_CROSS_REC doubles each
  int rec_cap, max_cross;
  int br_a, br_b, br_sg;  // bracket found by detect(), waiting for the root kernel (wavefront path)
  int pad2_;
  double s0[3];   // start point in spherical coordinates (r, θ, φ)
  double box_m;   // min_c |x0_c| · (1 − 1/1.0001): the start box's half width in its narrowest component
};
static_assert(sizeof(RayHot) == 192, "RayHot layout");
static_assert(sizeof(RayCold) == 96, "RayCold layout");

struct Ray : RayHot, RayCold {
  // transient: the attempted step between attempt() and apply()
  double un[6], s_new, dt_new, h;

  ART_HD double u7_of(double u6) const { return u7_0 + fac7 * (u6 - u6_0); }

  // ---- set-up: IC transform (RT.jl:179-216).  startup() then produces the first RHS, the initial
  // step size and the first condition value.
  ART_HD void init(const DevParams& P, const RayIn& in, double* rec_, int rec_cap_) {
    rec = rec_;
    rec_cap = rec_cap_;
    max_cross = in.max_cross;
    erg = in.erg;
    axion = in.species != SPECIES_PHOTON;
    for (int i = 0; i < 3; ++i) x0[i] = in.x0[i];
    ray_initial_state(P, in.x0, in.k0, erg, in.dOmega, u, u7_0);
    u6_0 = u[5];
    for (int i = 0; i < 3; ++i) s0[i] = u[i];
    box_m = fmin(fmin(fabs(x0[0]), fabs(x0[1])), fabs(x0[2])) * (1.0 - 1.0 / 1.0001) * (1.0 - 1e-6);
    left_box = 0;
    fac7 = axion ? 0.0 : P.Omega * erg / C_KM;
    s = log(fmax(in.t_start, P.t_floor));  // MR.jl:166
    dt = 0.0;
    s_new = s;
    dt_new = 0.0;
    h = 0.0;
    sgn_end = 0;
    n_cross = iters = n_acc = n_rej = n_rhs = n_cond = 0;
    br_a = br_b = br_sg = 0;
    pad2_ = 0;
    status = RAY_END;
    event_last = false;
    starting = true;
    for (int i = 0; i < 6; ++i) k1[i] = 0.0;
    qold = dp::qoldinit;
    done = !(s < P.ln_tend);
  }

  // func! / func_axion!, RT.jl:71-123
  // SP = −1: species from the ray (mixed batches); 0 / 1: compile-time axion / photon
  template <int SP = -1>
  ART_HD void f(const DevParams& P, const double v[6], double lnt, double dv[6]) const {
    ft<SP>(P, v, fexp(lnt), dv);
  }
  template <int SP = -1>
  ART_HD void ft(const DevParams& P, const double v[6], double t, double dv[6]) const {
    const bool ax = SP < 0 ? axion : (SP == SPECIES_AXION);
    const double einv = frcp(erg);
    if (ax)
      rhs_axion(P, v[0], v[1], v[3] * erg, v[4] * erg, v[5] * erg, t, erg, einv, dv);
    else
      rhs_photon(P, v[0], v[1], v[2], v[3] * erg, v[4] * erg, v[5] * erg, -u7_of(v[5]), t, einv, dv);
  }

  // `condition` of propagate (RT.jl:254-298) at a state v, time lnt.  Exact shortcut: for
  // r > Rcut the plasma frequency is below m_a in every direction, so the condition is strictly
  // negative whenever it is defined (NrmSq ≥ 0 ⇔ E² r ≥ m_a² (r − rs) in the exterior metric) —
  // 56 % of all evaluations of a GR run.
  ART_HD double cond(const DevParams& P, const double v[6], double lnt) const {
    const double E = u7_of(v[5]);
    const double r = v[0];
    if (r > P.Rcut && r > P.r_int) return (E * E * r >= P.ma2 * (r - P.rs_ray)) ? -1.0 : NAN;
    const double t = fexp(lnt);
    return thick_F(P, r, v[1], v[2] - P.Omega * t, v[3], v[4], v[5], E, P.rs_ray, true);
  }

  ART_HD static double rms7(const double a[6], double a7) {
    double q = a7 * a7;
    for (int i = 0; i < 6; ++i) q += a[i] * a[i];
    return sqrt(q / 7.0);
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

  // ---------------------------------------------------------------------------------------
  // One iteration = [startup() once] → attempt() → publish() → crossing scan of the accepted step →
  // finish_step().  On the GPU these are separate kernels over compacted ray lists
  // (art_wavefront.cuh); step() below is the sequential definition they must reproduce.
  // Only SIGNS of the condition are kept by the scan (three 64-bit masks per ray: positive /
  // negative / exactly zero; neither = NaN); the two bracket values the root finder starts from are
  // re-evaluated, which gives the same numbers.
  // ---------------------------------------------------------------------------------------

  // guard shared by start-up and attempted steps (OrdinaryDiffEq maxiters)
  ART_HD bool iter_guard(const DevParams& P) {
    if (++iters > P.maxiters) {
      status = RAY_MAXITERS;
      done = true;
      return false;
    }
    return true;
  }

  // startup(): first RHS, OrdinaryDiffEq ode_determine_initdt (Hairer II.4; order 5, 7-component RMS
  // norm), reference sign of the crossing condition at the start point.
  template <int SP = -1>
  ART_HD void startup(const DevParams& P) {
    const double s1 = P.ln_tend;
    f<SP>(P, u, s, k1);
    double sk[6], a[6], v[6], dv[6];
    for (int i = 0; i < 6; ++i) sk[i] = P.atol + fabs(u[i]) * P.rtol;
    const double sk7 = P.atol + fabs(u7_0) * P.rtol;
    for (int i = 0; i < 6; ++i) a[i] = u[i] / sk[i];
    const double d0 = rms7(a, u7_0 / sk7);
    for (int i = 0; i < 6; ++i) a[i] = k1[i] / sk[i];
    const double d1 = rms7(a, fac7 * k1[5] / sk7);
    double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
    dt0 = fmin(dt0, s1 - s);
    for (int i = 0; i < 6; ++i) v[i] = u[i] + dt0 * k1[i];
    f<SP>(P, v, s + dt0, dv);
    n_rhs += 2;
    for (int i = 0; i < 6; ++i) a[i] = (dv[i] - k1[i]) / sk[i];
    const double d2 = rms7(a, fac7 * (dv[5] - k1[5]) / sk7) / dt0;
    const double dm = fmax(d1, d2);
    const double dt1 = (dm <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : exp(-(2.0 + log10(dm)) / 5.0 * 2.302585092994046);
    dt = fmin(fmin(100.0 * dt0, dt1), s1 - s);
    const double g = cond(P, u, s);
    n_cond++;
    sgn_end = (g > 0.0) - (g < 0.0);
    starting = false;
  }

  // v = u + h Σ_{j<ST} a_{ST+1,j+1} k_j with the tableau row as constant-bank operands and the stage store
  // at static offsets (a j loop with a run-time bound spent 21 % of the stage kernel's instructions on
  // address and loop arithmetic — profiles/r01_wf10_inst_by_line.txt)
  template <int ST, class KS>
  ART_HD void combine(const KS K, double v[6]) const {
    for (int i = 0; i < 6; ++i) v[i] = u[i];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < ST; ++j) {
      if (ST == 6 && j == 1) continue;  // b2 = 0: the term is an exact zero
      const double c = h * DP_A[ST][j];
      for (int i = 0; i < 6; ++i) v[i] += c * K(j, i);
    }
  }

  // attempt(): one attempted DP5 step with error control (stage loop rolled: one copy of the
  // right-hand side).  Returns true if the step was accepted; un, s_new, dt_new, h and K then
  // describe it until apply().
  template <int SP = -1, class KS = KStore>
  ART_HD bool attempt(const DevParams& P, const KS K) {
    const double s1 = P.ln_tend;
    h = dt;
    bool last = false;
    if (s + h >= s1) {
      h = s1 - s;
      last = true;
    }
    s_new = last ? s1 : s + h;
    for (int i = 0; i < 6; ++i) K(0, i) = k1[i];
    // the six stage times are known up front: their exponentials are evaluated together (five independent
    // chains) instead of heading the dependent chain of every right-hand side
    if (KS::PARKS) {
      for (int st = 1; st <= 5; ++st) K.park(st - 1) = fexp(s + DP_C[st] * h);
      K.park(5) = fexp(s_new);
    }
    for (int st = 1; st <= 6; ++st) {
      double v[6];
      switch (st) {  // warp-uniform: the stage loop stays rolled (one copy of the right-hand side)
        case 1: combine<1>(K, v); break;
        case 2: combine<2>(K, v); break;
        case 3: combine<3>(K, v); break;
        case 4: combine<4>(K, v); break;
        case 5: combine<5>(K, v); break;
        default: combine<6>(K, v); break;
      }
      double dv[6];
      if (KS::PARKS) {
        ft<SP>(P, v, K.park(st - 1), dv);
      } else {
        const double ts = (st == 6) ? s_new : s + DP_C[st] * h;
        f<SP>(P, v, ts, dv);
      }
      for (int i = 0; i < 6; ++i) K(st, i) = dv[i];
      if (st == 6)
        for (int i = 0; i < 6; ++i) un[i] = v[i];
    }
    n_rhs += 6;

    // ---- error control (divisions as reciprocals, no special-case branches) ---------------------------
    double q = 0.0, e6 = 0.0;
    for (int i = 0; i < 6; ++i) {
      double e = DP_E[0] * K(0, i);  // e2 = 0
      e += DP_E[2] * K(2, i);
      e += DP_E[3] * K(3, i);
      e += DP_E[4] * K(4, i);
      e += DP_E[5] * K(5, i);
      e += DP_E[6] * K(6, i);
      e *= h;
      if (i == 5) e6 = e;
      const double r_ = e * frcp(P.atol + fmax(fabs(u[i]), fabs(un[i])) * P.rtol);
      q += r_ * r_;
    }
    const double r7 = fac7 * e6 * frcp(P.atol + fmax(fabs(u7_of(u[5])), fabs(u7_of(un[5]))) * P.rtol);
    q += r7 * r7;
    const double EEst = (q > 0.0) ? fsqrt(q * (1.0 / 7.0)) : q;  // 0 stays 0, NaN stays NaN
    // PI controller: q11 = EEst^β1, q = q11 / qold^β2
    double q11 = fexp(dp::beta1 * flog(EEst));
    double qq = q11 * fexp(-dp::beta2 * flog(qold));
    qq = fmax(1.0 / dp::qmax, fmin(1.0 / dp::qmin, qq * (1.0 / dp::gamma)));
    if (EEst == 0.0) qq = 1.0 / dp::qmax;
    bool accept = (EEst <= 1.0) || (h <= P.dtmin);  // force_dtmin
    if (!(EEst == EEst)) {
      accept = false;
      q11 = 1e30;
    }
    if (!accept) {
      n_rej++;
      dt = h * frcp(fmin(1.0 / dp::qmin, q11 * (1.0 / dp::gamma)));
      if (dt < P.dtmin) dt = P.dtmin;
      return false;
    }
    n_acc++;
    if (qq <= dp::qsteady_max && qq >= dp::qsteady_min) qq = 1.0;
    qold = fmax(EEst, dp::qoldinit);
    dt_new = h * frcp(qq);
    n_cond += P.M_ray - 1 + (event_last ? 1 : 0);
    return true;
  }

  // Dense-output coefficients of component i of the accepted step (Hairer contd5: rcont1 = u,
  // rcont2..5 = r1..r4)
  template <class KS>
  ART_HD void dense_comp(const KS K, int i, double r[4]) const {
    double d = 0.0;
    for (int j = 0; j < 7; ++j) d += DP_D[j] * K(j, i);
    const double a1 = un[i] - u[i];
    const double a2 = h * K(0, i) - a1;
    r[0] = a1;
    r[1] = a2;
    r[2] = a1 - h * K(6, i) - a2;
    r[3] = h * d;
  }

  // publish the dense-output segment of the accepted step
  template <class KS, class DS>
  ART_HD void publish(const KS K, const DS D) const {
    for (int i = 0; i < 6; ++i) {
      double r[4];
      dense_comp(K, i, r);
      D(SEG_U + i) = u[i];
      D(SEG_R1 + i) = r[0];
      D(SEG_R2 + i) = r[1];
      D(SEG_R3 + i) = r[2];
      D(SEG_R4 + i) = r[3];
    }
    D(SEG_S) = s;
    D(SEG_H) = h;
    D(SEG_U70) = u7_0;
    D(SEG_FAC7) = fac7;
    D(SEG_U60) = u6_0;
  }

  // Whole-step far-field test straight from the stage store (no segment needed): true ⇒ every scan
  // point of the accepted step, the nudge point included, is provably negative (SegQuick::all_far).
  template <class KS>
  ART_HD bool step_all_far(const DevParams& P, const KS K) const {
    double r0[4], r5[4];
    dense_comp(K, 0, r0);
    const double lo = fmin(u[0], u[0] + r0[0]) - 0.25 * (fabs(r0[1]) + fabs(r0[2]) + 0.25 * fabs(r0[3]));
    if (!(lo > fmax(P.Rcut, P.r_int) * (1.0 + 1e-9))) return false;
    dense_comp(K, 5, r5);
    const double e0 = u7_0 + fac7 * (u[5] - u6_0), e1 = e0 + fac7 * r5[0];
    const double de = fabs(fac7) * 0.25 * (fabs(r5[1]) + fabs(r5[2]) + 0.25 * fabs(r5[3]));
    const double emin = fmin(fabs(e0), fabs(e1)) - de;
    return emin > 0.0 && emin * emin >= P.ma2 * (1.0 + 1e-12);
  }

  // Direction-aware refinement of step_all_far: the sign of the condition along the WHOLE accepted step, when
  // it can be certified without the momentum:  F = base·(Q|b_z| − m_a² r³) − Q|b_z|·g^rr NrmSq·ww N'² with
  // 0 ≤ g^rr k∥² ≤ E² − m_a² g^rr (Cauchy–Schwarz), so
  //   Q|b_z| < m_a² r³            ⇒ F < 0   ("dark": the plasma frequency is below the axion mass), return −1
  //   Q|b_z| g^rr > E² r³         ⇒ F > 0   ("lit": well inside the conversion surface),            return +1
  // and only the thin shell in between (≈ 6 % in radius) needs the scan.  Photons leave the surface outwards
  // (dark until Rcut), back-traced axions run inwards (lit).  Rigorous bounds: r, θ, φ stay within chord ±
  // ¼(|r2|+|r3|+¼|r4|) of the dense output; ψ = φ − Ωt with t ∈ [e^s, e^{s_new}]; |b_z| within
  // |b_z(centre)| ± [3(|cosθm| + sinθm)·δθ + 1.5 sinθm·δψ] (sup of its partials); E within its chord ± dev.
  template <class KS>
  ART_HD int step_sign(const DevParams& P, const KS K, double t_new) const {
    double lo[3], hi[3];
    for (int i = 0; i < 3; ++i) {
      double r[4];
      dense_comp(K, i, r);
      const double dev = 0.25 * (fabs(r[1]) + fabs(r[2]) + 0.25 * fabs(r[3]));
      lo[i] = fmin(u[i], u[i] + r[0]) - dev;
      hi[i] = fmax(u[i], u[i] + r[0]) + dev;
    }
    if (!(lo[0] > P.r_int && lo[0] > P.rNS)) return 0;
    double r5[4];
    dense_comp(K, 5, r5);
    const double e0 = u7_0 + fac7 * (u[5] - u6_0), e1 = e0 + fac7 * r5[0];
    const double de = fabs(fac7) * 0.25 * (fabs(r5[1]) + fabs(r5[2]) + 0.25 * fabs(r5[3]));
    const double emin = fmin(fabs(e0), fabs(e1)) - de, emax = fmax(fabs(e0), fabs(e1)) + de;
    if (!(emin > 0.0 && emin * emin >= P.ma2 * (1.0 + 1e-12))) return 0;  // NrmSq ≥ 0 over the step
    const double t_old = fexp(s);
    const double p0 = lo[2] - P.Omega * t_new, p1 = hi[2] - P.Omega * t_old;
    const double thm = 0.5 * (lo[1] + hi[1]), dth = 0.5 * (hi[1] - lo[1]);
    const double psm = 0.5 * (p0 + p1), dps = 0.5 * (p1 - p0);
    double sn, cs, sp, cp;
    fsincos(thm, &sn, &cs);
    fsincos(psm, &sp, &cp);
    const double bzm = fabs(P.cm * (3.0 * cs * cs - 1.0) + 3.0 * P.sm * sn * cs * cp);
    const double slack = 3.0 * (fabs(P.cm) + fabs(P.sm)) * dth + 1.5 * fabs(P.sm) * dps;
    const double r3lo = lo[0] * lo[0] * lo[0], r3hi = hi[0] * hi[0] * hi[0];
    if (P.Q * (bzm + slack) < P.ma2 * r3lo * (1.0 - 1e-9)) return -1;
    const double grr_min = 1.0 - P.rs_ray / lo[0];
    if (!P.isotropic && P.Q * (bzm - slack) * grr_min > emax * emax * r3hi * (1.0 + 1e-9)) return 1;
    return 0;
  }

  // Does the whole accepted step lie inside the ±0.01 % box around the start point (RT.jl:307-311)?  Then every
  // sign change on it is discarded by the start-box rule and only the sign at its end point survives (as the
  // next step's reference sign): ONE evaluation instead of M−1, done by the stage kernel itself.  A ray spends its first
  // 10–15 steps there while the step size grows from the initial 1e-6 — more than half of all condition
  // evaluations of a run.  Rigorous bound: each spherical coordinate stays within its chord ± ¼(|r2|+|r3|+¼|r4|)
  // (as in step_all_far), and a Cartesian component moves by at most δr + r_max(δθ + δφ).
  template <class KS>
  ART_HD bool step_in_start_box(const KS K, const double s0_[3], double box_m_) const {
    double d[3];
    for (int i = 0; i < 3; ++i) {
      double r[4];
      dense_comp(K, i, r);
      const double dev = 0.25 * (fabs(r[1]) + fabs(r[2]) + 0.25 * fabs(r[3]));
      d[i] = fmax(fabs(u[i] - s0_[i]), fabs(u[i] + r[0] - s0_[i])) + dev;
    }
    return d[0] + (s0_[0] + d[0]) * (d[1] + d[2]) < box_m_;
  }

  // stages(): sequential form of start-up / attempted step (host check).  Returns true if a step was
  // accepted; its dense-output segment has then been published in D and must be scanned before
  // finish_step().
  template <int SP = -1, class KS = KStore, class DS = SegStore>
  ART_HD bool stages(const DevParams& P, const KS K, const DS D) {
    if (!iter_guard(P)) return false;
    if (starting) {
      startup<SP>(P);
      return false;
    }
    if (!attempt<SP>(P, K)) return false;
    publish(K, D);
    return true;
  }

  // is the dense-output point th still inside the ±0.01 % box around the start point (RT.jl:307-311)?
  template <class DS>
  ART_HD bool in_start_box(const DevParams& P, const DS D, double th) const {
    double w[6], x[3], k[3];
    seg_state(D, th, w);
    state_to_cartesian(P, w, erg, false, x, k);
    const double sb = 1.0001;
    bool in = true;
    for (int c = 0; c < 3; ++c) in = in && (fabs(x[c]) < fabs(x0[c]) * sb) && (fabs(x[c]) > fabs(x0[c]) / sb);
    return in;
  }

  // finish_step(): first sign change of the scanned step → root → callbacks → state update.
  template <class KS, class DS>
  ART_HD void finish_step(const DevParams& P, const KS K, const DS D, const ScanMasks m) {
    int ia = 0, ib = -1, sg_last = 0;
    const int ev = detect(P, D, m, ia, ib, sg_last);
    double th_ev = -1.0;
    if (ev == 1) th_ev = locate(P, D, ia, ib);
    else if (ev == 2) th_ev = 1.0;
    apply(P, K, D, th_ev, sg_last);
  }

  // detect(): first sign change of the scanned step.  0 = none, 1 = bracket (ia, ib) found,
  // 2 = the end point lies exactly on the surface (prev_sign · next_sign ≤ 0).
  template <class DS>
  ART_HD int detect(const DevParams& P, const DS D, const ScanMasks m, int& ia, int& ib, int& sg_last) const {
    const int M = P.M_ray;
    int cur = sgn_end;
    if (event_last)  // DiffEqBase repeat_nudge = 1/100: reference sign taken just after the event
      cur = (int)(m.pos & 1ull) - (int)(m.neg & 1ull);
    // grid points that take part in the comparison: 1..M−1 (after an event only those beyond the nudge)
    unsigned long long valid = ((M >= 64) ? ~0ull : ((1ull << M) - 1ull)) & ~1ull;
    if (event_last)
      for (int i = 1; i < M - 1 && (double)i * P.inv_Mrm1 <= 0.01; ++i) valid &= ~(1ull << i);
    ia = 0;   // 0 stands for the left end (θ = 0, or 0.01 after an event)
    ib = -1;
    unsigned long long from = valid;
    while (cur != 0) {
      const unsigned long long opp = (cur > 0 ? m.neg : m.pos) & from;
      if (!opp) break;
      const int i = ffs64(opp) - 1;
      const unsigned long long below = (cur > 0 ? m.pos : m.neg) & from & ((1ull << i) - 1ull);
      if (below) ia = 63 - clz64(below);
      // start-box rule: a sign change whose grid point is still inside the box
      // around the start point is discarded without locating it; the reference sign flips
      if (n_cross == 0 && in_start_box(P, D, seg_theta(P, i))) {
        cur = -cur;
        ia = i;
        from &= ~((2ull << i) - 1ull);
        continue;
      }
      ib = i;
      break;
    }
    sg_last = (int)((m.pos >> (M - 1)) & 1ull) - (int)((m.neg >> (M - 1)) & 1ull);
    if (ib >= 0) return 1;
    if (cur != 0 && ((m.zero >> (M - 1)) & 1ull)) return 2;
    return 0;
  }

  // locate(): bracketing root finder of the event time — Illinois regula falsi with Brent's minimum
  // step; bisection whenever the bracket has not halved over two iterations, and for good once three
  // iterations have not brought |f| down (a jump, e.g. ωp switched off at r = rNS, has no zero to
  // converge to).  Run to 1e-13 of the step (the condition's own round-off noise is wider); left-biased.
  // Values come from seg_cond_full (far-field shortcut values would not be on the scale of thick_F).
  template <class DS>
  ART_HD double locate(const DevParams& P, const DS D, int ia, int ib) {
    double ra = (ia == 0) ? (event_last ? 0.01 : 0.0) : seg_theta(P, ia), rb = seg_theta(P, ib);
    double fa = seg_cond_full(P, D, ra), fb = seg_cond_full(P, D, rb);
    n_cond += 2;
    // ωp is switched off inside the star (zeroIn, RT.jl:1165), so the condition JUMPS where the step
    // crosses r = rNS and a bracketing iteration could only bisect its way there (≈ 40 evaluations of
    // the field).  A bracket that straddles the stellar surface is therefore split at the surface first,
    // by bisection on the dense-output polynomial of r alone; two evaluations of the condition either
    // side then decide whether the jump is the event or which smooth half holds the sign change.
    {
      const double ga = seg_comp(D, 0, ra, 1.0 - ra) - P.rNS, gb = seg_comp(D, 0, rb, 1.0 - rb) - P.rNS;
      if ((ga <= 0.0) != (gb <= 0.0)) {  // inside ⇔ r ≤ rNS, as in thick_F
        double lo = ra, hi = rb;
        for (int k = 0; k < 64; ++k) {
          if (hi - lo <= ROOT_RTOL * fmax(fmax(fabs(lo), fabs(hi)), 0.01)) break;
          const double mid = 0.5 * (lo + hi);
          const double gm = seg_comp(D, 0, mid, 1.0 - mid) - P.rNS;
          if ((gm <= 0.0) == (ga <= 0.0)) lo = mid;
          else hi = mid;
        }
        const double flo = seg_cond_full(P, D, lo), fhi = seg_cond_full(P, D, hi);
        n_cond += 2;
        if ((flo < 0.0) != (fa < 0.0)) {
          rb = lo; fb = flo;
        } else if ((fhi < 0.0) != (flo < 0.0)) {
          return lo;
        } else {
          ra = hi; fa = fhi;
        }
      }
    }
    const double f0 = fmin(fabs(fa), fabs(fb));
    double w1 = 2.0, w2 = 2.0;
    bool pure = false;
    int side = 0;
    for (int it = 0; it < 200; ++it) {
      const double tol = ROOT_RTOL * fmax(fmax(fabs(ra), fabs(rb)), 0.01);  // 1e-13 of the step, floor 1e-15
      const double width = fabs(rb - ra);
      if (width <= tol) break;
      if (it == 3 && fmin(fabs(fa), fabs(fb)) > 0.25 * f0) pure = true;
      double c = (ra * fb - rb * fa) * frcp(fb - fa);
      const double lo = fmin(ra, rb), hi = fmax(ra, rb);
      if (pure || (it >= 3 && width > 0.5 * w2) || !(c > lo && c < hi) || width <= 2.0 * tol) c = 0.5 * (ra + rb);
      else if (fabs(c - rb) < tol) c = rb + copysign(tol, ra - rb);  // a converged iterate collapses the bracket
      w2 = w1;
      w1 = width;
      const double fc = seg_cond_full(P, D, c);
      n_cond++;
      if (fc == 0.0) return c;
      if ((fc < 0.0) != (fb < 0.0)) {
        ra = rb; fa = fb; side = 0;
      } else {
        if (side == 1) fa *= 0.5;
        side = 1;
      }
      rb = c; fb = fc;
    }
    return fmin(ra, rb);
  }

  // termination checks at the end of an accepted step: DiscreteCallback (photon hits the star,
  // RT.jl:352-359) and the end of the span
  ART_HD void settle(const DevParams& P) {
    if (!axion && u[0] < P.rNS101) {
      status = RAY_HIT_NS;
      done = true;
      return;
    }
    if (!(s < P.ln_tend)) done = true;
  }

  // apply(): callbacks and state update at the end of an accepted step (th_ev < 0: no event).
  // K(6, ·) = k7 of the step (FSAL).  After an event the derivative at the event point is recomputed
  // here unless the ray is finished (the reference's solver does it at the start of its next step).
  template <class KS, class DS>
  ART_HD void apply(const DevParams& P, const KS K, const DS D, double th_ev, int sg_last) {
    bool terminated = false;
    if (th_ev >= 0.0) {
      double ue[6];
      double se;
      if (th_ev >= 1.0) {
        for (int i = 0; i < 6; ++i) ue[i] = un[i];
        se = s_new;
      } else {
        seg_state(D, th_ev, ue);
        se = s + th_ev * h;
      }
      terminated = on_crossing(P, ue, se);
      for (int i = 0; i < 6; ++i) u[i] = ue[i];
      s = se;
      event_last = true;
    } else {
      for (int i = 0; i < 6; ++i) {
        u[i] = un[i];
        k1[i] = K(6, i);  // FSAL
      }
      s = s_new;
      sgn_end = sg_last;
      event_last = false;
    }
    dt = dt_new;
    if (terminated) done = true;
    else settle(P);
    if (th_ev >= 0.0 && !done) {  // FSAL invalidated by the event
      f<-1>(P, u, s, k1);
      n_rhs++;
    }
  }

  // sequential form (host check, and the definition the GPU kernels must reproduce)
  ART_HD void step(const DevParams& P, const KStore K, const SegStore D) {
    if (!stages(P, K, D)) return;
    ScanMasks m = {0ull, 0ull, 0ull};
    const int M = P.M_ray;
    for (int i = event_last ? 0 : 1; i < M; ++i) {
      const double g = seg_cond(P, D, seg_theta(P, i));
      if (g > 0.0) m.pos |= 1ull << i;
      else if (g < 0.0) m.neg |= 1ull << i;
      else if (g == 0.0) m.zero |= 1ull << i;
    }
    finish_step(P, K, D, m);
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
