// art_physics.cuh — closed-form device physics of the hot path (sm_100a, FP64 throughout).
//
// Everything the reference obtains through ForwardDiff dual numbers on freshly allocated
// 1×N matrices (RayTracer.jl:21,24,84-88) is evaluated here from hand-derived, algebraically
// reduced partial derivatives held in registers.  The functions are `__host__ __device__`
// so that tests/ can compile the very same source for the CPU and compare it with the oracle
// without a GPU (tests/hostcheck); the shipped library only ever calls them from kernels.
//
// Notation:  c = cosθ, s = sinθ, ψ = φ − Ωt,
//   b_r = 2(cosθm c + sinθm s cosψ), b_θ = cosθm s − sinθm c cosψ, b_φ = sinθm sinψ
//   (dipole components in units of Bn = B0 (rNS/r)^3 / 2,          RT.jl:1142-1147)
//   b_z = b_r c − b_θ s,  ωp² = Q |b_z| / r³                        (RT.jl:1151-1154)
//   D = b_r² + b_θ² + b_φ²,  N = √g^rr k_r b_r + k_θ b_θ / r + k_φ b_φ /(r s)
//   k∥² = N² / D                                                     (K_par, RT.jl:1044-1058)
//   H = ½[g^tt ω² + g^rr k_r² + k_θ²/r² + k_φ²/(r² s²) + ωp² (1 − g^rr N²/(D ω²))]
//                                                                    (hamiltonian, RT.jl:530-556)
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define ART_HD __host__ __device__ __forceinline__
#else
#define ART_HD inline
#endif

namespace art {

// Constants.jl:3-5 — these exact values define parity.
constexpr double C_KM = 2.99792e5;
constexpr double HBAR = 6.582119e-16;
constexpr double GNEW = 1.32712e11;
constexpr double PI = 3.14159265358979323846;

// ---------------------------------------------------------------------------------------
// Branch-free FP64 helpers for the hot loops.  CUDA's IEEE `/`, sqrt() and sincos() each carry a
// special-case branch (and sincos a Payne–Hanek path with a local-memory frame); on the ray equations
// those made 63 % of the executed instructions non-FP64 (profiles/r01_wf5_*).  Arguments here are
// finite, normal and of moderate size, so: MUFU seed + two Newton steps (≤ 1–2 ulp), and a
// Cody–Waite reduction with fdlibm's kernel polynomials for |x| < 1e5 (≤ 1 ulp).
// The host build (tests/hostcheck) uses the plain operators for 1/x and sqrt.
// ---------------------------------------------------------------------------------------
ART_HD double frcp(double x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
#else
  return 1.0 / x;
#endif
}
ART_HD double frsqrt(double x) {  // x > 0
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  double e = fma(-hx * y, y, 0.5);
  y = fma(y, e, y);
  e = fma(-hx * y, y, 0.5);
  return fma(y, e, y);
#else
  return 1.0 / sqrt(x);
#endif
}
ART_HD double fsqrt(double x) {  // x > 0
#if defined(__CUDA_ARCH__)
  const double y = frsqrt(x);
  double s = x * y;
  const double r = fma(-s, s, x);
  return fma(r, 0.5 * y, s);
#else
  return sqrt(x);
#endif
}
// Polynomial coefficients live in the constant bank: a 64-bit literal costs two move instructions per
// use (17 % of the stage kernel's instructions were IMAD.MOV/UMOV pairs feeding DFMA), a constant-bank
// operand none.
// (Host code compiled by nvcc must not read a __constant__ symbol — it would see the host shadow — so
// each table exists twice and the name resolves per compilation pass.)
#define ART_FM_SIN {-1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04, \
                    2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10}
#define ART_FM_COS {4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05, \
                    -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11}
#define ART_FM_PIO2 {0.63661977236758134308, 1.57079632679489655800e+00, 6.12323399573676603587e-17, 8.4784276603688985e-32}
// 1/n!, n = 2..13
#define ART_FM_EXP {1.0 / 2, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320, 1.0 / 362880, \
                    1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600, 1.0 / 6227020800.0}
#define ART_FM_LN2 {1.4426950408889634074, 6.93147180369123816490e-01, 1.90821492927058770002e-10}
// 1/(2n+1), n = 1..9 (atanh series)
#define ART_FM_LOG {1.0 / 3, 1.0 / 5, 1.0 / 7, 1.0 / 9, 1.0 / 11, 1.0 / 13, 1.0 / 15, 1.0 / 17, 1.0 / 19}
#if defined(__CUDACC__)
__device__ __constant__ double FM_SIN_D[6] = ART_FM_SIN;
__device__ __constant__ double FM_COS_D[6] = ART_FM_COS;
__device__ __constant__ double FM_PIO2_D[4] = ART_FM_PIO2;
__device__ __constant__ double FM_EXP_D[12] = ART_FM_EXP;
__device__ __constant__ double FM_LN2_D[3] = ART_FM_LN2;
__device__ __constant__ double FM_LOG_D[9] = ART_FM_LOG;
#endif
#if defined(__CUDA_ARCH__)
#define FM_SIN FM_SIN_D
#define FM_COS FM_COS_D
#define FM_PIO2 FM_PIO2_D
#define FM_EXP FM_EXP_D
#define FM_LN2 FM_LN2_D
#define FM_LOG FM_LOG_D
#else
static const double FM_SIN[6] = ART_FM_SIN;
static const double FM_COS[6] = ART_FM_COS;
static const double FM_PIO2[4] = ART_FM_PIO2;
static const double FM_EXP[12] = ART_FM_EXP;
static const double FM_LN2[3] = ART_FM_LN2;
static const double FM_LOG[9] = ART_FM_LOG;
#endif

ART_HD void fsincos(double x, double* sn, double* cs) {
  const double k = rint(x * FM_PIO2[0]);  // 2/π
  double r = fma(-k, FM_PIO2[1], x);
  r = fma(-k, FM_PIO2[2], r);
  r = fma(-k, FM_PIO2[3], r);
  const double z = r * r;
#if !defined(ART_ESTRIN)
  double ps = fma(z, FM_SIN[5], FM_SIN[4]);
  ps = fma(z, ps, FM_SIN[3]);
  ps = fma(z, ps, FM_SIN[2]);
  ps = fma(z, ps, FM_SIN[1]);
  ps = fma(z, ps, FM_SIN[0]);
  double pc = fma(z, FM_COS[5], FM_COS[4]);
  pc = fma(z, pc, FM_COS[3]);
  pc = fma(z, pc, FM_COS[2]);
  pc = fma(z, pc, FM_COS[1]);
  pc = fma(z, pc, FM_COS[0]);
#else
  // Estrin (tuning variant, -DART_ESTRIN): pairs in z, combined in z² — three dependent FMAs per polynomial instead of
  // five.  Measured 2 % SLOWER on the whole step than Horner (580 vs 567 ms per 3e6 events): the extra multiplies and live
  // values cost more than the shorter chains save
  const double z2 = z * z;
  const double s01 = fma(z, FM_SIN[1], FM_SIN[0]), s23 = fma(z, FM_SIN[3], FM_SIN[2]), s45 = fma(z, FM_SIN[5], FM_SIN[4]);
  const double c01 = fma(z, FM_COS[1], FM_COS[0]), c23 = fma(z, FM_COS[3], FM_COS[2]), c45 = fma(z, FM_COS[5], FM_COS[4]);
  const double ps = fma(z2, fma(z2, s45, s23), s01);
  const double pc = fma(z2, fma(z2, c45, c23), c01);
#endif
  const double S = fma(r * z, ps, r);
  const double C = fma(z * z, pc, fma(-0.5, z, 1.0));
  const int n = (int)k;
  const double a = (n & 1) ? C : S, b = (n & 1) ? S : C;
  *sn = (n & 2) ? -a : a;
  *cs = ((n + 1) & 2) ? -b : b;
}

// e^x for x < 700 (arguments below −700 are clamped there): x = k ln2 + r, |r| ≤ ln2/2, Taylor to r¹³ (truncation 4e-18), scaled through the
// exponent field.  Branch-free; ≤ 1 ulp.
ART_HD double fexp(double x) {
  x = fmax(x, -700.0);  // (the exponent-field scaling below has no underflow path: e^-700 = 1e-304 stands in for 0)
  const double k = rint(x * FM_LN2[0]);
  double r = fma(-k, FM_LN2[1], x);
  r = fma(-k, FM_LN2[2], r);
#if !defined(ART_ESTRIN)
  double p = fma(r, FM_EXP[11], FM_EXP[10]);
  p = fma(r, p, FM_EXP[9]);
  p = fma(r, p, FM_EXP[8]);
  p = fma(r, p, FM_EXP[7]);
  p = fma(r, p, FM_EXP[6]);
  p = fma(r, p, FM_EXP[5]);
  p = fma(r, p, FM_EXP[4]);
  p = fma(r, p, FM_EXP[3]);
  p = fma(r, p, FM_EXP[2]);
  p = fma(r, p, FM_EXP[1]);
  p = fma(r, p, FM_EXP[0]);
#else
  // Estrin: six pairs in r, three quadruples in r², combined in r⁴ — five dependent operations instead of eleven
  const double r2 = r * r, r4 = r2 * r2;
  const double e01 = fma(r, FM_EXP[1], FM_EXP[0]), e23 = fma(r, FM_EXP[3], FM_EXP[2]), e45 = fma(r, FM_EXP[5], FM_EXP[4]);
  const double e67 = fma(r, FM_EXP[7], FM_EXP[6]), e89 = fma(r, FM_EXP[9], FM_EXP[8]), eab = fma(r, FM_EXP[11], FM_EXP[10]);
  const double q0 = fma(r2, e23, e01), q1 = fma(r2, e67, e45), q2 = fma(r2, eab, e89);
  double p = fma(r4, fma(r4, q2, q1), q0);
#endif
  p = fma(r * r, p, r) + 1.0;
#if defined(__CUDA_ARCH__)
  return __hiloint2double(__double2hiint(p) + ((int)k << 20), __double2loint(p));
#else
  return ldexp(p, (int)k);
#endif
}

// ln x for finite x > 0 (0 gives ≈ −709, which every caller clamps): exponent split, m ∈ [√½, √2),
// ln m = 2 atanh(s), s = (m − 1)/(m + 1), series to s¹⁹ (truncation 2e-17).  Branch-free; ≤ 2 ulp.
ART_HD double flog(double x) {
#if defined(__CUDA_ARCH__)
  const int hi = __double2hiint(x);
  double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
#else
  unsigned long long bits;
  memcpy(&bits, &x, 8);
  const int hi = (int)(bits >> 32);
  bits = (bits & 0x000fffffffffffffull) | 0x3ff0000000000000ull;
  double m;
  memcpy(&m, &bits, 8);
#endif
  int e = (hi >> 20) - 1023;
  const bool up = m > 1.4142135623730951;
  m = up ? 0.5 * m : m;
  e += up ? 1 : 0;
  const double s = (m - 1.0) * frcp(m + 1.0);
  const double z = s * s;
  double p = fma(z, FM_LOG[8], FM_LOG[7]);
  p = fma(z, p, FM_LOG[6]);
  p = fma(z, p, FM_LOG[5]);
  p = fma(z, p, FM_LOG[4]);
  p = fma(z, p, FM_LOG[3]);
  p = fma(z, p, FM_LOG[2]);
  p = fma(z, p, FM_LOG[1]);
  p = fma(z, p, FM_LOG[0]);
  const double lm = fma(s * z, p, s);  // atanh(s)
  const double ed = (double)e;
  return fma(ed, FM_LN2[1], fma(ed, FM_LN2[2], 2.0 * lm));
}

// Run constants, derived once on the host from art_params (art_api.cu: make_dev_params).
struct DevParams {
  double cm, sm;      // cos θm, sin θm
  double Omega;       // ωPul
  double Q;           // ωp² = Q |b_z| / r³  [eV² km³]
  double Bn0;         // |B0| rNS³ / 2   (field magnitude scale, G km³)
  double rNS, rNS101; // rNS, 1.01 rNS
  double rs_ray;      // 2GM/c² seen by the ray equations / crossing condition (0 when flat)
  double rs_real;     // 2GM/c² of the star (k_norm_Cart, conversion probability, sampler)
  double r_int;       // interior-metric switch radius (g_schwartz kwarg rNS = 10.0, RT.jl:455)
  double ma, ma2;     // axion mass
  double Ax_g;
  double t0;          // field time for sampler / weights / conversion probability
  double rtol, atol, dtmin;
  double ln_t0, ln_tend, t_floor;  // MR.jl:411-413; t_floor = exp(ln_t0) (MR.jl:153,166)
  double maxR, Rcut;  // Find_Conversion_Surface; Rcut = maxR/1.01 = radius where max ωp = m_a
  double samp_dt;     // 0.5 km
  double rho_DM;
  int32_t maxiters, M_ray, M_samp, n_max;
  int32_t isotropic, flat;
  int32_t melrose, thick;  // hamiltonian's melrose switch (RT.jl:546-551); sampler's thick_surface switch (RT.jl:1548)
  int32_t num_cutoff, MC_nodes, max_nodes;
  int32_t integrator;  // IG_VERN6 / IG_DP5 (art_tracer.cuh)
  double prob_cutoff;
  double inv_Mrm1, inv_Msm1;  // 1/(M_ray − 1), 1/(M_samp − 1): dense-output grid spacing
  double cpsi0, spsi0, mh[3];  // cos/sin Ω t0 and the magnetic axis m̂ at the field time t0 (sampler)
  double T_samp;               // length of the sampler's line march, 2.2 maxR (RT.jl:1611)
  double inv_r_int;            // 1 / r_int
  // boundary-layer term of the plasma frequency (bndry_lyr > 0, RT.jl:1155-1162): for r ≥ rNS
  //   ωp = √(Q|b_z|/r³) + ωpole (rNS/r)^{3/2} exp(−(r − rmax·bl)/(0.1 rmax)),  ωpole = √(2Q/rNS³), rmax = rNS (ωpole/m_a)^{2/3}
  double bl;                   // bndry_lyr (≤ 0: off)
  double wpole;                // plasma frequency at the pole of the surface
  double bl_rmax;              // rmax with the run's axion mass (sampler, ray condition, du7, g_det, ∇ωp of the conversion probability)
  double bl_rmax_def;          // rmax with GJ_Model_ωp_vecSPH's DEFAULT Mass_a = 1e-5: get_Prob_nonAD's own ωp call does not pass
                               // the mass (MR.jl:99) — quirk kept
};

// boundary-layer term B(r) and its r-derivative (r ≥ rNS; the callers apply the r ≥ rNS / zeroIn rules)
// (the branch-free helpers: with the library's sqrt / exp / division the boundary-layer kernels spent a third of their
// instructions in slow paths)
ART_HD double bl_w(const DevParams& P, double r, double rmax) {
  const double q = P.rNS * frcp(r);
  return P.wpole * q * fsqrt(q) * fexp(-(r - rmax * P.bl) * frcp(0.1 * rmax));
}
ART_HD double bl_dw(const DevParams& P, double r, double rmax) { return bl_w(P, r, rmax) * (-1.5 * frcp(r) - frcp(0.1 * rmax)); }
// ωp² with the boundary layer from the Goldreich–Julian ωp0² (r ≥ rNS)
ART_HD double bl_wp2(const DevParams& P, double Pp0, double r, double rmax) {
  const double w = (Pp0 > 0.0 ? fsqrt(Pp0) : 0.0) + bl_w(P, r, rmax);
  return w * w;
}

struct Met {
  double gtt, grr, dgtt, dgrr;  // g^tt, g^rr and their r-derivatives
};

// g_schwartz (RT.jl:455-501) with d/dr.  rs = 2GM/c²; R = interior switch radius.
ART_HD Met metric(double r, double rs, double R) {
  Met m;
  if (r > R) {
    double A = 1.0 - rs / r;
    m.grr = A;
    m.gtt = -1.0 / A;
    m.dgrr = rs / (r * r);
    m.dgtt = m.dgrr / (A * A);
  } else {
    // rs_e = rs (r/R)^3 ;  a = rs_e / R ;  b = r² rs_e / R³        (RT.jl:463, 496-497)
    double q = r / R;
    double q3 = q * q * q;
    double a = rs * q3 / R;
    double b = a * q * q;
    double sa = sqrt(1.0 - a), sb = sqrt(1.0 - b);
    double den = 3.0 * sa - sb;
    m.grr = 1.0 - b;
    m.gtt = -4.0 / (den * den);
    m.dgrr = -5.0 * b / r;
    double dden = -4.5 * a / (r * sa) + 2.5 * b / (r * sb);
    m.dgtt = 8.0 * dden / (den * den * den);
  }
  return m;
}

// Angular structure of the rotating dipole and its θ/ψ partials.
struct Fld {
  double c, s;                 // cosθ, sinθ
  double br, bt, bp, bz;       // b_r, b_θ, b_φ, b_z
  double br_t, bt_t, bz_t;     // ∂θ   (∂θ b_φ = 0)
  double br_p, bt_p, bp_p, bz_p;  // ∂ψ
  double D, D_t, D_p;
};

// core: from (cosθ, sinθ, cosψ, sinψ) — the sampler gets these straight from Cartesian
// coordinates without any trigonometric call.
template <bool DERIV>
ART_HD Fld field_cs(const DevParams& P, double c, double s, double cp, double sp) {
  Fld f;
  f.c = c;
  f.s = s;
  const double cm = P.cm, sm = P.sm;
  double a = cm * f.c + sm * f.s * cp;  // m̂·r̂
  f.br = 2.0 * a;
  f.bt = cm * f.s - sm * f.c * cp;
  f.bp = sm * sp;
  f.bz = f.br * f.c - f.bt * f.s;
  f.D = f.br * f.br + f.bt * f.bt + f.bp * f.bp;
  if (DERIV) {
    f.br_t = -2.0 * f.bt;            // ∂θ(2(cm c + sm s cψ)) = 2(−cm s + sm c cψ) = −2 b_θ
    f.bt_t = a;                      // = b_r / 2
    f.bz_t = f.br_t * f.c - f.br * f.s - f.bt_t * f.s - f.bt * f.c;
    f.br_p = -2.0 * sm * f.s * sp;
    f.bt_p = sm * f.c * sp;
    f.bp_p = sm * cp;
    f.bz_p = f.br_p * f.c - f.bt_p * f.s;
    f.D_t = 2.0 * (f.br * f.br_t + f.bt * f.bt_t);
    f.D_p = 2.0 * (f.br * f.br_p + f.bt * f.bt_p + f.bp * f.bp_p);
  }
  return f;
}

template <bool DERIV>
ART_HD Fld field(const DevParams& P, double th, double psi) {
  double c, s, cp, sp;
  fsincos(th, &s, &c);
  fsincos(psi, &sp, &cp);
  return field_cs<DERIV>(P, c, s, cp, sp);
}

// Metric for the hot loops: as metric(), from 1/r, plus √g^rr and 1/√g^rr; the exterior branch has
// one rsqrt and no division.
struct MetHot {
  double gtt, grr, dgtt, dgrr, sq, isq;
};
ART_HD MetHot metric_hot(double r, double rinv, double rs, double R) {
  MetHot m;
  if (r > R) {
    const double A = 1.0 - rs * rinv;
    const double y = frsqrt(A);
    const double Ainv = y * y;
    m.grr = A;
    m.gtt = -Ainv;
    m.dgrr = rs * rinv * rinv;
    m.dgtt = m.dgrr * Ainv * Ainv;
    m.sq = A * y;
    m.isq = y;
  } else {
    // interior branch of metric() with the branch-free reciprocal / rsqrt (axions crossing the star: a few
    // lanes of a warp walk this path, and IEEE sqrt and division made it 12 % of the stage kernel's
    // instructions — profiles/r01_final_lines.txt)
    const double Rinv = frcp(R);
    const double q = r * Rinv, q2 = q * q;
    const double a = rs * Rinv * q2 * q;
    const double b = a * q2;
    const double ysa = frsqrt(1.0 - a), ysb = frsqrt(1.0 - b);
    const double sb = (1.0 - b) * ysb;
    const double dinv = frcp(3.0 * (1.0 - a) * ysa - sb);
    const double dinv2 = dinv * dinv;
    m.grr = 1.0 - b;
    m.gtt = -4.0 * dinv2;
    m.dgrr = -5.0 * b * rinv;
    m.dgtt = 8.0 * (2.5 * b * ysb - 4.5 * a * ysa) * rinv * dinv2 * dinv;
    m.sq = sb;
    m.isq = ysb;
  }
  return m;
}

// ---------------------------------------------------------------------------------------
// Photon right-hand side: func!, RT.jl:71-91, components 1..6 (du7 = (Ω·erg/c_km)·du6, see
// DESIGN.md §3.2 — H depends on t only through ψ = φ − Ωt, so ∂H/∂t = −Ω ∂H/∂φ).
// k = (k_r,k_θ,k_φ) = u[4:6]·erg, om = −u7, t = e^{ln t}, einv = 1/erg.
// Five reciprocals (r, sinθ, D, ω, via rsqrt: g^rr), no division.
// ---------------------------------------------------------------------------------------
// ddelta (bndry_lyr > 0 only, else null): d/d(ln t) of δ = u7 − [u7₀ + (Ω·erg/c_km)(u6 − u6₀)].  The reference differentiates
// the boundary-layer term in du7 ONLY (hamiltonian is called with bndry_lyr at RT.jl:88, without at RT.jl:84-85), so with
// ωp_b = ωp0 + B(r):  ∂ψ(ωp_b²) = (ωp_b/ωp0) ∂ψ(ωp0²),  W_b = ωp_b²/(Dω²)  and
//   dδ = −Ω t g^rr/ω · ½[(ωp_b/ωp0 − 1)·sg·∂ψb_z·X − (W_b − W) g^rr (2N ∂ψN − N² ∂ψD/D)]      (Melrose form).
ART_HD void rhs_photon(const DevParams& P, double r, double th, double ph, double kr, double kt,
                       double kp, double om, double t, double einv, double du[6], double* ddelta = nullptr) {
  if (r <= P.rNS101) {  // RT.jl:86
    for (int i = 0; i < 6; ++i) du[i] = 0.0;
    if (ddelta) *ddelta = 0.0;
    return;
  }
  const double rinv = frcp(r);
  const MetHot m = metric_hot(r, rinv, P.rs_ray, P.r_int);
  const Fld f = field<true>(P, th, ph - P.Omega * t);
  const double sinv = frcp(f.s);
  const double rinv2 = rinv * rinv, rinv3 = rinv2 * rinv;
  const double q3 = P.Q * rinv3;
  const double Pp = q3 * fabs(f.bz);
  const double sg = signbit(f.bz) ? -q3 : q3;  // ForwardDiff abs: signbit ? −d : d
  const double sq = m.sq;
  const double ominv = frcp(om);
  const double om2inv = ominv * ominv;
  const double n_r = sq * f.br, n_t = f.bt * rinv, n_p = f.bp * rinv * sinv;
  const double N = kr * n_r + kt * n_t + kp * n_p;
  const double Dinv = frcp(f.D);
  double W = Pp * Dinv * om2inv;  // ωp² / (D ω²)
  double X = 1.0 - m.grr * N * N * Dinv * om2inv;
  if (P.isotropic) {  // hamiltonian iso branch, RT.jl:542-544
    W = 0.0;
    X = 1.0;
  }
  const double gN = W * m.grr * N;
  const double s2inv = rinv2 * sinv * sinv;
  const double kp2 = kp * kp * s2inv;  // k_φ² /(r² s²)
  const double kt2 = kt * kt * rinv2;
  // ∂H/∂k
  const double Hk_r = m.grr * kr - gN * n_r;
  const double Hk_t = kt * rinv2 - gN * n_t;
  const double Hk_p = kp * s2inv - gN * n_p;
  // ∂N/∂x
  const double N_r = (0.5 * m.dgrr * m.isq) * kr * f.br - (kt * n_t + kp * n_p) * rinv;
  const double N_t = kr * sq * f.br_t + kt * f.bt_t * rinv - kp * n_p * f.c * sinv;
  const double N_p = kr * sq * f.br_p + kt * f.bt_p * rinv + kp * f.bp_p * rinv * sinv;
  const double fac = C_KM * t * m.grr * ominv;
  const double face = fac * einv;
  if (!P.melrose && !P.isotropic) {
    // hamiltonian's "original form" (melrose = false, RT.jl:546-548) with cosθ_B from Ctheta_B_sphere
    // (RT.jl:957-971):  H = ½[g^tt ω² + K² + G],  G = ωp² Wm (K² − κ)/(Wm K² − ωp² κ),
    //   K² = g^ij k_i k_j,  κ = N²/D (= K² cos²θ_B),  Wm = ω²/g^rr.
    // Chain rule through (ωp², κ, K², Wm) with the same ∂N/∂x, ∂D/∂x, ∂b_z/∂x as the Melrose form:
    //   G_P = Wm² K² (K² − κ)/den²,  G_κ = ωp² Wm K² (ωp² − Wm)/den²,
    //   G_K = ωp² Wm κ (Wm − ωp²)/den²,  G_W = −ωp⁴ κ (K² − κ)/den²,   den = Wm K² − ωp² κ.
    const double K2 = m.grr * kr * kr + kt2 + kp2;
    const double kap = N * N * Dinv;
    const double gri = m.isq * m.isq;  // 1/g^rr
    const double Wm = om * om * gri;
    const double dinv = frcp(Wm * K2 - Pp * kap);
    const double d2 = dinv * dinv;
    const double cP = Wm * Wm * K2 * (K2 - kap) * d2;
    const double ck = Pp * Wm * K2 * (Pp - Wm) * d2;
    const double cK1 = 1.0 + Pp * Wm * kap * (Wm - Pp) * d2;
    const double cW = -Pp * Pp * kap * (K2 - kap) * d2;
    const double cN = ck * N * Dinv;  // ½ G_κ ∂κ/∂N
    const double ckk = ck * kap * Dinv;
    const double Hr = 0.5 * (m.dgtt * om * om + cK1 * (m.dgrr * kr * kr - 2.0 * (kt2 + kp2) * rinv) - 3.0 * Pp * rinv * cP -
                             cW * om * om * m.dgrr * gri * gri) + cN * N_r;
    const double Ht = 0.5 * (-2.0 * f.c * sinv * kp2 * cK1 + sg * f.bz_t * cP - ckk * f.D_t) + cN * N_t;
    const double Hp = 0.5 * (sg * f.bz_p * cP - ckk * f.D_p) + cN * N_p;
    du[0] = (cK1 * m.grr * kr + cN * n_r) * fac;
    du[1] = (cK1 * kt * rinv2 + cN * n_t) * fac;
    du[2] = (cK1 * kp * s2inv + cN * n_p) * fac;
    du[3] = -Hr * face;
    du[4] = -Ht * face;
    du[5] = -Hp * face;
    return;
  }
  // ∂H/∂x
  const double H_r = 0.5 * (m.dgtt * om * om + m.dgrr * kr * kr - 2.0 * (kt2 + kp2) * rinv - 3.0 * Pp * rinv * X -
                            W * (m.dgrr * N * N + 2.0 * m.grr * N * N_r));
  const double H_t = 0.5 * (-2.0 * f.c * sinv * kp2 + sg * f.bz_t * X -
                            W * m.grr * (2.0 * N * N_t - N * N * f.D_t * Dinv));
  const double H_p = 0.5 * (sg * f.bz_p * X - W * m.grr * (2.0 * N * N_p - N * N * f.D_p * Dinv));
  du[0] = Hk_r * fac;
  du[1] = Hk_t * fac;
  du[2] = Hk_p * fac;
  du[3] = -H_r * face;
  du[4] = -H_t * face;
  du[5] = -H_p * face;
  if (ddelta) {
#if defined(__CUDA_ARCH__)
    const double w0i = Pp > 0.0 ? frsqrt(Pp) : INFINITY, w0 = Pp > 0.0 ? Pp * w0i : 0.0;  // (on the null surface: √0 and 1/0, as below)
    const double wb = w0 + bl_w(P, r, P.bl_rmax);                   // (r > 1.01 rNS here)
    const double ratio = wb * w0i - 1.0;
#else
    const double w0 = sqrt(Pp), wb = w0 + bl_w(P, r, P.bl_rmax);  // (r > 1.01 rNS here)
    const double ratio = wb / w0 - 1.0;
#endif
    // (ratio = ωp_b/ωp0 − 1: ∞ on the null surface, like the reference's √|b_z|)
    double dW = (wb * wb - Pp) * Dinv * om2inv;                     // W_b − W
    double Xp = X;
    if (P.isotropic) {
      dW = 0.0;
      Xp = 1.0;
    }
    const double dHp = 0.5 * (ratio * sg * f.bz_p * Xp - dW * m.grr * (2.0 * N * N_p - N * N * f.D_p * Dinv));
    *ddelta = -P.Omega * t * m.grr * ominv * dHp;
  }
}

// Axion right-hand side: func_axion!, RT.jl:95-123 (geodesic; interior metric inside the star).
ART_HD void rhs_axion(const DevParams& P, double r, double th, double kr, double kt, double kp,
                      double t, double erg, double einv, double du[6]) {
  const double rinv = frcp(r);
  const MetHot m = metric_hot(r, rinv, P.rs_ray, P.r_int);
  double s, c;
  fsincos(th, &s, &c);
  const double sinv = frcp(s);
  const double rinv2 = rinv * rinv;
  const double s2inv = rinv2 * sinv * sinv;
  const double kp2 = kp * kp * s2inv;
  const double kt2 = kt * kt * rinv2;
  const double H_r = 0.5 * (m.dgtt * erg * erg + m.dgrr * kr * kr - 2.0 * (kt2 + kp2) * rinv);
  const double H_t = -c * sinv * kp2;
  const double fac = C_KM * t * m.grr * einv;
  const double face = fac * einv;
  du[0] = m.grr * kr * fac;
  du[1] = kt * rinv2 * fac;
  du[2] = kp * s2inv * fac;
  du[3] = -H_r * face;
  du[4] = -H_t * face;
  du[5] = 0.0;
}

// ---------------------------------------------------------------------------------------
// Thick-surface level-crossing condition (propagate: RT.jl:262-296; sampler: RT.jl:1554-1581).
// (w_r,w_θ,w_φ) = any positive multiple of the covariant direction; E = energy at infinity.
//   NrmSq = (−E² g^tt − m_a²)/(g^ij w_i w_j);  k∥² = NrmSq N(w)²/D;  g^tt E² + NrmSq·(w·w) = −m_a²
//   ⇒ Ham = ½(−m_a² + ωp² (1 − g^rr k∥²/E²)) / E²
// zero_in: ωp = 0 for r ≤ rNS (GJ_Model_ωp_vecSPH default, RT.jl:1165; the sampler's Cartesian
// variant GJ_Model_ωp_vec has no such switch).
// ---------------------------------------------------------------------------------------
ART_HD double thick_condition_f(const DevParams& P, double r, const Fld& f, double wr, double wt,
                                double wp, double E, double rs, bool zero_in) {
  const Met m = metric(r, rs, P.r_int);
  const double rinv = 1.0 / r, sinv = 1.0 / f.s;
  double Pp = P.Q * rinv * rinv * rinv * fabs(f.bz);
  if (P.bl > 0.0 && r >= P.rNS) Pp = bl_wp2(P, Pp, r, P.bl_rmax);
  if (zero_in && r <= P.rNS) Pp = 0.0;
  const double wts = wt * rinv, wps = wp * rinv * sinv;
  const double ww = m.grr * wr * wr + wts * wts + wps * wps;
  const double E2 = E * E;
  const double NrmSq = (-E2 * m.gtt - P.ma2) / ww;
  // E below the local axion mass shell (only reachable in the reference's "flat" mode, where
  // g^tt = −1 and a photon that lost energy has E < m_a): the reference takes sqrt(NrmSq) and
  // Julia throws a DomainError there; this build reports NaN = "no crossing" (DESIGN.md §quirks).
  if (NrmSq < 0.0) return NAN;
  const double N = wr * sqrt(m.grr) * f.br + wts * f.bt + wps * f.bp;
  double kpar2 = NrmSq * N * N / f.D;
  if (P.isotropic) kpar2 = 0.0;
  return 0.5 * (Pp * (1.0 - m.grr * kpar2 / E2) - P.ma2) / E2;
}

ART_HD double thick_condition(const DevParams& P, double r, double th, double psi, double wr,
                              double wt, double wp, double E, double rs, bool zero_in) {
  const Fld f = field<false>(P, th, psi);
  return thick_condition_f(P, r, f, wr, wt, wp, E, rs, zero_in);
}

// Division-free form of the same condition for the ray tracer's crossing scan and root finder:
//   F = Q|b_z| (E² ww' D − g^rr NrmSq·ww N'²) − m_a² E² ww' D r³,   ww' = ww (r s)²,  N' = N r s,
// i.e. F = Ham · 2E² · (E² ww D) (r s)² r³ — a strictly positive multiple, so F has the sign, the zeros
// and the NaN set of Ham (same algebra as the sampler's condition, art_sampler.cuh).  Exterior metric:
// g^rr NrmSq·ww = E² − m_a² A.  One reciprocal, one square root, two sincos.
// SIGN_ONLY: the caller keeps nothing but the sign (crossing scan).  F = base·(Q|b_z| − m_a² r³) − Q|b_z|·aniso
// with base > 0 and aniso ≥ 0, so wherever the plasma frequency is below the axion mass in THIS direction
// (Q|b_z| < m_a² r³) the sign is negative whatever the momentum: return before the momentum-dependent half.
template <bool SIGN_ONLY = false>
ART_HD double thick_F(const DevParams& P, double r, double th, double psi, double wr, double wt,
                      double wp, double E, double rs, bool zero_in) {
  double c, s, cp, sp;
  fsincos(th, &s, &c);
  fsincos(psi, &sp, &cp);
  const Fld f = field_cs<false>(P, c, s, cp, sp);
  const double E2 = E * E;
  double grr, gn;
  if (r > P.r_int) {
    grr = 1.0 - rs * frcp(r);
    gn = E2 - P.ma2 * grr;
    if (gn < 0.0) return NAN;  // NrmSq < 0
  } else {
    const Met m = metric(r, rs, P.r_int);
    const double nrm = -E2 * m.gtt - P.ma2;
    if (nrm < 0.0) return NAN;
    grr = m.grr;
    gn = grr * nrm;
  }
  double Pq = P.Q * fabs(f.bz);
  if (P.bl > 0.0 && r >= P.rNS) {  // boundary layer: ωp = √(Q|b_z|/r³) + B(r), Pq = ωp² r³
    const double r3 = r * r * r;
    Pq = bl_wp2(P, Pq / r3, r, P.bl_rmax) * r3;
  }
  if (zero_in && r <= P.rNS) Pq = 0.0;
  const double mr3 = P.ma2 * (r * r * r);
  if (SIGN_ONLY && Pq < mr3) return -1.0;
  const double sq = fsqrt(grr);
  const double a = wr * r * s, b = wt * s;
  const double wwp = grr * a * a + b * b + wp * wp;
  const double Np = a * sq * f.br + b * f.bt + wp * f.bp;
  const double base = E2 * wwp * f.D;
  const double aniso = P.isotropic ? 0.0 : gn * Np * Np;
  return Pq * (base - aniso) - mr3 * base;
}

// Cartesian → spherical
ART_HD void to_sph(const double x[3], double& r, double& th, double& ph) {
  r = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  th = acos(x[2] / r);
  ph = atan2(x[1], x[0]);
}

// The "celerity" map Cartesian vector → covariant spherical components shared by k_norm_Cart
// (RT.jl:656-661), k_sphere (1001-1007), propagate (205-211), sampler condition (1563-1566).
ART_HD void celerity(const double x[3], double r, double st, const double kv[3], double AA,
                     double w[3]) {
  const double dr = (x[0] * kv[0] + x[1] * kv[1] + x[2] * kv[2]) / r;
  const double rs_ = r * st;
  const double v1 = (x[2] * dr - r * kv[2]) / rs_;
  const double v2 = (-x[1] * kv[0] + x[0] * kv[1]) / rs_;
  const double ai = 1.0 / AA;
  w[0] = dr / sqrt(AA) * ai;
  w[1] = v1 * r * ai;
  w[2] = v2 * rs_ * ai;
}

// k_norm_Cart on its live branches (axion / photon with ax_fix=true, RT.jl:667,680): scale
// factor √NrmSq that puts khat on the axion mass shell.  Real stellar mass even when flat.
ART_HD double k_norm_factor(const DevParams& P, const double x[3], const double khat[3], double erg,
                            double r, double th) {
  double st = sin(th);
  double w[3];
  celerity(x, r, st, khat, 1.0 - P.rs_real / r, w);
  const Met m = metric(r, P.rs_real, P.r_int);
  const double wt = w[1] / r, wp = w[2] / (r * st);
  const double NrmSq = (-erg * erg * m.gtt - P.ma2) / (m.grr * w[0] * w[0] + wt * wt + wp * wp);
  return sqrt(NrmSq);
}

// propagate's initial-condition transform, RT.jl:179-216.  Out: u[0..5], u7.
ART_HD void ray_initial_state(const DevParams& P, const double x0[3], const double k0_in[3],
                              double erg, double dOmega, double u[6], double& u7) {
  double r, th, ph;
  to_sph(x0, r, th, ph);
  const double f = k_norm_factor(P, x0, k0_in, erg, r, th);
  const double k0[3] = {f * k0_in[0], f * k0_in[1], f * k0_in[2]};
  double w[3];
  celerity(x0, r, sin(th), k0, 1.0 - P.rs_ray / r, w);
  u[0] = r;
  u[1] = th;
  u[2] = ph;
  u[3] = w[0] / erg;
  u[4] = w[1] / erg;
  u[5] = w[2] / erg;
  u7 = erg * dOmega;
}

// Cartesian "proper velocity" of a state, as recorded at crossings (RT.jl:335-342, no interior
// mass rescale) and at the end of propagate (RT.jl:393-416, enclosed mass rescaled for r < rNS).
ART_HD void state_to_cartesian(const DevParams& P, const double u[6], double erg, bool rescale_inside,
                               double x[3], double k[3]) {
  double rs = P.rs_ray;
  const double r = u[0];
  if (rescale_inside && r < P.rNS) rs *= r * r * r / (P.rNS * P.rNS * P.rNS);
  const double om = 1.0 - rs / r;
  double st, ct, sp, cp;
  sincos(u[1], &st, &ct);
  sincos(u[2], &sp, &cp);
  const double v0 = u[3] * erg * sqrt(om) * om;
  const double v1 = u[4] * erg / r * om;
  const double v2 = u[5] * erg / (r * st) * om;
  x[0] = r * st * cp;
  x[1] = r * st * sp;
  x[2] = r * ct;
  const double vt = st * v0 + ct * v1;
  k[0] = cp * vt - sp * v2;
  k[1] = sp * vt + cp * v2;
  k[2] = ct * v0 - st * v1;
}

// ---------------------------------------------------------------------------------------
// Per-sample scalars of MainRunner.jl:498-557.
// ---------------------------------------------------------------------------------------
struct SampleWeights {
  double k_init[3], cos_w, jac_gr, sln_prob, erg_inf, vel_eng;
};

ART_HD SampleWeights sample_weights(const DevParams& P, const double x[3], const double vloc[3],
                                    const double vinf[3]) {
  SampleWeights W;
  double r, th, ph;
  to_sph(x, r, th, ph);
  const double vI2 = vinf[0] * vinf[0] + vinf[1] * vinf[1] + vinf[2] * vinf[2];
  W.vel_eng = vI2 / (C_KM * C_KM) * 0.5;                                  // MR.jl:522 (sic: /c twice)
  W.erg_inf = P.ma * sqrt(1.0 + vI2 / (1.0 - vI2));                       // MR.jl:523-524
  const double fk = k_norm_factor(P, x, vloc, W.erg_inf, r, th);          // MR.jl:529
  for (int i = 0; i < 3; ++i) W.k_init[i] = fk * vloc[i];
  // k_sphere (MR.jl:533): flat ⇒ Mass_NS = 0 inside that function only
  const Fld f = field<true>(P, th, ph - P.Omega * P.t0);
  double ks[3];
  celerity(x, r, f.s, W.k_init, 1.0 - (P.flat ? 0.0 : P.rs_real) / r, ks);
  const Met m = metric(r, P.rs_real, P.r_int);
  const double rinv = 1.0 / r, sinv = 1.0 / f.s, rinv2 = rinv * rinv;
  const double gth = rinv2, gpp = rinv2 * sinv * sinv;
  // ---- cos_w: direction of ∇ω(x; k) for omega_function (RT.jl:558-589, 1370-1375).  Only the
  // direction enters, so the outer sqrt and the 1/√2 drop out: ∇ω ∝ ∇F,
  //   F = K + Pp + √S,  S = K² + 2 K Pp − 4 k∥² Pp + Pp²,  K = g^ij k_i k_j,  k∥² = N²/D.
  const double q3 = P.Q * rinv * rinv2;
  const double Pp = q3 * fabs(f.bz);
  const double sg = signbit(f.bz) ? -q3 : q3;
  const double P_x[3] = {-3.0 * Pp * rinv, sg * f.bz_t, sg * f.bz_p};
  const double K = m.grr * ks[0] * ks[0] + gth * ks[1] * ks[1] + gpp * ks[2] * ks[2];
  const double K_x[3] = {m.dgrr * ks[0] * ks[0] - 2.0 * rinv * (gth * ks[1] * ks[1] + gpp * ks[2] * ks[2]),
                         -2.0 * f.c * sinv * gpp * ks[2] * ks[2], 0.0};
  const double sq = sqrt(m.grr);
  const double n_t = f.bt * rinv, n_p = f.bp * rinv * sinv;
  const double N = ks[0] * sq * f.br + ks[1] * n_t + ks[2] * n_p;
  const double N_x[3] = {(0.5 * m.dgrr / sq) * ks[0] * f.br - (ks[1] * n_t + ks[2] * n_p) * rinv,
                         ks[0] * sq * f.br_t + ks[1] * f.bt_t * rinv - ks[2] * n_p * f.c * sinv,
                         ks[0] * sq * f.br_p + ks[1] * f.bt_p * rinv + ks[2] * f.bp_p * rinv * sinv};
  const double D_x[3] = {0.0, f.D_t, f.D_p};
  double kp2 = N * N / f.D;
  double go[3];
  if (P.isotropic) {
    for (int j = 0; j < 3; ++j) go[j] = K_x[j] + P_x[j];  // ω² = K + ωp²
  } else {
    const double S = K * K + 2.0 * K * Pp - 4.0 * kp2 * Pp + Pp * Pp;
    const double hs = 0.5 / sqrt(S);
    for (int j = 0; j < 3; ++j) {
      const double kp2_x = (2.0 * N * N_x[j] - kp2 * D_x[j]) / f.D;
      const double S_x = 2.0 * K * K_x[j] + 2.0 * (K_x[j] * Pp + K * P_x[j]) - 4.0 * (kp2_x * Pp + kp2 * P_x[j]) +
                         2.0 * Pp * P_x[j];
      go[j] = K_x[j] + P_x[j] + hs * S_x;
    }
  }
  const double gnorm = sqrt(m.grr * go[0] * go[0] + gth * go[1] * go[1] + gpp * go[2] * go[2]);
  const double kmag = sqrt(K);
  W.cos_w = fabs(m.grr * ks[0] * go[0] + gth * ks[1] * go[1] + gpp * ks[2] * go[2]) / (kmag * gnorm);
  // ---- g_det (RT.jl:734-754): dr_θ = ∂θωp/∂rωp = ∂θ(ωp²)/∂r(ωp²); metric switch radius = rNS there
  if (P.flat) {
    W.jac_gr = 1.0;
  } else {
    const Met mg = metric(r, P.rs_real, P.rNS);
    double dr_th = P_x[1] / P_x[0], dr_p = P_x[2] / P_x[0];
    if (P.bl > 0.0 && r >= P.rNS) {  // ∂ωp with the boundary layer (RT.jl:743): only ∂r changes
      const double h2 = 0.5 / sqrt(Pp);
      const double wr = P_x[0] * h2 + bl_dw(P, r, P.bl_rmax);
      dr_th = P_x[1] * h2 / wr;
      dr_p = P_x[2] * h2 / wr;
    }
    const double s2 = f.s * f.s;
    W.jac_gr = sqrt(s2 * (mg.grr * r * r + dr_th * dr_th) + dr_p * dr_p) /
               sqrt(s2 * (r * r + dr_th * dr_th) + dr_p * dr_p);
  }
  // ---- MR.jl:552-557
  const double dense_extra = 2.0 / sqrt(PI) * (C_KM / 220.0) * sqrt(P.rs_real / r);
  const double redshift = sqrt(1.0 - P.rs_real / r);
  const double phaseS = dense_extra * (2.0 * PI * P.maxR * P.maxR) * (P.rho_DM * 1e9) / P.ma * W.jac_gr;
  W.sln_prob = W.cos_w * redshift * phaseS * 1e10 * C_KM * 1e5 * (double)P.n_max;
  return W;
}

// ---------------------------------------------------------------------------------------
// Conversion probability: get_Prob_nonAD (MR.jl:67-124) + conversion_prob (RT.jl:1405-1473),
// anisotropic branch, closed-form gradients.  Returns P_nonAD.  Field at t0 (the reference
// hard-wires 0), always the real stellar mass in the metric (global Mass_NS, MR.jl:75) and in
// the Christoffel symbols (RT.jl:503-512), Mass_NS → 0 only inside GJ_Model_Sphereical/k_sphere
// when flat.  The sign of B0 (−1 on the back-trace) cancels in every term that reaches Prob
// except through cosθ_B·∂cosθ_B (even) — kept as an argument for clarity of the interface.
// ---------------------------------------------------------------------------------------
ART_HD double conv_prob_nonAD(const DevParams& P, const double pos[3], const double kpos[3],
                              double erg_scaled, double bsign) {
  double r, th, ph;
  to_sph(pos, r, th, ph);
  const Fld f = field<true>(P, th, ph - P.Omega * P.t0);
  const Met m = metric(r, P.rs_real, P.r_int);                   // spatial_dot metric
  const double rsB = P.flat ? 0.0 : P.rs_real;
  const Met mB = metric(r, rsB, P.r_int);                        // metric inside GJ_Model_Sphereical
  const double rinv = 1.0 / r, sinv = 1.0 / f.s, rinv2 = rinv * rinv;
  const double gth = rinv2, gpp = rinv2 * sinv * sinv;
  const double c195 = 1.95e-2;
  const double Bn = bsign * P.Bn0 * rinv * rinv2;                // B0 (rNS/r)³ / 2
  const double sqB = sqrt(mB.grr);
  // covariant B (RT.jl:1298) and covariant k (k_sphere)
  const double Bs[3] = {Bn * f.br / sqB, Bn * f.bt * r, Bn * f.bp * r * f.s};
  double ks[3];
  celerity(pos, r, f.s, kpos, 1.0 - rsB / r, ks);
  const double BB = m.grr * Bs[0] * Bs[0] + gth * Bs[1] * Bs[1] + gpp * Bs[2] * Bs[2];
  const double Bmag = sqrt(BB) * c195;                                                   // MR.jl:78
  const double kmag = sqrt(m.grr * ks[0] * ks[0] + gth * ks[1] * ks[1] + gpp * ks[2] * ks[2]);  // MR.jl:79
  double ctB = (m.grr * Bs[0] * ks[0] + gth * Bs[1] * ks[1] + gpp * Bs[2] * ks[2]) * c195 / (kmag * Bmag);
  double stB = sin(acos(ctB));                                                           // MR.jl:81
  const double wE = erg_scaled / sqrt(1.0 - P.rs_real / r);                              // MR.jl:94
  const double q3 = P.Q * rinv * rinv2;
  double Pp0 = q3 * fabs(f.bz);
  if (r <= P.rNS) Pp0 = 0.0;                                                             // zeroIn=true, MR.jl:99
  const double wp0 = sqrt(Pp0);
  const bool blon = P.bl > 0.0 && r > P.rNS;
  // boundary layer: the value comes from get_Prob_nonAD's own ωp call, which does not pass Mass_a (default 1e-5, MR.jl:99);
  // the gradient from conversion_prob's, which does (RT.jl:1427)
  const double wp = blon ? wp0 + bl_w(P, r, P.bl_rmax_def) : wp0;
  const double Pp = wp * wp;
  const double vloc = sqrt(wE * wE - P.ma2) / wE;                                        // RT.jl:1410
  const double wE2 = wE * wE, wE4 = wE2 * wE2;
  // ∂μ ωp = ∂μ(ωp0²)/(2ωp0) (+ B'(r))
  const double sg = signbit(f.bz) ? -q3 : q3;
  const double h = 0.5 / wp0;
  const double dwp[3] = {-3.0 * Pp0 * rinv * h + (blon ? bl_dw(P, r, P.bl_rmax) : 0.0), sg * f.bz_t * h, sg * f.bz_p * h};  // RT.jl:1427
  // ∂μ |B|·1.95e-2,  |B| = |Bn| √D                                                        RT.jl:1429
  const double sD = sqrt(f.D);
  const double aBn = fabs(Bn) * c195;
  const double dB[3] = {-3.0 * aBn * sD * rinv, aBn * 0.5 * f.D_t / sD, aBn * 0.5 * f.D_p / sD};
  // term1 = Σ_i k_i ∂μ B^i,  B^r = Bn b_r √g^rr_B, B^θ = Bn b_θ / r, B^φ = Bn b_φ /(r s)   RT.jl:1432
  const double Bc = Bn * c195;
  const double dBr[3] = {Bc * f.br * (-3.0 * rinv * sqB + 0.5 * mB.dgrr / sqB), Bc * sqB * f.br_t, Bc * sqB * f.br_p};
  const double dBt[3] = {-4.0 * Bc * f.bt * rinv2, Bc * f.bt_t * rinv, Bc * f.bt_p * rinv};
  const double dBp[3] = {-4.0 * Bc * f.bp * rinv2 * sinv, -Bc * f.bp * f.c * rinv * sinv * sinv, Bc * f.bp_p * rinv * sinv};
  // Christoffel terms (Cristoffel, RT.jl:503-527: real mass, ignores flat)                 RT.jl:1434-1436
  const double GM = 0.5 * P.rs_real;
  const double G_rrr = -GM / (r * (r - 2.0 * GM));
  const double G_rtt = -(r - 2.0 * GM);
  const double G_rpp = G_rtt * f.s * f.s;
  const double G_1r = rinv;  // G_trt = G_prp = G_ttr = G_ppr
  const double G_tpp = -f.s * f.c;
  const double G_cot = f.c * sinv;  // G_ptp = G_ppt
  const double Bu_r = m.grr * Bs[0] * c195, Bu_t = gth * Bs[1] * c195, Bu_p = gpp * Bs[2] * c195;
  const double t2[3] = {ks[0] * Bu_r * G_rrr + ks[1] * G_1r * Bu_t + ks[2] * G_1r * Bu_p,
                        ks[0] * Bu_t * G_rtt + ks[2] * G_cot * Bu_p + ks[1] * Bu_r * G_1r,
                        ks[0] * Bu_p * G_rpp + ks[1] * G_tpp * Bu_p + ks[2] * G_1r * Bu_r + ks[2] * G_cot * Bu_t};
  if (P.isotropic) {  // RT.jl:1420-1422 with kmag fixed: ω = √(kmag² + ωp²)
    ctB = 0.0;
    stB = 1.0;
  }
  const double wp2 = Pp, wp4 = Pp * Pp;
  const double preF = wp / fabs(wE4 * wE + ctB * ctB * wE * (wp4 - 2.0 * wp2 * wE2));     // RT.jl:1448
  double vhat = 0.0;
  const double gk[3] = {m.grr * ks[0] / kmag, gth * ks[1] / kmag, gpp * ks[2] / kmag};
  for (int j = 0; j < 3; ++j) {
    const double term1 = ks[0] * dBr[j] + ks[1] * dBt[j] + ks[2] * dBp[j];
    const double dct = (term1 + t2[j]) / (kmag * Bmag) - ctB * dB[j] / Bmag;             // RT.jl:1438
    double dE = preF * (wE4 * stB * stB * dwp[j] - wE2 * ctB * wp * (wE2 - wp2) * dct);  // RT.jl:1449
    // (isotropic: ∇ω of omega_function, which is called WITHOUT bndry_lyr, RT.jl:1420-1422)
    if (P.isotropic) dE = wp0 * (j == 0 ? -3.0 * Pp0 * rinv * h : dwp[j]) / sqrt(kmag * kmag + Pp0);
    vhat += gk[j] * dE;                                                                  // RT.jl:1460
  }
  const double prefactor = wE4 * stB * stB / (ctB * ctB * wp2 * (wp2 - 2.0 * wE2) + wE4);   // RT.jl:1467
  const double gB = P.Ax_g * 1e-9 * Bmag;
  return PI / 2.0 * prefactor * gB * gB / (fabs(vhat) * vloc * C_KM * HBAR);               // RT.jl:1468
}

}  // namespace art
