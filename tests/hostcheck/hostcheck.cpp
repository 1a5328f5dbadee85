#include <cstdio>
#include <cstdlib>
// tests/hostcheck/hostcheck.cpp — TEST INFRASTRUCTURE: compiles the product's device headers
// (adiabatic_raytracer_b200/csrc/*.cuh, all `__host__ __device__`) for the CPU so that the
// `-m "not gpu"` suite can compare the kernel arithmetic with the oracle without a GPU.
// It is NOT part of the shipped library and is never loaded by the package.
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../adiabatic_raytracer_b200/csrc/art_host.hpp"
#include "../../adiabatic_raytracer_b200/csrc/art_sampler.cuh"
#include "../../adiabatic_raytracer_b200/csrc/art_tracer.cuh"
#include "../../adiabatic_raytracer_b200/csrc/art_tree.cuh"

using namespace art;

static int dev(const art_params* p, DevParams& d) {
  std::string e;
  return make_dev_params(*p, d, e);
}

static void trace_one(const DevParams& P, const RayIn& in, double* rec, int cap, RayOut& o) {
  Ray R;
  double kbuf[K_ROWS + K_PARK + K_XROWS], sbuf[SEG_MAX];
  KStore K{kbuf, 1};
  SegStore D{sbuf, 1};
  R.init(P, in, rec, cap);
  while (!R.done) R.step(P, K, D);
  R.finish(P, o);
}

extern "C" {

// branch-free helpers of art_physics.cuh (portable parts: fsincos, fexp; thick_F vs thick_condition)
int hc_fastmath(int64_t n, const double* x, double* sn, double* cs, double* ex) {
  for (int64_t i = 0; i < n; ++i) {
    fsincos(x[i], sn + i, cs + i);
    ex[i] = fexp(x[i]);
  }
  return 0;
}

int hc_flog(int64_t n, const double* x, double* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = flog(x[i]);
  return 0;
}

// sign-equivalent division-free condition (same argument layout as hc_condition)
int hc_condition_F(const art_params* p, int64_t n, const double* in, double* out) {
  DevParams P;
  int rc = dev(p, P);
  if (rc) return rc;
  for (int64_t i = 0; i < n; ++i) {
    const double* a = in + 8 * i;
    out[i] = thick_F(P, a[0], a[1], a[2] - P.Omega * exp(a[7]), a[3], a[4], a[5], a[6], P.rs_ray, true);
  }
  return 0;
}

int hc_max_radius(const art_params* p, double* maxR, double* Rcut) {
  DevParams d;
  int rc = dev(p, d);
  *maxR = d.maxR;
  *Rcut = d.Rcut;
  return rc;
}

int hc_rhs(const art_params* p, int64_t n, const int32_t* species, const double* u, const double* erg,
           const double* lnt, double* du) {
  DevParams P;
  int rc = dev(p, P);
  if (rc) return rc;
  for (int64_t i = 0; i < n; ++i) {
    const double* v = u + 7 * i;
    double* o = du + 7 * i;
    double t = exp(lnt[i]);
    if (species[i] == SPECIES_PHOTON) {
      double dd = 0.0;
      rhs_photon(P, v[0], v[1], v[2], v[3] * erg[i], v[4] * erg[i], v[5] * erg[i], -v[6], t, 1.0 / erg[i], o, P.bl > 0.0 ? &dd : nullptr);
      o[6] = P.Omega * erg[i] / C_KM * o[5] + dd;
    } else {
      rhs_axion(P, v[0], v[1], v[3] * erg[i], v[4] * erg[i], v[5] * erg[i], t, erg[i], 1.0 / erg[i], o);
      o[6] = 0.0;
    }
  }
  return 0;
}

// thick condition in ray form: in = u[7], lnt
int hc_condition(const art_params* p, int64_t n, const double* in, double* out) {
  DevParams P;
  int rc = dev(p, P);
  if (rc) return rc;
  for (int64_t i = 0; i < n; ++i) {
    const double* a = in + 8 * i;
    out[i] = thick_condition(P, a[0], a[1], a[2] - P.Omega * exp(a[7]), a[3], a[4], a[5], a[6], P.rs_ray, true);
  }
  return 0;
}

int hc_sample_from_uniforms(const art_params* p, int64_t n, const double* u, double* x, double* v_loc,
                            double* v_inf, int32_t* accepted, int32_t* n_cross, int32_t* n_cond) {
  DevParams P;
  int rc = dev(p, P);
  if (rc) return rc;
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t i = 0; i < n; ++i) {
    SampleResult R;
    sample_attempt(P, u + 10 * i, R);
    accepted[i] = R.accepted;
    n_cross[i] = R.n_cross;
    if (n_cond) n_cond[i] = R.n_cond;
    for (int c = 0; c < 3; ++c) {
      x[c * n + i] = R.x[c];
      v_loc[c * n + i] = R.vloc[c];
      v_inf[c * n + i] = R.vinf[c];
    }
  }
  return 0;
}

// How the sampler's trips are spent (tuning aid): counts[mode at the start of the trip] += 1 for SCAN / ROOT / NUDGE / INIT / SKIP,
// counts[5] += condition evaluations, counts[6] += attempts that never enter the loop (the line misses the Rcut sphere),
// counts[7] / counts[8]: scanned steps that are pointwise dark / lit (a sharper momentum-free bound could pass them over).
int hc_sampler_trips(const art_params* p, int64_t n, const double* u, long long* counts) {
  DevParams P;
  int rc = dev(p, P);
  if (rc) return rc;
  for (int k = 0; k < 12; ++k) counts[k] = 0;
  for (int64_t i = 0; i < n; ++i) {
    Sampler S;
    double taus[SAMP_KEEP];
    S.taus = taus;
    S.taus_stride = 1;
    S.begin(P, u + 10 * i);
    if (S.done) counts[6]++;
    while (!S.done) {
      counts[S.mode]++;
      if (S.mode == Sampler::SCAN) {
        bool dark = true, lit = true;
        for (int k = 0; k <= P.M_samp - 1; ++k) {
          const double t = S.tau + (double)k / (P.M_samp - 1) * S.h;
          double x[3];
          S.point(t, x);
          const double r = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
          const double mx = (P.mh[0] * x[0] + P.mh[1] * x[1] + P.mh[2] * x[2]) / r;
          const double bz = 3.0 * mx * x[2] / r - P.cm;
          const double Pq = P.Q * fabs(bz) / (r * r * r);
          dark = dark && Pq < P.ma2;
          lit = lit && r > P.rNS && Pq * (1.0 - P.rs_real / r) > S.E2;
        }
        if (dark) counts[7]++;
        if (lit) counts[8]++;
      }
      S.trip(P, true);
    }
    counts[5] += S.n_cond;
  }
  return 0;
}

int hc_weights(const art_params* p, int64_t n, const double* x, const double* v_loc, const double* v_inf,
               double* k_init, double* cos_w, double* jac_gr, double* sln_prob, double* erg_inf, double* vel_eng) {
  DevParams P;
  int rc = dev(p, P);
  if (rc) return rc;
  for (int64_t i = 0; i < n; ++i) {
    double xx[3] = {x[i], x[n + i], x[2 * n + i]}, vl[3] = {v_loc[i], v_loc[n + i], v_loc[2 * n + i]},
           vi[3] = {v_inf[i], v_inf[n + i], v_inf[2 * n + i]};
    SampleWeights W = sample_weights(P, xx, vl, vi);
    for (int c = 0; c < 3; ++c) k_init[c * n + i] = W.k_init[c];
    cos_w[i] = W.cos_w; jac_gr[i] = W.jac_gr; sln_prob[i] = W.sln_prob; erg_inf[i] = W.erg_inf; vel_eng[i] = W.vel_eng;
  }
  return 0;
}

int hc_conv_prob(const art_params* p, int64_t n, const double* x, const double* k, const double* erg_scaled,
                 int32_t b_sign, double* out) {
  DevParams P;
  int rc = dev(p, P);
  if (rc) return rc;
  for (int64_t i = 0; i < n; ++i) {
    double xx[3] = {x[i], x[n + i], x[2 * n + i]}, kk[3] = {k[i], k[n + i], k[2 * n + i]};
    out[i] = conv_prob_nonAD(P, xx, kk, erg_scaled[i], b_sign < 0 ? -1.0 : 1.0);
  }
  return 0;
}

int hc_propagate(const art_params* p, int64_t n, const int32_t* species, const double* x0, const double* k0,
                 const double* erg, const double* t_start, const double* dOmega, int32_t b_sign,
                 int32_t max_crossings, double* x_f, double* k_f, double* u7_f, int32_t* status, int32_t* n_cross,
                 double* cross_rec, int32_t cross_cap, int32_t* counters) {
  DevParams P;
  int rc = dev(p, P);
  if (rc) return rc;
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t i = 0; i < n; ++i) {
    RayIn in;
    for (int c = 0; c < 3; ++c) { in.x0[c] = x0[c * n + i]; in.k0[c] = k0[c * n + i]; }
    in.erg = erg[i]; in.t_start = t_start[i]; in.dOmega = dOmega[i]; in.max_cross = max_crossings;
    in.species = species[i];
    RayOut o;
    double* rec = cross_rec + (size_t)i * cross_cap * 8;
    trace_one(P, in, rec, cross_cap, o);
    for (int c = 0; c < 3; ++c) { x_f[c * n + i] = o.xf[c]; k_f[c * n + i] = o.kf[c]; }
    u7_f[i] = o.u7f; status[i] = o.status; n_cross[i] = o.n_cross;
    if (counters) { counters[4 * i] = o.n_acc; counters[4 * i + 1] = o.n_rej; counters[4 * i + 2] = o.n_rhs; counters[4 * i + 3] = o.n_cond; }
  }
  return 0;
}

// the whole event pipeline with the product's per-event functions, sequential generations
int hc_run(const art_params* p, uint64_t seed, int64_t first_event, int64_t n_events, double* rows, int64_t rows_cap,
           int64_t* n_rows, int64_t* f_inx, int64_t* tot_count) {
  DevParams P;
  int rc = dev(p, P);
  if (rc) return rc;
  const int ncols = p->save_mode > 1 ? ART_ROW_COLS_FULL : (p->save_mode > 0 ? ART_ROW_COLS_LONG : ART_ROW_COLS_SHORT);
  std::vector<long long> keys((size_t)rows_cap);
  unsigned long long nr = 0, nfin[2] = {0, 0};
  RowSink S{rows, keys.data(), &nr, (unsigned long long)rows_cap, ncols, nfin};
  int64_t failed = 0, tc = 0;
  std::vector<double> rec((size_t)p->cross_cap * 8);
  std::vector<TreeNode> stack((size_t)p->max_nodes + 4);
  for (int64_t e = 0; e < n_events; ++e) {
    const uint64_t ev = (uint64_t)(first_event + e);
    SampleResult R;
    for (uint32_t j = 0;; ++j) {
      double U[10];
      sampler_uniforms(seed, ev, j, U);
      sample_attempt(P, U, R);
      if (R.accepted) break;
      failed++;
    }
    SampleWeights W = sample_weights(P, R.x, R.vloc, R.vinf);
    EventData E;
    for (int c = 0; c < 3; ++c) { E.x[c] = R.x[c]; E.k_init[c] = W.k_init[c]; E.vinf[c] = R.vinf[c]; }
    E.erg_inf = W.erg_inf; E.vel_eng = W.vel_eng; E.sln_prob = W.sln_prob; E.cos_w = W.cos_w;
    RayIn in;
    for (int c = 0; c < 3; ++c) { in.x0[c] = E.x[c]; in.k0[c] = -E.k_init[c]; }
    in.erg = E.erg_inf; in.t_start = 0.0; in.dOmega = -1.0; in.max_cross = 100000; in.species = SPECIES_AXION;
    RayOut o;
    trace_one(P, in, rec.data(), p->cross_cap, o);
    tree_backward(P, E, o, rec.data(), p->cross_cap);
    TreeState T;
    tree_forward_start(P, E, T);
    unsigned long long row0 = nr;
    while (T.active) {
      for (int c = 0; c < 3; ++c) {
        in.x0[c] = (&T.cur.x)[c];
        in.k0[c] = (&T.cur.kx)[c];
      }
      in.t_start = T.cur.t; in.dOmega = T.cur.dw; in.max_cross = -1; in.species = T.cur.species;
      trace_one(P, in, rec.data(), p->cross_cap, o);
      tree_forward_update(P, E, T, stack.data(), o, rec.data(), seed, ev, e, S);
    }
    tc += T.tree_len;
    for (unsigned long long r = row0; r < nr && r < (unsigned long long)rows_cap; ++r)
      patch_row(rows + r * ncols, ncols, E, T.count, T.info);
  }
  if (nr > (unsigned long long)rows_cap) return ART_ERR_CAPACITY;
  *n_rows = (int64_t)nr;
  *f_inx = failed + (int64_t)nfin[0];
  *tot_count = tc;
  return 0;
}

}  // extern "C"

// step log of one ray (diagnostics): out[k*4 + (s_before, dt_before, accepted, n_cross_after)]
extern "C" int hc_trace_ray(const art_params* p, int32_t species, const double* x0, const double* k0, double erg,
                            double t_start, double dOmega, int32_t max_crossings, double* out, int cap) {
  DevParams P;
  int rc = dev(p, P);
  if (rc) return rc;
  RayIn in;
  for (int c = 0; c < 3; ++c) { in.x0[c] = x0[c]; in.k0[c] = k0[c]; }
  in.erg = erg; in.t_start = t_start; in.dOmega = dOmega; in.max_cross = max_crossings; in.species = species;
  std::vector<double> rec((size_t)p->cross_cap * 8);
  int n = 0;
  Ray R;
  double kbuf[K_ROWS + K_PARK + K_XROWS], sbuf[SEG_MAX];
  KStore K{kbuf, 1};
  SegStore D{sbuf, 1};
  R.init(P, in, rec.data(), p->cross_cap);
  while (!R.done) {
    double s = R.s, dt = R.dt;
    int acc = R.n_acc;
    R.step(P, K, D);
    if (n < cap) { out[4 * n] = s; out[4 * n + 1] = dt; out[4 * n + 2] = R.n_acc - acc; out[4 * n + 3] = R.n_cross; }
    n++;
  }
  return n;
}

// Certified-sign shortcuts of the wavefront stage kernel (Ray::step_all_far / step_sign / step_in_start_box) against the
// literal scan of the same accepted step: counts[0] accepted steps, [1] steps with a certified sign, [2] certified steps
// whose literal scan holds a point of another sign, [3] steps certified "inside the start box", [4] of those, steps with
// a scan point outside the box.
template <int IG>
static void certify_ray(const DevParams& P, const RayIn& in, double* rec, int cap, long long counts[12]) {
  Ray R;
  double kbuf[K_ROWS + K_PARK + K_XROWS], sbuf[SEG_MAX];
  KStore K{kbuf, 1};
  SegStore D{sbuf, 1};
  R.init(P, in, rec, cap);
  double s0[3] = {R.s0[0], R.s0[1], R.s0[2]};
  while (!R.done) {
    if (!R.template stages<IG>(P, K, D)) continue;
    ScanMasks m = {0ull, 0ull, 0ull};
    const int M = P.M_ray;
    for (int i = R.event_last ? 0 : 1; i < M; ++i) {
      const double g = seg_cond<IG>(P, D, seg_theta(P, i));
      if (g > 0.0) m.pos |= 1ull << i;
      else if (g < 0.0) m.neg |= 1ull << i;
      else if (g == 0.0) m.zero |= 1ull << i;
    }
    counts[0]++;
    const int cs = R.template step_all_far<IG>(P, K) ? -1 : R.template step_sign<IG>(P, K, exp(R.s_new));
    if (cs != 0) {
      counts[1]++;
      // every scanned point must carry the certified sign: no other sign, no zero, no NaN (a NaN end point would leave the
      // literal path with reference sign 0 instead of cs)
      unsigned long long want = 0ull;
      for (int i = R.event_last ? 0 : 1; i < M; ++i) want |= 1ull << i;
      const unsigned long long got = cs < 0 ? m.neg : m.pos;
      if (got != want) counts[2]++;
    }
    const bool inbox = R.n_cross == 0 && R.template step_in_start_box<IG>(K, s0, R.box_m);
    if (cs == 0 && !inbox) {  // what the scan kernel gets: uniformly negative / positive / mixed, and where
      const unsigned long long all = m.pos | m.neg | m.zero;
      if (m.neg == all && all) counts[5]++;
      else if (m.pos == all && all) counts[6]++;
      else counts[7]++;
      if (R.event_last) counts[10]++;
      {  // would a momentum-free bound do?  pointwise dark (Q|bz| < m_a² r³) / lit (Q|bz| g^rr > E² r³) at every scan point
        bool dark = true, lit = true;
        for (int i = 1; i < M; ++i) {
          double w[6];
          const double th = seg_theta(P, i);
          seg_state<IG>(D, th, w);
          const double E = D(Seg<IG>::U70) + D(Seg<IG>::FAC7) * (w[5] - D(Seg<IG>::U60));
          const double t = exp(D(Seg<IG>::S) + th * D(Seg<IG>::H));
          const double ps = w[2] - P.Omega * t;
          const double bz = fabs(P.cm * (3.0 * cos(w[1]) * cos(w[1]) - 1.0) + 3.0 * P.sm * sin(w[1]) * cos(w[1]) * cos(ps));
          const double r3 = w[0] * w[0] * w[0];
          dark = dark && P.Q * bz < P.ma2 * r3;
          lit = lit && P.Q * bz * (1.0 - P.rs_ray / w[0]) > E * E * r3;
        }
        if (dark || lit) {
          counts[11]++;
          if (R.template step_sign<IG>(P, K, exp(R.s_new), 0.0) != 0) counts[8]++;  // slack is what stops it
          else if (R.u[0] < P.rNS || R.un[0] < P.rNS) counts[9]++;
        }
      }
    }
    if (inbox) {
      counts[3]++;
      bool all_in = true;
      for (int i = 1; i < M; ++i) all_in = all_in && R.template in_start_box<IG>(P, D, seg_theta(P, i));
      if (!all_in) counts[4]++;
    }
    R.template finish_step<IG>(P, K, D, m);
  }
}

extern "C" int hc_certificates(const art_params* p, int64_t n, const int32_t* species, const double* x0, const double* k0,
                               const double* erg, int32_t max_crossings, long long* counts) {
  DevParams P;
  int rc = dev(p, P);
  if (rc) return rc;
  for (int k = 0; k < 12; ++k) counts[k] = 0;
  std::vector<double> rec((size_t)p->cross_cap * 8);
  for (int64_t i = 0; i < n; ++i) {
    RayIn in;
    for (int c = 0; c < 3; ++c) { in.x0[c] = x0[c * n + i]; in.k0[c] = k0[c * n + i]; }
    in.erg = erg[i]; in.t_start = 0.0; in.dOmega = -1.0; in.max_cross = max_crossings; in.species = species[i];
    if (P.integrator == IG_DP5) certify_ray<IG_DP5>(P, in, rec.data(), p->cross_cap, counts);
    else if (P.bl > 0.0) certify_ray<IG_VERN6_BL>(P, in, rec.data(), p->cross_cap, counts);
    else certify_ray<IG_VERN6>(P, in, rec.data(), p->cross_cap, counts);
  }
  return 0;
}

// saveMode 3: trajectory of one node sampled at `ntimes` saveat points (Ray::step_sampled, the code k_node_traj runs per thread).
// out: rows of 4 doubles (x y z ln t), returns the number of rows.
extern "C" int hc_node_traj(const art_params* p, int32_t species, const double* x0, const double* k0, double erg, double t_start,
                            double dOmega, int32_t max_crossings, int32_t ntimes, double* out, int32_t out_cap) {
  DevParams P;
  int rc = dev(p, P);
  if (rc) return -1;
  RayIn in;
  for (int c = 0; c < 3; ++c) { in.x0[c] = x0[c]; in.k0[c] = k0[c]; }
  in.erg = erg; in.t_start = t_start; in.dOmega = dOmega; in.max_cross = max_crossings; in.species = species;
  std::vector<double> rec(64 * 8);
  Ray R;
  double kbuf[K_ROWS + K_PARK + K_XROWS], sbuf[SEG_MAX];
  KStore K{kbuf, 1};
  SegStore D{sbuf, 1};
  R.init(P, in, rec.data(), 64);
  const double s0 = R.s, delta = (P.ln_tend - s0) / (double)(ntimes > 1 ? ntimes - 1 : 1);
  int n_out = 0, kq = 1;
  {
    double x[3], k[3];
    state_to_cartesian(P, R.u, R.erg, false, x, k);
    out[0] = x[0]; out[1] = x[1]; out[2] = x[2]; out[3] = s0;
    n_out = 1;
  }
  while (!R.done) {
    if (P.integrator == IG_DP5) R.step_sampled<IG_DP5>(P, K, D, s0, delta, ntimes, kq, out, n_out, out_cap);
    else if (P.bl > 0.0) R.step_sampled<IG_VERN6_BL>(P, K, D, s0, delta, ntimes, kq, out, n_out, out_cap);
    else R.step_sampled<IG_VERN6>(P, K, D, s0, delta, ntimes, kq, out, n_out, out_cap);
  }
  if (R.s < P.ln_tend && n_out < out_cap && out[4 * (n_out - 1) + 3] < R.s) {
    double x[3], k[3];
    state_to_cartesian(P, R.u, R.erg, false, x, k);
    double* q = out + 4 * (size_t)n_out++;
    q[0] = x[0]; q[1] = x[1]; q[2] = x[2]; q[3] = R.s;
  }
  return n_out;
}
