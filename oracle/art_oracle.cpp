// oracle/art_oracle.cpp — TEST INFRASTRUCTURE ONLY.
//
// CPU oracle of the hot path of SamWitte/Adiabatic_RayTracer: surface sampler
// (find_samples_new, RayTracer.jl:1480-1653), per-sample weights (MainRunner.jl:498-557),
// tree bookkeeping (get_tree, MainRunner.jl:126-352) and row assembly
// (MainRunner.jl:581-761), on top of physics.hpp / tracer.hpp.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library; the product (adiabatic_raytracer_b200/) never does.
// Parity status: UNPINNED against the real Julia code (it cannot run here); pinned against
// sympy/mpmath/scipy golden vectors generated from the reference's formulas (tests/golden/).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../include/art.h"
#include "philox.hpp"
#include "tracer.hpp"

#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

static NSParams ns_of(const art_params& p) {
  NSParams n;
  n.thm = p.theta_m; n.w = p.omega_pul; n.B0 = p.B0; n.rNS = p.rNS; n.Mass_NS = p.Mass_NS;
  n.Mass_a = p.Mass_a; n.Ax_g = p.Ax_g; n.bndry_lyr = p.bndry_lyr;
  n.flat = p.flat != 0; n.isotropic = p.isotropic != 0; n.melrose = p.melrose != 0;
  n.t0 = p.t0_sample; n.r_int = p.metric_interior_radius;
  return n;
}

static RayParams ray_of(const art_params& p, double erg, double B0, bool is_axion) {
  RayParams r;
  r.ns = ns_of(p);
  r.erg = erg; r.B0 = B0;
  r.reltol = p.reltol; r.abstol = p.ode_err; r.dtmin = p.dtmin;
  r.maxiters = p.maxiters; r.interp_points = p.interp_points_ray;
  r.integrator = p.integrator;
  r.is_axion = is_axion;
  return r;
}

// ---------------------------------------------------------------------------------------
// find_samples_new, RayTracer.jl:1480-1653, one attempt driven by 10 uniforms.
// ---------------------------------------------------------------------------------------
struct SampleOut {
  int numV = 0, n_cross = 0, n_cond = 0;
  double x[3] = {0, 0, 0}, rRND = 0, vloc[3] = {0, 0, 0}, vinf[3] = {0, 0, 0};
};

static SampleOut find_sample(const double U[10], double maxR, const art_params& p) {
  const NSParams N = ns_of(p);
  SampleOut S;
  double thi = std::acos(1.0 - 2.0 * U[0]);     // RT.jl:1486
  double phi = U[1] * 2.0 * M_PI;               // RT.jl:1487
  double thl = std::acos(1.0 - 2.0 * U[2]);     // RT.jl:1490
  double phl = U[3] * 2.0 * M_PI;               // RT.jl:1491
  double phR = U[4] * 2.0 * M_PI;               // RT.jl:1494
  double rRND = std::sqrt(U[5]) * maxR;         // RT.jl:1497
  double vall[3] = {std::sin(thi) * std::cos(phi), std::sin(thi) * std::sin(phi), std::cos(thi)};  // RT.jl:1521
  double vl[3] = {std::sin(thl) * std::cos(phl), std::sin(thl) * std::sin(phl), std::cos(thl)};    // RT.jl:1524
  double x1 = rRND * std::cos(phR), x2 = rRND * std::sin(phR);
  double xs[3] = {x1 * std::cos(-phi) * std::cos(-thi) + x2 * std::sin(-phi),
                  x2 * std::cos(-phi) - x1 * std::sin(-phi) * std::cos(-thi), x1 * std::sin(-thi)};  // RT.jl:1529
  double vI[3];
  for (int i = 0; i < 3; ++i) vI[i] = (220.0 + U[6 + i] * 1.0e-5) / std::sqrt(3.0);  // RT.jl:1531
  double vI_mag = std::sqrt(vI[0] * vI[0] + vI[1] * vI[1] + vI[2] * vI[2]);
  double gammaA = 1.0 / std::sqrt(1.0 - sq(vI_mag / c_km));                     // RT.jl:1537
  double erg_inf = N.Mass_a * std::sqrt(1.0 + sq(vI_mag / c_km * gammaA));      // RT.jl:1538
  for (int i = 0; i < 3; ++i) xs[i] += vall[i] * (-maxR * 1.1);                 // RT.jl:1542

  const double r_s0 = 2.0 * N.Mass_NS * GNew / (c_km * c_km);
  auto point = [&](double tau, double u[3]) {
    for (int i = 0; i < 3; ++i) u[i] = xs[i] + tau * vall[i];
  };
  auto cond = [&](double tau) {  // RT.jl:1547-1581
    double u[3], s[3];
    point(tau, u);
    if (!p.thick_surface) {  // RT.jl:1548-1551
      S.n_cond++;
      return std::log(omega_p_cart(u, N.t0, N.thm, N.w, N.B0, N.rNS, N.bndry_lyr, N.Mass_a)) - std::log(N.Mass_a);
    }
    cart_to_sph(u, s);
    double AA = 1.0 - r_s0 / s[0];
    if (s[0] < N.rNS) AA = 1.0;
    double w[3];
    celerity(u, s, vl, AA, w);
    double omP = omega_p_cart(u, N.t0, N.thm, N.w, N.B0, N.rNS, N.bndry_lyr, N.Mass_a);
    S.n_cond++;
    return thick_condition(s, w, erg_inf, omP, N.t0, N.Mass_NS, N, N.B0);
  };
  std::vector<double> taus;  // kept crossings, in march order
  auto affect = [&](double tau) {  // RT.jl:1585-1597
    double u[3], s[3];
    point(tau, u);
    cart_to_sph(u, s);
    double omP = omega_p_cart(u, N.t0, N.thm, N.w, N.B0, N.rNS, N.bndry_lyr, N.Mass_a);
    Metric<double> g = g_schwartz<double>(s[0], s[1], N.Mass_NS);
    double ergL = erg_inf / std::sqrt(g.grr);
    if (s[0] > N.rNS && ergL > omP) taus.push_back(tau);
  };

  // Euler march dt = 0.5 over (0, 2.2 maxR) with the ContinuousCallback scan, RT.jl:1599-1613
  const double T = 2.2 * maxR, dt0 = p.sampler_dt;
  const int M = p.interp_points_sampler;
  double tau = 0.0;
  bool event_last = false;
  double g_prev = cond(0.0);
  while (tau < T) {
    double dt = dt0;
    bool last = false;
    if (tau + dt >= T) { dt = T - tau; last = true; }
    double tau_new = last ? T : tau + dt;
    double prev = g_prev, th_left = 0.0;
    if (event_last) { th_left = 0.01; prev = cond(tau + th_left * dt); }
    int sgn_prev = (prev > 0) - (prev < 0);
    double th_a = th_left, g_a = prev, th_ev = -1.0, g_end = 0.0;
    for (int i = 1; i < M; ++i) {
      double th = (i == M - 1) ? 1.0 : (double)i * (1.0 / (double)(M - 1));
      double gv = cond(th >= 1.0 ? tau_new : tau + th * dt);
      if (i == M - 1) g_end = gv;
      int sg = (gv > 0) - (gv < 0);
      if (sgn_prev != 0 && sgn_prev * sg < 0) {
        int ne = 0;
        auto f = [&](double thx) { return cond(tau + thx * dt); };
        th_ev = find_root(f, th_a, g_a, th, gv, &ne);
        break;
      }
      if (sg == sgn_prev) { th_a = th; g_a = gv; }
    }
    if (th_ev < 0 && sgn_prev != 0 && g_end == 0.0) th_ev = 1.0;
    if (th_ev >= 0) {
      double te = th_ev >= 1.0 ? tau_new : tau + th_ev * dt;
      affect(te);
      tau = te;
      event_last = true;
    } else {
      tau = tau_new;
      g_prev = g_end;
      event_last = false;
    }
  }
  S.n_cross = (int)taus.size();
  if (taus.empty()) return S;  // RT.jl:1618-1620
  int randInx = 1 + (int)std::floor(U[9] * p.n_maxSample);  // RT.jl:1623
  if (randInx > p.n_maxSample) randInx = p.n_maxSample;
  if (S.n_cross < randInx) return S;  // RT.jl:1628-1632
  point(taus[randInx - 1], S.x);
  double rmag = std::sqrt(S.x[0] * S.x[0] + S.x[1] * S.x[1] + S.x[2] * S.x[2]);
  double vmag_loc = std::sqrt(vI_mag * vI_mag + 2.0 * GNew * N.Mass_NS / rmag) / c_km;  // RT.jl:1638
  S.numV = 1;
  S.rRND = rRND;
  for (int i = 0; i < 3; ++i) {
    S.vloc[i] = vl[i] * vmag_loc;
    S.vinf[i] = vI[i] / c_km;
  }
  return S;
}

// ---------------------------------------------------------------------------------------
// Per-sample scalars, MainRunner.jl:498-557.
// ---------------------------------------------------------------------------------------
struct Weights {
  double k_init[3], cos_w, jac_gr, sln_prob, erg_inf, vel_eng;
};

static Weights sample_weights(const double x[3], const double vloc[3], const double vinf[3],
                              double maxR, const art_params& p) {
  const NSParams N = ns_of(p);
  Weights W;
  double rmag = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  double vI_mag = std::sqrt(vinf[0] * vinf[0] + vinf[1] * vinf[1] + vinf[2] * vinf[2]);          // MR.jl:521
  W.vel_eng = (sq(vinf[0] / c_km) + sq(vinf[1] / c_km) + sq(vinf[2] / c_km)) / 2.0;             // MR.jl:522 (sic)
  double gammaA = 1.0 / std::sqrt(1.0 - vI_mag * vI_mag);                                        // MR.jl:523
  W.erg_inf = N.Mass_a * std::sqrt(1.0 + sq(vI_mag * gammaA));                                   // MR.jl:524
  k_norm_cart(x, vloc, W.erg_inf, N.Mass_NS, N.Mass_a, W.k_init);                                // MR.jl:529
  double ks[3];
  k_sphere(x, W.k_init, N.Mass_NS, N.flat, ks);                                                  // MR.jl:533
  W.cos_w = cos_w_of(x, ks, N);                                                                  // MR.jl:536
  double s[3];
  cart_to_sph(x, s);
  W.jac_gr = g_det(s, N);                                                                        // MR.jl:550
  double dense_extra = 2.0 / std::sqrt(M_PI) * (1.0 / (220.0 / c_km)) *
                       std::sqrt(2.0 * N.Mass_NS * GNew / (c_km * c_km) / rmag);                 // MR.jl:552
  double redshift = std::sqrt(1.0 - 2.0 * GNew * N.Mass_NS / rmag / (c_km * c_km));              // MR.jl:554
  double phaseS = dense_extra * (2.0 * M_PI * maxR * maxR) * (p.rho_DM * 1e9) / N.Mass_a * W.jac_gr;  // MR.jl:555
  W.sln_prob = std::fabs(W.cos_w) * redshift * phaseS * (1e5 * 1e5) * c_km * 1e5 * (double)p.n_maxSample;  // MR.jl:556
  return W;
}

// ---------------------------------------------------------------------------------------
// get_tree, MainRunner.jl:126-352.
// ---------------------------------------------------------------------------------------
struct Node {
  double x, y, z, kx, ky, kz, t, dw;
  int species;  // ART_SPECIES_*
  double prob, weight, parent_weight, prob_conv, prob_conv0;
  bool is_final = false;
  double xe[3] = {0, 0, 0}, ke[3] = {0, 0, 0}, u7e = 0;
  int status = 0;
  std::vector<Crossing> cr;
  std::vector<double> Pc;
};

static Node make_node(double x, double y, double z, double kx, double ky, double kz, double t,
                      double dw, int species, double prob, double weight, double parent_weight,
                      double prob_conv, double prob_conv0) {
  Node n;
  n.x = x; n.y = y; n.z = z; n.kx = kx; n.ky = ky; n.kz = kz; n.t = t; n.dw = dw;
  n.species = species; n.prob = prob; n.weight = weight; n.parent_weight = parent_weight;
  n.prob_conv = prob_conv; n.prob_conv0 = prob_conv0;
  return n;
}

struct Counters {
  int64_t n_rays = 0, n_rays_photon = 0, n_acc = 0, n_rej = 0, n_rhs = 0, n_cond = 0;
  int64_t n_status[5] = {0, 0, 0, 0, 0};
};

struct TreeOut {
  std::vector<Node> tree;
  int count = 0, info = 1;
};

static TreeOut get_tree(Node first, double erg_inf_ini, const art_params& p, double B0,
                        int num_cutoff, int splittings_cutoff, int MC_nodes, int max_nodes,
                        uint64_t seed, uint64_t event_id, uint32_t* rng_k, Counters* C) {
  const NSParams N = ns_of(p);
  TreeOut out;
  {  // MR.jl:132-137
    double pos[3] = {first.x, first.y, first.z}, kpos[3] = {first.kx, first.ky, first.kz};
    double Pn = prob_nonAD(pos, kpos, erg_inf_ini * std::fabs(first.dw), N, B0);
    first.prob = 1.0 - std::exp(-Pn);
  }
  std::vector<Node> events;
  events.push_back(first);
  double tot_prob = 0.0;
  int count = 0, count_main = 0, info = 1;
  const double dt0 = std::exp(p.ln_t_start);               // MR.jl:153
  const double ln_tend = std::log(1.0 / p.omega_pul);       // MR.jl:412
  while (!events.empty()) {
    count++;
    Node ev = events.back();
    events.pop_back();
    double pos0[3] = {ev.x, ev.y, ev.z}, k0[3] = {ev.kx, ev.ky, ev.kz};
    double ln_ts = std::log(std::max(ev.t, dt0));          // MR.jl:166
    bool is_axion = ev.species != ART_SPECIES_PHOTON;
    RayParams rp = ray_of(p, erg_inf_ini, B0, is_axion);
    RayResult R = propagate(pos0, k0, rp, ln_ts, ln_tend, splittings_cutoff, ev.dw);  // MR.jl:175-191
    if (C) {
      C->n_rays++; C->n_rays_photon += !is_axion; C->n_acc += R.n_acc; C->n_rej += R.n_rej;
      C->n_rhs += R.n_rhs; C->n_cond += R.n_cond; C->n_status[R.status]++;
    }
    for (int i = 0; i < 3; ++i) { ev.xe[i] = R.xf[i]; ev.ke[i] = R.kf[i]; }
    ev.u7e = R.u7f;
    ev.status = R.status;
    std::vector<Crossing> cr = R.cr;
    if (cr.empty()) {  // MR.jl:200-207
      count_main++;
      tot_prob += ev.weight;
      if (std::sqrt(ev.xe[0] * ev.xe[0] + ev.xe[1] * ev.xe[1] + ev.xe[2] * ev.xe[2]) > N.rNS * 1.1) ev.is_final = true;
    } else {
      bool rare = false;  // MR.jl:213-224
      for (auto& c : cr) rare = rare || std::fabs(c.kx) > 1 || std::fabs(c.ky) > 1 || std::fabs(c.kz) > 1;
      if (rare) {
        out.tree.push_back(ev);
        tot_prob += ev.weight;
        continue;
      }
      if (cr.size() > 1) {  // MR.jl:227-245
        const double epsabs = 1e-5;
        std::vector<char> flag(cr.size(), 1);
        bool any = false;
        for (size_t j = 0; j + 1 < cr.size(); ++j) {
          double d = std::sqrt(sq(cr[j + 1].x - cr[j].x) + sq(cr[j + 1].y - cr[j].y) + sq(cr[j + 1].z - cr[j].z));
          if (d < epsabs) any = true;
          flag[j] = d > epsabs;
        }
        if (any) {
          std::vector<Crossing> keep;
          for (size_t j = 0; j < cr.size(); ++j)
            if (flag[j]) keep.push_back(cr[j]);
          cr.swap(keep);
        }
      }
      ev.cr = cr;
      const size_t Nc = cr.size();
      std::vector<double> Prob(Nc);
      for (size_t j = 0; j < Nc; ++j) {  // MR.jl:265-267 (per-ray-correct)
        double pc[3] = {cr[j].x, cr[j].y, cr[j].z}, kc[3] = {cr[j].kx, cr[j].ky, cr[j].kz};
        Prob[j] = 1.0 - std::exp(-prob_nonAD(pc, kc, erg_inf_ini * std::fabs(cr[j].dw), N, B0));
      }
      ev.Pc = Prob;
      int new_species = ev.species == ART_SPECIES_PHOTON ? ART_SPECIES_AXION : ART_SPECIES_PHOTON;
      const Crossing& c0 = cr[0];
      if (splittings_cutoff <= 0) {
        if (count > MC_nodes) {  // MR.jl:281-292
          double r = tree_uniform(seed, event_id, (*rng_k)++);
          if (r < Prob[0])
            events.push_back(make_node(c0.x, c0.y, c0.z, c0.kx, c0.ky, c0.kz, c0.t, c0.dw, new_species,
                                       Prob[0], ev.weight, ev.weight, Prob[0], Prob[0]));
          else
            events.push_back(make_node(c0.x, c0.y, c0.z, c0.kx, c0.ky, c0.kz, c0.t, c0.dw, ev.species,
                                       1 - Prob[0], ev.weight, ev.weight, Prob[0], ev.prob_conv));
        } else {  // MR.jl:296-303
          events.push_back(make_node(c0.x, c0.y, c0.z, c0.kx, c0.ky, c0.kz, c0.t, c0.dw, new_species,
                                     Prob[0], Prob[0] * ev.weight, ev.weight, Prob[0], Prob[0]));
          events.push_back(make_node(c0.x, c0.y, c0.z, c0.kx, c0.ky, c0.kz, c0.t, c0.dw, ev.species,
                                     1 - Prob[0], (1 - Prob[0]) * ev.weight, ev.weight, Prob[0], ev.prob_conv));
        }
      } else {  // MR.jl:307-316
        for (size_t j = 0; j < Nc; ++j) {
          events.push_back(make_node(cr[j].x, cr[j].y, cr[j].z, cr[j].kx, cr[j].ky, cr[j].kz, cr[j].t,
                                     cr[j].dw, new_species, Prob[j], Prob[j] * ev.weight, ev.weight,
                                     Prob[0], Prob[0]));
          ev.weight = ev.weight * (1 - Prob[j]);
        }
        tot_prob += ev.weight;
      }
    }
    out.tree.push_back(ev);
    if (tot_prob >= 1 - p.prob_cutoff) { info = 2; break; }          // MR.jl:324-327
    if (num_cutoff <= 0 && splittings_cutoff > 0) break;             // MR.jl:328-330
    if (count_main >= num_cutoff) { info = 3; break; }               // MR.jl:331-334
    if (count > max_nodes) { info = 4; break; }                      // MR.jl:336-339
    std::stable_sort(events.begin(), events.end(),
                     [](const Node& a, const Node& b) { return a.weight < b.weight; });  // MR.jl:342
  }
  if (count > MC_nodes) info = -std::abs(info);  // MR.jl:346-348
  out.count = count;
  out.info = info;
  return out;
}

// ---------------------------------------------------------------------------------------
// One event of main_runner_tree's loop, MainRunner.jl:450-747.
// ---------------------------------------------------------------------------------------
struct EventOut {
  std::vector<double> rows;  // row-major, ncols each
  int64_t n_failed = 0, n_final_photon = 0, n_final_axion = 0, tot_count = 0, n_cond_sampler = 0;
  Counters C;
  double event_line[ART_EVENT_COLS];  // numbers of the event_* text file of saveMode > 1, MR.jl:593-610, 733-735
};

static void run_event(const art_params& p, double maxR, uint64_t seed, uint64_t e, int ncols, EventOut& E) {
  SampleOut S;
  for (uint32_t j = 0;; ++j) {  // MR.jl:463-494
    double U[10];
    sampler_uniforms(seed, e, j, U);
    S = find_sample(U, maxR, p);
    E.n_cond_sampler += S.n_cond;
    if (S.numV) break;
    E.n_failed++;
  }
  Weights W = sample_weights(S.x, S.vloc, S.vinf, maxR, p);
  uint32_t rng_k = 0;
  // backward, MR.jl:581-589
  Node parent = make_node(S.x[0], S.x[1], S.x[2], -W.k_init[0], -W.k_init[1], -W.k_init[2], 0.0, -1.0,
                          ART_SPECIES_AXION, 1.0, 1.0, -1.0, -1.0, -1.0);
  TreeOut back = get_tree(parent, W.erg_inf, p, -p.B0, 0, 100000, 5, 50, seed, e, &rng_k, &E.C);
  Node& nb = back.tree[0];
  int c_bck = back.count;
  double Prob_nonAD_0 = nb.prob;                    // MR.jl:626
  double samp_back_weight = nb.prob * nb.weight;    // MR.jl:630
  // forward, MR.jl:653-664
  Node par2 = make_node(S.x[0], S.x[1], S.x[2], W.k_init[0], W.k_init[1], W.k_init[2], 0.0, -1.0,
                        ART_SPECIES_PHOTON, 1.0, 1.0, -1.0, -1.0, -1.0);
  TreeOut fw = get_tree(par2, W.erg_inf, p, p.B0, p.num_cutoff, -1, p.MC_nodes, p.max_nodes, seed, e,
                        &rng_k, &E.C);
  E.tot_count += (int64_t)fw.tree.size();  // MR.jl:667
  for (Node& n : fw.tree) {  // MR.jl:670-729
    if (!n.is_final) continue;
    double absf = std::sqrt(n.ke[0] * n.ke[0] + n.ke[1] * n.ke[1] + n.ke[2] * n.ke[2]);
    double absfX = std::sqrt(n.xe[0] * n.xe[0] + n.xe[1] * n.xe[1] + n.xe[2] * n.xe[2]);
    double phf = std::atan2(n.ke[1], n.ke[0]), phfX = std::atan2(n.xe[1], n.xe[0]);
    double thf = std::acos(n.ke[2] / absf), thfX = std::acos(n.xe[2] / absfX);
    int id = n.species == ART_SPECIES_PHOTON ? 1 : 0;
    n.weight *= samp_back_weight;  // MR.jl:686
    double weight_tmp = n.weight;  // opticalDepth = 0, weightC = 1
    double dOm = n.u7e / p.Mass_a + W.vel_eng;  // MR.jl:709
    if (id == 1) E.n_final_photon++; else E.n_final_axion++;
    // columns 29, 30 (save_mode > 1): |k_f| and the node's time, the two numbers of the final_* text
    // file (MR.jl:689-701) that are not in the 29-column row
    double row[ART_ROW_COLS_FULL] = {(double)e, (double)id, thf, phf, thfX, phfX, absfX, W.sln_prob, weight_tmp,
                                     S.x[0], S.x[1], S.x[2], dOm,
                                     n.weight, 0.0, 1.0, W.k_init[0], W.k_init[1], W.k_init[2], W.cos_w,
                                     (double)fw.count, (double)fw.info, n.prob, n.prob_conv, n.prob_conv0,
                                     samp_back_weight, absfX, (double)c_bck, Prob_nonAD_0, absf, n.t};
    E.rows.insert(E.rows.end(), row, row + ncols);
  }
  // event_* line, MR.jl:593-610 (nb.x … nb.kz are the back-trace node's initial conditions: the sample
  // point and −k_init) and 733-735 (count; the wall-clock field is not reproduced)
  const double line[ART_EVENT_COLS] = {(double)e, S.vinf[0], S.vinf[1], S.vinf[2], W.sln_prob, S.x[0], S.x[1], S.x[2],
                                       W.k_init[0], W.k_init[1], W.k_init[2], (double)fw.count};
  std::memcpy(E.event_line, line, sizeof(line));
}

}  // namespace orc

// =========================================================================================
// C interface for ctypes (tests/, bench.py cpu_baseline).  Mirrors include/art.h argument
// conventions (component-major 3-vectors) so the parity tests feed identical buffers.
// =========================================================================================
using namespace orc;

extern "C" {

int orc_max_radius(const art_params* p, double* maxR) {
  *maxR = find_conversion_surface(p->Mass_a, p->fix_time, p->theta_m, p->omega_pul, p->B0, p->rNS);
  return *maxR < p->rNS ? ART_ERR_NO_SURFACE : ART_OK;
}

void orc_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32_10(ctr, key, out); }
void orc_sampler_uniforms(uint64_t seed, uint64_t event, uint32_t attempt, double* U) {
  sampler_uniforms(seed, event, attempt, U);
}

// pointwise probes: which = 0 ωp_sph(zeroIn=false) ; 1 metric (4 outs) ; 2 hamiltonian photon ;
// 3 thick condition (ray form) ; 4 omega_function
int orc_pointwise(const art_params* p, int which, int64_t n, const double* in, double* out) {
  NSParams N = ns_of(*p);
  for (int64_t i = 0; i < n; ++i) {
    if (which == 0) {
      const double* a = in + 4 * i;  // r th ph t
      out[i] = omega_p_sph<double>(a[0], a[1], a[2], a[3], N.thm, N.w, N.B0, N.rNS, false, N.bndry_lyr, N.Mass_a);
    } else if (which == 1) {
      const double* a = in + 2 * i;
      Metric<double> g = g_schwartz<double>(a[0], a[1], N.flat ? 0.0 : N.Mass_NS);
      out[4 * i] = g.gtt; out[4 * i + 1] = g.grr; out[4 * i + 2] = g.gthth; out[4 * i + 3] = g.gpp;
    } else if (which == 2) {
      const double* a = in + 8 * i;  // r th ph kr kth kph t omega
      double k[3] = {a[3], a[4], a[5]};
      out[i] = hamiltonian<double>(a[0], a[1], a[2], k, a[6], a[7], N.thm, N.w, N.B0, N.rNS,
                                   N.flat ? 0.0 : N.Mass_NS, N.Mass_a, N.isotropic, false, -1.0, N.melrose);
    } else if (which == 3) {
      const double* a = in + 8 * i;  // u[7], lnt
      RayParams rp = ray_of(*p, 1.0, p->B0, false);
      out[i] = ray_condition(rp, a, a[7]);
    } else if (which == 4) {
      const double* a = in + 7 * i;  // r th ph k1 k2 k3 t
      double k[3] = {a[3], a[4], a[5]};
      out[i] = omega_function<double>(a[0], a[1], a[2], k, a[6], N.thm, N.w, N.B0, N.rNS, N.Mass_NS, N.Mass_a, N.isotropic);
    } else {
      return ART_ERR_ARG;
    }
  }
  return ART_OK;
}

int orc_rhs(const art_params* p, int64_t n, const int32_t* species, const double* u, const double* erg,
            const double* lnt, int32_t b_sign, double* du) {
  for (int64_t i = 0; i < n; ++i) {
    RayParams rp = ray_of(*p, erg[i], b_sign < 0 ? -p->B0 : p->B0, species[i] != ART_SPECIES_PHOTON);
    rhs(rp, u + 7 * i, lnt[i], du + 7 * i);
  }
  return ART_OK;
}

int orc_sample_from_uniforms(const art_params* p, int64_t n, const double* u, double* x, double* v_loc,
                             double* v_inf, int32_t* accepted, int32_t* n_cross) {
  double maxR;
  int rc = orc_max_radius(p, &maxR);
  if (rc) return rc;
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t i = 0; i < n; ++i) {
    SampleOut S = find_sample(u + 10 * i, maxR, *p);
    accepted[i] = S.numV;
    n_cross[i] = S.n_cross;
    for (int c = 0; c < 3; ++c) {
      x[c * n + i] = S.x[c];
      v_loc[c * n + i] = S.vloc[c];
      v_inf[c * n + i] = S.vinf[c];
    }
  }
  return ART_OK;
}

int orc_sample(const art_params* p, uint64_t seed, int64_t first_event, int64_t n, double* x, double* v_loc,
               double* v_inf, int32_t* n_cross, int32_t* n_failed) {
  double maxR;
  int rc = orc_max_radius(p, &maxR);
  if (rc) return rc;
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t i = 0; i < n; ++i) {
    SampleOut S;
    int nf = 0;
    for (uint32_t j = 0;; ++j) {
      double U[10];
      sampler_uniforms(seed, (uint64_t)(first_event + i), j, U);
      S = find_sample(U, maxR, *p);
      if (S.numV) break;
      nf++;
    }
    n_failed[i] = nf;
    n_cross[i] = S.n_cross;
    for (int c = 0; c < 3; ++c) {
      x[c * n + i] = S.x[c];
      v_loc[c * n + i] = S.vloc[c];
      v_inf[c * n + i] = S.vinf[c];
    }
  }
  return ART_OK;
}

int orc_weights(const art_params* p, int64_t n, const double* x, const double* v_loc, const double* v_inf,
                double* k_init, double* cos_w, double* jac_gr, double* sln_prob, double* erg_inf,
                double* vel_eng) {
  double maxR;
  int rc = orc_max_radius(p, &maxR);
  if (rc) return rc;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    double xx[3] = {x[i], x[n + i], x[2 * n + i]}, vl[3] = {v_loc[i], v_loc[n + i], v_loc[2 * n + i]},
           vi[3] = {v_inf[i], v_inf[n + i], v_inf[2 * n + i]};
    Weights W = sample_weights(xx, vl, vi, maxR, *p);
    for (int c = 0; c < 3; ++c) k_init[c * n + i] = W.k_init[c];
    cos_w[i] = W.cos_w; jac_gr[i] = W.jac_gr; sln_prob[i] = W.sln_prob; erg_inf[i] = W.erg_inf; vel_eng[i] = W.vel_eng;
  }
  return ART_OK;
}

int orc_conv_prob(const art_params* p, int64_t n, const double* x, const double* k, const double* erg_scaled,
                  int32_t b_sign, double* P_nonAD) {
  NSParams N = ns_of(*p);
  double B0 = b_sign < 0 ? -p->B0 : p->B0;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    double xx[3] = {x[i], x[n + i], x[2 * n + i]}, kk[3] = {k[i], k[n + i], k[2 * n + i]};
    P_nonAD[i] = prob_nonAD(xx, kk, erg_scaled[i], N, B0);
  }
  return ART_OK;
}

int orc_propagate(const art_params* p, int64_t n, const int32_t* species, const double* x0, const double* k0,
                  const double* erg, const double* t_start, const double* dOmega, int32_t b_sign,
                  int32_t max_crossings, double* x_f, double* k_f, double* u7_f, int32_t* status,
                  int32_t* n_cross, double* cross_rec, int32_t cross_cap, int32_t* counters) {
  const double dt0 = std::exp(p->ln_t_start), ln_tend = std::log(1.0 / p->omega_pul);
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t i = 0; i < n; ++i) {
    double xx[3] = {x0[i], x0[n + i], x0[2 * n + i]}, kk[3] = {k0[i], k0[n + i], k0[2 * n + i]};
    RayParams rp = ray_of(*p, erg[i], b_sign < 0 ? -p->B0 : p->B0, species[i] != ART_SPECIES_PHOTON);
    RayResult R = propagate(xx, kk, rp, std::log(std::max(t_start[i], dt0)), ln_tend, max_crossings,
                            dOmega[i], cross_cap);
    for (int c = 0; c < 3; ++c) { x_f[c * n + i] = R.xf[c]; k_f[c * n + i] = R.kf[c]; }
    u7_f[i] = R.u7f; status[i] = R.status; n_cross[i] = (int)R.cr.size();
    for (int j = 0; j < (int)R.cr.size() && j < cross_cap; ++j) {
      double* r = cross_rec + ((size_t)i * cross_cap + j) * ART_CROSS_REC;
      r[0] = R.cr[j].x; r[1] = R.cr[j].y; r[2] = R.cr[j].z; r[3] = R.cr[j].kx; r[4] = R.cr[j].ky;
      r[5] = R.cr[j].kz; r[6] = R.cr[j].t; r[7] = R.cr[j].dw;
    }
    if (counters) {
      counters[4 * i] = R.n_acc; counters[4 * i + 1] = R.n_rej; counters[4 * i + 2] = R.n_rhs; counters[4 * i + 3] = R.n_cond;
    }
  }
  return ART_OK;
}

// main_runner_tree's event loop over e ∈ [first_event, first_event+n_events); OpenMP over
// events (the moral equivalent of the reference's N-process replication, runner_example.sh).
int orc_run_events(const art_params* p, uint64_t seed, int64_t first_event, int64_t n_events, double* rows,
                   int64_t rows_cap, int64_t* n_rows, int64_t* f_inx, int64_t* tot_count, art_stats* st, double* events) {
  double maxR;
  int rc = orc_max_radius(p, &maxR);
  if (rc) return rc;
  const int ncols = p->save_mode > 1 ? ART_ROW_COLS_FULL : (p->save_mode > 0 ? ART_ROW_COLS_LONG : ART_ROW_COLS_SHORT);
  std::vector<EventOut> ev((size_t)n_events);
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t i = 0; i < n_events; ++i) run_event(*p, maxR, seed, (uint64_t)(first_event + i), ncols, ev[(size_t)i]);
  int64_t nr = 0, fi = 0, tc = 0;
  art_stats S;
  std::memset(&S, 0, sizeof(S));
  for (auto& E : ev) {
    int64_t k = (int64_t)E.rows.size() / ncols;
    if (nr + k > rows_cap) return ART_ERR_CAPACITY;
    std::memcpy(rows + nr * ncols, E.rows.data(), E.rows.size() * sizeof(double));
    nr += k;
    fi += E.n_failed + E.n_final_photon;
    tc += E.tot_count;
    S.n_events++; S.n_attempts += E.n_failed + 1; S.n_sampler_cond += E.n_cond_sampler;
    S.n_rays += E.C.n_rays; S.n_rays_photon += E.C.n_rays_photon; S.n_acc += E.C.n_acc; S.n_rej += E.C.n_rej;
    S.n_rhs += E.C.n_rhs; S.n_cond += E.C.n_cond; S.n_final_photon += E.n_final_photon;
    S.n_final_axion += E.n_final_axion;
    for (int k2 = 0; k2 < 5; ++k2) S.n_status[k2] += E.C.n_status[k2];
  }
  if (events)
    for (int64_t i = 0; i < n_events; ++i)
      std::memcpy(events + i * ART_EVENT_COLS, ev[(size_t)i].event_line, sizeof(double) * ART_EVENT_COLS);
  *n_rows = nr; *f_inx = fi; *tot_count = tc;
  if (st) *st = S;
  return ART_OK;
}

int orc_run(const art_params* p, uint64_t seed, int64_t first_event, int64_t n_events, double* rows,
            int64_t rows_cap, int64_t* n_rows, int64_t* f_inx, int64_t* tot_count, art_stats* st) {
  return orc_run_events(p, seed, first_event, n_events, rows, rows_cap, n_rows, f_inx, tot_count, st, nullptr);
}

// step log of one ray: out[k*5 + (s, dt, EEst, accepted, event)], returns number of steps
int orc_trace_ray(const art_params* p, int32_t species, const double* x0, const double* k0, double erg,
                  double t_start, double dOmega, int32_t b_sign, int32_t max_crossings, double* out, int cap) {
  const double dt0 = std::exp(p->ln_t_start), ln_tend = std::log(1.0 / p->omega_pul);
  RayParams rp = ray_of(*p, erg, b_sign < 0 ? -p->B0 : p->B0, species != ART_SPECIES_PHOTON);
  std::vector<StepTrace> tr;
  propagate(x0, k0, rp, std::log(std::max(t_start, dt0)), ln_tend, max_crossings, dOmega, 1 << 20, &tr);
  int n = std::min((int)tr.size(), cap);
  for (int i = 0; i < n; ++i) {
    out[5 * i] = tr[i].s; out[5 * i + 1] = tr[i].dt; out[5 * i + 2] = tr[i].EEst;
    out[5 * i + 3] = tr[i].accepted; out[5 * i + 4] = tr[i].event;
  }
  return (int)tr.size();
}

void orc_set_literal_start_events(int on) { g_literal_start_events = on != 0; }

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
