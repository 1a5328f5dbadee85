// art_tree.cuh — per-event tree / re-conversion bookkeeping around the ray tracer.
//
// Replaces get_tree (reference MainRunner.jl:126-352) and the row assembly of main_runner_tree
// (MainRunner.jl:581-761).  The reference walks one event at a time on the host, popping the
// highest-weight node, propagating it, and pushing children.  Here every event owns a small
// node stack in HBM; one thread per event applies the loop body of get_tree to the node that
// was just propagated, emits its output row if it is final, and hands the next node to the ray
// queue of the next generation (art_kernels.cuh).  Semantics kept: weight splitting at the first
// crossing (MR.jl:296-303), Monte-Carlo fallback after MC_nodes (281-292), stops on
// prob_cutoff / num_cutoff / max_nodes with the info codes 2/3/4 (324-339), highest weight first
// with a stable ascending sort (342), info sign flip (346-348), the "rare fail" guard (213-224).
#pragma once
#include "art_philox.cuh"
#include "art_physics.cuh"
#include "art_tracer.cuh"

namespace art {

struct TreeNode {  // RT.node, RayTracer.jl:126-163 (fields that reach the outputs)
  double x, y, z, kx, ky, kz, t, dw;
  double prob, weight, prob_conv, prob_conv0;
  double parent_weight;  // weight of the node this one split off (−1: root), written by saveNode (MR.jl:17-19)
  int species, pad_;
};

struct EventData {  // per-event scalars of MainRunner.jl:498-557 and of the back-trace
  double x[3], k_init[3], vinf[3];
  double erg_inf, vel_eng, sln_prob, cos_w;
  double back_weight, prob0;  // samp_back_weight (MR.jl:630), Prob_nonAD_0 (MR.jl:626)
  double back_node_w;         // weight of the back-traced node itself, Π_j (1 − P_j) (MR.jl:313) — saveMode 3
  int c_bck, n_failed;
};

// saveMode 3 (MR.jl:17-65, 574-578, 612, 669-670): every node of an event's two trees, as the tree kernel met it.
// pos = 0: the back-traced axion; pos = 1, 2, …: the forward tree in the order of `push!(tree, event)`.
struct NodeLogEntry {
  int event, pos, species, n_cross;
  double weight, prob, parent_weight;
  double x0[3], k0[3], t0, dw0, erg;  // what propagate was called with (MR.jl:175-191): the trajectory is re-traced from here
};
struct NodeLog {
  NodeLogEntry* e;            // null: logging off
  double* cross;              // [cap][cross_cap][4]: x y z t of the recorded crossings
  unsigned long long* count;
  long long cap;
  int cross_cap;
};

struct TreeState {
  TreeNode cur;     // node being propagated
  double tot_prob;
  int count, count_main, info, stack_n, rng_k, tree_len, active, back_status;
};

struct RowSink {
  double* rows;                  // row-major, ncols per row
  long long* keys;               // (local event index << 8) | position in tree, for ordering
  unsigned long long* n_rows;    // running count
  unsigned long long cap;
  int ncols;
  unsigned long long* n_final;   // [0] photons, [1] axions
};

ART_HD unsigned long long sink_add(unsigned long long* p, unsigned long long v) {
#if defined(__CUDA_ARCH__)
  return atomicAdd(p, v);
#else
  unsigned long long o = *p;
  *p += v;
  return o;
#endif
}

ART_HD TreeNode make_node(double x, double y, double z, double kx, double ky, double kz, double t, double dw,
                          int species, double prob, double weight, double prob_conv, double prob_conv0, double parent_weight) {
  TreeNode n;
  n.x = x; n.y = y; n.z = z; n.kx = kx; n.ky = ky; n.kz = kz; n.t = t; n.dw = dw;
  n.species = species; n.pad_ = 0; n.prob = prob; n.weight = weight; n.prob_conv = prob_conv; n.prob_conv0 = prob_conv0;
  n.parent_weight = parent_weight;
  return n;
}

// keep the stack sorted ascending by weight, stable (== push! … sort!(events, by=weight), MR.jl:342)
ART_HD void stack_insert(TreeNode* stack, int& n, const TreeNode& nd) {
  int pos = n;
  while (pos > 0 && stack[pos - 1].weight > nd.weight) {
    stack[pos] = stack[pos - 1];
    --pos;
  }
  stack[pos] = nd;
  ++n;
}

ART_HD double conv_P(const DevParams& P, const double pos[3], const double k[3], double erg, double bsign) {
  return 1.0 - exp(-conv_prob_nonAD(P, pos, k, erg, bsign));  // MR.jl:136,267
}

// ---- back-trace result (MR.jl:581-630): get_tree(num_cutoff=0, splittings_cutoff=1e5) ----
ART_HD void tree_backward(const DevParams& P, EventData& E, const RayOut& R, const double* rec, int rec_cap) {
  const double kb[3] = {-E.k_init[0], -E.k_init[1], -E.k_init[2]};
  const double P0 = conv_P(P, E.x, kb, E.erg_inf * 1.0, -1.0);  // first.prob, MR.jl:132-137 (Δω = −1)
  double weight = 1.0;
  int nc = R.n_cross < rec_cap ? R.n_cross : rec_cap;
  bool rare = false;  // MR.jl:213-224
  for (int j = 0; j < nc; ++j) {
    const double* q = rec + (size_t)j * 8;
    rare = rare || fabs(q[3]) > 1.0 || fabs(q[4]) > 1.0 || fabs(q[5]) > 1.0;
  }
  if (!rare) {
    // de-duplicate crossings closer than 1e-5 km (keep the later one), MR.jl:227-245
    bool any = false;
    for (int j = 0; j + 1 < nc; ++j) {
      const double* a = rec + (size_t)j * 8;
      const double* b = a + 8;
      const double d = sqrt((b[0] - a[0]) * (b[0] - a[0]) + (b[1] - a[1]) * (b[1] - a[1]) + (b[2] - a[2]) * (b[2] - a[2]));
      any = any || d < 1e-5;
    }
    for (int j = 0; j < nc; ++j) {
      const double* a = rec + (size_t)j * 8;
      if (any && j + 1 < nc) {
        const double* b = a + 8;
        const double d = sqrt((b[0] - a[0]) * (b[0] - a[0]) + (b[1] - a[1]) * (b[1] - a[1]) + (b[2] - a[2]) * (b[2] - a[2]));
        if (!(d > 1e-5)) continue;
      }
      const double Pj = conv_P(P, a, a + 3, E.erg_inf * fabs(a[7]), -1.0);  // MR.jl:265-267
      weight *= (1.0 - Pj);                                                 // MR.jl:313
    }
  }
  E.prob0 = P0;
  E.back_weight = P0 * weight;
  E.back_node_w = weight;
  E.c_bck = 1;
}

// ---- forward tree: first node (MR.jl:653-657, 132-141) -----------------------------------
ART_HD void tree_forward_start(const DevParams& P, const EventData& E, TreeState& T) {
  TreeNode first = make_node(E.x[0], E.x[1], E.x[2], E.k_init[0], E.k_init[1], E.k_init[2], 0.0, -1.0,
                             SPECIES_PHOTON, 1.0, 1.0, -1.0, -1.0, -1.0);
  first.prob = conv_P(P, E.x, E.k_init, E.erg_inf * 1.0, 1.0);
  T.cur = first;
  T.tot_prob = 0.0;
  T.count = 1;  // the pop of the first node
  T.count_main = 0;
  T.info = 1;
  T.stack_n = 0;
  T.rng_k = 0;
  T.tree_len = 0;
  T.active = 1;
}

ART_HD void emit_row(const DevParams& P, const EventData& E, const TreeState& T, const TreeNode& n,
                     const RayOut& R, double event_id, long long local_event, const RowSink& S) {
  const double absf = sqrt(R.kf[0] * R.kf[0] + R.kf[1] * R.kf[1] + R.kf[2] * R.kf[2]);
  const double absfX = sqrt(R.xf[0] * R.xf[0] + R.xf[1] * R.xf[1] + R.xf[2] * R.xf[2]);
  const int id = n.species == SPECIES_PHOTON ? 1 : 0;
  const double w = n.weight;  // × samp_back_weight (MR.jl:686) is applied by patch_row once the back-trace is in
  sink_add(S.n_final + (id == 1 ? 0 : 1), 1ULL);
  const unsigned long long slot = sink_add(S.n_rows, 1ULL);
  if (slot >= S.cap) return;
  double* row = S.rows + slot * (unsigned long long)S.ncols;
  S.keys[slot] = (local_event << 8) | (long long)(T.tree_len & 0xFF);
  row[0] = event_id;
  row[1] = (double)id;
  row[2] = acos(R.kf[2] / absf);       // θf   MR.jl:677
  row[3] = atan2(R.kf[1], R.kf[0]);    // ϕf   MR.jl:675
  row[4] = acos(R.xf[2] / absfX);      // θfX  MR.jl:678
  row[5] = atan2(R.xf[1], R.xf[0]);    // ϕfX  MR.jl:676
  row[6] = absfX;
  row[7] = E.sln_prob;                 // divided by f_inx later (MR.jl:749)
  row[8] = w;                          // weight_tmp (opticalDepth = 0, weightC = 1)
  row[9] = E.x[0];
  row[10] = E.x[1];
  row[11] = E.x[2];
  row[12] = R.u7f / P.ma + E.vel_eng;  // Δω   MR.jl:709
  if (S.ncols > 13) {                  // MR.jl:715
    row[13] = w;
    row[14] = 0.0;
    row[15] = 1.0;
    row[16] = E.k_init[0];
    row[17] = E.k_init[1];
    row[18] = E.k_init[2];
    row[19] = E.cos_w;
    row[20] = 0.0;  // c    — patched when the tree is complete
    row[21] = 0.0;  // info — patched when the tree is complete
    row[22] = n.prob;
    row[23] = n.prob_conv;
    row[24] = n.prob_conv0;
    row[25] = 0.0;  // samp_back_weight — patch_row
    row[26] = absfX;
    row[27] = 0.0;  // c_bck            — patch_row
    row[28] = 0.0;  // Prob_nonAD_0     — patch_row
  }
  if (S.ncols > 29) {  // save_mode ≥ 2: the final_* text file's extra numbers, MR.jl:689-701
    row[29] = absf;
    row[30] = n.t;
  }
}

// Completes a row once its event is finished: the forward tree and the back-trace of an event run
// concurrently, so the back-traced weight (MR.jl:630,686) and the tree's final count / info
// (MR.jl:715, columns c and info) are filled in afterwards.
ART_HD void patch_row(double* row, int ncols, const EventData& E, int count, int info) {
  row[8] *= E.back_weight;
  if (ncols > 13) {
    row[13] *= E.back_weight;
    row[20] = (double)count;
    row[21] = (double)info;
    row[25] = E.back_weight;
    row[27] = (double)E.c_bck;
    row[28] = E.prob0;
  }
}

// ---- loop body of get_tree after `propagate` returned (MR.jl:192-342), forward mode --------
// On return T.active tells whether T.cur holds the next node to propagate.
ART_HD void tree_forward_update(const DevParams& P, const EventData& E, TreeState& T, TreeNode* stack,
                                const RayOut& R, const double* rec, uint64_t seed, uint64_t event_id,
                                long long local_event, const RowSink& S) {
  const TreeNode ev = T.cur;
  bool skip_checks = false;
  if (R.n_cross < 1) {  // MR.jl:200-207
    T.count_main++;
    T.tot_prob += ev.weight;
    const double rf = sqrt(R.xf[0] * R.xf[0] + R.xf[1] * R.xf[1] + R.xf[2] * R.xf[2]);
    if (rf > P.rNS * 1.1) emit_row(P, E, T, ev, R, (double)event_id, local_event, S);
  } else {
    const double* c0 = rec;
    if (fabs(c0[3]) > 1.0 || fabs(c0[4]) > 1.0 || fabs(c0[5]) > 1.0) {  // MR.jl:213-224
      T.tot_prob += ev.weight;
      skip_checks = true;  // `continue`
    } else {
      const double Pc = conv_P(P, c0, c0 + 3, E.erg_inf * fabs(c0[7]), 1.0);  // MR.jl:265-267
      const int new_species = ev.species == SPECIES_PHOTON ? SPECIES_AXION : SPECIES_PHOTON;
      if (T.count > P.MC_nodes) {  // MR.jl:281-292
        const double r = tree_uniform(seed, event_id, (uint32_t)T.rng_k++);
        if (r < Pc)
          stack_insert(stack, T.stack_n, make_node(c0[0], c0[1], c0[2], c0[3], c0[4], c0[5], c0[6], c0[7],
                                                   new_species, Pc, ev.weight, Pc, Pc, ev.weight));
        else
          stack_insert(stack, T.stack_n, make_node(c0[0], c0[1], c0[2], c0[3], c0[4], c0[5], c0[6], c0[7],
                                                   ev.species, 1.0 - Pc, ev.weight, Pc, ev.prob_conv, ev.weight));
      } else {  // MR.jl:296-303
        stack_insert(stack, T.stack_n, make_node(c0[0], c0[1], c0[2], c0[3], c0[4], c0[5], c0[6], c0[7],
                                                 new_species, Pc, Pc * ev.weight, Pc, Pc, ev.weight));
        stack_insert(stack, T.stack_n, make_node(c0[0], c0[1], c0[2], c0[3], c0[4], c0[5], c0[6], c0[7],
                                                 ev.species, 1.0 - Pc, (1.0 - Pc) * ev.weight, Pc, ev.prob_conv, ev.weight));
      }
    }
  }
  T.tree_len++;  // push!(tree, event)
  bool stop = false;
  if (!skip_checks) {
    if (T.tot_prob >= 1.0 - P.prob_cutoff) {  // MR.jl:324-327
      T.info = 2;
      stop = true;
    } else if (T.count_main >= P.num_cutoff) {  // MR.jl:331-334
      T.info = 3;
      stop = true;
    } else if (T.count > P.max_nodes) {  // MR.jl:336-339
      T.info = 4;
      stop = true;
    }
  }
  if (!stop && T.stack_n > 0) {  // next iteration of `while length(events) > 0`
    T.count++;
    T.cur = stack[--T.stack_n];
    T.active = 1;
  } else {
    if (T.count > P.MC_nodes) T.info = -(T.info < 0 ? -T.info : T.info);  // MR.jl:346-348
    T.active = 0;
  }
}

}  // namespace art
