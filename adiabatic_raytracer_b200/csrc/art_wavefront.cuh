// art_wavefront.cuh — step-batched ("wavefront") ray tracer + in-loop tree for sm_100a.
//
// The persistent one-kernel tracer (k_trace, art_kernels.cuh) is instruction-fetch bound on
// B200: its per-step code path (right-hand side, controller, dense output, crossing condition,
// root finder, callbacks ≈ 5k SASS instructions) is walked once per step by 8 resident warps
// that sit in different places, and 60-90 % of its stall samples are stall_no_inst
// (profiles/r01_trace_v4_hotlines.txt).  Here every phase of a step is its own small kernel
// that ALL resident warps execute together, the ray state lives in HBM between phases
// (≈ 0.9 KB per ray and step — three orders of magnitude below the FP64 work it frames), and
// the thread ↔ ray mapping is rebuilt every step by ballot/popc compaction:
//
//   k_wf_init      once per batch: IC transform of every queued ray
//   k_wf_stages<S> one thread per ACTIVE ray of species S (axion / photon lists are kept apart,
//                  so a warp never executes both right-hand sides): start-up or the six DP5
//                  stages (stage derivatives in shared memory) + error control; accepted steps
//                  publish their dense-output segment, classify their M−1 scan points with the
//                  far-field test and append the points that need the full condition to a
//                  GLOBAL work list (one atomic per warp)
//   k_wf_scan      one thread per work item (ray, point): the thick-surface condition — every
//                  lane busy whatever the accept/reject pattern or the distance of its ray
//   k_wf_detect    one thread per active ray: first sign change (bit masks), callbacks and state
//                  update when there is no crossing; rays with a bracket go to the root list
//   k_wf_root      one thread per bracketed ray: Illinois root finder (all lanes iterate
//                  together), crossing callback, state update
//   k_wf_tree      one thread per FINISHED ray: get_tree's loop body (art_tree.cuh); the next
//                  node of the event is re-initialised in the same ray slot and joins the next
//                  step's active list — events of any tree depth share one loop, there are no
//                  per-generation launches or host round trips
//
// Survivors are appended to the next step's active lists inside detect/root/tree (warp-aggregated
// atomics); nothing waits on another kernel; the host reads one counter every few steps.
#pragma once
#include <cuda_runtime.h>

#include "art_kernels.cuh"

namespace art {

constexpr int WF_THREADS = 128;
#ifndef ART_STAGE_MINBLOCKS
#define ART_STAGE_MINBLOCKS 2  // resident CTAs per SM the stage kernels' register allocation is tuned for
#endif
constexpr int SEG_PAD = 36;  // doubles per published segment (SEG_N = 35, padded)
constexpr int KF_PAD = 12;   // k1 and k7 of the last accepted step (FSAL), 6 doubles each

enum { WC_ITEMS = 0, WC_ROOTS = 1, WC_FIN = 2, WC_ACT = 3 /* + parity*2 + species */, WC_N = 8 };

struct WfPool {
  Ray* ray;                   // [cap] ray state between kernels
  double* kf;                 // [cap][12]
  double* seg;                // [cap][36]
  unsigned long long* need;   // [cap] scan points awaiting the full condition
  unsigned long long* neg;    // [cap] scan points proven negative by the far-field test
  unsigned char* pend;        // [cap] 1 = accepted step waiting for k_wf_detect
  unsigned char* sign;        // [cap][64]
  int* act[2][2];             // [parity][species][cap] ray ids
  unsigned* items;            // [cap * (M−1+1)] work list: (ray << 6) | point
  int* roots;                 // [cap]
  int* fin;                   // [cap]
  unsigned long long* ctr;    // WC_*
};

struct WfTree {  // per-event tree context (null ev ⇒ plain ray batch, art_propagate)
  EventData* ev;
  TreeState* ts;
  TreeNode* stacks;
  int stack_stride;
  unsigned long long seed;
  long long first_event, chunk_base, n_events;
  RowSink sink;
  unsigned long long* tot_count;
  int* fin_count;
  int* fin_info;
};

// FSAL copy K(0) ← K(6) of Ray::apply goes through a view of which only rows 0 and 6 exist
struct KfView {
  double* kf;
  __device__ __forceinline__ double& operator()(int j, int i) const { return kf[(j == 0 ? 0 : 6) + i]; }
};

struct WfStat {
  unsigned long long c[11];
  __device__ __forceinline__ WfStat() {
#pragma unroll
    for (int k = 0; k < 11; ++k) c[k] = 0;
  }
  __device__ __forceinline__ void add(const RayOut& o, bool axion) {
    c[0] += o.n_acc; c[1] += o.n_rej; c[2] += o.n_rhs; c[3] += o.n_cond; c[4] += 1; c[5] += axion ? 0 : 1;
    c[6 + o.status] += 1;
  }
  __device__ __forceinline__ void flush(DevStats* stats) {
#pragma unroll
    for (int k = 0; k < 11; ++k) c[k] = warp_sum(c[k]);
    if ((threadIdx.x & 31) == 0 && stats) {
      if (c[0]) atomicAdd(&stats->n_acc, c[0]);
      if (c[1]) atomicAdd(&stats->n_rej, c[1]);
      if (c[2]) atomicAdd(&stats->n_rhs, c[2]);
      if (c[3]) atomicAdd(&stats->n_cond, c[3]);
      if (c[4]) atomicAdd(&stats->n_rays, c[4]);
      if (c[5]) atomicAdd(&stats->n_rays_photon, c[5]);
      for (int k = 0; k < 5; ++k)
        if (c[6 + k]) atomicAdd(&stats->n_status[k], c[6 + k]);
    }
  }
};

// Every lane of the warp calls this: finished rays write their result and join the finished list,
// live rays join the next step's active list of their species (stream compaction of survivors).
__device__ __forceinline__ void wf_route(const DevParams& P, const WfPool& W, int next, bool valid, int id,
                                         const Ray& R, RayOut* out, WfStat& st) {
  const bool fin = valid && R.done;
  if (fin) {
    RayOut o;
    R.finish(P, o);
    out[id] = o;
    st.add(o, R.axion);
  }
  const long long sf = warp_pop(W.ctr + WC_FIN, fin);
  if (fin) W.fin[sf] = id;
#pragma unroll
  for (int sp = 0; sp < 2; ++sp) {
    const bool mine = valid && !R.done && ((R.axion ? SPECIES_AXION : SPECIES_PHOTON) == sp);
    const long long s = warp_pop(W.ctr + WC_ACT + next * 2 + sp, mine);
    if (mine) W.act[next][sp][s] = id;
  }
}

__global__ void k_wf_reset(WfPool W, int next) {
  if (threadIdx.x == 0) {
    W.ctr[WC_ITEMS] = 0; W.ctr[WC_ROOTS] = 0; W.ctr[WC_FIN] = 0;
    W.ctr[WC_ACT + next * 2] = 0; W.ctr[WC_ACT + next * 2 + 1] = 0;
  }
}

// ---- start-up of a batch: IC transform; the first stages() call of a ray does the rest --------
__global__ void __launch_bounds__(WF_THREADS)
k_wf_init(const __grid_constant__ DevParams P, const QueuedRay* __restrict__ q, long long n, WfPool W,
          RayOut* __restrict__ out, double* __restrict__ rec, int rec_cap, DevStats* stats) {
  WfStat st;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  const long long nround = ((n + nthreads - 1) / nthreads) * nthreads;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nround; i += nthreads) {
    Ray R;
    R.done = false;
    R.axion = false;
    const bool valid = i < n;
    if (valid) {
      R.init(P, q[i].in, rec + (size_t)i * rec_cap * 8, rec_cap);
      W.pend[i] = 0;
      W.ray[i] = R;
    }
    wf_route(P, W, 0, valid, (int)i, R, out, st);
  }
  st.flush(stats);
}

// ---- one iteration (start-up or attempted step) per active ray of species SP --------------------
template <int SP>
__global__ void __launch_bounds__(WF_THREADS, ART_STAGE_MINBLOCKS)
k_wf_stages(const __grid_constant__ DevParams P, WfPool W, int cur) {
  extern __shared__ double ksm[];
  const KStoreT<WF_THREADS> K{ksm + threadIdx.x};
  const long long n = (long long)W.ctr[WC_ACT + cur * 2 + SP];
  const int* __restrict__ act = W.act[cur][SP];
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  const long long nround = ((n + nthreads - 1) / nthreads) * nthreads;
  const unsigned lane = threadIdx.x & 31;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < nround; t += nthreads) {
    unsigned long long need = 0ull;
    int id = -1;
    if (t < n) {
      id = act[t];
      Ray R = W.ray[id];
      double* kf = W.kf + (size_t)id * KF_PAD;
      const bool had_k1 = !R.need_k1;
      if (had_k1)
        for (int k = 0; k < 6; ++k) K(0, k) = kf[k];
      const SegAoS D{W.seg + (size_t)id * SEG_PAD};
      const bool pending = R.template stages<SP>(P, K, D);
      if (!had_k1)
        for (int k = 0; k < 6; ++k) kf[k] = K(0, k);
      if (pending) {
        unsigned long long neg = 0ull;
        for (int k = 0; k < 6; ++k) kf[6 + k] = K(6, k);
        const int M = P.M_ray;
        const SegQuick Q(D);
        if (Q.all_far(P)) {
          neg = ((M >= 64) ? ~0ull : ((1ull << M) - 1ull)) & (R.event_last ? ~0ull : ~1ull);
        } else {
          for (int i = R.event_last ? 0 : 1; i < M; ++i) {
            const int qk = Q.classify(P, seg_theta(P, i));
            if (qk == 0) need |= 1ull << i;
            else if (qk == -1) neg |= 1ull << i;
          }
        }
        W.need[id] = need;
        W.neg[id] = neg;
      }
      W.pend[id] = pending ? 1 : 0;
      W.ray[id] = R;
    }
    // append this warp's scan points to the global work list: one atomic per warp
    const int cnt = __popcll(need);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(FULL, incl, o);
      if ((int)lane >= o) incl += v;
    }
    const int total = __shfl_sync(FULL, incl, 31);
    if (total) {
      unsigned long long base = 0;
      if (lane == 31) base = atomicAdd(W.ctr + WC_ITEMS, (unsigned long long)total);
      base = __shfl_sync(FULL, base, 31);
      unsigned long long pos = base + (unsigned long long)(incl - cnt);
      unsigned long long m = need;
      while (m) {
        const int i = __ffsll((long long)m) - 1;
        W.items[pos++] = ((unsigned)id << 6) | (unsigned)i;
        m &= m - 1ull;
      }
    }
  }
}

// ---- crossing condition at every listed (ray, point) -----------------------------------------------
__global__ void __launch_bounds__(256)
k_wf_scan(const __grid_constant__ DevParams P, WfPool W) {
  const long long n = (long long)W.ctr[WC_ITEMS];
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += nthreads) {
    const unsigned code = W.items[k];
    const unsigned id = code >> 6, i = code & 63u;
    const SegAoS D{W.seg + (size_t)id * SEG_PAD};
    const double g = seg_cond_full(P, D, seg_theta(P, (int)i));
    W.sign[(size_t)id * 64 + i] = (g > 0.0) ? 1 : ((g < 0.0) ? 2 : ((g == 0.0) ? 3 : 0));
  }
}

// ---- events: detection, no-crossing update, routing --------------------------------------------------
__global__ void __launch_bounds__(WF_THREADS)
k_wf_detect(const __grid_constant__ DevParams P, WfPool W, int cur, RayOut* __restrict__ out, DevStats* stats) {
  WfStat st;
  const int next = 1 - cur;
  const long long n0 = (long long)W.ctr[WC_ACT + cur * 2 + 0], n1 = (long long)W.ctr[WC_ACT + cur * 2 + 1];
  const long long n = n0 + n1;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  const long long nround = ((n + nthreads - 1) / nthreads) * nthreads;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < nround; t += nthreads) {
    Ray R;
    R.done = false;
    R.axion = false;
    bool valid = t < n, to_root = false;
    int id = -1;
    if (valid) {
      id = (t < n0) ? W.act[cur][0][t] : W.act[cur][1][t - n0];
      R = W.ray[id];
      if (W.pend[id]) {
        ScanMasks m{0ull, W.neg[id], 0ull};
        unsigned long long need = W.need[id];
        const unsigned char* sg = W.sign + (size_t)id * 64;
        while (need) {
          const int i = __ffsll((long long)need) - 1;
          const unsigned v = sg[i];
          if (v == 1) m.pos |= 1ull << i;
          else if (v == 2) m.neg |= 1ull << i;
          else if (v == 3) m.zero |= 1ull << i;
          need &= need - 1ull;
        }
        const SegAoS D{W.seg + (size_t)id * SEG_PAD};
        int ia, ib, sgl;
        const int ev = R.detect(P, D, m, ia, ib, sgl);
        if (ev == 1) {
          R.br_a = ia; R.br_b = ib; R.br_sg = sgl;
          to_root = true;
        } else {
          const KfView K{W.kf + (size_t)id * KF_PAD};
          R.apply(P, K, D, ev == 2 ? 1.0 : -1.0, sgl);
        }
        W.pend[id] = 0;
        W.ray[id] = R;
      }
    }
    const long long sr = warp_pop(W.ctr + WC_ROOTS, to_root);
    if (to_root) W.roots[sr] = id;
    wf_route(P, W, next, valid && !to_root, id, R, out, st);
  }
  st.flush(stats);
}

__global__ void __launch_bounds__(WF_THREADS)
k_wf_root(const __grid_constant__ DevParams P, WfPool W, int next, RayOut* __restrict__ out, DevStats* stats) {
  WfStat st;
  const long long n = (long long)W.ctr[WC_ROOTS];
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  const long long nround = ((n + nthreads - 1) / nthreads) * nthreads;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < nround; t += nthreads) {
    Ray R;
    R.done = false;
    R.axion = false;
    const bool valid = t < n;
    int id = -1;
    if (valid) {
      id = W.roots[t];
      R = W.ray[id];
      double segl[SEG_PAD];  // the root finder evaluates the segment ~12 times: keep it on chip
      {
        const double* sgp = W.seg + (size_t)id * SEG_PAD;
#pragma unroll
        for (int k = 0; k < SEG_PAD; ++k) segl[k] = sgp[k];
      }
      const SegAoS D{segl};
      const KfView K{W.kf + (size_t)id * KF_PAD};
      const double th = R.locate(P, D, R.br_a, R.br_b);
      R.apply(P, K, D, th, R.br_sg);
      W.ray[id] = R;
    }
    wf_route(P, W, next, valid, id, R, out, st);
  }
  st.flush(stats);
}

// ---- tree: finished rays → get_tree loop body → next node re-initialised in the same slot ------------
// Returns true when the event continues: R then holds the freshly initialised ray of the next node.
__device__ __forceinline__ bool wf_tree_step(const DevParams& P, const WfTree& T, int id, const RayOut& ro,
                                             double* __restrict__ rec, int rec_cap, Ray& R) {
  const double* rc = rec + (size_t)id * rec_cap * 8;
  if (id < T.n_events) {  // back-traced axion of event id: samp_back_weight
    EventData E = T.ev[id];
    tree_backward(P, E, ro, rc, rec_cap);
    T.ev[id] = E;
    T.ts[id].back_status = ro.status;
    return false;
  }
  const int e = id - (int)T.n_events;
  const EventData E = T.ev[e];
  TreeState S = T.ts[e];
  tree_forward_update(P, E, S, T.stacks + (size_t)e * T.stack_stride, ro, rc, T.seed,
                      (unsigned long long)(T.first_event + T.chunk_base + e), T.chunk_base + e, T.sink);
  bool spawn = false;
  if (S.active) {
    RayIn in;
    in.x0[0] = S.cur.x; in.x0[1] = S.cur.y; in.x0[2] = S.cur.z;
    in.k0[0] = S.cur.kx; in.k0[1] = S.cur.ky; in.k0[2] = S.cur.kz;
    in.erg = E.erg_inf; in.t_start = S.cur.t; in.dOmega = S.cur.dw; in.max_cross = -1;
    in.species = S.cur.species;
    R.init(P, in, rec + (size_t)id * rec_cap * 8, rec_cap);
    spawn = true;
  } else {
    atomicAdd(T.tot_count, (unsigned long long)S.tree_len);
    T.fin_count[e] = S.count;
    T.fin_info[e] = S.info;
  }
  T.ts[e] = S;
  return spawn;
}

__global__ void __launch_bounds__(WF_THREADS)
k_wf_tree(const __grid_constant__ DevParams P, WfPool W, WfTree T, int next, RayOut* __restrict__ out,
          double* __restrict__ rec, int rec_cap, DevStats* stats) {
  WfStat st;
  const long long n = (long long)W.ctr[WC_FIN];
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  const long long nround = ((n + nthreads - 1) / nthreads) * nthreads;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < nround; t += nthreads) {
    Ray R;
    R.done = true;
    R.axion = false;
    bool spawn = false;
    int id = -1;
    if (t < n) {
      id = W.fin[t];
      spawn = wf_tree_step(P, T, id, out[id], rec, rec_cap, R);
      if (spawn) {
        W.pend[id] = 0;
        W.ray[id] = R;
      }
    }
    // a re-initialised ray that is already past the end of the span retires at once (never for t < 1/Ω)
    wf_route(P, W, next, spawn, id, R, out, st);
  }
  st.flush(stats);
}

// ---- tail: when few rays are left the step-batched loop is latency bound (six kernels per step
// for a handful of warps), so ONE persistent kernel finishes them: every lane takes an active ray,
// steps it sequentially to its end, walks on through its event's tree in place, then takes the next.
__global__ void __launch_bounds__(WF_THREADS)
k_wf_tail(const __grid_constant__ DevParams P, WfPool W, WfTree T, int cur, RayOut* __restrict__ out,
          double* __restrict__ rec, int rec_cap, DevStats* stats) {
  extern __shared__ double ksm[];
  const KStore K{ksm + threadIdx.x, (int)blockDim.x};
  WfStat st;
  const long long n0 = (long long)W.ctr[WC_ACT + cur * 2 + 0], n1 = (long long)W.ctr[WC_ACT + cur * 2 + 1];
  const long long n = n0 + n1;
  Ray R;
  R.done = true;
  bool have = false, exhausted = false;
  int id = -1;
  for (;;) {
    const long long got = warp_pop(W.ctr + WC_ITEMS, !have && !exhausted);
    if (!have && !exhausted) {
      if (got >= 0 && got < n) {
        id = (got < n0) ? W.act[cur][0][got] : W.act[cur][1][got - n0];
        R = W.ray[id];
        if (!R.need_k1)
          for (int k = 0; k < 6; ++k) K(0, k) = W.kf[(size_t)id * KF_PAD + k];
        have = true;
      } else {
        exhausted = true;
      }
    }
    if (__ballot_sync(FULL, have) == 0) break;
    if (have) {
      if (!R.done) {
        const SegStore D{W.seg + (size_t)id * SEG_PAD, 1};
        R.step(P, K, D);
      }
      if (R.done) {
        RayOut o;
        R.finish(P, o);
        out[id] = o;
        st.add(o, R.axion);
        have = false;
        if (T.ev) have = wf_tree_step(P, T, id, o, rec, rec_cap, R);
      }
    }
  }
  st.flush(stats);
}

}  // namespace art
