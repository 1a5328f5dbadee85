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
// Survivors are appended to the next step's active lists inside stages/detect/root/tree (warp-aggregated
// atomics); the host reads a few counters every 8–32 steps.  k_wf_root and k_wf_tree are latency-bound
// (≈ 30–50 µs whatever the list length), so they run on a second stream, one iteration behind: the roots
// and finished rays that iteration k produces are processed while iteration k+1's stage kernel runs, and
// rejoin the active list of iteration k+2.  Lists are double-buffered by iteration parity; the list COUNTERS
// come in four sets indexed by iteration mod 4: while iteration k runs, sets k−1 (roots / finished lists the
// side pass consumes), k and k+1 (active lists being filled) are live and set k+2 is dead, so the stage kernel
// of iteration k zeroes it — no reset launch.
//
// The fused pipeline keeps the loop FULL (art_api.cu, run_pool): events live in a pool of slots (slot s: ray s =
// back-traced axion, ray S+s = forward tree); when both halves of an event are through, its slot goes onto a
// device-side ring of free slots, and k_wf_inject — part of the side pass — starts freshly sampled events in
// them.  The thin tail of late tree nodes (which used to be half the tracer time of a batch) then rides on the
// bulk of the following events and only the very last block of a call drains alone.
#pragma once
#include <cuda_runtime.h>

#include "art_kernels.cuh"

namespace art {

#ifndef ART_WF_THREADS
#define ART_WF_THREADS 32   // threads per CTA of the init / detect / root / tree kernels (tracer per 1e6 events with
                            // one-warp stage CTAs: 256 threads 207.6 ms, 128 195.3, 64 188.2, 32 185.3)
#endif
constexpr int WF_THREADS = ART_WF_THREADS;
// The stage kernel runs ONE WARP per CTA: a CTA retires as soon as its warp is through, so a warp whose rays take all
// their attempts does not hold the slots of three finished ones.  Measured per 1e6 events (tracer, 255 registers, 8 warps
// per SM in every case): 256 threads x 1 CTA 221 ms, 128 x 2 199 ms, 64 x 4 189 ms, 32 x 8 185 ms.
#ifndef ART_STAGE_THREADS
#define ART_STAGE_THREADS 32
#endif
constexpr int STAGE_THREADS = ART_STAGE_THREADS;
#ifndef ART_STAGE_MINBLOCKS
#define ART_STAGE_MINBLOCKS 12 // resident one-warp CTAs per SM the stage kernel's register allocation is tuned for.  Vern6, 3e6 events
                               // through the refilled pool (profiles/r02_variant_sweep.txt): 8 CTAs (220 registers, no spills) 901 ms;
                               // 10 CTAs 936; 12 CTAs (168 registers, 140 B of spills) 872; with the rolled tableau loop 12 CTAs 847
#endif
#ifndef ART_STAGE_NSTEP
#define ART_STAGE_NSTEP 4      // attempted steps a ray may take inside one k_wf_stages launch
#endif

// counters: 4 sets of WC_SET (indexed by iteration & 3), then the free-slot ring's tail
enum { WC_ACT = 0 /* + species */, WC_PEND = 2, WC_ROOTS = 3, WC_FIN = 4, WC_SET = 8, WC_RING_TAIL = 32, WC_N = 40 };
__device__ __forceinline__ unsigned long long* wc(unsigned long long* ctr, int k, int which) { return ctr + (k & 3) * WC_SET + which; }

struct WfPool {
  RayHot* hot;                // [cap] per-step ray state
  RayCold* cold;              // [cap] callback-only ray state
  double* seg;                // [cap][Seg<IG>::N] dense-output segment of a pending (accepted, unscanned) step
  unsigned long long* mpos;   // [cap] scan points with a positive condition (bit i = grid point i)
  unsigned long long* mneg;   // [cap] … negative (far-field points included)
  unsigned long long* mzero;  // [cap] … exactly zero
  int* act[2][2];             // [parity][species][cap] ray ids
  int* pend;                  // [cap] rays with a pending step
  int* roots[2];              // [list parity][cap] rays with a bracketed sign change, produced in one iteration and
                              // consumed (k_wf_root) during the next, concurrently with its stage kernel
  int* fin[2];                // [list parity][cap] finished rays, consumed by k_wf_tree / k_wf_collect the same way
  unsigned long long* ctr;    // WC_*
};

struct WfTree {  // per-event tree context of the fused pipeline
  EventData* ev;          // [events of the call] per-event scalars, indexed by the event's index in the call
  int* fin_count;         // [events of the call] tree count / info of the finished forward tree
  int* fin_info;
  TreeState* ts;          // [slots]
  TreeNode* stacks;       // [slots][stack_stride]
  int* slot_event;        // [slots] event (index in the call) that occupies the slot
  int* slot_parts;        // [slots] halves of the event that are through (back-trace, forward tree): 2 ⇒ the slot is free
  int* ring;              // [slots] ring of free slots; tail = ctr[WC_RING_TAIL] (pushes), head kept by the host (pops)
  int stack_stride;
  int n_slots;            // S: ray id < S = back-trace of slot id, otherwise forward tree of slot id − S
  unsigned long long seed;
  EventIds ids;           // event ids of the call (art_kernels.cuh)
  RowSink sink;
  unsigned long long* tot_count;
  NodeLog nlog;           // saveMode 3: every node of every tree (e = null otherwise)
};

// apply()'s FSAL copy k1 ← K(6): on the device k7 is already in the hot record
struct K7View {
  const double* k7;
  double k7x;
  __device__ __forceinline__ double operator()(int, int i) const { return k7[i]; }
  __device__ __forceinline__ double kx(int) const { return k7x; }
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

// Stream compaction of survivors.  Every lane of the warp calls this: finished rays join the finished
// list (their result is written by k_wf_tree / k_wf_collect), rays with a pending step the pending
// list, live rays the next step's active list of their species — one atomic per warp and list.
// k = the iteration whose kernels call this: finished / pending lists of iteration k, active lists of iteration k + 1
__device__ __forceinline__ void wf_route(const WfPool& W, int k, bool valid, int id, bool done, bool pending, bool axion) {
  const bool fin = valid && done;
  const long long sf = warp_pop(wc(W.ctr, k, WC_FIN), fin);
  if (fin) ((k & 1) ? W.fin[1] : W.fin[0])[sf] = id;
  const bool pd = valid && !done && pending;
  const long long sp_ = warp_pop(wc(W.ctr, k, WC_PEND), pd);
  if (pd) W.pend[sp_] = id;
#pragma unroll
  for (int sp = 0; sp < 2; ++sp) {
    const bool mine = valid && !done && !pending && ((axion ? SPECIES_AXION : SPECIES_PHOTON) == sp);
    const long long s = warp_pop(wc(W.ctr, k + 1, WC_ACT + sp), mine);
    if (mine) (((k + 1) & 1) ? W.act[1][sp] : W.act[0][sp])[s] = id;
  }
}

// Thin lists: when a kernel has fewer items than lanes, its items are dealt one per `spread` lanes, so
// that a warp holds 1–4 of them instead of 32 — a warp walks the UNION of its lanes' control flow
// (root-finder iterations, tree branches), and in the thin tail of a batch only latency counts.
__device__ __forceinline__ int wf_spread(long long n) {
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  int s = 1;
  while (s < 32 && n * (2LL * s) <= nthreads) s <<= 1;
  return s;
}

// ---- start-up of a batch: IC transform, first RHS, initial step size, starting sign ----------------
template <int IG>
__global__ void __launch_bounds__(WF_THREADS)
k_wf_init(const __grid_constant__ DevParams P, const QueuedRay* __restrict__ q, long long n, WfPool W,
          double* __restrict__ rec, int rec_cap, DevStats* stats) {
  if (blockIdx.x == 0 && threadIdx.x == 0 && stats) atomicAdd(&stats->n_cond_eval, (unsigned long long)n);  // start-point signs
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  const long long nround = ((n + nthreads - 1) / nthreads) * nthreads;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nround; i += nthreads) {
    Ray R;
    R.done = false;
    R.axion = false;
    const bool valid = i < n;
    if (valid) {
      R.init(P, q[i].in, rec + (size_t)i * rec_cap * 8, rec_cap);
      if (!R.done && R.iter_guard(P)) R.template startup<IG, -1>(P);
      W.hot[i] = static_cast<const RayHot&>(R);
      W.cold[i] = static_cast<const RayCold&>(R);
    }
    // rays that start past the end of the span stay `done`: the first k_wf_stages launch retires them
    wf_route(W, -1, valid, (int)i, false, false, R.axion);  // active lists of iteration 0
  }
}

// ---- attempted steps of the active rays of species SP ------------------------------------------------
// One thread per ray: 6 right-hand sides per attempt with the stage derivatives in shared memory (one
// column per thread), error control, then one of three outcomes —
//   rejected                       → new dt, try again
//   accepted, whole step provably in the far field and no sign change possible
//                                  → state advanced in registers, try the next step
//   accepted otherwise             → state advanced optimistically, dense-output segment published,
//                                    scan points classified and listed; ray joins the pending list
// A warp keeps stepping (at most `nstep` attempts) while at least `min_go` of its lanes can go on, so
// far-field rays and rejected steps do not pay a state round trip per attempt (bulk: 4 attempts, half
// the warp; thin tail of a batch: 16 attempts, any lane — there only latency counts).
template <int IG, int SP>
__device__ __forceinline__ void wf_stages_body(const DevParams& P, const WfPool& W, int k, int nstep, int min_go,
                                               unsigned bid, unsigned nblocks, double* ksm, DevStats* stats) {
  constexpr int LAST = Scheme<IG>::LAST, NEXTRA = Scheme<IG>::NEXTRA, TNEW = IG == IG_DP5 ? 5 : 6;  // park slot of e^{s_new}
  unsigned long long n_full = 0;  // condition evaluations done here (end points of in-box steps)
  unsigned long long n_scan = 0;  // steps handed to the crossing scan
  const KStoreT<STAGE_THREADS> K{ksm + threadIdx.x};
  const long long n = (long long)*wc(W.ctr, k, WC_ACT + SP);
  const int* __restrict__ act = (k & 1) ? W.act[1][SP] : W.act[0][SP];  // (a dynamic index would copy the parameter struct to local memory)
  const long long nthreads = (long long)nblocks * blockDim.x;
  const long long nround = ((n + nthreads - 1) / nthreads) * nthreads;
  // the ids of this thread's next two trips are kept in flight and the hot record of the next trip is
  // pulled into L2 while the current ray is being stepped (8 resident warps cannot hide an HBM miss)
  long long t = bid * (long long)blockDim.x + threadIdx.x;
  int id_a = t < n ? act[t] : -1;
  int id_b = t + nthreads < n ? act[t + nthreads] : -1;
  for (; t < nround; t += nthreads) {
    const int id = id_a;
    bool pending = false, done = false;
    const bool valid = id >= 0;
    id_a = id_b;
    if (id_a >= 0) {
      const char* nx = (const char*)(W.hot + id_a);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + 128));
    }
    id_b = t + 2 * nthreads < n ? act[t + 2 * nthreads] : -1;
    Ray R;
    if (valid) static_cast<RayHot&>(R) = W.hot[id];
    if constexpr (Scheme<IG>::BL) {
      R.delta = valid ? W.cold[id].delta : 0.0;
      R.k1x = valid ? W.cold[id].k1x : 0.0;
      R.dn = R.delta;
    }
    bool go = valid && !R.done;
    done = valid && R.done;
    // Rounds.  A step round runs the LAST stages of one attempted step for every lane that can go on, at ONE call
    // site of the right-hand side (the stage loop is rolled), then error control and the certified-sign tests.  A lane
    // whose accepted step has to be scanned stops there with its stage store intact (pending).  The final round —
    // after nstep step rounds, or as soon as fewer than min_go lanes can go on — runs the lazy dense-output stages of
    // Vern6 for all pending lanes of the warp together (one round per launch at ≈ half the lanes, instead of one per
    // attempt at a sixth), through the same call site; then the segments are published.
    for (int it = 0;; ++it) {
      const bool final_round = it >= nstep || __popc(__ballot_sync(FULL, go)) < (it == 0 ? 1 : min_go);
      bool work = pending;
      int st0 = LAST + 1, st1 = LAST + NEXTRA;
      if (!final_round) {
        if (go && !R.iter_guard(P)) {
          go = false;
          done = true;
        }
        if (go) R.template begin_attempt<IG>(P, K);
        work = go;
        st0 = 1;
        st1 = LAST;
      }
#pragma unroll 1
      for (int st = st0; st <= st1; ++st)
        if (work) R.template stage<IG, SP>(P, K, st);
      if (final_round) break;
      if (go && R.template control<IG>(P, K)) {
        // certified sign of the condition along the whole step: −1 far field / dark, +1 lit, 0 unknown
        int cs = R.template step_all_far<IG>(P, K) ? -1 : R.template step_sign<IG>(P, K, K.park(TNEW));
        if (cs != 0 && !R.event_last && R.sgn_end * cs < 0) cs = 0;  // a sign change against the reference sign: scan
        bool settled = cs != 0;
        if (!settled && R.n_cross == 0 && !R.event_last && !R.left_box) {
          // a step inside the start box keeps nothing but the sign at its end point (step_in_start_box):
          // evaluated here, the ray goes on without a round trip through the scan
          const RayCold& C = W.cold[id];
          const double s0_[3] = {C.s0[0], C.s0[1], C.s0[2]};
          if (R.template step_in_start_box<IG>(K, s0_, C.box_m)) {
            const double E = R.u7_of(R.un[5]) + (Scheme<IG>::BL ? R.dn : 0.0);
            const double g = thick_F<true>(P, R.un[0], R.un[1], R.un[2] - P.Omega * K.park(TNEW), R.un[3], R.un[4], R.un[5], E,
                                           P.rs_ray, true);
            n_full++;
            if (g != 0.0) {  // (an exact zero at the end point is an event: left to the scan)
              cs = (g > 0.0) - (g < 0.0);  // NaN: 0
              settled = true;
            }
          } else {
            R.left_box = 1;
          }
        }
        if (settled) {
          // no crossing possible: apply(th_ev < 0, sg_last = cs)
          for (int i = 0; i < 6; ++i) {
            R.u[i] = R.un[i];
            R.k1[i] = K(LAST, i);
          }
          if constexpr (Scheme<IG>::BL) {
            R.delta = R.dn;
            R.k1x = K.kx(LAST);
          }
          R.s = R.s_new;
          R.sgn_end = cs;
          R.event_last = false;
          R.dt = R.dt_new;
          R.settle(P);
          if (R.done) {
            go = false;
            done = true;
          }
        } else {
          go = false;
          pending = true;
        }
      }
    }
    if (pending) {
      R.n_rhs += NEXTRA;
      n_scan++;
      const SegAoS D{W.seg + (size_t)id * Seg<IG>::N};
      R.template publish<IG>(K, D);
      D(Seg<IG>::NUDGE) = R.event_last ? 1.0 : 0.0;
      // optimistic advance; k_wf_detect / k_wf_root complete or override it
      for (int i = 0; i < 6; ++i) {
        R.u[i] = R.un[i];
        R.k1[i] = K(LAST, i);
      }
      if constexpr (Scheme<IG>::BL) {
        R.delta = R.dn;
        R.k1x = K.kx(LAST);
      }
      R.s = R.s_new;
      R.dt = R.dt_new;
    }
    if (valid) W.hot[id] = static_cast<const RayHot&>(R);
    if constexpr (Scheme<IG>::BL) {
      if (valid) {
        W.cold[id].delta = R.delta;
        W.cold[id].k1x = R.k1x;
      }
    }
    wf_route(W, k, valid, id, done, pending, SP == SPECIES_AXION);
  }
  n_full = warp_sum(n_full);
  n_scan = warp_sum(n_scan);
  if ((threadIdx.x & 31) == 0 && stats) {
    if (n_full) atomicAdd(&stats->n_cond_eval, n_full);
    if (n_scan) atomicAdd(&stats->n_scanned, n_scan);
  }
}

// One launch steps both species: the first half of the grid walks the axion list, the second half the
// photon list (a warp never executes both right-hand sides; in the thin tail of a batch the two lists
// are stepped concurrently instead of one launch after the other).
template <int IG>
#ifdef ART_STAGE_MAXNREG  // tuning: an explicit register cap instead of the cap implied by the resident-CTA count (tracer per
                          // 1e6 events: 224 registers / 9 warps 187.9 ms, 200 / 10 190.2, 184 / 11 190.3; 255 / 8: 185.3)
__global__ void __maxnreg__(ART_STAGE_MAXNREG)
#else
__global__ void __launch_bounds__(STAGE_THREADS, ART_STAGE_MINBLOCKS)
#endif
k_wf_stages(const __grid_constant__ DevParams P, WfPool W, int k, int nstep, int min_go, DevStats* stats) {
  extern __shared__ double ksm[];
  // counter set k + 2 is dead while iteration k runs (see the header): zero it for its next use
  if (blockIdx.x == 0 && threadIdx.x < WC_SET) wc(W.ctr, k + 2, 0)[threadIdx.x] = 0ull;
  const unsigned half = gridDim.x >> 1;
  if (blockIdx.x < half) wf_stages_body<IG, SPECIES_AXION>(P, W, k, nstep, min_go, blockIdx.x, half, ksm, stats);
  else wf_stages_body<IG, SPECIES_PHOTON>(P, W, k, nstep, min_go, blockIdx.x - half, half, ksm, stats);
}

// ---- crossing scan: one warp per group of pending steps -----------------------------------------------
// A step has M slots — its M−1 grid points and the nudge point that is only scanned after an event.  The
// slots of up to three steps (3·M ≤ 160) are laid end to end and dealt to the lanes, 32 per round, so that
// 94 % of the lanes hold a point (one step per warp would leave 22 % idle: 50 slots in 64).  Every lane first
// classifies its point with the far-field test (two polynomials); the lanes whose point needs it then
// evaluate the condition together; three ballots per round turn the signs into the bit masks k_wf_detect
// works on.  No per-ray work list, no per-point sign bytes: each segment is read once (staged in shared
// memory) and 24 bytes of masks are written per step.
#ifndef ART_SCAN_THREADS
#define ART_SCAN_THREADS 256
#endif
constexpr int SCAN_THREADS = ART_SCAN_THREADS;
#ifndef ART_SCAN_PACK
#define ART_SCAN_PACK 3
#endif
constexpr int SCAN_PACK = ART_SCAN_PACK;  // steps per warp and trip (at most)
constexpr int SCAN_ROUNDS = 5;  // ballot words kept per sign class

#ifndef ART_SCAN_MINBLOCKS
#define ART_SCAN_MINBLOCKS 3
#endif
template <int IG>
__global__ void __launch_bounds__(SCAN_THREADS, ART_SCAN_MINBLOCKS)
k_wf_scan(const __grid_constant__ DevParams P, WfPool W, int k, DevStats* stats) {
  constexpr int SEG_PAD = Seg<IG>::N;
  __shared__ double seg_sm[SCAN_THREADS / 32][SCAN_PACK][SEG_PAD];
  __shared__ double th_sm[SCAN_ROUNDS][SCAN_THREADS];        // per lane and round: dense-output parameter of its slot
  __shared__ short j_sm[SCAN_ROUNDS][SCAN_THREADS];          // … segment index, +8 for a nudge slot, −1 for no slot
  __shared__ unsigned bal_sm[SCAN_THREADS / 32][3][SCAN_ROUNDS + 2];  // ballot words of the current trip
  const long long n = (long long)*wc(W.ctr, k, WC_PEND);
  const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long nwarps = (long long)gridDim.x * (SCAN_THREADS / 32);
  const int M = P.M_ray;
  const int pack = min(SCAN_PACK, (32 * SCAN_ROUNDS) / M);  // M ≤ 64 ⇒ pack ≥ 2
  const int rounds = (pack * M + 31) >> 5;
  // what this lane does in each round does not depend on the steps: which of the (up to three) staged
  // segments it reads, at which dense-output parameter, and whether its slot is a nudge slot.  Kept in
  // shared memory so that the round loop can stay ROLLED (one copy of the condition: unrolled five times the
  // kernel spent 1.4 warps per issue waiting for instructions — profiles/r01_wf12_iteration10_keys.txt)
  for (int rd = 0; rd < SCAN_ROUNDS; ++rd) {
    const int slot = rd * 32 + (int)lane;
    const int j = slot / M, q = slot - j * M;
    const bool has = rd < rounds && j < pack;
    j_sm[rd][threadIdx.x] = (short)(has ? (q < M - 1 ? j : j + 8) : -1);
    th_sm[rd][threadIdx.x] = seg_theta(P, (q < M - 1) ? q + 1 : 0);
  }
  if (lane < 3 * (SCAN_ROUNDS + 2)) (&bal_sm[wib][0][0])[lane] = 0u;
  unsigned long long n_full = 0;
  for (long long p = (blockIdx.x * (long long)(SCAN_THREADS / 32) + wib) * pack; p < n; p += nwarps * pack) {
    const int cnt = (int)min((long long)pack, n - p);
    int my_id = -1;
    if ((int)lane < cnt) my_id = W.pend[p + lane];
    __syncwarp();
    for (int j = 0; j < cnt; ++j) {
      const int id = __shfl_sync(FULL, my_id, j);
      const double* src = W.seg + (size_t)id * SEG_PAD;
      seg_sm[wib][j][lane] = src[lane];
      if (lane < SEG_PAD - 32) seg_sm[wib][j][32 + lane] = src[32 + lane];
    }
    __syncwarp();
#pragma unroll 1
    for (int rd = 0; rd < rounds; ++rd) {
      int code = 0;  // 1 positive, 2 negative, 3 zero, 0 NaN / no point
      const int j = j_sm[rd][threadIdx.x];
      if (j >= 0 && (j & 7) < cnt) {
        const SegAoS D{seg_sm[wib][j & 7]};
        if (j < 8 || D(Seg<IG>::NUDGE) != 0.0) {
          const double th = th_sm[rd][threadIdx.x];
          const int qk = seg_quick<IG>(P, D, th);
          if (qk == -1) code = 2;
          else if (qk == 0) {
            const double g = seg_cond_full<IG, true>(P, D, th);
            n_full++;
            code = (g > 0.0) ? 1 : ((g < 0.0) ? 2 : ((g == 0.0) ? 3 : 0));
          }
        }
      }
      const unsigned bp = __ballot_sync(FULL, code == 1), bn = __ballot_sync(FULL, code == 2), bz = __ballot_sync(FULL, code == 3);
      if (lane == 0) {
        bal_sm[wib][0][rd] = bp; bal_sm[wib][1][rd] = bn; bal_sm[wib][2][rd] = bz;
      }
    }
    __syncwarp();
    if ((int)lane < cnt) {
      // bits [off, off + M) of each class, re-ordered to grid-point order: slot q < M−1 ↔ point q + 1,
      // slot M−1 ↔ the nudge point 0
      const int off = (int)lane * M, w = off >> 5, sh = off & 31;
      unsigned long long m3[3];
      for (int c = 0; c < 3; ++c) {
        const unsigned* b = bal_sm[wib][c];
        const unsigned long long lo = (unsigned long long)b[w] | ((unsigned long long)b[w + 1] << 32);
        unsigned long long r = lo >> sh;
        if (sh) r |= (unsigned long long)b[w + 2] << (64 - sh);
        if (M < 64) r &= (1ull << M) - 1ull;
        m3[c] = ((r & ((1ull << (M - 1)) - 1ull)) << 1) | ((r >> (M - 1)) & 1ull);
      }
      W.mpos[my_id] = m3[0]; W.mneg[my_id] = m3[1]; W.mzero[my_id] = m3[2];
    }
  }
  n_full = warp_sum(n_full);
  if (lane == 0 && stats && n_full) atomicAdd(&stats->n_cond_eval, n_full);
}

// Rebuilds the transient view of a pending step from the hot record (optimistically advanced) and the
// published segment (state at the start of the step).
template <int IG>
__device__ __forceinline__ void wf_load_pending(Ray& R, const SegAoS D) {
  for (int i = 0; i < 6; ++i) {
    R.un[i] = R.u[i];
    R.u[i] = D(Seg<IG>::U + i);
  }
  if constexpr (Scheme<IG>::BL) {
    R.dn = R.delta;
    R.delta = D(Seg<IG>::DELTA);
  }
  R.s_new = R.s;
  R.s = D(Seg<IG>::S);
  R.h = D(Seg<IG>::H);
  R.dt_new = R.dt;
}

// ---- events: detection over the pending rays; no crossing ⇒ the optimistic advance stands ------------
template <int IG>
__global__ void __launch_bounds__(WF_THREADS)
k_wf_detect(const __grid_constant__ DevParams P, WfPool W, int k) {
  const long long n = (long long)*wc(W.ctr, k, WC_PEND);
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  const long long nround = ((n + nthreads - 1) / nthreads) * nthreads;
  // (the id of this thread's next trip is kept in flight and that ray's hot record and sign masks are pulled into L2
  // while the current ray is being looked at: the kernel waits on HBM, long scoreboard 41 cycles per issue)
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  int id_next = t < n ? W.pend[t] : -1;
  for (; t < nround; t += nthreads) {
    const bool valid = t < n;
    bool to_root = false, done = false, axion = false;
    int id = id_next;
    id_next = t + nthreads < n ? W.pend[t + nthreads] : -1;
    if (id_next >= 0) {
      const char* nx = (const char*)(W.hot + id_next);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + 128));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(W.mpos + id_next));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(W.mneg + id_next));
    }
    if (valid) {
      Ray R;
      static_cast<RayHot&>(R) = W.hot[id];
      axion = R.axion;
      const ScanMasks m{W.mpos[id], W.mneg[id], W.mzero[id]};
      const SegAoS D{W.seg + (size_t)id * Seg<IG>::N};
      // detect() only reads the start-box reference point from the cold record (first crossing only)
      if (R.n_cross == 0) {
        const RayCold& C = W.cold[id];
        R.x0[0] = C.x0[0]; R.x0[1] = C.x0[1]; R.x0[2] = C.x0[2];
      }
      int ia, ib, sgl;
      const int ev = R.template detect<IG>(P, D, m, ia, ib, sgl);
      if (ev != 0) {
        RayCold& C = W.cold[id];
        C.br_a = ia; C.br_b = (ev == 2) ? -2 : ib; C.br_sg = sgl;
        to_root = true;
      } else {
        // apply(th_ev < 0): state, FSAL derivative, s and dt were advanced by k_wf_stages
        R.sgn_end = sgl;
        R.event_last = false;
        R.settle(P);
        done = R.done;
        RayHot& H = W.hot[id];
        H.sgn_end = R.sgn_end; H.status = R.status;
        H.event_last = false; H.done = R.done;
      }
    }
    const long long sr = warp_pop(wc(W.ctr, k, WC_ROOTS), to_root);
    if (to_root) ((k & 1) ? W.roots[1] : W.roots[0])[sr] = id;
    wf_route(W, k, valid && !to_root, id, done, false, axion);
  }
}

// ---- root finder (all lanes iterate together), crossing callback, state update ------------------------
template <int IG>
__global__ void __launch_bounds__(WF_THREADS)
k_wf_root(const __grid_constant__ DevParams P, WfPool W, int k, DevStats* stats) {
  // side pass of iteration k: the roots list iteration k − 1 produced; survivors join iteration k + 1
  unsigned long long n_full = 0;
  const long long n = (long long)*wc(W.ctr, k - 1, WC_ROOTS);
  const int* __restrict__ roots = ((k - 1) & 1) ? W.roots[1] : W.roots[0];
  const int spread = wf_spread(n);
  const long long nthreads = (long long)gridDim.x * blockDim.x / spread;
  const long long nround = ((n + nthreads - 1) / nthreads) * nthreads;
  const long long t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  for (long long t = t0 / spread; t < nround; t += nthreads) {
    const bool valid = t < n && (t0 % spread) == 0;
    bool done = false, axion = false;
    int id = -1;
    if (valid) {
      id = roots[t];
      Ray R;
      static_cast<RayHot&>(R) = W.hot[id];
      static_cast<RayCold&>(R) = W.cold[id];
      constexpr int SEG_PAD = Seg<IG>::N;
      double segl[SEG_PAD];  // the root finder evaluates the segment ~12 times: keep it on chip
      {
        const double* sgp = W.seg + (size_t)id * SEG_PAD;
#pragma unroll
        for (int k = 0; k < SEG_PAD; ++k) segl[k] = sgp[k];
      }
      const SegAoS D{segl};
      wf_load_pending<IG>(R, D);
      const K7View K{R.k1, R.k1x};
      const int c0 = R.n_cond;
      const double th = (R.br_b == -2) ? 1.0 : R.template locate<IG>(P, D, R.br_a, R.br_b);
      n_full += (unsigned long long)(R.n_cond - c0);
      R.template apply<IG>(P, K, D, th, R.br_sg);
      W.hot[id] = static_cast<const RayHot&>(R);
      if constexpr (Scheme<IG>::BL) {
        W.cold[id].delta = R.delta;
        W.cold[id].k1x = R.k1x;
      }
      done = R.done;
      axion = R.axion;
    }
    wf_route(W, k, valid, id, done, false, axion);
  }
  n_full = warp_sum(n_full);
  if ((threadIdx.x & 31) == 0 && stats && n_full) atomicAdd(&stats->n_cond_eval, n_full);
}

// result of a finished ray
__device__ __forceinline__ RayOut wf_result(const DevParams& P, const RayHot& H, const RayCold& C, WfStat& st) {
  Ray R;
  static_cast<RayHot&>(R) = H;
  R.delta = P.bl > 0.0 ? C.delta : 0.0;
  RayOut o;
  R.finish(P, o);
  st.add(o, R.axion);
  return o;
}

// plain ray batches (art_propagate): results of the finished rays
__global__ void __launch_bounds__(WF_THREADS)
k_wf_collect(const __grid_constant__ DevParams P, WfPool W, int k, RayOut* __restrict__ out, DevStats* stats) {
  WfStat st;
  const long long n = (long long)*wc(W.ctr, k - 1, WC_FIN);
  const int* __restrict__ fin = ((k - 1) & 1) ? W.fin[1] : W.fin[0];
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int id = fin[t];
    out[id] = wf_result(P, W.hot[id], W.cold[id], st);
  }
  st.flush(stats);
}

// ---- tree: finished rays → get_tree loop body → next node re-initialised in the same slot ------------
// One half of an event (its back-trace, or its forward tree) is through: the second one to arrive hands the slot back.
__device__ __forceinline__ void wf_release_half(const WfTree& T, WfPool& W, int slot) {
  if (atomicAdd(T.slot_parts + slot, 1) == 1) {
    T.slot_parts[slot] = 0;
    const unsigned long long at = atomicAdd(W.ctr + WC_RING_TAIL, 1ull);
    T.ring[at % (unsigned long long)T.n_slots] = slot;
  }
}

__device__ __forceinline__ void wf_log_node(const NodeLog& L, int event, int pos, int species, double weight, double prob, double parent_weight,
                                            const double x0[3], const double k0[3], double t0, double dw0, double erg, int n_cross, const double* rc) {
  const unsigned long long at = atomicAdd(L.count, 1ull);
  if ((long long)at >= L.cap) return;
  NodeLogEntry n;
  n.event = event; n.pos = pos; n.species = species; n.n_cross = n_cross < L.cross_cap ? n_cross : L.cross_cap;
  n.weight = weight; n.prob = prob; n.parent_weight = parent_weight;
  for (int c = 0; c < 3; ++c) { n.x0[c] = x0[c]; n.k0[c] = k0[c]; }
  n.t0 = t0; n.dw0 = dw0; n.erg = erg;
  L.e[at] = n;
  double* q = L.cross + at * (unsigned long long)L.cross_cap * 4;
  for (int j = 0; j < n.n_cross; ++j) {
    q[4 * j] = rc[8 * j]; q[4 * j + 1] = rc[8 * j + 1]; q[4 * j + 2] = rc[8 * j + 2]; q[4 * j + 3] = rc[8 * j + 6];
  }
}

// Returns true when the event continues: R then holds the freshly initialised ray of the next node.
__device__ __forceinline__ bool wf_tree_step(const DevParams& P, const WfTree& T, WfPool& W, int id, const RayOut& ro,
                                             double* __restrict__ rec, int rec_cap, Ray& R) {
  const double* rc = rec + (size_t)id * rec_cap * 8;
  if (id < T.n_slots) {  // back-traced axion of the event in slot id: samp_back_weight
    const int e = T.slot_event[id];
    EventData E = T.ev[e];
    tree_backward(P, E, ro, rc, rec_cap);
    T.ev[e] = E;
    if (T.nlog.e) {  // saveNode(f, nb), MR.jl:612: the parent axion of the back-trace with all its crossings
      const double kb[3] = {-E.k_init[0], -E.k_init[1], -E.k_init[2]};
      wf_log_node(T.nlog, e, 0, SPECIES_AXION, E.back_node_w, E.prob0, -1.0, E.x, kb, 0.0, -1.0, E.erg_inf, ro.n_cross, rc);
    }
    wf_release_half(T, W, id);
    return false;
  }
  const int slot = id - T.n_slots;
  const int e = T.slot_event[slot];
  const EventData E = T.ev[e];  // (the back-trace only writes back_weight / prob0 / c_bck, which the forward tree does not read)
  TreeState S = T.ts[slot];
  if (T.nlog.e) {  // saveNode(f, tree[ii]), MR.jl:669-670
    const TreeNode& n = S.cur;
    const double x0[3] = {n.x, n.y, n.z}, k0[3] = {n.kx, n.ky, n.kz};
    wf_log_node(T.nlog, e, S.tree_len + 1, n.species, n.weight, n.prob, n.parent_weight, x0, k0, n.t, n.dw, E.erg_inf, ro.n_cross, rc);
  }
  tree_forward_update(P, E, S, T.stacks + (size_t)slot * T.stack_stride, ro, rc, T.seed,
                      (unsigned long long)T.ids.id(e), (long long)e, T.sink);
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
  T.ts[slot] = S;
  if (!spawn) wf_release_half(T, W, slot);
  return spawn;
}

template <int IG>
__global__ void __launch_bounds__(WF_THREADS)
k_wf_tree(const __grid_constant__ DevParams P, WfPool W, WfTree T, int k, double* __restrict__ rec, int rec_cap, DevStats* stats) {
  // side pass of iteration k: the finished list iteration k − 1 produced; spawned nodes join iteration k + 1
  WfStat st;
  unsigned long long n_start = 0;
  const long long n = (long long)*wc(W.ctr, k - 1, WC_FIN);
  const int* __restrict__ fin = ((k - 1) & 1) ? W.fin[1] : W.fin[0];
  const int spread = wf_spread(n);
  const long long nthreads = (long long)gridDim.x * blockDim.x / spread;
  const long long nround = ((n + nthreads - 1) / nthreads) * nthreads;
  const long long t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  for (long long t = t0 / spread; t < nround; t += nthreads) {
    Ray R;
    R.done = true;
    R.axion = false;
    bool spawn = false;
    int id = -1;
    if (t < n && (t0 % spread) == 0) {
      id = fin[t];
      const RayOut o = wf_result(P, W.hot[id], W.cold[id], st);
      spawn = wf_tree_step(P, T, W, id, o, rec, rec_cap, R);
      if (spawn) {
        // a re-initialised ray that is already past the end of the span stays `done`: the next
        // k_wf_stages launch routes it back here (never happens for t < 1/Ω)
        if (!R.done && R.iter_guard(P)) {
          R.template startup<IG, -1>(P);
          n_start++;
        }
        W.hot[id] = static_cast<const RayHot&>(R);
        W.cold[id] = static_cast<const RayCold&>(R);
      }
    }
    // spawned rays join the next step's active list even when already done (see above)
    wf_route(W, k, spawn, id, false, false, R.axion);
  }
  st.flush(stats);
  n_start = warp_sum(n_start);
  if ((threadIdx.x & 31) == 0 && stats && n_start) atomicAdd(&stats->n_cond_eval, n_start);
}

// ---- injection of freshly sampled events into free slots (side pass of iteration k) -------------------
// Sample j of the injection (position (c0 + j) % fifo of the sample ring the sampler stream fills ahead of the loop, in
// completion order; the record names its event) takes the slot ring[(head + j) % S]: per-sample weights
// (MainRunner.jl:498-557), first tree node, and the IC transform / first right-hand side / initial step size /
// starting sign of its two rays — the back-traced axion in ray slot s (MR.jl:581-589: −k, t = 0, Δω = −1, all
// crossings) and the forward photon in ray slot S + s (MR.jl:653-657).  Both join iteration k + 1.
template <int IG>
__global__ void __launch_bounds__(WF_THREADS)
k_wf_inject(const __grid_constant__ DevParams P, WfPool W, WfTree T, int k, const EventSample* __restrict__ samples, long long fifo, int m,
            long long c0, unsigned long long head, double* __restrict__ rec, int rec_cap, DevStats* stats) {
  unsigned long long n_start = 0;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  const long long n2 = 2LL * m;
  const long long nround = ((n2 + nthreads - 1) / nthreads) * nthreads;
  // thread t < m: event t's weights + back-trace ray; thread m + t: its forward ray (reads the weights again: cheap
  // next to two right-hand sides, and it keeps the two rays of an event on different threads)
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < nround; t += nthreads) {
    Ray R;
    R.done = true;
    R.axion = false;
    const bool valid = t < n2;
    int id = -1;
    if (valid) {
      const bool fwd = t >= m;
      const int j = (int)(fwd ? t - m : t);
      const int slot = T.ring[(head + (unsigned long long)j) % (unsigned long long)T.n_slots];
      const EventSample smp = samples[(c0 + j) % fifo];
      const int e = (int)smp.e;
      const SampleWeights Wt = sample_weights(P, smp.x, smp.vloc, smp.vinf);
      RayIn in;
      for (int c = 0; c < 3; ++c) { in.x0[c] = smp.x[c]; in.k0[c] = fwd ? Wt.k_init[c] : -Wt.k_init[c]; }
      in.erg = Wt.erg_inf; in.t_start = 0.0; in.dOmega = -1.0;
      in.max_cross = fwd ? -1 : 100000;
      in.species = fwd ? SPECIES_PHOTON : SPECIES_AXION;
      if (!fwd) {
        EventData E;
        for (int c = 0; c < 3; ++c) { E.x[c] = smp.x[c]; E.k_init[c] = Wt.k_init[c]; E.vinf[c] = smp.vinf[c]; }
        E.erg_inf = Wt.erg_inf; E.vel_eng = Wt.vel_eng; E.sln_prob = Wt.sln_prob; E.cos_w = Wt.cos_w;
        E.back_weight = 0.0; E.prob0 = 0.0; E.c_bck = 0; E.n_failed = smp.n_failed;
        T.ev[e] = E;
        T.slot_event[slot] = e;
        T.slot_parts[slot] = 0;
      } else {
        EventData E;  // what tree_forward_start reads
        for (int c = 0; c < 3; ++c) { E.x[c] = smp.x[c]; E.k_init[c] = Wt.k_init[c]; }
        E.erg_inf = Wt.erg_inf;
        TreeState S;
        tree_forward_start(P, E, S);
        S.back_status = 0;
        T.ts[slot] = S;
      }
      id = fwd ? T.n_slots + slot : slot;
      R.init(P, in, rec + (size_t)id * rec_cap * 8, rec_cap);
      if (!R.done && R.iter_guard(P)) {
        R.template startup<IG, -1>(P);
        n_start++;
      }
      W.hot[id] = static_cast<const RayHot&>(R);
      W.cold[id] = static_cast<const RayCold&>(R);
    }
    wf_route(W, k, valid, id, false, false, R.axion);
  }
  n_start = warp_sum(n_start);
  if ((threadIdx.x & 31) == 0 && stats && n_start) atomicAdd(&stats->n_cond_eval, n_start);
}

__global__ void k_ring_init(int* ring, int* parts, int n, unsigned long long* ctr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    ring[i] = i;
    parts[i] = 0;
  }
  if (i == 0) ctr[WC_RING_TAIL] = (unsigned long long)n;
}

}  // namespace art
