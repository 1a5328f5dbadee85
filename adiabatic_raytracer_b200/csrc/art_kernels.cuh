// art_kernels.cuh — sm_100a kernels of the hot path.
//
//   K1 k_sample_events    surface sampler, persistent threads, one attempt per lane, events
//                         pulled from a device queue by warp-aggregated atomics (FP64 ALU)
//   K2 k_weights          per-sample flux weights (FP64, pointwise)
//   K3 k_wf_*             adaptive ray integrator, step-batched over compacted ray lists
//                         (art_wavefront.cuh) (FP64 ALU)
//   K4 k_wf_tree          get_tree loop body per finished ray, next node spawned in place
//   K5 k_hist             binned flux with shared-memory privatised bins (atomics)
//
// None of this has a dense contraction: tensor cores / TMA are deliberately not used
// (BASELINE.json north_star); the rules that matter are register residency, grid = SM count ×
// resident CTAs, coalesced SoA/record I/O once per ray.
#pragma once
#include <cuda_runtime.h>

#include "art_sampler.cuh"
#include "art_tracer.cuh"
#include "art_tree.cuh"

namespace art {

constexpr unsigned FULL = 0xffffffffu;
constexpr int SAMPLE_THREADS = 128;

struct DevStats {  // accumulated with atomics; mirrors the counters of art_stats
  unsigned long long n_attempts, n_sampler_cond, n_rays, n_rays_photon, n_acc, n_rej, n_rhs, n_cond;
  unsigned long long n_status[5];
  unsigned long long n_cond_eval;  // full evaluations of the crossing condition (scan + root finder + start-up)
  unsigned long long n_scanned;    // accepted steps handed to the crossing scan
};

struct QueuedRay {  // one node to propagate
  RayIn in;
  int owner;  // local event index (fused pipeline) or caller's ray index (art_propagate)
  int pad_;
};

// Pops one work item per requesting lane with a single atomic per warp.
__device__ __forceinline__ long long warp_pop(unsigned long long* head, bool want) {
  const unsigned m = __ballot_sync(FULL, want);
  if (m == 0) return -1;
  const unsigned lane = threadIdx.x & 31;
  const int leader = __ffs(m) - 1;
  unsigned long long base = 0;
  if ((int)lane == leader) base = atomicAdd(head, (unsigned long long)__popc(m));
  base = __shfl_sync(FULL, base, leader);
  return want ? (long long)(base + __popc(m & ((1u << lane) - 1u))) : -1;
}

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULL, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------
// K1: surface sampler (fused pipeline): event e draws attempts j = 0,1,… until accepted.
// ------------------------------------------------------------------------------------------
struct EventSample {
  double x[3], vloc[3], vinf[3];
  int n_cross, n_failed;
  long long e;  // index of the event in the call (compact output: events complete out of order)
};

// An event whose attempt failed when the launch had no fresh work left: the next launch goes on with attempt `attempt`.
struct SampleCarry {
  long long idx;
  unsigned attempt, pad_;
};

#ifndef ART_SAMPLE_MINBLOCKS
#define ART_SAMPLE_MINBLOCKS 5
#endif
// Event ids of a pipeline call: event e (index in the call) has id first + (e / bsize) bstride + e % bsize (batches of a
// pipelined job, art_run_device_batches); a plain range is bsize = bstride = anything ≥ the number of events.
struct EventIds {
  long long first, bsize, bstride;
  __host__ __device__ long long id(long long e) const { return first + (e / bsize) * bstride + e % bsize; }
};

// Work counters of one sampler launch (zeroed before it)
enum { SC_HEAD = 0, SC_OUT = 1, SC_CARRY = 2 };

// One launch works through n_carry carried events (carry_in) and then the fresh events e0 … e0 + n − 1 of the call.
//   compact = 0 (art_sample): the sample of event e0 + i goes to out[i]; every event is drawn until it is accepted.
//   compact = 1 (fused pipeline): accepted samples are APPENDED to the ring `out` (ring entries, first free position
//   out_base) in completion order, each carrying its event index.  An event needs a geometric number of attempts (≈ 4.6 on
//   average, 50 for the unluckiest of 10⁵), one after the other: a launch that drew every event to the end spent most of its
//   warp time on a long tail of warps with one or two live lanes (12.9 of 32 lanes per instruction, profiles/r02_*).  So
//   once the work queue is empty, an event whose attempt fails is handed to carry_out instead and the launch ends within
//   one attempt; the next launch starts with the carried events.  `drain` (last launch of a job) switches that off.
// Attempt j of event e draws from the Philox stream (seed, id(e), j) wherever and whenever it runs: same samples.
// ONE instantiation and ONE call site of Sampler::begin serve both modes and every attempt, carried or not: the bits of a
// sample do not depend on where its attempt ran.
//
// Scheduling inside a warp.  The lanes of a warp are at different places of their attempts (art_sampler.cuh: SCAN trips,
// ROOT / NUDGE / INIT point evaluations, SKIP tests) or between attempts (result to process, next work item to pop, begin()).
// Every round the warp runs ONE kind of work — the one most of its lanes are waiting for — for exactly those lanes; the
// others sit the round out (no instruction slots, only latency, and nothing waits for a single event here).  Run as one loop
// in which every lane did whatever it needed, each kind ran at 2–9 lanes and the kernel averaged 10 of 32 lanes per
// instruction.
// (ART_SAMP_FIN_BIAS: head start of the between-attempts work in that vote — it is cheap and feeds the other kinds; measured per
// 3e6 events: bias 0 548 ms, 4 544, 12 554; five resident CTAs per SM (102 registers) instead of four 544 vs 548, three 556)
#ifndef ART_SAMP_FIN_BIAS
#define ART_SAMP_FIN_BIAS 4
#endif
__global__ void __launch_bounds__(SAMPLE_THREADS, ART_SAMPLE_MINBLOCKS)
k_sample_events(const __grid_constant__ DevParams P, unsigned long long seed, EventIds ids, long long e0, long long n,
                const SampleCarry* __restrict__ carry_in, long long n_carry, SampleCarry* __restrict__ carry_out, int compact,
                int drain, EventSample* __restrict__ out, long long ring, long long out_base, unsigned long long* ctr,
                DevStats* stats, unsigned long long* n_failed_total) {
  __shared__ double taus_sm[SAMP_KEEP * SAMPLE_THREADS];
  Sampler S;
  S.taus = taus_sm + threadIdx.x;
  S.taus_stride = SAMPLE_THREADS;
  S.done = true;
  S.mode = Sampler::SCAN;
  bool have = false, exhausted = false;
  long long idx = -1;
  unsigned attempt = 0;
  double u_idx = 0.0;
  unsigned long long c_att = 0, c_cond = 0, c_fail = 0;
  const long long n_total = n_carry + n;
  const bool may_carry = compact && !drain;
  const int M = P.M_samp;
  for (;;) {
    const bool running = have && !S.done;
    const int n_between = __popc(__ballot_sync(FULL, (have && S.done) || (!have && !exhausted)));
    const int n_scan = __popc(__ballot_sync(FULL, running && S.mode == Sampler::SCAN));
    const int n_point = __popc(__ballot_sync(FULL, running && S.in_point_mode()));
    const int n_skip = __popc(__ballot_sync(FULL, running && S.mode == Sampler::SKIP));
    if (n_between + n_scan + n_point + n_skip == 0) break;
    const int top = max(max(n_between + ART_SAMP_FIN_BIAS * (n_between > 0), n_scan), max(n_point, n_skip));
    if (n_between > 0 && n_between + ART_SAMP_FIN_BIAS >= top) {
      // ---- between attempts: result of the finished one, next work item, begin() ----
      const bool queue_empty = __any_sync(FULL, exhausted);
      bool emit = false, carry = false, fresh = false;
      EventSample es;
      if (have && S.done) {
        SampleResult R;
        S.finish(P, u_idx, R);
        c_att++;
        c_cond += R.n_cond;
        if (R.accepted) {
          for (int c = 0; c < 3; ++c) { es.x[c] = R.x[c]; es.vloc[c] = R.vloc[c]; es.vinf[c] = R.vinf[c]; }
          es.n_cross = R.n_cross;
          es.n_failed = (int)attempt;
          es.e = idx;
          c_fail += attempt;
          emit = true;
        } else {  // MainRunner.jl:463-494: try again (f_inx counts the failure)
          attempt++;
          if (may_carry && queue_empty) carry = true;
          else fresh = true;
        }
      }
      if (compact) {
        const long long po = warp_pop(ctr + SC_OUT, emit);
        if (emit) out[(out_base + po) % ring] = es;
        const long long pc = warp_pop(ctr + SC_CARRY, carry);
        if (carry) carry_out[pc] = SampleCarry{idx, attempt, 0u};
      } else if (emit) {
        out[idx - e0] = es;
      }
      if (emit || carry) have = false;
      const long long got = warp_pop(ctr + SC_HEAD, !have && !exhausted);
      if (!have && !exhausted) {
        if (got >= 0 && got < n_total) {
          if (got < n_carry) {
            idx = carry_in[got].idx;
            attempt = carry_in[got].attempt;
          } else {
            idx = e0 + (got - n_carry);
            attempt = 0;
          }
          have = true;
          fresh = true;
        } else {
          exhausted = true;
        }
      }
      if (fresh) {
        double U[10];
        sampler_uniforms(seed, (unsigned long long)ids.id(idx), attempt, U);
        u_idx = U[9];
        S.begin(P, U);
      }
    } else if (n_scan == top) {
      if (running && S.mode == Sampler::SCAN) S.scan_trip(P);
    } else if (n_skip == top) {
      if (running && S.mode == Sampler::SKIP) S.skip_certified_steps(P);
    } else {
      // ---- point evaluations: until every lane that was in a point mode has left it (a root takes ≈ 12) ----
      bool mine = running && S.in_point_mode();
#pragma unroll 1
      for (int k = 1; k <= M - 1; ++k) {
        const bool w = mine && !S.done && S.in_point_mode();
        if (!__any_sync(FULL, w)) break;
        if (w) S.point_slot(P);
      }
    }
  }
  c_att = warp_sum(c_att);
  c_cond = warp_sum(c_cond);
  c_fail = warp_sum(c_fail);
  if ((threadIdx.x & 31) == 0) {
    if (stats) {
      atomicAdd(&stats->n_attempts, c_att);
      atomicAdd(&stats->n_sampler_cond, c_cond);
    }
    if (n_failed_total && c_fail) atomicAdd(n_failed_total, c_fail);
  }
}

// fine-grained API: one attempt per thread from caller uniforms
__global__ void k_sample_uniforms(const __grid_constant__ DevParams P, long long n, const double* __restrict__ U,
                                  double* x, double* vloc, double* vinf, int* accepted, int* n_cross) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double u[10];
  for (int k = 0; k < 10; ++k) u[k] = U[i * 10 + k];
  SampleResult R;
  sample_attempt(P, u, R);
  accepted[i] = R.accepted;
  n_cross[i] = R.n_cross;
  for (int c = 0; c < 3; ++c) {
    x[c * n + i] = R.x[c];
    vloc[c * n + i] = R.vloc[c];
    vinf[c * n + i] = R.vinf[c];
  }
}

// ------------------------------------------------------------------------------------------
// K2: weights — fine-grained entry points (the fused pipeline computes them in k_wf_inject, art_wavefront.cuh)
// ------------------------------------------------------------------------------------------
__global__ void k_weights_api(const __grid_constant__ DevParams P, long long n, const double* x, const double* vloc,
                              const double* vinf, double* k_init, double* cos_w, double* jac, double* sln,
                              double* erg, double* vel) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double xx[3] = {x[i], x[n + i], x[2 * n + i]}, vl[3] = {vloc[i], vloc[n + i], vloc[2 * n + i]},
               vi[3] = {vinf[i], vinf[n + i], vinf[2 * n + i]};
  const SampleWeights W = sample_weights(P, xx, vl, vi);
  for (int c = 0; c < 3; ++c) k_init[c * n + i] = W.k_init[c];
  cos_w[i] = W.cos_w; jac[i] = W.jac_gr; sln[i] = W.sln_prob; erg[i] = W.erg_inf; vel[i] = W.vel_eng;
}

__global__ void k_conv_prob_api(const __grid_constant__ DevParams P, long long n, const double* x, const double* k,
                                const double* erg, double bsign, double* out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double xx[3] = {x[i], x[n + i], x[2 * n + i]}, kk[3] = {k[i], k[n + i], k[2 * n + i]};
  out[i] = conv_prob_nonAD(P, xx, kk, erg[i], bsign);
}

__global__ void k_rhs_api(const __grid_constant__ DevParams P, long long n, const int* species, const double* u,
                          const double* erg, const double* lnt, double* du) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* v = u + 7 * i;
  double o[6];
  const double t = exp(lnt[i]);
  if (species[i] == SPECIES_PHOTON) {
    double dd = 0.0;
    rhs_photon(P, v[0], v[1], v[2], v[3] * erg[i], v[4] * erg[i], v[5] * erg[i], -v[6], t, 1.0 / erg[i], o, P.bl > 0.0 ? &dd : nullptr);
    du[7 * i + 6] = P.Omega * erg[i] / C_KM * o[5] + dd;
  } else {
    rhs_axion(P, v[0], v[1], v[3] * erg[i], v[4] * erg[i], v[5] * erg[i], t, erg[i], 1.0 / erg[i], o);
    du[7 * i + 6] = 0.0;
  }
  for (int c = 0; c < 6; ++c) du[7 * i + c] = o[c];
}

// ------------------------------------------------------------------------------------------
// K4: tree
// ------------------------------------------------------------------------------------------
// rows are completed (back-traced weight, tree count / info) when their event is finished
__global__ void k_patch_rows(long long n_rows, double* rows, const long long* keys, int ncols, long long chunk_base,
                             const EventData* __restrict__ ev, const int* fin_count, const int* fin_info,
                             long long row_begin) {
  const long long i = row_begin + blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  const long long e = (keys[i] >> 8) - chunk_base;
  patch_row(rows + i * ncols, ncols, ev[e], fin_count[e], fin_info[e]);
}

__global__ void k_iota(int* idx, long long n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) idx[i] = (int)i;
}

__global__ void k_gather_rows(const double* __restrict__ rows, const int* __restrict__ order, long long n, int ncols,
                              double* __restrict__ out) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n * ncols) return;
  const long long r = t / ncols;
  const int c = (int)(t - r * ncols);
  out[t] = rows[(long long)order[r] * ncols + c];
}

// ------------------------------------------------------------------------------------------
// saveMode 3: trajectory of every logged tree node, sampled like `solve(…, saveat = (ln t_end − ln t_start)/(ntimes − 1))`
// (RT.jl:174, 383).  One thread per node re-traces it with the sequential form of the stepper (Ray::step_sampled: the same
// attempt / error control / event code the wavefront kernels run, so the same trajectory).  A debugging and plotting mode
// (plot/plotTree.py): a few thousand nodes, no attempt at speed.
// traj: [node][out_cap][4] = x y z ln t; npts[node].
// ------------------------------------------------------------------------------------------
constexpr int NODE_SCRATCH = K_ROWS + K_PARK + K_XROWS + SEG_MAX;  // doubles of per-node scratch (stage store + segment)
template <int IG>
__global__ void k_node_traj(const __grid_constant__ DevParams P, const NodeLogEntry* __restrict__ nodes, long long n, int ntimes,
                            double* __restrict__ traj, int* __restrict__ npts, int out_cap, double* __restrict__ rec_scratch,
                            double* __restrict__ k_scratch) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const NodeLogEntry nd = nodes[i];
  RayIn in;
  for (int c = 0; c < 3; ++c) { in.x0[c] = nd.x0[c]; in.k0[c] = nd.k0[c]; }
  in.erg = nd.erg; in.t_start = nd.t0; in.dOmega = nd.dw0;
  in.max_cross = nd.pos == 0 ? 100000 : -1;   // back-trace collects every crossing, forward nodes stop at the first (MR.jl:589, 663)
  in.species = nd.species;
  Ray R;
  const KStore K{k_scratch + (size_t)i * NODE_SCRATCH, 1};
  const SegStore D{k_scratch + (size_t)i * NODE_SCRATCH + K_ROWS + K_PARK + K_XROWS, 1};
  R.init(P, in, rec_scratch + (size_t)i * 64 * 8, 64);
  const double s0 = R.s;
  const double delta = (P.ln_tend - s0) / (double)(ntimes > 1 ? ntimes - 1 : 1);
  double* out = traj + (size_t)i * out_cap * 4;
  int n_out = 0, kq = 0;
  {  // the initial point is saved by every solve
    double x[3], k[3];
    state_to_cartesian(P, R.u, R.erg, false, x, k);
    out[0] = x[0]; out[1] = x[1]; out[2] = x[2]; out[3] = s0;
    n_out = 1;
    kq = 1;
  }
  while (!R.done) R.template step_sampled<IG>(P, K, D, s0, delta, ntimes, kq, out, n_out, out_cap);
  if (R.s < P.ln_tend && n_out < out_cap && out[4 * (n_out - 1) + 3] < R.s) {  // terminated early (NS hit): the last state is saved too
    double x[3], k[3];
    state_to_cartesian(P, R.u, R.erg, false, x, k);
    double* q = out + 4 * (size_t)n_out++;
    q[0] = x[0]; q[1] = x[1]; q[2] = x[2]; q[3] = R.s;
  }
  npts[i] = n_out;
}

// ------------------------------------------------------------------------------------------
// K5: binned flux / spectra.  bins: [species(2)][n_ct][n_ph][n_dw][4] = Σ w·sln, Σ (w·sln)², Σ w, count
// ------------------------------------------------------------------------------------------
__global__ void k_hist(long long row_begin, long long n_rows, const double* __restrict__ rows, int ncols, int n_ct,
                       int n_ph, int n_dw, double dw_lo, double dw_scale, double* __restrict__ hist, int use_smem) {
  extern __shared__ double sh[];
  const int nb = 2 * n_ct * n_ph * n_dw * 4;
  if (use_smem) {
    for (int k = threadIdx.x; k < nb; k += blockDim.x) sh[k] = 0.0;
    __syncthreads();
  }
  double* H = use_smem ? sh : hist;
  for (long long i = row_begin + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_rows;
       i += (long long)gridDim.x * blockDim.x) {
    const double* r = rows + i * ncols;
    const int id = (int)r[1];
    const double ct = cos(r[2]);
    int ic = (int)floor((ct + 1.0) * 0.5 * n_ct);
    ic = ic < 0 ? 0 : (ic >= n_ct ? n_ct - 1 : ic);
    int ip = (int)floor((r[3] + PI) / (2.0 * PI) * n_ph);
    ip = ip < 0 ? 0 : (ip >= n_ph ? n_ph - 1 : ip);
    const double w = r[8] * r[7];
    int iw = 0;
    if (n_dw > 1) {
      iw = (int)floor((r[12] - dw_lo) * dw_scale);
      iw = iw < 0 ? 0 : (iw >= n_dw ? n_dw - 1 : iw);
    }
    double* b = H + (((size_t)(id * n_ct + ic) * n_ph + ip) * n_dw + iw) * 4;
    atomicAdd(b + 0, w);
    atomicAdd(b + 1, w * w);
    atomicAdd(b + 2, r[8]);
    atomicAdd(b + 3, 1.0);
  }
  if (use_smem) {
    __syncthreads();
    for (int k = threadIdx.x; k < nb; k += blockDim.x)
      if (sh[k] != 0.0) atomicAdd(hist + k, sh[k]);
  }
}

// ------------------------------------------------------------------------------------------
// FP64 DFMA peak probe (roofline denominator; MEASURED_PEAKS.json carries no FP64 figure)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

}  // namespace art
