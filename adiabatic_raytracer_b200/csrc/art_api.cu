// art_api.cu — C ABI of include/art.h: handle, device memory, pipeline driver, NCCL reduce.
//
// The pipeline driver is the body of main_runner_tree's event loop (reference
// MainRunner.jl:450-747), restructured from "one event at a time on the host" into
// generations of batched device work over a chunk of events:
//   sample → weights → back-trace (axions) → back-finish → { trace photons, trace axions,
//   tree-update } × generations → (patch rows) → histogram.
// The host only reads two queue counters per generation.  There is no CPU compute path: every
// entry point that computes needs a CUDA device and fails with ART_ERR_CUDA otherwise.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/art.h"
#include "art_host.hpp"
#include "art_kernels.cuh"
#include "art_wavefront.cuh"

using namespace art;

namespace {

std::string g_create_error;

struct DBuf {  // grow-only device buffer
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() { return (T*)p; }
};

// ---- NCCL through dlopen: no build-time dependency; a Julia host picks up the system
// libnccl.so.2, a torchrun-launched bench the copy torch already loaded. -------------------
struct Id128 {  // ncclUniqueId is passed BY VALUE to ncclCommInitRank
  char b[128];
};
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, Id128, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

bool nccl_load(std::string& err) {
  if (g_nccl.lib) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) {
    err = std::string("cannot dlopen libnccl.so.2: ") + dlerror();
    return false;
  }
  g_nccl.GetUniqueId = (int (*)(void*))dlsym(g_nccl.lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(void**, int, Id128, int))dlsym(g_nccl.lib, "ncclCommInitRank");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(g_nccl.lib, "ncclAllReduce");
  g_nccl.CommDestroy = (int (*)(void*))dlsym(g_nccl.lib, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
    err = "libnccl lacks a required symbol";
    return false;
  }
  return true;
}

}  // namespace

// RAII for the CUDA events of one call (error paths return early: nothing may leak)
struct EvGuard {
  std::vector<cudaEvent_t> ev;
  cudaEvent_t make(unsigned flags = cudaEventDefault) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreateWithFlags(&e, flags) != cudaSuccess) return nullptr;
    ev.push_back(e);
    return e;
  }
  ~EvGuard() {
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
  }
};

// Working set of the wavefront loop: three streams (main: stages → scan → detect; side: root finder, tree and injection,
// one iteration behind; sampler: the next block of events, concurrently with the loop) and the ray / event-slot pools.
struct Lane {
  cudaStream_t stream = nullptr, stream2 = nullptr, stream3 = nullptr;
  cudaEvent_t ev_main = nullptr, ev_side = nullptr, ev_samp = nullptr;
  DBuf samples, tstate, stacks, q0, rout, rec, ctr;
  DBuf carry[2];                          // sampler carry-over lists (k_sample_events)
  unsigned long long* hctr = nullptr;     // page-locked copy of the sampler's counters, valid once ev_samp has completed
  DBuf slot_event, slot_parts, ring;
  DBuf wf_hot, wf_cold, wf_seg, wf_need, wf_neg, wf_pend, wf_sign, wf_act[4], wf_roots[2], wf_fin[2], wf_ctr;
  float ms_sampler = 0, ms_weights = 0, ms_tracer = 0;
  float ms_k[3] = {0, 0, 0};  // ART_TIME_KERNELS: stage kernel, scan, detect
  long long n_blocks = 0;
  void release() {
    DBuf* all[] = {&samples, &tstate, &stacks, &q0, &rout, &rec, &ctr, &carry[0], &carry[1], &slot_event, &slot_parts, &ring, &wf_hot, &wf_cold, &wf_seg,
                   &wf_need, &wf_neg, &wf_pend, &wf_sign, &wf_act[0], &wf_act[1], &wf_act[2], &wf_act[3], &wf_roots[0], &wf_roots[1],
                   &wf_fin[0], &wf_fin[1], &wf_ctr};
    for (DBuf* b : all) b->release();
    if (hctr) { cudaFreeHost(hctr); hctr = nullptr; }
    for (cudaStream_t* st : {&stream, &stream2, &stream3})
      if (*st) { cudaStreamDestroy(*st); *st = nullptr; }
    for (cudaEvent_t* e : {&ev_main, &ev_side, &ev_samp})
      if (*e) { cudaEventDestroy(*e); *e = nullptr; }
  }
};

struct art_handle {
  art_params params;
  DevParams dp;
  int device = 0, n_sm = 0;
  int blocks_sample = 0;
  cudaStream_t stream = nullptr;  // main stream (fine-grained entry points, patch / histogram / collectives)
  std::string err;
  std::mutex err_mu;
  std::vector<Lane> lanes;
  // shared by all lanes of a pipeline call
  DBuf events, fin_count, fin_info, gctr, stats;
  DBuf rows, keys, hist, hist_global, scratch[12];
  DBuf nlog, nlog_cross, nlog_count;  // saveMode 3: node log of the last pipeline call
  long long nlog_cap = 0;
  bool have_global = false;       // hist_global holds the result of an all-reduce of the CURRENT local block
  DBuf q[2];                      // fine-grained art_propagate queue
  long long rows_cap = 0;         // rows held on device
  long long n_rows_dev = 0;
  long long last_first_event = 0, last_n_events = 0;  // events of the last pipeline call (art_fetch_events)
  long long last_batch_events = 1, last_batch_stride = 1;
  int ncols = 13;
  int n_ct = 16, n_ph = 16, n_dw = 1;
  double dw_lo = 0.0, dw_hi = 0.0;
  double host_counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  void* comm = nullptr;
  int nranks = 1, rank = 0;
  std::atomic<long long> launches{0};  // kernels launched since the current pipeline call began
  int bulk_burst = 8;              // iterations between host checks while the pool is full (ART_BULK_BURST)
  int bulk_nstep = ART_STAGE_NSTEP;  // attempts per stage launch in the bulk (ART_BULK_NSTEP)
  int scan_grid_per_sm = ART_SCAN_MINBLOCKS;  // grid of the scan kernel = one resident wave (ART_SCAN_GRID_PER_SM; tracer per 1e6
                                  // events: 3 CTAs per SM 182.6 ms, 6 183.4, 8 185.3, 12 185.7, 24 189.0)
  int detect_grid_per_sm = 24;    // grid cap of the detect kernel (HBM-latency bound, 80 registers: 24 one-warp CTAs per SM are
                                  // one resident wave; ART_DETECT_GRID_PER_SM; tracer per 1e6 events: 4 → 189.5 ms, 8 → 183.2,
                                  // 16 → 181.8, 24 → 181.8, 48 → 180.5)
  int stage_grid_per_sm = ART_STAGE_MINBLOCKS;  // … of each species half of the stage kernel's grid: two resident waves in all (ART_STAGE_GRID_PER_SM)
  int wf_grid_per_sm = 8;         // grid cap of the wavefront kernels, CTAs per SM (ART_WF_GRID_PER_SM)
  long long tail_below = 0;       // rays in flight below which the stage kernel switches to its tail policy (ART_TAIL_BELOW; default 256 per SM)
  int bulk_min_go = 16;           // a warp goes on to its next attempt while this many lanes can (ART_BULK_MIN_GO)
  int tail_nstep = 4;             // attempts per stage launch once fewer rays than one resident wave are left (ART_TAIL_NSTEP)
  long long pool_slots = 1 << 20;  // event slots of the pool = events in flight at a time (ART_POOL_SLOTS); 2 rays per slot
  long long refill_min = 0;        // free slots at which sampled events are injected (ART_REFILL_MIN; default pool / 16)
  long long sample_block = 1 << 16; // smallest sampler launch, events (ART_SAMPLE_BLOCK)
};

namespace {

int fail(art_handle* h, int code, const std::string& msg) {
  if (h) {
    std::lock_guard<std::mutex> g(h->err_mu);
    h->err = msg;
  } else {
    g_create_error = msg;
  }
  return code;
}

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(h, ART_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));      \
  } while (0)

inline int nblk(long long n, int t) { return (int)((n + t - 1) / t); }

int upload(art_handle* h, DBuf& b, const void* src, size_t bytes) {
  CU(b.ensure(bytes ? bytes : 8));
  if (bytes) CU(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, h->stream));
  return ART_OK;
}

int ensure_lanes(art_handle* h, int n) {
  while ((int)h->lanes.size() < n) {
    h->lanes.emplace_back();
    CU(cudaStreamCreateWithFlags(&h->lanes.back().stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&h->lanes.back().stream2, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&h->lanes.back().stream3, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&h->lanes.back().ev_main, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->lanes.back().ev_side, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->lanes.back().ev_samp, cudaEventDisableTiming));
  }
  return ART_OK;
}

// Events the loop still has to start (fused pipeline): sampled block by block on the sampler stream while the
// loop runs, injected into free slots by the side pass.
struct PoolJob {
  uint64_t seed = 0;
  EventIds ids{0, 1, 1};
  long long n_events = 0;
  long long S = 0;                 // event slots
  unsigned long long head = 0;     // free-slot ring: slots [head, tail) are free; tail lives on the device
  // sample FIFO (capacity fifo = 2 blocks): events [0, produced) have been handed to the sampler stream, [0, sampled) are
  // known to be complete, [0, consumed) have been injected into the loop.  The sampler always works on whole blocks —
  // small launches leave its persistent threads one or two events each and cost three times as much per event.
  long long block = 0, fifo = 0, produced = 0, sampled = 0, consumed = 0;
  // An event needs a geometric number of attempts; a launch does not wait for its unluckiest events: once its work queue is
  // empty, failed attempts are carried over to the next launch (k_sample_events).  So `produced` counts the fresh events
  // handed out, `sampled` / `consumed` count completed samples in the ring (in completion order, each with its event index),
  // and carry_n events wait in carry list carry_cur.  The launch that takes the last fresh events draws everything to the end.
  long long carry_n = 0;
  int carry_cur = 0;
  bool busy = false;               // a sampler launch is in flight
  int m = 0;                       // events the next side pass injects (0: none)
  unsigned long long* gctr = nullptr;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> samp_ev;  // timing of the sampler launches
  long long n_blocks = 0, n_injections = 0;
};

int wf_alloc(art_handle* h, Lane& L, long long n_rays, int seg_doubles, WfPool& W) {
  CU(L.wf_hot.ensure((size_t)n_rays * sizeof(RayHot)));
  CU(L.wf_cold.ensure((size_t)n_rays * sizeof(RayCold)));
  CU(L.wf_seg.ensure((size_t)n_rays * seg_doubles * sizeof(double)));
  CU(L.wf_need.ensure((size_t)n_rays * 8));
  CU(L.wf_neg.ensure((size_t)n_rays * 8));
  CU(L.wf_sign.ensure((size_t)n_rays * 8));
  for (int k = 0; k < 4; ++k) CU(L.wf_act[k].ensure((size_t)n_rays * 4));
  CU(L.wf_pend.ensure((size_t)n_rays * 4));
  for (int k = 0; k < 2; ++k) {
    CU(L.wf_roots[k].ensure((size_t)n_rays * 4));
    CU(L.wf_fin[k].ensure((size_t)n_rays * 4));
  }
  CU(L.wf_ctr.ensure(WC_N * 8));
  W.hot = L.wf_hot.as<RayHot>(); W.cold = L.wf_cold.as<RayCold>(); W.seg = L.wf_seg.as<double>();
  W.mpos = L.wf_need.as<unsigned long long>(); W.mneg = L.wf_neg.as<unsigned long long>();
  W.mzero = L.wf_sign.as<unsigned long long>();
  for (int par = 0; par < 2; ++par)
    for (int sp = 0; sp < 2; ++sp) W.act[par][sp] = L.wf_act[par * 2 + sp].as<int>();
  W.pend = L.wf_pend.as<int>();
  for (int k = 0; k < 2; ++k) { W.roots[k] = L.wf_roots[k].as<int>(); W.fin[k] = L.wf_fin[k].as<int>(); }
  W.ctr = L.wf_ctr.as<unsigned long long>();
  return ART_OK;
}

// The wavefront loop (art_wavefront.cuh).  n_act0 rays are already on the active lists of iteration 0 (plain batches,
// art_propagate); with a tree context every finished ray walks its event's tree, and with a job the loop keeps
// starting new events in the slots that finished ones hand back.
template <int IG>
int wf_loop(art_handle* h, Lane& L, WfPool W, long long n_act0, RayOut* out, double* rec, int rec_cap, DevStats* stats,
            const WfTree* tree, PoolJob* job, EvGuard& evg) {
  cudaStream_t st = L.stream;
  const size_t ksmem = (size_t)(K_ROWS + K_PARK + (Scheme<IG>::BL ? K_XROWS : 0)) * STAGE_THREADS * sizeof(double);
  const int gmax = h->n_sm * h->wf_grid_per_sm;
  auto grid = [&](long long work, int threads) { return (int)std::max<long long>(1, std::min<long long>(gmax, (work + threads - 1) / threads)); };
  long long n_act = n_act0;  // upper bound of the rays in flight, refreshed at every host check
  const bool trace = getenv("ART_WF_TRACE") != nullptr;
  // ART_SIDE_STREAM=0 keeps everything on one stream (profilers serialise kernels and make stream waits costly)
  const bool side = !(getenv("ART_SIDE_STREAM") && atoi(getenv("ART_SIDE_STREAM")) == 0);
  cudaStream_t s2 = side ? L.stream2 : st;
  cudaStream_t s3 = side ? L.stream3 : st;
  const bool timek = getenv("ART_TIME_KERNELS") != nullptr;  // CUDA events around every launch of the main stream
  std::vector<cudaEvent_t> tk;
  const long long refill_min = job ? (h->refill_min > 0 ? h->refill_min : std::max<long long>(job->S / 16, 1)) : 0;
  long long iter_cap = 64LL * h->params.maxiters;
  if (job) iter_cap *= (job->n_events / std::max<long long>(job->S, 1) + 2);
  unsigned long long ring_tail = job ? (unsigned long long)job->S : 0ull;  // as of the last counter read (k_ring_init: all slots free)
  for (long long iter = 0;;) {
    // ---- host check: keep the sampler one block ahead; hand freshly sampled events to the next side pass ---------------
    if (job) {
      if (job->busy) {
        // nothing in flight and nothing to inject: wait for the block; otherwise pick it up as soon as it is there
        const bool must_wait = n_act == 0 && job->sampled == job->consumed;
        const cudaError_t q = must_wait ? cudaEventSynchronize(L.ev_samp) : cudaEventQuery(L.ev_samp);
        if (q == cudaSuccess) {
          job->sampled += (long long)L.hctr[SC_OUT];
          job->carry_n = (long long)L.hctr[SC_CARRY];
          job->carry_cur ^= 1;
          job->busy = false;
        } else if (q != cudaErrorNotReady) {
          return fail(h, ART_ERR_CUDA, std::string("sampler: ") + cudaGetErrorString(q));
        }
      }
      const long long m_next = std::min(job->block, job->n_events - job->produced);
      if (!job->busy && (m_next > 0 || job->carry_n > 0) && job->sampled - job->consumed + m_next + job->carry_n <= job->fifo) {
        const long long m = m_next;
        const int drain = job->produced + m >= job->n_events;
        cudaEvent_t ea = evg.make(), eb = evg.make();
        if (!ea || !eb) return fail(h, ART_ERR_CUDA, "cudaEventCreate");
        if (side) CU(cudaStreamWaitEvent(s3, L.ev_side, 0));  // (every injection enqueued so far has read its part of the FIFO)
        CU(cudaMemsetAsync(L.ctr.p, 0, 64, s3));
        CU(cudaEventRecord(ea, s3));
        const int blocks = (int)std::min<long long>(h->blocks_sample, (m + job->carry_n + SAMPLE_THREADS - 1) / SAMPLE_THREADS);
        k_sample_events<<<blocks, SAMPLE_THREADS, 0, s3>>>(h->dp, job->seed, job->ids, job->produced, m,
                                                                 L.carry[job->carry_cur].as<SampleCarry>(), job->carry_n,
                                                                 L.carry[job->carry_cur ^ 1].as<SampleCarry>(), 1, drain,
                                                                 L.samples.as<EventSample>(), job->fifo, job->sampled,
                                                                 L.ctr.as<unsigned long long>(), stats, job->gctr);
        CU(cudaGetLastError());
        CU(cudaEventRecord(eb, s3));
        CU(cudaMemcpyAsync(L.hctr, L.ctr.p, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s3));
        CU(cudaEventRecord(L.ev_samp, s3));
        job->samp_ev.emplace_back(ea, eb);
        job->n_blocks++;
        h->launches++;
        job->produced += m;
        job->busy = true;
        if (n_act == 0 && job->sampled == job->consumed) continue;  // (first block: wait for it at the top of the loop)
      }
      const long long n_free = (long long)(ring_tail - job->head);
      const long long avail = job->sampled - job->consumed;
      const long long left = job->n_events - job->consumed;
      job->m = 0;
      if (avail > 0 && n_free >= std::min(std::min(left, avail), refill_min)) job->m = (int)std::min(avail, n_free);
      if (n_act == 0 && job->m == 0 && !job->busy && left == 0) break;
    } else if (n_act == 0) {
      break;
    }
    const long long n_plan = n_act + (job ? 2LL * job->m : 0);  // … with the events about to be injected
    const int burst = n_plan > 500000 ? h->bulk_burst : (n_plan > 65536 ? 16 : 32);  // steps between host checks (deep queue: host jitter must not idle the GPU)
    const bool tail_mode = n_plan < h->tail_below;                        // thin tail: any lane keeps its warp stepping
    for (int b = 0; b < burst; ++b, ++iter) {
      const int k = (int)(iter & 0x3fffffff);
      // main stream: the side pass of iteration k − 1 has filled the active lists of iteration k
      if (side) {
        CU(cudaStreamWaitEvent(st, L.ev_side, 0));
        CU(cudaEventRecord(L.ev_main, st));   // detect(k − 1) is through: its roots / finished lists are complete
        CU(cudaStreamWaitEvent(s2, L.ev_main, 0));
      }
      // side stream: root finder, tree and injection on the lists iteration k − 1 produced, concurrent with stages(k)
      k_wf_root<IG><<<grid(n_act / 2 + 1, WF_THREADS), WF_THREADS, 0, s2>>>(h->dp, W, k, stats);
      if (tree)
        k_wf_tree<IG><<<grid(n_act / 2 + 1, WF_THREADS), WF_THREADS, 0, s2>>>(h->dp, W, *tree, k, rec, rec_cap, stats);
      else
        k_wf_collect<<<grid(n_act / 4 + 1, WF_THREADS), WF_THREADS, 0, s2>>>(h->dp, W, k, out, stats);
      h->launches += 2;
      if (job && job->m > 0) {
        k_wf_inject<IG><<<grid(2LL * job->m, WF_THREADS), WF_THREADS, 0, s2>>>(h->dp, W, *tree, k, L.samples.as<EventSample>(),
                                                                               job->fifo, job->m, job->consumed, job->head, rec,
                                                                               rec_cap, stats);
        h->launches++;
        job->head += (unsigned long long)job->m;
        job->consumed += job->m;
        job->n_injections++;
        n_act += 2LL * job->m;
        job->m = 0;
      }
      if (side) CU(cudaEventRecord(L.ev_side, s2));
      const int gs = (int)std::max<long long>(1, std::min<long long>((long long)h->n_sm * h->stage_grid_per_sm, (n_act + STAGE_THREADS - 1) / STAGE_THREADS));
      auto mark = [&]() {
        if (timek) {
          cudaEvent_t e = evg.make();
          if (e) cudaEventRecord(e, st);
          tk.push_back(e);
        }
      };
      mark();
      k_wf_stages<IG><<<2 * gs, STAGE_THREADS, ksmem, st>>>(h->dp, W, k, tail_mode ? h->tail_nstep : h->bulk_nstep, tail_mode ? 1 : h->bulk_min_go, stats);
      mark();
      k_wf_scan<IG><<<(int)std::max<long long>(1, std::min<long long>((long long)h->n_sm * h->scan_grid_per_sm, (n_act * 11 + SCAN_THREADS - 1) / SCAN_THREADS)),
                      SCAN_THREADS, 0, st>>>(h->dp, W, k, stats);
      mark();
      k_wf_detect<IG><<<(int)std::max<long long>(1, std::min<long long>((long long)h->n_sm * h->detect_grid_per_sm, (n_act + WF_THREADS - 1) / WF_THREADS)),
                        WF_THREADS, 0, st>>>(h->dp, W, k);
      mark();
      h->launches += 3;
    }
    CU(cudaGetLastError());
    if (side) CU(cudaStreamWaitEvent(st, L.ev_side, 0));
    unsigned long long c[WC_N];
    CU(cudaMemcpyAsync(c, W.ctr, sizeof(c), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    // next iteration: its active lists, plus the roots / finished lists of the last one (the side pass picks them up)
    const int kn = (int)(iter & 3), kl = (int)((iter - 1) & 3);
    const unsigned long long na0 = c[kn * WC_SET + WC_ACT], na1 = c[kn * WC_SET + WC_ACT + 1];
    n_act = (long long)(na0 + na1 + c[kl * WC_SET + WC_ROOTS] + c[kl * WC_SET + WC_FIN]);
    ring_tail = c[WC_RING_TAIL];
    if (trace) {
      timespec ts;
      clock_gettime(CLOCK_MONOTONIC, &ts);
      const double now = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
      static thread_local double t_prev = 0.0;
      fprintf(stderr, "[wf] iter %lld in flight %lld (ax %llu ph %llu roots %llu finished %llu) started %lld free %lld +%.2f ms = %.0f us/iteration\n",
              iter, n_act, na0, na1, c[kl * WC_SET + WC_ROOTS], c[kl * WC_SET + WC_FIN], job ? job->consumed : 0,
              job ? (long long)(c[WC_RING_TAIL] - job->head) : 0, t_prev > 0 ? now - t_prev : 0.0, t_prev > 0 ? (now - t_prev) * 1e3 / burst : 0.0);
      t_prev = now;
    }
    if (iter > iter_cap) return fail(h, ART_ERR_CUDA, "wavefront tracer did not terminate");
  }
  if (job) L.n_blocks += job->n_injections;
  for (size_t i = 0; i + 3 < tk.size(); i += 4)  // (the loop ends with a synchronised stream: every event has completed)
    for (int j = 0; j < 3; ++j) {
      float ms = 0;
      if (tk[i + j] && tk[i + j + 1] && cudaEventElapsedTime(&ms, tk[i + j], tk[i + j + 1]) == cudaSuccess) L.ms_k[j] += ms;
    }
  return ART_OK;
}

// step-batched tracer over a device queue of plain rays (art_propagate)
template <int IG>
int launch_trace_wf_t(art_handle* h, Lane& L, const QueuedRay* q, long long n, RayOut* out, double* rec, int rec_cap,
                      DevStats* stats) {
  if (n <= 0) return ART_OK;
  cudaStream_t st = L.stream;
  WfPool W;
  if (int rc = wf_alloc(h, L, n, Seg<IG>::N, W)) return rc;
  CU(cudaMemsetAsync(W.ctr, 0, WC_N * 8, st));
  const int gmax = h->n_sm * h->wf_grid_per_sm;
  k_wf_init<IG><<<(int)std::max<long long>(1, std::min<long long>(gmax, (n + WF_THREADS - 1) / WF_THREADS)), WF_THREADS, 0, st>>>(
      h->dp, q, n, W, rec, rec_cap, stats);
  h->launches++;
  CU(cudaGetLastError());
  CU(cudaEventRecord(L.ev_side, st));  // (the first iteration waits for "the previous side pass")
  EvGuard evg;
  return wf_loop<IG>(h, L, W, n, out, rec, rec_cap, stats, nullptr, nullptr, evg);
}

int launch_trace_wf(art_handle* h, Lane& L, const QueuedRay* q, long long n, RayOut* out, double* rec, int rec_cap,
                    DevStats* stats) {
  if (h->dp.integrator == IG_DP5) return launch_trace_wf_t<IG_DP5>(h, L, q, n, out, rec, rec_cap, stats);
  if (h->dp.bl > 0.0) return launch_trace_wf_t<IG_VERN6_BL>(h, L, q, n, out, rec, rec_cap, stats);
  return launch_trace_wf_t<IG_VERN6>(h, L, q, n, out, rec, rec_cap, stats);
}

}  // namespace

extern "C" {

const char* art_version(void) { return "adiabatic_raytracer_b200 0.2 (sm_100a; integrators: Vern6 [default], DP5)"; }

int art_params_default(art_params* p) {
  if (!p) return ART_ERR_ARG;
  memset(p, 0, sizeof(*p));
  p->Mass_a = 1e-5; p->Ax_g = 1e-12; p->theta_m = 0.0; p->omega_pul = 1.0; p->B0 = 1e14; p->rNS = 10.0; p->Mass_NS = 1.0;
  p->bndry_lyr = -1.0; p->ode_err = 1e-6; p->reltol = 1e-7; p->fix_time = 0.0; p->rho_DM = 0.45; p->vmean_ax = 220.0;
  p->prob_cutoff = 1e-10; p->ln_t_start = -30.0; p->t0_sample = 0.0; p->metric_interior_radius = 10.0;
  p->sampler_dt = 0.5; p->dtmin = 1e-13;
  p->flat = 0; p->isotropic = 0; p->melrose = 1; p->thick_surface = 1; p->n_maxSample = 6; p->num_cutoff = 5;
  p->MC_nodes = 5; p->max_nodes = 50; p->ntimes = 3; p->save_mode = 0; p->interp_points_ray = 50;
  p->interp_points_sampler = 20; p->maxiters = 100000; p->cross_cap = 16; p->integrator = ART_INTEGRATOR_VERN6;
  return ART_OK;
}

int art_max_radius(const art_params* p, double* maxR) {
  if (!p || !maxR) return ART_ERR_ARG;
  DevParams d;
  std::string e;
  int rc = make_dev_params(*p, d, e);
  if (rc == ART_OK || rc == ART_ERR_NO_SURFACE) *maxR = d.maxR;
  if (rc != ART_OK) g_create_error = e;
  return rc;
}

const char* art_last_error(const art_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int art_create(const art_params* p, int device_id, art_handle** out) {
  art_handle* h = nullptr;
  if (!p || !out) return fail(nullptr, ART_ERR_ARG, "null argument");
  *out = nullptr;
  DevParams d;
  std::string e;
  int rc = make_dev_params(*p, d, e);
  if (rc != ART_OK) return fail(nullptr, rc, e);
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    return fail(nullptr, ART_ERR_CUDA,
                std::string("no CUDA device (this library has no CPU fallback): ") + cudaGetErrorString(ce));
  if (device_id < 0 || device_id >= ndev) return fail(nullptr, ART_ERR_ARG, "device_id out of range");
  ce = cudaSetDevice(device_id);
  if (ce != cudaSuccess) return fail(nullptr, ART_ERR_CUDA, cudaGetErrorString(ce));
  h = new art_handle();
  h->params = *p;
  h->dp = d;
  h->device = device_id;
  h->ncols = p->save_mode > 1 ? ART_ROW_COLS_FULL : (p->save_mode > 0 ? ART_ROW_COLS_LONG : ART_ROW_COLS_SHORT);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device_id);
  h->n_sm = prop.multiProcessorCount;
  int occ = 0;
  cudaFuncSetAttribute(k_wf_stages<IG_VERN6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((K_ROWS + K_PARK) * STAGE_THREADS * sizeof(double)));
  cudaFuncSetAttribute(k_wf_stages<IG_VERN6>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(k_wf_stages<IG_DP5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((K_ROWS + K_PARK) * STAGE_THREADS * sizeof(double)));
  cudaFuncSetAttribute(k_wf_stages<IG_DP5>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(k_wf_stages<IG_VERN6_BL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((K_ROWS + K_PARK + K_XROWS) * STAGE_THREADS * sizeof(double)));
  cudaFuncSetAttribute(k_wf_stages<IG_VERN6_BL>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sample_events, SAMPLE_THREADS, 0);
  h->blocks_sample = h->n_sm * std::max(occ, 1);
  if (const char* t = getenv("ART_BULK_NSTEP")) h->bulk_nstep = std::max(1, atoi(t));
  if (const char* t = getenv("ART_BULK_BURST")) h->bulk_burst = std::max(1, std::min(64, atoi(t)));
  if (const char* t = getenv("ART_TAIL_NSTEP")) h->tail_nstep = std::max(1, atoi(t));
  if (const char* t = getenv("ART_BULK_MIN_GO")) h->bulk_min_go = std::max(1, std::min(32, atoi(t)));
  if (const char* t = getenv("ART_WF_GRID_PER_SM")) h->wf_grid_per_sm = std::max(1, atoi(t));
  if (const char* t = getenv("ART_SCAN_GRID_PER_SM")) h->scan_grid_per_sm = std::max(1, atoi(t));
  if (const char* t = getenv("ART_DETECT_GRID_PER_SM")) h->detect_grid_per_sm = std::max(1, atoi(t));
  if (const char* t = getenv("ART_STAGE_GRID_PER_SM")) h->stage_grid_per_sm = std::max(1, atoi(t));
  h->tail_below = 256LL * h->n_sm;
  if (const char* t = getenv("ART_TAIL_BELOW")) h->tail_below = std::max(0LL, atoll(t));
  if (const char* t = getenv("ART_POOL_SLOTS")) h->pool_slots = std::max(32LL, atoll(t));
  if (const char* t = getenv("ART_REFILL_MIN")) h->refill_min = std::max(1LL, atoll(t));
  if (const char* t = getenv("ART_SAMPLE_BLOCK")) h->sample_block = std::max(32LL, atoll(t));
  ce = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (ce != cudaSuccess) {
    std::string m = cudaGetErrorString(ce);
    delete h;
    return fail(nullptr, ART_ERR_CUDA, m);
  }
  *out = h;
  return ART_OK;
}

void art_destroy(art_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  for (Lane& L : h->lanes) L.release();
  DBuf* all[] = {&h->events, &h->fin_count, &h->fin_info, &h->gctr, &h->stats, &h->rows, &h->keys, &h->hist, &h->hist_global, &h->q[0], &h->q[1], &h->nlog, &h->nlog_cross, &h->nlog_count};
  for (DBuf* b : all) b->release();
  for (DBuf& b : h->scratch) b.release();
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

// ---- fine-grained entry points -------------------------------------------------------------

int art_sample_from_uniforms(art_handle* h, int64_t n, const double* u, double* x, double* v_loc, double* v_inf,
                             int32_t* accepted, int32_t* n_cross) {
  if (!h || n < 0) return ART_ERR_ARG;
  CU(cudaSetDevice(h->device));
  if (n == 0) return ART_OK;
  DBuf* s = h->scratch;
  if (upload(h, s[0], u, n * 10 * sizeof(double))) return ART_ERR_CUDA;
  for (int k = 1; k <= 3; ++k) CU(s[k].ensure(3 * n * sizeof(double)));
  CU(s[4].ensure(n * sizeof(int)));
  CU(s[5].ensure(n * sizeof(int)));
  k_sample_uniforms<<<nblk(n, 128), 128, 0, h->stream>>>(h->dp, n, s[0].as<double>(), s[1].as<double>(), s[2].as<double>(),
                                                         s[3].as<double>(), s[4].as<int>(), s[5].as<int>());
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(x, s[1].p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(v_loc, s[2].p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(v_inf, s[3].p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(accepted, s[4].p, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(n_cross, s[5].p, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return ART_OK;
}

int art_sample(art_handle* h, uint64_t seed, int64_t first_event, int64_t n, double* x, double* v_loc, double* v_inf,
               int32_t* n_cross, int32_t* n_failed) {
  if (!h || n < 0) return ART_ERR_ARG;
  CU(cudaSetDevice(h->device));
  if (n == 0) return ART_OK;
  if (int rc = ensure_lanes(h, 1)) return rc;
  Lane& L = h->lanes[0];
  CU(L.samples.ensure(n * sizeof(EventSample)));
  CU(L.ctr.ensure(64));
  CU(cudaMemsetAsync(L.ctr.p, 0, 64, h->stream));
  const int blocks = (int)std::min<long long>(h->blocks_sample, (n + SAMPLE_THREADS - 1) / SAMPLE_THREADS);
  k_sample_events<<<blocks, SAMPLE_THREADS, 0, h->stream>>>(h->dp, seed, EventIds{first_event, n, n}, 0, n, nullptr, 0, nullptr, 0, 1,
                                                                   L.samples.as<EventSample>(), n, 0, L.ctr.as<unsigned long long>(),
                                                                   nullptr, nullptr);
  CU(cudaGetLastError());
  std::vector<EventSample> host((size_t)n);
  CU(cudaMemcpyAsync(host.data(), L.samples.p, n * sizeof(EventSample), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  for (int64_t i = 0; i < n; ++i) {
    for (int c = 0; c < 3; ++c) {
      x[c * n + i] = host[i].x[c];
      v_loc[c * n + i] = host[i].vloc[c];
      v_inf[c * n + i] = host[i].vinf[c];
    }
    n_cross[i] = host[i].n_cross;
    n_failed[i] = host[i].n_failed;
  }
  return ART_OK;
}

int art_weights(art_handle* h, int64_t n, const double* x, const double* v_loc, const double* v_inf, double* k_init,
                double* cos_w, double* jac_gr, double* sln_prob, double* erg_inf, double* vel_eng) {
  if (!h || n < 0) return ART_ERR_ARG;
  CU(cudaSetDevice(h->device));
  if (n == 0) return ART_OK;
  DBuf* s = h->scratch;
  if (upload(h, s[0], x, 3 * n * 8) || upload(h, s[1], v_loc, 3 * n * 8) || upload(h, s[2], v_inf, 3 * n * 8)) return ART_ERR_CUDA;
  CU(s[3].ensure(3 * n * 8));
  for (int k = 4; k <= 8; ++k) CU(s[k].ensure(n * 8));
  k_weights_api<<<nblk(n, 128), 128, 0, h->stream>>>(h->dp, n, s[0].as<double>(), s[1].as<double>(), s[2].as<double>(),
                                                     s[3].as<double>(), s[4].as<double>(), s[5].as<double>(),
                                                     s[6].as<double>(), s[7].as<double>(), s[8].as<double>());
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(k_init, s[3].p, 3 * n * 8, cudaMemcpyDeviceToHost, h->stream));
  double* outs[5] = {cos_w, jac_gr, sln_prob, erg_inf, vel_eng};
  for (int k = 0; k < 5; ++k) CU(cudaMemcpyAsync(outs[k], s[4 + k].p, n * 8, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return ART_OK;
}

int art_conv_prob(art_handle* h, int64_t n, const double* x, const double* k, const double* erg_scaled, int32_t b_sign,
                  double* P_nonAD) {
  if (!h || n < 0) return ART_ERR_ARG;
  CU(cudaSetDevice(h->device));
  if (n == 0) return ART_OK;
  DBuf* s = h->scratch;
  if (upload(h, s[0], x, 3 * n * 8) || upload(h, s[1], k, 3 * n * 8) || upload(h, s[2], erg_scaled, n * 8)) return ART_ERR_CUDA;
  CU(s[3].ensure(n * 8));
  k_conv_prob_api<<<nblk(n, 128), 128, 0, h->stream>>>(h->dp, n, s[0].as<double>(), s[1].as<double>(), s[2].as<double>(),
                                                       b_sign < 0 ? -1.0 : 1.0, s[3].as<double>());
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(P_nonAD, s[3].p, n * 8, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return ART_OK;
}

int art_rhs(art_handle* h, int64_t n, const int32_t* species, const double* u, const double* erg, const double* lnt,
            double* du) {
  if (!h || n < 0) return ART_ERR_ARG;
  CU(cudaSetDevice(h->device));
  if (n == 0) return ART_OK;
  DBuf* s = h->scratch;
  if (upload(h, s[0], species, n * 4) || upload(h, s[1], u, 7 * n * 8) || upload(h, s[2], erg, n * 8) ||
      upload(h, s[3], lnt, n * 8))
    return ART_ERR_CUDA;
  CU(s[4].ensure(7 * n * 8));
  k_rhs_api<<<nblk(n, 128), 128, 0, h->stream>>>(h->dp, n, s[0].as<int>(), s[1].as<double>(), s[2].as<double>(),
                                                 s[3].as<double>(), s[4].as<double>());
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(du, s[4].p, 7 * n * 8, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return ART_OK;
}

int art_propagate(art_handle* h, int64_t n, const int32_t* species, const double* x0, const double* k0,
                  const double* erg, const double* t_start, const double* dOmega, int32_t b_sign,
                  int32_t max_crossings, double* x_f, double* k_f, double* u7_f, int32_t* status, int32_t* n_cross,
                  double* cross_rec, int32_t cross_cap, int32_t* counters) {
  (void)b_sign;  // the sign of B0 drops out of the ray equations and of the crossing condition
  if (!h || n < 0 || cross_cap < 1) return ART_ERR_ARG;
  CU(cudaSetDevice(h->device));
  if (n == 0) return ART_OK;
  std::vector<QueuedRay> qs((size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    QueuedRay& r = qs[(size_t)i];
    for (int c = 0; c < 3; ++c) { r.in.x0[c] = x0[c * n + i]; r.in.k0[c] = k0[c * n + i]; }
    r.in.erg = erg[i]; r.in.t_start = t_start[i]; r.in.dOmega = dOmega[i]; r.in.max_cross = max_crossings;
    r.in.species = species[i] == ART_SPECIES_PHOTON ? SPECIES_PHOTON : SPECIES_AXION;
    r.owner = (int)i; r.pad_ = 0;
  }
  if (int rc = ensure_lanes(h, 1)) return rc;
  Lane& L = h->lanes[0];
  if (upload(h, h->q[0], qs.data(), n * sizeof(QueuedRay))) return ART_ERR_CUDA;
  CU(L.rout.ensure(n * sizeof(RayOut)));
  CU(L.rec.ensure((size_t)n * cross_cap * 8 * sizeof(double)));
  CU(cudaMemsetAsync(L.rec.p, 0, (size_t)n * cross_cap * 8 * sizeof(double), h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (int rc = launch_trace_wf(h, L, h->q[0].as<QueuedRay>(), n, L.rout.as<RayOut>(), L.rec.as<double>(), cross_cap, nullptr))
    return rc;
  std::vector<RayOut> ro((size_t)n);
  CU(cudaMemcpyAsync(ro.data(), L.rout.p, n * sizeof(RayOut), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(cross_rec, L.rec.p, (size_t)n * cross_cap * 8 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  for (int64_t i = 0; i < n; ++i) {
    const RayOut& o = ro[(size_t)i];
    for (int c = 0; c < 3; ++c) { x_f[c * n + i] = o.xf[c]; k_f[c * n + i] = o.kf[c]; }
    u7_f[i] = o.u7f; status[i] = o.status; n_cross[i] = o.n_cross;
    if (counters) {
      counters[4 * i] = o.n_acc; counters[4 * i + 1] = o.n_rej; counters[4 * i + 2] = o.n_rhs; counters[4 * i + 3] = o.n_cond;
    }
  }
  return ART_OK;
}

// ---- fused pipeline ----------------------------------------------------------------------------

int art_hist_config(art_handle* h, int32_t n_costheta, int32_t n_phi) {
  return art_hist_config_spectrum(h, n_costheta, n_phi, 1, 0.0, 0.0);
}

int art_hist_config_spectrum(art_handle* h, int32_t n_costheta, int32_t n_phi, int32_t n_domega, double domega_lo,
                             double domega_hi) {
  if (!h || n_costheta < 1 || n_phi < 1 || n_domega < 1 || (long long)n_costheta * n_phi * n_domega > (1 << 20)) return ART_ERR_ARG;
  if (n_domega > 1 && !(domega_hi > domega_lo)) return ART_ERR_ARG;
  h->n_ct = n_costheta;
  h->n_ph = n_phi;
  h->n_dw = n_domega;
  h->dw_lo = domega_lo;
  h->dw_hi = domega_hi;
  return art_hist_reset(h);
}

int64_t art_hist_len(const art_handle* h) { return h ? (int64_t)2 * h->n_ct * h->n_ph * h->n_dw * 4 + 8 : 0; }

int art_hist_reset(art_handle* h) {
  if (!h) return ART_ERR_ARG;
  CU(cudaSetDevice(h->device));
  const size_t bytes = (size_t)art_hist_len(h) * sizeof(double);
  CU(h->hist.ensure(bytes));
  CU(cudaMemsetAsync(h->hist.p, 0, bytes, h->stream));
  for (double& c : h->host_counters) c = 0.0;
  h->have_global = false;
  CU(cudaStreamSynchronize(h->stream));
  return ART_OK;
}

static int hist_push_counters(art_handle* h) {
  const size_t off = (size_t)(art_hist_len(h) - 8) * sizeof(double);
  CU(cudaMemcpyAsync((char*)h->hist.p + off, h->host_counters, 8 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  return ART_OK;
}

// This GPU's block: bins accumulated by every pipeline call since the last reset + this GPU's trailing counters.
int art_hist_get(art_handle* h, double* out, int64_t cap) {
  if (!h || !out || cap < art_hist_len(h)) return ART_ERR_ARG;
  CU(cudaSetDevice(h->device));
  if (!h->hist.p) {
    int rc = art_hist_reset(h);
    if (rc) return rc;
  }
  int rc = hist_push_counters(h);
  if (rc) return rc;
  CU(cudaMemcpyAsync(out, h->hist.p, (size_t)art_hist_len(h) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return ART_OK;
}

// The all-reduced block of the last art_comm_allreduce_hist (sum over the ranks of what art_hist_get returns on each).
int art_hist_get_global(art_handle* h, double* out, int64_t cap) {
  if (!h || !out || cap < art_hist_len(h)) return ART_ERR_ARG;
  CU(cudaSetDevice(h->device));
  if (!h->have_global) return fail(h, ART_ERR_ARG, "art_hist_get_global: no all-reduce since the last pipeline call / reset");
  CU(cudaMemcpyAsync(out, h->hist_global.p, (size_t)art_hist_len(h) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return ART_OK;
}

}  // extern "C"

// The fused pipeline over the events of one call: one continuously refilled pool (art_wavefront.cuh).
template <int IG>
static int run_pool(art_handle* h, Lane& L, uint64_t seed, int64_t first_event, long long n_events, long long batch_size,
                    long long batch_stride, const RowSink& sink, unsigned long long* gctr, DevStats* dstats, EvGuard& evg) {
  const int cap = h->params.cross_cap;
  const int stack_stride = h->params.max_nodes + 4;
  const long long S = std::max<long long>(32, std::min<long long>(h->pool_slots, n_events));
  cudaStream_t st = L.stream;
  // sampler blocks: half a pool, at least 2^16 events (smaller launches are inefficient), never more than the call
  const long long block = std::min<long long>(n_events, std::max<long long>(S / 2, h->sample_block));
  const long long n_lanes = (long long)h->blocks_sample * SAMPLE_THREADS;  // (a launch carries at most one event per lane)
  CU(L.samples.ensure((2 * block + n_lanes) * sizeof(EventSample)));
  for (int c = 0; c < 2; ++c) CU(L.carry[c].ensure(n_lanes * sizeof(SampleCarry)));
  if (!L.hctr) CU(cudaHostAlloc((void**)&L.hctr, 8 * sizeof(unsigned long long), cudaHostAllocDefault));
  CU(L.tstate.ensure(S * sizeof(TreeState)));
  CU(L.stacks.ensure((size_t)S * stack_stride * sizeof(TreeNode)));
  CU(L.rec.ensure((size_t)2 * S * cap * 8 * sizeof(double)));
  CU(L.ctr.ensure(64));
  CU(L.slot_event.ensure(S * sizeof(int)));
  CU(L.slot_parts.ensure(S * sizeof(int)));
  CU(L.ring.ensure(S * sizeof(int)));
  WfPool W;
  if (int rc = wf_alloc(h, L, 2 * S, Seg<IG>::N, W)) return rc;
  CU(cudaMemsetAsync(W.ctr, 0, WC_N * 8, st));
  k_ring_init<<<nblk(S, 256), 256, 0, st>>>(L.ring.as<int>(), L.slot_parts.as<int>(), (int)S, W.ctr);
  h->launches++;
  CU(cudaGetLastError());
  WfTree T;
  T.ev = h->events.as<EventData>(); T.fin_count = h->fin_count.as<int>(); T.fin_info = h->fin_info.as<int>();
  T.ts = L.tstate.as<TreeState>(); T.stacks = L.stacks.as<TreeNode>(); T.slot_event = L.slot_event.as<int>();
  T.slot_parts = L.slot_parts.as<int>(); T.ring = L.ring.as<int>(); T.stack_stride = stack_stride; T.n_slots = (int)S;
  T.seed = seed; T.ids = EventIds{first_event, batch_size, batch_stride}; T.sink = sink;
  T.tot_count = gctr + 1;
  T.nlog = NodeLog{nullptr, nullptr, nullptr, 0, 0};
  if (h->params.save_mode >= 3) {
    T.nlog = NodeLog{h->nlog.as<NodeLogEntry>(), h->nlog_cross.as<double>(), h->nlog_count.as<unsigned long long>(), h->nlog_cap,
                     h->params.cross_cap};
    CU(cudaMemsetAsync(h->nlog_count.p, 0, 8, st));
  }
  PoolJob job;
  job.seed = seed; job.ids = T.ids; job.n_events = n_events; job.S = S; job.gctr = gctr + 0;
  job.block = block; job.fifo = 2 * block + n_lanes;
  cudaEvent_t t0 = evg.make(), t1 = evg.make();
  if (!t0 || !t1) return fail(h, ART_ERR_CUDA, "cudaEventCreate");
  CU(cudaEventRecord(t0, st));
  const int rc = wf_loop<IG>(h, L, W, 0, nullptr, L.rec.as<double>(), cap, dstats, &T, &job, evg);
  CU(cudaEventRecord(t1, st));
  // every stream is drained before anything is reported or freed, also on the error path
  cudaStreamSynchronize(L.stream3);
  cudaStreamSynchronize(L.stream2);
  CU(cudaStreamSynchronize(st));
  if (rc) return rc;
  float ms = 0;
  cudaEventElapsedTime(&ms, t0, t1);
  L.ms_tracer += ms;
  for (auto& pr : job.samp_ev) {
    float m2 = 0;
    cudaEventElapsedTime(&m2, pr.first, pr.second);
    L.ms_sampler += m2;
  }
  return ART_OK;
}

extern "C" {

int art_run_device_batches(art_handle* h, uint64_t seed, int64_t first_event, int64_t batch_events, int64_t n_batches,
                           int64_t batch_stride, int64_t* n_rows_out, int64_t* f_inx_out, int64_t* tot_count_out,
                           art_stats* st_out) {
  if (!h || batch_events < 0 || n_batches < 1 || batch_stride < batch_events) return ART_ERR_ARG;
  const long long n_events = batch_events * n_batches;
  if (n_events >= (1LL << 31) - 1024) return fail(h, ART_ERR_ARG, "more than 2^31 events in one call");
  CU(cudaSetDevice(h->device));
  const int ncols = h->ncols;
  if (!h->hist.p) {
    int rc = art_hist_reset(h);
    if (rc) return rc;
  }
  const long long ne = std::max<long long>(n_events, 1);
  if (int rc = ensure_lanes(h, 1)) return rc;
  Lane& L = h->lanes[0];
  CU(h->events.ensure(ne * sizeof(EventData)));
  CU(h->fin_count.ensure(ne * sizeof(int)));
  CU(h->fin_info.ensure(ne * sizeof(int)));
  CU(h->gctr.ensure(16 * sizeof(unsigned long long)));
  CU(h->stats.ensure(sizeof(DevStats)));
  if (h->params.save_mode >= 3) {
    // saveMode 3 (one tree_* text file per event, MR.jl:574-578): every node of every tree is logged.  A debugging /
    // plotting mode — sized for the worst case (max_nodes + 2 nodes per event) while that stays below 4 GB
    const size_t per = sizeof(NodeLogEntry) + (size_t)h->params.cross_cap * 4 * sizeof(double);
    long long cap = ne * (h->params.max_nodes + 2);
    if ((size_t)cap * per > ((size_t)4 << 30)) cap = std::max<long long>(ne * 16 + 4096, (long long)(((size_t)4 << 30) / per));
    CU(h->nlog.ensure((size_t)cap * sizeof(NodeLogEntry)));
    CU(h->nlog_cross.ensure((size_t)cap * h->params.cross_cap * 4 * sizeof(double)));
    CU(h->nlog_count.ensure(8));
    h->nlog_cap = cap;
  }
  // Rows: the reference appends without bound.  A forward tree emits at most min(num_cutoff, max_nodes + 1) rows
  // (every row counts towards num_cutoff, every node towards max_nodes); the buffer starts at min(that, 3) rows per event
  // and, should a run need more, is re-allocated at the exact size and the call re-run (per-event Philox streams make
  // the re-run identical).
  const long long bound = std::max(1, std::min(h->params.num_cutoff, h->params.max_nodes + 1));
  long long want_rows = n_events * std::min<long long>(bound, 3) + 1024;
  if (const char* t = getenv("ART_ROWS_INIT_FRAC"))  // test hook: start with rows = frac · events, so that the re-run path is taken
    want_rows = std::max<long long>(1, (long long)(atof(t) * (double)n_events));
  h->n_rows_dev = 0;
  h->have_global = false;
  h->last_first_event = first_event;
  h->last_batch_events = std::max<long long>(batch_events, 1);
  h->last_batch_stride = batch_stride;
  h->last_n_events = 0;
  unsigned long long* gctr = h->gctr.as<unsigned long long>();
  DevStats* dstats = h->stats.as<DevStats>();
  EvGuard evg;
  cudaEvent_t e0 = evg.make(), e1 = evg.make();
  if (!e0 || !e1) return fail(h, ART_ERR_CUDA, "cudaEventCreate");
  unsigned long long hc[16];
  float ms_tot = 0;
  for (int attempt = 0;; ++attempt) {
    if (want_rows > h->rows_cap) {
      CU(h->rows.ensure((size_t)want_rows * ncols * sizeof(double)));
      CU(h->keys.ensure((size_t)want_rows * sizeof(long long)));
      h->rows_cap = want_rows;
    }
    h->launches = 0;
    // gctr layout: [0] failed sampler attempts, [1] tot_count, [6] n_rows, [7,8] n_final (photon, axion)
    CU(cudaMemsetAsync(gctr, 0, 16 * sizeof(unsigned long long), h->stream));
    CU(cudaMemsetAsync(h->stats.p, 0, sizeof(DevStats), h->stream));
    RowSink sink{h->rows.as<double>(), h->keys.as<long long>(), gctr + 6, (unsigned long long)h->rows_cap, ncols, gctr + 7};
    CU(cudaEventRecord(e0, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    L.ms_sampler = L.ms_weights = L.ms_tracer = 0;
    L.ms_k[0] = L.ms_k[1] = L.ms_k[2] = 0;
    L.n_blocks = 0;
    if (n_events > 0) {
      const int rc = h->dp.integrator == IG_DP5
                         ? run_pool<IG_DP5>(h, L, seed, first_event, n_events, h->last_batch_events, batch_stride, sink, gctr, dstats, evg)
                         : (h->dp.bl > 0.0
                                ? run_pool<IG_VERN6_BL>(h, L, seed, first_event, n_events, h->last_batch_events, batch_stride, sink, gctr, dstats, evg)
                                : run_pool<IG_VERN6>(h, L, seed, first_event, n_events, h->last_batch_events, batch_stride, sink, gctr, dstats, evg));
      if (rc) return rc;
    }
    CU(cudaMemcpy(hc, gctr, sizeof(hc), cudaMemcpyDeviceToHost));
    if ((long long)hc[6] <= h->rows_cap) break;
    if (attempt >= 1) return fail(h, ART_ERR_CAPACITY, "device row buffer overflow after re-allocation");
    want_rows = (long long)hc[6] + 1024;  // exact: the counter counts every row, stored or not
  }
  const long long nr = (long long)hc[6];
  if (nr > 0) {
    k_patch_rows<<<nblk(nr, 256), 256, 0, h->stream>>>(nr, h->rows.as<double>(), h->keys.as<long long>(), ncols, 0,
                                                       h->events.as<EventData>(), h->fin_count.as<int>(),
                                                       h->fin_info.as<int>(), 0);
    const int nb = 2 * h->n_ct * h->n_ph * h->n_dw * 4;
    const int use_smem = nb * (int)sizeof(double) <= 48 * 1024;
    const int blocks = (int)std::min<long long>(h->n_sm * 4, (nr + 255) / 256);
    k_hist<<<blocks, 256, use_smem ? nb * sizeof(double) : 0, h->stream>>>(0, nr, h->rows.as<double>(), ncols, h->n_ct, h->n_ph, h->n_dw, h->dw_lo,
                                                                           h->n_dw > 1 ? (double)h->n_dw / (h->dw_hi - h->dw_lo) : 0.0,
                                                                           h->hist.as<double>(), use_smem);
    h->launches += 2;
    CU(cudaGetLastError());
  }
  CU(cudaEventRecord(e1, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  cudaEventElapsedTime(&ms_tot, e0, e1);

  DevStats ds;
  CU(cudaMemcpy(&ds, dstats, sizeof(ds), cudaMemcpyDeviceToHost));
  h->n_rows_dev = nr;
  h->last_n_events = n_events;
  const long long f_inx = (long long)hc[0] + (long long)hc[7];
  if (n_rows_out) *n_rows_out = h->n_rows_dev;
  if (f_inx_out) *f_inx_out = f_inx;
  if (tot_count_out) *tot_count_out = (long long)hc[1];
  art_stats S;
  memset(&S, 0, sizeof(S));
  S.n_events = n_events; S.n_attempts = (long long)ds.n_attempts; S.n_sampler_cond = (long long)ds.n_sampler_cond;
  S.n_rays = (long long)ds.n_rays; S.n_rays_photon = (long long)ds.n_rays_photon; S.n_acc = (long long)ds.n_acc;
  S.n_rej = (long long)ds.n_rej; S.n_rhs = (long long)ds.n_rhs; S.n_cond = (long long)ds.n_cond;
  S.n_final_photon = (long long)hc[7]; S.n_final_axion = (long long)hc[8];
  for (int k = 0; k < 5; ++k) S.n_status[k] = (long long)ds.n_status[k];
  // ms_sampler: sum of the sampler launches (they overlap the loop on their own stream); ms_tracer: the loop, start to drain
  S.ms_sampler = L.ms_sampler; S.ms_weights = 0.0; S.ms_tracer = L.ms_tracer;
  S.ms_tree = 0.0;
  S.ms_total = ms_tot;
  S.n_launches = h->launches.load();
  S.n_cond_eval = (long long)ds.n_cond_eval;
  S.n_scanned_steps = (long long)ds.n_scanned;
  S.n_pool_blocks = L.n_blocks;
  S.ms_k_stages = L.ms_k[0]; S.ms_k_scan = L.ms_k[1]; S.ms_k_detect = L.ms_k[2];
  if (st_out) *st_out = S;
  // trailing histogram counters: [f_inx, n_rows, n_events, n_acc, n_rej, n_rhs, n_cond, n_rays]
  h->host_counters[0] += (double)f_inx; h->host_counters[1] += (double)h->n_rows_dev; h->host_counters[2] += (double)n_events;
  h->host_counters[3] += (double)S.n_acc; h->host_counters[4] += (double)S.n_rej; h->host_counters[5] += (double)S.n_rhs;
  h->host_counters[6] += (double)S.n_cond; h->host_counters[7] += (double)S.n_rays;
  return ART_OK;
}

int art_run_device(art_handle* h, uint64_t seed, int64_t first_event, int64_t n_events, int64_t* n_rows_out,
                   int64_t* f_inx_out, int64_t* tot_count_out, art_stats* st_out) {
  if (!h || n_events < 0) return ART_ERR_ARG;
  return art_run_device_batches(h, seed, first_event, n_events, 1, n_events, n_rows_out, f_inx_out, tot_count_out, st_out);
}

int art_fetch_rows(art_handle* h, double* rows, int64_t rows_cap, int64_t* n_rows) {
  if (!h || !rows) return ART_ERR_ARG;
  CU(cudaSetDevice(h->device));
  const long long n = h->n_rows_dev;
  if (n_rows) *n_rows = n;
  if (n > rows_cap) return fail(h, ART_ERR_CAPACITY, "rows_cap too small (n_rows holds the number of rows: enlarge the buffer and call art_fetch_rows again)");
  if (n == 0) return ART_OK;
  const int ncols = h->ncols;
  // order rows as the reference writes them — by event, then by position in the tree — on the device:
  // radix sort of (key, row index) pairs (CUB, a utility outside the hot path), gather, one copy out
  DBuf* s = h->scratch;
  CU(s[0].ensure((size_t)n * 8));            // keys out
  CU(s[1].ensure((size_t)n * 4));            // index in
  CU(s[2].ensure((size_t)n * 4));            // index out
  CU(s[3].ensure((size_t)n * ncols * 8));    // ordered rows
  k_iota<<<nblk(n, 256), 256, 0, h->stream>>>(s[1].as<int>(), n);
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, h->keys.as<long long>(), s[0].as<long long>(), s[1].as<int>(),
                                  s[2].as<int>(), (int)n, 0, 64, h->stream);
  CU(s[4].ensure(tmp_bytes));
  CU(cub::DeviceRadixSort::SortPairs(s[4].p, tmp_bytes, h->keys.as<long long>(), s[0].as<long long>(), s[1].as<int>(),
                                     s[2].as<int>(), (int)n, 0, 64, h->stream));
  k_gather_rows<<<nblk(n * ncols, 256), 256, 0, h->stream>>>(h->rows.as<double>(), s[2].as<int>(), n, ncols, s[3].as<double>());
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(rows, s[3].p, (size_t)n * ncols * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return ART_OK;
}

int art_fetch_events(art_handle* h, double* events, int64_t events_cap, int64_t* n_events) {
  if (!h || !events || !n_events) return ART_ERR_ARG;
  CU(cudaSetDevice(h->device));
  const long long n = h->last_n_events;
  *n_events = n;
  if (n > events_cap) return fail(h, ART_ERR_CAPACITY, "event buffer too small");
  if (n == 0) return ART_OK;
  std::vector<EventData> ev((size_t)n);
  std::vector<int> cnt((size_t)n);
  CU(cudaMemcpy(ev.data(), h->events.p, (size_t)n * sizeof(EventData), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(cnt.data(), h->fin_count.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
  for (long long i = 0; i < n; ++i) {
    const EventData& E = ev[(size_t)i];
    double* o = events + i * ART_EVENT_COLS;
    o[0] = (double)(h->last_first_event + (i / h->last_batch_events) * h->last_batch_stride + i % h->last_batch_events);
    for (int c = 0; c < 3; ++c) { o[1 + c] = E.vinf[c]; o[5 + c] = E.x[c]; o[8 + c] = E.k_init[c]; }
    o[4] = E.sln_prob;
    o[11] = (double)cnt[(size_t)i];
  }
  return ART_OK;
}

int art_host_alloc(void** p, int64_t bytes) {
  if (!p || bytes <= 0) return ART_ERR_ARG;
  *p = nullptr;
  const cudaError_t e = cudaMallocHost(p, (size_t)bytes);
  if (e != cudaSuccess) return fail(nullptr, ART_ERR_CUDA, std::string("cudaMallocHost: ") + cudaGetErrorString(e));
  return ART_OK;
}

int art_host_free(void* p) {
  if (!p) return ART_OK;
  return cudaFreeHost(p) == cudaSuccess ? ART_OK : ART_ERR_CUDA;
}

int art_run(art_handle* h, uint64_t seed, int64_t first_event, int64_t n_events, double* rows, int64_t rows_cap,
            int64_t* n_rows, int64_t* f_inx, int64_t* tot_count, art_stats* stats) {
  int rc = art_run_device(h, seed, first_event, n_events, n_rows, f_inx, tot_count, stats);
  if (rc) return rc;
  return art_fetch_rows(h, rows, rows_cap, n_rows);
}

int art_normalize_rows(double* rows, int64_t n_rows, int32_t ncols, int64_t f_inx_total) {
  if (!rows || ncols < 8 || f_inx_total <= 0) return ART_ERR_ARG;
  const double f = (double)f_inx_total;
  for (int64_t i = 0; i < n_rows; ++i) rows[i * ncols + 7] /= f;  // MainRunner.jl:749 (saveAll[:, 8])
  return ART_OK;
}

// Julia's string(::Float64) (Ryu shortest round-trip; fixed notation iff -4 < point <= 6,
// otherwise d.ddde±x) for the file name, MainRunner.jl:750-760.
static std::string julia_float(double v) {
  if (v == 0.0) return "0.0";
  char buf[64];
  for (int prec = 1; prec <= 17; ++prec) {
    snprintf(buf, sizeof buf, "%.*e", prec - 1, v);
    if (strtod(buf, nullptr) == v) break;
  }
  std::string s(buf);
  const size_t epos = s.find('e');
  std::string mant = s.substr(0, epos);
  const int ex = atoi(s.c_str() + epos + 1);
  bool neg = false;
  if (mant[0] == '-') {
    neg = true;
    mant = mant.substr(1);
  }
  std::string digits;
  for (char c : mant)
    if (c != '.') digits.push_back(c);
  while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
  const int nd = (int)digits.size();
  const int pt = ex + 1;  // position of the decimal point relative to the digit string
  std::string out;
  if (pt > -4 && pt <= 6) {
    if (pt <= 0) out = "0." + std::string((size_t)(-pt), '0') + digits;
    else if (pt >= nd) out = digits + std::string((size_t)(pt - nd), '0') + ".0";
    else out = digits.substr(0, (size_t)pt) + "." + digits.substr((size_t)pt);
  } else {
    out = digits.substr(0, 1) + "." + (nd > 1 ? digits.substr(1) : std::string("0")) + "e" + std::to_string(ex);
  }
  return (neg ? "-" : "") + out;
}

int art_npy_name(const art_params* p, int64_t Ntajs, const char* dir_tag, const char* file_tag, char* out, int64_t out_cap) {
  if (!p || !out) return ART_ERR_ARG;
  std::string s = std::string(dir_tag ? dir_tag : "results") + "/npy/tree_";
  s += "MassAx_" + julia_float(p->Mass_a) + "_AxionG_" + julia_float(p->Ax_g);
  s += "_ThetaM_" + julia_float(p->theta_m) + "_rotPulsar_" + julia_float(p->omega_pul) + "_B0_" + julia_float(p->B0);
  s += "_Ax_trajs_" + std::to_string((long long)Ntajs);
  s += "_N_Times_" + std::to_string(p->ntimes);
  s += "_num_cutoff_" + std::to_string(p->num_cutoff);
  s += "_MC_nodes_" + std::to_string(p->MC_nodes);
  s += "_max_nodes_" + std::to_string(p->max_nodes);
  s += std::string("_") + (file_tag ? file_tag : "") + ".npy";
  if ((int64_t)s.size() + 1 > out_cap) return ART_ERR_CAPACITY;
  memcpy(out, s.c_str(), s.size() + 1);
  return ART_OK;
}

int art_write_event_files(const char* dir_tag, const char* file_tag, const double* rows, int64_t n_rows,
                          const double* events, int64_t n_events) {
  if ((!rows && n_rows > 0) || (!events && n_events > 0)) return ART_ERR_ARG;
  const std::string dir = std::string(dir_tag ? dir_tag : "results") + "/event/", tag = file_tag ? file_tag : "";
  auto jint = [](double v) { return std::to_string((long long)v); };
  FILE* f = fopen((dir + "final_" + tag).c_str(), "w");
  if (!f) return ART_ERR_IO;
  for (int64_t i = 0; i < n_rows; ++i) {  // MR.jl:689-701
    const double* r = rows + i * ART_ROW_COLS_FULL;
    // the root node's time is the integer literal 0 of MR.jl:656 (prints "0"); every other node's is a Float64
    const std::string t = r[30] == 0.0 ? std::string("0") : julia_float(r[30]);
    fprintf(f, "%s %s %s %s %s %s %s %s %s %s\n", jint(r[0]).c_str(), julia_float(r[13]).c_str(), jint(r[1]).c_str(),
            julia_float(r[2]).c_str(), julia_float(r[3]).c_str(), julia_float(r[29]).c_str(), julia_float(r[4]).c_str(),
            julia_float(r[5]).c_str(), julia_float(r[6]).c_str(), t.c_str());
  }
  fclose(f);
  f = fopen((dir + "event_" + tag).c_str(), "w");
  if (!f) return ART_ERR_IO;
  for (int64_t i = 0; i < n_events; ++i) {  // MR.jl:593-610 and 733-735
    const double* e = events + i * ART_EVENT_COLS;
    fprintf(f, "%s %s %s %s %s", jint(e[0]).c_str(), julia_float(e[1]).c_str(), julia_float(e[2]).c_str(),
            julia_float(e[3]).c_str(), julia_float(e[4]).c_str());
    for (int c = 0; c < 3; ++c) fprintf(f, " %s", julia_float(e[5 + c]).c_str());    // nb.x, nb.y, nb.z: the sample point
    for (int c = 0; c < 3; ++c) fprintf(f, " %s", julia_float(-e[8 + c]).c_str());   // nb.kx …: −k_init (back-trace node)
    for (int c = 0; c < 3; ++c) fprintf(f, " %s", julia_float(e[5 + c]).c_str());    // xpos_flat
    for (int c = 0; c < 3; ++c) fprintf(f, " %s", julia_float(e[8 + c]).c_str());    // k_init
    fprintf(f, " 0.0 %s\n", jint(e[11]).c_str());  // wall-clock seconds of the event (not reproduced), count
  }
  fclose(f);
  return ART_OK;
}

}  // extern "C"

template <int IG>
static int node_trajectories(art_handle* h, const NodeLogEntry* d_nodes, long long n, int ntimes, int out_cap, double* d_traj, int* d_npts,
                             double* d_rec, double* d_k) {
  k_node_traj<IG><<<nblk(n, 64), 64, 0, h->stream>>>(h->dp, d_nodes, n, ntimes, d_traj, d_npts, out_cap, d_rec, d_k);
  CU(cudaGetLastError());
  return ART_OK;
}

extern "C" {

int art_write_tree_files(art_handle* h, const char* dir_tag, const char* file_tag, int64_t* n_files) {
  if (!h) return ART_ERR_ARG;
  if (h->params.save_mode < 3) return fail(h, ART_ERR_ARG, "art_write_tree_files needs a handle created with save_mode = 3");
  CU(cudaSetDevice(h->device));
  if (n_files) *n_files = 0;
  unsigned long long cnt = 0;
  if (!h->nlog_count.p || h->last_n_events <= 0) return ART_OK;
  CU(cudaMemcpy(&cnt, h->nlog_count.p, 8, cudaMemcpyDeviceToHost));
  if ((long long)cnt > h->nlog_cap) return fail(h, ART_ERR_CAPACITY, "node log overflow: saveMode 3 is a debugging mode, run fewer events per call");
  const long long n = (long long)cnt;
  if (n == 0) return ART_OK;
  const int cc = h->params.cross_cap;
  const int ntimes = std::max(h->params.ntimes, 2);
  const int out_cap = ntimes + 64 + 2;
  const size_t traj_bytes = (size_t)n * out_cap * 4 * sizeof(double);
  if (traj_bytes > ((size_t)16 << 30)) return fail(h, ART_ERR_CAPACITY, "saveMode 3: trajectories of this run exceed 16 GB (fewer events or a smaller ntimes)");
  // the log in tree order: by event, then by position
  std::vector<NodeLogEntry> nodes((size_t)n);
  std::vector<double> cross((size_t)n * cc * 4);
  CU(cudaMemcpy(nodes.data(), h->nlog.p, (size_t)n * sizeof(NodeLogEntry), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(cross.data(), h->nlog_cross.p, cross.size() * sizeof(double), cudaMemcpyDeviceToHost));
  DBuf* s = h->scratch;
  CU(s[5].ensure(traj_bytes));
  CU(s[6].ensure((size_t)n * sizeof(int)));
  CU(s[7].ensure((size_t)n * 64 * 8 * sizeof(double)));
  CU(s[8].ensure((size_t)n * NODE_SCRATCH * sizeof(double)));
  const int rc = h->dp.integrator == IG_DP5
                     ? node_trajectories<IG_DP5>(h, h->nlog.as<NodeLogEntry>(), n, ntimes, out_cap, s[5].as<double>(), s[6].as<int>(), s[7].as<double>(), s[8].as<double>())
                     : (h->dp.bl > 0.0
                            ? node_trajectories<IG_VERN6_BL>(h, h->nlog.as<NodeLogEntry>(), n, ntimes, out_cap, s[5].as<double>(), s[6].as<int>(), s[7].as<double>(), s[8].as<double>())
                            : node_trajectories<IG_VERN6>(h, h->nlog.as<NodeLogEntry>(), n, ntimes, out_cap, s[5].as<double>(), s[6].as<int>(), s[7].as<double>(), s[8].as<double>()));
  if (rc) return rc;
  std::vector<double> traj((size_t)n * out_cap * 4);
  std::vector<int> npts((size_t)n);
  CU(cudaMemcpyAsync(traj.data(), s[5].p, traj_bytes, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(npts.data(), s[6].p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  std::vector<long long> order((size_t)n);
  for (long long i = 0; i < n; ++i) order[(size_t)i] = i;
  std::sort(order.begin(), order.end(), [&](long long a, long long b) {
    const NodeLogEntry &A = nodes[(size_t)a], &B = nodes[(size_t)b];
    return A.event != B.event ? A.event < B.event : A.pos < B.pos;
  });
  const std::string dir = std::string(dir_tag ? dir_tag : "results") + "/tree/tree_" + (file_tag ? file_tag : "");
  const EventIds ids{h->last_first_event, h->last_batch_events, h->last_batch_stride};
  FILE* f = nullptr;
  int cur_event = -1;
  long long files = 0;
  for (long long oi = 0; oi < n; ++oi) {
    const long long i = order[(size_t)oi];
    const NodeLogEntry& nd = nodes[(size_t)i];
    if (nd.event != cur_event) {  // MR.jl:575: one file per event, named by the event number
      if (f) fclose(f);
      f = fopen((dir + std::to_string((long long)ids.id(nd.event))).c_str(), "w");
      if (!f) return fail(h, ART_ERR_IO, "cannot open " + dir + "… (the directory <dir_tag>/tree must exist, as in the reference)");
      cur_event = nd.event;
      files++;
    }
    // saveNode, MR.jl:17-65
    fprintf(f, "%s %s %s %s\n", nd.species == SPECIES_PHOTON ? "photon" : "axion", julia_float(nd.weight).c_str(),
            julia_float(nd.prob).c_str(), julia_float(nd.parent_weight).c_str());
    if (nd.n_cross > 0) {
      const double* q = cross.data() + (size_t)i * cc * 4;
      for (int c = 0; c < 4; ++c) {  // xc, yc, zc, tc
        for (int j = 0; j < nd.n_cross; ++j) fprintf(f, "  %s", julia_float(q[4 * j + c]).c_str());
        if (c < 3) fputc('\n', f);
      }
    } else {
      fputs("-\n-\n-", f);
    }
    fputc('\n', f);
    const double* t = traj.data() + (size_t)i * out_cap * 4;
    for (int c = 0; c < 4; ++c) {  // traj[:, 1], traj[:, 2], traj[:, 3], times (ln t)
      for (int j = 0; j < npts[(size_t)i]; ++j) fprintf(f, "  %s", julia_float(t[4 * j + c]).c_str());
      fputc('\n', f);
    }
  }
  if (f) fclose(f);
  if (n_files) *n_files = files;
  return ART_OK;
}

int art_write_npy(const char* path, const double* rows, int64_t n_rows, int32_t ncols) {
  if (!path || (!rows && n_rows > 0)) return ART_ERR_ARG;
  FILE* f = fopen(path, "wb");
  if (!f) return ART_ERR_IO;
  char dict[256];
  int len = snprintf(dict, sizeof dict, "{'descr': '<f8', 'fortran_order': True, 'shape': (%lld, %d), }", (long long)n_rows, ncols);
  int total = 10 + len + 1;
  int pad = (64 - total % 64) % 64;
  unsigned short hlen = (unsigned short)(len + pad + 1);
  fwrite("\x93NUMPY\x01\x00", 1, 8, f);
  fwrite(&hlen, 2, 1, f);
  fwrite(dict, 1, (size_t)len, f);
  for (int i = 0; i < pad; ++i) fputc(' ', f);
  fputc('\n', f);
  // NPZ.jl writes the Julia (column-major) array as is: column by column
  std::vector<double> col((size_t)n_rows);
  for (int c = 0; c < ncols; ++c) {
    for (int64_t i = 0; i < n_rows; ++i) col[(size_t)i] = rows[i * ncols + c];
    fwrite(col.data(), sizeof(double), (size_t)n_rows, f);
  }
  fclose(f);
  return ART_OK;
}

// ---- multi-GPU -------------------------------------------------------------------------------

int art_comm_unique_id(char* id128) {
  std::string e;
  if (!id128) return ART_ERR_ARG;
  if (!nccl_load(e)) return fail(nullptr, ART_ERR_NCCL, e);
  int rc = g_nccl.GetUniqueId(id128);
  return rc == 0 ? ART_OK : fail(nullptr, ART_ERR_NCCL, "ncclGetUniqueId failed");
}

int art_comm_init(art_handle* h, int32_t nranks, int32_t rank, const char* id128) {
  if (!h || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return ART_ERR_ARG;
  std::string e;
  if (!nccl_load(e)) return fail(h, ART_ERR_NCCL, e);
  CU(cudaSetDevice(h->device));
  Id128 id;
  memcpy(id.b, id128, 128);
  int rc = g_nccl.CommInitRank(&h->comm, nranks, id, rank);
  if (rc != 0) return fail(h, ART_ERR_NCCL, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
  h->nranks = nranks;
  h->rank = rank;
  return ART_OK;
}

int art_comm_allreduce_hist(art_handle* h) {
  if (!h) return ART_ERR_ARG;
  CU(cudaSetDevice(h->device));
  if (!h->hist.p) {
    int rc = art_hist_reset(h);
    if (rc) return rc;
  }
  int rc = hist_push_counters(h);
  if (rc) return rc;
  const size_t bytes = (size_t)art_hist_len(h) * sizeof(double);
  CU(h->hist_global.ensure(bytes));
  if (h->comm && h->nranks > 1) {
    // ncclAllReduce(sum, f64) over [bins | counters], OUT OF PLACE — the only collective of the path.  The local
    // block keeps accumulating untouched, so the call can be repeated (after every pipeline call, or once at the
    // end): each time the global block is the sum over ranks of the local blocks as they stand.
    rc = g_nccl.AllReduce(h->hist.p, h->hist_global.p, (size_t)art_hist_len(h), /*ncclFloat64*/ 8, /*ncclSum*/ 0, h->comm, h->stream);
    if (rc != 0) return fail(h, ART_ERR_NCCL, std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
  } else {
    CU(cudaMemcpyAsync(h->hist_global.p, h->hist.p, bytes, cudaMemcpyDeviceToDevice, h->stream));
  }
  CU(cudaStreamSynchronize(h->stream));
  h->have_global = true;
  return ART_OK;
}

int art_comm_destroy(art_handle* h) {
  if (!h) return ART_ERR_ARG;
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  h->comm = nullptr;
  h->nranks = 1;
  h->rank = 0;
  return ART_OK;
}

int art_fp64_peak(art_handle* h, double* tflops, double* ms_out) {
  if (!h || !tflops) return ART_ERR_ARG;
  CU(cudaSetDevice(h->device));
  const int blocks = h->n_sm * 8, threads = 256, iters = 4096;
  CU(h->scratch[0].ensure((size_t)blocks * threads * sizeof(double)));
  cudaEvent_t a, b;
  CU(cudaEventCreate(&a));
  CU(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CU(cudaEventRecord(a, h->stream));
    k_fp64_peak<<<blocks, threads, 0, h->stream>>>(h->scratch[0].as<double>(), iters, 0.999999, 1e-9);
    CU(cudaEventRecord(b, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (rep > 0) best = std::min(best, ms);
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  const double flops = 2.0 * 64.0 * iters * (double)blocks * threads;  // 8×8 FMAs per iteration
  *tflops = flops / (best * 1e-3) / 1e12;
  if (ms_out) *ms_out = best;
  return ART_OK;
}

}  // extern "C"
