/*
 * art.h — C ABI of the B200-native hot path of SamWitte/Adiabatic_RayTracer.
 *
 * The reference has no FFI layer: its seam is the Julia function boundary
 *   main_runner_tree(Mass_a, Ax_g, θm, ωPul, B0, rNS, Mass_NS, ωProp, Ntajs, gammaF; kwargs...)
 * (reference src/MainRunner.jl:355-363, called from src/Gen_Samples.jl:181-191).
 * A Julia host keeps that function and `ccall`s the entry points below
 * (see INTEGRATION.md and julia/MainRunnerGPU.jl).  Every function is plain C:
 * POD parameters, caller-owned host buffers, `int` status (0 = OK, <0 = error,
 * message via art_last_error).  No exceptions cross the boundary.  There is no
 * CPU fallback: every compute entry point requires a CUDA device.
 *
 * Conventions
 *  - units as in the reference: eV, 1/GeV, rad, 1/s, G, km, M_sun.
 *  - 3-vectors for n rays are component-major: v[c*n + i], c = 0..2
 *    (== a Julia `Matrix{Float64}(undef, n, 3)`; numpy shape (3, n) C-order).
 *  - a handle owns the device memory of ONE GPU and is not thread-safe;
 *    distinct handles are independent.
 */
#ifndef ART_H_
#define ART_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes --------------------------------------------------------------------- */
#define ART_OK 0
#define ART_ERR_ARG (-1)        /* bad argument / unsupported switch                        */
#define ART_ERR_CUDA (-2)       /* CUDA runtime error (no device, launch failure, OOM ...)   */
#define ART_ERR_NO_SURFACE (-3) /* maxR < rNS: "Too small Max R" (MainRunner.jl:389-396)     */
#define ART_ERR_CAPACITY (-4)   /* caller buffer too small                                  */
#define ART_ERR_NCCL (-5)       /* NCCL unavailable / failed                                */
#define ART_ERR_IO (-6)

/* ---- per-ray termination status (replaces the reference's prints) ----------- */
#define ART_RAY_END 0       /* reached ln t = log(1/ωPul): the "observer sphere"            */
#define ART_RAY_HIT_NS 1    /* photon terminated at r < 1.01 rNS  (RayTracer.jl:352-359)    */
#define ART_RAY_CROSSING 2  /* stopped at max_crossings level crossings (RT.jl:345-349)     */
#define ART_RAY_MAXITERS 3  /* maxiters exhausted (RT.jl:383-384)                           */
#define ART_RAY_CROSS_OVERFLOW 4 /* more crossings than the record capacity                 */

#define ART_INTEGRATOR_VERN6 0
#define ART_INTEGRATOR_DP5 1

#define ART_SPECIES_AXION 0
#define ART_SPECIES_PHOTON 1

/* number of doubles in one crossing record: x y z kx ky kz t Δω  (RayTracer.jl:325-342)   */
#define ART_CROSS_REC 8
/* row widths written by main_runner_tree (MainRunner.jl:715,720)                            */
#define ART_ROW_COLS_SHORT 13
#define ART_ROW_COLS_LONG 29
/* save_mode >= 2: the 29 columns + |k_f| and the time of the node, the two numbers of the final_* text file
 * (MainRunner.jl:689-701) that the .npy row does not hold; the .npy file still gets columns 0..28        */
#define ART_ROW_COLS_FULL 31
/* per-event numbers of the event_* text file (MainRunner.jl:593-610, 733-735):
 * event id, v_inf[3] (units of c), sln_prob (not normalised), sample point x[3], k_init[3], count       */
#define ART_EVENT_COLS 12

/*
 * Every physics and numerics knob of main_runner_tree's signature and of the fixed block
 * Gen_Samples.jl:158-174, carried explicitly (the reference leaks some through globals,
 * e.g. Mass_NS in get_Prob_nonAD, MainRunner.jl:75).
 */
typedef struct art_params {
  /* positional arguments of main_runner_tree (MainRunner.jl:355) */
  double Mass_a;    /* axion mass [eV]                     --MassA  default 1e-5           */
  double Ax_g;      /* axion-photon coupling [1/GeV]       --Axg    default 1e-12          */
  double theta_m;   /* magnetic inclination [rad]          --ThetaM default 0.0            */
  double omega_pul; /* rotation angular frequency [1/s]    --rotW   default 1.0            */
  double B0;        /* polar field [G]                     --B0     default 1e14           */
  double rNS;       /* stellar radius [km]                 --rNS    default 10             */
  double Mass_NS;   /* stellar mass [M_sun]                --Mass_NS default 1             */
  /* keyword arguments */
  double bndry_lyr;   /* --bndry_lyr (Gen_Samples.jl:127-130): > 0 adds the boundary-layer term to ωp (RT.jl:1155-1162);
                         needs melrose = 1 and integrator = Vern6; <= 0 (default −1): off      */
  double ode_err;     /* abstol of the ray integrator      Gen_Samples.jl:163 = 1e-6        */
  double reltol;      /* reltol of the ray integrator      RayTracer.jl:383   = 1e-7        */
  double fix_time;    /* field time used for maxR          Gen_Samples.jl:162 = 0           */
  double rho_DM;      /* GeV/cm^3                          Gen_Samples.jl:172 = 0.45        */
  double vmean_ax;    /* km/s (only the literal 220 is used by the reference, RT.jl:1531)   */
  double prob_cutoff; /* --probCutoff default 1e-10                                        */
  double ln_t_start;  /* MainRunner.jl:411  = -30                                           */
  double t0_sample;   /* field time seen by sampler / conversion probability / weights;
                         the reference hard-wires 0 (RT.jl:1480, MR.jl:74); phase-resolved
                         runs (BASELINE config 4) set 2π j/(64 ωPul)                       */
  double metric_interior_radius; /* g_schwartz kwarg rNS=10.0 that no caller overrides
                                    (RT.jl:455) — quirk kept; must be 10.0
                                    (any other value is refused with ART_ERR_ARG)          */
  double sampler_dt;  /* Euler march step of the surface sampler [km], RT.jl:1613 = 0.5     */
  double dtmin;       /* RT.jl:384 = 1e-13 (forced)                                         */
  int32_t flat;          /* Gen_Samples.jl:165 = false ("mixed" flat mode)       */
  int32_t isotropic;     /* Gen_Samples.jl:166 = false                                      */
  int32_t melrose;       /* Gen_Samples.jl:167 = true; false = hamiltonian's "original form" (RT.jl:546-548) */
  int32_t thick_surface; /* Gen_Samples.jl:173 = true; false = sampler condition log ωp − log m_a (RT.jl:1548) */
  int32_t n_maxSample;   /* Gen_Samples.jl:174 = 6                                          */
  int32_t num_cutoff;    /* --numCutoff default 5                                           */
  int32_t MC_nodes;      /* --MCNodes   default 5                                           */
  int32_t max_nodes;     /* --maxNodes  default 50                                          */
  int32_t ntimes;        /* Gen_Samples.jl:164 = 3: file name; points per saved trajectory (save_mode 3) */
  int32_t save_mode;     /* 0 → 13-column rows, 1 → 29-column rows (MR.jl:714-721), 2 → 31 columns + event table
                            (clear-text final_* / event_* files, MR.jl:435-444, 593-610, 689-701); 3 → as 2, and the tree
                            kernel logs every node for art_write_tree_files (one tree_* file per event with the
                            trajectories sampled at `ntimes` points, MR.jl:17-65)                          */
  int32_t interp_points_ray;     /* ContinuousCallback interp_points, RT.jl:357 = 50; 2..64  */
  int32_t interp_points_sampler; /* RT.jl:1599 = 20; 2..32                                  */
  int32_t maxiters;      /* RT.jl:384 = 1e5                                                 */
  int32_t cross_cap;     /* capacity of the per-ray crossing record list in the fused
                            pipeline (the reference's list is unbounded; overflow is
                            reported as ART_RAY_CROSS_OVERFLOW)                            */
  int32_t integrator;    /* ART_INTEGRATOR_VERN6 (0, default): Verner's 9-stage 6(5) pair, the
                            reference's solve(prob, Vern6(), …) (RT.jl:383); ART_INTEGRATOR_DP5 (1):
                            Dormand–Prince 5(4), round 1's scheme, kept for comparison              */
  int32_t reserved_;     /* keeps sizeof(art_params) a multiple of 8; must be 0                    */
} art_params;

/* counters accumulated by the tracer / sampler kernels; they feed the roofline (SURVEY §8 d) */
typedef struct art_stats {
  int64_t n_events;       /* events completed                                              */
  int64_t n_attempts;     /* sampler attempts                                              */
  int64_t n_sampler_cond; /* thick-surface condition evaluations in the sampler            */
  int64_t n_rays;         /* node propagations (backward axion + forward tree nodes)       */
  int64_t n_rays_photon;  /* … of which photons                                            */
  int64_t n_acc;          /* accepted RK steps                                             */
  int64_t n_rej;          /* rejected RK steps                                             */
  int64_t n_rhs;          /* right-hand-side evaluations                                   */
  int64_t n_cond;         /* crossing-condition evaluations in the tracer                  */
  int64_t n_final_photon; /* rows with id = 1 (photons that reached the observer sphere)   */
  int64_t n_final_axion;  /* rows with id = 0                                              */
  int64_t n_status[5];    /* ray termination statuses, index = ART_RAY_*                   */
  double ms_sampler;      /* device time (CUDA events), ms: sum of the sampler launches — they run on
                             their own stream, concurrently with the tracer loop            */
  double ms_weights;      /* 0 (weights are computed where events are injected into the loop) */
  double ms_tracer;       /* the wavefront loop from its first launch to the last drain      */
  double ms_tree;         /* 0 (the tree is walked inside the loop)                          */
  double ms_total;        /* the whole pipeline call                                         */
  int64_t n_launches;     /* kernels launched by the pipeline call                          */
  int64_t n_cond_eval;    /* … of n_cond, evaluations actually performed (the rest were decided by the
                             exact far-field shortcut r > Rcut)                             */
  int64_t n_scanned_steps; /* accepted steps that went through the crossing scan (= records the compaction pass
                              k_wf_detect read; the others carried a certified sign)         */
  int64_t n_pool_blocks;  /* blocks of events injected into the pool by the call             */
  double ms_k_stages;     /* per-kernel device time, ms — only measured when the environment variable       */
  double ms_k_scan;       /* ART_TIME_KERNELS is set (CUDA events around every launch of the main stream,   */
  double ms_k_detect;     /* which costs a few per cent): stage kernel, crossing scan, compaction pass      */
} art_stats;

typedef struct art_handle art_handle;

/* Fill *p with the reference defaults (Gen_Samples.jl:18-131 flags, :158-174 fixed block,
 * MainRunner.jl:356-363 kwargs).                                                          */
int art_params_default(art_params* p);

/* Find_Conversion_Surface (RayTracer.jl:1250-1263): maxR = 1.01 rNS (ωp(rNS,θ*)/m_a)^(2/3).
 * Host scalar, once per run; ART_ERR_NO_SURFACE if maxR < rNS (MainRunner.jl:389-396).     */
int art_max_radius(const art_params* p, double* maxR);

/* Create a handle on CUDA device `device_id`.  Fails (ART_ERR_CUDA) without a device.       */
int art_create(const art_params* p, int device_id, art_handle** out);
void art_destroy(art_handle* h);
const char* art_last_error(const art_handle* h); /* h may be NULL: last create() error       */
const char* art_version(void);

/* ---- fine-grained entry points (one per reference function, for parity tests) ---------- */

/* find_samples_new (RayTracer.jl:1480-1653) driven by caller-supplied uniforms, mirroring
 * its `pre_randomized` hook (RT.jl:1482-1518) but with the default √u radius law.
 * u: n×10 row-major uniforms in [0,1): u1..u6 (θi, ϕi, θloc, ϕloc, ϕRND, r²RND),
 * 3 for vIfty, 1 for randInx.  Outputs (component-major 3-vectors):
 * accepted[i] ∈ {0,1} (numV), n_cross[i] (number of kept crossings), x (km),
 * v_loc (units c), v_inf (units c).                                                         */
int art_sample_from_uniforms(art_handle* h, int64_t n, const double* u, double* x, double* v_loc,
                             double* v_inf, int32_t* accepted, int32_t* n_cross);

/* Same with the library's Philox4x32-10 streams: event e ∈ [first_event, first_event+n)
 * draws attempts j = 0,1,… (counter = (block, j, e), key = seed) until one is accepted
 * (MainRunner.jl:463-494).  n_failed[i] = rejected attempts of event i (→ f_inx).           */
int art_sample(art_handle* h, uint64_t seed, int64_t first_event, int64_t n, double* x,
               double* v_loc, double* v_inf, int32_t* n_cross, int32_t* n_failed);

/* Per-sample scalars of MainRunner.jl:498-557: k_norm_Cart (RT.jl:643), k_sphere (RT.jl:983),
 * dwp_ds → cos_w (RT.jl:1370-1375), g_det (RT.jl:734), sln_prob (MR.jl:552-557, before the
 * 1/f_inx normalisation), erg_inf_ini (MR.jl:524), vel_eng (MR.jl:522).                     */
int art_weights(art_handle* h, int64_t n, const double* x, const double* v_loc,
                const double* v_inf, double* k_init, double* cos_w, double* jac_gr,
                double* sln_prob, double* erg_inf, double* vel_eng);

/* get_Prob_nonAD (MainRunner.jl:67-124) + conversion_prob (RayTracer.jl:1405-1473).
 * erg_scaled = erg_inf_ini·|Δω|; b_sign = +1 (forward) or −1 (back-trace passes −B0).
 * Output P_nonAD (the caller forms 1 − exp(−P), MR.jl:136,267).                             */
int art_conv_prob(art_handle* h, int64_t n, const double* x, const double* k,
                  const double* erg_scaled, int32_t b_sign, double* P_nonAD);

/* propagate (RayTracer.jl:171-452) for n independent nodes.  species[i] ∈ ART_SPECIES_*;
 * x0,k0 Cartesian start position [km] / momentum [eV]; erg = erg_inf_ini per ray;
 * t_start = node.t [s]; dOmega = node.Δω; b_sign as above; max_crossings as passed by
 * get_tree (−1 → stop at first, 100000 → collect all).  Outputs: x_f, k_f (RT.jl:393-416
 * back-transform of the last state), u7_f, status (ART_RAY_*), n_cross, and up to
 * cross_cap records of ART_CROSS_REC doubles each, cross_rec[(i*cross_cap + j)*8 + c].
 * counters (optional, may be NULL): n_acc, n_rej, n_rhs, n_cond per ray as int32[4*n].      */
int art_propagate(art_handle* h, int64_t n, const int32_t* species, const double* x0,
                  const double* k0, const double* erg, const double* t_start,
                  const double* dOmega, int32_t b_sign, int32_t max_crossings, double* x_f,
                  double* k_f, double* u7_f, int32_t* status, int32_t* n_cross,
                  double* cross_rec, int32_t cross_cap, int32_t* counters);

/* One evaluation of the ray right-hand side func!/func_axion! (RayTracer.jl:71-123) per ray:
 * u,du are n×7 row-major.  For parity tests of the closed-form gradients.                    */
int art_rhs(art_handle* h, int64_t n, const int32_t* species, const double* u,
            const double* erg, const double* lnt, double* du);

/* ---- fused pipeline: the body of main_runner_tree's event loop (MR.jl:450-747) --------- */

/* Runs events e ∈ [first_event, first_event + n_events): sampling, weights, back-trace,
 * forward tree, row assembly.  rows: row-major n_rows × ncols (ncols = 13, 29 or 31 by
 * save_mode: ART_ROW_COLS_SHORT / _LONG / _FULL) into caller memory of rows_cap rows; column 0 is the event id e (the
 * reference numbers events photon_trajs = 1,2,…), column 7 (sln_prob) is NOT yet divided by
 * f_inx — that normalisation needs the global count (MR.jl:749): use art_normalize_rows.
 * f_inx = failed sampler attempts + final photon rows of these events (MR.jl:469,477,713).
 * tot_count = Σ length(tree) (the reference's return value, MR.jl:667,763).
 * The per-GPU flux histogram (see below) is accumulated on the device as a side effect.     */
int art_run(art_handle* h, uint64_t seed, int64_t first_event, int64_t n_events, double* rows,
            int64_t rows_cap, int64_t* n_rows, int64_t* f_inx, int64_t* tot_count,
            art_stats* stats);

/* Device-resident variant used by bench.py's kernel-only timing: identical work, rows stay in
 * HBM (fetch later with art_fetch_rows).                                                     */
int art_run_device(art_handle* h, uint64_t seed, int64_t first_event, int64_t n_events,
                   int64_t* n_rows, int64_t* f_inx, int64_t* tot_count, art_stats* stats);
/* Several batches of events as ONE pipelined job: batch b = events first_event + b·batch_stride + [0, batch_events),
 * b = 0 … n_batches−1 (batch_stride ≥ batch_events; art_run_device is the case n_batches = 1).  The engine keeps its
 * pool of ray slots full across batch boundaries — slots freed by the late tree nodes of one batch are refilled with
 * events of the next — so only the last batch drains alone.  Results are those of n_batches separate calls (per-event
 * random streams); rows come out in batch order.  This is what config 3–5 style production runs and bench.py's
 * back-to-back steps call.                                                                                     */
int art_run_device_batches(art_handle* h, uint64_t seed, int64_t first_event, int64_t batch_events, int64_t n_batches,
                           int64_t batch_stride, int64_t* n_rows, int64_t* f_inx, int64_t* tot_count, art_stats* stats);
/* Rows of the last pipeline call, ordered as the reference writes them (by event, then by position in the tree).  If
 * rows_cap is too small: ART_ERR_CAPACITY with *n_rows = rows needed — enlarge the buffer and call again (nothing is
 * recomputed).  On the device the row buffer grows by itself (the call is re-run once at the exact size).        */
int art_fetch_rows(art_handle* h, double* rows, int64_t rows_cap, int64_t* n_rows);

/* Event table of the last art_run / art_run_device call: n_events × ART_EVENT_COLS, row-major, in event order
 * (what main_runner_tree writes per event into <dir_tag>/event/event_<file_tag> when saveMode > 1).          */
int art_fetch_events(art_handle* h, double* events, int64_t events_cap, int64_t* n_events);

/* saveMode > 1 clear-text files (MainRunner.jl:435-444): <dir_tag>/event/final_<file_tag> — one line per final
 * particle: event, weight, species id, θf ϕf |k_f|, θfX ϕfX |x_f|, time of the node (MR.jl:689-701) — from
 * ART_ROW_COLS_FULL-column rows in event order, and <dir_tag>/event/event_<file_tag> — one line per event
 * (MR.jl:593-610, 733-735) from the event table.  Numbers are printed like Julia's string(); the per-event
 * wall-clock field of the reference is written as 0.0.  The directory must exist (as in the reference).     */
int art_write_event_files(const char* dir_tag, const char* file_tag, const double* rows, int64_t n_rows,
                          const double* events, int64_t n_events);

/* saveMode 3 (MainRunner.jl:17-65, 574-578, 612, 669-670): one text file <dir_tag>/tree/tree_<file_tag><event> per event of the
 * last pipeline call, holding every node of the event's back-trace and forward tree in saveNode's layout — species, weight,
 * prob, parent_weight; the level crossings xc / yc / zc / tc (or three lines "-"); the trajectory x / y / z and its times (ln t)
 * at the `ntimes` saveat points of `solve(…, saveat = …)` (RT.jl:174, 383), read off the dense output, plus the located
 * events.  Needs a handle created with save_mode = 3 (the tree kernel then logs its nodes); the directory must exist.
 * A debugging / plotting mode (plot/plotTree.py).  Differences from the reference's files: an event point is saved once
 * (DiffEq saves it before and after the callback), and start-point events the start-box rule discards are not saved.  */
int art_write_tree_files(art_handle* h, const char* dir_tag, const char* file_tag, int64_t* n_files);

/* Page-locked host memory for the row buffers (optional: any host pointer works, a pinned one makes the
 * device→host copy of the rows a single DMA transfer at PCIe speed instead of a staged pageable copy).
 * A Julia host wraps the pointer with unsafe_wrap(Array, Ptr{Float64}(p), (ncols, cap)).              */
int art_host_alloc(void** p, int64_t bytes);
int art_host_free(void* p);

/* rows[:,7] /= f_inx_total  (MainRunner.jl:749)                                              */
int art_normalize_rows(double* rows, int64_t n_rows, int32_t ncols, int64_t f_inx_total);

/* Write rows as the reference's .npy (NPZ.jl layout: Float64, fortran_order True).
 * art_npy_name builds "<dir_tag>/npy/tree_MassAx_…_<file_tag>.npy" (MainRunner.jl:750-760)
 * with Julia's string(Float64) formatting.                                                   */
int art_write_npy(const char* path, const double* rows, int64_t n_rows, int32_t ncols);
int art_npy_name(const art_params* p, int64_t Ntajs, const char* dir_tag, const char* file_tag,
                 char* out, int64_t out_cap);

/* ---- binned flux (plot/flux.py:38-52, jonas_test_analyses/analysis.py:86-100) ----------
 * Histogram over (species, cosθf bin, ϕf bin) of the final momentum direction; per bin
 * 4 doubles: Σ w·sln_prob, Σ (w·sln_prob)², Σ w, count.  Trailing block of 8 counters:
 * [f_inx, n_rows, n_events, n_acc, n_rej, n_rhs, n_cond, n_rays].  sln_prob is the
 * un-normalised value; divide the first two moments by f_inx (f_inx²) after the reduce.      */
int art_hist_config(art_handle* h, int32_t n_costheta, int32_t n_phi);
/* Binned spectra (BASELINE config 3): a third axis of n_domega equal bins of Δω (row column 12, MainRunner.jl:709)
 * on [domega_lo, domega_hi); values outside fall in the end bins.  Layout [species(2)][n_costheta][n_phi][n_domega][4].
 * art_hist_config(h, a, b) is art_hist_config_spectrum(h, a, b, 1, 0, 0).                                   */
int art_hist_config_spectrum(art_handle* h, int32_t n_costheta, int32_t n_phi, int32_t n_domega, double domega_lo,
                             double domega_hi);
int64_t art_hist_len(const art_handle* h); /* doubles */
int art_hist_reset(art_handle* h);
int art_hist_get(art_handle* h, double* out, int64_t cap);        /* THIS GPU's block (local bins + local counters)   */
/* the all-reduced block of the last art_comm_allreduce_hist: Σ over ranks of what art_hist_get returns on each;
 * ART_ERR_ARG if no all-reduce has run since the last pipeline call or reset                                  */
int art_hist_get_global(art_handle* h, double* out, int64_t cap);

/* ---- multi-GPU: one process (or thread) per GPU; only collective = all-reduce of hist --- */
int art_comm_unique_id(char* id128);                       /* ncclGetUniqueId, 128 bytes      */
int art_comm_init(art_handle* h, int32_t nranks, int32_t rank, const char* id128);
/* ncclAllReduce(sum, f64) of [bins | counters], out of place: the local block is left untouched (so the call may be
 * repeated, e.g. after every pipeline call), the result is read with art_hist_get_global.  Without a communicator
 * (single GPU) the global block is a copy of the local one.                                                    */
int art_comm_allreduce_hist(art_handle* h);
int art_comm_destroy(art_handle* h);

/* FP64 DFMA peak probe used as the roofline denominator (MEASURED_PEAKS.json has no FP64
 * entry): runs a register-resident DFMA chain on all SMs, returns TFLOP/s.                    */
int art_fp64_peak(art_handle* h, double* tflops, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* ART_H_ */
