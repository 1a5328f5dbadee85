#!/usr/bin/env python
"""bench.py — rays traced to the observer sphere per second (GR mode), BASELINE.json's metric.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...   # CPU arm (oracle port, all host cores)

One "step" = one pass of the fused hot path (surface sampling → weights → back-trace → forward
tree → rows → binned flux [→ NCCL all-reduce of the histogram]) over a batch of synthetic events:
BASELINE.json configs[1], "same NS, Schwarzschild/GR ray equations, 1e6 rays on 1×B200" (each
rank runs its own 1e6-event batch: weak scaling).  A ray is counted when a photon reaches the end
of the integration span t = 1/ωPul, i.e. r ≈ c/ωPul — the reference's "observer sphere"
 — so value = Σ_ranks (final photon rows) / time.

Timing: CUDA events recorded on the library's launching stream around the whole pipeline
(art_stats.ms_total), max over ranks, bracketed by barrier + synchronize; the working set of a
step (≈ 8 KB/event ≈ 8 GB) is far larger than the 126 MB L2, so no explicit flush is needed.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# algorithmic flop per evaluation (mathematical operations after CSE)
F_RHS = 250.0
F_COND = 150.0
F_SAMPLER_COND = 150.0

NS = dict(Mass_a=1e-5, Ax_g=1e-12, theta_m=0.2, omega_pul=1.0, B0=1e14, rNS=10.0, Mass_NS=1.0, flat=0, integrator=0)
SEED = 1769

# the same string in both arms (the reference arm runs a bounded sample of it)
WORKLOAD = ("BASELINE configs[1]: GR (flat=false), m_a=1e-5 eV, Ax_g=1e-12, B0=1e14 G, rotW=1.0 (the reference's default; "
            "P = 1 s is value_p1s), theta_m=0.2, M=1 Msun, R=10 km; 1e6 events per GPU per step (sampler + back-trace "
            "+ forward tree)")



# hardware view of the kernels (ncu --set full / launch list of THIS build, written by tools/ncu_summary.py): loaded, never
# hard-coded; absent file ⇒ the keys are null
NCU_SUMMARY = os.path.join(ROOT, "profiles", "r02_ncu_summary.json")
# algorithmic bytes the compaction pass (k_wf_detect) moves per scanned step: hot record 192 B + sign masks 24 B + list id
# 4 B in, start-box reference point (one 32-byte sector) in, 16 B of the record (one sector) + survivor-list id 4 B out
DETECT_BYTES_PER_STEP = 192 + 24 + 4 + 32 + 32 + 4
HBM_FALLBACK_GBS = 6650.0  # B200_PROFILING.md, only if MEASURED_PEAKS.json is absent


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--events", type=int, default=1_000_000, help="events per rank per step")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--integrator", default="vern6", choices=["vern6", "dp5"],
                    help="vern6 = the reference's scheme (default); dp5 = round 1's scheme, for comparison")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(self.index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def use_all_host_threads():
    """The CPU arms run on every host core this process may use: torchrun exports OMP_NUM_THREADS=1 to its workers, which
    would time the OpenMP oracle on ONE core.  Must run before the oracle library (libgomp) is first loaded."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = os.environ.get("ART_CPU_THREADS", str(n))


def cpu_baseline(seconds, events_hint=None):
    """Oracle (CPU port of the reference path, OpenMP over events) on a bounded sample of the SAME
    workload; rays-to-observer per second on the host cores of this box."""
    use_all_host_threads()
    from adiabatic_raytracer_b200.params import ArtParams
    from oracle import loader as oracle
    p = ArtParams.default(**NS)
    cores = oracle.num_threads()
    t0 = time.perf_counter()
    oracle.run(p, SEED, 1, 64 * cores)
    dt = time.perf_counter() - t0
    rate = 64 * cores / dt
    n = events_hint or max(64 * cores, int(rate * seconds))
    t0 = time.perf_counter()
    out = oracle.run(p, SEED, 1, n)
    dt = time.perf_counter() - t0
    st = out["stats"]
    return {"value": st["n_final_photon"] / dt, "unit": "rays/s", "cores": cores, "kind": "port",
            "sample": f"{n} events of the same workload (GR, seed {SEED}, Vern6), oracle/ (C++ -O2 -fopenmp port of the "
                      f"reference's algorithm incl. dual-number gradients), {dt:.1f} s.  A FASTER-than-reference baseline: the "
                      f"Julia code allocates per right-hand side and runs GC.gc() per event (MainRunner.jl:745); it cannot run "
                      f"here (no Julia in the image)",
            "events_per_s": n / dt}, out, dt, n


def run_reference(args):
    """--impl reference: the reference's own CPU implementation cannot run (Julia absent), so the
    oracle port stands in (tier rule ④), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    use_all_host_threads()
    from adiabatic_raytracer_b200.params import ArtParams
    from oracle import loader as oracle
    p = ArtParams.default(**NS)
    cores = oracle.num_threads()
    n = max(256, 96 * cores)
    for w in range(args.warmup):
        oracle.run(p, SEED, 1 + w * n, max(64, n // 8))
    rays = 0
    t0 = time.perf_counter()
    for s in range(args.steps):
        out = oracle.run(p, SEED, 1 + s * n, n)
        rays += out["stats"]["n_final_photon"]
    dt = time.perf_counter() - t0
    v = rays / dt
    line = {
        "impl": "reference", "metric": "rays traced to observer/sec (GR mode)", "value": v, "unit": "rays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "events_per_step": n, "seed": SEED,
                   "sample": f"bounded sample of the workload: {n} events per step on the host CPU",
                   "integrator": "Vern6 (the oracle's restatement of the reference's solve(prob, Vern6()), RT.jl:383)"},
        "cpu_baseline": {"value": v, "unit": "rays/s", "cores": cores, "kind": "port",
                         "sample": f"{n} events/step x {args.steps} steps, oracle/ OpenMP on {cores} threads (Vern6; a faster-than-"
                                   "reference port: no per-call allocation, no GC per event)"},
        "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


_JSON_FD = None


def emit(line):
    """The ONE JSON line goes to the process's original stdout; everything else printed on fd 1 while the bench runs
    (NCCL's version banner, library chatter) is diverted to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    args = parse()
    NS["integrator"] = 0 if args.integrator == "vern6" else 1
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from adiabatic_raytracer_b200 import lib
    from adiabatic_raytracer_b200.params import ArtParams

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    p = ArtParams.default(**NS)
    h = lib.Handle(p, local)
    h.hist_config(16, 32)
    if world > 1:  # the library's own NCCL communicator; torch only ships the 128-byte id around
        uid = [lib.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        h.comm_init(world, rank, uid[0])

    E = args.events
    ncols = h.ncols

    def first_event(step):  # disjoint event-id ranges per (step, rank): per-event Philox streams
        return 1 + (step * world + rank) * E

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- warm-up ------------------------------------------------------------------------------
    for w in range(args.warmup):
        h.run_device(SEED, first_event(1000 + w), E)
        h.comm_allreduce_hist()
    h.run_device_batches(SEED, first_event(1100), E, 2, world * E)
    fp64_peak, _ = h.fp64_peak()

    # ---- timed: device-resident (value) ----------------------------------------------------------
    # The K steps are submitted as ONE pipelined job (art_run_device_batches): step s of this rank = events
    # first_event(s) + [0, E); the engine refills the ray slots that the late tree nodes of step s free with events of
    # step s + 1, so only the last step drains alone.  One all-reduce of the binned flux closes the job.
    clocks = ClockSampler(local)
    barrier()
    if rank == 0:
        clocks.start()
    h.hist_reset()
    t_wall0 = time.perf_counter()
    out = h.run_device_batches(SEED, first_event(0), E, args.steps, world * E)
    ta = time.perf_counter()
    h.comm_allreduce_hist()          # the path's only collective
    ms_coll = (time.perf_counter() - ta) * 1e3
    agg = out["stats"]
    ms_dev = agg["ms_total"] + ms_coll
    barrier()
    wall_dev = (time.perf_counter() - t_wall0) * 1e3
    clk = clocks.stop() if rank == 0 else None
    ms_dev = max_over_ranks(ms_dev)
    wall_dev = max_over_ranks(wall_dev)
    rays_local = agg["n_final_photon"]
    rays_total = sum_over_ranks(rays_local)
    value = rays_total / (ms_dev * 1e-3)

    # ---- hist_check (outside the timed region): the NCCL-reduced block must equal the sum of the per-rank blocks ----
    loc = torch.from_numpy(h.hist_get()).cuda()
    glob = h.hist_get_global()
    if world > 1:
        dist.all_reduce(loc, op=dist.ReduceOp.SUM)
    ref = loc.cpu().numpy()
    nb = ref.size - 8
    counts_ok = np.array_equal(ref[3:nb:4], glob[3:nb:4]) and np.array_equal(ref[nb:], glob[nb:])
    sums_ok = np.allclose(ref[:nb], glob[:nb], rtol=1e-12, atol=0.0)
    hist_check = "ok" if (counts_ok and sums_ok and glob[nb + 2] == world * E * args.steps) else "MISMATCH"

    # one step as a single call (its tail drains alone), for comparison with the pipelined figure
    single = h.run_device(SEED, first_event(2000), E)["stats"]["ms_total"]

    # ---- timed: end to end through the C ABI with HOST buffers (e2e) ---------------------------------------------
    # the call a user makes for a production run: batches as one job, rows ordered on the device and copied into a
    # page-locked host buffer, all-reduced histogram read back.  Chunks of at most 8 steps bound the host buffer (≈ 2 GB).
    chunk = min(args.steps, 8)
    pinned = lib.PinnedRows(int(chunk * E * 2.3) + 1024, ncols)  # e2e result buffer: page-locked host memory (art_host_alloc)
    host_rows = pinned.array
    # untimed: one chunk at full size, so that the device-side ordering buffers and row buffers have their final size
    o = h.run_device_batches(SEED, first_event(3000), E, chunk, world * E)
    h.fetch_rows(o["n_rows"], rows_out=host_rows)
    barrier()
    t0 = time.perf_counter()
    e2e_rays = 0
    d2h = 0
    done = 0
    e2e_parts = {"run_s": 0.0, "fetch_rows_s": 0.0, "reduce_and_hist_s": 0.0}
    while done < args.steps:
        nb_ = min(chunk, args.steps - done)
        ta = time.perf_counter()
        o = h.run_device_batches(SEED, first_event(done), E, nb_, world * E)
        tb = time.perf_counter()
        rows = h.fetch_rows(o["n_rows"], rows_out=host_rows)
        tc = time.perf_counter()
        h.comm_allreduce_hist()
        hist = h.hist_get_global()
        td = time.perf_counter()
        e2e_parts["run_s"] += tb - ta
        e2e_parts["fetch_rows_s"] += tc - tb
        e2e_parts["reduce_and_hist_s"] += td - tc
        e2e_rays += o["stats"]["n_final_photon"]
        d2h += rows.shape[0] * ncols * 8 + hist.nbytes + 16 * 8
        done += nb_
    barrier()
    wall_e2e = max_over_ranks(time.perf_counter() - t0)
    e2e_total = sum_over_ranks(e2e_rays)
    e2e_value = e2e_total / wall_e2e

    # ---- comparison figures (untimed region, rank-local; only at N = 1 to keep the scaling runs short) ---------------
    extra = {}
    if world == 1:
        ks = min(args.steps, 3)
        # BASELINE configs' literal "P = 1 s" (rotW = 2π) GR workload
        hp = lib.Handle(p.copy(omega_pul=2 * np.pi), local)
        hp.run_device(SEED, 1, E)
        op = hp.run_device_batches(SEED, 1 + E, E, ks)
        hp.close()
        extra["workload_p1s"] = "same as workload but P = 1 s (rotW = 2π)"
        extra["value_p1s"] = op["stats"]["n_final_photon"] / (op["stats"]["ms_total"] * 1e-3)
        extra["ms_per_step_p1s"] = op["stats"]["ms_total"] / ks
        # the other integrator
        other = 1 - p.integrator
        ho = lib.Handle(p.copy(integrator=other), local)
        ho.run_device(SEED, 1, E)
        oo = ho.run_device_batches(SEED, 1 + E, E, ks)
        ho.close()
        extra["value_other_integrator"] = {"integrator": "DP5" if other == 1 else "Vern6",
                                           "value": oo["stats"]["n_final_photon"] / (oo["stats"]["ms_total"] * 1e-3),
                                           "ms_per_step": oo["stats"]["ms_total"] / ks,
                                           "n_rhs_per_ray": oo["stats"]["n_rhs"] / oo["stats"]["n_rays"]}

    # ---- per-kernel device times of one job (CUDA events around every launch of the main stream; a separate, untimed run) --
    os.environ["ART_TIME_KERNELS"] = "1"
    tk = h.run_device_batches(SEED, first_event(4000), E, min(args.steps, 2), world * E)["stats"]
    del os.environ["ART_TIME_KERNELS"]
    hbm, hbm_src = hbm_peak()
    det_bytes = tk["n_scanned_steps"] * DETECT_BYTES_PER_STEP
    det_gbs = det_bytes / (tk["ms_k_detect"] * 1e-3) / 1e9 if tk["ms_k_detect"] > 0 else None
    kernels = {
        "run": f"{min(args.steps, 2)} steps with ART_TIME_KERNELS=1 (events around every main-stream launch), untimed region",
        "ms_total": tk["ms_total"], "ms_k_stages": tk["ms_k_stages"], "ms_k_scan": tk["ms_k_scan"], "ms_k_detect": tk["ms_k_detect"],
        "ms_sampler_overlapped": tk["ms_sampler"],
        "share_of_step": {k: tk[k] / tk["ms_total"] for k in ("ms_k_stages", "ms_k_scan", "ms_k_detect")},
        "compaction": {"kernel": "k_wf_detect (ballot/popc compaction of the scanned rays into next-step / root / finished lists)",
                       "scanned_steps": tk["n_scanned_steps"], "algorithmic_bytes": det_bytes,
                       "bytes_per_step": DETECT_BYTES_PER_STEP, "ms": tk["ms_k_detect"], "GBps": det_gbs,
                       "frac_of_hbm_peak": det_gbs / hbm if det_gbs else None, "hbm_peak_GBps": hbm, "hbm_peak_source": hbm_src},
    }

    # ---- roofline: algorithmic FP64 flop of the path's FP64-bound kernels over the device time of the job ------------------
    # tracer: k_wf_stages (8 RHS per attempted Vern6 step + 3 dense-output RHS per scanned step) + k_wf_scan / k_wf_root
    # (crossing conditions actually EVALUATED: n_cond_eval; the reference's remaining per-step evaluations are decided by
    # the certified-sign shortcuts and are NOT credited); sampler: k_sample_events, which now overlaps the tracer on its
    # own stream, so the two cannot be timed apart: flop of both over the job's total device time.
    n_cond_eval = agg.get("n_cond_eval", 0) or agg["n_cond"]
    flop_trace = agg["n_rhs"] * F_RHS + n_cond_eval * F_COND
    flop_samp = agg["n_sampler_cond"] * F_SAMPLER_COND
    t_job = agg["ms_total"] * 1e-3
    achieved = (flop_trace + flop_samp) / t_job / 1e12
    ncu = None
    if os.path.exists(NCU_SUMMARY):
        try:
            ncu = json.load(open(NCU_SUMMARY))
        except Exception:
            ncu = None
    roofline = {
        "bound": "fp64_alu", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
        "traffic": (ncu or {}).get("dram_bytes_per_step"),
        "kernel": "the FP64-bound kernels of the job: wavefront tracer (k_wf_stages + k_wf_scan + k_wf_root) and surface sampler "
                  "(k_sample_events), which overlap on separate streams; sum of their algorithmic flop over the job's device time",
        "peak_source": "measured in this run by art_fp64_peak (register-resident DFMA chain, all SMs); "
                       "MEASURED_PEAKS.json has no FP64 entry (hbm_gbs / bf16 only)",
        "algorithmic": f"F_rhs={F_RHS:.0f}, F_cond={F_COND:.0f} flop (SURVEY §8d) x measured n_rhs, n_cond_eval, n_sampler_cond "
                       "(conditions actually evaluated)",
        "flop_per_step": {"tracer": flop_trace / args.steps, "sampler": flop_samp / args.steps},
        "frac_tracer_only": flop_trace / t_job / 1e12 / fp64_peak,
        "per_ray": {"n_acc": agg["n_acc"] / agg["n_rays"], "n_rej": agg["n_rej"] / agg["n_rays"],
                    "n_rhs": agg["n_rhs"] / agg["n_rays"], "n_cond_reference": agg["n_cond"] / agg["n_rays"],
                    "n_cond_eval": n_cond_eval / agg["n_rays"], "flop": flop_trace / agg["n_rays"],
                    "scanned_steps": agg.get("n_scanned_steps", 0) / agg["n_rays"]},
        "frac_if_every_reference_evaluation_counted": (agg["n_rhs"] * F_RHS + agg["n_cond"] * F_COND + flop_samp) / t_job / 1e12 / fp64_peak,
        "sampler": {"cond_per_attempt": agg["n_sampler_cond"] / agg["n_attempts"], "ms_kernels_per_step": agg["ms_sampler"] / args.steps},
        "ms_per_step": {"total": agg["ms_total"] / args.steps, "loop": agg["ms_tracer"] / args.steps},
        "kernels": kernels,
        # hardware view — NOT measured in this run: ncu --set full / launch list of this build, profiles/r02_ncu_summary.json
        "ncu": ncu,
    }

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _, _, _ = cpu_baseline(args.cpu_seconds)

    if rank == 0:
        cfg = {
            "workload": WORKLOAD,
            "events_per_gpu_per_step": E, "seed": SEED, "reltol": p.reltol, "abstol": p.ode_err,
            "interp_points_ray": p.interp_points_ray, "interp_points_sampler": p.interp_points_sampler,
            "integrator": ("Vern6 (Verner 9-stage 6(5), the reference's solve(prob, Vern6()) RT.jl:383), lazy 6th-order dense output"
                           if p.integrator == 0 else "DP5(4) FSAL (round 1's scheme)") + ", OrdinaryDiffEq PI controller",
            "l2": "working set ~9 KB per event slot x 2^20 slots ~ 9 GB >> 126 MB L2, no flush needed",
            "events_per_s": agg["n_events"] * world / (ms_dev * 1e-3),
            "ray_propagations_per_s": agg["n_rays"] * world / (ms_dev * 1e-3),
            "wall_ms_per_step": wall_dev / args.steps,
            "submission": f"the {args.steps} steps are one pipelined job (art_run_device_batches): ray slots freed by step s are "
                          "refilled with events of step s+1, only the last step drains alone",
            "ms_per_step_single_call": single,
            "pool_blocks_per_step": agg.get("n_pool_blocks", 0) / args.steps,
        }
        cfg.update(extra)
        line = {
            "metric": "rays traced to observer/sec (GR mode)", "value": value, "unit": "rays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "hist_check": hist_check,
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": 216,
                    "d2h_bytes_per_step": int(d2h / args.steps), "parts_rank0_s": e2e_parts,
                    "note": "art_run_device_batches + art_fetch_rows + art_comm_allreduce_hist + art_hist_get_global through the C ABI: "
                            "inputs are the 216-byte art_params + seed (the path has no bulk input), outputs = rows (ordered on the "
                            "device, one DMA copy into a page-locked host buffer) + histogram; chunks of <= 8 steps"},
            "gpu_launches": int(agg["n_launches"] + 1),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "clocks": clk,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    host_rows = None
    pinned.close()
    h.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
