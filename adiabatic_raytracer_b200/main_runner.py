"""Host-side mirror of the reference's entry point ``main_runner_tree`` (MainRunner.jl:355-765).

Same positional arguments, keyword arguments, units, return value (``tot_count``) and output
file (``<dir_tag>/npy/tree_MassAx_..._<file_tag>.npy``, Float64, 13 or 29 columns, consumed by
Combine_Files.py).  The reference's toolchain (Julia) is not installed in this image, so this
mirror is Python; ``julia/MainRunnerGPU.jl`` is the same thing written with ``ccall`` for a Julia
host.  All arithmetic happens in libart_b200.so on the GPU.

Differences a user can observe:
  * random numbers come from per-event Philox streams, not Julia's global RNG;
  * ``iseed < 0`` draws a seed from os.urandom like the reference draws rand(0:1e8);
  * saveMode = 2 writes <dir_tag>/event/final_<file_tag> and event_<file_tag> once, after the run (the
    reference appends per event); the per-event wall-clock field of event_* is 0.0;
  * saveMode = 3 writes <dir_tag>/tree/tree_<file_tag><event> in saveNode's layout (MainRunner.jl:17-65, 574-578) with the
    trajectories read off the dense output at the ``ntimes`` saveat points; an event point is saved once (DiffEq: twice);
  * ``v_NS`` is accepted and, as in the reference (MainRunner.jl:418-422 computes two angles that nothing
    reads; the velocity shift at :497 is commented out), has no effect.
"""
import math
import os
import time

import numpy as np

from . import lib as _lib
from . import sharding as _sharding
from .params import ArtParams, ART_ERR_NO_SURFACE


def params_from_call(Mass_a, Ax_g, theta_m, omega_pul, B0, rNS, Mass_NS, *, flat=True, isotropic=False,
                     melrose=False, thick_surface=True, bndry_lyr=-1, ode_err=1e-6, fix_time=0.0, ntimes=1000,
                     rho_DM=0.45, vmean_ax=220.0, saveMode=0, n_maxSample=6, num_cutoff=5, MC_nodes=5,
                     max_nodes=50, prob_cutoff=1e-10, **extra) -> ArtParams:
    """Keyword defaults are those of main_runner_tree's own signature (MainRunner.jl:356-363) —
    NOT Gen_Samples.jl's (which passes flat=false, melrose=true explicitly)."""
    if saveMode < 3:
        ntimes = 3  # MainRunner.jl:379-381
    return ArtParams.default(
        Mass_a=Mass_a, Ax_g=Ax_g, theta_m=theta_m, omega_pul=omega_pul, B0=B0, rNS=rNS, Mass_NS=Mass_NS,
        flat=int(bool(flat)), isotropic=int(bool(isotropic)), melrose=int(bool(melrose)),
        thick_surface=int(bool(thick_surface)), bndry_lyr=float(bndry_lyr), ode_err=ode_err, fix_time=fix_time,
        ntimes=int(ntimes), rho_DM=rho_DM, vmean_ax=vmean_ax, save_mode=int(saveMode), n_maxSample=int(n_maxSample),
        num_cutoff=int(num_cutoff), MC_nodes=int(MC_nodes), max_nodes=int(max_nodes), prob_cutoff=prob_cutoff, **extra)

def _check_crossing_capacity(stats, cross_cap):
    """The reference keeps every crossing of a ray (an unbounded list); the library keeps cross_cap records per ray and reports
    rays that had more as ART_RAY_CROSS_OVERFLOW.  A back-trace weight built from a truncated list would silently differ from
    the reference's, so the runner refuses the result instead (rerun with a larger cross_cap, at most 64)."""
    n_over = int(stats["n_status"][4])
    if n_over:
        raise RuntimeError(f"{n_over} rays crossed the conversion surface more than cross_cap = {cross_cap} times: their records "
                           "were truncated (ART_RAY_CROSS_OVERFLOW); rerun with a larger cross_cap")


def main_runner_tree(Mass_a, Ax_g, theta_m, omega_pul, B0, rNS, Mass_NS, omega_prop, Ntajs, gammaF, *,
                     flat=True, isotropic=False, melrose=False, thick_surface=True, bndry_lyr=-1, ode_err=1e-6,
                     cutT=100000, fix_time=0.0, CLen_Scale=True, file_tag="", ntimes=1000, v_NS=(0, 0, 0),
                     rho_DM=0.45, vmean_ax=220.0, saveMode=0, ntimes_ax=1000, dir_tag="results", n_maxSample=6,
                     iseed=-1, num_cutoff=5, MC_nodes=5, max_nodes=50, prob_cutoff=1e-10, device=0,
                     return_rows=False):
    """Drop-in for ``main_runner_tree`` (MainRunner.jl:355).  Returns tot_count (or None when
    there is no conversion surface, MainRunner.jl:389-396)."""
    if iseed < 0:  # MainRunner.jl:365-372
        iseed = int.from_bytes(os.urandom(4), "little") % 100000001
    elif iseed == 0:
        iseed = int.from_bytes(os.urandom(8), "little")
    print(f"Using seed {iseed}")
    p = params_from_call(Mass_a, Ax_g, theta_m, omega_pul, B0, rNS, Mass_NS, flat=flat, isotropic=isotropic,
                         melrose=melrose, thick_surface=thick_surface, bndry_lyr=bndry_lyr, ode_err=ode_err,
                         fix_time=fix_time, ntimes=ntimes, rho_DM=rho_DM, vmean_ax=vmean_ax, saveMode=saveMode,
                         n_maxSample=n_maxSample, num_cutoff=num_cutoff, MC_nodes=MC_nodes, max_nodes=max_nodes,
                         prob_cutoff=prob_cutoff)
    rc, maxR = _lib.max_radius(p)
    if rc == ART_ERR_NO_SURFACE:
        print("Too small Max R.... quitting.... ")
        print(f"Max radius found...\t{maxR}")
        return None
    h = _lib.Handle(p, device)
    try:
        # `while photon_trajs < desired_trajs` with photon_trajs starting at 1 ⇒ Ntajs−1 events,
        # numbered 1 … Ntajs−1 (MainRunner.jl:399,450)
        n_events = max(int(Ntajs) - 1, 0)
        out = h.run(iseed, 1, n_events)
        _check_crossing_capacity(out["stats"], p.cross_cap)
        events = h.fetch_events(n_events) if saveMode > 1 else None
        if saveMode > 2:  # MainRunner.jl:574-578: one tree_<file_tag><event> file per event
            os.makedirs(os.path.join(dir_tag, "tree"), exist_ok=True)
            h.write_tree_files(dir_tag, file_tag)
    finally:
        pass
    rows = np.ascontiguousarray(out["rows"])
    if saveMode > 1:  # MainRunner.jl:435-444, 593-610, 689-701 (the directory must exist, as in the reference)
        os.makedirs(os.path.join(dir_tag, "event"), exist_ok=True)
        _lib.write_event_files(dir_tag, file_tag, rows, events)
        rows = np.ascontiguousarray(rows[:, :29])
    if rows.shape[0] == 0:
        h.close()
        raise RuntimeError("no final particles: the reference fails here too (saveAll is nothing, MainRunner.jl:749)")
    _lib.normalize_rows(rows, out["f_inx"])  # MainRunner.jl:749
    fname = _lib.npy_name(p, int(Ntajs), dir_tag, file_tag)
    os.makedirs(os.path.dirname(fname), exist_ok=True)
    _lib.write_npy(fname, rows)
    h.close()
    if return_rows:
        return out["tot_count"], rows, fname, out
    return out["tot_count"]


def _exchange_comm_id(path: str, rank: int, timeout_s: float = 120.0) -> bytes:
    """NCCL unique id from rank 0 to the other ranks through a file (the reference's own multi-process runs share nothing
    but the file system, runner_GR_tasks.sh:22-28); any other channel (MPI, sockets, torch.distributed) works as well."""
    if rank == 0:
        uid = _lib.comm_unique_id()
        tmp = path + ".tmp"
        with open(tmp, "wb") as f:
            f.write(uid)
        os.replace(tmp, path)
        return uid
    t0 = time.time()
    while not (os.path.exists(path) and os.path.getsize(path) == 128):
        if time.time() - t0 > timeout_s:
            raise TimeoutError(f"no NCCL id at {path}")
        time.sleep(0.05)
    return open(path, "rb").read()


def main_runner_tree_multi(Mass_a, Ax_g, theta_m, omega_pul, B0, rNS, Mass_NS, omega_prop, Ntajs, gammaF, *, rank, world,
                           comm_id=None, device=None, normalisation="global", write_files=True, fetch_rows=True,
                           hist_bins=(16, 32), file_tag="", dir_tag="results", iseed=1769, **kwargs):
    """One job of ``Ntajs − 1`` events over ``world`` GPUs, one process per GPU (rank = this process).

    The reference's "multi-device" path is N independent single-process runs plus Combine_Files.py
    (runner_GR_tasks.sh:22-28, Gen_Samples.jl:195-239).  Here the events 1 … Ntajs−1 of ONE seeded job are split in
    contiguous ranges over the ranks (per-event random streams: the rows do not depend on ``world``), every rank writes its
    own ``tree_…_<file_tag><rank>.npy``, and the only collective is the all-reduce of the binned flux whose trailing
    counters carry the global ``f_inx`` (MainRunner.jl:749).

    normalisation = "global":   column 8 (sln_prob) is divided by the all-reduced f_inx — concatenating the files gives
                                exactly the single-GPU file of the same job;
                  = "per_rank": divided by this rank's own f_inx, i.e. every file is a self-contained run as the
                                reference's processes write them, to be combined by Combine_Files.py / combine_files
                                (which divide by the number of files).
    comm_id: the 128-byte NCCL id from ``lib.comm_unique_id()`` of rank 0, or None to exchange it through
    ``<dir_tag>/.nccl_id_<file_tag>``.  Returns a dict (tot_count of this rank, global counters, reduced histogram,
    file name); None when there is no conversion surface.
    """
    for k in ("saveMode", "cutT", "CLen_Scale", "v_NS", "ntimes_ax"):
        kwargs.setdefault(k, {"saveMode": 0, "cutT": 100000, "CLen_Scale": True, "v_NS": (0, 0, 0), "ntimes_ax": 1000}[k])
    saveMode = kwargs.pop("saveMode")
    for k in ("cutT", "CLen_Scale", "v_NS", "ntimes_ax"):
        kwargs.pop(k)
    if saveMode > 1:
        raise NotImplementedError("main_runner_tree_multi writes the .npy rows (saveMode 0 / 1)")
    p = params_from_call(Mass_a, Ax_g, theta_m, omega_pul, B0, rNS, Mass_NS, saveMode=saveMode, **kwargs)
    rc, maxR = _lib.max_radius(p)
    if rc == ART_ERR_NO_SURFACE:
        if rank == 0:
            print("Too small Max R.... quitting.... ")
        return None
    n_total = max(int(Ntajs) - 1, 0)
    first, count = _sharding.split_events(n_total, world)[rank]
    h = _lib.Handle(p, rank if device is None else device)
    h.hist_config(*hist_bins)
    if world > 1:
        if comm_id is None:
            os.makedirs(dir_tag, exist_ok=True)
            comm_id = _exchange_comm_id(os.path.join(dir_tag, f".nccl_id_{file_tag}"), rank)
        h.comm_init(world, rank, comm_id)
    t0 = time.perf_counter()
    out = h.run_device(iseed, first, count)
    _check_crossing_capacity(out["stats"], p.cross_cap)
    h.comm_allreduce_hist()            # the only collective: [bins | f_inx, n_rows, n_events, …]
    G = h.hist_get_global()
    wall = time.perf_counter() - t0
    f_inx_total = int(G[-_sharding.N_TRAILING])
    rows = None
    fname = None
    if fetch_rows or write_files:
        rows = np.ascontiguousarray(h.fetch_rows(out["n_rows"]))
        f_norm = f_inx_total if normalisation == "global" else out["f_inx"]
        if rows.shape[0] and f_norm > 0:
            _lib.normalize_rows(rows, f_norm)  # MainRunner.jl:749
        if write_files:
            fname = _lib.npy_name(p, int(Ntajs), dir_tag, f"{file_tag}{rank}")
            os.makedirs(os.path.dirname(fname), exist_ok=True)
            _lib.write_npy(fname, rows)
    res = dict(rank=rank, world=world, first_event=first, n_events=count, n_rows=out["n_rows"], f_inx=out["f_inx"],
               tot_count=out["tot_count"], f_inx_total=f_inx_total, n_rows_total=int(G[-_sharding.N_TRAILING + 1]),
               n_events_total=int(G[-_sharding.N_TRAILING + 2]), hist=G, hist_local=h.hist_get(), stats=out["stats"],
               wall_s=wall, file=fname, rows=rows)
    h.close()
    return res
