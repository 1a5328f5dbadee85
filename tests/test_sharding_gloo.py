"""World-size-2 gloo test of the multi-GPU host logic on CPU: disjoint event shards, all-reduce of the binned
flux block, global f_inx normalisation.  The per-rank compute is done by the CPU oracle (test infrastructure);
on a GPU box the same logic runs with libart_b200.so + NCCL (bench.py --gpus N)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from adiabatic_raytracer_b200 import sharding
from adiabatic_raytracer_b200.params import ArtParams

NCT, NPH = 6, 8
N_PER_RANK = 48
SEED = 1769


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import loader as oracle
    p = ArtParams.default(theta_m=0.2)
    first, n = sharding.event_range(rank, world, 0, N_PER_RANK)
    out = oracle.run(p, SEED, first, n)
    H = sharding.binned_flux(out["rows"], out["f_inx"], NCT, NPH)
    t = torch.from_numpy(H)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)   # the path's only collective
    np.save(os.path.join(out_dir, f"hist_{rank}.npy"), t.numpy())
    np.save(os.path.join(out_dir, f"rows_{rank}.npy"), out["rows"])
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    H0 = np.load(tmp_path / "hist_0.npy")
    H1 = np.load(tmp_path / "hist_1.npy")
    assert np.array_equal(H0, H1)  # all ranks hold the reduced block
    # single process over the union of the two shards
    from oracle import loader as oracle
    p = ArtParams.default(theta_m=0.2)
    ref = oracle.run(p, SEED, 1, world * N_PER_RANK)
    Href = sharding.binned_flux(ref["rows"], ref["f_inx"], NCT, NPH)
    assert H0[-8] == ref["f_inx"] and H0[-7] == len(ref["rows"])
    assert np.allclose(H0[:-8], Href[:-8], rtol=1e-12, atol=0)
    rows = np.vstack([np.load(tmp_path / "rows_0.npy"), np.load(tmp_path / "rows_1.npy")])
    assert np.array_equal(rows, ref["rows"])  # per-event streams: shards concatenate to the single-process run
    flux = sharding.normalised_flux(H0, NCT, NPH)
    assert np.isclose(flux[..., 0].sum(), (ref["rows"][:, 8] * ref["rows"][:, 7]).sum() / ref["f_inx"], rtol=1e-12)


def test_event_ranges_are_disjoint():
    seen = set()
    for step in range(3):
        for rank in range(4):
            f, n = sharding.event_range(rank, 4, step, 10)
            ids = set(range(f, f + n))
            assert not (ids & seen)
            seen |= ids
    assert min(seen) == 1 and len(seen) == 120
    parts = sharding.split_events(103, 8)
    assert sum(c for _, c in parts) == 103 and parts[0][0] == 1
    assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(7))
