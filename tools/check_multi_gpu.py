#!/usr/bin/env python
"""Multi-GPU correctness of the product entry point on real hardware (run under torchrun on N GPUs of one box):

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/check_multi_gpu.py [events]

Every rank runs its share of ONE seeded job through main_runner_tree_multi (library-owned NCCL communicator, id through a
file) and writes its tree_*.npy and its local histogram block.  Rank 0 then checks, on numbers produced by CUDA + NCCL:
  * the all-reduced histogram block (bins and trailing counters) = the sum of the per-rank blocks: counts bit-exact,
    sums to 1e-12;
  * a second all-reduce without a reset gives the same block (no compounding);
  * the concatenation of the per-rank files = the file of the same job run on one GPU, bit for bit
    (per-event random streams; global f_inx normalisation, MainRunner.jl:749).
Prints one JSON line; exit code 0 only if everything holds.
"""
import json
import os
import shutil
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adiabatic_raytracer_b200 import main_runner_tree_multi  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
base = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/check_multi"
d = os.path.join(base, f"w{world}")
if rank == 0:
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d, exist_ok=True)
    open(os.path.join(d, "go"), "w").close()
while not os.path.exists(os.path.join(d, "go")):
    time.sleep(0.05)
args = (1e-5, 1e-12, 0.2, 1.0, 1e14, 10.0, 1.0, "Simple", n + 1, [1.0, 1.0])
kw = dict(flat=False, melrose=True, ode_err=1e-6, ntimes=3, iseed=1769)
r = main_runner_tree_multi(*args, rank=rank, world=world, device=local, file_tag="m_", dir_tag=d, **kw)
np.save(os.path.join(d, f"hist_local_{rank}.npy"), r["hist_local"])
np.save(os.path.join(d, f"hist_global_{rank}.npy"), r["hist"])
open(os.path.join(d, f"done_{rank}"), "w").close()
if rank != 0:
    sys.exit(0)
for q in range(world):
    while not os.path.exists(os.path.join(d, f"done_{q}")):
        time.sleep(0.05)
loc = [np.load(os.path.join(d, f"hist_local_{q}.npy")) for q in range(world)]
glo = [np.load(os.path.join(d, f"hist_global_{q}.npy")) for q in range(world)]
S = np.sum(loc, axis=0)
nb = S.size - 8
G = r["hist"]
checks = {
    "counts_exact": bool(np.array_equal(S[3:nb:4], G[3:nb:4])),
    "counters_exact": bool(np.array_equal(S[nb:], G[nb:])),
    "sums_1e-12": bool(np.allclose(S[:nb], G[:nb], rtol=1e-12, atol=0)),
    "same_block_on_every_rank": bool(all(np.array_equal(g, G) for g in glo)),
    "events_total": bool(G[-6] == n),
}
# the same job on one GPU (this process, device `local`)
d1 = os.path.join(base, f"w{world}_single")
shutil.rmtree(d1, ignore_errors=True)
r1 = main_runner_tree_multi(*args, rank=0, world=1, device=local, file_tag="m_", dir_tag=d1, **kw)
parts = [np.load(os.path.join(d, "npy", f)) for f in sorted(os.listdir(os.path.join(d, "npy")), key=lambda s: int(s.rsplit("_", 1)[1][:-4]))]
cat = np.vstack(parts)
one = np.load(r1["file"])
checks["files_concatenate_to_the_single_gpu_file"] = bool(cat.shape == one.shape and np.array_equal(cat, one))
checks["f_inx_total_equal"] = bool(r["f_inx_total"] == r1["f_inx_total"])
checks["single_gpu_hist_counts_equal"] = bool(np.array_equal(r1["hist"][3:nb:4], G[3:nb:4]) and np.array_equal(r1["hist"][nb:], G[nb:]))
checks["single_gpu_hist_sums_1e-10"] = bool(np.allclose(r1["hist"][:nb], G[:nb], rtol=1e-10, atol=0))
ok = all(checks.values())
print(json.dumps({"check": "multi_gpu", "world": world, "events": n, "ok": ok, **checks, "rows": int(cat.shape[0]),
                  "f_inx_total": r["f_inx_total"]}), flush=True)
sys.exit(0 if ok else 1)
