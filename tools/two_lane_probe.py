"""Probe: do two independent pools on one GPU (two handles, two host threads) overlap well enough to beat one pool?
python tools/two_lane_probe.py [events]"""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adiabatic_raytracer_b200 import lib
from adiabatic_raytracer_b200.params import ArtParams

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000000
p = ArtParams.default(theta_m=0.2)
h = [lib.Handle(p, 0) for _ in range(2)]
for x in h:
    x.run_device(1769, 1, 300000)   # allocations
def wall(f):
    torch.cuda.synchronize(); t = time.perf_counter(); f(); torch.cuda.synchronize(); return (time.perf_counter() - t) * 1e3
for rep in range(2):
    t1 = wall(lambda: h[0].run_device(1769, 1, n))
    def both():
        th = [threading.Thread(target=lambda i=i: h[i].run_device(1769, 1 + i * (n // 2), n // 2)) for i in range(2)]
        [t.start() for t in th]; [t.join() for t in th]
    t2 = wall(both)
    print(f"one pool {n} events: {t1:.1f} ms wall; two pools x {n//2} events concurrently: {t2:.1f} ms wall", flush=True)
for pool in ("524288",):
    os.environ["ART_POOL_SLOTS"] = pool
    g = [lib.Handle(p, 0) for _ in range(2)]
    for x in g: x.run_device(1769, 1, 300000)
    def both2():
        th = [threading.Thread(target=lambda i=i: g[i].run_device(1769, 1 + i * (n // 2), n // 2)) for i in range(2)]
        [t.start() for t in th]; [t.join() for t in th]
    for rep in range(2):
        print(f"two pools of {pool} slots: {wall(both2):.1f} ms wall", flush=True)
