"""Developer tool (GPU box): time the trace-kernel scheduling variants on the C4 scene in ONE process.
   gpurun -- 'python tools/sched_sweep.py > gpurun_out/sched_sweep.log'
TRB_TRACE_SCHED / TRB_REFILL are re-read by the library at every launch, so the scene is built once."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tray_rust_b200 import _ffi as F, api, scenebuild as SB  # noqa: E402

W, H, SPP_STEP, STEPS = 1920, 1080, 8, 2
dev = torch.device("cuda:0")
film = torch.zeros(H, W, 4, dtype=torch.float32, device=dev)
stats = torch.zeros(10, dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def scene(depth0):
    b = SB.scene_c4(1_000_000, W, H, 4096)
    if depth0:
        b.integrator = (0, 0, 0)
    g = api.Scene(b.finish(), 0)
    g.update_frame(0, 0.0, 0.0)
    return g


def measure(g, sched, refill=8, flags=0, quads=1, occ=7, sst=16):
    os.environ["TRB_SMEM_STACK"] = str(sst)
    os.environ["TRB_TRACE_SCHED"] = str(sched)
    os.environ["TRB_REFILL"] = str(refill)
    os.environ["TRB_TRACE_QUADS"] = str(quads)
    os.environ["TRB_TRACE_OCC"] = str(occ)
    g.render_device(film.data_ptr(), stats.data_ptr(), None, spp=4096, sample_first=0, sample_count=SPP_STEP, seed=1, flags=flags)
    torch.cuda.synchronize(); stats.zero_()
    ms = 0.0
    for i in range(1, 1 + STEPS):
        flush.fill_(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.render_device(film.data_ptr(), stats.data_ptr(), None, spp=4096, sample_first=i * SPP_STEP, sample_count=SPP_STEP, seed=1, flags=flags)
        e1.record(); torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    s = stats.cpu().numpy()
    return float(s[1:5].sum()) / ms / 1e3, ms / STEPS


def sched(quorum, burst):
    return quorum | burst << 8


# (name, sched word, refill, quads, trace occ, smem stack entries)
CASES = [("flat pairs", 0, 8, 0, 7, 16), ("quorum 6 (default)", 6, 8, 0, 7, 16), ("quorum 4", 4, 8, 0, 7, 16), ("quorum 8", 8, 8, 0, 7, 16), ("quorum 6 sst12", 6, 8, 0, 7, 12),
         ("quorum 6 occ8", 6, 8, 0, 8, 16), ("quorum 6 refill 6", 6, 6, 0, 7, 16), ("quorum 6 refill 10", 6, 10, 0, 7, 16), ("quads quorum 6", 6, 8, 1, 7, 16)]
if __name__ == "__main__":
    full, direct = scene(False), scene(True)
    for name, sc, r, q, occ, sst in CASES:
        v, ms = measure(full, sc, r, quads=q, occ=occ, sst=sst)
        v0, ms0 = measure(direct, sc, r, quads=q, occ=occ, sst=sst)
        print("%-24s full path %7.1f Mrays/s (%6.1f ms/step)   primary+shadow %7.1f Mrays/s (%5.1f ms/step)" % (name, v, ms, v0, ms0), flush=True)
