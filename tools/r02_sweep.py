"""Developer tool (GPU box): A/B the round-2 launch options on the bench workload (C4, 1920x1080, 8 spp per step) in ONE
process: ray-queue sorting (mode / grid bits / first round), split shading. Prints one JSON line per variant
with the step time, the trace-kernel share and a bit-exactness check of the per-sample radiance against the first variant.
   gpurun -- 'python tools/r02_sweep.py > gpurun_out/r02_sweep.log'"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tray_rust_b200 import _ffi as F, api, scenebuild as SB  # noqa: E402

W, H, SPP_STEP, STEPS = 1920, 1080, 8, 3
dev = torch.device("cuda:0")
film = torch.zeros(H, W, 4, dtype=torch.float32, device=dev)
stats = torch.zeros(10, dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
g = api.Scene(SB.scene_c4(1_000_000, W, H, 4096).finish(), 0)
g.update_frame(0, 0.0, 0.0)
DEFAULTS = {"sort.mode": 0, "sort.bits": 5, "sort.min_round": 1, "shade.split": 0, "trace.refill": 8, "trace.sched": 6, "trace.grid": 0,
            "trace.pipe": 36, "shade.kind": 1}
ref = None


def measure(name, **opts):
    global ref
    cfg = dict(DEFAULTS)
    cfg.update({k.replace("_", ".", 1): v for k, v in opts.items()})
    for k, v in cfg.items():
        g.set_option(k, v)
    s, _ = g.render_samples(block_start=9000, block_count=600, sample_first=0, sample_count=4, seed=1)
    same = None
    if ref is None:
        ref = s.tobytes()
    else:
        same = s.tobytes() == ref
    g.render_device(film.data_ptr(), stats.data_ptr(), None, spp=4096, sample_first=0, sample_count=SPP_STEP, seed=1)
    torch.cuda.synchronize(); stats.zero_()
    ms = 0.0
    for i in range(1, 1 + STEPS):
        flush.fill_(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.render_device(film.data_ptr(), stats.data_ptr(), None, spp=4096, sample_first=i * SPP_STEP, sample_count=SPP_STEP, seed=1)
        e1.record(); torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    rays = int(stats.cpu().numpy()[1:5].sum())
    for i in range(1, 1 + STEPS):
        g.render_device(film.data_ptr(), stats.data_ptr(), None, spp=4096, sample_first=i * SPP_STEP, sample_count=SPP_STEP, seed=1, flags=F.RENDER_TIME_TRACE)
    torch.cuda.synchronize()
    tms, tn = g.trace_time()
    print(json.dumps({"variant": name, "opts": opts, "ms_per_step": ms / STEPS, "mrays_s": rays / ms / 1e3, "trace_ms_per_step": tms / STEPS,
                      "other_ms_per_step": (ms - tms) / STEPS, "bit_exact_vs_first": same}), flush=True)


VARIANTS = os.environ.get("SWEEP", "shade").split(",")
measure("default configuration")
if "shade" in VARIANTS:
    measure("split shade", shade_split=1)
    measure("trace refill 12", trace_refill=12)
if "sort" in VARIANTS:
    measure("split shade", shade_split=1)
    for bits in (4, 5, 6):
        measure("sort octant-major bits=%d" % bits, sort_mode=1, sort_bits=bits)
    measure("sort cell-major bits=5", sort_mode=2, sort_bits=5)
    measure("sort octant-major bits=5 from round 2", sort_mode=1, sort_bits=5, sort_min_round=2)
    measure("sort octant-major bits=5 + split", sort_mode=1, sort_bits=5, shade_split=1)
if "pipe" in VARIANTS:  # trace.pipe: kernel variant (trb_api.cu Tuning::pipe)
    for pipe in (38, 39, 36):
        measure("trace.pipe=%d" % pipe, trace_pipe=pipe)
    measure("default again")
if "final" in VARIANTS:  # the round's trace-kernel steps side by side, in one process
    measure("trace.pipe=0 (round-1 kernel)", trace_pipe=0)
    measure("trace.pipe=1 (+ box_hit_finite)", trace_pipe=1)
    measure("trace.pipe=33 (+ RayHome + fused non-node chains, 7 CTAs per SM)", trace_pipe=33)
    measure("trace.pipe=34 (8 CTAs per SM)", trace_pipe=34)
    measure("trace.pipe=36 (9 CTAs per SM) = default", trace_pipe=36)
    measure("split shade", shade_split=1)
if "kind" in VARIANTS:  # the matte instantiations of the split shade kernels against the generic split kernels and the fused kernel (C4 is all matte)
    measure("split + matte instantiations = per-scene default", shade_split=-1, shade_kind=1)
    measure("split, generic kernels", shade_split=1, shade_kind=0)
    measure("fused shade kernel", shade_split=0)
