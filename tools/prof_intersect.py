"""Traversal microbench: primary rays of C4 and a diffuse bounce-1 ray set through trb_intersect_device."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tray_rust_b200 import _ffi as F, api, scenebuild as SB
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--tris", type=int, default=1000000)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--w", type=int, default=1920)
ap.add_argument("--h", type=int, default=1080)
a = ap.parse_args()
t0 = time.time()
g = api.Scene(SB.scene_c4(a.tris, a.w, a.h, 4096).finish())
print("scene_create %.2fs" % (time.time() - t0), flush=True)
g.update_frame(0, 0.0, 0.0)
rays, _ = g.camera_rays(sample_first=0, sample_count=1, seed=1)
hits, st = g.intersect(rays)
n = len(rays)
print("primary: n=%d node/ray=%.1f tri/ray=%.2f inst/ray=%.2f hit%%=%.1f" % (n, st.node_tests / n, st.tri_tests / n, st.inst_tests / n, 100 * np.mean(hits["inst"] != F.MISS)))
rng = np.random.default_rng(5)
hit = np.nonzero(hits["inst"] != F.MISS)[0]
sec = np.zeros(len(hit), F.RAY_DTYPE)
sec["o"] = rays["o"][hit] + rays["d"][hit] * hits["t"][hit, None]
d = rng.normal(size=(len(hit), 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
sec["d"] = d; sec["min_t"] = 0.001; sec["max_t"] = np.inf
h2, st2 = g.intersect(sec)
m = len(sec)
print("bounce1: n=%d node/ray=%.1f tri/ray=%.2f inst/ray=%.2f hit%%=%.1f" % (m, st2.node_tests / m, st2.tri_tests / m, st2.inst_tests / m, 100 * np.mean(h2["inst"] != F.MISS)))
for name, r, s in (("primary", rays, st), ("bounce1", sec, st2)):
    k = len(r)
    d_rays = torch.from_numpy(r.view(np.float32).reshape(-1, 8).copy()).cuda()
    d_hits = torch.zeros((k, 4), dtype=torch.int32, device="cuda")
    best = 1e9
    for it in range(a.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.intersect_device(k, d_rays.data_ptr(), d_hits.data_ptr()); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    alg = 48 * k + 32 * s.node_tests + 48 * s.tri_tests + 64 * s.inst_tests
    print("%s: %.3f ms  %.1f Mrays/s  algorithmic %.1f GB/s" % (name, best, k / best / 1e3, alg / best / 1e6), flush=True)
