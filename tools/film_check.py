"""Developer tool (GPU box): the atomic-free film kernel (TRB_FILM_V2=1) against the default one — equality within the film
tolerance on small scenes (image borders, wide Gaussian filter) and timing on the C4 workload.
   gpurun -- 'python tools/film_check.py > gpurun_out/film_check.log'"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tray_rust_b200 import _ffi as F, api, scenebuild as SB  # noqa: E402
from oracle import pyoracle as O  # noqa: E402


def film(g, v2, **kw):
    os.environ["TRB_FILM_V2"] = "1" if v2 else "0"
    f, _ = g.render(**kw)
    return f


out = {"small": []}
for name, b in (("zoo64", SB.scene_materials_zoo(64, 64, 8, SB.synthetic_merl_table())), ("tiny16", SB.scene_smallpt_like(16, 16, 8)), ("c4_20k_wide", SB.scene_c4(20000, 128, 72, 8))):
    if name == "c4_20k_wide":
        b.film.update(filter_type=F.FILTER_GAUSSIAN, filter_w=3.0, filter_h=2.5, filter_b=0.5, filter_c=0.0)
    desc = b.finish()
    g, o = api.Scene(desc), O.OracleScene(desc)
    f1, f2 = film(g, False, seed=3), film(g, True, seed=3)
    fo, _ = o.render(seed=3)
    n = lambda f: f[..., :3] / np.maximum(f[..., 3:], 1e-6)
    out["small"].append({"scene": name, "v2_vs_v1_max_abs": float(np.abs(f1 - f2).max()), "v2_vs_v1_allclose": bool(np.allclose(f1, f2, rtol=1e-4, atol=1e-5)),
                         "v1_rmse_vs_oracle": float(np.sqrt(np.mean((n(f1) - n(fo)) ** 2))), "v2_rmse_vs_oracle": float(np.sqrt(np.mean((n(f2) - n(fo)) ** 2))),
                         "weights_equal": bool(np.allclose(f1[..., 3], f2[..., 3], rtol=1e-5, atol=1e-6))})
    g.close()

dev = torch.device("cuda:0")
W, H = 1920, 1080
fbuf = torch.zeros(H, W, 4, dtype=torch.float32, device=dev)
stats = torch.zeros(10, dtype=torch.int64, device=dev)
for tag, depth0 in (("primary_shadow", True), ("full_path", False)):
    b = SB.scene_c4(1_000_000, W, H, 4096)
    if depth0:
        b.integrator = (0, 0, 0)
    g = api.Scene(b.finish(), 0)
    g.update_frame(0, 0.0, 0.0)
    res = {}
    for v2 in (0, 1):
        os.environ["TRB_FILM_V2"] = str(v2)
        g.render_device(fbuf.data_ptr(), stats.data_ptr(), None, spp=4096, sample_first=0, sample_count=8, seed=1)
        torch.cuda.synchronize(); stats.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(1, 4):
            g.render_device(fbuf.data_ptr(), stats.data_ptr(), None, spp=4096, sample_first=8 * i, sample_count=8, seed=1)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        s = stats.cpu().numpy()
        res["v2" if v2 else "v1"] = {"ms_per_step": ms / 3, "mrays_s": float(s[1:5].sum()) / ms / 1e3}
    out[tag] = res
    g.close()
print(json.dumps(out))
