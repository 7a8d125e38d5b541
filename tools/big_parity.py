"""SURVEY 8(d) parity bar (i) at scale, on a GPU box: >= 1e8 rays of the full-size C4 workload (1 M triangles, 1920x1080),
every camera sample's radiance compared bit for bit with the CPU oracle, plus the ray and box/triangle/instance test counters.
   gpurun -- 'python tools/big_parity.py > gpurun_out/big_parity.log'"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tray_rust_b200 import _ffi as F, api, scenebuild as SB  # noqa: E402
from oracle import pyoracle as O  # noqa: E402

SPP = int(sys.argv[1]) if len(sys.argv) > 1 else 6
KEYS = ["camera_samples", "rays_primary", "rays_shadow", "rays_mis", "rays_continuation", "node_tests", "tri_tests", "inst_tests"]
t0 = time.time()
desc = SB.scene_c4(1_000_000, 1920, 1080, 4096).finish()
g, o = api.Scene(desc), O.OracleScene(desc)
g.update_frame(0, 0.0, 0.0); o.update_frame(0, 0.0, 0.0)
t1 = time.time()
kw = dict(sample_first=0, sample_count=SPP, seed=1)
gs, gst = g.render_samples(flags=F.RENDER_STATS | F.RENDER_REFERENCE_SHADOW, **kw)   # reference-equivalent closest-hit shadow rays: counters comparable
t2 = time.time()
gs2, _ = g.render_samples(**kw)                                                      # product default (any-hit shadow rays, phased trace kernel)
t3 = time.time()
os_, ost = o.render_samples(**kw)
t4 = time.time()
same = gs.tobytes() == os_.tobytes()
same2 = gs2.tobytes() == os_.tobytes()
cnt = {k: (int(getattr(gst, k)), int(getattr(ost, k))) for k in KEYS}
rays = sum(cnt[k][1] for k in KEYS[1:5])
bad = int(np.count_nonzero(gs.view(np.uint8).reshape(len(gs), -1) != os_.view(np.uint8).reshape(len(os_), -1))) if not same else 0
print(json.dumps({"workload": "C4 full size: 1M triangles, 1920x1080, %d spp" % SPP, "camera_samples": int(len(gs)), "rays_compared": rays,
                  "radiance_bit_exact_reference_shadow_mode": bool(same), "radiance_bit_exact_default_mode": bool(same2), "differing_bytes": bad,
                  "counters_gpu_vs_oracle": cnt, "counters_equal": all(a == b for a, b in cnt.values()),
                  "seconds": {"scene_build_both": round(t1 - t0, 1), "gpu_stats_mode": round(t2 - t1, 2), "gpu_default": round(t3 - t2, 2), "oracle": round(t4 - t3, 1)}}))
sys.exit(0 if (same and same2 and all(a == b for a, b in cnt.values())) else 1)
