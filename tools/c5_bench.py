"""Developer tool (GPU box): the C5-shaped keyframed scene (tests/golden/scenes/c5_tr15_like.json, BASELINE configs[4] stand-in)
at 1920x1080 through the keyframed kernel variants: parity spot-check against the oracle on a block range + a timing line.
   gpurun -- 'python tools/c5_bench.py > gpurun_out/c5.log'"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
from tray_rust_b200 import _ffi as F, api  # noqa: E402
from oracle import pyoracle as O  # noqa: E402
import make_scenes  # noqa: E402

W, H, SPP_STEP, FRAME = 1920, 1080, 8, 12
merl = os.path.join(REPO, "tests", "golden", "scenes", "merl", "synthetic.binary")
if not os.path.exists(merl):
    make_scenes.write_synthetic_merl(merl)
lib = F.load_trb()
d = C.POINTER(F.SceneDesc)()
assert lib.trb_desc_load_json(os.path.join(REPO, "tests", "golden", "scenes", "c5_tr15_like.json").encode(), W, H, 2048, C.byref(d)) == 0
desc = d.contents
step = desc.film.scene_time / desc.film.frames
g, o = api.Scene(desc, 0), O.OracleScene(desc)
g.update_frame(FRAME, FRAME * step, (FRAME + 1) * step); o.update_frame(FRAME, FRAME * step, (FRAME + 1) * step)
kw = dict(block_start=12000, block_count=64, sample_first=0, sample_count=4, seed=1)
gs, _ = g.render_samples(**kw); os_, _ = o.render_samples(**kw)
parity = gs.tobytes() == os_.tobytes()
dev = torch.device("cuda:0")
film = torch.zeros(H, W, 4, dtype=torch.float32, device=dev)
stats = torch.zeros(10, dtype=torch.int64, device=dev)
g.render_device(film.data_ptr(), stats.data_ptr(), None, spp=2048, sample_first=0, sample_count=SPP_STEP, seed=1)
torch.cuda.synchronize(); stats.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(1, 4):
    g.render_device(film.data_ptr(), stats.data_ptr(), None, spp=2048, sample_first=i * SPP_STEP, sample_count=SPP_STEP, seed=1)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
s = stats.cpu().numpy()
print(json.dumps({"workload": "c5_tr15_like.json 1920x1080 frame %d (keyframed camera/instances/emission, 12 instances, MERL), 3 passes x %d spp" % (FRAME, SPP_STEP),
                  "bit_exact_vs_oracle_on_64_blocks": bool(parity), "mrays_s": float(s[1:5].sum()) / ms / 1e3, "msamples_s": float(s[0]) / ms / 1e3,
                  "ms_per_pass": ms / 3, "rays": {"primary": int(s[1]), "shadow": int(s[2]), "mis": int(s[3]), "continuation": int(s[4])}}))
