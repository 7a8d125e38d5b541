"""Developer tool (GPU box): BASELINE configs[4] — the reference's tr15.json with stand-in assets (tests/golden/make_tr15.py) and the
smaller tr15-shaped scene — at 1920x1080 through the keyframed kernel variants: parity spot-check against the oracle on a block
range, then timing with the per-path transform table on and off (option "anim.table").
   gpurun -- 'python tools/c5_bench.py > gpurun_out/c5.log'"""
import ctypes as C
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
from tray_rust_b200 import _ffi as F, api  # noqa: E402
from oracle import pyoracle as O  # noqa: E402
import make_scenes  # noqa: E402
import make_tr15  # noqa: E402

W, H, SPP_STEP = 1920, 1080, 4
SCENES = os.path.join(REPO, "tests", "golden", "scenes")
make_tr15.write_assets()
make_scenes.write_synthetic_merl(os.path.join(SCENES, "merl", "synthetic.binary"))
lib = F.load_trb()
dev = torch.device("cuda:0")
film = torch.zeros(H, W, 4, dtype=torch.float32, device=dev)
stats = torch.zeros(10, dtype=torch.int64, device=dev)

QUICK = "--quick" in sys.argv   # one scene, table on only (for ncu launch lists)
for name, frame in ((("c5_tr15.json", 300),) if QUICK else (("c5_tr15.json", 300), ("c5_tr15_like.json", 12))):
    d = C.POINTER(F.SceneDesc)()
    assert lib.trb_desc_load_json(os.path.join(SCENES, name).encode(), W, H, 2048, C.byref(d)) == 0, lib.trb_last_error()
    desc = d.contents
    step = desc.film.scene_time / desc.film.frames
    g, o = api.Scene(desc, 0), O.OracleScene(desc)
    g.update_frame(frame, frame * step, (frame + 1) * step); o.update_frame(frame, frame * step, (frame + 1) * step)
    kw = dict(block_start=12000, block_count=48, sample_first=0, sample_count=2, seed=1)
    os_, _ = o.render_samples(**kw)
    for table, split, occ, pipe, msort in (((2, -1, 4, 34, 1),) if QUICK else ((2, 0, 4, 34, 1), (2, 1, 4, 34, 1), (2, 1, 4, 34, 0), (1, 1, 4, 34, 1), (2, 1, 4, 0, 1), (0, 0, 4, 34, 1))):
        g.set_option("anim.table", table); g.set_option("shade.split", split); g.set_option("shade.anim_occupancy", occ); g.set_option("trace.pipe", pipe); g.set_option("shade.sort", msort)
        parity = g.render_samples(**kw)[0].tobytes() == os_.tobytes()
        g.render_device(film.data_ptr(), stats.data_ptr(), None, spp=2048, sample_first=0, sample_count=SPP_STEP, seed=1)
        torch.cuda.synchronize(); stats.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(1, 3):
            g.render_device(film.data_ptr(), stats.data_ptr(), None, spp=2048, sample_first=i * SPP_STEP, sample_count=SPP_STEP, seed=1)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        s = stats.cpu().numpy()
        print(json.dumps({"workload": "%s 1920x1080 frame %d (%d instances, %d meshes, %d MERL tables), 2 passes x %d spp" % (name, frame, desc.n_instances, desc.n_meshes, desc.n_merl, SPP_STEP),
                          "per_path_transform_table": table, "split_shade": split, "trace_variant": pipe, "material_buckets": bool(msort), "shade_ctas_per_sm": occ, "bit_exact_vs_oracle_on_48_blocks": bool(parity), "mrays_s": float(s[1:5].sum()) / ms / 1e3,
                          "msamples_s": float(s[0]) / ms / 1e3, "ms_per_pass": ms / 2,
                          "rays": {"primary": int(s[1]), "shadow": int(s[2]), "mis": int(s[3]), "continuation": int(s[4])}}), flush=True)
    g.close(); o.close()
    lib.trb_desc_free(d)
