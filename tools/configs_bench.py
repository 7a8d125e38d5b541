"""Developer tool (GPU box): throughput of every BASELINE.json configuration on ONE GPU, scene resident, film on the device —
C1 the reference's cornell_box.json 400x400, C2 smallpt-shaped 512x512, C3 Cornell + 69 451-triangle mesh 800x600, C4 1 M triangles
1920x1080 (the bench workload), C5 the reference's tr15.json with stand-in assets 1920x1080 — with the split-shading material buckets on
and off where the scene uses the split kernels. One JSON line per configuration.
   gpurun -- 'python tools/configs_bench.py > gpurun_out/configs.log'"""
import ctypes as C
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
from tray_rust_b200 import _ffi as F, api, scenebuild as SB  # noqa: E402
import make_scenes  # noqa: E402
import make_tr15  # noqa: E402

SCENES = os.path.join(REPO, "tests", "golden", "scenes")
make_tr15.write_assets()
make_scenes.write_synthetic_merl(os.path.join(SCENES, "merl", "synthetic.binary"))
lib = F.load_trb()
dev = torch.device("cuda:0")


def from_json(name, w, h, spp):
    d = C.POINTER(F.SceneDesc)()
    assert lib.trb_desc_load_json(os.path.join(SCENES, name).encode(), w, h, spp, C.byref(d)) == 0, lib.trb_last_error()
    return d.contents


CONFIGS = [
    ("C1 scenes/cornell_box.json 400x400 (64 spp frame)", lambda: from_json("c1_cornell_box.json", 400, 400, 64), 400, 400, 64, 16, 0),
    ("C2 scenes/smallpt.json 512x512 (1024 spp frame)", lambda: from_json("c2_smallpt.json", 512, 512, 1024), 512, 512, 1024, 32, 0),
    ("C3 Cornell + 69 451-triangle mesh 800x600 (2048 spp frame)", lambda: SB.scene_c3(800, 600, 2048).finish(), 800, 600, 2048, 32, 0),
    ("C4 1 M random triangles in Cornell walls 1920x1080 (4096 spp frame)", lambda: SB.scene_c4(1_000_000, 1920, 1080, 4096).finish(), 1920, 1080, 4096, 8, 0),
    ("C5 scenes/tr15.json, stand-in assets, frame 300, 1920x1080 (2048 spp frame)", lambda: from_json("c5_tr15.json", 1920, 1080, 2048), 1920, 1080, 2048, 8, 300),
]
only = [a for a in sys.argv[1:] if not a.startswith("-")]
for name, mk, w, h, spp, spp_step, frame in CONFIGS:
    if only and not any(name.startswith(o) for o in only):
        continue
    if name.startswith("C1") and not os.path.exists(os.path.join(SCENES, "c1_cornell_box.json")):
        continue
    desc = mk()
    g = api.Scene(desc, 0)
    step = desc.film.scene_time / max(desc.film.frames, 1)
    g.update_frame(frame, frame * step, (frame + 1) * step)
    film = torch.zeros(h, w, 4, dtype=torch.float32, device=dev)
    stats = torch.zeros(10, dtype=torch.int64, device=dev)
    for buckets, kind in ((1, 1), (1, 0), (0, 0)):
        g.set_option("shade.sort", buckets); g.set_option("shade.kind", kind)
        if os.environ.get("SPLIT"): g.set_option("shade.split", int(os.environ["SPLIT"]))
        g.render_device(film.data_ptr(), stats.data_ptr(), None, spp=spp, sample_first=0, sample_count=spp_step, seed=1)
        torch.cuda.synchronize(); stats.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(1, 4):
            g.render_device(film.data_ptr(), stats.data_ptr(), None, spp=spp, sample_first=i * spp_step, sample_count=spp_step, seed=1)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        s = stats.cpu().numpy()
        print(json.dumps({"config": name, "material_buckets": bool(buckets), "matte_instantiation": bool(kind), "shade_split_option": os.environ.get("SPLIT", "per scene"), "mrays_s": float(s[1:5].sum()) / ms / 1e3, "msamples_s": float(s[0]) / ms / 1e3,
                          "ms_per_pass": ms / 3, "spp_per_pass": spp_step, "instances": desc.n_instances, "meshes": desc.n_meshes}), flush=True)
    g.close()
