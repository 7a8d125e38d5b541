"""Generates the committed golden vectors from the CPU oracle (detmath build):  python tests/golden/make_golden.py
The reference holds no vectors for this path and cannot be built here, so these pin the oracle against
regressions and give the GPU tests a fixed target that does not depend on running the oracle on the GPU box."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tray_rust_b200 import _ffi as F, api, scenebuild as SB  # noqa: E402
from oracle import pyoracle as O  # noqa: E402


def golden_scenes():
    """name -> (desc factory, render kwargs)"""
    def zoo():
        return SB.scene_materials_zoo(16, 16, 4, SB.synthetic_merl_table()).finish()

    def c4():
        return SB.scene_c4(5000, 32, 16, 4).finish()

    def cornell():
        lib = F.load_trb()
        d = C.POINTER(F.SceneDesc)()
        assert lib.trb_desc_load_json(os.path.join(HERE, "scenes", "c1_cornell_box.json").encode(), 24, 16, 4, C.byref(d)) == 0
        return d.contents  # leaked on purpose (tiny)

    def smallpt():
        lib = F.load_trb()
        d = C.POINTER(F.SceneDesc)()
        assert lib.trb_desc_load_json(os.path.join(HERE, "scenes", "c2_smallpt.json").encode(), 16, 16, 4, C.byref(d)) == 0
        return d.contents
    def anim():
        return SB.scene_animated(24, 16, 4, frames=4, scene_time=1.0).finish()

    def tr15_like():
        import make_scenes
        merl = os.path.join(HERE, "scenes", "merl", "synthetic.binary")
        if not os.path.exists(merl):
            make_scenes.write_synthetic_merl(merl)
        lib = F.load_trb()
        d = C.POINTER(F.SceneDesc)()
        assert lib.trb_desc_load_json(os.path.join(HERE, "scenes", "c5_tr15_like.json").encode(), 24, 16, 4, C.byref(d)) == 0
        return d.contents
    return {"zoo": (zoo, dict(seed=11)), "c4_5k": (c4, dict(seed=12)), "c1_cornell": (cornell, dict(seed=13)), "c2_smallpt": (smallpt, dict(seed=14)),
            "anim_f2": (anim, dict(seed=15)), "c5_tr15_like_f12": (tr15_like, dict(seed=16))}


# Scene::update_frame arguments (frame, start, end) the vectors were made with; (0, 0, 0) unless listed
FRAMES = {"anim_f2": (2, 0.5, 0.75), "c5_tr15_like_f12": (12, 6.0, 6.5)}


def frame_of(name):
    return FRAMES.get(name, (0, 0.0, 0.0))


if __name__ == "__main__":
    for name, (mk, kw) in golden_scenes().items():
        o = O.OracleScene(mk())
        o.update_frame(*frame_of(name))
        rays, xy = o.camera_rays(**kw)
        hits, hst = o.intersect(rays)
        samples, st = o.render_samples(**kw)
        film, _ = o.render(threads=1, flags=F.RENDER_NO_UPDATE, **kw)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), rays=rays, xy=xy, hits=hits, samples=samples, film=film,
                            counters=np.array([st.rays_primary, st.rays_shadow, st.rays_mis, st.rays_continuation, st.node_tests, st.tri_tests, st.inst_tests], np.uint64))
        print(name, len(rays), "samples; mean radiance", samples["r"].mean())
    lib = O.load_oracle("det")
    pts = np.zeros((3, 64, 2), np.float32)
    for k, scr in enumerate(((0, 0), (0x9E3779B9, 0x7F4A7C15), (0xFFFFFFFE, 1))):
        for i in range(64):
            lib.orc_sample_02(i, scr[0], scr[1], F.ptr(pts[k, i]))
    perm = np.array([[lib.orc_permute(i % l, l, 0xC0FFEE ^ l) for i in range(16)] for l in (16, 9, 13)], np.uint32)  # i must be < l
    rng = np.array([lib.orc_rng(1, p, s, d) for p in (0, 1, 77) for s in (0, 5, 0xFFFFFFFF) for d in (0, 3, 14, 33)], np.uint32)
    np.savez_compressed(os.path.join(HERE, "sampler.npz"), sample_02=pts, permute=perm, rng=rng)
    print("done")
