"""Small scenes under compute-sanitizer (memcheck / racecheck / initcheck): wavefront and megakernel, film and samples."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tray_rust_b200 import _ffi as F, api, scenebuild as SB
for name, b in (("zoo", SB.scene_materials_zoo(32, 32, 4, SB.synthetic_merl_table())), ("c4_5k", SB.scene_c4(5000, 64, 32, 4))):
    g = api.Scene(b.finish())
    g.update_frame(0, 0.0, 0.0)
    film, st = g.render(seed=3)                          # default film kernel (k_wf_film_v2: per-warp private tiles, lockstep RMW)
    g.set_option("film.v2", 0)
    film_a, _ = g.render(seed=3)                         # shared-memory-atomics film kernel
    g.set_option("film.v2", 1)
    g.set_option("pass.paths", 4096)                     # several internal passes per call
    film_p, _ = g.render(seed=3)
    g.set_option("pass.paths", 1 << 24)
    assert np.allclose(film, film_a, rtol=1e-4, atol=1e-5) and np.allclose(film, film_p, rtol=1e-4, atol=1e-5)
    s, _ = g.render_samples(seed=3, flags=F.RENDER_STATS)
    s2, _ = g.render_samples(seed=3, flags=F.RENDER_MEGAKERNEL)
    f2, _ = g.render(seed=3, flags=F.RENDER_MEGAKERNEL)
    rays, _ = g.camera_rays(seed=3)
    h, _ = g.intersect(rays)
    print(name, "ok", st.rays_total(), s.tobytes() == s2.tobytes(), float(np.abs(film - f2).max()), g.to_srgb8(film).mean())
    g.close()
# a filter whose reach exceeds filter_pixel_width (the 2x2 lock-block sample rule) through both film kernels
bw = SB.scene_smallpt_like(32, 32, 4)
bw.film.update(filter_type=F.FILTER_GAUSSIAN, filter_w=3.0, filter_h=2.5, filter_b=0.5, filter_c=0.0)
g = api.Scene(bw.finish())
fa, _ = g.render(seed=3)
g.set_option("film.v2", 0)
fb, _ = g.render(seed=3)
print("wide filter ok", bool(np.allclose(fa, fb, rtol=1e-4, atol=1e-5)))
g.close()
# keyframed kernels (ANIM variants) and the optional DQuad records
g = api.Scene(SB.scene_animated(32, 32, 4).finish())
for fr in (0, 2):
    g.update_frame(fr, fr * 0.25, (fr + 1) * 0.25)
    film, st = g.render(seed=3, flags=F.RENDER_NO_UPDATE)
    s, _ = g.render_samples(seed=3, flags=F.RENDER_STATS)
    s2, _ = g.render_samples(seed=3, flags=F.RENDER_MEGAKERNEL)
    print("animated frame", fr, "ok", st.rays_total(), s.tobytes() == s2.tobytes())
g.close()
g = api.Scene(SB.scene_c4(5000, 64, 32, 4).finish())
g.update_frame(0, 0.0, 0.0)
g.set_option("trace.quads", 1)
a, _ = g.render_samples(seed=3)
g.set_option("trace.quads", 0)
b, _ = g.render_samples(seed=3)
print("quads ok", a.tobytes() == b.tobytes())
g.close()
