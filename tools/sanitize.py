"""Small scenes under compute-sanitizer (memcheck / racecheck / initcheck): wavefront and megakernel, film and samples."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tray_rust_b200 import _ffi as F, api, scenebuild as SB
for name, b in (("zoo", SB.scene_materials_zoo(32, 32, 4, SB.synthetic_merl_table())), ("c4_5k", SB.scene_c4(5000, 64, 32, 4))):
    g = api.Scene(b.finish())
    g.update_frame(0, 0.0, 0.0)
    film, st = g.render(seed=3)
    s, _ = g.render_samples(seed=3, flags=F.RENDER_STATS)
    s2, _ = g.render_samples(seed=3, flags=F.RENDER_MEGAKERNEL)
    f2, _ = g.render(seed=3, flags=F.RENDER_MEGAKERNEL)
    rays, _ = g.camera_rays(seed=3)
    h, _ = g.intersect(rays)
    print(name, "ok", st.rays_total(), s.tobytes() == s2.tobytes(), float(np.abs(film - f2).max()), g.to_srgb8(film).mean())
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
os.environ["TRB_TRACE_QUADS"] = "1"
g = api.Scene(SB.scene_c4(5000, 64, 32, 4).finish())
g.update_frame(0, 0.0, 0.0)
a, _ = g.render_samples(seed=3)
os.environ["TRB_TRACE_QUADS"] = "0"
b, _ = g.render_samples(seed=3)
print("quads ok", a.tobytes() == b.tobytes())
g.close()
