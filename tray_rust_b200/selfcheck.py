"""__graft_entry__.smoke(): one small invocation of the hot path on cuda:0, checked against the oracle."""
import numpy as np

from . import _ffi as F, api, scenebuild as SB


def smoke():
    desc = SB.scene_materials_zoo(32, 32, 4, SB.synthetic_merl_table()).finish()
    g = api.Scene(desc, 0)                      # raises if libtrb.so or the GPU is missing: no CPU fallback
    o = api.OracleScene(desc)                   # the checker (test infrastructure)
    g.update_frame(0, 0.0, 0.0); o.update_frame(0, 0.0, 0.0)
    gs, gst = g.render_samples(seed=5)
    os_, ost = o.render_samples(seed=5)
    assert gs.tobytes() == os_.tobytes(), "per-sample radiance differs from the oracle"
    gf, st = g.render(seed=5)
    of, _ = o.render(seed=5)
    ig = gf[..., :3] / np.maximum(gf[..., 3:], 1e-6); io = of[..., :3] / np.maximum(of[..., 3:], 1e-6)
    rmse = float(np.sqrt(np.mean((ig - io) ** 2)))
    assert rmse < 1e-5, rmse
    # the keyframed kernel variants (per-ray AnimatedTransform evaluation) on one frame of a small animated scene
    desc2 = SB.scene_animated(32, 32, 2, animated_fov=True).finish()
    g2, o2 = api.Scene(desc2, 0), api.OracleScene(desc2)
    g2.update_frame(1, 0.25, 0.5); o2.update_frame(1, 0.25, 0.5)
    assert g2.render_samples(seed=5)[0].tobytes() == o2.render_samples(seed=5)[0].tobytes(), "keyframed scene differs from the oracle"
    print("smoke ok: %d camera samples bit-exact vs oracle (+ %d of a keyframed scene), film rmse %.2e, %d rays, kernel %.2f ms"
          % (len(gs), 32 * 32 * 2, rmse, st.rays_total(), st.kernel_ms))
