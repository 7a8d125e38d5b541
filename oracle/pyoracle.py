"""Python bindings of the CPU oracle (oracle/_build/liboracle_*.so) — TEST INFRASTRUCTURE.

Only tests/, tools/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference`` legs may import
this module, and only as the checker (or the timed CPU arm), never as part of the product path: nothing under
``tray_rust_b200/`` imports it. ``OracleScene`` has the same methods as ``tray_rust_b200.api.Scene`` so parity tests
read the same on both sides; the ctypes struct mirrors of include/trb.h are shared with the product bindings.
"""
import ctypes as C
import os

import numpy as np

from tray_rust_b200 import _ffi as F
from tray_rust_b200.api import _Base, _cfg, TrbError
from tray_rust_b200._ffi import SceneDesc, Stats, RenderCfg, Material, Keyframe, u32, f32  # noqa: F401

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_oracle = {}


def oracle_path(kind="det"):
    return os.path.join(REPO, "oracle", "_build", "liboracle_%s.so" % kind)


def load_oracle(kind="det"):
    """Load the CPU oracle (test infrastructure). kind: 'det' (detmath) or 'sys' (glibc libm)."""
    if kind in _oracle:
        return _oracle[kind]
    p = oracle_path(kind)
    if not os.path.exists(p):
        raise RuntimeError("oracle library missing (%s): run oracle/build.sh" % p)
    lib = C.CDLL(p)
    vp, sz = C.c_void_p, C.c_size_t
    lib.orc_last_error.restype = C.c_char_p
    lib.orc_scene_create.argtypes = [C.POINTER(SceneDesc), C.POINTER(vp)]
    lib.orc_scene_destroy.argtypes = [vp]
    lib.orc_scene_destroy.restype = None
    lib.orc_scene_update_frame.argtypes = [vp, u32, f32, f32]
    lib.orc_set_baseline_mode.argtypes = [vp, C.c_int]
    lib.orc_set_baseline_mode.restype = None
    lib.orc_block_list.argtypes = [vp, u32, u32, C.POINTER(u32), vp, u32]
    lib.orc_scene_get_bvh.argtypes = [vp, C.c_int, C.POINTER(u32), vp, C.POINTER(u32), vp]
    lib.orc_scene_get_transform.argtypes = [vp, u32, vp, vp]
    lib.orc_scene_get_filter_table.argtypes = [vp, vp]
    lib.orc_intersect.argtypes = [vp, sz, vp, vp, C.POINTER(Stats)]
    lib.orc_render.argtypes = [vp, C.POINTER(RenderCfg), vp, C.POINTER(Stats), C.c_int]
    lib.orc_render_samples.argtypes = [vp, C.POINTER(RenderCfg), sz, vp, C.POINTER(Stats), C.c_int]
    lib.orc_camera_rays.argtypes = [vp, C.POINTER(RenderCfg), sz, vp, vp]
    lib.orc_film_to_srgb8.argtypes = [vp, vp, vp]
    lib.orc_detmath.argtypes = [C.c_int, sz, vp, vp, vp]
    lib.orc_detmath.restype = None
    lib.orc_rng.argtypes = [u32, u32, u32, u32]
    lib.orc_rng.restype = u32
    lib.orc_permute.argtypes = [u32, u32, u32]
    lib.orc_permute.restype = u32
    lib.orc_sample_02.argtypes = [u32, u32, u32, vp]
    lib.orc_sample_02.restype = None
    lib.orc_morton2.argtypes = [u32, u32]
    lib.orc_morton2.restype = u32
    lib.orc_bsdf_probe.argtypes = [C.POINTER(Material), vp, vp, vp, u32, vp, vp]
    lib.orc_m4_mul.argtypes = [vp, vp, vp]
    lib.orc_m4_mul.restype = None
    lib.orc_m4_inverse.argtypes = [vp, vp]
    lib.orc_m4_inverse.restype = None
    lib.orc_keyframe_transform.argtypes = [C.POINTER(Keyframe), vp, vp]
    lib.orc_keyframe_transform.restype = None
    lib.orc_partition_even.argtypes = [vp, sz]
    lib.orc_partition_even.restype = sz
    lib.orc_libm_kind.restype = C.c_int
    lib.orc_cross_dot.argtypes = [vp, vp, vp]
    lib.orc_cross_dot.restype = None
    lib.orc_xf_apply.argtypes = [C.POINTER(Keyframe), vp, vp]
    lib.orc_xf_apply.restype = None
    _oracle[kind] = lib
    return lib



class OracleScene(_Base):
    """CPU oracle with the same surface (TEST INFRASTRUCTURE ONLY)."""
    _pfx = "orc_"

    def __init__(self, desc, libm="det", baseline=False):
        self._lib = load_oracle(libm)
        self._desc = desc
        h = C.c_void_p()
        self._h = None
        self._check(self._lib.orc_scene_create(C.byref(desc), C.byref(h)))
        self._h = h
        self.width, self.height = desc.film.width, desc.film.height
        self.spp = 1 << (max(1, desc.film.samples) - 1).bit_length()
        if baseline:
            self._lib.orc_set_baseline_mode(h, 1)

    def _check(self, rc):
        if rc != F.TRB_OK:
            raise TrbError(rc, (self._lib.orc_last_error() or b"").decode())

    def close(self):
        if self._h is not None:
            self._lib.orc_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def render(self, film=None, threads=0, **kw):
        cfg = _cfg(**kw)
        if film is None:
            film = np.zeros((self.height, self.width, 4), np.float32)
        st = F.Stats()
        self._check(self._lib.orc_render(self._h, C.byref(cfg), F.ptr(film), C.byref(st), threads))
        return film, st

    def render_samples(self, threads=0, **kw):
        cfg = _cfg(**kw)
        n = self._n_samples(cfg)
        out = np.zeros(n, F.SAMPLE_DTYPE)
        st = F.Stats()
        self._check(self._lib.orc_render_samples(self._h, C.byref(cfg), n, F.ptr(out), C.byref(st), threads))
        return out, st

    def camera_rays(self, **kw):
        cfg = _cfg(**kw)
        n = self._n_samples(cfg)
        rays, xy = np.zeros(n, F.RAY_DTYPE), np.zeros((n, 2), np.float32)
        self._check(self._lib.orc_camera_rays(self._h, C.byref(cfg), n, F.ptr(rays), F.ptr(xy)))
        return rays, xy

    def intersect(self, rays):
        rays = np.ascontiguousarray(rays, dtype=F.RAY_DTYPE)
        hits = np.zeros(len(rays), F.HIT_DTYPE)
        st = F.Stats()
        self._check(self._lib.orc_intersect(self._h, len(rays), F.ptr(rays), F.ptr(hits), C.byref(st)))
        return hits, st

    def to_srgb8(self, film):
        film = np.ascontiguousarray(film, dtype=np.float32)
        out = np.zeros((self.height, self.width, 3), np.uint8)
        self._check(self._lib.orc_film_to_srgb8(self._h, F.ptr(film), F.ptr(out)))
        return out



def smoke():
    """One small invocation of the hot path on cuda:0, checked bit-for-bit against this oracle (__graft_entry__.smoke)."""
    from tray_rust_b200 import api as _api, scenebuild as _SB
    desc = _SB.scene_materials_zoo(32, 32, 4, _SB.synthetic_merl_table()).finish()
    g = _api.Scene(desc, 0)                      # raises if libtrb.so or the GPU is missing: no CPU fallback
    o = OracleScene(desc)                   # the checker (test infrastructure)
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
    desc2 = _SB.scene_animated(32, 32, 2, animated_fov=True).finish()
    g2, o2 = _api.Scene(desc2, 0), OracleScene(desc2)
    g2.update_frame(1, 0.25, 0.5); o2.update_frame(1, 0.25, 0.5)
    assert g2.render_samples(seed=5)[0].tobytes() == o2.render_samples(seed=5)[0].tobytes(), "keyframed scene differs from the oracle"
    print("smoke ok: %d camera samples bit-exact vs oracle (+ %d of a keyframed scene), film rmse %.2e, %d rays, kernel %.2f ms"
          % (len(gs), 32 * 32 * 2, rmse, st.rays_total(), st.kernel_ms))
