"""Python host side above the C ABI.

``Scene`` wraps a ``trb_scene`` (GPU, product). The CPU oracle's wrapper with the same methods
(``oracle.pyoracle.OracleScene``) is test infrastructure and lives outside this package.

The reference-shaped mirror (``Config``, ``RenderTarget``, ``Exec.render`` with the argument
meaning of /root/reference/src/exec/mod.rs:17-49) lives in ``tray_rust_b200.exec``.
"""
import ctypes as C

import numpy as np

from . import _ffi as F


class TrbError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("trb status %d: %s" % (status, msg))
        self.status = status


def _cfg(spp=0, sample_first=0, sample_count=0, block_start=0, block_count=0, current_frame=0, seed=1, flags=0,
         shard_index=0, shard_count=0, shard_chunk=0):
    return F.RenderCfg(spp, sample_first, sample_count, block_start, block_count, current_frame, seed, flags,
                       shard_index, shard_count, shard_chunk)


class _Base:
    """Shared helpers; subclasses provide self._lib, self._h, self._pfx and self._check."""

    def _n_samples(self, cfg):
        nb = self.n_blocks(cfg.block_start, cfg.block_count)
        if cfg.shard_count > 1:
            ch = max(1, cfg.shard_chunk)
            nb = sum(1 for j in range(nb) if (j // ch) % cfg.shard_count == cfg.shard_index)
        spp = self.spp if cfg.spp == 0 else 1 << (max(1, cfg.spp) - 1).bit_length()
        cnt = cfg.sample_count if cfg.sample_count else spp - cfg.sample_first
        return nb * 64 * cnt

    def n_blocks(self, start=0, count=0):
        n = F.u32()
        self._check(getattr(self._lib, self._pfx + "block_list")(self._h, start, count, C.byref(n), None, 0))
        return n.value

    def block_list(self, start=0, count=0):
        n = self.n_blocks(start, count)
        xy = np.zeros((n, 2), np.uint32)
        m = F.u32()
        self._check(getattr(self._lib, self._pfx + "block_list")(self._h, start, count, C.byref(m), F.ptr(xy), n))
        return xy

    def bvh(self, which=-1):
        nn, no = F.u32(), F.u32()
        f = getattr(self._lib, self._pfx + "scene_get_bvh")
        self._check(f(self._h, which, C.byref(nn), None, C.byref(no), None))
        nodes = np.zeros(nn.value, F.NODE_DTYPE)
        order = np.zeros(no.value, np.uint32)
        self._check(f(self._h, which, C.byref(nn), F.ptr(nodes), C.byref(no), F.ptr(order)))
        return nodes, order

    def transform(self, inst):
        m, i = np.zeros(16, np.float32), np.zeros(16, np.float32)
        self._check(getattr(self._lib, self._pfx + "scene_get_transform")(self._h, inst, F.ptr(m), F.ptr(i)))
        return m.reshape(4, 4), i.reshape(4, 4)

    def filter_table(self):
        t = np.zeros(256, np.float32)
        self._check(getattr(self._lib, self._pfx + "scene_get_filter_table")(self._h, F.ptr(t)))
        return t.reshape(16, 16)

    def update_frame(self, frame=0, start=0.0, end=0.0):
        self._check(getattr(self._lib, self._pfx + "scene_update_frame")(self._h, frame, start, end))


class Scene(_Base):
    """A scene resident on one B200 (product path)."""
    _pfx = "trb_"

    def __init__(self, desc, device=0):
        self._lib = F.load_trb()
        self._desc = desc  # keep arrays alive
        h = C.c_void_p()
        self._h = None
        self._check(self._lib.trb_scene_create(C.byref(desc), device, C.byref(h)))
        self._h = h
        self.device = device
        w, hh, spp, nb, ni, nl = (F.u32() for _ in range(6))
        self._check(self._lib.trb_scene_info(h, *(C.byref(x) for x in (w, hh, spp, nb, ni, nl))))
        self.width, self.height, self.spp, self.total_blocks, self.n_instances, self.n_lights = (x.value for x in (w, hh, spp, nb, ni, nl))

    def _check(self, rc):
        if rc != F.TRB_OK:
            raise TrbError(rc, (self._lib.trb_last_error() or b"").decode())

    def close(self):
        if self._h is not None:
            self._lib.trb_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def render(self, film=None, **kw):
        """trb_render: host film buffer, accumulated into. Returns (film, Stats)."""
        cfg = _cfg(**kw)
        if film is None:
            film = np.zeros((self.height, self.width, 4), np.float32)
        st = F.Stats()
        self._check(self._lib.trb_render(self._h, C.byref(cfg), F.ptr(film), C.byref(st)))
        return film, st

    def render_device(self, d_film_ptr, d_stats_ptr=None, stream=None, **kw):
        cfg = _cfg(**kw)
        self._check(self._lib.trb_render_device(self._h, C.byref(cfg), d_film_ptr, d_stats_ptr, stream))

    def set_option(self, name, value):
        """trb_scene_set_option: launch-shape options (never change results)."""
        self._check(self._lib.trb_scene_set_option(self._h, name.encode(), int(value)))

    def check_error(self):
        """trb_scene_check_error: drain the device, raise on a latched traversal-stack overflow."""
        self._check(self._lib.trb_scene_check_error(self._h))

    def trace_time(self):
        """(total ms, launches) of the trace kernel for launches made with RENDER_TIME_TRACE since the last call."""
        ms, n = F.f32(), F.u32()
        self._check(self._lib.trb_scene_trace_time(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def render_samples(self, **kw):
        cfg = _cfg(**kw)
        n = self._n_samples(cfg)
        out = np.zeros(n, F.SAMPLE_DTYPE)
        st = F.Stats()
        self._check(self._lib.trb_render_samples(self._h, C.byref(cfg), n, F.ptr(out), C.byref(st)))
        return out, st

    def camera_rays(self, **kw):
        cfg = _cfg(**kw)
        n = self._n_samples(cfg)
        rays, xy = np.zeros(n, F.RAY_DTYPE), np.zeros((n, 2), np.float32)
        self._check(self._lib.trb_camera_rays(self._h, C.byref(cfg), n, F.ptr(rays), F.ptr(xy)))
        return rays, xy

    def intersect(self, rays):
        rays = np.ascontiguousarray(rays, dtype=F.RAY_DTYPE)
        hits = np.zeros(len(rays), F.HIT_DTYPE)
        st = F.Stats()
        self._check(self._lib.trb_intersect(self._h, len(rays), F.ptr(rays), F.ptr(hits), C.byref(st)))
        return hits, st

    def intersect_device(self, n, d_rays, d_hits, d_stats=None, stream=None):
        self._check(self._lib.trb_intersect_device(self._h, n, d_rays, d_hits, d_stats, stream))

    def to_srgb8(self, film):
        film = np.ascontiguousarray(film, dtype=np.float32)
        out = np.zeros((self.height, self.width, 3), np.uint8)
        self._check(self._lib.trb_film_to_srgb8(self._h, F.ptr(film), F.ptr(out)))
        return out


class Comm:
    """One rank of a multi-GPU job (one process per GPU): NCCL communicator owned by libtrb (trb_comm_*). The 128-byte
    unique id is created by rank 0 with ``Comm.unique_id()`` and shipped to the other ranks by any transport."""

    def __init__(self, uid, n_ranks, rank, device):
        self._lib = F.load_trb()
        h = C.c_void_p()
        self._h = None
        buf = (C.c_char * 128).from_buffer_copy(bytes(uid))
        rc = self._lib.trb_comm_create(buf, n_ranks, rank, device, C.byref(h))
        if rc != F.TRB_OK:
            raise TrbError(rc, (self._lib.trb_last_error() or b"").decode())
        self._h, self.n_ranks, self.rank, self.device = h, n_ranks, rank, device

    @staticmethod
    def unique_id():
        lib = F.load_trb()
        buf = (C.c_char * 128)()
        rc = lib.trb_nccl_unique_id(buf)
        if rc != F.TRB_OK:
            raise TrbError(rc, (lib.trb_last_error() or b"").decode())
        return bytes(buf)

    def _check(self, rc):
        if rc != F.TRB_OK:
            raise TrbError(rc, (self._lib.trb_last_error() or b"").decode())

    def reduce_film(self, d_film_ptr, n_floats, root=0, stream=None):
        """SUM-reduce a device film to `root` (in place), enqueued on `stream` (trb_comm_reduce_film: ncclReduce)."""
        self._check(self._lib.trb_comm_reduce_film(self._h, d_film_ptr, n_floats, root, stream))

    def render_sharded(self, scene, film=None, root=0, **kw):
        """trb_render_sharded: this rank's tile shard at full spp, ONE film reduce, root adds into its host film."""
        cfg = _cfg(**kw)
        if film is None and self.rank == root:
            film = np.zeros((scene.height, scene.width, 4), np.float32)
        st = F.Stats()
        self._check(self._lib.trb_render_sharded(scene._h, self._h, C.byref(cfg), root, F.ptr(film) if film is not None else None, C.byref(st)))
        return film, st

    def close(self):
        if self._h is not None:
            self._lib.trb_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Group:
    """One process driving several GPUs (trb_group_*): a scene replica per device, tile-sharded render, one film reduce."""

    def __init__(self, desc, devices):
        self._lib = F.load_trb()
        self._desc = desc
        devs = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        self._h = None
        rc = self._lib.trb_group_create(C.byref(desc), devs, len(devices), C.byref(h))
        if rc != F.TRB_OK:
            raise TrbError(rc, (self._lib.trb_last_error() or b"").decode())
        self._h, self.devices = h, list(devices)
        self.width, self.height = desc.film.width, desc.film.height

    def render(self, film=None, **kw):
        cfg = _cfg(**kw)
        if film is None:
            film = np.zeros((self.height, self.width, 4), np.float32)
        st = F.Stats()
        rc = self._lib.trb_group_render(self._h, C.byref(cfg), F.ptr(film), C.byref(st))
        if rc != F.TRB_OK:
            raise TrbError(rc, (self._lib.trb_last_error() or b"").decode())
        return film, st

    def close(self):
        if self._h is not None:
            self._lib.trb_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def bits(a):
    """uint32 view of a float32 array (bit-exact comparisons)."""
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
