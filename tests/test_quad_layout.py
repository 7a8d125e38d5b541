"""The two-level DQuad node records (csrc/trb_device.h) walk exactly the leaves a literal BVH::intersect (bvh.rs:81-130)
walks, in the same order, with the same max_t history. Host-only self-check of the layout + visit-order logic that the
trace kernel runs (quad_visit is shared host/device code); the GPU parity tests then compare real hits bit for bit."""
import ctypes as C

import numpy as np
import pytest

from tray_rust_b200 import _ffi as F, scenebuild as SB


def build(trb, boxes, max_geom):
    boxes = np.ascontiguousarray(boxes, np.float32)
    n = len(boxes)
    nn = F.u32()
    assert trb.trb_host_build_bvh(F.ptr(boxes), n, max_geom, C.byref(nn), None, None) == F.TRB_OK
    nodes, order = np.zeros(nn.value, F.NODE_DTYPE), np.zeros(n, np.uint32)
    assert trb.trb_host_build_bvh(F.ptr(boxes), n, max_geom, C.byref(nn), F.ptr(nodes), F.ptr(order)) == F.TRB_OK
    return nodes


def tri_boxes(n, seed, jitter):
    pos, _, _, idx = SB.random_triangle_mesh(n, seed, jitter=jitter)
    v = pos[idx.reshape(-1, 3)]
    return np.concatenate([v.min(axis=1), v.max(axis=1)], axis=1)


def rays_for(boxes, n, seed, axis_aligned=0):
    rng = np.random.default_rng(seed)
    lo, hi = boxes[:, :3].min(axis=0), boxes[:, 3:].max(axis=0)
    rays = np.zeros(n, F.RAY_DTYPE)
    rays["o"] = rng.uniform(lo - 2, hi + 2, size=(n, 3)).astype(np.float32)
    tgt = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    d = tgt - rays["o"]
    rays["d"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    if axis_aligned:  # tiny but non-zero components: finite 1/d of magnitude up to 1e30
        k = rng.integers(0, 3, size=axis_aligned)
        rays["d"][np.arange(axis_aligned), k] = rng.choice([1e-30, -1e-30, 1e-12, -3e-20], size=axis_aligned).astype(np.float32)
    rays["min_t"] = rng.choice([0.0, 0.001, 5.0], size=n).astype(np.float32)
    rays["max_t"] = rng.choice([np.inf, 30.0, 0.999], size=n).astype(np.float32)
    return rays


@pytest.mark.parametrize("n,max_geom,jitter,seed", [(1, 16, 0.2, 1), (2, 16, 0.2, 2), (3, 1, 0.2, 3), (17, 16, 0.5, 4), (40, 4, 0.3, 5),
                                                     (5000, 16, 0.15, 6), (5000, 1, 2.0, 7), (60000, 16, 1.0, 8)])
def test_quad_walk_equals_reference_walk(trb, n, max_geom, jitter, seed):
    boxes = tri_boxes(n, seed, jitter)
    nodes = build(trb, boxes, max_geom)
    rays = rays_for(boxes, 4000 if n < 5000 else 20000, seed + 100, axis_aligned=200)
    bad, nl, nq = F.u32(), C.c_uint64(), C.c_uint64()
    assert trb.trb_host_quad_check(F.ptr(nodes), len(nodes), F.ptr(rays), len(rays), C.byref(bad), C.byref(nl), C.byref(nq)) == F.TRB_OK
    assert bad.value == 0
    if n >= 5000:
        assert nl.value > 3000 and nq.value > 100000   # the walks really went through the tree


def test_quad_walk_on_instance_boxes(trb):
    """TLAS-shaped input: a handful of large overlapping boxes, max_geom 4 (scene.rs:138)."""
    rng = np.random.default_rng(9)
    c = rng.uniform(-10, 10, size=(23, 3)); h = rng.uniform(0.1, 8, size=(23, 3))
    boxes = np.concatenate([c - h, c + h], axis=1).astype(np.float32)
    nodes = build(trb, boxes, 4)
    rays = rays_for(boxes, 20000, 10)
    bad = F.u32()
    assert trb.trb_host_quad_check(F.ptr(nodes), len(nodes), F.ptr(rays), len(rays), C.byref(bad), None, None) == F.TRB_OK
    assert bad.value == 0
