"""Geometry known answers for the oracle: analytic shapes, triangle test, BVH traversal against a float64
brute force, SAH build invariants, and the product's host BVH builder against the oracle's."""
import ctypes as C

import numpy as np

from tray_rust_b200 import _ffi as F, api, scenebuild as SB
from oracle import pyoracle as O


def mk_rays(o, d, tmin=0.0, tmax=np.inf):
    o = np.atleast_2d(np.asarray(o, np.float32)); d = np.atleast_2d(np.asarray(d, np.float32))
    r = np.zeros(len(o), F.RAY_DTYPE)
    r["o"], r["d"], r["min_t"], r["max_t"] = o, d, tmin, tmax
    return r


def simple_scene(add):
    b = SB.SceneBuilder(8, 8, 1)
    m = b.add_material(F.MAT_MATTE, (0.5, 0.5, 0.5), roughness=1.0)
    add(b, m)
    b.area_light(F.SHAPE_RECT, m, [SB.trs(t=(0, 1000, 0))], (1, 1, 1), p0=1, p1=1)
    b.add_camera([SB.trs(t=(0, 0, -5))])
    o = O.OracleScene(b.finish())
    o.update_frame(0, 0.0, 0.0)
    return o


def test_sphere_known_answers():  # sphere.rs:33-81 through a translate+scale instance
    o = simple_scene(lambda b, m: b.receiver(F.SHAPE_SPHERE, m, [SB.trs(t=(0, 0, 10), s=2)], p0=1.0))
    h, _ = o.intersect(mk_rays([[0, 0, 0], [0, 0, 10], [0, 5, 0], [0, 0, 0]], [[0, 0, 1], [0, 0, 1], [0, 0, 1], [0, 0, 2]]))
    assert h["inst"].tolist() == [0, 0, F.MISS, 0]
    assert h["t"][0] == 8.0   # front of the radius-2 sphere
    assert h["t"][1] == 2.0   # from the centre: second root
    assert np.isinf(h["t"][2])
    assert h["t"][3] == 4.0   # direction of length 2: t is shared between spaces (receiver.rs:29-43)


def test_rect_and_disk_known_answers():  # rectangle.rs:38-64, disk.rs:42-76
    def add(b, m):
        b.receiver(F.SHAPE_RECT, m, [SB.trs(t=(0, 0, 5))], p0=2.0, p1=4.0)
        b.receiver(F.SHAPE_DISK, m, [SB.trs(t=(10, 0, 5))], p0=2.0, p1=1.0)
    o = simple_scene(add)
    rays = mk_rays([[0, 0, 0], [0.5, 2, 0], [1.01, 0, 0], [0, 0, 0], [10, 1.5, 0], [10, 0.5, 0], [10, 2.5, 0], [1, 0.5, 0]],
                   [[0, 0, 1]] * 3 + [[1, 0, 0]] + [[0, 0, 1]] * 4)
    h, _ = o.intersect(rays)
    # inclusive y edge hits; outside; parallel ray; annulus hit / hole / outside; and the reference's asymmetric NaN
    # slab (bbox.rs:77-78,103): a ray exactly on the x face with d.x == 0 has tmax = 0*inf = NaN in the x
    # slab, which is never overwritten, so the instance box is rejected although the rectangle test is inclusive.
    assert h["inst"].tolist() == [0, 0, F.MISS, F.MISS, 1, F.MISS, F.MISS, F.MISS]
    assert h["t"][0] == 5.0 and h["t"][4] == 5.0


def test_box_cull_is_strict_even_though_shapes_are_inclusive():
    """Q10: BBox::fast_intersect ends with `tmin < max_t && tmax > min_t` (bbox.rs:103), so a hit exactly at
    min_t or max_t is culled by the (zero-thickness) instance box before the inclusive shape test
    (rectangle.rs:46 `t < min_t || t > max_t`) can accept it."""
    o = simple_scene(lambda b, m: b.receiver(F.SHAPE_RECT, m, [SB.trs(t=(0, 0, 5))], p0=2.0, p1=2.0))
    h, _ = o.intersect(mk_rays([[0, 0, 0]] * 4, [[0, 0, 1]] * 4, tmin=[5.0, 0.0, 5.0001, 4.999], tmax=[9.0, 5.0, 9.0, 5.001]))
    assert h["inst"].tolist() == [F.MISS, F.MISS, F.MISS, 0]


def brute_force(pos, idx, o, d, tmin, tmax):
    """float64 Moeller-Trumbore over all triangles; returns (tri, t)."""
    pa, pb, pc = (pos[idx[:, k]].astype(np.float64) for k in range(3))
    e0, e1 = pb - pa, pc - pa
    best_t, best = tmax, -1
    s0 = np.cross(d, e1)
    dd = np.einsum("ij,ij->i", s0, e0)
    ok = dd != 0
    div = np.where(ok, 1.0 / np.where(ok, dd, 1), 0)
    dv = o - pa
    b1 = np.einsum("ij,ij->i", dv, s0) * div
    s1 = np.cross(dv, e0)
    b2 = np.einsum("ij,j->i", s1, d) * div
    t = np.einsum("ij,ij->i", e1, s1) * div
    hit = ok & (b1 >= 0) & (b1 <= 1) & (b2 >= 0) & (b1 + b2 <= 1) & (t >= tmin) & (t <= tmax)
    if hit.any():
        k = np.argmin(np.where(hit, t, np.inf))
        best, best_t = int(k), float(t[k])
    return best, best_t


def test_bvh_traversal_matches_brute_force():
    pos, nrm, uv, idx = SB.random_triangle_mesh(3000, 11, lo=(-2, -2, 4), hi=(2, 2, 8), jitter=0.4)
    def add(b, m):
        me = b.add_mesh(pos, nrm, uv, idx)
        b.receiver(F.SHAPE_MESH, m, [SB.trs()], mesh=me)
    o = simple_scene(add)
    rng = np.random.default_rng(4)
    n = 300
    org = rng.uniform(-1, 1, size=(n, 3)).astype(np.float32) * [2, 2, 0]
    d = rng.normal(size=(n, 3)).astype(np.float32) * [0.3, 0.3, 0] + [0, 0, 1]
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    h, st = o.intersect(mk_rays(org, d))
    assert st.tri_tests < 3000 * n / 5  # the BVH actually culls
    n_hit = 0
    for i in range(n):
        tri, t = brute_force(pos, idx, org[i].astype(np.float64), d[i].astype(np.float64), 0.0, np.inf)
        if tri < 0:
            assert h["inst"][i] == F.MISS or h["t"][i] > 1e3
            continue
        n_hit += 1
        assert h["inst"][i] == 0
        assert abs(h["t"][i] - t) <= 1e-4 * max(1.0, t)
        if h["prim"][i] != tri:  # only near-ties may pick another triangle
            tri2, t2 = brute_force(pos, idx[[h["prim"][i]]], org[i].astype(np.float64), d[i].astype(np.float64), 0.0, np.inf)
            assert abs(t2 - t) < 1e-4
    assert n_hit > 50


def check_bvh_invariants(nodes, order, boxes, max_geom):
    n = len(boxes)
    assert sorted(order.tolist()) == list(range(n))  # every primitive in exactly one leaf
    seen = np.zeros(len(nodes), bool)

    def walk(i):
        seen[i] = True
        nd = nodes[i]
        if nd["b"] & F.BVH_LEAF:
            cnt = int(nd["b"] & ~np.uint32(F.BVH_LEAF))
            prims = order[nd["a"]:nd["a"] + cnt]
            assert 1 <= cnt <= max(max_geom, 1)
            lo = boxes[prims][:, :3].min(0); hi = boxes[prims][:, 3:].max(0)
            assert np.array_equal(lo, nd["bmin"]) and np.array_equal(hi, nd["bmax"])
            return lo, hi, cnt
        assert nd["b"] in (0, 1, 2)
        l = walk(i + 1)               # first child follows its parent (bvh.rs:248-267)
        r = walk(int(nd["a"]))
        lo, hi = np.minimum(l[0], r[0]), np.maximum(l[1], r[1])
        assert np.array_equal(lo, nd["bmin"]) and np.array_equal(hi, nd["bmax"])
        return lo, hi, l[2] + r[2]
    import sys
    sys.setrecursionlimit(10000)
    assert walk(0)[2] == n and seen.all()


def test_host_bvh_builder_matches_oracle_and_invariants(trb):
    """BVH::build (bvh.rs:139-267): the product's host builder and the oracle's must emit identical arrays."""
    for n, seed in ((1, 1), (2, 2), (4, 3), (5, 4), (16, 5), (17, 6), (300, 7), (20000, 8)):
        pos, nrm, uv, idx = SB.random_triangle_mesh(n, seed)
        if n == 300:  # coincident centroids exercise bvh.rs:156-166
            pos[: 3 * 40] = np.tile(pos[:3], (40, 1))
        tri = pos[idx.reshape(-1)].reshape(-1, 3, 3)
        boxes = np.concatenate([tri.min(1), tri.max(1)], axis=1).astype(np.float32)
        nn = F.u32()
        assert trb.trb_host_build_bvh(F.ptr(boxes), n, 16, C.byref(nn), None, None) == F.TRB_OK
        nodes = np.zeros(nn.value, F.NODE_DTYPE); order = np.zeros(n, np.uint32)
        trb.trb_host_build_bvh(F.ptr(boxes), n, 16, C.byref(nn), F.ptr(nodes), F.ptr(order))
        b = SB.SceneBuilder(8, 8, 1)
        m = b.add_material(F.MAT_MATTE, (0.5, 0.5, 0.5), roughness=1.0)
        b.receiver(F.SHAPE_MESH, m, [SB.trs()], mesh=b.add_mesh(pos, nrm, uv, idx))
        b.area_light(F.SHAPE_RECT, m, [SB.trs(t=(0, 99, 0))], (1, 1, 1), p0=1, p1=1)
        b.add_camera([SB.trs()])
        on, oo = O.OracleScene(b.finish()).bvh(0)
        assert nodes.tobytes() == on.tobytes() and np.array_equal(order, oo), n
        check_bvh_invariants(on, oo, boxes, 16)


def test_host_bvh_rejects_empty(trb):
    nn = F.u32()
    assert trb.trb_host_build_bvh(None, 0, 4, C.byref(nn), None, None) == F.TRB_INVALID_ARG  # bvh.rs:35 assert!


def test_point_lights_are_never_hit():  # emitter.rs:119-120
    b = SB.SceneBuilder(8, 8, 1)
    b.point_light([SB.trs(t=(0, 0, 5))], (1, 1, 1, 10))
    b.add_camera([SB.trs()])
    o = O.OracleScene(b.finish())
    o.update_frame(0, 0.0, 0.0)
    h, st = o.intersect(mk_rays([[0, 0, 0]], [[0, 0, 1]]))
    assert h["inst"][0] == F.MISS
