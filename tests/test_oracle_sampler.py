"""Sampler / RNG / detmath known answers for the oracle (src/sampler/{ld,morton,block_queue}.rs and the
DESIGN.md determinism contract)."""
import numpy as np

from tray_rust_b200 import _ffi as F, api, scenebuild as SB
from oracle import pyoracle as O


def s02(oracle, n, s0=0, s1=0):
    out = np.zeros(2, np.float32)
    oracle.orc_sample_02(n, s0, s1, F.ptr(out))
    return out


def test_radical_inverses(oracle):
    # van der Corput base 2 and the Sobol' second dimension, unscrambled (ld.rs:95-119; PBRT 7.4.3)
    assert [s02(oracle, i)[0] for i in range(5)] == [0.0, 0.5, 0.25, 0.75, 0.125]
    assert [s02(oracle, i)[1] for i in range(5)] == [0.0, 0.5, 0.75, 0.25, 0.625]  # direction numbers 0x80.., 0xC0.., 0xA0.. (hand-run of ld.rs:109-119)


def test_sample_cap_is_one_minus_epsilon(oracle):  # Q21
    v = s02(oracle, 0xFFFFFFFF, 0, 0)
    assert v[0] <= np.float32(1.0) - np.float32(1.1920929e-7)


def test_02_sequence_stratification(oracle):
    """(0,2)-sequence: the first 2^k points put exactly one point in every elementary interval, for any scramble."""
    for scr in ((0, 0), (0xDEADBEEF, 0x12345678)):
        for k in (4, 6):
            n = 1 << k
            pts = np.array([s02(oracle, i, *scr) for i in range(n)])
            for a in range(k + 1):
                nx, ny = 1 << a, 1 << (k - a)
                cells = (np.floor(pts[:, 0] * nx).astype(int) * ny + np.floor(pts[:, 1] * ny).astype(int))
                assert len(set(cells.tolist())) == n


def test_permute_is_bijection(oracle):
    for l in list(range(1, 40)) + [64, 100, 256, 1000, 4096]:
        for p in (0, 1, 0xABCDEF01, 0xFFFFFFFF):
            img = sorted(oracle.orc_permute(i, l, p) for i in range(l))
            assert img == list(range(l)), (l, p)


def test_rng_is_a_pure_function_and_mixes(oracle):
    a = oracle.orc_rng(1, 2, 3, 4)
    assert a == oracle.orc_rng(1, 2, 3, 4)
    vals = np.array([oracle.orc_rng(7, p, 0, d) for p in range(64) for d in range(16)], dtype=np.uint64)
    assert len(set(vals.tolist())) == len(vals)
    u = vals.astype(np.float64) / 2 ** 32
    assert abs(u.mean() - 0.5) < 0.03 and 0.07 < u.var() < 0.1


def test_morton(oracle):  # morton.rs
    assert oracle.orc_morton2(0, 0) == 0 and oracle.orc_morton2(1, 0) == 1 and oracle.orc_morton2(0, 1) == 2
    assert oracle.orc_morton2(3, 5) == 0b100111
    assert oracle.orc_morton2(0xFFFF, 0xFFFF) == 0xFFFFFFFF


def test_block_queue_order_and_selection(oracle):  # block_queue.rs:28-46
    b = SB.SceneBuilder(32, 24, 1)
    m = b.add_material(F.MAT_MATTE, (0.5, 0.5, 0.5), roughness=1.0)
    b.area_light(F.SHAPE_RECT, m, [SB.trs()], (1, 1, 1), p0=1, p1=1)
    b.add_camera([SB.trs(t=(0, 0, -5))])
    o = O.OracleScene(b.finish())
    bl = o.block_list()
    assert len(bl) == 4 * 3
    codes = [oracle.orc_morton2(int(x), int(y)) for x, y in bl]
    assert codes == sorted(codes)
    assert bl[:4].tolist() == [[0, 0], [1, 0], [0, 1], [1, 1]]
    sel = o.block_list(5, 3)
    assert np.array_equal(sel, bl[5:8])
    assert len(o.block_list(10, 100)) == 2  # skip/take past the end


def _ulps(a, b):
    a = np.asarray(a, np.float32); b = np.asarray(b, np.float64)
    spacing = np.spacing(np.maximum(np.abs(b), 1e-30).astype(np.float32)).astype(np.float64)
    return np.abs(a.astype(np.float64) - b) / spacing


def test_detmath_accuracy(oracle):
    rng = np.random.default_rng(2)

    def run(op, a, b=None):
        a = np.ascontiguousarray(a, np.float32)
        b = np.ascontiguousarray(a if b is None else b, np.float32)
        out = np.zeros_like(a)
        oracle.orc_detmath(op, len(a), F.ptr(a), F.ptr(b), F.ptr(out))
        return out
    x = rng.uniform(-7, 7, 20000).astype(np.float32)
    assert np.abs(run(0, x) - np.sin(x.astype(np.float64))).max() < 3e-7
    assert np.abs(run(1, x) - np.cos(x.astype(np.float64))).max() < 3e-7
    c = rng.uniform(-1, 1, 20000).astype(np.float32)
    assert np.abs(run(2, c) - np.arccos(c.astype(np.float64))).max() < 6e-7
    y, xx = rng.normal(size=20000).astype(np.float32), rng.normal(size=20000).astype(np.float32)
    assert np.abs(run(3, y, xx) - np.arctan2(y.astype(np.float64), xx.astype(np.float64))).max() < 8e-7
    e = rng.uniform(-80, 10, 20000).astype(np.float32)
    assert _ulps(run(4, e), np.exp(e.astype(np.float64))).max() < 4
    l = np.exp(rng.uniform(-20, 20, 20000)).astype(np.float32)
    assert np.abs(run(5, l) - np.log(l.astype(np.float64))).max() < 3e-6
    p = rng.uniform(0.0031308, 1.0, 20000).astype(np.float32)
    g = np.full_like(p, 1 / 2.4)
    assert _ulps(run(6, p, g), p.astype(np.float64) ** (1 / 2.4)).max() < 8
    # special values
    assert run(0, [0.0])[0] == 0 and run(1, [0.0])[0] == 1 and run(2, [1.0])[0] == 0 and run(4, [0.0])[0] == 1 and run(5, [1.0])[0] == 0
    assert np.isnan(run(0, [np.inf])[0]) and run(4, [-200.0])[0] == 0 and np.isinf(run(4, [100.0])[0])
    assert run(3, [0.0], [0.0])[0] == 0


def test_camera_rays_cover_the_image(oracle):
    b = SB.scene_smallpt_like(16, 16, 4)
    o = O.OracleScene(b.finish())
    o.update_frame(0, 0.0, 0.0)
    rays, xy = o.camera_rays(seed=3)
    assert len(rays) == 16 * 16 * 4
    assert xy.min() >= 0 and xy[:, 0].max() <= 16 and xy[:, 1].max() <= 16
    # every pixel receives exactly spp samples inside [x, x+1] x [y, y+1]
    px = np.minimum(np.floor(xy).astype(int), 15)
    counts = np.zeros((16, 16), int)
    np.add.at(counts, (px[:, 1], px[:, 0]), 1)
    assert (counts == 4).all()
    assert np.allclose(np.linalg.norm(rays["d"], axis=1), 1.0, atol=1e-6)
    assert np.allclose(rays["o"], [0, 12, -60])
    # the centre of the image looks down +z
    c = np.argmin(np.abs(xy - 8).sum(1))
    assert rays["d"][c][2] > 0.99
