"""The reference's own unit tests for this path (the only vectors it holds), restated
against the oracle AND against the product's host code where a host-only entry point exists:
  src/linalg/matrix4.rs:265-305 (mul), src/linalg/transform.rs:284-379 (identity/translate/scale/rotate_*),
  src/linalg/mod.rs:129-142 (cross, dot), src/partition.rs:41-53 (partition)."""
import ctypes as C
import math

import numpy as np

from tray_rust_b200 import _ffi as F
from tray_rust_b200.scenebuild import quat_axis_angle


def kf(t=(0, 0, 0), q=(0, 0, 0, 1), s=(1, 1, 1)):
    k = F.Keyframe()
    k.translation[:] = t
    k.rotation[:] = q
    k.scaling[:] = s
    return k


def apply(oracle, k, v):
    out = np.zeros(9, np.float32)
    v = np.asarray(v, np.float32)
    oracle.orc_xf_apply(C.byref(k), F.ptr(v), F.ptr(out))
    return out[0:3], out[3:6], out[6:9]


def test_matrix_mul_literal(oracle):  # matrix4.rs:289-305
    a = np.array([1, 2, 1, 0, 3, 1, 4, 2, 1, 2, -5, 4, 3, 2, 4, 1], np.float32)
    b = np.array([8, 0, 2, 3, -2, 1, 0, 1, 5, -2, 3, 1, 0, 0, 4, 1], np.float32)
    c = np.array([9, 0, 5, 6, 42, -7, 26, 16, -21, 12, 3, 4, 40, -6, 22, 16], np.float32)
    out = np.zeros(16, np.float32)
    oracle.orc_m4_mul(F.ptr(a), F.ptr(b), F.ptr(out))
    assert np.array_equal(out, c)
    eye = np.eye(4, dtype=np.float32).reshape(-1)
    oracle.orc_m4_mul(F.ptr(eye), F.ptr(eye), F.ptr(out))
    assert np.array_equal(out, eye)


def test_matrix_inverse(oracle):
    rng = np.random.default_rng(0)
    for _ in range(20):
        m = rng.normal(size=(4, 4)).astype(np.float32)
        out = np.zeros(16, np.float32)
        oracle.orc_m4_inverse(F.ptr(m), F.ptr(out))
        assert np.allclose(out.reshape(4, 4) @ m, np.eye(4), atol=2e-3)


def test_mult_sanity(oracle):  # transform.rs:284-293
    p, v, n = apply(oracle, kf(), (1, 2, 3))
    assert np.array_equal(p, [1, 2, 3]) and np.array_equal(v, [1, 2, 3]) and np.array_equal(n, [1, 2, 3])


def test_translate(oracle):  # transform.rs:294-304
    p, _, _ = apply(oracle, kf(t=(1, 2, 3)), (1, 2, -1))
    assert np.array_equal(p, [2, 4, 2])
    _, v, n = apply(oracle, kf(t=(1, 2, 3)), (1, 0, 1))
    assert np.array_equal(v, [1, 0, 1]) and np.array_equal(n, [1, 0, 1])


def test_scale(oracle):  # transform.rs:305-315
    p, v, _ = apply(oracle, kf(s=(0.5, 0.1, 2.0)), (10, 20, 30))
    exp = np.array([10, 20, 30], np.float32) * np.array([0.5, 0.1, 2.0], np.float32)
    assert np.array_equal(p, exp) and np.array_equal(v, exp)
    _, _, n = apply(oracle, kf(s=(0.5, 0.1, 2.0)), (1, 2, 10))
    assert np.allclose(n, [2, 20, 5], rtol=1e-6)


def test_rotations(oracle):  # transform.rs:316-369 (tolerances as in the reference)
    p, v, n = apply(oracle, kf(q=quat_axis_angle((1, 0, 0), 90)), (0, 1, 0))
    for x in (p, v, n):
        assert abs(x[0]) < 1e-6 and abs(x[1]) < 1e-4 and abs(x[2] - 1) < 1e-6
    p, v, n = apply(oracle, kf(q=quat_axis_angle((0, 1, 0), -90)), (1, 0, 0))
    for x in (p, v, n):
        assert abs(x[0]) < 1e-4 and abs(x[1]) < 1e-6 and abs(x[2] - 1) < 1e-6
    p, v, n = apply(oracle, kf(q=quat_axis_angle((0, 0, 1), 90)), (1, 0, 0))
    for x in (p, v, n):
        assert abs(x[0]) < 1e-4 and abs(x[1] - 1) < 1e-6 and abs(x[2]) < 1e-6


def test_cross_dot(oracle):  # linalg/mod.rs:129-142
    out = np.zeros(4, np.float32)
    a, b = np.array([1, 0, 0], np.float32), np.array([0, 1, 0], np.float32)
    oracle.orc_cross_dot(F.ptr(a), F.ptr(b), F.ptr(out))
    assert np.array_equal(out[:3], [0, 0, 1])
    a, b = np.array([1, 2, 3], np.float32), np.array([4, 5, 6], np.float32)
    oracle.orc_cross_dot(F.ptr(a), F.ptr(b), F.ptr(out))
    assert out[3] == np.float32(1 * 4 + 2 * 5 + 3 * 6)


def test_partition(oracle):  # partition.rs:41-53
    v = np.array([1, 2, 3, 4, 5, 6], np.uint32)
    idx = oracle.orc_partition_even(F.ptr(v), len(v))
    assert idx == 3
    assert all(x % 2 == 0 for x in v[:3]) and all(x % 2 == 1 for x in v[3:])
    assert list(v) == [6, 2, 4, 3, 5, 1]  # the two-ended swap order of the reference's algorithm


def test_keyframe_transform_host_matches_oracle(oracle, trb):
    """Keyframe::transform (keyframe.rs:60-63): product host code vs oracle, bit for bit."""
    rng = np.random.default_rng(1)
    for _ in range(200):
        ax = rng.normal(size=3)
        k = kf(t=rng.normal(size=3) * 10, q=quat_axis_angle(ax, rng.uniform(-180, 180)), s=rng.uniform(0.1, 5, size=3))
        om, oi, hm, hi = (np.zeros(16, np.float32) for _ in range(4))
        oracle.orc_keyframe_transform(C.byref(k), F.ptr(om), F.ptr(oi))
        assert trb.trb_host_keyframe_transform(C.byref(k), F.ptr(hm), F.ptr(hi)) == F.TRB_OK
        assert np.array_equal(om.view(np.uint32), hm.view(np.uint32))
        assert np.array_equal(oi.view(np.uint32), hi.view(np.uint32))
        assert np.allclose(om.reshape(4, 4) @ oi.reshape(4, 4), np.eye(4), atol=1e-4)


def test_quaternion_matrix_is_rotation(oracle):
    k = kf(q=quat_axis_angle((1, 2, 3), 40))
    m, i = np.zeros(16, np.float32), np.zeros(16, np.float32)
    oracle.orc_keyframe_transform(C.byref(k), F.ptr(m), F.ptr(i))
    r = m.reshape(4, 4)[:3, :3]
    assert np.allclose(r @ r.T, np.eye(3), atol=1e-6) and abs(np.linalg.det(r) - 1) < 1e-6
    a = np.array([1, 2, 3]) / math.sqrt(14)
    assert np.allclose(r @ a, a, atol=1e-6)
