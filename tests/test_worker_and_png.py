"""the wire-compatible `--worker` replacement (tray_rust_b200/lib/trb_worker, plain C++ over the C ABI) driven
by a stand-in for the reference master that speaks its bincode protocol (/root/reference/src/exec/distrib/mod.rs:51-100,
master.rs:217-236, worker.rs:60-89), and the PNG writer that replaces image::save_buffer (main.rs:95-103)."""
import os
import socket
import struct
import subprocess
import time
import zlib

import numpy as np
import pytest

from tray_rust_b200 import _ffi as F

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENE = os.path.join(REPO, "tests", "golden", "scenes", "c1_cornell_box.json")
WORKER = os.path.join(REPO, "tray_rust_b200", "lib", "trb_worker")


def encode_instructions(scene, frames, block_start, block_count):
    """bincode(Infinite) of exec::distrib::Instructions: u64 size | String | (usize, usize) | usize | usize, little-endian."""
    s = scene.encode()
    body = struct.pack("<Q", len(s)) + s + struct.pack("<QQQQ", frames[0], frames[1], block_start, block_count)
    return struct.pack("<Q", 8 + len(body)) + body


def decode_frame(buf):
    size, frame, bw, bh, nb = struct.unpack_from("<QQQQQ", buf, 0)
    o = 40
    blocks = np.frombuffer(buf, "<u8", 2 * nb, o).reshape(-1, 2); o += 16 * nb
    (npx,) = struct.unpack_from("<Q", buf, o); o += 8
    pixels = np.frombuffer(buf, "<f4", npx, o); o += 4 * npx
    assert size == len(buf) == o
    return frame, (bw, bh), blocks, pixels


def start_worker(port, *extra):
    p = subprocess.Popen([WORKER, "--worker", "--port", str(port), "--seed", "5", "--spp", "4", *extra], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert "listening for master" in p.stdout.readline()
    return p


def recv_exact(sock, n):
    out = b""
    while len(out) < n:
        chunk = sock.recv(n - len(out))
        if not chunk:
            break
        out += chunk
    return out


def test_png_writer_roundtrip(tmp_path, trb):
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, size=(37, 211, 3), dtype=np.uint8)      # > 65535 bytes of scanlines: several stored blocks
    img2 = rng.integers(0, 256, size=(301, 199, 3), dtype=np.uint8)
    for k, a in enumerate((img, img2)):
        p = str(tmp_path / ("t%d.png" % k))
        assert trb.trb_write_png(p.encode(), F.ptr(np.ascontiguousarray(a)), a.shape[1], a.shape[0]) == F.TRB_OK
        raw = open(p, "rb").read()
        assert raw[:8] == b"\x89PNG\r\n\x1a\n"
        o, chunks = 8, []
        while o < len(raw):
            (n,) = struct.unpack(">I", raw[o:o + 4]); ty = raw[o + 4:o + 8]; data = raw[o + 8:o + 8 + n]
            (crc,) = struct.unpack(">I", raw[o + 8 + n:o + 12 + n])
            assert crc == zlib.crc32(ty + data) & 0xffffffff
            chunks.append((ty, data)); o += 12 + n
        assert [c[0] for c in chunks] == [b"IHDR", b"IDAT", b"IEND"]
        w, h, depth, ctype, comp, flt, inter = struct.unpack(">IIBBBBB", chunks[0][1])
        assert (w, h, depth, ctype, comp, flt, inter) == (a.shape[1], a.shape[0], 8, 2, 0, 0, 0)
        scan = np.frombuffer(zlib.decompress(chunks[1][1]), np.uint8).reshape(h, 1 + 3 * w)
        assert (scan[:, 0] == 0).all() and np.array_equal(scan[:, 1:].reshape(h, w, 3), a)
    assert trb.trb_write_png(str(tmp_path / "no" / "dir.png").encode(), F.ptr(img), 211, 37) == F.TRB_IO


def test_worker_parses_the_masters_instructions_and_refuses_to_render_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by test_worker_renders_the_masters_block_range")
    port = 40000 + os.getpid() % 20000
    p = start_worker(port)
    with socket.create_connection(("127.0.0.1", port), timeout=10) as s:
        s.sendall(encode_instructions(SCENE, (0, 0), 100, 50))
        out, err = p.communicate(timeout=60)
    assert 'scene: "%s", frames: (0, 0), block_start: 100, block_count: 50' % SCENE in out
    assert p.returncode == 3 and "no CPU fallback" in err       # TRB_NO_DEVICE: never a host render


@pytest.mark.gpu
def test_worker_renders_the_masters_block_range():
    """The stand-in master sends Instructions for blocks [100, 150) of frame 0 and reads one Frame back: the 2x2 lock blocks whose
    weights are all non-zero (render_target.rs:215-241), bit-compatible with the film of the same trb_render call."""
    from tray_rust_b200 import api, exec as X
    port = 40000 + os.getpid() % 20000
    p = start_worker(port)
    with socket.create_connection(("127.0.0.1", port), timeout=30) as s:
        s.sendall(encode_instructions(SCENE, (0, 0), 100, 50))
        s.settimeout(120)
        head = recv_exact(s, 8)
        (size,) = struct.unpack("<Q", head)
        buf = head + recv_exact(s, size - 8)
    out, err = p.communicate(timeout=60)
    assert p.returncode == 0, err
    frame, bsize, blocks, pixels = decode_frame(buf)
    assert frame == 0 and bsize == (2, 2) and len(pixels) == 16 * len(blocks) and len(blocks) > 0
    scene, rt, spp, fi = X.Scene.load_file(SCENE, 0, 0, 0, 4)
    film, _ = scene.gpu.render(spp=4, seed=5, block_start=100, block_count=50)
    h, w = film.shape[:2]
    wt = film[..., 3].reshape(h // 2, 2, w // 2, 2)
    full = (wt != 0).all(axis=(1, 3))
    ys, xs = np.nonzero(full)
    assert np.array_equal(blocks, np.stack([2 * xs, 2 * ys], axis=1).astype(np.uint64))          # row-major over the lock-block grid
    want = np.stack([film[2 * ys + dy, 2 * xs + dx] for dy in (0, 1) for dx in (0, 1)], axis=1)    # (n, 4 pixels, RGBW)
    assert np.allclose(pixels.reshape(-1, 4, 4), want, rtol=2e-4, atol=2e-5)
