"""N>1 path on CPU: world_size-2 gloo. Each rank renders its share of the Morton block list (the reference's own
sharding, master.rs:88-120) and the films are combined by one SUM reduce (image.rs:21-50). The renderer here is the
oracle because there is no GPU in this container; the sharding + reduce plumbing is the product's (tray_rust_b200.dist)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_blocks_matches_master_rs():
    from tray_rust_b200.dist import shard_blocks
    for n in (1, 7, 12, 32400):
        for w in (1, 2, 3, 8):
            parts = [shard_blocks(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            for (s0, c0), (s1, _) in zip(parts, parts[1:]):
                assert s0 + c0 == s1
            assert all(c == n // w for _, c in parts[:-1])          # blocks_per_worker, remainder to the last (master.rs:91-93,218-224)
    from tray_rust_b200.dist import shard_is_empty
    assert [shard_is_empty(shard_blocks(1, r, 2)) for r in range(2)] == [True, False]   # fewer blocks than ranks: idle ranks, never "all blocks"


def _worker(rank, world, port, out):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tray_rust_b200 import _ffi as F, api, scenebuild as SB
    from oracle import pyoracle as O
    from tray_rust_b200.dist import shard_blocks, reduce_film, max_over_ranks, sum_over_ranks
    o = O.OracleScene(SB.scene_smallpt_like(32, 24, 4).finish())
    o.update_frame(0, 0.0, 0.0)
    nb = o.n_blocks()
    start, count = shard_blocks(nb, rank, world)
    assert count > 0  # an empty shard must skip the render: block_count 0 means "all blocks" in the ABI (block_queue.rs:39-41)
    film, st = o.render(threads=1, flags=F.RENDER_NO_UPDATE, block_start=start, block_count=count, seed=3)
    t = torch.from_numpy(film)
    reduce_film(t, dst=0)
    rays = sum_over_ranks([st.rays_primary + st.rays_shadow + st.rays_mis + st.rays_continuation, st.camera_samples], "cpu")
    tmax = max_over_ranks(rank + 1.0, "cpu")
    if rank == 0:
        np.save(out, t.numpy())
        np.save(out + ".meta.npy", np.array(rays + [tmax]))
    dist.destroy_process_group()


def test_two_rank_tile_sharding_sums_to_the_full_frame(tmp_path):
    sys.path.insert(0, REPO)
    from tray_rust_b200 import _ffi as F, scenebuild as SB
    from oracle import pyoracle as O
    out = str(tmp_path / "film.npy")
    mp.spawn(_worker, args=(2, 29533, out), nprocs=2, join=True)
    got = np.load(out)
    o = O.OracleScene(SB.scene_smallpt_like(32, 24, 4).finish())
    o.update_frame(0, 0.0, 0.0)
    full, st = o.render(threads=1, flags=F.RENDER_NO_UPDATE, seed=3)
    assert np.allclose(got, full, rtol=1e-5, atol=1e-6)             # fp32 sum order differs between 1 and 2 ranks
    meta = np.load(out + ".meta.npy")
    assert meta[0] == st.rays_primary + st.rays_shadow + st.rays_mis + st.rays_continuation and meta[1] == st.camera_samples
    assert meta[2] == 2.0
