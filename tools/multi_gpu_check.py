"""Multi-GPU paths of the library on a real box (N >= 2 GPUs):
   torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py   # one process per GPU
   python tools/multi_gpu_check.py --group 2                                                                       # one process, 2 GPUs
Checks trb_render_sharded (interleaved and reference-style contiguous sharding, ONE ncclReduce per frame issued by libtrb) and
trb_group_render against the single-GPU render of the same frame: same ray counts, film equal up to float addition order."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tray_rust_b200 import _ffi as F, api, scenebuild as SB  # noqa: E402

W, H, SPP = 640, 360, 4


def desc():
    return SB.scene_c4(100_000, W, H, 64).finish()


def img(f):
    return f[..., :3] / np.maximum(f[..., 3:], 1e-6)


def group_mode(n):
    g1 = api.Scene(desc(), 0)
    ref, st1 = g1.render(spp=SPP, seed=3)
    grp = api.Group(desc(), list(range(n)))
    film, st = grp.render(spp=SPP, seed=3)
    film2, _ = grp.render(spp=SPP, seed=3)            # a second frame through the same communicators
    ok = st.rays_total() == st1.rays_total() and st.camera_samples == st1.camera_samples and np.allclose(film, ref, rtol=2e-4, atol=2e-5) and np.allclose(film2, film, rtol=2e-4, atol=2e-5)
    print(json.dumps({"mode": "trb_group_render", "devices": n, "rays": st.rays_total(), "rays_single": st1.rays_total(),
                      "rmse": float(np.sqrt(np.mean((img(film) - img(ref)) ** 2))), "kernel_ms": st.kernel_ms, "kernel_ms_single": st1.kernel_ms, "ok": bool(ok)}))
    return ok


def rank_mode():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")                      # plumbing only (ships the unique id, sums the counters)
    ids = [api.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, 0)
    comm = api.Comm(ids[0], world, rank, local)
    g = api.Scene(desc(), local)
    ok = True
    for name, kw in (("interleaved", {}), ("contiguous (master.rs:91-93)", dict(shard_count=0xffffffff))):
        film, st = comm.render_sharded(g, None, 0, spp=SPP, seed=3, **kw)
        t = torch.tensor([st.rays_total(), st.camera_samples], dtype=torch.int64)
        dist.all_reduce(t)
        if rank == 0:
            ref, st1 = g.render(spp=SPP, seed=3)
            good = int(t[0]) == st1.rays_total() and int(t[1]) == st1.camera_samples and np.allclose(film, ref, rtol=2e-4, atol=2e-5)
            print(json.dumps({"mode": "trb_render_sharded " + name, "ranks": world, "rays": int(t[0]), "rays_single": st1.rays_total(),
                              "rmse": float(np.sqrt(np.mean((img(film) - img(ref)) ** 2))), "ok": bool(good)}))
            ok = ok and good
    dist.barrier()
    comm.close()
    dist.destroy_process_group()
    return ok


if __name__ == "__main__":
    if "--group" in sys.argv:
        sys.exit(0 if group_mode(int(sys.argv[sys.argv.index("--group") + 1])) else 1)
    sys.exit(0 if rank_mode() else 1)
