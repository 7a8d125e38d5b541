"""Developer loop on a GPU box: parity of every stage against the oracle + quick timings.
Run: gpurun -- 'python tools/gpu_check.py [--big]'   (writes gpurun_out/gpu_check.json)"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tray_rust_b200 import _ffi as F, api, scenebuild as SB  # noqa: E402
from oracle import pyoracle as O  # noqa: E402

RES = {}


def report(name, ok, extra=""):
    RES[name] = bool(ok)
    print(("PASS " if ok else "FAIL ") + name + (" " + str(extra) if extra != "" else ""), flush=True)


def compare_scene(tag, desc, spp_pass=4, n_rand=20000, film=True):
    g = api.Scene(desc)
    o = O.OracleScene(desc)
    g.update_frame(0, 0.0, 0.0)
    o.update_frame(0, 0.0, 0.0)
    gn, go = g.bvh(-1)
    on, oo = o.bvh(-1)
    report(tag + ":tlas", gn.tobytes() == on.tobytes() and np.array_equal(go, oo))
    for i in range(desc.n_instances):
        gm, gi = g.transform(i)
        om, oi = o.transform(i)
        if not (np.array_equal(api.bits(gm), api.bits(om)) and np.array_equal(api.bits(gi), api.bits(oi))):
            report(tag + ":xf%d" % i, False)
            break
    else:
        report(tag + ":transforms", True)
    report(tag + ":filter_table", np.array_equal(api.bits(g.filter_table()), api.bits(o.filter_table())))
    report(tag + ":blocks", np.array_equal(g.block_list(), o.block_list()))
    kw = dict(sample_first=0, sample_count=spp_pass, seed=7)
    gr, gxy = g.camera_rays(**kw)
    orr, oxy = o.camera_rays(**kw)
    report(tag + ":camera_rays", gr.tobytes() == orr.tobytes() and gxy.tobytes() == oxy.tobytes(),
           "n=%d" % len(gr))
    gh, gst = g.intersect(orr)
    oh, ost = o.intersect(orr)
    same = gh.tobytes() == oh.tobytes()
    report(tag + ":intersect_primary", same, "hit%%=%.1f" % (100.0 * np.mean(oh["inst"] != F.MISS)))
    if not same:
        bad = np.nonzero((gh["inst"] != oh["inst"]) | (gh["prim"] != oh["prim"]) | (api.bits(gh["t"]) != api.bits(oh["t"])))[0]
        print("   mismatches", len(bad), "first", bad[:5], gh[bad[:3]], oh[bad[:3]])
    report(tag + ":intersect_counters", (gst.node_tests, gst.tri_tests, gst.inst_tests) == (ost.node_tests, ost.tri_tests, ost.inst_tests),
           (gst.node_tests, ost.node_tests, gst.tri_tests, ost.tri_tests, gst.inst_tests, ost.inst_tests))
    # random secondary-like rays from hit points
    rng = np.random.default_rng(3)
    hit = np.nonzero(oh["inst"] != F.MISS)[0]
    if len(hit):
        sel = rng.choice(hit, size=min(n_rand, len(hit)), replace=False)
        rays = np.zeros(len(sel), F.RAY_DTYPE)
        p = orr["o"][sel] + orr["d"][sel] * oh["t"][sel, None]
        d = rng.normal(size=(len(sel), 3)).astype(np.float32)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        rays["o"] = p; rays["d"] = d; rays["min_t"] = 0.001; rays["max_t"] = np.inf
        gh2, gs2 = g.intersect(rays)
        oh2, os2 = o.intersect(rays)
        report(tag + ":intersect_random", gh2.tobytes() == oh2.tobytes(), "n=%d hit%%=%.1f" % (len(sel), 100.0 * np.mean(oh2["inst"] != F.MISS)))
        report(tag + ":intersect_random_counters", (gs2.node_tests, gs2.tri_tests) == (os2.node_tests, os2.tri_tests))
    # per-sample radiance, bit-exact (reference shadow mode so counters are comparable)
    t0 = time.time()
    osamp, ost = o.render_samples(flags=0, **kw)
    t_or = time.time() - t0
    gsamp, gst = g.render_samples(flags=F.RENDER_STATS | F.RENDER_REFERENCE_SHADOW, **kw)
    same = gsamp.tobytes() == osamp.tobytes()
    report(tag + ":samples_bitexact", same, "n=%d oracle %.2fs gpu %.1fms" % (len(gsamp), t_or, gst.kernel_ms))
    if not same:
        neq = np.nonzero((api.bits(gsamp["r"]) != api.bits(osamp["r"])) | (api.bits(gsamp["g"]) != api.bits(osamp["g"])) | (api.bits(gsamp["b"]) != api.bits(osamp["b"])))[0]
        pos_bad = np.sum((api.bits(gsamp["x"]) != api.bits(osamp["x"])) | (api.bits(gsamp["y"]) != api.bits(osamp["y"])))
        print("   radiance mismatches %d / %d, pos mismatches %d" % (len(neq), len(gsamp), pos_bad))
        for i in neq[:5]:
            print("   ", i, gsamp[i], osamp[i])
        nan_g = np.isnan(gsamp["r"]).sum(); nan_o = np.isnan(osamp["r"]).sum()
        print("   nan gpu/oracle", nan_g, nan_o, "max abs diff", np.nanmax(np.abs(gsamp["r"] - osamp["r"])))
    gd, od = gst.as_dict(), ost.as_dict()
    keys = ["camera_samples", "rays_primary", "rays_shadow", "rays_mis", "rays_continuation", "node_tests", "tri_tests", "inst_tests"]
    report(tag + ":render_counters", all(gd[k] == od[k] for k in keys), {k: (gd[k], od[k]) for k in keys})
    # default (any-hit shadow) mode must give the same radiance
    gsamp2, _ = g.render_samples(flags=0, **kw)
    report(tag + ":anyhit_same_radiance", gsamp2.tobytes() == gsamp.tobytes())
    # the megakernel execution shape must agree bit for bit, counters included
    gsamp3, gst3 = g.render_samples(flags=F.RENDER_MEGAKERNEL | F.RENDER_STATS | F.RENDER_REFERENCE_SHADOW, **kw)
    report(tag + ":megakernel_same", gsamp3.tobytes() == gsamp.tobytes() and all(getattr(gst3, k) == getattr(gst, k) for k in keys))
    if film:
        gf, gs = g.render(flags=0, **kw)
        of, _ = o.render(flags=0, threads=0, **kw)
        wdiff = np.abs(gf[..., 3] - of[..., 3]).max()
        denom = np.maximum(np.abs(of), 1e-3)
        rel = np.abs(gf - of) / denom
        img_g = gf[..., :3] / np.maximum(gf[..., 3:], 1e-6)
        img_o = of[..., :3] / np.maximum(of[..., 3:], 1e-6)
        rmse = float(np.sqrt(np.mean((img_g - img_o) ** 2)))
        report(tag + ":film", rmse < 1e-5 and np.isfinite(gf).all(), "rmse=%.3g maxrel=%.3g wdiff=%.3g kernel_ms=%.2f" % (rmse, rel.max(), wdiff, gs.kernel_ms))
        s8g, s8o = g.to_srgb8(of), o.to_srgb8(of)
        report(tag + ":srgb8", np.array_equal(s8g, s8o), "maxdiff=%d" % int(np.abs(s8g.astype(int) - s8o.astype(int)).max()))
    g.close(); o.close()


FLAGS = 0


def timing(n_tris):
    import torch
    b = SB.scene_c4(n_tris, 1920, 1080, 4096)
    t0 = time.time()
    g = api.Scene(b.finish())
    print("scene_create(%d tris): %.2fs" % (n_tris, time.time() - t0), flush=True)
    g.update_frame(0, 0.0, 0.0)
    film = torch.zeros((1080, 1920, 4), dtype=torch.float32, device="cuda")
    stats = torch.zeros(10, dtype=torch.int64, device="cuda")
    out = {}
    for spp_pass in (1, 4, 8):
        for it in range(3):
            stats.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            g.render_device(film.data_ptr(), stats.data_ptr(), None, sample_first=it * spp_pass, sample_count=spp_pass, seed=1, flags=FLAGS)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            st = stats.cpu().numpy()
            rays = int(st[1:5].sum())
            print("render pass spp=%d: %.1f ms, %.1f Msamples/s, %.1f Mrays/s (p/s/m/c = %s)" % (spp_pass, ms, st[0] / ms / 1e3, rays / ms / 1e3, st[1:5]), flush=True)
            out["render_spp%d" % spp_pass] = dict(ms=ms, msamples=st[0] / ms / 1e3, mrays=rays / ms / 1e3)
    # primary-ray intersect microbench
    rays, _ = g.camera_rays(sample_first=0, sample_count=1, seed=1)
    d_rays = torch.from_numpy(rays.view(np.float32).reshape(-1, 8)).cuda()
    d_hits = torch.zeros((len(rays), 4), dtype=torch.int32, device="cuda")
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.intersect_device(len(rays), d_rays.data_ptr(), d_hits.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("intersect primary %d rays: %.2f ms = %.1f Mrays/s" % (len(rays), ms, len(rays) / ms / 1e3), flush=True)
        out["intersect_primary_mrays"] = len(rays) / ms / 1e3
    RES["timing_%d" % n_tris] = out
    g.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    a = ap.parse_args()
    import traceback
    if not a.no_parity:
        for tag, mk, sp in [("zoo", lambda: SB.scene_materials_zoo(64, 64, 16, SB.synthetic_merl_table()), 8),
                            ("smallpt", lambda: SB.scene_smallpt_like(64, 64, 16), 4),
                            ("c4_20k", lambda: SB.scene_c4(20000, 128, 72, 8), 2),
                            ("c3", lambda: SB.scene_c3(96, 72, 8, subdiv=4), 2)]:
            try:
                compare_scene(tag, mk().finish(), spp_pass=sp)
            except Exception:
                traceback.print_exc()
                report(tag + ":exception", False)
    try:
        timing(100000)
        if a.big:
            timing(1000000)
    except Exception:
        traceback.print_exc()
        report("timing:exception", False)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(RES, open("gpurun_out/gpu_check.json", "w"), indent=1, default=str)
    fails = [k for k, v in RES.items() if v is False]
    print("FAILED:", fails)
    sys.exit(1 if fails else 0)
