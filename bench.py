#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json): Mrays/s over all logical ray types and frame ms,
bun69k 4K (3840x2160) 16 spp, on N B200s, beside the CPU restatement of the reference on the host cores.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one full frame of the workload through the wavefront kernels.
  value    : whole-job throughput with the scene resident in HBM and the frame left on the device (tile gather included for N>1)
  e2e      : same metric through the C ABI with HOST buffers: scene re-upload (H2D) + render + frame read-back (D2H) every step
  roofline : dominant kernel (k_trace): algorithmic bytes touched per launch / CUDA-event duration vs the measured HBM copy peak
  cpu_baseline : the oracle ("port" of the reference; the Java reference cannot run here) on a bounded tile of the same workload
N>1: one process per GPU (torchrun), frame split in interleaved 8-row chunks (strong scaling: the frame is fixed), NCCL all-gather of chunks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1] (the configuration the metric is quoted on): bun69k in the p3_t09 wrapper
    "bun69k": dict(scene="p3_t09.cli", cols=3840, rows=2160, spp=16, photons=-1, metric="Mrays/s (all ray types), bun69k 4K 16spp",
                   desc="configs[1]: bun69k in the p3_t09 wrapper; mesh = bun500 subdivided 3x, 61824 tris, stand-in for the missing bun69k.cli"),
    "t01": dict(scene="t01.cli", cols=3840, rows=2160, spp=16, photons=-1, metric="Mrays/s (all ray types), t01 4K 16spp", desc="configs[0]: two spheres, one point light"),
    "sierp": dict(scene="p3_t11_sierp.cli", cols=3840, rows=2160, spp=16, photons=-1, metric="Mrays/s (all ray types), p3_t11_sierp 4K 16spp",
                  desc="configs[2]: 21845 instances of the bun69k stand-in through the reference-topology instance BVH"),
    "planets": dict(scene="plnts3ColsBunnies.cli", cols=3840, rows=2160, spp=16, photons=-1, metric="Mrays/s (all ray types), plnts3ColsBunnies 4K 16spp",
                    desc="configs[3]: textured planets, bunny BVHs, reflection/refraction, spot + point lights"),
    "box_caustics": dict(scene="box_caustics.cli", cols=3840, rows=2160, spp=16, photons=4000000, metric="Mrays/s (all ray types incl. photon segments), box caustics 4K 16spp",
                         desc="configs[4]: Cornell wrapper around data/box.cli, caustic photon map k=80 r=0.05; photon pass inside every step"),
    "t11": dict(scene="t11.cli", cols=3840, rows=2160, spp=16, photons=1000000, metric="Mrays/s (all ray types incl. photon segments), t11 Cornell box 4K 16spp",
                desc="the reference scene behind its shipped box-caustics renders (t11c.png, boxCaustics*.png): Cornell box, mirror + glass spheres, point + disk light, diffuse photon map k=200 r=0.1"),
    "box_gi": dict(scene="box_gi.cli", cols=3840, rows=2160, spp=16, photons=1000000, metric="Mrays/s (all ray types incl. photon segments), box diffuse-GI 4K 16spp",
                   desc="configs[4] diffuse variant: diffuse photon map k=200 r=0.1; photon pass inside every step"),
}
WORKLOAD = dict(WORKLOADS["bun69k"])
METRIC = WORKLOAD["metric"]
CPU_TILE = (1728, 972, 1728 + 384, 972 + 216)      # bounded CPU sample: centre 384x216 tile of the 4K frame at 16 spp


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu, enabled=True):
        self.gpu, self.rows, self.stop, self.t, self.enabled = gpu, [], False, None, enabled

    def _run(self):
        while not self.stop:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        if self.enabled:
            self.t = threading.Thread(target=self._run, daemon=True); self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.t is not None:
            self.t.join(timeout=6)

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.rows)}


def cpu_leg(threads, steps=1, warmup=0):
    """Times the oracle (CPU restatement of the reference) on the bounded tile of the benchmark workload."""
    from oracle import orc
    orc.build()
    w = WORKLOAD
    o = orc.OracleScene(w["scene"], cols=w["cols"], rows=w["rows"], spp=w["spp"], photons=min(w["photons"], 200000) if w["photons"] >= 0 else -1)
    best = None
    for i in range(warmup + steps):
        r = o.render(rect=CPU_TILE, threads=threads, want=("argb",))
        rays = sum(r["stats"][k] for k in ("primary", "shadow", "reflect", "refract"))
        if i >= warmup and (best is None or r["seconds"] < best[0]):
            best = (r["seconds"], rays)
    return best[1] / best[0] / 1e6, best[0], best[1]


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # each step = the bounded tile, all host threads
    from oracle import orc
    orc.build()
    w = WORKLOAD
    o = orc.OracleScene(w["scene"], cols=w["cols"], rows=w["rows"], spp=w["spp"], photons=min(w["photons"], 200000) if w["photons"] >= 0 else -1)
    secs, rays = [], 0
    for i in range(args.warmup + args.steps):
        r = o.render(rect=CPU_TILE, threads=threads, want=("argb",))
        if i >= args.warmup:
            secs.append(r["seconds"]); rays = sum(r["stats"][k] for k in ("primary", "shadow", "reflect", "refract"))
    ms = 1e3 * sum(secs) / len(secs)
    value = rays / (ms / 1e3) / 1e6
    sample = "oracle (C++ restatement of the Java reference; no JDK on the box), %d threads, tile x[%d,%d) y[%d,%d) of the 4K frame at 16 spp = %d rays per step" % (threads, CPU_TILE[0], CPU_TILE[2], CPU_TILE[1], CPU_TILE[3], rays)
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": "%s %dx%d %dspp (bounded sample: centre 384x216 tile)" % (w["scene"], w["cols"], w["rows"], w["spp"])},
                      "cpu_baseline": {"value": round(value, 4), "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample},
                      "e2e": {"value": round(value, 4), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--accel", type=int, default=int(os.environ.get("DRT_ACCEL", "1")), help="0 reference-topology literal order, 1 reference-topology near-first (bit-identical results, default), 2 GPU LBVH")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--workload", default="bun69k", choices=sorted(WORKLOADS) + ["synth"])
    ap.add_argument("--synth", default="soup:65536", help="with --workload synth: soup:N (N random triangles) or grid:K (K^3 bunny instances), see tools/make_synth.py")
    ap.add_argument("--photons", type=int, default=None, help="photons cast per light for the photon workloads (sweep 1M..64M)")
    ap.add_argument("--res", default=None, help="COLSxROWS override")
    ap.add_argument("--spp", type=int, default=None)
    args = ap.parse_args()
    global METRIC
    if args.workload == "synth":
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import make_synth
        kind, n = args.synth.split(":")
        path = getattr(make_synth, kind)(int(n)) if int(os.environ.get("RANK", "0")) == 0 or not os.path.exists(os.path.join(make_synth.GEN, "%s_%s.cli" % (kind, n))) else os.path.join(make_synth.GEN, "%s_%s.cli" % (kind, n))
        WORKLOADS["synth"] = dict(scene="gen/" + os.path.basename(path), cols=3840, rows=2160, spp=16, photons=-1, metric="Mrays/s (all ray types), synthetic %s 4K 16spp" % args.synth,
                                  desc="SURVEY 8(d) synthetic scaling scene %s under the config-2 camera and lights" % args.synth)
    WORKLOAD.clear(); WORKLOAD.update(WORKLOADS[args.workload]); METRIC = WORKLOAD["metric"]
    if args.photons is not None and WORKLOAD["photons"] >= 0:
        WORKLOAD["photons"] = args.photons
    if args.res:
        WORKLOAD["cols"], WORKLOAD["rows"] = (int(x) for x in args.res.lower().split("x"))
    if args.spp:
        WORKLOAD["spp"] = args.spp
    if args.impl == "reference":
        return reference_arm(args)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import distraytracer_old_b200 as drt
    from distraytracer_old_b200 import dist as D

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = WORKLOAD
    ctx = drt.Context(device=local, cols=w["cols"], rows=w["rows"])
    scene = drt.Scene.from_cli(ctx, w["scene"], spp=w["spp"], photons=w["photons"], accel=args.accel)
    has_photons = scene.info()["photon_kind"] != 0
    npix = w["cols"] * w["rows"]
    frame = torch.zeros(D.padded_pixels(w["rows"], w["cols"], world), dtype=torch.int32, device="cuda")   # padded to whole chunks per rank: the gather packs by view
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        if has_photons:      # the reference emits the map inside draw() (myScene.initRender, :1096-1099): it is part of the frame
            stored, extra_launches = D.photon_pass(scene, w["photons"], world, rank, dist)
        else:
            extra_launches = 0
        if world == 1:
            st = scene.draw_device(0, npix, frame.data_ptr())
            out = frame[:npix]
        else:
            st = scene.draw_device_chunks(world, rank, D.CHUNK_ROWS, frame.data_ptr())
            out = D.gather_frame(frame, w["rows"], w["cols"], world, rank, dist)
        st.kernel_launches += extra_launches
        return st, out

    for _ in range(args.warmup):
        flush.fill_(1); step()
    barrier()
    gpu_ms, trace_ms, rays, launches, trace_launch_count = [], 0.0, 0, 0, 0
    with ClockSampler(local, enabled=(rank == 0)) as cs:          # one sampler per job: concurrent nvidia-smi queries from every rank serialise on the driver
        for _ in range(args.steps):
            flush.fill_(1)                       # L2 flush between timed iterations
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            st, out = step()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            t = torch.tensor([ms, float(st.rays_total), float(st.kernel_launches), st.ms_trace], dtype=torch.float64, device="cuda")
            if dist is not None:
                mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX); sm = t.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
                ms, r, l, tr = mx[0].item(), sm[1].item(), sm[2].item(), mx[3].item()
            else:
                r, l, tr = t[1].item(), t[2].item(), t[3].item()
            gpu_ms.append(ms); rays = int(r); launches = int(l); trace_ms += tr
            stage = {"trace": st.ms_trace, "shade": st.ms_shade, "light": st.ms_light, "other": st.ms_other, "render_total": st.ms_total}
    clocks = cs.summary()
    frame_crc = None
    if rank == 0 and out is not None:      # identical for every GPU count (sampler keyed by absolute pixel, canonical photon order)
        import zlib
        frame_crc = "%08x" % (zlib.crc32(out.cpu().numpy().tobytes()) & 0xffffffff)
    ms_per_step = sum(gpu_ms) / len(gpu_ms)
    value = rays / (ms_per_step / 1e3) / 1e6

    # ---- e2e through the C ABI with host buffers (rank-local share for N>1 is not meaningful: measured at N==1 semantics on every rank's full frame)
    e2e = None
    if not args.no_e2e:
        host = torch.empty(npix, dtype=torch.int32).pin_memory()
        scene_bytes = scene.accel_info()["scene_bytes"]            # what drt_scene_reupload copies host -> device
        for _ in range(2):
            scene.reupload(); scene.draw_into(host.data_ptr())
        barrier(); t0 = time.perf_counter(); tot_r = 0
        for _ in range(args.steps):
            ta = time.perf_counter()
            scene.reupload()                        # H2D of the flattened scene
            tb = time.perf_counter()
            st = scene.draw_into(host.data_ptr())   # kernels + D2H of the ARGB frame into pinned host memory
            tot_r += st.rays_total
            if os.environ.get("DRT_BENCH_DEBUG"):
                print("e2e step: reupload %.1f ms, draw_into %.1f ms (gpu %.1f)" % ((tb - ta) * 1e3, (time.perf_counter() - tb) * 1e3, st.ms_total), file=sys.stderr, flush=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        e2e_val = tot_r / dt / 1e6
        if dist is not None:     # every rank rendered a full frame here; report the slowest rank's single-GPU figure
            t = torch.tensor([e2e_val], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN); e2e_val = t.item() * world
        e2e = {"value": round(e2e_val, 3), "unit": "Mrays/s", "h2d_bytes_per_step": int(scene_bytes), "d2h_bytes_per_step": npix * 4,
               "note": "drt_scene_reupload + drt_render into pinned host memory, wall clock" + (" (N independent full frames)" if world > 1 else "")}

    # ---- roofline of the dominant kernel (k_trace, primary level): algorithmic bytes touched per ray (SURVEY 8(d) form, this build's record sizes)
    roof = None; cpu = None
    if rank == 0:
        peaks, which = measured_peaks()
        cctx = drt.Context(device=local, cols=480, rows=270, counters=True)          # same camera/scene at 1/8 linear size: per-ray averages
        cs2 = drt.Scene.from_cli(cctx, w["scene"], spp=w["spp"], photons=min(w["photons"], 100000) if w["photons"] >= 0 else -1, accel=args.accel)
        _, cst = cs2.draw()
        r_all = cst.rays_primary + cst.rays_reflect + cst.rays_refract          # rays traced by k_trace (closest hit)
        box_per_ray, prim_per_ray = cst.box_tests_closest / r_all, cst.prim_tests_closest / r_all
        closest_share = r_all / cst.rays_total
        cctx.close()
        tri_bytes = 104 if args.accel == 0 else 128                                  # one winding state of the FP64 pool / one packed 128-byte record
        bytes_per_ray = 96 + 96 + 64 * box_per_ray + tri_bytes * prim_per_ray           # ray rec in, hit rec out, 64 B per box (128 B node = 2 boxes), triangle bytes
        n_trace_launches = max(1, -(-(npix * w["spp"]) // (8 << 20)))                  # primary-level launches per step (one per batch)
        ms_trace_step = trace_ms / len(gpu_ms)
        closest_rays_per_step = rays * closest_share                                   # primary + reflection + refraction rays of one frame
        achieved = closest_rays_per_step * bytes_per_ray / (ms_trace_step / 1e3) / 1e9
        roof = {"bound": "hbm", "kernel": "k_trace", "achieved": round(achieved, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": round(achieved / peaks["hbm_gbs"], 4),
                "traffic": None, "peak_source": which + " (burst copy)", "per_ray": {"box_tests": round(box_per_ray, 2), "prim_tests": round(prim_per_ray, 2), "bytes": round(bytes_per_ray, 1)},
                "ms_trace_per_step": round(ms_trace_step, 3), "launches_per_step": n_trace_launches,
                "note": "traversal is issue/latency bound with an L2-resident scene (SURVEY 8(d)); bytes are algorithmic bytes touched, not DRAM traffic; see profiles/ for ncu dram bytes and issue-slot utilisation"}
        # the same kernel against the bytes that MUST cross HBM (record in + record out; the scene is cache resident) and against the issue-slot
        # roofline (ncu, profiles/r1c_ncu_k_trace.md): these two say what actually bounds it
        compulsory = closest_rays_per_step * 192.0 / (ms_trace_step / 1e3) / 1e9
        roof["compulsory_hbm"] = {"achieved": round(compulsory, 1), "unit": "GB/s", "frac": round(compulsory / peaks["hbm_gbs"], 4), "bytes_per_ray": 192}
        roof["issue_slots_ncu"] = {"util": 0.446, "fp64_pipe": 0.284, "source": "profiles/r1c_ncu_k_trace.md (sm__inst_issued.avg.pct_of_peak_sustained_active)"}
        tr_path = os.path.join(ROOT, "profiles", "trace_traffic.json")
        if os.path.exists(tr_path):
            try:
                roof["traffic"] = json.load(open(tr_path)).get("dram_bytes_per_launch")
            except Exception:
                pass
        if world == 1 and not args.no_cpu:
            v, secs, crays = cpu_leg(1)
            cpu = {"value": round(v, 4), "unit": "Mrays/s", "cores": 1, "kind": "port",
                   "sample": "oracle (C++ restatement; the Java reference has no JDK here), single thread like the reference, centre 384x216 tile of the 4K/16spp frame: %d rays in %.1f s" % (crays, secs)}

    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 3), "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3),
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "%s %dx%d %dspp%s (%s)" % (w["scene"], w["cols"], w["rows"], w["spp"], (", %d photons cast per light" % w["photons"]) if has_photons else "", w["desc"]),
                           "accel": ["reference-topology literal", "reference-topology fast", "lbvh"][args.accel], "l2": "flushed between timed iterations (256 MiB fill)", "partition": "interleaved 8-row chunks" if world > 1 else "single GPU",
                           "rays_per_frame": rays},
                "frame_ms": round(ms_per_step, 3), "frame_crc32": frame_crc, "accel_info": scene.accel_info(), "stages_ms_rank0_last_step": {k: round(v, 3) for k, v in stage.items()}, "gpu_launches": launches, "clocks": clocks, "e2e": e2e, "roofline": roof, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
