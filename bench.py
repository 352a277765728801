#!/usr/bin/env python
"""bench.py — Mrays/s and ms/frame of the Scene::render hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--scene cfg4_shotgun_1080] [--impl reference]

One "step" = one frame: pass 1 (launchWorkers) + Sobel + SSAA re-trace (launchSSAA) of the workload
scene at 1920x1080.  Rays = the number of Render::trace invocations the REFERENCE makes for that
frame (primary + secondary + shadow + SSAA) — the CUDA path reproduces that count
exactly (tests/test_gpu_parity.py), so both arms share the numerator.

ours:       value  = rays / device time, scene and queues resident in HBM, framebuffer left in HBM,
                     CUDA events on the render stream, L2 flushed between timed frames.
            e2e    = the same frame through the public host-buffer call Scene::render() makes
                     (rtb_set_camera + rtb_render_bgr8 into pinned HOST memory: the frame arrives as the BMP's pixel
                     bytes, H2D and D2H inside the call — with a pinned buffer the bytes cross PCIe beside the Sobel /
                     SSAA kernels and the re-traced pixels are rewritten in place); e2e_float = rtb_render with a float32
                     host framebuffer; e2e_pipelined = the begin / end halves with two host buffers in turn.
            gpu_launches = every kernel launched inside the timed region (RtbStats.kernelLaunches).
            roofline = the dominant kernel (k_tile, pass 1): frac = PHYSICAL DRAM bytes per launch (ncu capture at
                     HEAD, profiles/traffic.json) over its live CUDA-event time (handle created with
                     RTB_CREATE_KERNEL_TIMING) against the measured HBM peak; roofline.issue = the issue-slot
                     roofline that actually binds it; roofline.work_normalised = SURVEY.md 8d's yardstick
                     (reference work counts), reported but not a roofline fraction.
            parity_sha_ok = sha256 of the last timed frame equals the committed golden digest.
reference:  oracle/_ref/ref_driver (the unmodified reference sources, compiled by oracle/Makefile)
            timing launchWorkers + launchSSAA on all host cores; the oracle port when that binary
            is absent.
N > 1:      rows are dealt to ranks in cyclic strips (counted from the first row that can contain geometry), each rank
            renders its strips and its output kernel stores them over NVLink straight into rank 0's double-buffered
            symmetric-memory frame, one device-side barrier closes the frame (--transport nccl: one NCCL gather instead);
            strong scaling.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mrays/s"
DEFAULT_SCENE = "cfg4_shotgun_1080"
GOLDEN_KEY = {"cfg4_shotgun_1080": "cfg4_1080", "cfg2_smooth_shading_1024": "cfg2_1024",
              "cfg3_reflective_refractive_1080": "cfg3_1080", "cfg1_simple_shapes_256": "cfg1_256",
              "cfg5_shotgun_2160": "cfg5_2160", "cfgD_dragon_1080": "cfgD_1080"}
L2_FLUSH_BYTES = 256 << 20


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scene", default=DEFAULT_SCENE)
    ap.add_argument("--strip-rows", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--transport", default="auto", choices=["auto", "p2p", "nccl"],
                    help="N>1: how the strips reach rank 0 (p2p = NVLink stores into rank 0's symmetric-memory frame)")
    return ap.parse_args()


def golden_rays(scene):
    try:
        g = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))
        return int(g[GOLDEN_KEY[scene]]["rays"])
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline
# ------------------------------------------------------------------------------------------------
def run_reference(scene, frames, warmup, rays):
    """Times the reference's own CPU implementation.  Returns (ms per frame list, info dict)."""
    cores = os.cpu_count() or 1
    drv = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    scenes_dir = os.path.join(ROOT, "scenes")
    if os.path.exists(drv):
        out = subprocess.run([drv, "bench", scene + ".scene", str(frames + warmup), str(cores)], cwd=scenes_dir, check=True,
                             capture_output=True, text=True).stdout.strip().splitlines()[-1]
        info = json.loads(out)
        ms = info["frame_ms"][warmup:]
        if rays is None:
            st = json.loads(subprocess.run([drv, "stats", scene + ".scene", str(cores)], cwd=scenes_dir, check=True,
                                           capture_output=True, text=True).stdout.strip().splitlines()[-1])
            rays = st["rays"]
        return ms, {"kind": "reference", "cores": info["workers"], "rays": rays,
                    "sample": f"{len(ms)} full frames of {scene} (launchWorkers + launchSSAA), std::thread tile renderer, "
                              f"{info['workers']} workers, after {warmup} warm-up frame(s)"}
    # fall back to the oracle port (plain-C restatement, pthreads over rows)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import rendering_b200 as rb
    from helpers import oracle_render
    sc = rb.Scene(rb.scene_path(scene))
    ms = []
    for i in range(frames + warmup):
        t = time.perf_counter()
        _, _, cnt = oracle_render(sc, threads=cores)
        if i >= warmup:
            ms.append((time.perf_counter() - t) * 1e3)
    return ms, {"kind": "port", "cores": cores, "rays": cnt["rays"],
                "sample": f"{len(ms)} full frames of {scene}, oracle/rtb_oracle.c with {cores} pthreads"}


def scene_size(scene):
    """width, height of a bundled config scene (from its [options] block), without loading any asset"""
    w = h = None
    for ln in open(os.path.join(ROOT, "scenes", scene + ".scene")):
        ln = ln.strip()
        if ln.startswith("width="):
            w = int(ln.split("=")[1])
        elif ln.startswith("height="):
            h = int(ln.split("=")[1])
    return w, h


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rays = golden_rays(args.scene)
    # bounded: a frame of the 250k-triangle scene takes the reference tens of seconds; cap the run at a few minutes
    probe, info = run_reference(args.scene, 1, 0, rays)
    budget_frames = int(max(1, min(args.steps, 120000.0 / max(probe[0], 1e-3))))
    if budget_frames > 1 or args.warmup > 0 and probe[0] < 30000.0:
        ms, info = run_reference(args.scene, budget_frames, 1 if probe[0] < 30000.0 else 0, rays)
    else:
        ms = probe
    rays = info.pop("rays")
    mean_ms = sum(ms) / len(ms)
    value = rays / (mean_ms * 1e-3) / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": mean_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "reference assets (scenes/input), no randomness",
        "config": workload_config(args.scene, *scene_size(args.scene), rays),
        "cpu_baseline": dict(info, value=value, unit="Mrays/s"),
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.stop = threading.Event()
        self.thread = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.05)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# ours
# ------------------------------------------------------------------------------------------------
def workload_config(scene, w, h, rays):
    """The `config` object: identical in both arms (the driver compares them)."""
    return {"workload": f"{scene}: {w}x{h} frame, pass 1 + Sobel + 4x SSAA re-trace", "rays_per_frame": rays}


def ours(args):
    import hashlib
    import numpy as np
    import torch
    import torch.distributed as dist
    import rendering_b200 as rb
    from rendering_b200 import dist as rdist
    from rendering_b200._ffi import KERNEL_KINDS

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl ours) needs a CUDA device: the renderer has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    t0 = time.perf_counter()
    sc = rb.Scene(rb.scene_path(args.scene))          # .scene / .obj / .bmp loaders + reference tree build (host)
    load_ms = (time.perf_counter() - t0) * 1e3
    h, w = sc.height, sc.width
    stream = torch.cuda.Stream()
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")

    # reference-defined work of the frame (outside the timed region): the literal reference walk with counters
    counted = rb.Renderer(sc, device=local, counters=True)
    _, cst = counted.render()
    counted.close()
    rays = cst["rays"]
    yardstick_bytes = 48 * rays + 32 * cst["boxTests"] + 48 * cst["triTests"] + 12 * w * h      # SURVEY.md 8d, whole frame

    # the fast path's OWN work for the same frame (search-BVH nodes fetched, triangles tested), also untimed
    wst_r = rb.Renderer(sc, device=local, walk_stats=True)
    _, wst = wst_r.render()
    wst_r.close()
    own_bytes = sum(64 * wst["walkNodes"][k] + 48 * wst["walkTris"][k] + 64 * wst["walkEligibility"][k] for k in (0, 1)) + 48 * rays

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = rb.Renderer(sc, device=local)                  # upload + search-BVH build + eligibility tables (rtb_create)
    create_ms = (time.perf_counter() - t0) * 1e3
    out = torch.empty((h, w, 3), dtype=torch.float32, device="cuda") if world == 1 else None
    exchange = (rdist.FrameExchange(h, w, args.strip_rows, rank, world, torch.device("cuda", local), args.transport, origin=r.strip_origin())
                if world > 1 else None)

    def step(rr, end_event=None):
        """One frame, device-resident result.  Everything is enqueued before the host waits: frame, exchange barrier (N > 1) and
        the caller's closing event, so no host latency sits between the last kernel and the event."""
        if world == 1:
            rr.render_device_begin(out.data_ptr(), stream=stream.cuda_stream)
            if end_event is not None:
                end_event.record(stream)
            return rr.render_end(), out
        return exchange.render(rr, stream, end_event)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed_loop(rr, steps):
        """K frames, each bracketed by CUDA events on the render stream, L2 flushed (untimed) before each."""
        total, launches, dev_ms = 0.0, 0, 0.0
        kms = [0.0] * len(KERNEL_KINDS)
        kl = [0] * len(KERNEL_KINDS)
        last = None
        barrier()
        for _ in range(steps):
            with torch.cuda.stream(stream):
                flush.zero_()                      # evict L2 (256 MiB written, L2 is 126 MB); not timed
            if exchange is not None:
                exchange.align(stream)             # all ranks start the frame together (untimed device-side rendezvous)
            with torch.cuda.stream(stream):
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record(stream)
            st, last = step(rr, e1)
            e1.synchronize()
            total += e0.elapsed_time(e1)
            dev_ms += st["msTotal"]              # this rank's own frame (fill .. output kernel), without the exchange
            launches += st["kernelLaunches"]
            for k in range(len(KERNEL_KINDS)):
                kms[k] += st["msKernel"][k]
                kl[k] += st["launchesKernel"][k]
        barrier()
        timed_loop.rank_ms = dev_ms / steps
        return total, launches, kms, kl, st, last

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step(r)
    barrier()

    sampler = ClockSampler(local)
    with sampler:
        total_ms, launches, _, _, fst, last_frame = timed_loop(r, args.steps)
        per_rank = torch.zeros(world, dtype=torch.float64, device="cuda")
        per_rank[rank] = timed_loop.rank_ms
        if world > 1:
            dist.all_reduce(per_rank)
        per_rank = [float(x) for x in per_rank.tolist()]
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        ms_per_step = total_ms / args.steps
        value = rays / (ms_per_step * 1e-3) / 1e6

        # ---- parity of the frame just timed: sha256 of the last timed frame against the committed golden digest ----
        parity = None
        if rank == 0:
            digest = hashlib.sha256(last_frame.cpu().numpy().tobytes()).hexdigest()
            try:
                g = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))[GOLDEN_KEY[args.scene]]
                parity = {"frame_sha256": digest, "golden": "tests/golden/golden.json:" + GOLDEN_KEY[args.scene],
                          "ok": digest == g["final_sha256_stateless"], "equals_reference_digest": digest == g["final_sha256"],
                          "pixels_where_the_reference_differs": g.get("stateful_pixels", 0),
                          "note": "golden = digest of the unmodified reference's framebuffer (oracle/_ref, tests/golden/make_golden.py); where a "
                                  "normal-map texel is fetched twice the reference's in-place normalisation (objects.cpp:148) drifts by an ulp and the pin "
                                  "is the stateless digest, attributed pixel for pixel by tests/test_oracle.py"}
            except Exception as e:   # noqa: BLE001
                parity = {"frame_sha256": digest, "ok": None, "note": f"no golden digest: {e}"}

        # ---- per-kernel durations: the same frames on a handle that brackets every launch with CUDA events
        #      (RTB_CREATE_KERNEL_TIMING; the events themselves cost stream time, so `value` is taken without them)
        rt_ = rb.Renderer(sc, device=local, kernel_timing=True)
        for _ in range(warm):
            step(rt_)
        ksteps = min(args.steps, 50)
        ktotal_ms, _, kernel_ms, kernel_launches, _, _ = timed_loop(rt_, ksteps)
        rt_.close()

        # ---- end to end through the public host-buffer calls --------------------------------------
        # headline: rtb_render_bgr8, what Scene::render() calls (the frame arrives as the BMP's pixel bytes);
        # also reported: rtb_render with a float32 host framebuffer (4x the bytes over PCIe).
        # N > 1: no host-level barrier between frames — the exchange's device-side barrier is the only synchronisation.
        row_bytes = (w * 3 + 3) & ~3
        host_px = torch.empty((h, row_bytes), dtype=torch.uint8).pin_memory()
        host_fb = torch.empty((h, w, 3), dtype=torch.float32).pin_memory()
        e2e = {}
        for name in ("bgr8", "float"):
            h2d = d2h = 0

            def frame():
                # the step's input is the camera: uploaded (host -> device) inside the timed region, as a frame loop does
                r.set_camera_raw(sc.desc.camera)
                if world == 1:
                    _, est = r.render_bgr8(out=host_px.numpy()) if name == "bgr8" else r.render(out=host_fb.numpy())
                    return est["h2dBytes"], est["d2hBytes"]
                est, fr = step(r)
                extra = 0
                if rank == 0:
                    if name == "bgr8":
                        r.frame_to_bgr8(fr.data_ptr(), host_px.numpy(), stream=stream.cuda_stream)
                        extra = host_px.numel()
                    else:
                        with torch.cuda.stream(stream):
                            host_fb.copy_(fr, non_blocking=True)
                        stream.synchronize()
                        extra = host_fb.numel() * 4
                return est["h2dBytes"], est["d2hBytes"] + extra

            for _ in range(2):
                frame()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                h2d, d2h = frame()
            barrier()
            e2e_ms = (time.perf_counter() - t0) * 1e3
            te = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            per = float(te.item()) / args.steps
            e2e[name] = {"value": rays / (per * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": per,
                         "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)}

        # ---- the same loop with the output pipelined (frame loop at steady state): every frame still uploads its
        #      camera and delivers its BMP bytes to pinned host memory inside the timed region, but the device-to-host copy of frame i
        #      overlaps the kernels of frame i+1 (rtb_render_bgr8_begin / rtb_render_end, two host buffers in turn)
        e2e_pipe = None
        if world == 1:
            pair = [torch.empty((h, row_bytes), dtype=torch.uint8).pin_memory() for _ in range(2)]
            for i in range(4):
                r.set_camera_raw(sc.desc.camera)
                r.render_bgr8_begin(pair[i & 1].numpy())
                r.render_end()
            r.output_sync()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(args.steps):
                r.set_camera_raw(sc.desc.camera)
                r.render_bgr8_begin(pair[i & 1].numpy())
                est = r.render_end()
            r.output_sync()
            per = (time.perf_counter() - t0) * 1e3 / args.steps
            ok = bool((pair[(args.steps - 1) & 1] == host_px).all())          # the last pipelined frame equals the synchronous call's bytes
            e2e_pipe = {"value": rays / (per * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": per, "h2d_bytes_per_step": int(est["h2dBytes"]),
                        "d2h_bytes_per_step": int(est["d2hBytes"]), "bytes_equal_synchronous_call": ok,
                        "what": "frame loop through rtb_set_camera + rtb_render_bgr8_begin + rtb_render_end with two pinned host buffers in turn: "
                                "the copy of frame i overlaps the kernels of frame i+1; wall clock over all frames incl. the last copy"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    k_tile, k_ssaa = KERNEL_KINDS.index("tile"), KERNEL_KINDS.index("tile_ssaa")
    dom, dom_name = (k_tile, "k_tile") if kernel_ms[k_tile] >= kernel_ms[k_ssaa] else (k_ssaa, "k_tile_ssaa")
    dom_launches = max(1, kernel_launches[dom])
    avg_launch_ms = kernel_ms[dom] / dom_launches
    tile_ms_per_frame = (kernel_ms[k_tile] + kernel_ms[k_ssaa]) / ksteps
    prof = {}
    try:   # physical DRAM bytes and issue statistics per launch from the committed ncu --set full capture (profiles/traffic.json)
        prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[args.scene][dom_name]
    except Exception:
        pass
    traffic = prof.get("dram_bytes_per_launch")
    traffic_rank = traffic / max(1, world) if traffic else None           # per rank share, approx. for N>1
    achieved = traffic_rank / (avg_launch_ms * 1e-3) / 1e9 if traffic_rank and avg_launch_ms > 0 else None
    issue = dict(prof.get("issue", {}))
    if issue and avg_launch_ms > 0:
        # issue-slot roofline: warp instructions of the launch / (148 SMs x 4 schedulers x SM clock) over the LIVE launch time
        issue["frac_of_issue_peak_live"] = issue["inst_executed"] / max(1, world) / (148 * 4 * 1.965e9) / (avg_launch_ms * 1e-3)
    traced = rays - fst["shadowRaysSkipped"] - fst["backgroundPixels"] if world == 1 else None
    head = e2e["bgr8"]
    line = {
        "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "reference assets (scenes/input), no randomness",
        "config": workload_config(args.scene, w, h, rays),
        "method": {"l2": "flushed between timed frames (256 MiB write)",
                   "partition": (f"cyclic strips of {args.strip_rows} rows counted from the first geometry row; exchange: "
                                 + ("NVLink stores into rank 0's double-buffered symmetric-memory frame + 1 device barrier" if exchange.transport == "p2p" else "1 NCCL gather"))
                   if world > 1 else "single GPU",
                   "timing": "CUDA events on the render stream per frame, summed; max over ranks",
                   "pipeline": "tile (k_tile: whole recursion per 256-ray tile inside one persistent kernel per pass; 64-ray tiles for passes that cannot fill the machine)"},
        "parity_sha_ok": parity["ok"] if parity else None, "parity": parity,
        "rays": {"reference_equivalent": rays, "note": "numerator of value / e2e: the reference's Render::trace call count for this frame (SURVEY.md 8d)",
                 "traced": traced, "background_prefilled_primaries": fst["backgroundPixels"] if world == 1 else None,
                 "dead_shadow_rays_not_traced": fst["shadowRaysSkipped"] if world == 1 else None,
                 "value_traced": traced / (ms_per_step * 1e-3) / 1e6 if traced else None},
        "e2e": dict(head, what=("per frame: camera constants uploaded (rtb_set_camera), rtb_render_bgr8 into a pinned host buffer (the call Scene::render() makes; BMP pixel bytes: copied to the host beside the Sobel / SSAA kernels, the re-traced pixels rewritten in place, DESIGN.md 4 'Early output'), wall clock"
                                if world == 1 else f"per frame: camera uploaded on every rank, strips rendered, exchanged to rank 0 ({exchange.transport}), converted to BMP pixel bytes there and copied to pinned host memory; wall clock over all frames, no host barrier between frames")),
        "e2e_float": e2e["float"], "e2e_pipelined": e2e_pipe,
        "gpu_launches": launches, "launches_per_frame": launches / args.steps,
        "per_rank_render_ms": per_rank,
        "roofline": {"bound": "hbm", "kernel": dom_name + (" (pass 1: ray generation, traversal, surface, shadow, shade of every tile)" if dom == k_tile else " (SSAA samples)"),
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved and peak else None,
                     "traffic": traffic, "peak_source": peak_src, "avg_launch_ms": avg_launch_ms, "launches_per_step": dom_launches / ksteps,
                     "what": "PHYSICAL: dram__bytes_read+write per launch of that kernel (ncu --set full at HEAD, profiles/traffic.json) over its live CUDA-event "
                             "duration.  The scene is L1/L2-resident, so the kernel is nowhere near the HBM roof; it is bound by instruction issue (see issue)",
                     "issue": issue or None,
                     "work_normalised": {"bytes_per_frame": yardstick_bytes, "achieved": yardstick_bytes / max(1, world) / (tile_ms_per_frame * 1e-3) / 1e9 if tile_ms_per_frame > 0 else None,
                                         "unit": "GB/s", "over": "both k_tile launches of a frame",
                                         "note": "SURVEY.md 8d yardstick: 48 B/ray + 32 B/box test + 48 B/triangle test with the REFERENCE's work counts + 12 B/pixel; the fast "
                                                 "path tests ~1/100 of the reference's triangles, so this exceeds the HBM peak and is NOT a roofline fraction"},
                     "own_work": {"bytes_per_frame": own_bytes, "achieved": own_bytes / max(1, world) / (tile_ms_per_frame * 1e-3) / 1e9 if tile_ms_per_frame > 0 else None, "unit": "GB/s",
                                  "nodes_fetched": sum(wst["walkNodes"]), "triangles_tested": sum(wst["walkTris"]), "eligibility_evaluations": sum(wst["walkEligibility"]),
                                  "note": "what THIS algorithm touches: 64 B per search-BVH node + 48 B per triangle tested + 64 B per eligibility evaluation + 48 B per ray, "
                                          "served from L1/L2: a cache-bandwidth figure"},
                     "timing_note": f"per-launch durations from a handle with CUDA events around every launch ({ksteps} frames, {ktotal_ms / ksteps:.4f} ms/frame with the events)"},
        "kernel_ms_per_step": {KERNEL_KINDS[k]: kernel_ms[k] / ksteps for k in range(len(KERNEL_KINDS)) if kernel_launches[k]},
        "setup_ms": {"scene_load_host": load_ms, "rtb_create": create_ms, "note": "outside the timed region: loaders + reference tree (host); upload, search BVH, eligibility tables"},
        "clocks": sampler.summary(),
    }
    if args.gpus == 1 and not args.no_cpu_baseline:
        try:
            # bounded sample: one probe frame tells how many full frames fit in ~20 s of CPU time
            probe, info = run_reference(args.scene, 1, 0, rays)
            frames = int(max(1, min(8, 20000.0 / max(probe[0], 1e-3))))
            if frames > 1:
                ms, info = run_reference(args.scene, frames, 1, rays)
            else:
                ms = probe
                info["sample"] = info["sample"].replace("after 0 warm-up frame(s)", "single frame, no warm-up (one frame takes %.0f s)" % (probe[0] / 1e3))
            info.pop("rays", None)
            mean_ms = sum(ms) / len(ms)
            line["cpu_baseline"] = dict(info, value=rays / (mean_ms * 1e-3) / 1e6, unit="Mrays/s", ms_per_frame=mean_ms)
        except Exception as e:   # the baseline is reported, never required for the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "Mrays/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {e}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
