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
                     (rtb_render_bgr8 into pinned HOST memory: the frame arrives as the BMP's pixel
                     bytes, D2H inside the call); e2e_float = rtb_render with a float32 host framebuffer.
            roofline = the dominant traversal kernel (k_walk closest-hit or shadow), algorithmic bytes
                     from the reference's own work counts: 48 B/ray + 32 B/box test +
                     48 B/triangle test, over its CUDA-event time measured on a handle created with
                     RTB_CREATE_KERNEL_TIMING over the same frames (per-launch events cost stream
                     time, so `value` is taken on a handle without them).
reference:  oracle/_ref/ref_driver (the unmodified reference sources, compiled by oracle/Makefile)
            timing launchWorkers + launchSSAA on all host cores; the oracle port when that binary
            is absent.
N > 1:      rows are dealt to ranks in cyclic strips, each rank renders its strips, one
            torch.distributed gather (NCCL) assembles the frame on rank 0; strong scaling.
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
              "cfg3_reflective_refractive_1080": "cfg3_1080", "cfg1_simple_shapes_256": "cfg1_256"}
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
        "config": {"workload": f"{args.scene}: 1920x1080 frame, pass 1 + Sobel + 4x SSAA re-trace", "rays_per_frame": rays},
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
def ours(args):
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

    sc = rb.Scene(rb.scene_path(args.scene))
    h, w = sc.height, sc.width
    stream = torch.cuda.Stream()
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")

    # reference-defined work of the frame, per kernel kind (outside the timed region)
    counted = rb.Renderer(sc, device=local, counters=True)
    _, cst = counted.render()
    counted.close()
    rays = cst["rays"]
    closest_rays = cst["primaryRays"] + cst["secondaryRays"]
    trace_bytes = 48 * closest_rays + 32 * (cst["boxTests"] - cst["boxTestsShadow"]) + 48 * (cst["triTests"] - cst["triTestsShadow"])
    shadow_bytes = 48 * cst["shadowRays"] + 32 * cst["boxTestsShadow"] + 48 * cst["triTestsShadow"]

    # the fast path's OWN work for the same frame (search-BVH nodes fetched, triangles tested), also untimed
    wst_r = rb.Renderer(sc, device=local, walk_stats=True)
    _, wst = wst_r.render()
    wst_r.close()
    own_bytes = [64 * wst["walkNodes"][k] + 48 * wst["walkTris"][k] + 64 * wst["walkEligibility"][k] for k in (0, 1)]
    own_bytes[0] += 48 * closest_rays
    own_bytes[1] += 48 * cst["shadowRays"]

    r = rb.Renderer(sc, device=local)
    out = torch.empty((h, w, 3), dtype=torch.float32, device="cuda") if world == 1 else None
    exchange = rdist.FrameExchange(h, w, args.strip_rows, rank, world, torch.device("cuda", local), args.transport) if world > 1 else None

    def step(rr):
        if world == 1:
            return rr.render_device(out.data_ptr(), stream=stream.cuda_stream), None
        return exchange.render(rr, stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed_loop(rr, steps):
        """K frames, each bracketed by CUDA events on the render stream, L2 flushed (untimed) before each."""
        total, launches = 0.0, 0
        kms = [0.0] * len(KERNEL_KINDS)
        kl = [0] * len(KERNEL_KINDS)
        barrier()
        for _ in range(steps):
            with torch.cuda.stream(stream):
                flush.zero_()                      # evict L2 (256 MiB written, L2 is 126 MB); not timed
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record(stream)
            st, _ = step(rr)
            e1.record(stream)
            e1.synchronize()
            total += e0.elapsed_time(e1)
            launches += st["kernelLaunches"]
            for k in range(len(KERNEL_KINDS)):
                kms[k] += st["msKernel"][k]
                kl[k] += st["launchesKernel"][k]
        barrier()
        return total, launches, kms, kl

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step(r)
    barrier()

    sampler = ClockSampler(local)
    with sampler:
        total_ms, launches, _, _ = timed_loop(r, args.steps)
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        ms_per_step = total_ms / args.steps
        value = rays / (ms_per_step * 1e-3) / 1e6

        # ---- per-kernel durations: the same frames on a handle that brackets every launch with CUDA events
        #      (RTB_CREATE_KERNEL_TIMING; the events themselves cost stream time, so `value` is taken without them)
        rt_ = rb.Renderer(sc, device=local, kernel_timing=True)
        for _ in range(warm):
            step(rt_)
        ksteps = min(args.steps, 50)
        ktotal_ms, _, kernel_ms, kernel_launches = timed_loop(rt_, ksteps)
        rt_.close()

        # ---- end to end through the public host-buffer calls --------------------------------------
        # headline: rtb_render_bgr8, what Scene::render() calls (the frame arrives as the BMP's pixel bytes);
        # also reported: rtb_render with a float32 host framebuffer (4x the bytes over PCIe)
        row_bytes = (w * 3 + 3) & ~3
        host_px = torch.empty((h, row_bytes), dtype=torch.uint8).pin_memory()
        host_fb = torch.empty((h, w, 3), dtype=torch.float32).pin_memory()
        e2e = {}
        for name in ("bgr8", "float"):
            e2e_ms = 0.0
            h2d = d2h = 0
            for i in range(args.steps + 2):
                barrier()
                t0 = time.perf_counter()
                # the step's input is the camera: uploaded (host -> device) inside the timed region, as a frame loop does
                r.set_camera_raw(sc.desc.camera)
                if world == 1:
                    if name == "bgr8":
                        _, est = r.render_bgr8(out=host_px.numpy())
                    else:
                        _, est = r.render(out=host_fb.numpy())
                    h2d_i, d2h_i = est["h2dBytes"], est["d2hBytes"]
                else:
                    est, frame = step(r)
                    h2d_i, d2h_i = est["h2dBytes"], est["d2hBytes"]
                    if rank == 0:
                        stream.synchronize()
                        if name == "bgr8":
                            r.frame_to_bgr8(frame.data_ptr(), host_px.numpy())
                            d2h_i += host_px.numel()
                        else:
                            host_fb.copy_(frame, non_blocking=False)
                            d2h_i += host_fb.numel() * 4
                barrier()
                if i > 1:
                    e2e_ms += (time.perf_counter() - t0) * 1e3
                    h2d, d2h = h2d_i, d2h_i
            te = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            per = float(te.item()) / args.steps
            e2e[name] = {"value": rays / (per * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": per,
                         "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)}

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
    k_trace = KERNEL_KINDS.index("trace")
    k_shadow = KERNEL_KINDS.index("shadow")
    dom = k_trace if kernel_ms[k_trace] >= kernel_ms[k_shadow] else k_shadow
    dom_bytes = (trace_bytes if dom == k_trace else shadow_bytes) / max(1, world)   # per rank share, approx. for N>1
    dom_launches = max(1, kernel_launches[dom])
    launches_per_step = dom_launches / ksteps
    avg_launch_ms = kernel_ms[dom] / dom_launches
    achieved = (dom_bytes / launches_per_step) / (avg_launch_ms * 1e-3) / 1e9 if avg_launch_ms > 0 else 0.0
    traffic = None
    try:   # DRAM bytes per launch of that kernel from the committed ncu --set full capture (profiles/traffic.json)
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tr[args.scene]["k_" + KERNEL_KINDS[dom]]["dram_bytes_per_launch"]
    except Exception:
        pass

    kdom = 0 if dom == k_trace else 1
    own = own_bytes[kdom] / max(1, world)
    own_achieved = (own / launches_per_step) / (avg_launch_ms * 1e-3) / 1e9 if avg_launch_ms > 0 else 0.0
    head = e2e["bgr8"] if e2e.get("bgr8") else e2e["float"]
    line = {
        "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "reference assets (scenes/input), no randomness",
        "config": {"workload": f"{args.scene}: {w}x{h} frame, pass 1 + Sobel + 4x SSAA re-trace", "rays_per_frame": rays,
                   "l2": "flushed between timed frames (256 MiB write)", "partition": (f"cyclic strips of {args.strip_rows} rows; exchange: " + ("NVLink stores into rank 0's symmetric-memory frame + 1 device barrier" if exchange.transport == "p2p" else "1 NCCL gather")) if world > 1 else "single GPU",
                   "timing": "CUDA events on the render stream per frame, summed; max over ranks"},
        "e2e": dict(head, what=("per frame: camera constants uploaded (rtb_set_camera), rtb_render_bgr8 into a pinned host buffer (the call Scene::render() makes; BMP pixel bytes), wall clock"
                                if world == 1 else f"strips rendered per rank, exchanged to rank 0 ({exchange.transport}), converted to BMP pixel bytes there and copied to pinned host memory, wall clock")),
        "e2e_float": e2e["float"] if e2e.get("bgr8") else None,
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": "k_walk (" + ("closest hit" if dom == k_trace else "shadow / any hit") + ")",
                     "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if peak else None, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": dom_bytes / launches_per_step, "avg_launch_ms": avg_launch_ms,
                     "launches_per_step": launches_per_step,
                     "own_work": {"bytes_per_launch": own / launches_per_step, "achieved": own_achieved, "unit": "GB/s",
                                  "frac_of_hbm_peak": own_achieved / peak if peak else None,
                                  "nodes_fetched": wst["walkNodes"][kdom], "triangles_tested": wst["walkTris"][kdom],
                                  "eligibility_evaluations": wst["walkEligibility"][kdom],
                                  "note": "what THIS kernel's algorithm touches: 64 B per search-BVH node fetched + 48 B per triangle "
                                          "tested + 64 B per eligibility evaluation + 48 B per ray; served almost entirely from L1/L2 "
                                          "(see traffic for the DRAM share), so it is a cache-bandwidth figure quoted against the HBM peak"},
                     "note": "bytes = 48/ray + 32/box test + 48/triangle test with the REFERENCE's work counts (SURVEY.md 8d): a work-"
                             "normalised yardstick, not DRAM traffic (the fast path tests ~1/100 of the reference's triangles and the "
                             "geometry is L2-resident), hence frac > 1; durations from a handle with per-launch CUDA events "
                             f"({ksteps} frames, {ktotal_ms / ksteps:.4f} ms/frame with the events)"},
        "kernel_ms_per_step": {KERNEL_KINDS[k]: kernel_ms[k] / ksteps for k in range(len(KERNEL_KINDS))},
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
