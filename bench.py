#!/usr/bin/env python
"""bench.py — FAST + BRIEF + match throughput on synthetic KITTI-shaped frames (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

A step = one pass of the front end over one batch of B (default 1024) consecutive 1241x376 8-bit frames per GPU:
detect (FAST segment test + Harris) -> exact top-2000 -> BRIEF -> Hamming match of frame f-1 -> f.
`value`  frames/s with the batch already resident in HBM (device timed, CUDA events on the
         context's stream, max over ranks);
`e2e`    frames/s through the C ABI with pinned HOST buffers: H2D of the pixels and D2H of keypoints,
         descriptors and matches inside the timed region;
`roofline` HBM roofline of the detect+describe kernels from live per-kernel CUDA-event times,
`roofline_popc` the POPC-pipe roofline of the match kernel; `cpu_baseline` the CPU oracle on this box.
--impl reference times the reference's CPU algorithm (oracle port; the literal reference cannot be
built here: OpenCV C++ is absent) on all host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W, MAX_KP = 376, 1241, 2000
METRIC = "frames_per_s_fast_brief_match_1241x376"
UNIT = "frames/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback", 1965.0


def make_frames(n, seed0, kind="G30"):
    from ya_vo_b200 import synth
    return synth.synth_batch(n, kind, seed0, H, W)


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, via NVML (nvidia-smi's own source) sampled from a
    background thread every 20 ms; the ctypes calls of the timed loop release the GIL."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index):
        import threading
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._th = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.nv is None:
            return
        import threading
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()

    def stop(self):
        if self._th is not None:
            self._stop.set()
            self._th.join(timeout=2)
        if not self.samples:
            # fall back to one nvidia-smi query (idle clocks; flagged by samples == 0)
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=10).stdout.splitlines()[0].split(",")
                return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": [], "samples": 0}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def workload_config(batch, kind):
    """Shared by both arms so the driver compares like with like."""
    return {"workload": "configs[1] batched (configs[2] shape): consecutive 1241x376 8-bit frames, FAST+BRIEF on each + "
                        "brute-force Hamming match of frame f-1 -> f",
            "frames_per_step_per_gpu": batch, "frame": [H, W], "max_keypoints": MAX_KP,
            "input": "%s seed 1000+f" % kind}


def bind_to_gpu_numa_node(gpu_index):
    """Pins this rank to the CPUs NVML reports as local to its GPU, so the pinned frame / result buffers it
    allocates afterwards live on that socket (matters for the host-fed e2e leg at 4-8 GPUs)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1 and 64 * w + b < ncpu]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return len(allowed)
    except Exception:
        pass
    return None


def cpu_baseline(sample_frames, offsets, cores):
    """The oracle (kind 'port') on a bounded sample of the same workload, all requested host threads."""
    from oracle import pyoracle as po
    po.build()
    n = sample_frames.shape[0]
    t0 = time.perf_counter()
    po.pipeline(sample_frames, offsets, MAX_KP, True, nthreads=cores, outputs=False)
    dt = time.perf_counter() - t0
    return n / dt, dt


def run_reference(args):
    """--impl reference: the reference's CPU algorithm for the same metric/config on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from ya_vo_b200 import synth
    cores = os.cpu_count() or 1
    offsets = synth.brief_offsets()
    n = max(8, min(args.batch, 8 * cores))  # bounded sample per step (~0.5 s wall per step on the box's cores)
    frames = make_frames(n, 1000, args.kind)
    from oracle import pyoracle as po
    po.build()
    for _ in range(args.warmup):
        po.pipeline(frames[: max(2, n // 4)], offsets, MAX_KP, True, nthreads=cores, outputs=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        po.pipeline(frames, offsets, MAX_KP, True, nthreads=cores, outputs=False)
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    sample = "%d of the %d frames of a step per timed step, oracle port of FastDetector/Brief, %d threads" % (n, args.batch, cores)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": dict(workload_config(args.batch, args.kind), note="literal reference cannot be built here (OpenCV C++ "
                       "absent); this is the CPU oracle port with identical outputs, all host threads"),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="frames per step per GPU (1024 x 0.48 MB pitched = 493 MB >> 126 MB L2)")
    ap.add_argument("--kind", default="G30")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sub-batch", type=int, default=-1, help="frames per kernel sub-batch (-1 = library default)")
    ap.add_argument("--chunk", type=int, default=-1, help="frames per copy/compute pipeline stage of the e2e leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from ya_vo_b200 import capi, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep stdout to the one JSON line
    numa = bind_to_gpu_numa_node(local)  # pinned buffers are first-touched on the GPU's own socket
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    B = args.batch
    capi.build()
    offsets = synth.brief_offsets()
    # frames of this rank's shard: frame f of the job is seed 1000+f
    frames = make_frames(B, 1000 + rank * B, args.kind)
    pinned = torch.empty((B, H, W), dtype=torch.uint8).pin_memory()
    pinned.numpy()[:] = frames
    frames_pinned = pinned.numpy()

    ctx = capi.Context(device=local, n_slots=B, max_rows=H, max_cols=W, max_kp=MAX_KP)
    ctx.set_brief_offsets(offsets)
    if args.sub_batch >= 0:
        ctx.set_sub_batch(args.sub_batch)
    if args.chunk > 0:
        ctx.set_pipeline_chunk(args.chunk)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))

    # ---- device-resident leg: `value` ---------------------------------------------------------------
    ctx.upload_batch(0, frames_pinned)
    ctx.sync()
    for _ in range(max(args.warmup, 3)):
        ctx.frontend_batch(0, B, True)
    ctx.sync()
    sampler = ClockSampler(local)
    ctx.set_profiling(True)
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    l0 = ctx.kernel_launches
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        ctx.frontend_batch(0, B, True)
    e1.record(stream)
    ctx.sync()
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop()
    launches = ctx.kernel_launches - l0
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    prof = ctx.profile_collect()
    ctx.set_profiling(False)
    value = world * B * args.steps / (dev_ms * 1e-3)

    # ---- end-to-end leg: host buffers through the C ABI ----------------------------------------------
    K = MAX_KP
    outs = {}
    shapes = dict(n_kp=((B,), torch.int32), rows=((B, K), torch.int32), cols=((B, K), torch.int32),
                  scores=((B, K), torch.float32), desc=((B, K, 32), torch.uint8), match_idx=((B, K), torch.int32),
                  match_dist=((B, K), torch.int32))
    keep = []
    for k, (shp, dt) in shapes.items():
        t = torch.empty(shp, dtype=dt).pin_memory()
        keep.append(t)
        outs[k] = t.numpy()
    h2d = B * H * W
    d2h = sum(int(np.prod(s)) * torch.empty((), dtype=d).element_size() for s, d in shapes.values())
    # two sets of pinned output arrays: step i's results travel back while step i+1 is uploaded and computed
    outs2 = {}
    for k, (shp, dt) in shapes.items():
        t = torch.empty(shp, dtype=dt).pin_memory()
        keep.append(t)
        outs2[k] = t.numpy()
    obuf = [outs, outs2]
    for _ in range(max(args.warmup, 3)):
        ctx.process_host_batch(frames_pinned, True, outs)
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tickets = []
    seen = 0
    for i in range(args.steps):
        # every step: H2D of its pinned frames, kernels, D2H of keypoints + descriptors + matches.  Software
        # pipelined one step deep: step i-1's results are waited for (and touched) while step i is in flight.
        _, t = ctx.submit_host_batch(frames_pinned, True, obuf[i & 1])
        tickets.append(t)
        if i >= 1:
            ctx.wait_batch(tickets[i - 1])
            seen += int(obuf[(i - 1) & 1]["n_kp"][0])
    ctx.wait()  # the last step's results are in host memory too
    seen += int(obuf[(args.steps - 1) & 1]["n_kp"][0])
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e = world * B * args.steps / e2e_s
    n_kp_mean = float(outs["n_kp"].mean())

    # ---- rooflines -------------------------------------------------------------------------------------
    peak_gbs, peak_kind, sm_max = load_peaks()
    steps = args.steps
    ms = {k: v[0] / max(v[1], 1) for k, v in prof.items()}  # average launch duration per class
    ms_step = {k: v[0] / steps for k, v in prof.items()}    # device time per step per class (all its launches)
    dd_ms = ms_step["detect_blur"] + ms_step["compact_score"] + ms_step["select_topk"] + ms_step["brief"]
    b_frame = W * H + 44 * n_kp_mean  # pixels read once + (row,col,score,descriptor) per keypoint
    achieved = B * b_frame / (dd_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("detect_describe_dram_bytes_per_frame") * B  # per step, like `achieved`
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "detect_blur+compact_score+select_topk+brief (device time of their launches in one step)",
                "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                "peak_kind": peak_kind, "traffic": traffic, "ms_per_step": dd_ms,
                "algorithmic_bytes_per_step": B * b_frame,
                "note": "ALU-issue bound, not HBM bound: ~1 byte/pixel compulsory traffic vs tens of integer ops/pixel"}
    # issue-slot roofline of the detect kernel: its instruction count per frame (ncu, profiles/traffic.json) against
    # 4 warp instructions per clock per SM — the limit that actually binds this ALU-heavy kernel
    roofline_issue = None
    try:
        wi = json.load(open(tp))["warp_insts_per_frame_by_kernel"]["detect_blur_kernel<1, 1>"]
        clk = (clocks.get("sm_mhz") or sm_max) * 1e6
        peak_issue = 148 * 4 * clk
        ach = wi * B / (ms_step["detect_blur"] * 1e-3)
        roofline_issue = {"bound": "issue", "kernel": "detect_blur", "achieved": ach / 1e9, "peak": peak_issue / 1e9,
                          "unit": "G warp-inst/s", "frac": ach / peak_issue, "warp_insts_per_frame": wi,
                          "lane_insts_per_pixel": wi * 32 / (H * W)}
    except Exception:
        pass
    pairs = float((outs["n_kp"][:-1].astype(np.float64) * outs["n_kp"][1:].astype(np.float64)).sum())
    m_ms = prof["match_partial"][0] / steps  # all match launches of one step
    gpairs = pairs / (m_ms * 1e-3) / 1e9 if m_ms > 0 else 0.0
    sm_mhz = clocks.get("sm_mhz") or sm_max
    popc_peak = 148 * 16 * sm_mhz * 1e6 / 8 / 1e9  # 16 POPC.32/clk/SM, 8 POPC per 256-bit pair
    roofline_popc = {"bound": "popc", "kernel": "match_partial", "achieved": gpairs, "peak": popc_peak,
                     "unit": "Gpairs/s", "frac": gpairs / popc_peak if popc_peak else None,
                     "peak_kind": "148 SM x 16 POPC/clk/SM x %.0f MHz (sampled) / 8 POPC per pair" % sm_mhz,
                     "ms_per_step": m_ms, "pairs_per_step": pairs}
    total_kernel_ms = sum(v[0] for v in prof.values()) / steps
    kernel_share = {k: (v[0] / steps) / total_kernel_ms for k, v in prof.items() if total_kernel_ms > 0}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": dict(workload_config(B, args.kind), **{"mean_keypoints": n_kp_mean,
                   "l2": "inputs larger than L2 (%d frames x %.2f MB pitched = %.0f MB > 126 MB)" % (B, 1280 * H / 1e6, B * 1280 * H / 1e6),
                   "parallelism": "frame-sharded, %d rank(s), no collective on the data path" % world,
                   "cpus_bound_per_rank": numa}),
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * e2e_s / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "roofline_popc": roofline_popc,
        "roofline_issue": roofline_issue,
        "kernel_ms_per_launch": ms,
        "kernel_ms_per_step": ms_step,
        "kernel_launches_per_step": {k: v[1] / steps for k, v in prof.items()},
        "kernel_share_of_step": kernel_share,
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            os.sched_setaffinity(0, range(os.cpu_count() or 1))  # the CPU baseline may use every host core
        except Exception:
            pass
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        ns = max(8, min(B, 16 * cores))  # ~15-30 s of CPU work (about 65 ms per frame and thread)
        v, dt = cpu_baseline(frames[:ns], offsets, cores)
        v1, dt1 = cpu_baseline(frames[:8], offsets, 1)  # the reference itself is single-threaded
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "first %d frames of the step (%.1f s wall, %d threads), oracle port" % (ns, dt, cores),
                                "single_thread_value": v1, "single_thread_sample": "first 8 frames, %.1f s" % dt1}
    elif rank == 0:
        line["cpu_baseline"] = None
    ctx.close()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
