#!/usr/bin/env python
"""Benchmark of the hot path: JPEG XL tile encode of BASELINE.json configs[1]
(4096x4096 sRGB8, tile mode, 256 independent 256x256 groups) on N B200s.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's own CPU encoder on the host cores

One step = one pass of the hot path over one 4096x4096 image per GPU (weak scaling: every rank
encodes its own image; rank 0 then gathers all codestreams with one NCCL gather).  Prints ONE JSON
line on rank 0.  See DESIGN.md "Measurement" for the definition of every field.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH = HEIGHT = 4096
CHANNELS = 3
METRIC = "Mpixels/s encoded"
UNIT = "Mpx/s"
WORKLOAD = "4096x4096 sRGB8 lossy (reference's fixed quantiser), tile mode shift 0/0, 256 independent 256x256 groups"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# reference arm: the unmodified reference library (oracle/_ref) on the host cores
# ------------------------------------------------------------------------------------------------
def reference_encode_threaded(lib, image: np.ndarray, threads: int) -> tuple[float, int]:
    """Encode all tiles of `image` with `threads` independent reference encoders (tile frames are
    independent, gaps are legal: libhydrium.h:240).  Returns (seconds, bytes)."""
    from hydrium_b200.encoder import encode_cli_loop
    h, w, _ = image.shape
    nty, ntx = (h + 255) // 256, (w + 255) // 256
    tiles = [(x, y) for y in range(nty) for x in range(ntx)]
    parts = [tiles[i::threads] for i in range(threads)]
    sizes = [0] * threads

    def work(i):
        if parts[i]:
            sizes[i] = len(encode_cli_loop(lib, image, tiles=parts[i], is_last=0))

    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    return time.perf_counter() - t0, sum(sizes)


def run_reference(args) -> int:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from hydrium_b200.synth import synth_image
    from oracle.pyoracle import have_ref, ref_library
    if not have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (needs /root/reference at build time)"}))
        return 0
    lib = ref_library("Os")
    cores = os.cpu_count() or 1
    image = synth_image(WIDTH, HEIGHT, 8)
    for _ in range(args.warmup):
        reference_encode_threaded(lib, image[:1024], cores)
    times = []
    for _ in range(args.steps):
        dt, _ = reference_encode_threaded(lib, image, cores)
        times.append(dt)
    mpx = WIDTH * HEIGHT / 1e6
    ms = 1e3 * sum(times) / len(times)
    value = mpx / (ms / 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "reference libhydrium (-Os, project flags) on host cores; "
                   "one encoder per thread, tiles dealt round-robin"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": f"full {WIDTH}x{HEIGHT} image per step, {cores} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def bind_to_gpu_cpus(device_index: int):
    """Pin this process to the CPUs NVML reports as local to its GPU (one process per GPU: the page-locked
    host buffers of the end-to-end path should sit on the GPU's NUMA node).  Returns a description."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(visible.split(",")[device_index]) if visible and visible.split(",")[device_index].isdigit() else device_index
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} GPU-local CPUs (NVML)"
    except Exception as e:   # noqa: BLE001 - best effort, the bench runs unpinned otherwise
        return f"unpinned ({type(e).__name__})"
    return "unpinned"


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args) -> int:
    import torch
    import torch.distributed as dist

    from hydrium_b200.abi import HYD_UINT8
    from hydrium_b200.dist import PeerGather, gather_spans
    from hydrium_b200.engine import Engine, output_bound

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: hydrium_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    affinity = bind_to_gpu_cpus(local_rank)   # before any page-locked allocation: host buffers land on the GPU's NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    eng = Engine(device=local_rank, max_batch_tiles=256)
    ext = torch.cuda.ExternalStream(eng.stream, device=dev)
    n_in = WIDTH * HEIGHT * CHANNELS
    cap = output_bound(WIDTH, HEIGHT)
    d_in = torch.empty(n_in, dtype=torch.uint8, device=dev)
    d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    # every rank gets its own image (different seed): weak scaling over independent images
    eng.synth_fill(d_in.data_ptr(), WIDTH, HEIGHT, bits=8, seed=rank)

    def flush_l2():
        with torch.cuda.stream(ext):
            flush_buf.fill_(rank & 0xFF)

    # rank 0 receives every rank's codestream here (sized for the synthetic's ~0.35 B/px with head-room)
    gbuf = torch.empty(world * n_in // 4, dtype=torch.uint8, device=dev) if (world > 1 and rank == 0) else None
    enc_done = [None]
    # The gather.  Preferred: every rank's compaction kernel writes its span straight into
    # rank 0's HBM over NVLink (CUDA IPC peer memory), then one tiny all-reduce as the barrier and one
    # kernel on rank 0 that closes the gaps.  Fallback: NCCL send/recv of the spans (gather_spans).
    pg = None
    gather_kind = None
    if world > 1:
        ok = torch.ones(1, dtype=torch.int32, device=dev)
        try:
            pg = PeerGather(eng, cap)
        except RuntimeError as e:
            print(f"[bench] rank {rank}: peer-memory gather unavailable ({e}); using NCCL send/recv", file=sys.stderr)
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            if pg is not None:
                pg.close()
            pg = None
        gather_kind = ("peer memory: k_gather_frames writes into rank 0's HBM over NVLink (CUDA IPC), all-reduce barrier, "
                       "k_compact_regions on rank 0") if pg is not None else "NCCL: all-gather of lengths + grouped send/recv"

    def step(gather: bool):
        if gather and pg is not None:
            n = eng.encode_image_device(d_in.data_ptr(), WIDTH, HEIGHT, CHANNELS, sample_fmt=HYD_UINT8,
                                        d_out=pg.d_out, d_out_cap=pg.d_out_cap)
            with torch.cuda.stream(ext):
                if enc_done[0] is not None:
                    enc_done[0].record()
                pg.finish(n, gbuf.data_ptr() if gbuf is not None else 0, gbuf.numel() if gbuf is not None else 0)
            return n
        n = eng.encode_image_device(d_in.data_ptr(), WIDTH, HEIGHT, CHANNELS, sample_fmt=HYD_UINT8,
                                    d_out=d_out.data_ptr(), d_out_cap=cap)
        if gather and world > 1:
            with torch.cuda.stream(ext):
                if enc_done[0] is not None:
                    enc_done[0].record()
                gather_spans(d_out[:n], dst=0, out=gbuf)
        return n

    for _ in range(max(args.warmup, 3)):
        step(True)
    torch.cuda.synchronize()

    # ---- device-resident throughput ("value") -------------------------------------------------
    launches0 = eng.launch_count
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    wall0 = time.perf_counter()
    total_ms = 0.0
    step_ms = []
    encode_ms = []
    out_bytes = 0
    for _ in range(args.steps):
        flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        enc_done[0] = torch.cuda.Event(enable_timing=True) if world > 1 else None
        with torch.cuda.stream(ext):
            e0.record()
        out_bytes = step(True)
        with torch.cuda.stream(ext):
            e1.record()
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        total_ms += step_ms[-1]
        if world > 1:
            encode_ms.append(e0.elapsed_time(enc_done[0]))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    wall_ms = 1e3 * (time.perf_counter() - wall0)
    enc_done[0] = None
    launches = eng.launch_count - launches0
    # per-kernel durations: same workload on the plain single-stream sequence (the band pipeline of
    # the timed loop overlaps kernels of different bands, so events could not bracket one kernel)
    eng.enable_timing(True)
    eng.stage_ms()
    for _ in range(args.steps):
        flush_l2()
        step(False)
    torch.cuda.synchronize()
    stages = eng.stage_ms()
    eng.enable_timing(False)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    mpx_total = world * WIDTH * HEIGHT / 1e6
    value = mpx_total / (ms_per_step / 1e3)

    # ---- the smooth variant of the synthetic input (noise amplitude / 8; fewer symbols) -----
    smooth = None
    if rank == 0:
        d_smooth = torch.empty(n_in, dtype=torch.uint8, device=dev)
        eng.synth_fill(d_smooth.data_ptr(), WIDTH, HEIGHT, bits=8, seed=0, smooth=True)
        sm_ms = []
        sm_bytes = 0
        for i in range(3 + min(args.steps, 5)):
            flush_l2()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(ext):
                e0.record()
            sm_bytes = eng.encode_image_device(d_smooth.data_ptr(), WIDTH, HEIGHT, CHANNELS, sample_fmt=HYD_UINT8,
                                               d_out=d_out.data_ptr(), d_out_cap=cap)
            with torch.cuda.stream(ext):
                e1.record()
            e1.synchronize()
            if i >= 3:
                sm_ms.append(e0.elapsed_time(e1))
        smooth = {"value": WIDTH * HEIGHT / 1e6 / (float(np.mean(sm_ms)) / 1e3), "unit": UNIT, "ms_per_step": float(np.mean(sm_ms)),
                  "bytes_out_per_px": sm_bytes / (WIDTH * HEIGHT), "note": "same image, noise amplitude / 8 (device-resident, 1 GPU)"}
        del d_smooth
        step(False)   # leave the engine's output buffer holding the headline image again
        torch.cuda.synchronize()

    # ---- end to end through the C ABI with host buffers ("e2e") --------------------------------
    lib = eng.lib
    h_in_p = lib.hydb_host_alloc(n_in)
    h_out_p = lib.hydb_host_alloc(cap)
    h_in = np.ctypeslib.as_array(C.cast(h_in_p, C.POINTER(C.c_uint8)), shape=(n_in,))
    h_in[:] = d_in.cpu().numpy()
    e2e_ms = 0.0
    e2e_bytes = 0
    n64 = C.c_uint64(0)
    for i in range(max(args.warmup, 3) + args.steps):
        if world > 1:
            dist.barrier()
        flush_l2()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rc = lib.hydb_encode_image_host(eng._h, h_in_p, WIDTH, HEIGHT, CHANNELS, HYD_UINT8, 0, h_out_p, cap, C.byref(n64))
        dt = 1e3 * (time.perf_counter() - t0)
        if rc != 0:
            raise RuntimeError(f"hydb_encode_image_host failed: {eng.error()}")
        if i >= max(args.warmup, 3):
            e2e_ms += dt
            e2e_bytes = int(n64.value)
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms_per_step = float(t.item()) / args.steps
    e2e_value = mpx_total / (e2e_ms_per_step / 1e3)
    h_out = bytes(np.ctypeslib.as_array(C.cast(h_out_p, C.POINTER(C.c_uint8)), shape=(e2e_bytes,)))
    dev_out = bytes(d_out[:out_bytes].cpu().numpy()) if world == 1 else None

    # ---- the nine-symbol libhydrium API, batched (rank 0, informational) -------------------------
    hyd_api = None
    if rank == 0:
        from hydrium_b200.encoder import encode_cli_loop
        os.environ["HYDRIUM_B200_BATCH"] = "256"
        os.environ["HYDRIUM_B200_DEVICE"] = str(local_rank)
        img = h_in.reshape(HEIGHT, WIDTH, CHANNELS)
        encode_cli_loop(lib, img[:512])
        t0 = time.perf_counter()
        api_out = encode_cli_loop(lib, img)
        api_ms = 1e3 * (time.perf_counter() - t0)
        hyd_api = {"value": WIDTH * HEIGHT / 1e6 / (api_ms / 1e3), "unit": UNIT, "ms": api_ms, "batch_tiles": 256,
                   "identical_to_engine_output": api_out == h_out}

    # ---- one-frame mode, the reference CLI's default (rank 0, informational) ------------------------
    one_frame = None
    if rank == 0:
        os.environ.pop("HYDRIUM_B200_BATCH", None)
        encode_cli_loop(lib, img[:2048, :2048], shift_x=-1, shift_y=-1)   # engine with multi-group slots
        t0 = time.perf_counter()
        of_out = encode_cli_loop(lib, img, shift_x=-1, shift_y=-1)
        of_ms = 1e3 * (time.perf_counter() - t0)
        one_frame = {"value": WIDTH * HEIGHT / 1e6 / (of_ms / 1e3), "unit": UNIT, "ms": of_ms, "bytes": len(of_out),
                     "note": "tile_size_shift -1: one frame of 4 LF groups x 64 groups, nine-symbol API, host buffers"}

    # ---- CPU baseline + parity (rank 0, N = 1 only) ------------------------------------------------
    cpu_baseline = None
    parity = None
    if rank == 0 and world == 1:
        from oracle.pyoracle import have_ref, ref_library
        from hydrium_b200.encoder import encode_cli_loop
        img = h_in.reshape(HEIGHT, WIDTH, CHANNELS)
        if have_ref():
            ref = ref_library("Os")
            t0 = time.perf_counter()
            ref_out = encode_cli_loop(ref, img)
            dt = time.perf_counter() - t0
            cpu_baseline = {"value": WIDTH * HEIGHT / 1e6 / dt, "unit": UNIT, "cores": 1, "kind": "reference",
                            "sample": f"full {WIDTH}x{HEIGHT} image once, single thread, reference built -Os with its own flags"}
        else:
            from oracle.pyoracle import Oracle
            t0 = time.perf_counter()
            ref_out = Oracle().encode_image(img)
            dt = time.perf_counter() - t0
            cpu_baseline = {"value": WIDTH * HEIGHT / 1e6 / dt, "unit": UNIT, "cores": 1, "kind": "port",
                            "sample": f"full {WIDTH}x{HEIGHT} image once, oracle restatement"}
        parity = {"device_path_identical": dev_out == ref_out, "host_path_identical": h_out == ref_out,
                  "bytes": len(ref_out)}
        if have_ref():
            t0 = time.perf_counter()
            ref_of = encode_cli_loop(ref, img, shift_x=-1, shift_y=-1)
            one_frame["reference_ms_one_thread"] = 1e3 * (time.perf_counter() - t0)
            parity["one_frame_identical"] = of_out == ref_of
    lib.hydb_host_free(h_in_p)
    lib.hydb_host_free(h_out_p)

    if rank == 0:
        peak, peak_src = measured_peaks()
        nb = max(stages["batches"], 1.0)
        ans_ms = stages["ans_chain"] / nb
        b_in, b_out = float(n_in), float(out_bytes)
        achieved = (b_in + b_out) / (ans_ms / 1e3) / 1e9 if ans_ms > 0 else 0.0
        traffic = xyb_traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            traffic = tj.get("k_ans_chain_dram_bytes_per_launch")
            xyb_traffic = tj.get("k_xyb_dct_dram_bytes_per_launch")
        per_stage = {k: stages[k] / nb for k in ("xyb_dct_quant", "hf_tokens", "lf_group", "ans_chain", "ans_pack", "gather")}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "ms_best": float(min(step_ms)),
            "ms_median": float(np.median(step_ms)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "images_per_step": world, "tiles_per_gpu": 256,
                       "l2": "flushed between timed steps (512 MB write)", "bytes_in_per_px": 3,
                       "cpu_affinity": affinity,
                       "bytes_out_per_px": out_bytes / (WIDTH * HEIGHT),
                       "timing": "CUDA events on the engine stream, per step, max over ranks",
                       "pipeline": "4 bands of tile rows on separate streams; per-kernel ms from a second, single-stream pass"},
            "roofline": {"bound": "hbm", "kernel": "k_ans_chain", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": b_in + b_out,
                         "kernel_ms": ans_ms,
                         "hbm_read_roofline_frac_whole_step": (b_in / (ms_per_step / 1e3) / 1e9) / peak,
                         "note": "serial rANS chain per tile, bound by in-order issue / dependent latency of one warp "
                                 "(54.6 cycles per symbol, profiles/r01_chain_source_v3.txt); see DESIGN.md"},
            # the stage the north star asks an HBM fraction for: RGB in, int16 coefficients out
            "roofline_xyb_dct_quant": (lambda ms: {
                "bound": "hbm", "kernel": "k_xyb_dct_quant", "achieved": b_in / (ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": b_in / (ms / 1e3) / 1e9 / peak, "traffic": xyb_traffic, "algorithmic_bytes_per_launch": b_in,
                "kernel_ms": ms,
                "note": "FP32 / issue bound (issue slots 82 % busy, profiles/r01b_ncu_summary.json): the format fixes the "
                        "DCT's summation order, ~380 thread instructions per pixel"})(per_stage["xyb_dct_quant"])
            if per_stage["xyb_dct_quant"] > 0 else None,
            "stages_ms": per_stage,
            "xyb_dct_quant_gbs": (b_in / (per_stage["xyb_dct_quant"] / 1e3) / 1e9) if per_stage["xyb_dct_quant"] > 0 else None,
            "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n_in, "d2h_bytes_per_step": e2e_bytes,
                    "ms_per_step": e2e_ms_per_step, "api": "hydb_encode_image_host (C ABI, pinned host buffers)"},
            "gather": ({"how": gather_kind, "ms_encode_rank0": float(np.mean(encode_ms)), "ms_step_rank0": float(np.mean(step_ms)),
                        "note": "rank 0's own encode vs its whole step (waiting for the slowest rank + the NCCL gather)"}
                       if world > 1 else None),
            "smooth_variant": smooth,
            "e2e_hyd_api": hyd_api,
            "e2e_hyd_api_one_frame": one_frame,
            "parity": parity,
            "gpu_launches": int(launches),
            "wall_ms_timed_region": wall_ms,
            "clocks": clocks,
        }
        print(json.dumps(line))
    if pg is not None:
        pg.close()
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    args = ap.parse_args()
    # The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner, for one) write to
    # file descriptor 1 behind Python's back, so everything but our line is sent to stderr.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    try:
        if args.impl == "reference":
            return run_reference(args)
        return run_ours(args)
    finally:
        real_stdout.flush()


if __name__ == "__main__":
    sys.exit(main())
