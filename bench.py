#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native PixelLink head.

Metric (BASELINE.json): images/s of the PixelLink-4s head step = loss fwd+bwd (OHEM 3:1,
nets/model.py:204-261) + inference decode (test_pixellink_fast.py:110-202) on 512x512
inputs (128x128 maps, 2 pixel + 16 link channels), batch 32 PER GPU (weak scaling: images
shard by batch).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path
                                                           # (numpy/OpenCV oracle port), rank 0

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B, H, W = 32, 128, 128                     # BASELINE.json configs[1]: batch 32 at 512x512 -> 128x128 maps
PX = B * H * W
BYTES_LOSS, BYTES_DECODE_EXTRA = 180, 4    # loss fwd+bwd 180 B/px; fused decode adds 4 B/px labels
METRIC = "img/s PixelLink-4s head (loss fwd+bwd+decode) 512^2 b32"
WORKLOAD = "PixelLink-4s head step: loss fwd+bwd (OHEM 3:1) + decode, batch 32 at 512x512 (128x128 maps), per GPU"
CONFIG_ID = 2
NSETS = 6                                  # rotating input sets: 6 x (56.6 MB in + 37.7 MB grads) = 566 MB >> 126 MB L2
REDUCE_EVERY = int(os.environ.get("BENCH_REDUCE_EVERY", "10"))  # N > 1: loss all-reduce cadence (multigpu_train.py:179)
REDUCE_ROUNDS = max(1, (REDUCE_EVERY + NSETS - 1) // NSETS)      # ... in whole rounds when a graph launch replays a round


def _peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the GPU is under load."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz, self.ok = None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def result(self):
        self.stop_flag = True
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------- CPU arm (oracle)
def _cpu_decode_one(args):
    from oracle import decode as D
    pl, ll = args
    lab, boxes, sizes, _ = D.decode_pixellink(pl, ll)
    return len(boxes)


def cpu_head_step(batch, pool):
    """The reference's CPU path for one head step on `batch` (numpy/OpenCV restatement):
    loss fwd+bwd in this process (numpy, vectorised over the batch), decode per image in `pool`."""
    from oracle import pixellink_loss as O
    r = O.loss_model(batch["pix_lab"], batch["pix_logits"], batch["link_lab"], batch["link_logits"])
    n = batch["pix_logits"].shape[0]
    args = [(batch["pix_logits"][b], batch["link_logits"][b]) for b in range(n)]
    nb = pool.map(_cpu_decode_one, args) if pool is not None else [_cpu_decode_one(a) for a in args]
    return float(r["loss"]), int(sum(nb))


def time_cpu(sample_images, steps, warmup, cores):
    import multiprocessing as mp
    from tensorflow_ocr_b200 import synth
    batch = synth.make_batch(CONFIG_ID, sample_images, H, W, "C")
    pool = mp.get_context("fork").Pool(cores) if cores > 1 else None
    try:
        for _ in range(warmup):
            cpu_head_step(batch, pool)
        t0 = time.perf_counter()
        for _ in range(steps):
            cpu_head_step(batch, pool)
        dt = time.perf_counter() - t0
    finally:
        if pool is not None:
            pool.close()
            pool.join()
    return sample_images * steps / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # bounded sample so that (K + W) steps end within ~2 minutes: ~0.12 s/image single core
    budget_s, per_img = 120.0, 0.12
    sample = int(max(1, min(B, budget_s / max(1, args.steps + args.warmup) / per_img * min(cores, 4))))
    v, dt = time_cpu(sample, args.steps, args.warmup, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "img/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "maps": [H, W], "batch_per_step": sample,
                   "note": "reference's TF1.4 cannot be installed offline; this is its numpy/OpenCV restatement "
                           "(oracle/) timed on the host cores"},
        "cpu_baseline": {"value": v, "unit": "img/s", "cores": cores, "kind": "port",
                         "sample": "%d images/step x %d steps of the config-2 workload; loss in one numpy process, "
                                   "decode in a %d-process pool" % (sample, args.steps, cores)},
        "e2e": {"value": v, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from tensorflow_ocr_b200 import _lib, head, synth
    from tensorflow_ocr_b200 import dist as pdist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the single JSON line
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    lcfg, dcfg = head.LossConfig(), head.DecodeConfig(max_boxes=128)

    sampler = ClockSampler(local)
    sampler.start()

    # ---- synthetic inputs (family C), NSETS rotating sets.  Weak scaling = the work per GPU
    # is fixed: every rank draws the SAME 32 synthetic images (in a rank-dependent order).  The step time is
    # data dependent (number and size of components, selection rounds); with rank-specific images
    # (first_image=rank*B) rank 1's set measured 6 us/step heavier than rank 0's on the same GPU, and the
    # loss all-reduce then throttles every rank to the slowest one — that was the whole N>1 "overhead".
    keys = ("pix_logits", "link_logits", "pix_lab", "link_lab")
    base = synth.make_batch(CONFIG_ID, B, H, W, "C", first_image=0)
    host_sets, dev_sets = [], []
    for s in range(NSETS):
        hs = {k: torch.from_numpy(np.ascontiguousarray(np.roll(base[k], s + rank, axis=0))).pin_memory() for k in keys}
        host_sets.append(hs)
        dev_sets.append({k: v.to(dev) for k, v in hs.items()})
    outs = [{} for _ in range(NSETS)]

    def step(i):
        d = dev_sets[i % NSETS]
        head.loss_and_decode_raw(d["pix_logits"], d["link_logits"], d["pix_lab"], d["link_lab"], lcfg, dcfg,
                                 outs[i % NSETS], want_rects=False)

    # Everything below runs on one dedicated stream (plus the library wrapper's auxiliary stream for
    # the concurrent decode branch), so that the cached workspaces are the same in direct and graph mode.
    main_stream = torch.cuda.Stream(dev, priority=int(os.environ.get('BENCH_MAIN_PRIO', '0')))
    torch.cuda.set_stream(main_stream)

    # first calls: allocate outputs / workspaces, set kernel attributes
    for i in range(NSETS):
        step(i)
    torch.cuda.synchronize()

    # ---- one CUDA graph per input set (the step is 7 small launches on two branches)
    graphs = None
    if not args.no_graphs:
        graphs = []
        for i in range(NSETS):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=main_stream):
                step(i)
            graphs.append(g)
        torch.cuda.synchronize()
        torch.cuda.set_stream(main_stream)
    # ... and one graph holding a whole round of NSETS consecutive steps: a graph launch costs the host
    # ~10-20 us and the device a ~2.5 us gap, which is a visible fraction of a 65 us step (and makes the
    # loop host-bound as soon as a process group's helper threads compete for the interpreter).
    # Steps that do not fill a round use the per-set graphs.
    round_graph = None
    if graphs is not None and not os.environ.get("BENCH_STEP_GRAPHS"):
        round_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(round_graph, stream=main_stream):
            for i in range(NSETS):
                step(i)
        torch.cuda.synchronize()
        torch.cuda.set_stream(main_stream)

    launches_per_step = [0]
    # training mode, N > 1: the only exchange of the path is the tower-summed loss for reporting, which the
    # reference produces every 10 steps (multigpu_train.py:179-183).  Same cadence here: one tiny NCCL
    # all-reduce of the loss scalars every REDUCE_EVERY steps on a side stream; the
    # compute stream only waits for the all-reduce that last read the stats buffer it is about to
    # overwrite.  (Issuing it every step costs no device time either, but its ~40 us of host-side launch
    # work per step makes a 66 us step host-bound.)
    side_stream = torch.cuda.Stream(dev) if world > 1 else None
    reducers = [pdist.LossStatsReducer(dev, stream=side_stream) for _ in range(NSETS)] if world > 1 else None

    def run_step(i):
        if reducers is not None and reducers[i % NSETS].pending:
            main_stream.wait_event(reducers[i % NSETS].event)
            reducers[i % NSETS].pending = False
        if graphs is not None:
            graphs[i % NSETS].replay()
        else:
            step(i)
        if reducers is not None and i % REDUCE_EVERY == 0:
            reducers[i % NSETS].submit(outs[i % NSETS]["stats"])

    def run_round(i):
        """NSETS steps (i .. i+NSETS-1, i a multiple of NSETS) as one graph launch."""
        if reducers is not None:
            for r in reducers:
                if r.pending:
                    main_stream.wait_event(r.event)
                    r.pending = False
        round_graph.replay()
        if reducers is not None and (i // NSETS) % REDUCE_ROUNDS == 0:
            reducers[NSETS - 1].submit(outs[NSETS - 1]["stats"])

    def run_steps(first, n):
        i = first
        while i < first + n:
            if round_graph is not None and i % NSETS == 0 and i + NSETS <= first + n:
                run_round(i)
                i += NSETS
            else:
                run_step(i)
                i += 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # kernels per step (counted once on a direct run: graph replays do not go through the library's counter)
    c0 = lib.plh_launch_count()
    step(0)
    torch.cuda.synchronize()
    launches_per_step[0] = int(lib.plh_launch_count() - c0)

    run_steps(0, args.warmup)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    host_t0 = time.perf_counter()
    run_steps(0, args.steps)
    host_us_per_step = (time.perf_counter() - host_t0) / args.steps * 1e6
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    per_rank_ms = [float(ms.item())]
    if world > 1:
        allms = [torch.zeros_like(ms) for _ in range(world)]
        dist.all_gather(allms, ms)
        per_rank_ms = [float(t.item()) for t in allms]
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- roofline of the dominant kernel (loss_main_kernel)
    # (1) the kernel alone: every input set's workspace is prepared by one full loss call, then the main
    #     pass is relaunched back to back over the rotating sets (> L2), CUDA events around the whole
    #     sequence on the launching stream, duration = elapsed / launches;
    # (2) the same kernel inside the step: events recorded by the library around that launch
    #     (plh_profile_begin/end); this one also contains the launch latency of an event-bracketed kernel.
    roofline = None
    if rank == 0:
        nbytes = lib.plh_workspace_bytes(_lib.OP_LOSS, B, H, W, 0)
        wss = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(NSETS)]
        routs = [{} for _ in range(NSETS)]
        only = head.LossConfig(main_only=True)

        def loss_call(i, cfg):
            d = dev_sets[i % NSETS]
            head.pixellink_loss_raw(d["pix_logits"], d["link_logits"], d["pix_lab"], d["link_lab"], cfg,
                                    not os.environ.get("BENCH_RF_NOGRAD"), False, None, routs[i % NSETS],
                                    wss[i % NSETS])

        for i in range(NSETS):
            loss_call(i, lcfg)
        for i in range(2 * NSETS):
            loss_call(i, only)
        torch.cuda.synchronize()
        # one graph holding the NSETS relaunches (a Python call per launch would be host-bound at ~20 us)
        rg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(rg, stream=main_stream):
            for i in range(NSETS):
                loss_call(i, only)
        torch.cuda.synchronize()
        torch.cuda.set_stream(main_stream)
        reps = max(1, min(args.steps, 4096) // NSETS)
        nprof = reps * NSETS
        for _ in range(3):
            rg.replay()
        torch.cuda.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for _ in range(reps):
            rg.replay()
        r1.record()
        torch.cuda.synchronize()
        per_launch_s = r0.elapsed_time(r1) * 1e-3 / nprof
        ref_loss = routs[0]["stats"][0].item()

        _lib.check(lib.plh_profile_begin(min(nprof, 4096)), "plh_profile_begin")
        for i in range(min(nprof, 4096)):
            step(i)
        tot, n, win = ctypes.c_float(0), ctypes.c_int(0), ctypes.c_float(0)
        _lib.check(lib.plh_profile_kernel_window(ctypes.byref(win)), "plh_profile_kernel_window")
        _lib.check(lib.plh_profile_end(ctypes.byref(tot), ctypes.byref(n)), "plh_profile_end")
        in_step_us = tot.value * 1e3 / max(1, n.value)
        in_step_window_us = win.value * 1e3 / max(1, n.value)
        # (3) device-side window of the kernel alone (first CTA start .. last CTA end, %globaltimer):
        #     what is left of (1) after subtracting it is launch + drain latency between dependent launches
        nwin = min(nprof, 600)
        _lib.check(lib.plh_profile_begin(nwin), "plh_profile_begin")
        for i in range(nwin):
            loss_call(i, only)
        _lib.check(lib.plh_profile_kernel_window(ctypes.byref(win)), "plh_profile_kernel_window")
        _lib.check(lib.plh_profile_end(ctypes.byref(tot), ctypes.byref(n)), "plh_profile_end")
        alone_window_us = win.value * 1e3 / max(1, n.value)
        assert abs(outs[0]["stats"][0].item() - ref_loss) <= 1e-6 * abs(ref_loss)   # main-only reruns compute the same loss

        peak, peak_src = _peaks()
        alg = BYTES_LOSS * PX
        ach = alg / per_launch_s / 1e9
        roofline = {"bound": "hbm", "kernel": "loss_main_kernel", "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": 61.0e6, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg, "us_per_launch": per_launch_s * 1e6,
                    "launches_timed": nprof,
                    "method": "kernel relaunched alone back to back (CUDA graph of %d launches over the rotating "
                              "input sets, replayed), CUDA events around the sequence / launches" % NSETS,
                    "us_per_launch_inside_step_event_bracketed": in_step_us,
                    "us_device_window_alone": alone_window_us,
                    "us_device_window_inside_step": in_step_window_us,
                    "frac_device_window_alone": alg / (alone_window_us * 1e-6) / 1e9 / peak,
                    "window_note": "device window = first CTA start to last CTA end (%globaltimer); us_per_launch minus "
                                   "it is the launch/drain latency between dependent launches, which the step hides "
                                   "behind the preceding kernel (programmatic dependent launch)",
                    "traffic_note": "ncu dram__bytes_read+write per launch (profiles/): 57.2 MB read + 3.9 MB written "
                                    "to DRAM; the 37.7 MB of gradients are still dirty in the 126 MB L2 at kernel end",
                    "frac_of_nominal_8TBs": ach / 8000.0,
                    "whole_step_GBs": (BYTES_LOSS + BYTES_DECODE_EXTRA) * PX / (ms_total * 1e-3 / args.steps) / 1e9}
    if world > 1:
        dist.barrier()

    # ---- e2e: host buffers in, host results out, copies inside the timed region
    hres = [{"stats": torch.empty(_lib.STATS_FLOATS + B, dtype=torch.float32).pin_memory(),
             "n_boxes": torch.empty(B, dtype=torch.int32).pin_memory(),
             "boxes": torch.empty((B, dcfg.max_boxes, 4, 2), dtype=torch.int32).pin_memory()} for _ in range(2)]
    # double-buffered staging: the H2D copy of step i+1 (copy stream) overlaps the head of step i
    stages = [{k: torch.empty_like(dev_sets[0][k]) for k in keys} for _ in range(2)]
    eouts = [{}, {}]
    copy_stream = torch.cuda.Stream(dev)
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def e2e_copy(i):
        j = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[j])          # the head of step i-2 has finished reading this buffer
            hs = host_sets[i % NSETS]
            for k in keys:
                stages[j][k].copy_(hs[k], non_blocking=True)
            copied[j].record(copy_stream)

    def e2e_step(i, last):
        j = i % 2
        if not last:
            e2e_copy(i + 1)
        main_stream.wait_event(copied[j])
        st = stages[j]
        head.loss_and_decode_raw(st["pix_logits"], st["link_logits"], st["pix_lab"], st["link_lab"], lcfg, dcfg,
                                 eouts[j], want_rects=False)
        consumed[j].record(main_stream)
        hr = hres[j]
        for k in ("stats", "n_boxes", "boxes"):
            hr[k].copy_(eouts[j][k], non_blocking=True)

    e2e_steps = max(3, min(args.steps, 200))
    for j in range(2):
        consumed[j].record(main_stream)
    e2e_copy(0)
    for i in range(3):
        e2e_step(i, i == 2)
    barrier()
    # timed: the copy of step 0 is issued inside the region
    e0.record()
    e2e_copy(3)
    for i in range(3, 3 + e2e_steps):
        e2e_step(i, i == 2 + e2e_steps)
    e1.record()
    barrier()
    ems = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / (float(ems.item()) * 1e-3)
    h2d = sum(host_sets[0][k].numel() * host_sets[0][k].element_size() for k in keys)
    d2h = sum(v.numel() * v.element_size() for v in hres[0].values())
    assert np.isfinite(hres[(2 + e2e_steps) % 2]["stats"][0].item())

    clocks = sampler.result()
    launches = torch.tensor([launches_per_step[0] * args.steps], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(launches)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "maps": [H, W], "batch_per_gpu": B, "global_batch": B * world,
                       "parallelism": "batch shards, dp%d" % world,
                       "l2": "inputs rotate over %d distinct sets (%.0f MB) > 126 MB L2" % (
                           NSETS, NSETS * (108 + 72) * PX / 1e6),
                       "data_per_rank": "the same 32 synthetic images on every rank, order rolled by rank "
                                        "(weak scaling: fixed work per GPU)",
                       "cuda_graphs": graphs is not None,
                       "us_per_step_per_rank": [round(t / args.steps * 1e3, 2) for t in per_rank_ms],
                       "host_enqueue_us_per_step": round(host_us_per_step, 2),
                       "steps_per_graph_launch": NSETS if round_graph is not None else 1,
                       "collective": ("async NCCL all-reduce of 64 loss scalars every %d steps (about the reference's reporting "
                                      "cadence, multigpu_train.py:179)" % (REDUCE_ROUNDS * NSETS if round_graph is not None
                                                                                else REDUCE_EVERY)) if world > 1 else "none"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "img/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                    "note": "pinned host logits+labels -> device (double-buffered, copy stream overlaps the previous step), head step, loss stats + boxes -> pinned host"},
            "gpu_launches": int(launches.item()),
            "roofline": roofline,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = 1
            v, dt = time_cpu(8, 45, 1, cores)
            line["cpu_baseline"] = {"value": v, "unit": "img/s", "cores": cores, "kind": "port",
                                    "sample": "8 images x 45 steps of the same workload (numpy/OpenCV oracle port, "
                                              "one process, %.1f s)" % dt}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-graphs", action="store_true", help="direct launches instead of CUDA graphs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
