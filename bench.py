#!/usr/bin/env python
"""bench.py — benchmarks of the B200-native PixelLink / EAST head.

  python bench.py --gpus N --steps K --warmup W [--config 2|3a|3b|4|5]   # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...                 # the reference's CPU path
                                                                          # (numpy/OpenCV oracle port), rank 0

--config selects a BASELINE.json configuration (default 2 = the headline the driver runs):
  2   PixelLink-4s head step = loss fwd+bwd (OHEM 3:1, nets/model.py:204-261) + decode
      (test_pixellink_fast.py:110-202), batch 32 PER GPU at 512x512 (128x128 maps); weak scaling
  3a  PixelLink-2s decode on ICDAR-shaped maps (768x1280 input -> 384x640 maps, scale 2.0 / 1.875),
      batch 64 in total, sharded over the GPUs (strong scaling)
  3b  same on 256x256 maps (512x512 input at 2s)
  4   EAST RBOX head: dice + IoU/angle loss fwd+bwd, restore_rectangle + locality-aware NMS, batch 32 per GPU
      at 512x512 (128x128 maps); NMS time reported separately (sequential fold, not roofline bound)
  5   loss ablation on PixelLink-4s maps at 768x768 (192x192 maps): softmax-CE+OHEM, focal, dice head, batch 256
      over 8 GPUs = 32 per GPU, one NCCL all-reduce of the loss scalars per step

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

NSETS = 6      # rotating input sets: their total size exceeds the 126 MB L2 for every configuration
KEYS = ("pix_logits", "link_logits", "pix_lab", "link_lab")


def _peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _ncu_traffic(kernel_substr):
    """dram__bytes_read + dram__bytes_write per launch of a kernel, from the committed ncu summary
    (profiles/r02_ncu_full_all_kernels.txt, written by tools/ncu_kernels.py from an `ncu --set full` capture)."""
    path = os.path.join(ROOT, "profiles", "r02_ncu_full_all_kernels.txt")
    try:
        vals = []
        for line in open(path):
            if kernel_substr in line:
                f = line.split()
                i = next(k for k, tok in enumerate(f) if tok.replace(".", "", 1).isdigit())
                vals.append((float(f[i + 1]) + float(f[i + 2])) * 1e6)     # columns: us, rd MB, wr MB
        return (sum(vals) / len(vals), os.path.relpath(path, ROOT)) if vals else (None, None)
    except Exception:
        return None, None


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


# ----------------------------------------------------------------------------- workloads
class Workload:
    """One BASELINE configuration: synthetic inputs, the GPU step, the CPU (oracle) step, the roofline kernel."""
    cid = "2"
    metric = ""
    workload = ""
    scaling = "weak"
    H = W = 128
    total_batch = None          # strong scaling: the batch is split over the ranks
    batch_per_gpu = 32
    bytes_per_px = 0            # algorithmic bytes of the whole step
    cpu_s_per_image = 0.03      # one core, for sizing the bounded CPU sample (~20 s)
    collective = False          # a per-step all-reduce of the loss scalars at N > 1

    def batch(self, world):
        return self.batch_per_gpu if self.total_batch is None else max(1, self.total_batch // world)

    def config(self, world):
        """The `config` object of the JSON line: the SAME object in both arms (the driver compares them); what is
        specific to a run (per-rank times, graph mode, collective ...) goes to the line's `run` object."""
        B = self.batch(world)
        return {"workload": self.workload, "config_id": self.cid, "maps": [self.H, self.W],
                "batch_per_gpu": B, "global_batch": B * world, "parallelism": "batch shards, dp%d" % world,
                "l2": "GPU arm: inputs rotate over %d distinct sets, larger than the 126 MB L2 in total" % NSETS}

    # --- host data: list of NSETS dicts of numpy arrays (rank-dependent order, same images on every rank)
    def host_sets(self, B, rank):
        raise NotImplementedError

    def setup(self, dev, B):
        pass

    def step(self, d, out):           # enqueue the step on the current stream; d = device tensors of one set
        raise NotImplementedError

    def results(self, out):           # tensors a caller reads back every step (e2e D2H)
        raise NotImplementedError

    def loss_stats(self, out):        # the 64 loss scalars for the all-reduce (or None)
        return None

    def cpu_step(self, batch, pool):  # the reference's CPU path on a host batch (dict of numpy arrays)
        raise NotImplementedError

    def roofline(self, ctx):          # -> dict (without peak / frac, added by the harness)
        raise NotImplementedError


def _pixellink_sets(cid, B, H, W, rank, keys=KEYS):
    from tensorflow_ocr_b200 import synth
    base = synth.make_batch(int(cid[0]), B, H, W, "C", first_image=0)
    return [{k: np.ascontiguousarray(np.roll(base[k], s + rank, axis=0)) for k in keys} for s in range(NSETS)]


def _cpu_decode_one(args):
    from oracle import decode as D
    pl, ll, scale = args
    lab, boxes, sizes, _ = D.decode_pixellink(pl, ll, scale=scale)
    return len(boxes)


class HeadStep(Workload):
    cid = "2"
    metric = "img/s PixelLink-4s head (loss fwd+bwd+decode) 512^2 b32"
    workload = "PixelLink-4s head step: loss fwd+bwd (OHEM 3:1) + decode, batch 32 at 512x512 (128x128 maps), per GPU"
    bytes_per_px = 184          # loss fwd+bwd 180 B/px + 4 B/px label map (logits shared)
    collective = True
    scale = (4.0, 3.75)

    def host_sets(self, B, rank):
        return _pixellink_sets(self.cid, B, self.H, self.W, rank)

    def setup(self, dev, B):
        from tensorflow_ocr_b200 import head
        self.lcfg, self.dcfg = head.LossConfig(), head.DecodeConfig(max_boxes=128, scale=self.scale)

    def step(self, d, out):
        from tensorflow_ocr_b200 import head
        head.loss_and_decode_raw(d["pix_logits"], d["link_logits"], d["pix_lab"], d["link_lab"], self.lcfg, self.dcfg,
                                 out, want_rects=False)

    def results(self, out):
        return [out["stats"], out["n_boxes"], out["boxes"]]

    def loss_stats(self, out):
        return out["stats"]

    def cpu_step(self, batch, pool):
        from oracle import pixellink_loss as O
        r = O.loss_model(batch["pix_lab"], batch["pix_logits"], batch["link_lab"], batch["link_logits"])
        args = [(batch["pix_logits"][b], batch["link_logits"][b], self.scale) for b in range(len(batch["pix_logits"]))]
        nb = pool.map(_cpu_decode_one, args) if pool is not None else [_cpu_decode_one(a) for a in args]
        return float(r["loss"]), int(sum(nb))

    def roofline(self, ctx):
        return _roofline_loss_main(self, ctx, self.lcfg, step_bytes=self.bytes_per_px)


def _roofline_loss_main(wl, ctx, lcfg, step_bytes):
    """The dominant kernel of the loss: loss_main_kernel (180 B/px).  (1) relaunched alone back to back over the
    rotating (cold) input sets as a replayed CUDA graph, CUDA events around the sequence; (2) bracketed by
    events inside the step; (3) its device-side window (first CTA start .. last CTA end, %globaltimer)."""
    import dataclasses
    import torch
    from tensorflow_ocr_b200 import _lib, head
    lib, dev, B, H, W = ctx["lib"], ctx["dev"], ctx["B"], wl.H, wl.W
    PX = B * H * W
    nbytes = lib.plh_workspace_bytes(_lib.OP_LOSS, B, H, W, 0)
    wss = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(NSETS)]
    routs = [{} for _ in range(NSETS)]
    only = dataclasses.replace(lcfg, main_only=True)
    dev_sets, main_stream = ctx["dev_sets"], ctx["main_stream"]

    def loss_call(i, cfg):
        d = dev_sets[i % NSETS]
        head.pixellink_loss_raw(d["pix_logits"], d["link_logits"], d["pix_lab"], d["link_lab"], cfg, True, False, None,
                                routs[i % NSETS], wss[i % NSETS])

    for i in range(NSETS):
        loss_call(i, lcfg)
    for i in range(2 * NSETS):
        loss_call(i, only)
    torch.cuda.synchronize()
    rg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(rg, stream=main_stream):
        for i in range(NSETS):
            loss_call(i, only)
    torch.cuda.synchronize()
    torch.cuda.set_stream(main_stream)
    reps = max(1, min(ctx["steps"], 4096) // NSETS)
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

    tot, n, win = ctypes.c_float(0), ctypes.c_int(0), ctypes.c_float(0)
    extra = {}
    if ctx.get("step_fn") is not None:
        _lib.check(lib.plh_profile_begin(min(nprof, 4096)), "plh_profile_begin")
        for i in range(min(nprof, 4096)):
            ctx["step_fn"](i)
        _lib.check(lib.plh_profile_kernel_window(ctypes.byref(win)), "plh_profile_kernel_window")
        _lib.check(lib.plh_profile_end(ctypes.byref(tot), ctypes.byref(n)), "plh_profile_end")
        extra["us_per_launch_inside_step_event_bracketed"] = tot.value * 1e3 / max(1, n.value)
        extra["us_device_window_inside_step"] = win.value * 1e3 / max(1, n.value)
    nwin = min(nprof, 600)
    _lib.check(lib.plh_profile_begin(nwin), "plh_profile_begin")
    for i in range(nwin):
        loss_call(i, only)
    _lib.check(lib.plh_profile_kernel_window(ctypes.byref(win)), "plh_profile_kernel_window")
    _lib.check(lib.plh_profile_end(ctypes.byref(tot), ctypes.byref(n)), "plh_profile_end")
    alone_window_us = win.value * 1e3 / max(1, n.value)

    alg = 180 * PX
    traffic, src = _ncu_traffic("loss_main_kernel") if (B, H, W) == (32, 128, 128) else (None, None)
    rf = {"bound": "hbm", "kernel": "loss_main_kernel", "achieved": alg / per_launch_s / 1e9, "unit": "GB/s",
          "traffic": traffic, "algorithmic_bytes_per_launch": alg, "us_per_launch": per_launch_s * 1e6,
          "launches_timed": nprof,
          "method": "kernel relaunched alone back to back (CUDA graph of %d launches over the rotating input sets, "
                    "replayed), CUDA events around the sequence / launches" % NSETS,
          "us_device_window_alone": alone_window_us,
          "window_note": "device window = first CTA start to last CTA end (%globaltimer)",
          "traffic_note": ("ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, read from %s (one `ncu --set "
                           "full` capture); below the algorithmic bytes because part of the 37.7 MB of gradients is "
                           "still dirty in the 126 MB L2 when the kernel ends" % src) if traffic else
                          "no ncu capture committed for this shape"}
    rf.update(extra)
    return rf


class DecodeOnly(Workload):
    scaling = "strong"
    total_batch = 64
    bytes_per_px = 76           # 72 B/px logits + 4 B/px label map
    cpu_s_per_image = 0.14

    def __init__(self, cid):
        self.cid = cid
        if cid == "3a":
            self.H, self.W, self.scale = 384, 640, (2.0, 1.875)
            self.workload = ("PixelLink-2s decode, ICDAR2015-shaped 768x1280 inputs (384x640 maps, scale 2.0 / 1.875), "
                             "batch 64 sharded over the GPUs")
        else:
            self.H, self.W, self.scale = 256, 256, (2.0, 2.0)
            self.workload = "PixelLink-2s decode, 512x512 inputs (256x256 maps, scale 2.0), batch 64 sharded over the GPUs"
        self.metric = "img/s PixelLink-2s decode %dx%d maps b64" % (self.H, self.W)

    def host_sets(self, B, rank):
        return _pixellink_sets(self.cid, B, self.H, self.W, rank, ("pix_logits", "link_logits"))

    def setup(self, dev, B):
        from tensorflow_ocr_b200 import head
        self.dcfg = head.DecodeConfig(max_boxes=512, scale=self.scale)

    def step(self, d, out):
        from tensorflow_ocr_b200 import head
        head.decode_raw(d["pix_logits"], d["link_logits"], self.dcfg, out, want_rects=False)

    def results(self, out):
        return [out["n_boxes"], out["boxes"]]

    def cpu_step(self, batch, pool):
        args = [(batch["pix_logits"][b], batch["link_logits"][b], self.scale) for b in range(len(batch["pix_logits"]))]
        nb = pool.map(_cpu_decode_one, args) if pool is not None else [_cpu_decode_one(a) for a in args]
        return 0.0, int(sum(nb))

    def roofline(self, ctx):
        """decode_tile_cc_kernel<FROM_LOGITS>: thresholds + tile labelling, the one decode kernel that streams the
        72 B/px of logits (the rest of the chain moves ~10 B/px and is latency bound)."""
        import dataclasses
        import torch
        from tensorflow_ocr_b200 import head
        dev_sets, main_stream, B = ctx["dev_sets"], ctx["main_stream"], ctx["B"]
        tile = dataclasses.replace(self.dcfg, phase=4, form="tiled")
        outs = [{} for _ in range(NSETS)]

        def call(i):
            d = dev_sets[i % NSETS]
            head.decode_raw(d["pix_logits"], d["link_logits"], tile, outs[i % NSETS], False)
        for i in range(NSETS):
            call(i)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=main_stream):
            for i in range(NSETS):
                call(i)
        torch.cuda.synchronize()
        torch.cuda.set_stream(main_stream)
        reps = max(1, min(ctx["steps"], 1200) // NSETS)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for _ in range(reps):
            g.replay()
        r1.record()
        torch.cuda.synchronize()
        per = r0.elapsed_time(r1) * 1e-3 / (reps * NSETS)
        alg = 76 * B * self.H * self.W          # 72 B/px read + 2 B/px flags + 4 B/px forest written ~ the 76 B/px of the decode
        return {"bound": "hbm", "kernel": "decode_tile_cc_kernel<FROM_LOGITS>", "achieved": alg / per / 1e9, "unit": "GB/s",
                "traffic": None, "algorithmic_bytes_per_launch": alg, "us_per_launch": per * 1e6,
                "launches_timed": reps * NSETS,
                "method": "kernel relaunched alone back to back (CUDA graph of %d launches over the rotating input "
                          "sets, replayed), CUDA events around the sequence / launches" % NSETS,
                "traffic_note": "no ncu capture committed for this shape"}


class EastHead(Workload):
    cid = "4"
    metric = "img/s EAST RBOX head (dice + IoU/angle loss fwd+bwd, restore_rectangle + locality-aware NMS) 512^2 b32"
    workload = ("EAST RBOX head: dice + IoU/angle loss fwd+bwd (1 score + 5 geometry channels at 1/4 res), restore_rectangle "
                "+ locality-aware NMS of the pixels with score > 0.8, batch 32 at 512x512 (128x128 maps), per GPU")
    bytes_per_px = 76
    cpu_s_per_image = 0.85
    THRESH = 0.8

    def host_sets(self, B, rank):
        from tensorflow_ocr_b200 import synth
        base = synth.make_east_batch(4, B, self.H, self.W, consistent=True)
        sets = []
        for s in range(NSETS):
            d = {k: np.ascontiguousarray(np.roll(v, s + rank, axis=0)) for k, v in base.items()}
            # candidate pixels (score > thresh), row-major per image, as EAST's detect() extracts them on the host
            # (np.argwhere on the score map); origin = (x, y) * 4, geometry = predicted (d0..d3, theta) there
            offs, org, geo, sc = [0], [], [], []
            for b in range(B):
                ys, xs = np.nonzero(d["score_pred"][b, :, :, 0] > self.THRESH)
                org.append(np.stack([xs, ys], 1).astype(np.float32) * 4.0)
                geo.append(d["geo_pred"][b, ys, xs, :])
                sc.append(d["score_pred"][b, ys, xs, 0])
                offs.append(offs[-1] + len(ys))
            d["cand_origin"] = np.concatenate(org).astype(np.float32)
            d["cand_geo"] = np.concatenate(geo).astype(np.float32)
            d["cand_score"] = np.concatenate(sc).astype(np.float64)
            d["cand_offsets"] = np.asarray(offs, np.int32)
            sets.append(d)
        return sets

    def step(self, d, out, nms=True):
        import torch
        from tensorflow_ocr_b200 import head
        outv, gs, gg = head.east_loss_raw(d["score_gt"], d["score_pred"], d["geo_gt"], d["geo_pred"], d["training_mask"])
        out["stats"], out["grad_score"], out["grad_geo"] = outv, gs, gg
        quads, idx = head.restore_rectangle_raw(d["cand_origin"], d["cand_geo"], want_index=True)
        # rows come out theta >= 0 first (datasets/icdar.py:479): back to scan order, score appended (torch index glue)
        n = quads.shape[0]
        polys = out.get("polys")
        if polys is None or polys.shape[0] != n:
            polys = out["polys"] = torch.empty((n, 9), dtype=torch.float64, device=quads.device)
        polys[:, :8].index_copy_(0, idx.long(), quads.view(n, 8))
        polys[:, 8] = d["cand_score"]
        if nms:
            out["nms"], out["n_out"] = head.lanms_raw(polys, d["cand_offsets"], 0.3)

    def results(self, out):
        return [out["stats"], out["n_out"], out["nms"]]

    def cpu_step(self, batch, pool):
        from oracle import east as E
        r = E.east_loss(batch["score_gt"], batch["score_pred"], batch["geo_gt"], batch["geo_pred"], batch["training_mask"])
        nb = 0
        for b in range(len(batch["score_pred"])):
            ys, xs = np.nonzero(batch["score_pred"][b, :, :, 0] > self.THRESH)
            if len(ys) == 0:
                continue
            geo = batch["geo_pred"][b, ys, xs, :]
            quads = E.restore_rectangle_rbox(np.stack([xs, ys], 1).astype(np.float32) * 4.0, geo)
            order = np.concatenate([np.nonzero(geo[:, 4] >= 0)[0], np.nonzero(geo[:, 4] < 0)[0]])
            polys = np.zeros((len(ys), 9))
            polys[order, :8] = quads.reshape(-1, 8)
            polys[:, 8] = batch["score_pred"][b, ys, xs, 0]
            nb += len(E.nms_locality(polys, 0.3))
        return float(r["loss"]), nb

    def roofline(self, ctx):
        """plh_east_loss (reduce + gradient kernels, 76 B/px); restore_rectangle and the NMS are timed beside it."""
        import torch
        from tensorflow_ocr_b200 import head
        dev_sets, main_stream, B = ctx["dev_sets"], ctx["main_stream"], ctx["B"]

        def timed(fn):
            for i in range(NSETS):
                fn(dev_sets[i])
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=main_stream):
                for i in range(NSETS):
                    fn(dev_sets[i])
            torch.cuda.synchronize()
            torch.cuda.set_stream(main_stream)
            reps = max(1, min(ctx["steps"], 600) // NSETS)
            g.replay()
            torch.cuda.synchronize()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record()
            for _ in range(reps):
                g.replay()
            r1.record()
            torch.cuda.synchronize()
            return r0.elapsed_time(r1) * 1e-3 / (reps * NSETS)
        t_loss = timed(lambda d: head.east_loss_raw(d["score_gt"], d["score_pred"], d["geo_gt"], d["geo_pred"], d["training_mask"]))
        t_rest = timed(lambda d: head.restore_rectangle_raw(d["cand_origin"], d["cand_geo"], want_index=True))
        scratch = {}
        t_all = timed(lambda d: self.step(d, scratch, nms=True))
        t_nonms = timed(lambda d: self.step(d, scratch, nms=False))
        alg = 76 * B * self.H * self.W
        return {"bound": "hbm", "kernel": "east_reduce_kernel + east_grad_kernel (plh_east_loss)", "achieved": alg / t_loss / 1e9,
                "unit": "GB/s", "traffic": None, "algorithmic_bytes_per_launch": alg, "us_per_launch": t_loss * 1e6,
                "us_restore_rectangle": t_rest * 1e6, "us_locality_aware_nms": (t_all - t_nonms) * 1e6,
                "candidates_per_step": int(dev_sets[0]["cand_origin"].shape[0]),
                "method": "each op relaunched alone back to back (CUDA graph over the rotating input sets, replayed), "
                          "CUDA events around the sequence / launches; NMS = step with minus step without it",
                "traffic_note": "no ncu capture committed for this shape"}


class LossAblation(Workload):
    cid = "5"
    H = W = 192
    metric = "img/s PixelLink-4s loss ablation (softmax-CE+OHEM, focal, dice head; fwd+bwd each) 768^2 b256/8 GPUs"
    workload = ("loss ablation on PixelLink-4s maps: softmax-CE with OHEM 3:1, focal (alpha 0.25, gamma 2) and the dice head, "
                "fwd+bwd each, batch 256 at 768x768 (192x192 maps) over 8 GPUs = 32 per GPU, NCCL all-reduce of the loss scalars")
    bytes_per_px = 180 + 180 + 184      # CE 180, focal 180, dice head 108 + 4 (mask) + 72
    collective = True
    cpu_s_per_image = 0.11

    def host_sets(self, B, rank):
        sets = _pixellink_sets(self.cid, B, self.H, self.W, rank)
        for d in sets:
            # the dice head takes PROBABILITIES (sigmoid outputs, nets/model_vgg_16.py:129-131) and a training mask
            pl, ll = d["pix_logits"], d["link_logits"].reshape(B, self.H, self.W, 8, 2)
            d["pix_prob"] = (1.0 / (1.0 + np.exp(-(pl[..., 1:2] - pl[..., 0:1])))).astype(np.float32)
            d["link_prob"] = (1.0 / (1.0 + np.exp(-(ll[..., 1] - ll[..., 0])))).astype(np.float32)
            d["training_mask"] = np.ones((B, self.H, self.W, 1), np.float32)
        return sets

    def setup(self, dev, B):
        from tensorflow_ocr_b200 import _lib, head
        self.ce, self.focal = head.LossConfig(), head.LossConfig(term=_lib.TERM_FOCAL)

    def step(self, d, out):
        from tensorflow_ocr_b200 import head
        a = (d["pix_logits"], d["link_logits"], d["pix_lab"], d["link_lab"])
        head.pixellink_loss_raw(*a, self.ce, True, False, None, out.setdefault("ce", {}))
        head.pixellink_loss_raw(*a, self.focal, True, False, None, out.setdefault("focal", {}))
        out["dice"] = head.dice_head_raw(d["pix_lab"], d["pix_prob"], d["link_lab"], d["link_prob"], d["training_mask"])
        out["stats"] = out["ce"]["stats"]

    def results(self, out):
        return [out["ce"]["stats"], out["focal"]["stats"], out["dice"][0]]

    def loss_stats(self, out):
        return out["ce"]["stats"]

    def cpu_step(self, batch, pool):
        from oracle import pixellink_loss as O
        a = (batch["pix_lab"], batch["pix_logits"], batch["link_lab"], batch["link_logits"])
        r = O.loss_model(*a)
        O.loss_model(*a, term="focal")
        O.loss_vgg16_dice(batch["pix_lab"], batch["pix_prob"], batch["link_lab"], batch["link_prob"], batch["training_mask"])
        return float(r["loss"]), 0

    def roofline(self, ctx):
        return _roofline_loss_main(self, ctx, self.ce, step_bytes=self.bytes_per_px)


class LogitProducer(Workload):
    """SURVEY 8f N3 (not a BASELINE configuration): the fusion that turns the backbone's feature maps into the 2 + 16
    logits (nets/pixellink.py:56-67) at the PixelLink-4s shapes of config 2 with VGG-16 widths, batch 8 per GPU."""
    cid = "n3"
    batch_per_gpu = 8
    metric = "img/s PixelLink-4s logit producer (fc7 + conv5_3 + conv4_3 + conv3_3 -> 2+16 logits) 512^2 b8"
    workload = ("PixelLink-4s logit producer: 1x1 fuse convolutions + bilinear x2 unpool + output convolutions over fc7 "
                "(32x32x1024), conv5_3 (32x32x512), conv4_3 (64x64x512), conv3_3 (128x128x256) -> logits 128x128x(2+16), batch 8 per GPU")
    bytes_per_px = (256 + 512 // 4 + (1024 + 512) // 16) * 4 + 72     # activations read per output pixel + the logits written
    cpu_s_per_image = 0.15
    CH = (("fc7", 1024, 4), ("conv5_3", 512, 4), ("conv4_3", 512, 2), ("conv3_3", 256, 1))

    def _params(self):
        rng = np.random.default_rng(31)
        p = {}
        for kind, n, last in (("pixel", 2, "text_predication"), ("link", 16, "link_predication")):
            for (name, K, _), st in zip(self.CH, (6, 5, 4, 3)):
                p["stage_%d_%s_fuse" % (st, kind)] = ((rng.standard_normal((K, n)) / np.sqrt(K)).astype(np.float32),
                                                      (0.1 * rng.standard_normal(n)).astype(np.float32))
            p[last] = ((rng.standard_normal((n, n)) / np.sqrt(n)).astype(np.float32), (0.1 * rng.standard_normal(n)).astype(np.float32))
        return p

    def host_sets(self, B, rank):
        rng = np.random.default_rng(77)
        base = {name: rng.standard_normal((B, self.H // d, self.W // d, K), dtype=np.float32) for name, K, d in self.CH}
        return [{k: np.ascontiguousarray(np.roll(v, s + rank, axis=0)) for k, v in base.items()} for s in range(NSETS)]

    def setup(self, dev, B):
        import torch
        p = self._params()
        cat = lambda a, b: torch.as_tensor(np.concatenate([a, b], -1)).to(dev)
        self.w = {name: cat(p["stage_%d_pixel_fuse" % st][0], p["stage_%d_link_fuse" % st][0]) for (name, _, _), st in zip(self.CH, (6, 5, 4, 3))}
        self.b = {name: cat(p["stage_%d_pixel_fuse" % st][1], p["stage_%d_link_fuse" % st][1]) for (name, _, _), st in zip(self.CH, (6, 5, 4, 3))}
        w_out = np.zeros((18, 18), np.float32)
        w_out[:2, :2], w_out[2:, 2:] = p["text_predication"][0], p["link_predication"][0]
        b_out = np.concatenate([p["text_predication"][1], p["link_predication"][1]])
        self.w_out = torch.as_tensor(w_out).to(dev)
        # no activation on the fuse convolutions: the output matrix goes one level up and into the conv3_3 weights
        self.w3 = (self.w["conv3_3"].double() @ self.w_out.double()).float().contiguous()
        self.b3 = (self.b["conv3_3"].double() @ self.w_out.double() + torch.as_tensor(b_out).to(dev).double()).float().contiguous()

    def step(self, d, out):
        from tensorflow_ocr_b200 import head
        s1 = head.head_fuse_level_raw([(d["fc7"], self.w["fc7"], None, self.b["fc7"], False),
                                       (d["conv5_3"], self.w["conv5_3"], None, self.b["conv5_3"], False)])
        s2 = head.head_fuse_level_raw([(d["conv4_3"], self.w["conv4_3"], None, self.b["conv4_3"], False)], prev=s1,
                                      w_out=self.w_out, logits=False)
        out["s2"] = s2
        out["pix"], out["link"] = head.head_fuse_level_raw([(d["conv3_3"], self.w3, None, self.b3, False)], prev=s2, logits=True)

    def results(self, out):
        return [out["pix"], out["link"]]

    def cpu_step(self, batch, pool):
        from oracle import head_logits as OH
        pix, link = OH.pixellink_layers(batch, self._params())
        return float(np.abs(link).mean()), 0

    def roofline(self, ctx):
        """head_fuse_tc_kernel on the conv3_3 level (the largest: 55 % of the bytes), relaunched alone over the rotating sets."""
        import torch
        from tensorflow_ocr_b200 import head
        dev_sets, B = ctx["dev_sets"], ctx["B"]
        s2 = ctx["outs"][0]["s2"] if "outs" in ctx else None
        if s2 is None:
            o = {}
            self.step(dev_sets[0], o)
            s2 = o["s2"]
        call = lambda i: head.head_fuse_level_raw([(dev_sets[i % NSETS]["conv3_3"], self.w3, None, self.b3, False)], prev=s2, logits=True)
        for i in range(NSETS):
            call(i)
        torch.cuda.synchronize()
        reps = 5 * NSETS
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for i in range(reps):
            call(i)
        r1.record()
        torch.cuda.synchronize()
        per = r0.elapsed_time(r1) * 1e-3 / reps
        alg = B * (self.H * self.W * (256 + 18) + (self.H // 2) * (self.W // 2) * 18) * 4
        return {"bound": "hbm", "kernel": "head_fuse_tc_kernel (conv3_3 level)", "achieved": alg / per / 1e9, "unit": "GB/s",
                "traffic": None, "algorithmic_bytes_per_launch": alg, "us_per_launch": per * 1e6, "launches_timed": reps,
                "method": "the level relaunched alone back to back over the rotating input sets, CUDA events around the sequence / launches",
                "traffic_note": "ncu: profiles/r02_headfuse.txt (batch 32: 546 MB read + 8 MB written to DRAM per launch)"}


def get_workload(cid) -> Workload:
    if cid == "n3":
        return LogitProducer()
    if cid == "2":
        return HeadStep()
    if cid in ("3a", "3b"):
        return DecodeOnly(cid)
    if cid == "4":
        return EastHead()
    if cid == "5":
        return LossAblation()
    raise SystemExit("unknown --config %r" % cid)


# ----------------------------------------------------------------------------- CPU arm (oracle)
def time_cpu(wl, sample_images, steps, warmup, cores):
    import multiprocessing as mp
    batch = wl.host_sets(sample_images, 0)[0]
    pool = mp.get_context("fork").Pool(cores) if cores > 1 else None
    try:
        for _ in range(warmup):
            wl.cpu_step(batch, pool)
        t0 = time.perf_counter()
        for _ in range(steps):
            wl.cpu_step(batch, pool)
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
    wl = get_workload(args.config)
    world = max(1, args.gpus)
    cores = os.cpu_count() or 1
    # bounded sample so that (K + W) steps end within ~2 minutes
    budget_s = 120.0
    sample = int(max(1, min(wl.batch(world), budget_s / max(1, args.steps + args.warmup) / wl.cpu_s_per_image * min(cores, 4))))
    v, dt = time_cpu(wl, sample, args.steps, args.warmup, cores)
    cfg = wl.config(world)
    line = {
        "impl": "reference", "metric": wl.metric, "value": v, "unit": "img/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": wl.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": v, "unit": "img/s", "cores": cores, "kind": "port",
                         "sample": "%d images/step x %d steps of the config-%s workload; numpy loss in one process, "
                                   "per-image decode / NMS in a %d-process pool" % (sample, args.steps, wl.cid, cores)},
        "e2e": {"value": v, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "run": {"note": "reference arm: TF1.4 cannot be installed offline; this is its numpy/OpenCV restatement "
                        "(oracle/) timed on the host cores", "images_per_step": sample},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from tensorflow_ocr_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_note = None
    if world > 1:
        # pin the rank to the CPUs NVML reports as local to its GPU before any pinned host buffer is allocated
        # (first touch then places the staging buffers on the GPU's NUMA node)
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = [64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1]
            if cpus:
                os.sched_setaffinity(0, cpus)
                numa_note = "cpus %d-%d" % (min(cpus), max(cpus))
        except Exception as e:      # not fatal: the box may hide the topology
            numa_note = "affinity not set (%s)" % type(e).__name__
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    wl = get_workload(args.config)
    B = wl.batch(world)
    PX = B * wl.H * wl.W
    wl.setup(dev, B)

    sampler = ClockSampler(local)
    sampler.start()

    # ---- synthetic inputs (family C), NSETS rotating sets.  Every rank draws the SAME images
    # (in a rank-dependent order): the step time is data dependent (number and size of components, selection
    # rounds), and with rank-specific images the per-step collective throttles every rank to the slowest one.
    host_np = wl.host_sets(B, rank)
    host_sets = [{k: torch.from_numpy(v).pin_memory() for k, v in hs.items()} for hs in host_np]
    dev_sets = [{k: v.to(dev) for k, v in hs.items()} for hs in host_sets]
    outs = [{} for _ in range(NSETS)]
    in_bytes = sum(v.numel() * v.element_size() for v in host_sets[0].values())

    # Everything below runs on one dedicated stream (plus the library wrapper's auxiliary stream for the concurrent
    # decode branch), so that the cached workspaces are the same in direct and graph mode.
    main_stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(main_stream)

    # training mode, N > 1: the only exchange of the path is the sum of the towers' loss scalars (SURVEY.md section
    # 8e) — ONE all-reduce of 64 floats PER STEP.  It is enqueued inside the step (so it is captured into the CUDA
    # graph: no host work per step) on a side branch that the next use of its buffer waits for, off the
    # critical path of the kernels.
    use_coll = world > 1 and wl.collective
    side_stream = torch.cuda.Stream(dev) if use_coll else None
    red_bufs = [torch.zeros(_lib.STATS_FLOATS, dtype=torch.float32, device=dev) for _ in range(NSETS)] if use_coll else None

    copy_done = [torch.cuda.Event() for _ in range(NSETS)] if use_coll else None

    def step(i, join=True, direct=False):
        """join: the side branch rejoins the main stream at the end of this step (needed at the end of a captured
        graph; inside the round graph only the last step joins, so no kernel ever waits for a collective).
        direct (no graph): instead of joining, the main stream waits, one round later, for the snapshot copy that
        read the stats buffer it is about to overwrite."""
        s = i % NSETS
        if use_coll and direct:
            main_stream.wait_event(copy_done[s])
        wl.step(dev_sets[s], outs[s])
        if use_coll:
            side_stream.wait_stream(main_stream)
            with torch.cuda.stream(side_stream):
                red_bufs[s].copy_(wl.loss_stats(outs[s])[:_lib.STATS_FLOATS], non_blocking=True)
                if direct:
                    copy_done[s].record(side_stream)
                dist.all_reduce(red_bufs[s], op=dist.ReduceOp.SUM)
            if join and not direct:
                main_stream.wait_stream(side_stream)

    # first calls: allocate outputs / workspaces, set kernel attributes
    for i in range(NSETS):
        step(i)
    torch.cuda.synchronize()

    # ---- one CUDA graph per input set, and one graph holding a whole round of NSETS consecutive steps (a graph
    # launch costs the host ~10-20 us and the device a ~2.5 us gap, a visible fraction of a ~60 us step)
    # (with the per-step collective a round is two passes over the sets: the side branch has to rejoin at the end
    # of a captured graph, and that one wait is then shared by 12 steps)
    ROUND = NSETS * (2 if use_coll else 1)
    graphs, round_graph = None, None
    if not args.no_graphs:
        graphs = []
        for i in range(NSETS):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=main_stream):
                step(i)
            graphs.append(g)
        torch.cuda.synchronize()
        torch.cuda.set_stream(main_stream)
        if not os.environ.get("BENCH_STEP_GRAPHS"):
            round_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(round_graph, stream=main_stream):
                for i in range(ROUND):
                    step(i, join=(i == ROUND - 1))
            torch.cuda.synchronize()
            torch.cuda.set_stream(main_stream)

    # The timed run is `steps` consecutive steps.  Short runs (the driver's 20 steps) are captured WHOLE into one
    # graph: a graph launch costs the device a ~2.5 us gap, and with the per-step collective a graph has to wait for
    # its last all-reduce before it ends — once per run instead of once per round.  Long runs replay round graphs
    # and one tail graph for the remainder.
    run_graph, tail_graph, n_tail = None, None, 0
    if round_graph is not None and args.steps <= 96:
        run_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(run_graph, stream=main_stream):
            for i in range(args.steps):
                step(i, join=(i == args.steps - 1))
        torch.cuda.synchronize()
        torch.cuda.set_stream(main_stream)
    elif round_graph is not None and use_coll and args.steps % ROUND > 1:
        n_tail = args.steps % ROUND
        t0 = args.steps - n_tail
        tail_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(tail_graph, stream=main_stream):
            for i in range(t0, t0 + n_tail):
                step(i, join=(i == t0 + n_tail - 1))
        torch.cuda.synchronize()
        torch.cuda.set_stream(main_stream)

    def run_steps(first, n):
        if run_graph is not None and first == 0 and n == args.steps:
            run_graph.replay()
            return
        i = first
        while i < first + n:
            if round_graph is not None and i % NSETS == 0 and i + ROUND <= first + n:
                round_graph.replay()
                i += ROUND
            elif tail_graph is not None and first == 0 and n == args.steps and i == args.steps - n_tail:
                tail_graph.replay()
                i += n_tail
            else:
                if graphs is not None:
                    graphs[i % NSETS].replay()
                else:
                    step(i, direct=True)
                i += 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # kernels per step (counted once on a direct run: graph replays do not go through the library's counter)
    c0 = lib.plh_launch_count()
    step(0)
    torch.cuda.synchronize()
    launches_per_step = int(lib.plh_launch_count() - c0)

    run_steps(0, args.warmup)
    if run_graph is not None:        # one untimed replay of the whole-run graph (first launch of a graph is slower)
        run_graph.replay()
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

    # ---- the same K steps with TWO batches in flight: consecutive steps are independent batches, and a step is a
    # latency-bound chain of small kernels that leaves most of the machine idle, so a caller that has the next
    # batch ready (inference, or several towers per GPU) can overlap two of them.  Reported beside `value`, never
    # instead of it: `value` keeps one batch in flight (a training loop cannot start step i+1 before step i ends).
    two_in_flight = None
    if wl.cid == "2" and not args.no_graphs and not use_coll:
        lanes = [torch.cuda.Stream(dev) for _ in range(2)]
        for f, ls in enumerate(lanes):             # workspaces / aux streams of the lanes are created outside capture
            ls.wait_stream(main_stream)
            with torch.cuda.stream(ls):
                for i in range(f, NSETS, 2):
                    wl.step(dev_sets[i], outs[i])
            main_stream.wait_stream(ls)
        torch.cuda.synchronize()
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g2, stream=main_stream):
            for f, ls in enumerate(lanes):
                ls.wait_stream(main_stream)
                with torch.cuda.stream(ls):
                    for i in range(f, NSETS, 2):
                        wl.step(dev_sets[i], outs[i])
            for ls in lanes:
                main_stream.wait_stream(ls)
        torch.cuda.synchronize()
        torch.cuda.set_stream(main_stream)
        reps = max(1, args.steps // NSETS)
        for _ in range(3):
            g2.replay()
        barrier()
        e0.record()
        for _ in range(reps):
            g2.replay()
        e1.record()
        barrier()
        t2 = e0.elapsed_time(e1) * 1e-3 / (reps * NSETS)
        two_in_flight = {"value": B / t2, "unit": "img/s", "us_per_step": t2 * 1e6, "steps": reps * NSETS,
                         "note": "two independent batches in flight on two stream pairs (throughput of independent "
                                 "batches; not the latency of a step)"}

    # ---- roofline of the dominant kernel
    roofline = None
    if rank == 0:
        ctx = {"lib": lib, "dev": dev, "B": B, "dev_sets": dev_sets, "main_stream": main_stream, "steps": args.steps,
               "step_fn": (lambda i: wl.step(dev_sets[i % NSETS], outs[i % NSETS])) if wl.cid == "2" else None}
        roofline = wl.roofline(ctx)
        peak, peak_src = _peaks()
        roofline["peak"], roofline["peak_source"] = peak, peak_src
        roofline["frac"] = roofline["achieved"] / peak
        roofline["frac_of_nominal_8TBs"] = roofline["achieved"] / 8000.0
        roofline["whole_step_GBs"] = wl.bytes_per_px * PX / (ms_total * 1e-3 / args.steps) / 1e9
    if world > 1:
        dist.barrier()

    # ---- e2e: host buffers in, host results out, copies inside the timed region
    step(0)
    torch.cuda.synchronize()
    res0 = wl.results(outs[0])
    hres = [[torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in res0] for _ in range(2)]
    # double-buffered staging: the H2D copy of step i+1 (copy stream) overlaps the head of step i
    stages = [{k: torch.empty_like(v) for k, v in dev_sets[0].items()} for _ in range(2)]
    eouts = [{}, {}]
    copy_stream = torch.cuda.Stream(dev)
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def e2e_copy(i):
        j = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[j])          # the step i-2 has finished reading this buffer
            hs = host_sets[i % NSETS]
            for k, v in hs.items():
                stages[j][k].copy_(v, non_blocking=True)
            copied[j].record(copy_stream)

    def e2e_step(i, last):
        j = i % 2
        if not last:
            e2e_copy(i + 1)
        main_stream.wait_event(copied[j])
        wl.step(stages[j], eouts[j])
        consumed[j].record(main_stream)
        for h, t in zip(hres[j], wl.results(eouts[j])):
            h.copy_(t, non_blocking=True)

    e2e_steps = max(3, min(args.steps, 200))
    for j in range(2):
        consumed[j].record(main_stream)
    e2e_copy(0)
    for i in range(3):
        e2e_step(i, i == 2)
    barrier()
    e0.record()
    e2e_copy(3)      # the copy of the first timed step is issued inside the region
    for i in range(3, 3 + e2e_steps):
        e2e_step(i, i == 2 + e2e_steps)
    e1.record()
    barrier()
    ems = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    e2e_ms = float(ems.item())
    e2e_value = world * B * e2e_steps / (e2e_ms * 1e-3)
    d2h = sum(t.numel() * t.element_size() for t in hres[0])
    e2e = {"value": e2e_value, "unit": "img/s", "h2d_bytes_per_step": int(in_bytes), "d2h_bytes_per_step": int(d2h),
           "steps": e2e_steps, "aggregate_h2d_GBs": world * in_bytes * e2e_steps / (e2e_ms * 1e-3) / 1e9,
           "note": "pinned host inputs -> device (double-buffered, copy stream overlaps the previous step), the step, "
                   "its results (loss scalars, box lists) -> pinned host; bound by the host->device copy "
                   "(aggregate_h2d_GBs over all ranks: every GPU of this box hangs off the same host memory)"}

    # ---- second e2e figure (config 2): the reference-named API with numpy arrays, gradients returned to the host
    if wl.cid == "2" and rank == 0 and not args.no_api_e2e:
        from tensorflow_ocr_b200.nets import model as M
        hn = host_np[0]
        for _ in range(2):
            r = M.loss_with_stats(hn["pix_lab"], hn["pix_logits"], hn["link_lab"], hn["link_logits"], None, want_mask=False)
        n_api = 8
        t0 = time.perf_counter()
        for i in range(n_api):
            hn = host_np[i % NSETS]
            r = M.loss_with_stats(hn["pix_lab"], hn["pix_logits"], hn["link_lab"], hn["link_logits"], None, want_mask=False)
        dt = time.perf_counter() - t0
        e2e["api_numpy"] = {"value": B * n_api / dt, "unit": "img/s", "steps": n_api,
                            "call": "nets.model.loss_with_stats(numpy...) -> loss scalars + grad_pixel + grad_link as numpy",
                            "h2d_bytes_per_step": int(in_bytes),
                            "d2h_bytes_per_step": int(sum(v.nbytes for v in r.values())),
                            "note": "py_func-style host round trip through pageable numpy arrays, loss only (no decode), "
                                    "synchronous; the 37.7 MB of gradients come back to the host every call"}
    if world > 1:
        dist.barrier()

    clocks = sampler.result()
    launches = torch.tensor([launches_per_step * args.steps], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(launches)

    if rank == 0:
        cfg = wl.config(world)
        run = {}
        run.update({
            "input_sets_MB": round(NSETS * in_bytes / 1e6, 1),
            "data_per_rank": "the same synthetic images on every rank, order rolled by rank (fixed work per GPU)",
            "cuda_graphs": graphs is not None,
            "us_per_step_per_rank": [round(t / args.steps * 1e3, 2) for t in per_rank_ms],
            "host_enqueue_us_per_step": round(host_us_per_step, 2),
            "steps_per_graph_launch": args.steps if run_graph is not None else (ROUND if round_graph is not None else 1),
            "launches_per_step": launches_per_step,
            "cpu_affinity_rank0": numa_note,
            "two_batches_in_flight": two_in_flight,
            "collective": ("NCCL all-reduce of the 64 loss scalars EVERY step, captured inside the step's CUDA graph "
                           "(side branch)") if use_coll else "none"})
        line = {
            "metric": wl.metric, "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": wl.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches.item()), "roofline": roofline,
            "run": run,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = 1
            sample = max(1, min(8, B))
            nsteps = max(1, int(20.0 / (wl.cpu_s_per_image * sample)))
            v, dt = time_cpu(wl, sample, nsteps, 1, cores)
            line["cpu_baseline"] = {"value": v, "unit": "img/s", "cores": cores, "kind": "port",
                                    "sample": "%d images x %d steps of the same workload (numpy/OpenCV oracle port, "
                                              "one process, %.1f s)" % (sample, nsteps, dt)}
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tear down in order: graphs that hold captured NCCL work first, then the process group.  A watchdog ends
        # the process if the communicator teardown does not return (seen with captured collectives): the line above
        # is already out.
        sys.stdout.flush()
        killer = threading.Timer(30.0, lambda: os._exit(0))
        killer.daemon = True
        killer.start()
        del round_graph, graphs, tail_graph, run_graph
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        killer.cancel()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="2", choices=["2", "3a", "3b", "4", "5", "n3"])
    ap.add_argument("--no-graphs", action="store_true", help="direct launches instead of CUDA graphs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-api-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
