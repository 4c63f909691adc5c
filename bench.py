#!/usr/bin/env python
"""Benchmark of the SE_UNet hot path on B200 (BASELINE.json metric: CT voxels/s, sliding-window inference).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU

One "step" = sliding-window inference of ONE synthetic 512x512x400 CT volume (BASELINE config 2:
294 windows of 128^3 at stride 64, prediction.py:65-111) per GPU: HU windowing, 294 SE_UNet forwards (batched),
sigmoid, overlap mean, 0.5 threshold.  `value` = volume voxels / s with the stored CT volume already resident in
HBM; `e2e` = the same through the public API (SlidingWindowPredictor.predict) with a pinned HOST volume in and
the HOST mask out.  N > 1: one process per GPU (torchrun), every rank segments its own volume, no data-path
collective (windows/volumes are independent) -> weak scaling; timing is max over ranks.

Synthetic data, random-init weights (no datasets/checkpoints in the sandbox).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

VOL = (512, 512, 400)
CUBE, STRIDE = 128, 64
FLOP_PER_PATCH = 6.306e11   # forward, one 128^3 patch, in_channel=2, dead dc62 excluded
METRIC = "ct_voxels_per_sec_sliding_window_inference"


def synthetic_ct(shape, seed=777):
    """int16 stored CT values (HU + 1024): soft tissue ~ N(424, 400) clipped to [0, 4095] plus a few air tubes."""
    g = torch.Generator().manual_seed(seed)
    v = torch.empty(shape, dtype=torch.float32)
    v.normal_(424.0, 400.0, generator=g)
    X, Y, Z = shape
    for i in range(6):
        cx, cy, r = int(X * (0.25 + 0.1 * i)), int(Y * (0.3 + 0.07 * i)), 3 + i
        v[cx - r:cx + r, cy - r:cy + r, :] = 30.0
    return v.clamp_(0, 4095).round_().to(torch.int16)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["bf16_tflops_sustained"], d["hbm_gbs"], "measured (MEASURED_PEAKS.json, sustained bf16 GEMM)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm = [float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        if sm:
            out["sm_mhz"] = statistics.median(sm)
            out["sm_max_mhz"] = float(rows[0][2])
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for i, n in enumerate(names):
                if any(r[5 + i].strip().lower().startswith("active") for r in rows if len(r) >= 9):
                    out["reasons"].append(n)
        return out


def cpu_reference_patch_seconds(n_patches, threads):
    """Times the oracle's fp32 CPU restatement of SE_UNet.forward (eval, no_grad, in_channel=2) on 128^3 patches."""
    from oracle import seunet_oracle as oracle
    torch.set_num_threads(threads)
    sd = oracle.init_params(2, 1, seed=777)
    x = torch.rand(1, 2, CUBE, CUBE, CUBE, generator=torch.Generator().manual_seed(1))
    ts = []
    with torch.no_grad():
        for _ in range(n_patches):
            t0 = time.perf_counter()
            oracle.forward(sd, x)
            ts.append(time.perf_counter() - t0)
    return ts


def n_windows():
    from se_unet_airseg_b200.inference import window_starts
    n = 1
    for L in VOL:
        n *= len(window_starts(L, CUBE, STRIDE))
    return n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    nwin = n_windows()
    vox = VOL[0] * VOL[1] * VOL[2]
    ts = cpu_reference_patch_seconds(args.warmup + args.steps, threads)[args.warmup:]
    t_patch = sum(ts) / len(ts)
    value = vox / (t_patch * nwin)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "voxels/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_patch * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "sliding-window SE_UNet inference of one synthetic 512x512x400 CT volume (294 windows 128^3, stride 64)",
                   "note": "each step = ONE 128^3 window forward of the fp32 CPU oracle (torch CPU ops, the reference's own code path); "
                           "volume throughput extrapolated as voxels / (294 * t_window)"},
        "cpu_baseline": {"value": value, "unit": "voxels/s", "cores": threads, "kind": "port",
                         "sample": f"{len(ts)} forward passes of one 1x2x128^3 window, fp32, eval mode"},
        "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_ours(args):
    import torch.distributed as dist
    from se_unet_airseg_b200 import SE_UNet, _lib
    from se_unet_airseg_b200.inference import SlidingWindowPredictor

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU path for the product")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()

    torch.manual_seed(777)
    model = SE_UNet(2, 1).to(dev).eval()
    sw = SlidingWindowPredictor(model, CUBE, STRIDE, batch=args.batch, streams=args.streams)
    img_host = synthetic_ct(VOL, seed=777 + rank).pin_memory()
    img_dev = img_host.to(dev)
    vox = VOL[0] * VOL[1] * VOL[2]
    nwin = n_windows()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ device-resident timing
    for _ in range(args.warmup):
        sw.predict_device(img_dev)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        mask = sw.predict_device(img_dev)
        ev[i + 1].record()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms_total = ev[0].elapsed_time(ev[-1])
    fg_frac = float(mask.float().mean().item())

    # ------------------------------------------------------------------ end-to-end (host in, host out)
    for _ in range(min(2, args.warmup)):
        sw.predict(img_host)
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        sw.predict(img_host)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)

    # ------------------------------------------------------------------ per-kernel timing (CUDA events inside the plan)
    plan = model._plan(args.batch, CUBE, CUBE, CUBE, 0, dev)
    L.seunet_plan_set_timing(plan.handle, 1)
    conv_ms, conv_flops, other_ms, nconv = 0.0, 0.0, 0.0, 0
    per_layer = {}
    for _ in range(2):
        x = torch.rand(args.batch, 2, CUBE, CUBE, CUBE, device=dev)
        with torch.no_grad():
            model(x)
        torch.cuda.synchronize()
        conv_ms, conv_flops, other_ms, nconv = 0.0, 0.0, 0.0, 0
        for i in range(L.seunet_plan_timing_count(plan.handle)):
            lab, ms, fl = ctypes.c_char_p(), ctypes.c_float(), ctypes.c_double()
            _lib.check(L.seunet_plan_timing_get(plan.handle, i, ctypes.byref(lab), ctypes.byref(ms), ctypes.byref(fl)), "timing")
            name = lab.value.decode()
            per_layer[name] = ms.value
            if name.startswith("conv:"):
                conv_ms += ms.value
                conv_flops += fl.value
                nconv += 1
            else:
                other_ms += ms.value
    L.seunet_plan_set_timing(plan.handle, 0)

    # ------------------------------------------------------------------ secondary metric: DP training step (BASELINE config 3)
    train = None
    if not args.no_train and 8 % world == 0:
        from se_unet_airseg_b200.trainer import DataParallelTrainer
        del sw
        model._runtime().plans.clear(); model._runtime().order.clear()
        torch.cuda.empty_cache()
        bt = 8 // world                                   # global batch 8 (train.py:167), sharded over the ranks
        model.train()
        tr = DataParallelTrainer(model, stage=2)
        gt = torch.Generator(device=dev).manual_seed(1234 + rank)
        xt = torch.rand(bt, 2, CUBE, CUBE, CUBE, device=dev, generator=gt)
        lab = (torch.rand(bt, 1, CUBE, CUBE, CUBE, device=dev, generator=gt) > 0.98).float()
        wgt = torch.where(lab > 0, torch.rand(lab.shape, device=dev, generator=gt) * 2 + 0.5, torch.ones_like(lab))
        for _ in range(2):
            tr.step(xt, lab, wgt)
        barrier()
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0e.record()
        tsteps = 3
        for _ in range(tsteps):
            loss_t = tr.step(xt, lab, wgt)
        t1e.record()
        barrier()
        ms_train = t0e.elapsed_time(t1e) / tsteps
        if world > 1:
            tt = torch.tensor([ms_train], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms_train = tt.item()
        train = {"metric": "train_patches_per_sec", "value": 8 / (ms_train * 1e-3), "unit": "patches/s", "ms_per_step": ms_train,
                 "global_batch": 8, "per_rank_batch": bt, "patch": "128^3", "stage": 2, "scaling": "strong",
                 "step": "forward + GUL loss sums (+NCCL all-reduce) + backward + gradient SUM all-reduce + fused AdamW",
                 "loss": float(loss_t.item()), "tflops": 1.89e12 * 8 / (ms_train * 1e-3) / 1e12}
        model.eval()
    if world > 1:
        t = torch.tensor([ms_total, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e = t.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_step = ms_total / args.steps
    value = world * vox / (ms_step * 1e-3)
    e2e_value = world * vox / (ms_e2e / e2e_steps * 1e-3)
    peak_tf, peak_gbs, peak_src = peaks()
    achieved_tf = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "conv_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))   # one ncu --set full capture at batch 1 (profiles/r01_conv_full_per_layer.txt)
        traffic = tj.get("dram_bytes_per_launch") * args.batch / max(1, tj.get("batch", 1))
    whole_tf = nwin * FLOP_PER_PATCH / (ms_step * 1e-3) / 1e12

    # CPU baseline on this box's host cores: bounded sample of the same workload (2 windows)
    threads = os.cpu_count() or 1
    cpu = None
    if not args.no_cpu_baseline:
        ts = cpu_reference_patch_seconds(3, threads)[1:]
        t_patch = sum(ts) / len(ts)
        cpu = {"value": vox / (t_patch * nwin), "unit": "voxels/s", "cores": threads, "kind": "port",
               "sample": f"{len(ts)} timed forwards (+1 warm-up) of one 1x2x128^3 window with the fp32 CPU oracle; "
                         f"{t_patch:.2f} s/window, extrapolated to 294 windows/volume"}

    line = {
        "metric": METRIC, "value": value, "unit": "voxels/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16" if L.seunet_act_dtype() == 0 else "bf16", "data": "synthetic",
        "config": {"workload": "sliding-window SE_UNet inference of one synthetic 512x512x400 CT volume per GPU "
                               "(294 windows of 128^3 at stride 64, eval mode, threshold 0.5)",
                   "in_channel": 2, "window_batch": args.batch, "streams": args.streams, "windows_per_volume": nwin,
                   "l2": "inputs larger than L2: every window forward streams > 1 GB of activations through a 126 MB L2",
                   "accumulate": "fp32", "storage": "fp16 activations/weights, fp32 accumulate/statistics"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "voxels/s", "h2d_bytes_per_step": int(img_host.numel() * 2),
                "d2h_bytes_per_step": int(vox), "ms_per_step": ms_e2e / e2e_steps},
        "gpu_launches": int(args.steps * ((nwin + args.batch - 1) // args.batch) * (24 + 18 + 6 + 3 + 3 + 1) + args.steps * 2),
        "roofline": {"bound": "tensor", "kernel": "conv_tc_kernel (tcgen05 implicit-GEMM conv, all 24 launches of one forward)",
                     "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                     "traffic": traffic, "peak_source": peak_src,
                     "launches_per_forward": nconv, "conv_ms_per_forward": conv_ms, "other_ms_per_forward": other_ms,
                     "whole_step_tflops": whole_tf, "whole_step_frac": whole_tf / peak_tf,
                     "top_layers_ms": dict(sorted(per_layer.items(), key=lambda kv: -kv[1])[:8])},
        "cpu_baseline": cpu,
        "mask_foreground_fraction": fg_frac,
        "train": train,
        "patches_per_s": world * nwin / (ms_step * 1e-3),
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=7, help="windows per forward (294 = 7 * 42; 7 fills the 148 persistent conv CTAs slightly better than 6)")
    ap.add_argument("--streams", type=int, default=3, help="CUDA streams alternating over window batches (overlaps HBM-bound and tensor-bound kernels)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the secondary training-step measurement")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
