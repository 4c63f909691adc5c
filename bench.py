#!/usr/bin/env python
"""Benchmark of the SE_UNet hot path on B200 (BASELINE.json metric: CT voxels/s, sliding-window inference).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU

One "step" = sliding-window inference of ONE synthetic 512x512x400 CT volume (BASELINE config 2:
294 windows of 128^3 at stride 64, prediction.py:65-111) per GPU: HU windowing, 294 SE_UNet forwards (batched),
sigmoid, overlap mean, 0.5 threshold.  `value` = volume voxels / s with the stored CT volume already resident in
HBM; `e2e` = the same through the public API (SlidingWindowPredictor.predict) with a pinned HOST volume in and
the HOST mask out.  N > 1: one process per GPU (torchrun); the 294 windows of the SAME volume are sharded by patch over
the ranks, the partial fixed-point planes go point-to-point (NCCL over NVLink) to the rank that owns them, every rank finalizes
1/N of the planes and rank 0 collects the mask -> strong scaling of the latency of one volume (the volume-per-rank sweep without any collective is reported as the secondary key `sweep`); max over ranks.

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


def peaks_all():
    """Roofline denominators: the driver-measured cuBLAS bf16 GEMM peaks of this pool's B200s - burst (a kernel timed alone;
    the figure SURVEY 8d specifies) and sustained (inside a long power-capped step) - and the HBM copy bandwidth."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"burst": d["bf16_tflops"], "sustained": d["bf16_tflops_sustained"], "hbm_gbs": d["hbm_gbs"],
                "source": "measured (MEASURED_PEAKS.json: cuBLAS bf16 8192^3 burst / sustained)"}
    return {"burst": 1590.0, "sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


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
        "warmup": args.warmup, "ms_per_step": t_patch * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "sliding-window SE_UNet inference of one synthetic 512x512x400 CT volume (294 windows of 128^3 at stride 64, "
                               "eval mode, threshold 0.5)",
                   "note": "each step = ONE 128^3 window forward of the fp32 CPU oracle (torch CPU ops, the reference's own code path); "
                           "volume throughput extrapolated as voxels / (294 * t_window)"},
        "cpu_baseline": {"value": value, "unit": "voxels/s", "cores": threads, "kind": "port",
                         "sample": f"{len(ts)} forward passes of one 1x2x128^3 window, fp32, eval mode"},
        "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def cpu_train_step_seconds(threads):
    """BASELINE config 1: SE_UNet forward+backward on one 1x1x128^3 fp32 patch with the reference's PyTorch CPU path
    (oracle port, eval mode, stage-1 Dice loss, train.py:597-602).  One step, no warm-up: ~20-30 s of CPU work."""
    from oracle import seunet_oracle as oracle
    torch.set_num_threads(threads)
    sd = {k: v.requires_grad_(True) for k, v in oracle.init_params(1, 1, seed=777).items()}
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 1, CUBE, CUBE, CUBE, generator=g)
    label = (torch.rand(1, 1, CUBE, CUBE, CUBE, generator=g) > 0.97).float()
    t0 = time.perf_counter()
    p0, p1 = oracle.forward(sd, x)
    t1 = time.perf_counter()
    oracle.stage_loss(1, p0, p1, label).backward()
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1


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
    # N = 1: one volume on one GPU.  N > 1: the SAME single volume, its 294 windows sharded by patch over the ranks (every
    # rank holds the host volume, e.g. the memory-mapped file; it copies only the planes its windows read) and the partial
    # fixed-point probability volumes summed by one NCCL reduce -> strong scaling of the latency of one volume.
    img_host = synthetic_ct(VOL, seed=777).pin_memory()
    img_dev = img_host.to(dev)
    vox = VOL[0] * VOL[1] * VOL[2]
    nwin = n_windows()
    sharded = world > 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        return sw.predict_device_sharded(img_dev, reuse_output=True) if sharded else sw.predict_device(img_dev, reuse_output=True)

    def host_step():
        return sw.predict_sharded(img_host, reuse_output=True) if sharded else sw.predict(img_host, reuse_output=True)

    def max_over_ranks(*vals):
        if world == 1:
            return list(vals)
        t = torch.tensor(list(vals), device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    # ------------------------------------------------------------------ device-resident timing
    for _ in range(args.warmup):
        device_step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        mask = device_step()
    e1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms_total = e0.elapsed_time(e1)
    fg_frac = float(mask.float().mean().item()) if mask is not None else 0.0

    # ------------------------------------------------------------------ end-to-end (host volume in, host mask out)
    for _ in range(min(2, args.warmup)):
        host_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        host_step()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    h2d_bytes = int(img_host.numel() * 2)
    if sharded:
        xa, xb = sw.shard_planes(VOL, rank, world)
        h2d_bytes = int((xb - xa) * VOL[1] * VOL[2] * 2)

    # ------------------------------------------------------------------ end-to-end incl. GPU post-processing (N = 1)
    ms_pp = None
    if not sharded:
        for _ in range(2):
            sw.predict_postprocessed_device(img_dev)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            out_pp = sw.predict_postprocessed_device(img_dev)
        e1.record()
        torch.cuda.synchronize()
        ms_pp = e0.elapsed_time(e1) / 2

    # ------------------------------------------------------------------ secondary (N > 1): volume sweep, one volume per rank, no collective
    sweep = None
    if sharded:
        for _ in range(2):
            sw.predict_device(img_dev, reuse_output=True)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            sw.predict_device(img_dev, reuse_output=True)
        e1.record()
        barrier()
        (ms_sweep,) = max_over_ranks(e0.elapsed_time(e1) / 3)
        sweep = {"metric": "ct_voxels_per_sec_volume_sweep", "value": world * vox / (ms_sweep * 1e-3), "unit": "voxels/s",
                 "ms_per_volume_per_gpu": ms_sweep, "scaling": "weak",
                 "note": "BASELINE config 4 sharded by VOLUME: every rank segments its own volume, no data-path collective"}

    # ------------------------------------------------------------------ per-kernel timing (CUDA events inside the plan)
    plan = model._plan(args.batch, CUBE, CUBE, CUBE, 0, dev)
    L.seunet_plan_set_timing(plan.handle, 1)
    conv_ms, conv_flops, other_ms, nconv = 0.0, 0.0, 0.0, 0
    per_layer = {}
    # the benched step: seunet_forward_window (head 0 skipped, head 1 -> sigmoid -> accumulator), here on a scratch accumulator
    flat, wgen = model._weights(model._param_tensors())
    plan.pack(flat, wgen)
    ones0, ones1 = torch.ones(args.batch, 24, device=dev), torch.ones(args.batch, 12, device=dev)
    acc_scratch = torch.zeros((CUBE, CUBE, CUBE), dtype=torch.int32, device=dev)
    starts0 = (ctypes.c_int * (3 * args.batch))(*([0] * (3 * args.batch)))
    # 2 untimed + 6 timed forwards back to back, per-launch times averaged over the 6 (one sample right after the host-side
    # set-up measured whatever clock the idle GPU was ramping through: 849 vs 1001 TFLOP/s on two boxes of the same pool)
    x = torch.rand(args.batch, 2, CUBE, CUBE, CUBE, device=dev)
    strides = (ctypes.c_int64 * 5)(*x.stride())
    n_timed, layer_sum, flops_of = 6, {}, {}
    for it in range(2 + n_timed):
        _lib.check(L.seunet_forward_window(plan.handle, _lib.ptr(x), strides, None, _lib.ptr(flat), _lib.ptr(ones0), _lib.ptr(ones1),
                                           starts0, _lib.ptr(acc_scratch), CUBE, CUBE, CUBE, 20, _lib.stream_ptr()), "seunet_forward_window")
        torch.cuda.synchronize()
        if it < 2:
            continue
        for i in range(L.seunet_plan_timing_count(plan.handle)):
            lab, ms, fl = ctypes.c_char_p(), ctypes.c_float(), ctypes.c_double()
            _lib.check(L.seunet_plan_timing_get(plan.handle, i, ctypes.byref(lab), ctypes.byref(ms), ctypes.byref(fl)), "timing")
            name = lab.value.decode()
            layer_sum[name] = layer_sum.get(name, 0.0) + ms.value
            flops_of[name] = fl.value
    for name, tot in layer_sum.items():
        per_layer[name] = tot / n_timed
        if name.startswith("conv:"):
            conv_ms += per_layer[name]
            conv_flops += flops_of[name]
            nconv += 1
        else:
            other_ms += per_layer[name]
    L.seunet_plan_set_timing(plan.handle, 0)

    # ------------------------------------------------------------------ secondary metric: DP training step (BASELINE configs 3 and 5)
    train, train5 = None, None
    if not args.no_train:
        from se_unet_airseg_b200.trainer import DataParallelTrainer
        del sw
        model._runtime().plans.clear(); model._runtime().order.clear()
        torch.cuda.empty_cache()

        def time_train(stage, global_batch, S, tsteps, graph=True):
            bt = global_batch // world
            model.train()
            # graph=True: forward + losses + backward + both NCCL exchanges replayed as ONE CUDA graph per step (new inputs are
            # copied into static buffers, AdamW runs eagerly behind it) - the product's own option, same kernels, same work
            tr = DataParallelTrainer(model, stage=stage, graph=graph)
            gt = torch.Generator(device=dev).manual_seed(1234 + rank)
            xt = torch.rand(bt, 2, S, S, S, device=dev, generator=gt)
            lab = (torch.rand(bt, 1, S, S, S, device=dev, generator=gt) > 0.98).float()
            wgt = torch.where(lab > 0, torch.rand(lab.shape, device=dev, generator=gt) * 2 + 0.5, torch.ones_like(lab))
            skel = lab * (torch.rand(lab.shape, device=dev, generator=gt) > 0.5).float() if stage == 3 else None
            for _ in range(3):                                       # eager warm-up, capture, first replay
                tr.step(xt, lab, wgt, skel)
            barrier()
            t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0e.record()
            for _ in range(tsteps):
                loss_t = tr.step(xt, lab, wgt, skel)
            t1e.record()
            barrier()
            (ms_t,) = max_over_ranks(t0e.elapsed_time(t1e) / tsteps)
            model.eval()
            model._runtime().plans.clear(); model._runtime().order.clear()
            del tr, xt, lab, wgt, skel
            torch.cuda.empty_cache()
            flop_patch = 1.89e12 * (S / 128.0) ** 3
            return {"metric": "train_patches_per_sec", "value": global_batch / (ms_t * 1e-3), "unit": "patches/s", "ms_per_step": ms_t,
                    "global_batch": global_batch, "per_rank_batch": bt, "patch": f"{S}^3", "stage": stage, "scaling": "strong",
                    "step": "forward + loss sums (+NCCL all-reduce of 16 fp64 sums) + backward + gradient SUM all-reduce (6.08 MB)" +
                            (", replayed as one CUDA graph (DataParallelTrainer(graph=True))" if graph else "") + " + fused AdamW",
                    "loss": float(loss_t.item()), "tflops": flop_patch * global_batch / (ms_t * 1e-3) / 1e12}

        def time_train_guarded(*a):
            # the secondary metric must never cost the headline line: a failed graph capture falls back to the eager step, and
            # the capture with NCCL inside was only verified on 1, 2 and 4 GPUs in this round (GPU budget) - 8 ranks run eagerly
            if world > 4:
                return time_train(*a, graph=False)
            try:
                return time_train(*a, graph=True)
            except Exception as e:          # noqa: BLE001
                note = f"{type(e).__name__}: {e}"[:200]
            r = time_train(*a, graph=False)
            r["graph_fallback"] = note
            return r

        if 8 % world == 0:
            train = time_train_guarded(2, 8, CUBE, 3)               # config 3: batch 8 x 128^3, GUL loss
        if args.config5 and 16 % world == 0 and 16 // world <= 8:      # 7.6 GB of workspace per 160^3 patch: <= 8 patches per GPU
            train5 = time_train_guarded(3, 16, 160, 2)              # config 5: batch 16 x 160^3, stage-3 loss (needs ~8.2 GB per patch)
    (ms_total, ms_e2e) = max_over_ranks(ms_total, ms_e2e)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_step = ms_total / args.steps
    value = vox / (ms_step * 1e-3)
    e2e_value = vox / (ms_e2e / e2e_steps * 1e-3)
    pk = peaks_all()
    achieved_tf = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    traffic, traffic_note = None, None
    tp = os.path.join(ROOT, "profiles", "conv_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))   # ncu --set full capture of the 24 conv launches of one forward AT THE BENCHED WINDOW BATCH
        if int(tj.get("batch", 0)) == args.batch:
            traffic = tj.get("dram_bytes_per_launch")
            traffic_note = tj.get("source")
    whole_tf = nwin * FLOP_PER_PATCH / (ms_step * 1e-3) / 1e12

    # CPU baselines on this box's host cores: bounded samples of the same workloads
    threads = os.cpu_count() or 1
    cpu = None
    if not args.no_cpu_baseline:
        ts = cpu_reference_patch_seconds(3, threads)[1:]
        t_patch = sum(ts) / len(ts)
        cpu = {"value": vox / (t_patch * nwin), "unit": "voxels/s", "cores": threads, "kind": "port",
               "sample": f"{len(ts)} timed forwards (+1 warm-up) of one 1x2x128^3 window with the fp32 CPU oracle; "
                         f"{t_patch:.2f} s/window, extrapolated to 294 windows/volume"}
        if train is not None:
            tf, tb = cpu_train_step_seconds(threads)
            train["cpu_baseline"] = {"value": 1.0 / (tf + tb), "unit": "patches/s", "cores": threads, "kind": "port",
                                     "sample": f"BASELINE config 1: one forward ({tf:.1f} s) + backward ({tb:.1f} s) of a 1x1x128^3 fp32 patch, "
                                               "Dice loss, eval mode, no warm-up"}

    workload = ("sliding-window SE_UNet inference of one synthetic 512x512x400 CT volume (294 windows of 128^3 at stride 64, eval "
                "mode, threshold 0.5)")
    if sharded:
        workload += f", windows sharded by patch over {world} GPUs, partial planes exchanged point-to-point with their owner ranks (NCCL over NVLink), mask collected on rank 0"
    line = {
        "metric": METRIC, "value": value, "unit": "voxels/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f16" if L.seunet_act_dtype() == 0 else "bf16", "data": "synthetic",
        "config": {"workload": workload,
                   "in_channel": 2, "window_batch": args.batch, "streams": args.streams, "windows_per_volume": nwin,
                   "l2": "inputs larger than L2: every window forward streams > 1 GB of activations through a 126 MB L2",
                   "accumulate": "32-bit fixed point (order-independent), fp32 mean",
                   "storage": "fp16 activations/weights, fp32 accumulate/statistics",
                   "output": "predictor-owned pinned result buffer reused across calls (reuse_output=True)"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "voxels/s", "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": int(vox), "ms_per_step": ms_e2e / e2e_steps,
                "note": "SlidingWindowPredictor.predict%s: pinned host int16 volume in, host uint8 mask out" % ("_sharded" if sharded else "")},
        # per window batch: nconv tcgen05 conv launches (22: two of the 24 convs run inside a fused apply pass) + 18 gate/norm passes
        # + 6 CAT + 3 up-sampling + head (which also accumulates into the volume) + input prep + head weights; per volume: HU windows,
        # accumulator clear, finalize
        "gpu_launches": int(args.steps * (-(-(-(-nwin // world)) // args.batch)) * (nconv + 18 + 6 + 3 + 3) + args.steps * 3),
        "roofline": {"bound": "tensor", "kernel": "conv_tc_kernel (tcgen05 implicit-GEMM conv, all %d launches of one forward)" % nconv,
                     "achieved": achieved_tf, "peak": pk["burst"], "unit": "TFLOP/s", "frac": achieved_tf / pk["burst"],
                     "frac_of_sustained": achieved_tf / pk["sustained"], "peak_sustained": pk["sustained"],
                     "traffic": traffic, "traffic_source": traffic_note, "peak_source": pk["source"],
                     "launches_per_forward": nconv, "conv_ms_per_forward": conv_ms, "other_ms_per_forward": other_ms,
                     "whole_step_tflops": whole_tf, "whole_step_frac": whole_tf / (world * pk["burst"]),
                     "whole_step_frac_of_sustained": whole_tf / (world * pk["sustained"]),
                     "top_layers_ms": dict(sorted(per_layer.items(), key=lambda kv: -kv[1])[:8])},
        "cpu_baseline": cpu,
        "mask_foreground_fraction": fg_frac,
        "patches_per_s": nwin / (ms_step * 1e-3),
    }
    if ms_pp is not None:
        line["e2e_postprocessed"] = {"value": vox / (ms_pp * 1e-3), "unit": "voxels/s", "ms_per_step": ms_pp,
                                     "note": "device-resident volume -> sliding window -> double-threshold sweep, border crop, "
                                             "largest component, hole filling on the GPU (prediction.py:78-116)",
                                     "foreground_voxels": int(out_pp.sum().item())}
    if sweep is not None:
        line["sweep"] = sweep
    line["train"] = train
    if train5 is not None:
        line["train_s160"] = train5
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
    ap.add_argument("--config5", action="store_true", help="also time BASELINE config 5 (stage 3, batch 16 x 160^3; needs >= 2 GPUs)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
