#!/usr/bin/env python
"""Benchmark of the cooperative-training hot path (BASELINE.json metric:
"cooperative-training samples/sec at 1/2/4/8 B200; masking-kernel HBM GB/s").

    python bench.py --gpus N --steps K --warmup W            # this build, one rank per GPU
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)
    torchrun --nproc-per-node N bench.py --gpus N ...        # N > 1
    python bench.py --size 256                               # BASELINE.json configs[2]'s shape (64 x 1x256x256 per GPU)
    python bench.py --workload predict                       # BASELINE.json configs[4]: inference on 10x256x256 stacks

Workload `train` (default; config.workload): BASELINE.json configs[1] -- ACDC cooperative_training config, synthetic
1x224x224 slices, 4 classes, batch 64 PER GPU, bf16, channel(image code)+spatial(shape code) targeted soft masking with
random thresholds.  A "step" is one full cooperative step (three passes + hard-example generation + backward + gradient
all-reduce + Adam) over one batch.  Weak scaling: per-GPU batch is fixed.
Workload `predict`: a "step" is one 10-slice 256x256 stack through FTN + STN refinement (predict, n_iter=2) with the
arg-max label map and the confusion matrix produced on the device; replicas only (no collective).

One JSON line on stdout (rank 0).  See DESIGN.md section 6 for how each field is measured.
"""
import argparse
import json
import os
import random
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GF_PER_SAMPLE = {224: 64.01, 256: 83.61, 192: 47.03}      # conv GFLOP per sample-step
IMAGE_CFG = {"loss_name": "mse", "mask_type": "channel", "max_threshold": 0.5, "random_threshold": True,
             "if_soft": True}
SEG_CFG = {"loss_name": "ce", "mask_type": "spatial", "max_threshold": 0.5, "random_threshold": True,
           "if_soft": True}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
                "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def synthetic_batch(n, size, seed, device=None, pin=False):
    """images U[0,1) f32 [n,1,H,W], labels randint{0..3} i64 [n,H,W]."""
    g = torch.Generator().manual_seed(seed)
    img = torch.rand((n, 1, size, size), generator=g, dtype=torch.float32)
    lab = torch.randint(0, 4, (n, size, size), generator=g, dtype=torch.int64)
    if pin:
        img, lab = img.pin_memory(), lab.pin_memory()
    if device is not None:
        img, lab = img.to(device), lab.to(device)
    return img, lab


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1]))
                    mx.append(float(parts[2]))
                except ValueError:
                    continue
                for name, val in zip(names, parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:  # noqa: BLE001
            pass
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_steps(steps, warmup, batch, size, threads, workload="train"):
    """The reference's own CPU path (oracle/model_oracle.py restates it; the reference is Python and cannot
    travel to the GPU box) on a bounded sample of the workload: same step, same mask config, batch `batch`."""
    from oracle import model_oracle
    torch.set_num_threads(threads)
    random.seed(0)
    np.random.seed(0)
    torch.manual_seed(0)
    solver = model_oracle.OracleSolver(num_classes=4, learning_rate=1e-4, seed=0)
    img, lab = synthetic_batch(batch, size, seed=0)
    if workload == "predict":
        solver.eval()
        fn = lambda: solver.predict(img, n_iter=2).max(1)[1].numpy()       # noqa: E731  (test...solver.py:104)
    else:
        fn = lambda: solver.cooperative_step(img, lab, IMAGE_CFG, SEG_CFG)  # noqa: E731
    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    predict = args.workload == "predict"
    batch = 10 if predict else args.ref_batch
    size = args.size
    value, dt = cpu_reference_steps(args.steps, args.warmup, batch, size, threads, args.workload)
    unit = "slices/s" if predict else "samples/s"
    sample = ("oracle/model_oracle.py (torch fp32 CPU restatement of the reference), %s, batch %d of 1x%dx%d per step, "
              "%d timed steps, %d host threads"
              % ("predict(n_iter=2) + arg-max" if predict else "one cooperative step", batch, size, size, args.steps, threads))
    cfg = workload_config(args, batch_per_gpu=batch)
    # this arm is the CPU port: say so instead of repeating the GPU arm's execution mode
    cfg.update({"precision": "fp32 (torch CPU)", "step_mode": "eager torch ops on the host cores", "parallelism": "cpu",
                "global_batch": batch, "sampled_batch": batch,
                "note": "bounded sample of the GPU arm's workload: batch %d per step instead of %d per GPU"
                        % (batch, args.batch)})
    line = {
        "impl": "reference", "metric": metric_name(args), "value": value, "unit": unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def metric_name(args):
    return "inference slices/sec (FTN + STN refinement)" if args.workload == "predict" else "cooperative-training samples/sec"


def workload_config(args, batch_per_gpu):
    if args.workload == "predict":
        return {"workload": "BASELINE.json configs[4]: inference-only FTN+STN refinement (predict, n_iter=2) on synthetic 3D "
                            "stacks, %d slices x 1x%dx%d per stack (one stack per step and GPU), arg-max label map + "
                            "confusion matrix on the device" % (batch_per_gpu, args.size, args.size),
                "batch_per_gpu": batch_per_gpu, "global_batch": batch_per_gpu * args.gpus, "image": [1, args.size, args.size],
                "num_classes": 4, "precision": args.precision, "parallelism": "replicas x%d (no collective)" % args.gpus,
                "step_mode": "CUDA-graph replay (inference.GraphedPredictor)",
                "l2_policy": "L2 flushed between timed steps is NOT applied: every step reads a different stack from a "
                             "ring of stacks larger than L2 (>= 160 MB of inputs + activations per GPU)"}
    which = "configs[1]" if args.size == 224 else ("configs[2] (per-GPU shard: global batch 512 over 8 GPUs)"
                                                   if args.size == 256 else "configs[1] at a non-standard size")
    return {"workload": "BASELINE.json %s: ACDC cooperative_training step, synthetic 1x%dx%d slices, 4 classes, "
                        "batch %d per GPU, channel(image code)+spatial(shape code) targeted soft masking, random "
                        "thresholds (max 0.5)" % (which, args.size, args.size, batch_per_gpu),
            "batch_per_gpu": batch_per_gpu, "global_batch": batch_per_gpu * args.gpus, "image": [1, args.size, args.size],
            "num_classes": 4, "precision": args.precision, "parallelism": "dp%d" % args.gpus,
            "step_mode": "eager launches" if getattr(args, "no_graph", False) else
                         "CUDA-graph replay (ONE graph: fwd+bwd, NCCL all-reduce, multi-tensor Adam)",
            "l2_policy": "inputs larger than L2: one step touches several GB of activations (126 MB L2); the masking "
                         "microbench reads/writes 308 MB per call and additionally flushes L2 between iterations"}


# ------------------------------------------------------------------------------------------------ masking microbench
def masking_microbench(pkg, peaks, iters=20):
    """BASELINE.json configs[3]: [512,64,28,28] fp32 latent codes.  Times the K1+K2 pair (one fused C-ABI call,
    two kernels) with CUDA events on the launching stream, L2 flushed between iterations."""
    N, C, H, W = 512, 64, 28, 28
    gen = torch.Generator(device="cuda").manual_seed(0)
    z = torch.relu(torch.randn(N, C, H, W, device="cuda", generator=gen))
    g = 1e-5 * torch.randn(N, C, H, W, device="cuda", generator=gen)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    numel = N * C * H * W
    out = {}

    def timed(fn, n_iter):
        ts = []
        for _ in range(3):
            fn()
        for _ in range(n_iter):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e-3)
        return statistics.mean(ts), min(ts)

    rng = pkg.ops.NativeRNG(0)
    for name, mode, n in (("channel", pkg.ops.MODE_CHANNEL, C), ("spatial", pkg.ops.MODE_SPATIAL, H * W)):
        for p in (0.1, 0.3, 0.5):
            k = int(n * p)
            mean_t, best_t = timed(lambda: pkg.ops.saliency_mask_apply(g, z, mode, k, soft=True, rng=rng), iters)
            algo = 12.0 * numel + 8.0 * N * n                     # g + z read, z~ written (fp32) + s, mask
            out["%s_p%02d" % (name, int(p * 100))] = {"us": mean_t * 1e6, "best_us": best_t * 1e6,
                                                     "GBps": algo / mean_t / 1e9, "algo_bytes": algo}
    s = pkg.ops.saliency_reduce(g, pkg.ops.MODE_CHANNEL)
    mean_t, _ = timed(lambda: pkg.ops.saliency_reduce(g, pkg.ops.MODE_CHANNEL), iters)
    out["k1_channel"] = {"us": mean_t * 1e6, "GBps": 4.0 * numel / mean_t / 1e9}
    mean_t, _ = timed(lambda: pkg.ops.saliency_reduce(g, pkg.ops.MODE_SPATIAL), iters)
    out["k1_spatial"] = {"us": mean_t * 1e6, "GBps": 4.0 * numel / mean_t / 1e9}
    mean_t, _ = timed(lambda: pkg.ops.topp_mask_apply(s, z, pkg.ops.MODE_CHANNEL, 19, soft=True, rng=rng), iters)
    out["k2_channel"] = {"us": mean_t * 1e6, "GBps": 8.0 * numel / mean_t / 1e9}
    mean_t, _ = timed(lambda: pkg.ops.channel_dropout(z, 0.5, rng=rng, want_mask=False), iters)
    out["dropout_p50"] = {"us": mean_t * 1e6, "GBps": 8.0 * numel / mean_t / 1e9}
    mean_t, _ = timed(lambda: pkg.ops.channel_dropout(z, 0.5, rng=rng, want_mask=True), iters)
    out["dropout_p50_with_quirk_mask"] = {"us": mean_t * 1e6, "GBps": 12.0 * numel / mean_t / 1e9}
    head = out["channel_p30"]
    traffic = None          # dram bytes per call from the committed ncu --set full capture (profiles/, not a live value)
    try:
        with open(os.path.join(ROOT, "profiles", "r1_masking_dram_traffic.json")) as f:
            traffic = float(json.load(f)["traffic_bytes_per_call"])
    except (OSError, KeyError, ValueError):
        pass
    roofline = {"kernel": "ctl_saliency_mask_apply (K1 saliency_channel + K2 topp_mask_apply), [512,64,28,28] fp32, "
                          "channel mode, p=0.3, soft, native Philox",
                "bound": "hbm", "achieved": head["GBps"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": head["GBps"] / peaks["hbm_gbs"], "traffic": traffic,
                "traffic_source": "profiles/r1_masking_dram_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                "algorithmic_bytes_per_launch": head["algo_bytes"], "peak_source": peaks["source"] + " (burst copy)",
                "avg_launch_us": head["us"]}
    return roofline, out


# ------------------------------------------------------------------------------------------------ conv microbench
def conv_microbench(pkg, peaks, batch, iters=5, sizes=((16, 16, 224), (32, 32, 112), (64, 64, 56), (128, 128, 28))):
    """K3 (forward / dgrad kernel) and K3w (weight gradient) on the 3x3 layer classes of FCN_16_standard at the bench
    batch: TFLOP/s against the measured bf16 peak (tensor-pipe utilisation) next to GB/s over the algorithmic bytes
    (bf16 in + out) against the measured HBM peak -- the 16/32-channel layers sit below the ridge, so
    their bound is HBM.  CUDA events, L2 flushed between iterations."""
    ops = pkg.ops
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    out = {}

    def timed(fn):
        for _ in range(2):
            fn()
        ts = []
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e-3)
        return statistics.mean(ts)

    def streamed(make_fn, n_sets=3, reps=21):
        """The same launch back to back: a CUDA-graph replay of `reps` launches cycling over `n_sets` input / output sets
        (together larger than L2), the way the kernel runs inside the step (chained launches, no idle ramp between)."""
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            fns = [make_fn() for _ in range(n_sets)]
            for f in fns:
                f()
            st.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                for i in range(reps):
                    fns[i % n_sets]()
            g.replay()
            ts = []
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(st); g.replay(); b.record(st)
                st.synchronize()
                ts.append(a.elapsed_time(b) * 1e-3 / reps)
        return sorted(ts)[len(ts) // 2]

    for cin, cout, size in sizes:
        x = ops.nchw_to_c8(torch.randn(batch, cin, size, size, device="cuda"))
        dy = ops.nchw_to_c8(torch.randn(batch, cout, size, size, device="cuda") * 0.1)
        wp = ops.pack_conv_weight(torch.randn(cout, cin, 3, 3, device="cuda") * 0.05)
        shift = torch.randn(cout, device="cuda")
        dw = torch.zeros(cout, cin, 3, 3, device="cuda")
        flops = 2.0 * batch * size * size * cout * cin * 9
        byts = 2.0 * batch * size * size * (cin + cout)
        for name, fn in (("fwd", lambda: ops.conv2d_c8(x, wp, cout, 9, shift=shift, act=ops.ACT_LRELU)),
                         ("wgrad", lambda: ops.conv_wgrad_c8(x, dy, 9, out=dw, layout='conv'))):
            t = timed(fn)
            out["%s_%dto%d_3x3_%d" % (name, cin, cout, size)] = {
                "us": round(t * 1e6, 1), "TFLOPs": round(flops / t / 1e12, 1),
                "tensor_frac": round(flops / t / 1e12 / peaks["bf16_tflops"], 3),
                "GBps": round(byts / t / 1e9, 0), "hbm_frac": round(byts / t / 1e9 / peaks["hbm_gbs"], 3)}
        if 3 * byts > 2.5 * 126e6:              # only where three input / output sets exceed L2 by a wide margin
            def make_fwd():
                xs = ops.nchw_to_c8(torch.randn(batch, cin, size, size, device="cuda"))
                return lambda: ops.conv2d_c8(xs, wp, cout, 9, shift=shift, act=ops.ACT_LRELU)
            ts = streamed(make_fwd)
            e = out["fwd_%dto%d_3x3_%d" % (cin, cout, size)]
            e["us_back_to_back"] = round(ts * 1e6, 1)
            e["hbm_frac_back_to_back"] = round(byts / ts / 1e9 / peaks["hbm_gbs"], 3)
            e["back_to_back"] = "median of 5 CUDA-graph replays of 21 launches cycling over 3 input / output sets (> L2)"
    return out


# ------------------------------------------------------------------------------------------------ GPU arm
def shutdown_distributed(dist, world, trainer=None):
    """Tears the process group down at a point ALL ranks reach together (right after the timed loops): graphs that
    recorded NCCL kernels go first -- a communicator still referenced by live CUDA graphs cannot be destroyed -- then
    the group.  A watchdog ends the process if the teardown blocks anyway."""
    if world <= 1:
        return
    import threading
    sys.stdout.flush()
    dog = threading.Timer(60.0, lambda: os._exit(3))
    dog.daemon = True
    dog.start()
    if trainer is not None and hasattr(trainer, "close"):
        trainer.close()
    dist.barrier()
    dist.destroy_process_group()
    dog.cancel()


def eager_gpu_baseline(args, img_d, lab_d):
    """The bar SURVEY.md section 2 / BASELINE.md name: the reference's own op sequence as eager PyTorch fp32 on THIS
    GPU (the reference runs `.to('cuda')` fp32 eager, advanced...model.py:133-138) -- the torch-ops yardstick
    (yardstick/torch_modes.py, 'fp32' = cuDNN true fp32, no TF32; and 'bf16' = cuDNN bf16 NHWC), same batch and
    mask configuration, eager launches, device-resident inputs, 2 warm-up + 4 timed steps each."""
    import cooperative_training_and_latent_space_data_augmentation_b200 as pkg
    import yardstick
    yardstick.install()
    out = {}
    try:
        for mode in ("fp32", "bf16"):
            pkg.conv_blocks.set_precision(mode)
            torch.manual_seed(0)
            solver = pkg.AdvancedTripletReconSegmentationModel('FCN_16_standard', num_classes=4, learning_rate=1e-4)
            trainer = pkg.CooperativeTrainer(solver, args.batch, seed=0, image_cfg=IMAGE_CFG, seg_cfg=SEG_CFG)
            for _ in range(2):
                trainer.step(img_d, lab_d)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(4):
                trainer.step(img_d, lab_d)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 4
            out[mode] = {"samples_per_s": args.batch * 1e3 / ms, "ms_per_step": ms}
            del trainer, solver
            torch.cuda.empty_cache()
    finally:
        pkg.conv_blocks.set_precision("kernel")
    return {"what": "eager PyTorch on the same B200, the reference's op sequence (torch / cuDNN convolutions, autograd, "
                    "torch-fp32 parity mode and cuDNN-bf16 mode of yardstick/torch_modes.py), batch %d of 1x%dx%d, "
                    "2 warm-up + 4 timed steps, device-resident inputs" % (args.batch, args.size, args.size),
            "fp32": out.get("fp32"), "bf16_cudnn": out.get("bf16")}


def run_predict_arm(args, world, rank, local_rank, dist):
    import cooperative_training_and_latent_space_data_augmentation_b200 as pkg
    from cooperative_training_and_latent_space_data_augmentation_b200 import _lib
    torch.manual_seed(0)
    solver = pkg.AdvancedTripletReconSegmentationModel('FCN_16_standard', num_classes=4)
    S, size = args.slices, args.size
    predictor = pkg.GraphedPredictor(solver, (S, 1, size, size), n_iter=2)
    # a ring of stacks larger than L2, so that no step finds its input (or the previous step's activations) cached
    ring = max(4, min(64, (160 << 20) // (S * size * size * 4) + 1))
    stacks = []
    for i in range(ring):
        img, lab = synthetic_batch(S, size, seed=5000 + 97 * rank + i, pin=False)
        stacks.append((img.pin_memory(), lab.to(torch.uint8).pin_memory()))
    dev_stacks = [(a.cuda(), b.cuda().long()) for a, b in stacks]
    out_host = [torch.empty((S, size, size), dtype=torch.uint8).pin_memory() for _ in range(2)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    state = {"i": 0}

    def device_step():
        a, b = dev_stacks[state["i"] % ring]
        state["i"] += 1
        return predictor.predict_chunk(a, b)

    def e2e_step():
        a, b = stacks[state["i"] % ring]                 # pinned HOST image (fp32) and ground truth (uint8)
        o = out_host[state["i"] % 2]
        state["i"] += 1
        labels = predictor.predict_chunk(a, b)
        o.copy_(labels, non_blocking=True)               # D2H of the label map (test...solver.py:104)
        torch.cuda.current_stream().synchronize()        # the caller reads it (numpy) before the next chunk
        return int(o[0, 0, 0])

    def timed_loop(fn, steps):
        barrier()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(steps):
            last = fn()
        stop.record()
        barrier()
        t = torch.tensor([start.elapsed_time(stop)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) * 1e-3, last

    n_warm = max(args.warmup, 3)
    for _ in range(n_warm):
        device_step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    steps = args.steps * 10                               # a step is ~1 ms: keep the timed region long enough to clock
    launches0 = _lib.LAUNCHES["count"]
    kernels_per_step = None
    t_dev, _ = timed_loop(device_step, steps)
    for _ in range(2):
        e2e_step()
    t_e2e, _ = timed_loop(e2e_step, steps)
    clocks = sampler.stop() if rank == 0 else None
    # kernels per replayed chunk: counted once eagerly on the same path (graph replays do not pass through ctypes)
    n0 = _lib.LAUNCHES["count"]
    with torch.no_grad():
        predictor._forward(dev_stacks[0][0], dev_stacks[0][1])
    kernels_per_step = _lib.LAUNCHES["count"] - n0
    scores, cls_iu = predictor.scores()
    shutdown_distributed(dist, world)
    if rank != 0:
        return
    peaks = measured_peaks()
    conv_layers = conv_microbench(pkg, peaks, S, sizes=((16, 16, size),))
    top = conv_layers["fwd_16to16_3x3_%d" % size]
    gf = {224: 5.63, 256: 7.36}.get(size)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, dt = cpu_reference_steps(3, 1, S, size, threads, "predict")
        cpu = {"value": v, "unit": "slices/s", "cores": threads, "kind": "port",
               "sample": "oracle/model_oracle.py predict(n_iter=2) + arg-max on one %d-slice stack, 1 warm-up + 3 timed "
                         "(%.1f s)" % (S, dt)}
    value = S * world * steps / t_dev
    line = {
        "metric": metric_name(args), "value": value, "unit": "slices/s", "n_gpus": world, "steps": steps,
        "warmup": args.warmup, "warmup_run": n_warm, "ms_per_step": 1e3 * t_dev / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args, S),
        "e2e": {"value": S * world * steps / t_e2e, "unit": "slices/s",
                "h2d_bytes_per_step": int(S * size * size * 5) * world, "d2h_bytes_per_step": int(S * size * size) * world,
                "ms_per_step": 1e3 * t_e2e / steps,
                "api": "GraphedPredictor.predict_chunk on pinned host tensors (fp32 image + uint8 ground truth H2D, graph "
                       "replay, uint8 label map D2H, stream synchronised every chunk)"},
        "gpu_launches": kernels_per_step * steps,
        "roofline": {"kernel": "conv_small_kernel (K3s) 16->16 3x3 @%d^2, batch %d (dominant kernel class of the inference pass)"
                               % (size, S), "bound": "hbm", "achieved": top["GBps"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": top["hbm_frac"], "traffic": None, "avg_launch_us": top["us"],
                     "peak_source": peaks["source"] + " (burst copy)"},
        "conv_blocks": conv_layers,
        "step_tensor_frac": (value / world * gf * 1e9 / (peaks["bf16_tflops_sustained"] * 1e12)) if gf else None,
        "metric_scores": {k.strip(): float(v) for k, v in scores.items()},
        "cpu_baseline": cpu, "clocks": clocks,
    }
    print(json.dumps(line), flush=True)


def run_gpu_arm(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this build has no CPU fallback (use --impl reference for the "
                         "CPU baseline)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if world != args.gpus and rank == 0:
        print("bench.py: warning: --gpus %d but WORLD_SIZE %d" % (args.gpus, world), file=sys.stderr)
    args.gpus = world
    if args.workload == "predict":
        return run_predict_arm(args, world, rank, local_rank, dist)

    import cooperative_training_and_latent_space_data_augmentation_b200 as pkg
    from cooperative_training_and_latent_space_data_augmentation_b200 import _lib
    if args.precision != "kernel":
        import yardstick
        yardstick.install()
    pkg.conv_blocks.set_precision(args.precision)
    torch.manual_seed(0)
    solver = pkg.AdvancedTripletReconSegmentationModel('FCN_16_standard', num_classes=4, learning_rate=1e-4)
    global_batch = args.batch * world
    if args.no_graph:
        trainer = pkg.CooperativeTrainer(solver, global_batch, seed=0, image_cfg=IMAGE_CFG, seg_cfg=SEG_CFG)
    else:       # product path: the step replayed from ONE CUDA graph (training.GraphedCooperativeTrainer)
        trainer = pkg.GraphedCooperativeTrainer(solver, global_batch, seed=0, image_cfg=IMAGE_CFG, seg_cfg=SEG_CFG,
                                                eager_steps=1 if args.profile else 3)

    # every rank draws its own slice of the global synthetic batch
    img_h, lab_h = synthetic_batch(args.batch, args.size, seed=1000 + rank, pin=True)
    img_d, lab_d = img_h.cuda(non_blocking=True), lab_h.cuda(non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        return trainer.step(img_d, lab_d)

    def e2e_step():
        if args.no_graph:
            out = trainer.step(img_h.cuda(non_blocking=True), lab_h.cuda(non_blocking=True))
        else:
            out = trainer.step(img_h, lab_h)        # pinned host tensors -> the trainer's static device buffers (H2D)
            trainer.prefetch(img_h, lab_h)          # the NEXT step's H2D copy starts now, under this step's compute
        return float(out['loss'].item())            # D2H read of the step's result

    def timed_loop(fn, steps):
        barrier()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(steps):
            last = fn()
        stop.record()
        barrier()
        ms = start.elapsed_time(stop)
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) * 1e-3, last

    n_warm = args.warmup if args.profile else max(args.warmup, 3)
    n_setup = 0
    if not args.no_graph:
        n_setup = trainer.eager_steps + 1           # eager library warm-up + the capture step, then n_warm replays
    for _ in range(n_warm + n_setup):
        device_step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.LAUNCHES["count"]
    if args.profile:                        # `ncu --profile-from-start off`: only the timed (graph-replayed) steps
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    t_dev, _ = timed_loop(device_step, args.steps)
    if args.profile:
        torch.cuda.profiler.stop()
    launches = _lib.LAUNCHES["count"] - launches0
    if args.profile:                        # ncu launch-list pass: the device loop only
        if rank == 0:
            sampler.stop()
            print(json.dumps({"profile_only": True, "ms_per_step": 1e3 * t_dev / args.steps,
                              "kernels_per_step": launches // max(1, args.steps)}), flush=True)
        shutdown_distributed(dist, world, trainer)
        return
    for _ in range(2):
        e2e_step()
    t_e2e, last_loss = timed_loop(e2e_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    # data-parallel result checks (every rank takes part): bit-identical parameters on all ranks after the timed steps,
    # and one extra eager step without the optimizer whose all-reduced gradient must equal the mean of the ranks' local
    # gradients (gathered)
    in_sync = trainer.params_in_sync()
    grad_check = None
    if world > 1:
        from cooperative_training_and_latent_space_data_augmentation_b200.training import cooperative_step
        cooperative_step(solver, img_d, lab_d, IMAGE_CFG, SEG_CFG, optimize=False)
        local = solver.flat_adam.flat_grads.clone()
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        want = torch.stack(gathered).double().mean(0)
        trainer.bucket.all_reduce_sum()
        got = trainer.bucket.mean_gradients().double()
        grad_check = {"max_err_over_max_abs": float((got - want).abs().max() / want.abs().max()),
                      "what": "all-reduced flat gradient x 1/world vs the mean of the %d ranks' gathered local gradients "
                              "(one eager step, optimizer off)" % world}

    shutdown_distributed(dist, world, trainer)          # every rank; rank 0 carries on alone with the microbenchmarks
    if rank != 0:
        return

    peaks = measured_peaks()
    value = global_batch * args.steps / t_dev
    e2e_value = global_batch * args.steps / t_e2e
    roofline, sweep = masking_microbench(pkg, peaks)
    conv_layers = conv_microbench(pkg, peaks, args.batch)
    gf = GF_PER_SAMPLE.get(args.size)
    cpu = None
    eager = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, dt = cpu_reference_steps(3, 1, args.ref_batch, args.size, threads)
        cpu = {"value": v, "unit": "samples/s", "cores": threads, "kind": "port",
               "sample": "oracle/model_oracle.py cooperative step, batch %d of 1x%dx%d, 1 warm-up + 3 timed steps "
                         "(%.1f s)" % (args.ref_batch, args.size, args.size, dt)}
    if world == 1 and not args.no_eager_baseline and args.precision == "kernel":
        del trainer
        torch.cuda.empty_cache()
        eager = eager_gpu_baseline(args, img_d, lab_d)
        if eager.get("fp32"):
            eager["speedup_vs_eager_fp32"] = value / eager["fp32"]["samples_per_s"]
        if eager.get("bf16_cudnn"):
            eager["speedup_vs_eager_cudnn_bf16"] = value / eager["bf16_cudnn"]["samples_per_s"]
    top = conv_layers.get("fwd_16to16_3x3_%d" % args.size) or {}
    line = {
        "metric": "cooperative-training samples/sec", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "warmup_run": n_warm,
        "setup_steps": n_setup, "ms_per_step": 1e3 * t_dev / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
        "config": workload_config(args, args.batch),
        "e2e": {"value": e2e_value, "unit": "samples/s",
                "h2d_bytes_per_step": int(img_h.numel() * 4 + lab_h.numel() * 8) * world,
                "d2h_bytes_per_step": 4 * world, "ms_per_step": 1e3 * t_e2e / args.steps,
                "api": "%s.step on pinned host tensors (next batch prefetched on a copy stream: every step's H2D copy "
                       "is inside the timed region, overlapped with the previous step) + loss.item()"
                       % ("GraphedCooperativeTrainer" if not args.no_graph else "CooperativeTrainer")},
        "gpu_launches": launches,
        "roofline": roofline,
        # the dominant kernel of the STEP (not of the metric's masking half): HBM-bound 16-channel 3x3 conv class
        "roofline_conv": {"kernel": "conv_small_kernel (K3s, warp-level tensor path) 16->16 3x3 @%d^2, batch %d (largest share of the step's device time)"
                                    % (args.size, args.batch), "bound": "hbm", "achieved": top.get("GBps"),
                          "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": top.get("hbm_frac"),
                          "tensor_frac_of_burst_peak": top.get("tensor_frac"), "avg_launch_us": top.get("us"),
                          # isolated launches after an L2 flush (above) pay the ~6 us ramp and tail of a lone kernel; in the
                          # step the kernel runs back to back with its neighbours (chained launches):
                          "back_to_back_us": top.get("us_back_to_back"), "back_to_back_frac": top.get("hbm_frac_back_to_back"),
                          # one ncu --set full capture of this kernel at this shape (dram read 102.8 MB + write 55.8 MB: part of
                          # the 103 MB output is still in L2 when the kernel ends); algorithmic bytes are 2 x 102.8 MB
                          "traffic": 158626048.0 if (args.size == 224 and args.batch == 64) else None,
                          "traffic_source": "profiles/r2_k3s_ncu_summary.csv (plain variant)",
                          "algorithmic_bytes_per_launch": 2.0 * args.batch * 16 * args.size * args.size * 2},
        "peak_convention": "kernels timed alone (roofline, roofline_conv, conv_blocks, masking_*) are divided by the BURST "
                           "peaks of MEASURED_PEAKS.json (hbm_gbs %.1f GB/s, bf16_tflops %.1f); the whole step "
                           "(step_tensor_frac) by the SUSTAINED bf16 peak (%.1f)"
                           % (peaks["hbm_gbs"], peaks["bf16_tflops"], peaks["bf16_tflops_sustained"]),
        "masking_GBps": {k: round(v["GBps"], 1) for k, v in sweep.items()},
        "masking_us": {k: round(v["us"], 2) for k, v in sweep.items()},
        "conv_blocks": conv_layers,
        "step_tensor_frac": (value * gf * 1e9 / (world * peaks["bf16_tflops_sustained"] * 1e12)) if gf else None,
        "params_in_sync": in_sync,
        "dp_gradient_check": grad_check,
        "cpu_baseline": cpu,
        "eager_gpu_baseline": eager,
        "clocks": clocks,
        "last_loss": last_loss,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="batch per GPU (configs[1]: 64)")
    ap.add_argument("--size", type=int, default=224, help="slice height = width (configs[1]: 224)")
    ap.add_argument("--precision", default="kernel", choices=["kernel", "bf16", "fp32"],
                    help="kernel: the sm_100a kernels of this build (product path); bf16/fp32: library convolutions")
    ap.add_argument("--ref-batch", type=int, default=8, help="bounded CPU sample: batch per CPU step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue the step's launches one by one (no CUDA graphs)")
    ap.add_argument("--profile", action="store_true", help="device loop only, honour --warmup < 3 (for ncu)")
    ap.add_argument("--workload", default="train", choices=["train", "predict"],
                    help="train: cooperative step (configs[1], --size 256 for configs[2]); predict: configs[4]")
    ap.add_argument("--slices", type=int, default=10, help="predict workload: slices per stack (configs[4]: 10)")
    ap.add_argument("--no-eager-baseline", action="store_true", help="skip the eager-PyTorch-on-this-GPU baseline leg")
    args = ap.parse_args()
    if args.workload == "predict" and args.size == 224 and "--size" not in " ".join(sys.argv):
        args.size = 256                          # configs[4] is 256 x 256
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
