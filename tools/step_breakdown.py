"""Per-entry-point GPU time of one eager cooperative step (BASELINE.json configs[1] shapes), measured with CUDA events
around every C-ABI call (no profiler): which calls of libctl_b200.so the step spends its device time in, keyed by
entry point and -- for the conv kernels -- by layer class.  Everything between two calls (torch's own kernels: losses,
optimizers, glue) is reported as 'torch / gaps'.  usage: python tools/step_breakdown.py [batch] [steps] [sync]
('sync' drains the queue before every call: slower, but an interval then never contains time the GPU waited for Python)"""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cooperative_training_and_latent_space_data_augmentation_b200 as pkg  # noqa: E402
from cooperative_training_and_latent_space_data_augmentation_b200 import _lib  # noqa: E402
import bench  # noqa: E402

RECORDS = []
SYNC = len(sys.argv) > 3 and sys.argv[3] == 'sync'


def key_of(name, a):
    if name == "ctl_conv2d_c8_bf16":       # x, N, H, W, Cin, w, Cout, taps, subsample, up2x, scale, shift, res, rs, rb, act, out, stats
        return "conv %d->%d k%d @%d%s%s%s%s" % (a[4], a[6], 3 if a[7] == 9 else 1, a[2], " s2" if a[8] == 2 else "",
                                              " up2x" if a[9] else "", " +res" if a[12] else "", " +stats" if a[17] else "")
    if name == "ctl_conv_wgrad_c8_bf16":   # x, dy, N, H, W, Cin, Cout, taps
        return "wgrad %d->%d k%d @%d" % (a[5], a[6], 3 if a[7] == 9 else 1, a[3])
    if name in ("ctl_bn_bwd_reduce_c8", "ctl_bn_bwd_apply_c8"):   # dy, h, a, N, C, H, W
        return "%s C%d @%d%s" % (name[4:], a[4], a[5], "" if a[1] else " (no h)")
    if name in ("ctl_scale_shift_act_c8", "ctl_bn_batch_affine_c8", "ctl_channel_sums_c8", "ctl_upsample2x_c8",
                "ctl_downsample2x_sum_c8", "ctl_zero_stuff2x_c8", "ctl_split_parity2x2_c8"):   # x, N, C, H, W
        return "%s C%d @%d" % (name[4:], a[2], a[3])
    return name[4:]


def instrument():
    lib = _lib.load()
    for name in _lib.KERNELS_PER_CALL:
        fn = getattr(lib, name)

        def timed(*args, _fn=fn, _name=name):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if SYNC:
                torch.cuda.synchronize()        # empty queue: the interval is launch latency + kernel, no host gap
            a.record()
            rc = _fn(*args)
            b.record()
            RECORDS.append((key_of(_name, args), a, b))
            return rc
        setattr(lib, name, timed)


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    pkg.conv_blocks.set_precision("kernel")
    torch.manual_seed(0)
    solver = pkg.AdvancedTripletReconSegmentationModel('FCN_16_standard', num_classes=4, learning_rate=1e-4)
    trainer = pkg.CooperativeTrainer(solver, batch, seed=0, image_cfg=bench.IMAGE_CFG, seg_cfg=bench.SEG_CFG)
    img, lab = bench.synthetic_batch(batch, 224, seed=1000, device="cuda")
    for _ in range(2):
        trainer.step(img, lab)
    instrument()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        trainer.step(img, lab)
    t1.record()
    torch.cuda.synchronize()
    total = t0.elapsed_time(t1) * 1e3 / steps
    agg = collections.defaultdict(lambda: [0, 0.0])
    for key, a, b in RECORDS:
        agg[key][0] += 1
        agg[key][1] += a.elapsed_time(b) * 1e3
    ours = sum(v[1] for v in agg.values()) / steps
    print("eager step %.1f us; inside C-ABI calls %.1f us (%.1f %%); torch / gaps %.1f us" %
          (total, ours, 100 * ours / total, total - ours))
    print("%-46s %6s %10s %7s %8s" % ("call", "n/step", "us/step", "share%", "avg us"))
    for key, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-46s %6.1f %10.1f %7.2f %8.1f" % (key, n / steps, us / steps, 100 * us / steps / total, us / n))


if __name__ == "__main__":
    main()
