"""K3 microbench: every conv layer class of FCN_16_standard at batch 64 / 224x224 input (BASELINE.json configs[1]).
Reports time, TFLOP/s (tensor roofline) and GB/s over algorithmic bytes (bf16 in + out), next to cuDNN's
bf16 channels_last convolution of the same layer.  CUDA-event timing, L2 flushed between iterations."""
import json
import os
import statistics
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import cooperative_training_and_latent_space_data_augmentation_b200 as pkg  # noqa: E402

ops = pkg.ops
LAYERS = [(16, 16, 3, 1, 224), (16, 16, 1, 1, 224), (16, 16, 3, 2, 224), (16, 32, 3, 1, 112), (32, 32, 3, 1, 112),
          (32, 32, 3, 2, 112), (32, 64, 3, 1, 56), (64, 64, 3, 1, 56), (64, 128, 3, 1, 28), (128, 128, 3, 1, 28),
          (128, 128, 3, 1, 14), (128, 64, 3, 1, 28), (64, 32, 3, 1, 56), (32, 16, 3, 1, 112), (128, 64, 1, 1, 28)]


def timed(fn, flush, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return statistics.mean(ts)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    peaks = bench.measured_peaks()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    rows = []
    for cin, cout, k, stride, size in LAYERS:
        x = torch.randn(B, cin, size, size, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        w = torch.randn(cout, cin, k, k, device="cuda") * 0.05
        wp = ops.pack_conv_weight(w)
        wb = w.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        shift = torch.randn(cout, device="cuda")
        xc = ops.nchw_to_c8(x.contiguous())
        t_mine = timed(lambda: ops.conv2d_c8(xc, wp, cout, k * k, subsample=stride, shift=shift, act=ops.ACT_LRELU), flush)
        t_lib = timed(lambda: F.leaky_relu(F.conv2d(x, wb, shift.to(torch.bfloat16), stride=stride, padding=k // 2), 0.2), flush)
        so = size // stride
        flops = 2.0 * B * so * so * cout * cin * k * k
        byts = 2.0 * B * (size * size * cin + so * so * cout)
        rows.append({"layer": "%d->%d %dx%d s%d @%d" % (cin, cout, k, k, stride, size), "us": round(t_mine * 1e6, 1),
                     "TFLOPs": round(flops / t_mine / 1e12, 1), "GBps": round(byts / t_mine / 1e9, 0),
                     "hbm_frac": round(byts / t_mine / 1e9 / peaks["hbm_gbs"], 3),
                     "cudnn_conv+lrelu_us": round(t_lib * 1e6, 1), "speedup_vs_cudnn": round(t_lib / t_mine, 2)})
        print(json.dumps(rows[-1]), flush=True)


if __name__ == "__main__":
    main()
