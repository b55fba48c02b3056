"""Profiling by elimination for the tcgen05 conv kernels (K3 forward / K3w weight gradient).

Each layer class is timed with stages of the kernel pipeline switched off through the CTL_DIAG_SKIP environment
variable (a `make DIAG=1` build only; csrc/ctl_runtime.cu: 1 = no MMA issue, 2 = no TMA loads, 4 = no epilogue memory traffic (K3w: no epilogue),
8 = no epilogue at all): whichever removal makes the time collapse names the bounding stage.  Outputs with any bit set are garbage;
the tool never checks them.  CUDA-event timing, L2 flushed between launches.  usage: python tools/diag_conv.py [B]"""
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cooperative_training_and_latent_space_data_augmentation_b200 as pkg  # noqa: E402

ops = pkg.ops
LAYERS = [(16, 16, 3, 224), (16, 16, 1, 224), (32, 32, 3, 112), (64, 64, 3, 56), (128, 128, 3, 28), (128, 128, 3, 14)]
FWD_FLAGS = [0, 1, 2, 4, 8, 1 | 4, 2 | 4, 1 | 2, 1 | 2 | 8]
WG_FLAGS = [0, 1, 2, 3, 4, 7]


def timed(fn, flush, iters=8):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return round(statistics.median(ts), 1)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for cin, cout, k, size in LAYERS:
        x = ops.nchw_to_c8(torch.randn(B, cin, size, size, device="cuda"))
        dy = ops.nchw_to_c8(torch.randn(B, cout, size, size, device="cuda") * 0.1)
        res = ops.nchw_to_c8(torch.randn(B, cout, size, size, device="cuda"))
        w = ops.pack_conv_weight(torch.randn(cout, cin, k, k, device="cuda") * 0.05)
        shift = torch.randn(cout, device="cuda")
        row = {"layer": "%d->%d %dx%d @%d B%d" % (cin, cout, k, k, size, B), "fwd_us": {}, "fwd_res_us": {}, "wgrad_us": {},
               "wgrad_kernel_layout_us": {}}
        for f in FWD_FLAGS:
            os.environ["CTL_DIAG_SKIP"] = str(f)
            row["fwd_us"][f] = timed(lambda: ops.conv2d_c8(x, w, cout, k * k, shift=shift, act=ops.ACT_LRELU), flush)
        for f in (0, 4):
            os.environ["CTL_DIAG_SKIP"] = str(f)
            row["fwd_res_us"][f] = timed(lambda: ops.conv2d_c8(x, w, cout, k * k, shift=shift, res=res, act=ops.ACT_LRELU), flush)
        for f in WG_FLAGS:
            os.environ["CTL_DIAG_SKIP"] = str(f)
            row["wgrad_us"][f] = timed(lambda: ops.conv_wgrad_c8(x, dy, k * k, layout='conv'), flush)
        for f in (0, 3):                    # [taps][Cin][Cout] output: direct 16-byte reductions instead of the staged block
            os.environ["CTL_DIAG_SKIP"] = str(f)
            row["wgrad_kernel_layout_us"][f] = timed(lambda: ops.conv_wgrad_c8(x, dy, k * k, layout='kernel'), flush)
        os.environ["CTL_DIAG_SKIP"] = "0"
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
