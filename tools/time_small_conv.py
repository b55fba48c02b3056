"""Times the 16 -> 16 channel 3x3 convolution (batch 64 @224^2; plain / statistics / residual / BatchNorm-backward
epilogues) on K3s (warp-level tensor path) and, with CTL_CONV_SMALL=0, on the tcgen05 kernel.  CUDA-event timing of a
CUDA-graph replay of 20 calls with inputs cycling over 3 buffer sets (> L2).  usage: [CTL_CONV_SMALL=0] python tools/time_small_conv.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cooperative_training_and_latent_space_data_augmentation_b200 as pkg  # noqa: E402

ops = pkg.ops


def main():
    C, S = 16, int(sys.argv[1]) if len(sys.argv) > 1 else 224
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    sets = []
    for _ in range(3):
        sets.append(dict(x=ops.nchw_to_c8(torch.randn(B, C, S, S, device="cuda")),
                         a=ops.nchw_to_c8(torch.randn(B, C, S, S, device="cuda"))))
    w = ops.pack_conv_weight(torch.randn(C, C, 3, 3, device="cuda") * 0.05)
    shift = torch.randn(C, device="cuda")
    scale = torch.rand(C, device="cuda") + 0.5
    stats = torch.zeros(2, C, device="cuda", dtype=torch.float64)
    variants = {
        "plain": lambda d: ops.conv2d_c8(d["x"], w, C, 9, shift=shift, act=ops.ACT_LRELU),
        "stats": lambda d: ops.conv2d_c8(d["x"], w, C, 9, shift=shift, stats=stats),
        "res": lambda d: ops.conv2d_c8(d["x"], w, C, 9, res=d["a"]),
        "bnbwd": lambda d: ops.conv2d_c8_bnbwd(d["x"], w, C, d["a"], scale, shift, ops.ACT_LRELU, stats.view(-1)),
    }
    only = os.environ.get("ONLY")
    if only:
        variants = {k: v for k, v in variants.items() if k in only.split(",")}
    out = {"size": S, "batch": B, "small": os.environ.get("CTL_CONV_SMALL", "1"), "diag": os.environ.get("CTL_DIAG_SKIP", "0")}
    st = torch.cuda.Stream()
    for name, fn in variants.items():
        with torch.cuda.stream(st):
            for d in sets:
                fn(d)
            st.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                for i in range(21):
                    fn(sets[i % 3])
            g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ts = []
            for _ in range(5):
                e0.record(st); g.replay(); e1.record(st)
                st.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3 / 21)
        out[name + "_us"] = round(sorted(ts)[2], 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
