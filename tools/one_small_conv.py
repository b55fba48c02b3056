"""One launch of each K3s epilogue form (16 -> 16 channel 3x3, batch 64 @224^2) for an ncu capture:
ncu --set full --import-source on --clock-control none -k regex:conv_small -o out python tools/one_small_conv.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cooperative_training_and_latent_space_data_augmentation_b200 as pkg  # noqa: E402

ops = pkg.ops
x = ops.nchw_to_c8(torch.randn(64, 16, 224, 224, device="cuda"))
a = ops.nchw_to_c8(torch.randn(64, 16, 224, 224, device="cuda"))
w = ops.pack_conv_weight(torch.randn(16, 16, 3, 3, device="cuda") * 0.05)
shift = torch.randn(16, device="cuda")
scale = torch.rand(16, device="cuda") + 0.5
stats = torch.zeros(2, 16, device="cuda", dtype=torch.float64)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(2):
    flush.zero_()
    ops.conv2d_c8(x, w, 16, 9, shift=shift, act=ops.ACT_LRELU)
    flush.zero_()
    ops.conv2d_c8(x, w, 16, 9, shift=shift, stats=stats)
    flush.zero_()
    ops.conv2d_c8(x, w, 16, 9, res=a)
    flush.zero_()
    ops.conv2d_c8_bnbwd(x, w, 16, a, scale, shift, ops.ACT_LRELU, stats.view(-1))
torch.cuda.synchronize()
