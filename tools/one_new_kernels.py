"""One launch of each kernel added late in round 2 (batch 64 @224^2) for an ncu capture:
ncu --set full --clock-control none -k regex:'stem_dgrad_small|apply_planes|reduce_items|conv_small_s2|act_planes' ..."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cooperative_training_and_latent_space_data_augmentation_b200 as pkg  # noqa: E402

ops = pkg.ops
N, S, C = 64, 224, 16
a = ops.nchw_to_c8(torch.randn(N, C, S, S, device="cuda"))
dy = ops.nchw_to_c8(torch.randn(N, C, S, S, device="cuda") * 0.1)
gamma, beta = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
scale, shift, mean, var = ops.bn_batch_affine_c8(a, gamma, beta, 1e-5, want_stats=True)
h = ops.scale_shift_act_c8(a, scale, shift, ops.ACT_LRELU)
w = torch.randn(16, 4, 3, 3, device="cuda") * 0.1
x4 = torch.randn(N, 4, S, S, device="cuda")
ws2 = ops.pack_conv_weight_s2(torch.randn(16, 16, 3, 3, device="cuda") * 0.05)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(2):
    flush.zero_()
    totals = torch.zeros(2 * C, device="cuda", dtype=torch.float64)
    ops.bn_act_bwd_c8(dy, None, a, ops.ACT_LRELU, mean, var, 1e-5, gamma, act_affine=(scale, shift), totals=totals)
    flush.zero_()
    totals = torch.zeros(2 * C, device="cuda", dtype=torch.float64)
    ops.bn_act_bwd_c8(dy, h, a, ops.ACT_LRELU, mean, var, 1e-5, gamma, want_dv=True, totals=totals)
    flush.zero_()
    ops.scale_shift_act_c8(a, scale, shift, ops.ACT_LRELU)
    flush.zero_()
    ops.stem_dgrad_c8(dy, x4, w, in_mode=1, temperature=2.0)
    flush.zero_()
    ops.conv2d_c8(a, ws2, 16, 9, subsample=2, act=ops.ACT_LRELU)
torch.cuda.synchronize()
