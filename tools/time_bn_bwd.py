"""One BatchNorm + activation backward (reduce -> totals -> apply) and one forward-statistics pass at the 16-channel 224^2
and 32-channel 112^2 levels (batch 64), for `ncu --metrics gpu__time_duration.sum` launch lists."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cooperative_training_and_latent_space_data_augmentation_b200 as pkg  # noqa: E402

ops = pkg.ops
for C, S in ((16, 224), (32, 112)):
    a = ops.nchw_to_c8(torch.randn(64, C, S, S, device="cuda"))
    dy = ops.nchw_to_c8(torch.randn(64, C, S, S, device="cuda") * 0.1)
    gamma = torch.ones(C, device="cuda")
    beta = torch.zeros(C, device="cuda")
    scale, shift, mean, var = ops.bn_batch_affine_c8(a, gamma, beta, 1e-5, want_stats=True)
    for _ in range(3):
        totals = torch.zeros(2 * C, device="cuda", dtype=torch.float64)
        ops.bn_act_bwd_c8(dy, None, a, ops.ACT_LRELU, mean, var, 1e-5, gamma, act_affine=(scale, shift), totals=totals)
        ops.scale_shift_act_c8(a, scale, shift, ops.ACT_LRELU)
torch.cuda.synchronize()
