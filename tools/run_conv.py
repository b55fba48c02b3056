"""Launches the K3 conv kernel on a few layer classes at batch 64 (BASELINE.json configs[1] shapes); meant to be wrapped
by ncu (profiles/README.md).  argv: reps"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cooperative_training_and_latent_space_data_augmentation_b200 as pkg  # noqa: E402

ops = pkg.ops
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
LAYERS = [(16, 16, 3, 224), (16, 16, 1, 224), (32, 32, 3, 112), (64, 64, 3, 56), (128, 128, 3, 28)]
B = 64
for cin, cout, k, size in LAYERS:
    x = ops.nchw_to_c8(torch.randn(B, cin, size, size, device="cuda"))
    w = ops.pack_conv_weight(torch.randn(cout, cin, k, k, device="cuda") * 0.05)
    shift = torch.randn(cout, device="cuda")
    for _ in range(reps):
        y = ops.conv2d_c8(x, w, cout, k * k, shift=shift, act=ops.ACT_LRELU)
    torch.cuda.synchronize()
print("done")
