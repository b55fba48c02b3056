"""Launches the K3 conv kernel (and K3w weight gradient) on a few layer classes at batch 64 (BASELINE.json configs[1]
shapes); meant to be wrapped by ncu (profiles/README.md).  argv: reps [layer filter: cin,cout,k,size ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cooperative_training_and_latent_space_data_augmentation_b200 as pkg  # noqa: E402

ops = pkg.ops
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
LAYERS = [(16, 16, 3, 224), (16, 16, 1, 224), (32, 32, 3, 112), (64, 64, 3, 56), (128, 128, 3, 28)]
if len(sys.argv) > 2:
    LAYERS = [tuple(int(v) for v in a.split(",")) for a in sys.argv[2:]]
B = 64
for cin, cout, k, size in LAYERS:
    x = ops.nchw_to_c8(torch.randn(B, cin, size, size, device="cuda"))
    dy = ops.nchw_to_c8(torch.randn(B, cout, size, size, device="cuda") * 0.1)
    w = ops.pack_conv_weight(torch.randn(cout, cin, k, k, device="cuda") * 0.05)
    shift = torch.randn(cout, device="cuda")
    for _ in range(reps):
        y = ops.conv2d_c8(x, w, cout, k * k, shift=shift, act=ops.ACT_LRELU)
        dw = ops.conv_wgrad_c8(x, dy, k * k, layout='conv')
    torch.cuda.synchronize()
print("done")
