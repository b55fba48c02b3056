"""Launches each masking kernel a few times at BASELINE.json configs[3] ([512,64,28,28] fp32) and at the
real-model shape ([64,128,14,14]); meant to be wrapped by ncu (see profiles/README.md)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cooperative_training_and_latent_space_data_augmentation_b200 as pkg  # noqa: E402

ops = pkg.ops
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for shape in ((512, 64, 28, 28), (64, 128, 14, 14)):
    N, C, H, W = shape
    gen = torch.Generator(device="cuda").manual_seed(0)
    z = torch.relu(torch.randn(*shape, device="cuda", generator=gen))
    g = 1e-5 * torch.randn(*shape, device="cuda", generator=gen)
    rng = ops.NativeRNG(0)
    for _ in range(reps):
        ops.saliency_mask_apply(g, z, ops.MODE_CHANNEL, int(C * 0.3), soft=True, rng=rng)
        ops.saliency_mask_apply(g, z, ops.MODE_SPATIAL, int(H * W * 0.3), soft=True, rng=rng)
        ops.channel_dropout(z, 0.5, rng=rng, want_mask=False)
        ops.channel_dropout(z, 0.5, rng=rng, want_mask=True)
    torch.cuda.synchronize()
print("done")
