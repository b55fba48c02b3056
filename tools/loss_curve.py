"""Loss trajectory of the cooperative step in a given precision mode (kernel / bf16 / fp32) on a fixed synthetic batch:
the product path must train like the library path.  argv: mode steps batch size"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import cooperative_training_and_latent_space_data_augmentation_b200 as pkg  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "kernel"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 16
size = int(sys.argv[4]) if len(sys.argv) > 4 else 224
pkg.conv_blocks.set_precision(mode)
torch.manual_seed(0)
solver = pkg.AdvancedTripletReconSegmentationModel('FCN_16_standard', num_classes=4, learning_rate=1e-4)
trainer = pkg.CooperativeTrainer(solver, batch, seed=0, image_cfg=bench.IMAGE_CFG, seg_cfg=bench.SEG_CFG)
img, lab = bench.synthetic_batch(batch, size, seed=1000, device="cuda")
out = []
for i in range(steps):
    r = trainer.step(img, lab)
    out.append((round(float(r['loss']), 4), round(float(r['loss/standard/total']), 4), round(float(r['loss/hard/total']), 4)))
print(mode, out)
gn = {k: round(float(torch.cat([p.grad.reshape(-1) for p in m.parameters()]).norm()), 5) for k, m in solver.model.items()}
print(mode, "last grad norms", gn)
