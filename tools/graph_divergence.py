"""Diagnostic: where does a CUDA-graph replay of the cooperative step start to differ from the eager step?
Runs the tests/test_graph_gpu.py scenario (4 x 64x64, lr 1e-3, 10 steps) as: eager trainer twice (run-to-run noise),
graphed trainer with every step eager, and graphed trainers that switch to replay after 8 / 5 / 2 eager steps.  Prints one
loss per step and run; the first step at which a replay run leaves the eager band names the faulty phase
(same step = forward of the replay, next step = its backward / optimizer)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cooperative_training_and_latent_space_data_augmentation_b200 as pkg  # noqa: E402
from oracle import weights  # noqa: E402   (diagnostic tool, not the product path)

CFG = ({"loss_name": "mse", "mask_type": "channel", "max_threshold": 0.5, "random_threshold": True, "if_soft": True},
       {"loss_name": "ce", "mask_type": "spatial", "max_threshold": 0.5, "random_threshold": True, "if_soft": True})
PERT = []
KEYS = ('loss', 'loss/standard/seg', 'loss/standard/shape', 'loss/hard/seg', 'loss/hard/image')


def run(cls, steps, lr, **kw):
    torch.manual_seed(0)
    solver = pkg.AdvancedTripletReconSegmentationModel('FCN_16_standard', num_classes=4, learning_rate=lr)
    for name, m in solver.model.items():
        m.load_state_dict(weights.synthetic_state_dict(m, 5, prefix=name + "."))
    solver.set_optimizers(capturable=True)
    trainer = cls(solver, 4, seed=3, image_cfg=CFG[0], seg_cfg=CFG[1], **kw)
    img, lab, noise = weights.synthetic_batch(4, 64, 64, seed=2)
    img, lab, noise = img.cuda(), lab.cuda(), noise.cuda()
    rows, pert = [], []
    for _ in range(steps):
        out = trainer.step(img, lab, noise)
        rows.append([float(out[k]) for k in KEYS])
        pert.append((out['perturbed_image'].float().clone(), out['perturbed_seg'].float().clone()))
    torch.cuda.synchronize()
    PERT.append(pert)
    return rows


def main():
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    lr = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-3
    pkg.conv_blocks.set_precision("kernel")
    runs = [("eager A", run(pkg.CooperativeTrainer, steps, lr)), ("eager B", run(pkg.CooperativeTrainer, steps, lr))]
    runs.append(("eager C", run(pkg.CooperativeTrainer, steps, lr)))
    for e in (steps, 8, 5, 2, 2, 2):
        runs.append(("graphed, %d eager steps" % e, run(pkg.GraphedCooperativeTrainer, steps, lr, eager_steps=e)))
    for ki, key in enumerate(KEYS):
        print("== %s (lr %g)" % (key, lr))
        for name, rows in runs:
            print("%-26s" % name, " ".join("%9.5f" % r[ki] for r in rows))
    # perturbed examples against run 0 (eager A): per step, the MAX over the samples of the relative L2 difference
    for which, label in ((0, "perturbed image"), (1, "perturbed seg")):
        print("== %s: max-over-samples relative difference to eager A" % label)
        for (name, _), pert in zip(runs[1:], PERT[1:]):
            vals = []
            for a, b in zip(pert, PERT[0]):
                d = (a[which] - b[which]).flatten(1).norm(dim=1) / b[which].flatten(1).norm(dim=1).clamp_min(1e-6)
                vals.append(float(d.max()))
            print("%-26s" % name, " ".join("%9.5f" % v for v in vals))


if __name__ == "__main__":
    main()

