"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference
(imported from /root/reference with stub modules, oracle/ref_import.py) on CPU.

Run in the build container:  python -m oracle.make_golden
The fixtures are small (inputs are regenerated from seeds; only outputs are stored).
"""
import contextlib
import io
import os
import sys
import random

import numpy as np
import torch

from . import weights
from .ref_import import import_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

MASK_CASES = [
    # name, N, C, H, W, mode, p, random, soft, seed
    ("ch_p30_hard", 4, 128, 14, 14, "channel", 0.3, False, False, 1),
    ("ch_p50_soft", 4, 128, 14, 14, "channel", 0.5, False, True, 2),
    ("ch_rand_soft", 3, 64, 28, 28, "channel", 0.5, True, True, 3),
    ("ch_k0", 2, 128, 14, 14, "channel", 0.005, False, False, 4),
    ("sp_p10_hard", 4, 128, 14, 14, "spatial", 0.1, False, False, 5),
    ("sp_p40_soft", 2, 64, 28, 28, "spatial", 0.4, False, True, 6),
    ("sp_rand_soft", 4, 128, 16, 16, "spatial", 0.5, True, True, 7),
    ("sp_n1", 1, 128, 12, 12, "spatial", 0.2, False, False, 8),
    ("ch_odd", 5, 24, 7, 9, "channel", 1 / 3.0, False, True, 9),
    ("sp_odd", 5, 24, 7, 9, "spatial", 1 / 3.0, False, False, 10),
]


def mask_inputs(N, C, H, W, seed):
    """z = relu(N(0,1)), g = 1e-5 * N(0,1)  (SURVEY.md 8d microbench distribution)."""
    rs = np.random.RandomState(1000 + seed)
    z = np.maximum(rs.standard_normal((N, C, H, W)), 0).astype(np.float32)
    g = (1e-5 * rs.standard_normal((N, C, H, W))).astype(np.float32)
    return z, g


def seed_all(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def gen_masking(ref_mu):
    for name, N, C, H, W, mode, p, rnd, soft, seed in MASK_CASES:
        z, g = mask_inputs(N, C, H, W, seed)
        # identity decoder + 'corr' loss + 4-D label  =>  dL/dz = label / numel  (model_util.py:212-214)
        # so an arbitrary gradient g can be pushed through the real reference function.
        numel = float(N * C * H * W)
        label = torch.from_numpy(g) * numel
        zt = torch.from_numpy(z)
        code = zt.clone().requires_grad_(True)
        g_ref = torch.autograd.grad(torch.mean(code * label), [code])[0].numpy()
        seed_all(seed)
        fn = ref_mu.mask_latent_code_channel_wise if mode == "channel" else ref_mu.mask_latent_code_spatial_wise
        masked, mask = fn(zt, lambda c: c, label, num_classes=4, percentile=p, random=rnd,
                          loss_type="corr", if_detach=True, if_soft=soft)
        # replay the RNG streams to record what the reference consumed
        seed_all(seed)
        n = C if mode == "channel" else H * W
        p_eff = np.random.rand() * p if rnd else p
        k = int(n * p_eff)
        rand = torch.rand(N, n).numpy() if soft else np.zeros((0,), np.float32)
        np.savez_compressed(
            os.path.join(OUT, "masking_%s.npz" % name),
            shape=np.array([N, C, H, W]), mode=mode, p=p, random=rnd, soft=soft, seed=seed,
            g_ref_checksum=np.float64(g_ref.astype(np.float64).sum()),
            g_ref_probe=g_ref.reshape(-1)[:: max(1, g_ref.size // 64)][:64],
            k=k, rand=rand.astype(np.float32),
            mask=mask.detach().numpy().astype(np.float32),
            masked_checksum=np.float64(masked.detach().numpy().astype(np.float64).sum()),
            masked_probe=masked.detach().numpy().reshape(-1)[:: max(1, z.size // 257)][:257],
            masked_dtype=str(masked.dtype), mask_shape=np.array(mask.shape))
        print("masking", name, "k=%d" % k, "masked_frac=%.4f" % float((mask.detach().numpy() != 1).mean()))

    # error semantics (SURVEY.md section 4 items 2,3)
    z, g = mask_inputs(2, 16, 4, 4, 99)
    label = torch.from_numpy(g) * float(z.size)
    try:
        ref_mu.mask_latent_code_channel_wise(torch.from_numpy(z), lambda c: c, label, percentile=1.0,
                                             loss_type="corr")
        err = "none"
    except Exception as e:  # noqa: BLE001
        err = type(e).__name__
    np.savez_compressed(os.path.join(OUT, "masking_errors.npz"), p1_error=err)
    print("p=1.0 ->", err)


def gen_dropout(RefSolver):
    with contextlib.redirect_stdout(io.StringIO()):
        solver = RefSolver("FCN_16_standard", num_classes=4, use_gpu=False)
    for name, N, C, H, W, p, seed in [("drop_p50", 4, 128, 14, 14, 0.5, 11), ("drop_p30", 3, 64, 28, 28, 0.3, 12),
                                      ("drop_p0", 2, 16, 4, 4, 0.0, 13)]:
        z, _ = mask_inputs(N, C, H, W, seed)
        seed_all(seed)
        masked, mask = solver.perturb_latent_code(torch.from_numpy(z), None, perturb_type="dropout", threshold=p)
        seed_all(seed)
        keep = torch.empty(N, C, 1, 1).bernoulli_(1 - p).view(N, C).numpy() if p > 0 else np.ones((N, C), np.float32)
        mk = masked.numpy()
        np.savez_compressed(os.path.join(OUT, "masking_%s.npz" % name), shape=np.array([N, C, H, W]), p=p, seed=seed,
                            keep=keep.astype(np.float32),
                            masked_checksum=np.float64(mk.astype(np.float64).sum()),
                            masked_probe=mk.reshape(-1)[:: max(1, z.size // 257)][:257],
                            mask_checksum=np.float64(mask.numpy().astype(np.float64).sum()),
                            mask_shape=np.array(mask.shape))
        print("dropout", name, "kept=%.3f" % keep.mean(), "mask_ones=%.3f" % mask.numpy().mean())


def _probe(t, n=512):
    a = t.detach().numpy().reshape(-1)
    return a[:: max(1, a.size // n)][:n].astype(np.float32)


def gen_model(RefSolver, N=2, H=64, W=64, name="model_step.npz", full_latents=True, steps=2):
    """Forward/backward of every sub-network, hard_example_generation with fixed mask types and
    full cooperative steps, all on weights regenerated from oracle/weights.py.  The default is the small
    fixture; `model_step_224.npz` is BASELINE.json configs[0] (batch 8 of 1x224x224), probes / checksums only."""
    wseed = 7
    with contextlib.redirect_stdout(io.StringIO()):
        solver = RefSolver("FCN_16_standard", num_classes=4, use_gpu=False, learning_rate=1e-4)
    for k, m in solver.model.items():
        m.load_state_dict(weights.synthetic_state_dict(m, wseed, prefix=k + "."))
    img, lab, noise = weights.synthetic_batch(N, H, W, seed=21)
    out = {"N": N, "H": H, "W": W, "weight_seed": wseed, "data_seed": 21}

    # ---- eval-mode forward of every module + predict(n_iter=2)
    solver.eval()
    with torch.no_grad():
        z_i, z_s = solver.model["image_encoder"](img)
        seg = solver.model["segmentation_decoder"](z_s)
        rec = solver.model["image_decoder"](z_i)
        pred2 = solver.predict(img, n_iter=2)
    out.update(eval_z_i=z_i.numpy() if full_latents else _probe(z_i, 4096),
               eval_z_s=z_s.numpy() if full_latents else _probe(z_s, 4096),
               eval_z_i_sum=np.float64(z_i.double().sum()), eval_z_s_sum=np.float64(z_s.double().sum()),
               eval_pred2_labels_hist=np.bincount(pred2.max(1)[1].numpy().reshape(-1), minlength=4),
               eval_seg=_probe(seg, 4096), eval_rec=_probe(rec, 4096),
               eval_pred2=_probe(pred2, 4096), eval_seg_sum=np.float64(seg.double().sum()),
               eval_pred2_sum=np.float64(pred2.double().sum()))

    cfg_i = {"loss_name": "mse", "mask_type": "channel", "max_threshold": 0.5, "random_threshold": True, "if_soft": True}
    cfg_s = {"loss_name": "ce", "mask_type": "spatial", "max_threshold": 0.5, "random_threshold": True, "if_soft": True}

    # ---- two cooperative steps (train...triplet.py:171-237), fixed channel(image)+spatial(shape)
    seed_all(5)
    for step in range(steps):
        solver.train()
        solver.reset_all_optimizers()
        noisy = torch.clamp(img + noise, 0, 1)
        s = solver.standard_training(img, lab, perturbed_image=noisy, separate_training=False)
        standard = s[0] + s[1] + s[3] + s[2]
        solver.reset_all_optimizers()
        # record what the two host RNG draws will be (numpy global), then rewind
        st = np.random.get_state()
        tst = torch.get_rng_state()
        p_img, p_seg = solver.hard_example_generation(img.detach().clone(), lab.detach().clone(),
                                                      corrupted_image_DA_config=cfg_i, corrupted_seg_DA_config=cfg_s)
        h = solver.hard_example_training(perturbed_image=p_img, perturbed_seg=p_seg, clean_image_l=img,
                                         label_l=lab, separate_training=False, use_gpu=False)
        hard = h[0] + h[1] + h[2] + h[3]
        loss = standard + hard
        solver.reset_all_optimizers()
        loss.backward()
        gn = {}
        for k, m in solver.model.items():
            gn[k] = np.array([float(p.grad.double().norm()) if p.grad is not None else -1.0 for p in m.parameters()])
        solver.optimize_all_params()
        out.update({
            "step%d_standard" % step: np.array([float(x) for x in s]),
            "step%d_hard" % step: np.array([float(x) for x in h]),
            "step%d_loss" % step: np.float64(loss.item()),
            "step%d_p_img" % step: _probe(p_img, 4096), "step%d_p_seg" % step: _probe(p_seg, 4096),
            "step%d_p_img_sum" % step: np.float64(p_img.double().sum()),
            "step%d_p_seg_sum" % step: np.float64(p_seg.double().sum()),
            "step%d_np_state_pos" % step: np.int64(st[2]),
        })
        for k in gn:
            out["step%d_gradnorm_%s" % (step, k)] = gn[k]
    # parameters after the two Adam steps: one checksum per module
    for k, m in solver.model.items():
        out["final_param_sum_" + k] = np.float64(sum(float(p.double().sum()) for p in m.parameters()))
        out["final_bn_tracked_" + k] = np.array([int(b) for n_, b in m.named_buffers() if n_.endswith("num_batches_tracked")])
        out["final_running_mean_sum_" + k] = np.float64(sum(float(b.double().sum()) for n_, b in m.named_buffers() if n_.endswith("running_mean")))
    np.savez_compressed(os.path.join(OUT, name), **out)
    print("model:", name, "losses", [float(out["step%d_loss" % i]) for i in range(steps)])


def gen_metrics():
    """runningScore (medseg/common_utils/metrics.py:12-57) of the unmodified reference on seeded label maps, including
    out-of-range true labels (ignored by _fast_hist's mask) and a class that never occurs (NaN IoU -> nanmean)."""
    from medseg.common_utils.metrics import runningScore
    rs = np.random.RandomState(11)
    cases = {}
    for name, n_cls, absent in (("a", 4, None), ("b", 4, 3), ("c", 2, None)):
        m = runningScore(n_cls)
        gts, preds = [], []
        for _ in range(3):
            gt = rs.randint(0, n_cls, size=(5, 32, 48)).astype(np.int64)
            pr = np.where(rs.rand(5, 32, 48) < 0.7, gt, rs.randint(0, n_cls, size=(5, 32, 48))).astype(np.int64)
            if absent is not None:
                gt[gt == absent] = 0
                pr[pr == absent] = 1
            gt[0, :2, :3] = 255 if name == "a" else gt[0, :2, :3]     # out-of-range labels are skipped
            m.update(gt, pr)
            gts.append(gt); preds.append(pr)
        scores, cls_iu = m.get_scores()
        cases.update({name + "_n": n_cls, name + "_gt": np.stack(gts).astype(np.uint8), name + "_pred": np.stack(preds).astype(np.uint8),
                      name + "_hist": m.confusion_matrix, name + "_scores": np.array(list(scores.values()), np.float64),
                      name + "_score_keys": np.array(list(scores.keys())),
                      name + "_cls_iu": np.array([cls_iu[i] for i in range(n_cls)], np.float64)})
    np.savez_compressed(os.path.join(OUT, "metrics_scores.npz"), **cases)
    print("metrics: mean IoU", cases["a_scores"][3], cases["b_scores"][3], cases["c_scores"][3])


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)          # fixtures must not depend on the thread count of the box
    ref_mu, RefSolver = import_reference()
    only = sys.argv[1:]
    if not only or "masking" in only:
        gen_masking(ref_mu)
        gen_dropout(RefSolver)
    if not only or "model" in only:
        gen_model(RefSolver)
    if not only or "model224" in only:
        # BASELINE.json configs[0]: batch 8 of 1x224x224 -- one cooperative step (CPU, single thread: minutes)
        gen_model(RefSolver, N=8, H=224, W=224, name="model_step_224.npz", full_latents=False, steps=1)
    if not only or "metrics" in only:
        gen_metrics()


if __name__ == "__main__":
    main()
