"""CPU: the oracle restatement replayed against fixtures produced by the unmodified reference
(oracle/make_golden.py).  This is what pins the oracle (the reference has no tests of its own)."""
import glob
import os
import random

import numpy as np
import pytest
import torch

from oracle import masking_oracle as mo
from oracle import model_oracle, weights
from oracle.make_golden import MASK_CASES, mask_inputs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _probe(a, n):
    a = np.asarray(a).reshape(-1)
    return a[:: max(1, a.size // n)][:n]


@pytest.mark.parametrize("case", [c[0] for c in MASK_CASES])
def test_masking_matches_reference(case):
    f = np.load(os.path.join(GOLDEN, "masking_%s.npz" % case))
    N, C, H, W = [int(x) for x in f["shape"]]
    mode = mo.MODE_CHANNEL if str(f["mode"]) == "channel" else mo.MODE_SPATIAL
    z, g0 = mask_inputs(N, C, H, W, int(f["seed"]))
    # the gradient the reference saw: label * fp32(1/numel) (identity decoder + 'corr' loss; mean backward)
    g = ((g0 * np.float32(z.size)) * np.float32(1.0 / z.size)).astype(np.float32)
    np.testing.assert_array_equal(_probe(g, 64)[:64], f["g_ref_probe"])
    k = int(f["k"])
    soft = bool(f["soft"])
    rand = f["rand"] if soft else None
    # (1) reference summation order -> every bit of mask and masked code must match
    s_ref = mo.saliency_reduce_reference_order(g, mode)
    zt, m, _, _ = mo.mask_given_gradient(z, g, mode, k, soft, rand, s=s_ref)
    assert tuple(m.shape) == tuple(int(x) for x in f["mask_shape"])
    np.testing.assert_array_equal(m, f["mask"])
    np.testing.assert_array_equal(_probe(zt, 257)[:257], f["masked_probe"])
    assert np.float64(zt.astype(np.float64).sum()) == f["masked_checksum"]
    assert int((m != 1).reshape(N, -1).sum(1).max()) == k and int((m != 1).reshape(N, -1).sum(1).min()) == k
    # (2) order-independent f64 saliency: same mask here (no near-ties in these fixtures)
    zt2, m2, s64, thr = mo.mask_given_gradient(z, g, mode, k, soft, rand)
    np.testing.assert_allclose(s64, s_ref, rtol=2e-5, atol=1e-12)
    np.testing.assert_array_equal(m2, f["mask"])


def test_threshold_index_semantics():
    # SURVEY.md section 4 items 1-3
    assert [mo.threshold_index(128, p)[0] for p in (0.1, 0.3, 0.5)] == [12, 38, 64]
    assert mo.threshold_index(128, 0.005)[0] == 0
    s = np.random.RandomState(0).standard_normal((3, 16)).astype(np.float32)
    thr = mo.topp_threshold(s, 0)
    assert (mo.build_mask(s, thr) == 1).all()                 # k = 0 -> nothing masked
    err = str(np.load(os.path.join(GOLDEN, "masking_errors.npz"))["p1_error"])
    assert err == "IndexError"
    with pytest.raises(IndexError):
        mo.topp_threshold(s, mo.threshold_index(16, 1.0)[0])
    # ranking is on the SIGNED mean (item 4)
    s2 = np.array([[-5.0, 1.0, 0.5, -0.1]], np.float32)
    assert mo.build_mask(s2, mo.topp_threshold(s2, 1)).tolist() == [[1.0, 0.0, 1.0, 1.0]]


@pytest.mark.parametrize("case", ["drop_p50", "drop_p30", "drop_p0"])
def test_dropout_matches_reference(case):
    f = np.load(os.path.join(GOLDEN, "masking_%s.npz" % case))
    N, C, H, W = [int(x) for x in f["shape"]]
    z, _ = mask_inputs(N, C, H, W, int(f["seed"]))
    out, mask = mo.channel_dropout(z, f["keep"], float(f["p"]))
    assert tuple(mask.shape) == tuple(int(x) for x in f["mask_shape"]) == (N, C, H, W)
    np.testing.assert_array_equal(_probe(out, 257)[:257], f["masked_probe"])
    assert np.float64(out.astype(np.float64).sum()) == f["masked_checksum"]
    assert np.float64(mask.astype(np.float64).sum()) == f["mask_checksum"]
    if float(f["p"]) > 0:       # the quirk: mask == (z == 0), not the dropout pattern
        np.testing.assert_array_equal(mask, (z == 0).astype(np.float32))


def test_philox_known_answer():
    # Random123 known-answer vectors for philox4x32-10
    out = mo.philox4x32_10(np.zeros((1, 4), np.uint32), np.zeros((1, 2), np.uint32))[0]
    assert [hex(int(x)) for x in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    ones = np.full((1, 4), 0xFFFFFFFF, np.uint32)
    out = mo.philox4x32_10(ones, np.full((1, 2), 0xFFFFFFFF, np.uint32))[0]
    assert [hex(int(x)) for x in out] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    pi_ctr = np.array([[0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344]], np.uint32)
    pi_key = np.array([[0xa4093822, 0x299f31d0]], np.uint32)
    out = mo.philox4x32_10(pi_ctr, pi_key)[0]
    assert [hex(int(x)) for x in out] == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]
    u = mo.native_rand(123, 5, 4, 100)
    assert u.min() >= 0 and u.max() < 1
    # shard invariance: rows 2..3 of the full draw == a shard starting at sample 2
    np.testing.assert_array_equal(u[2:], mo.native_rand(123, 5, 2, 100, first_sample=2))


def _seed_all(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


@pytest.mark.parametrize("fixture", ["model_step.npz", "model_step_224.npz"])
def test_model_oracle_matches_reference_fixture(fixture):
    # fixtures were generated single-threaded: oneDNN's summation order depends on the thread
    # count, and a 1-ulp change of dL/dz is enough to flip a top-k near-tie in step 1.
    # model_step_224.npz is BASELINE.json configs[0] (batch 8 of 1x224x224, one step; probes / checksums only).
    prev = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        _model_fixture_body(fixture)
    finally:
        torch.set_num_threads(prev)


def _model_fixture_body(fixture):
    f = np.load(os.path.join(GOLDEN, fixture))
    full = fixture == "model_step.npz"
    N, H, W = int(f["N"]), int(f["H"]), int(f["W"])
    solver = model_oracle.OracleSolver(num_classes=4, learning_rate=1e-4)
    for k, m in solver.model.items():
        m.load_state_dict(weights.synthetic_state_dict(m, int(f["weight_seed"]), prefix=k + "."))
    img, lab, noise = weights.synthetic_batch(N, H, W, seed=int(f["data_seed"]))

    solver.eval()
    with torch.no_grad():
        z_i, z_s = solver.model["image_encoder"](img)
        seg = solver.model["segmentation_decoder"](z_s)
    np.testing.assert_allclose(z_i.numpy() if full else _probe(z_i.numpy(), 4096)[:4096], f["eval_z_i"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(z_s.numpy() if full else _probe(z_s.numpy(), 4096)[:4096], f["eval_z_s"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(float(z_i.double().sum()), float(f["eval_z_i_sum"]), rtol=1e-5)
    np.testing.assert_allclose(_probe(seg.numpy(), 4096)[:4096], f["eval_seg"], rtol=1e-4, atol=1e-4)
    pred2 = solver.predict(img, n_iter=2)
    np.testing.assert_allclose(_probe(pred2.numpy(), 4096)[:4096], f["eval_pred2"], rtol=1e-4, atol=1e-4)
    hist = np.bincount(pred2.max(1)[1].numpy().reshape(-1), minlength=4)
    assert np.abs(hist - f["eval_pred2_labels_hist"]).sum() <= 1e-4 * hist.sum()       # arg-max label map of predict

    cfg_i = {"loss_name": "mse", "mask_type": "channel", "max_threshold": 0.5, "random_threshold": True, "if_soft": True}
    cfg_s = {"loss_name": "ce", "mask_type": "spatial", "max_threshold": 0.5, "random_threshold": True, "if_soft": True}
    _seed_all(5)
    n_steps = sum(1 for k in f.files if k.endswith("_loss") and k.startswith("step"))
    for step in range(n_steps):
        r = solver.cooperative_step(img, lab, cfg_i, cfg_s, noise=noise)
        std = [r["standard/seg"], r["standard/image"], r["standard/gt_shape"], r["standard/shape"]]
        hard = [r["hard/seg"], r["hard/image"], r["hard/shape"], r["hard/perturbed_shape"]]
        np.testing.assert_allclose([float(x) for x in std], f["step%d_standard" % step], rtol=2e-6)
        np.testing.assert_allclose([float(x) for x in hard], f["step%d_hard" % step], rtol=2e-6)
        np.testing.assert_allclose(float(r["loss"]), float(f["step%d_loss" % step]), rtol=2e-6)
        np.testing.assert_allclose(_probe(r["perturbed_image"].numpy(), 4096)[:4096], f["step%d_p_img" % step],
                                   rtol=1e-3, atol=1e-4)
        np.testing.assert_allclose(_probe(r["perturbed_seg"].numpy(), 4096)[:4096], f["step%d_p_seg" % step],
                                   rtol=1e-3, atol=2e-3)
    for k, m in solver.model.items():
        tracked = [int(b) for n_, b in m.named_buffers() if n_.endswith("num_batches_tracked")]
        assert tracked == list(f["final_bn_tracked_" + k]), k
        psum = sum(float(p.double().sum()) for p in m.parameters())
        np.testing.assert_allclose(psum, float(f["final_param_sum_" + k]), rtol=1e-5, atol=1e-3)
        rm = sum(float(b.double().sum()) for n_, b in m.named_buffers() if n_.endswith("running_mean"))
        np.testing.assert_allclose(rm, float(f["final_running_mean_sum_" + k]), rtol=1e-4, atol=1e-4)


def test_golden_files_present():
    assert len(glob.glob(os.path.join(GOLDEN, "*.npz"))) >= 17
