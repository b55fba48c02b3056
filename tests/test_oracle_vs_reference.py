"""CPU, build container only: runs the oracle restatement and the UNMODIFIED imported reference side by side on fresh
seeds (not the committed fixtures).  Skipped where /root/reference does not exist (the GPU box)."""
import contextlib
import io
import random

import numpy as np
import pytest
import torch

from oracle import masking_oracle as mo
from oracle import model_oracle, weights
from oracle.ref_import import import_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present")


def _seed(s):
    random.seed(s); np.random.seed(s); torch.manual_seed(s)


@pytest.fixture(scope="module")
def ref():
    return import_reference()


@pytest.mark.parametrize("mode,shape,p,rnd,soft", [("channel", (3, 32, 6, 5), 0.4, True, True),
                                                   ("spatial", (2, 16, 9, 7), 0.25, False, False),
                                                   ("spatial", (1, 8, 5, 5), 0.5, True, True)])
def test_masking_oracle_equals_reference_on_fresh_inputs(ref, mode, shape, p, rnd, soft):
    ref_mu, _ = ref
    N, C, H, W = shape
    rs = np.random.RandomState(123)
    z = np.maximum(rs.standard_normal(shape), 0).astype(np.float32)
    g = (1e-5 * rs.standard_normal(shape)).astype(np.float32)
    label = torch.from_numpy(g) * float(z.size)        # identity decoder + 'corr' loss => dL/dz == g
    fn = ref_mu.mask_latent_code_channel_wise if mode == "channel" else ref_mu.mask_latent_code_spatial_wise
    _seed(17)
    masked, mask = fn(torch.from_numpy(z), lambda c: c, label, num_classes=4, percentile=p, random=rnd,
                      loss_type="corr", if_detach=True, if_soft=soft)
    _seed(17)
    n = C if mode == "channel" else H * W
    k = int(n * (np.random.rand() * p if rnd else p))
    rand = torch.rand(N, n).numpy() if soft else None
    m = mo.MODE_CHANNEL if mode == "channel" else mo.MODE_SPATIAL
    want_z, want_m, _, _ = mo.mask_given_gradient(z, (label / float(z.size)).numpy(), m, k, soft=soft, rand=rand)
    assert np.array_equal(mask.detach().numpy().reshape(want_m.shape), want_m)
    assert np.array_equal(masked.detach().numpy(), want_z)


def test_model_oracle_step_equals_reference(ref):
    _, RefSolver = ref
    prev = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        N, H, W = 2, 32, 32
        with contextlib.redirect_stdout(io.StringIO()):
            rsolver = RefSolver("FCN_16_standard", num_classes=4, use_gpu=False, learning_rate=1e-4)
        osolver = model_oracle.OracleSolver(num_classes=4, learning_rate=1e-4)
        for k, m in rsolver.model.items():
            sd = weights.synthetic_state_dict(m, 3, prefix=k + ".")
            m.load_state_dict(sd)
            osolver.model[k].load_state_dict(sd)
        img, lab, noise = weights.synthetic_batch(N, H, W, seed=4)
        cfg_i = {"loss_name": "mse", "mask_type": "channel", "max_threshold": 0.5, "random_threshold": True, "if_soft": True}
        cfg_s = {"loss_name": "ce", "mask_type": "spatial", "max_threshold": 0.5, "random_threshold": True, "if_soft": True}
        _seed(8)
        rsolver.train(); rsolver.reset_all_optimizers()
        s = rsolver.standard_training(img, lab, perturbed_image=torch.clamp(img + noise, 0, 1), separate_training=False)
        p_img, p_seg = rsolver.hard_example_generation(img.detach().clone(), lab.detach().clone(),
                                                       corrupted_image_DA_config=cfg_i, corrupted_seg_DA_config=cfg_s)
        h = rsolver.hard_example_training(perturbed_image=p_img, perturbed_seg=p_seg, clean_image_l=img, label_l=lab,
                                          separate_training=False, use_gpu=False)
        ref_loss = float(sum(s) + sum(h))
        _seed(8)
        r = osolver.cooperative_step(img, lab, cfg_i, cfg_s, noise=noise, optimize=False)
        np.testing.assert_allclose(float(r["loss"]), ref_loss, rtol=1e-5)
        np.testing.assert_allclose(r["perturbed_image"].numpy(), p_img.numpy(), rtol=1e-4, atol=1e-5)
    finally:
        torch.set_num_threads(prev)
