"""GPU parity of the fused 2-D cross entropy (csrc/loss.cu: ctl_ce2d_fwd / ctl_ce2d_bwd) against the reference's
formulation in torch fp32 -- log_softmax -> nll_loss(sum) / region (medseg/models/custom_loss.py:706-741,
medseg/models/model_util.py:104-135) -- and against the CPU oracle's restatement of it.  Tolerances (fp32, different
summation order): loss 2e-6 relative, gradient 1e-6 of its largest magnitude."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

SHAPES = [(2, 4, 16, 16), (3, 4, 7, 9), (2, 2, 8, 8), (1, 8, 4, 4), (2, 3, 5, 12), (8, 4, 224, 224)]


@pytest.fixture(scope="module")
def pkg():
    import cooperative_training_and_latent_space_data_augmentation_b200 as pkg
    return pkg


@pytest.mark.parametrize("shape", SHAPES)
def test_fused_ce_matches_torch(pkg, shape):
    N, C, H, W = shape
    gen = torch.Generator(device="cuda").manual_seed(N * 1000 + C * 100 + H)
    x = (3.0 * torch.randn(shape, device="cuda", generator=gen)).requires_grad_(True)
    t = torch.randint(0, C, (N, H, W), device="cuda", generator=gen)
    scale = 1.0 / (N * H * W)
    for rep in range(3):                                    # the 16-byte workspace must come back zeroed every time
        got = pkg.ops.cross_entropy_2d(x, t, scale)
        want = F.nll_loss(F.log_softmax(x.detach().double(), dim=1), t, reduction='sum') * scale
        assert got.shape == () and got.dtype == torch.float32
        assert abs(float(got) - float(want)) <= 2e-6 * abs(float(want)), (rep, float(got), float(want))
    (g,) = torch.autograd.grad(got * 1.7, [x])
    xr = x.detach().clone().requires_grad_(True)
    ref = F.nll_loss(F.log_softmax(xr, dim=1), t, reduction='sum') * scale * 1.7
    (gr,) = torch.autograd.grad(ref, [xr])
    assert float((g - gr).abs().max()) <= 1e-6 * float(gr.abs().max()) + 1e-12


def test_ignored_labels_and_unaligned_views(pkg):
    gen = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(2, 4, 6, 10, device="cuda", generator=gen, requires_grad=True)
    t = torch.randint(0, 4, (2, 6, 10), device="cuda", generator=gen)
    t[0, 1, 2] = -100                                       # nll_loss's ignore_index
    t[1, 5, 9] = -100
    got = pkg.ops.cross_entropy_2d(x, t, 1.0)
    want = F.nll_loss(F.log_softmax(x, dim=1), t, reduction='sum')
    assert abs(float(got) - float(want)) <= 2e-6 * abs(float(want))
    (g,) = torch.autograd.grad(got, [x])
    (gr,) = torch.autograd.grad(want, [x])
    assert float((g - gr).abs().max()) <= 1e-6
    assert float(g[0, :, 1, 2].abs().max()) == 0.0


def test_solver_losses_route_through_the_fused_kernel(pkg):
    """losses.cross_entropy_2D (training CE) and model_util.cross_entropy_2D (saliency CE) == the oracle's values."""
    from oracle import model_oracle
    from cooperative_training_and_latent_space_data_augmentation_b200 import _lib, losses, model_util
    gen = torch.Generator().manual_seed(11)
    x = torch.randn(4, 4, 32, 32, generator=gen)
    t = torch.randint(0, 4, (4, 32, 32), generator=gen)
    before = _lib.LAUNCHES["count"]
    got_train = losses.basic_loss_fn(x.cuda(), t.cuda(), 'cross entropy')
    got_sal = model_util.cross_entropy_2D(x.cuda(), t.cuda())
    assert _lib.LAUNCHES["count"] - before == 2             # two fused forward launches, nothing else of ours
    want_train = model_oracle.ce_training(x, t)
    want_sal = model_oracle.ce_saliency(x, t)
    assert abs(float(got_train) - float(want_train)) <= 2e-6 * abs(float(want_train))
    assert abs(float(got_sal) - float(want_sal)) <= 2e-6 * abs(float(want_sal))
