"""CPU: host-side logic of the product package that needs no device."""
import torch


def test_packed_weight_cache_follows_every_kind_of_update():
    """Fused optimizers update parameters WITHOUT bumping Tensor._version; the packed bf16 copies must still be rebuilt
    (a stale forward weight silently stops training).  Also: in-place edits, storage replacement, per-object cache."""
    from cooperative_training_and_latent_space_data_augmentation_b200 import fastpath
    p = torch.nn.Parameter(torch.randn(4, 4))
    calls = []

    def pack(w):
        calls.append(1)
        return w.detach().clone()

    a = fastpath._packed(p, pack)
    assert fastpath._packed(p, pack) is a and len(calls) == 1
    assert fastpath._packed(p, pack, tag='dgrad') is not a and len(calls) == 2          # tags are separate entries
    opt = torch.optim.Adam([p], lr=1e-2, fused=True)
    p.grad = torch.ones_like(p)
    opt.step()
    b = fastpath._packed(p, pack)
    assert len(calls) == 3 and not torch.equal(a, b) and torch.equal(b, p.detach())
    with torch.no_grad():
        p.mul_(2.0)
    assert torch.equal(fastpath._packed(p, pack), p.detach()) and len(calls) == 4
    p.data = torch.zeros(4, 4)
    assert torch.equal(fastpath._packed(p, pack), torch.zeros(4, 4)) and len(calls) == 5
    q = torch.nn.Parameter(torch.ones(4, 4))                                               # a different object never hits
    assert torch.equal(fastpath._packed(q, pack), torch.ones(4, 4)) and len(calls) == 6
    fastpath.weights_changed()
    fastpath._packed(q, pack)
    assert len(calls) == 7
