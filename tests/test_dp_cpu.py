"""CPU, world size 2 over gloo: the host side of the batch-sharded step -- shard bounds, the flat gradient
buffer of optim.FlatAdam and its all-reduce (the 1/world average is folded into the Adam kernel), parameter broadcast, identical host RNG draws on every rank, and the
shard-invariance of the native Philox stream (restated by the oracle)."""
import os
import random
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cooperative_training_and_latent_space_data_augmentation_b200 import optim, training
        torch.manual_seed(100 + rank)                       # different initial weights per rank on purpose
        net = nn.Sequential(nn.Conv2d(2, 3, 3), nn.BatchNorm2d(3), nn.Conv2d(3, 1, 1))
        net(torch.ones(1, 2, 5, 5))                         # a forward BEFORE the trainer exists (stale packed weights case)
        flat = optim.FlatAdam({"net": net}, lr=1e-3)        # parameters / gradients become views of flat buffers
        training.broadcast_module_state([net], src=0)
        w_after_bcast = torch.cat([p.detach().reshape(-1) for p in net.parameters()]).clone()
        assert flat.attached()                              # the broadcast wrote THROUGH the views
        exchange = training._GradExchange(flat, None, world)
        assert flat.grad_scale == 1.0 / world               # the average is taken inside the Adam kernel
        # a rank-dependent "backward": loss = (rank+1) * sum(net(x))
        x = torch.ones(2, 2, 5, 5)
        ((rank + 1.0) * net(x).sum()).backward()
        local = exchange.flat.clone()
        net[0].weight.grad = None                           # something dropped a grad: reattach must restore the view
        assert not exchange.attached()
        exchange.reattach()
        exchange.all_reduce_sum()
        assert net[0].weight.grad is not None and net[0].weight.grad.data_ptr() == exchange.flat.data_ptr()
        reduced = exchange.mean_gradients()
        lo, hi = training.shard_bounds(8, world, rank)
        training.seed_host_rng(7)
        draws = (random.random(), float(np.random.rand()))
        out[rank] = {"w": w_after_bcast, "local": local, "reduced": reduced.clone(), "bounds": (lo, hi),
                     "draws": draws}
    finally:
        dist.destroy_process_group()


def test_flat_bucket_allreduce_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    assert torch.equal(r0["w"], r1["w"])                                   # broadcast made the replicas identical
    assert r0["bounds"] == (0, 4) and r1["bounds"] == (4, 8)
    assert r0["draws"] == r1["draws"]                                      # same host RNG stream on every rank
    assert torch.equal(r0["reduced"], r1["reduced"])
    # rank 0 dropped conv0.weight.grad before the exchange (reattach zero-fills it); everything else is the mean
    n0 = 3 * 2 * 3 * 3
    want = 0.5 * (r0["local"] + r1["local"])
    torch.testing.assert_close(r0["reduced"][n0:], want[n0:], rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(r0["reduced"][:n0], torch.zeros(n0), rtol=0, atol=0)
    # gradient of rank r is (r+1) x the base gradient -> mean = 1.5 x base
    torch.testing.assert_close(r1["local"][n0:], 2.0 * r0["local"][n0:], rtol=1e-5, atol=1e-7)


def test_shard_bounds_reject_uneven_split():
    from cooperative_training_and_latent_space_data_augmentation_b200 import training
    assert training.shard_bounds(512, 8, 3) == (192, 256)
    with pytest.raises(ValueError):
        training.shard_bounds(10, 4, 0)


def test_native_rng_is_shard_invariant():
    """A rank holding samples [lo, hi) draws exactly rows lo..hi of what the full batch would draw."""
    from oracle import masking_oracle as mo
    full = mo.native_rand(5, 3, 8, 196, first_sample=0)
    for lo, hi in ((0, 4), (4, 8)):
        part = mo.native_rand(5, 3, hi - lo, 196, first_sample=lo)
        assert np.array_equal(part, full[lo:hi])
    keep_full = mo.native_keep(5, 1, 8, 128, 0.5, first_sample=0)
    assert np.array_equal(mo.native_keep(5, 1, 4, 128, 0.5, first_sample=4), keep_full[4:8])


def test_latent_da_config_block_is_read_verbatim():
    from cooperative_training_and_latent_space_data_augmentation_b200 import training
    opt = {"learning": {"latent_DA": True},
           "latent_DA": {"mask_scope": ["image code", "shape code"],
                         "image code": {"loss_name": "mse", "mask_type": "random", "max_threshold": 0.5,
                                        "random_threshold": True, "if_soft": True},
                         "shape code": {"loss_name": "ce", "mask_type": "random", "max_threshold": 0.5,
                                        "random_threshold": True, "if_soft": True}}}
    gi, ic, gs, sc = training.latent_da_configs(opt)
    assert gi and gs and ic["loss_name"] == "mse" and sc["loss_name"] == "ce"
    assert training.latent_da_configs({"learning": {"latent_DA": False}}) == (False, None, False, None)
