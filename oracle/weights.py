"""TEST INFRASTRUCTURE ONLY -- deterministic synthetic weights.

There is no network for checkpoints and a 10 MB state_dict is too big to commit, so golden
model fixtures are generated from weights that any box can regenerate bit-for-bit: numpy's
legacy RandomState (stable across numpy versions) keyed by (seed, crc32(tensor name)).
Shapes and key names come from the module itself (reference module in make_golden.py,
oracle/product modules in the tests), so the same call fills all three identically.
"""
import zlib

import numpy as np
import torch


def synthetic_tensor(name, shape, seed):
    rs = np.random.RandomState((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFF)
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "num_batches_tracked":
        return torch.zeros((), dtype=torch.long)
    if leaf == "running_mean":
        return torch.from_numpy((0.1 * rs.standard_normal(shape)).astype(np.float32))
    if leaf == "running_var":
        return torch.from_numpy((1.0 + 0.2 * rs.random_sample(shape)).astype(np.float32))
    if len(shape) == 4:                                   # conv / conv-transpose weight
        fan_in = shape[1] * shape[2] * shape[3]
        return torch.from_numpy((np.sqrt(2.0 / fan_in) * rs.standard_normal(shape)).astype(np.float32))
    if leaf == "weight":                                  # BN gamma
        return torch.from_numpy((1.0 + 0.1 * rs.standard_normal(shape)).astype(np.float32))
    return torch.from_numpy((0.05 * rs.standard_normal(shape)).astype(np.float32))   # biases, BN beta


def synthetic_state_dict(module, seed, prefix=""):
    return {k: synthetic_tensor(prefix + k, tuple(v.shape), seed) for k, v in module.state_dict().items()}


def synthetic_batch(N, H, W, seed, num_classes=4):
    """ACDC-shaped synthetic data: image U[0,1) f32 [N,1,H,W], label {0..3} i64."""
    rs = np.random.RandomState(seed)
    img = torch.from_numpy(rs.random_sample((N, 1, H, W)).astype(np.float32))
    lab = torch.from_numpy(rs.randint(0, num_classes, size=(N, H, W)).astype(np.int64))
    noise = torch.from_numpy((0.05 * rs.standard_normal((N, 1, H, W))).astype(np.float32))
    return img, lab, noise
