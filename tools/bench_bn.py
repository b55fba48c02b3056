"""Microbench of the BatchNorm(train)+activation backward pair (ctl_bn_bwd_reduce_c8 + ctl_bn_bwd_apply_c8) and the
forward apply on the layer classes of the step at batch 64: CUDA events, L2 flushed between launches, GB/s over the
algorithmic bytes.  usage: python tools/bench_bn.py [B]"""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cooperative_training_and_latent_space_data_augmentation_b200 as pkg  # noqa: E402
from cooperative_training_and_latent_space_data_augmentation_b200 import _lib  # noqa: E402

ops = pkg.ops


def timed(fn, flush, iters=8):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    lib = _lib.load()
    for C, size in ((16, 224), (32, 112), (64, 56), (128, 28), (128, 14)):
        a = ops.nchw_to_c8(torch.randn(B, C, size, size, device="cuda"))
        dy = ops.nchw_to_c8(torch.randn(B, C, size, size, device="cuda") * 0.1)
        gamma = torch.ones(C, device="cuda"); beta = torch.zeros(C, device="cuda")
        scale, shift, mean, var = ops.bn_batch_affine_c8(a, gamma, beta, 1e-5, want_stats=True)
        h = ops.scale_shift_act_c8(a, scale, shift, ops.ACT_LRELU)
        nbytes = a.numel() * 2
        ws = torch.empty(lib.ctl_reduce_workspace_bytes(B, C), device="cuda", dtype=torch.uint8)
        coef = torch.empty((3, C), device="cuda"); pg = torch.empty((2, C), device="cuda"); da = torch.empty_like(a)
        st = torch.cuda.current_stream().cuda_stream

        def reduce(hp, sc, sh):
            _lib.check(lib.ctl_bn_bwd_reduce_c8(dy.data_ptr(), hp, a.data_ptr(), B, C, size, size, ops.ACT_LRELU,
                                                mean.data_ptr(), var.data_ptr(), 1e-5, gamma.data_ptr(), ws.data_ptr(), 0,
                                                coef.data_ptr(), pg[0].data_ptr(), pg[1].data_ptr(), sc, sh, st))

        def apply(hp, sc, sh):
            _lib.check(lib.ctl_bn_bwd_apply_c8(dy.data_ptr(), hp, a.data_ptr(), B, C, size, size, ops.ACT_LRELU,
                                               coef.data_ptr(), da.data_ptr(), sc, sh, st))
        rows = [("reduce  (h read)", lambda: reduce(h.data_ptr(), 0, 0), 3),
                ("reduce  (sign from a)", lambda: reduce(0, scale.data_ptr(), shift.data_ptr()), 2),
                ("apply   (h read)", lambda: apply(h.data_ptr(), 0, 0), 4),
                ("apply   (sign from a)", lambda: apply(0, scale.data_ptr(), shift.data_ptr()), 3),
                ("forward scale_shift_act", lambda: ops.scale_shift_act_c8(a, scale, shift, ops.ACT_LRELU), 2)]
        for name, fn, passes in rows:
            t = timed(fn, flush)
            print("C%-3d @%-3d %-26s %7.1f us  %6.0f GB/s (%d tensor passes of %.1f MB)"
                  % (C, size, name, t, passes * nbytes / t / 1e3, passes, nbytes / 1e6), flush=True)


if __name__ == "__main__":
    main()
