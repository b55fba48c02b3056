"""Times the training-path masking tail at the model's real latent shapes ([64,128,14,14] @224^2, [64,128,16,16] @256^2):
ctl_saliency_sums_mask_apply (one kernel: s from the epilogue's sums, select, mask, apply, NCHW + C8 outputs) next to the
materialised chain it replaces (c8_to_nchw + K1 + select + K2 + nchw_to_c8).  CUDA events around a graph of 40 back-to-back calls."""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cooperative_training_and_latent_space_data_augmentation_b200 as pkg  # noqa: E402

ops = pkg.ops


def timed(fn, iters=20, reps=40):
    """Per-call device time with the launches replayed from a CUDA graph (the host cost of the Python wrappers --
    several allocations and a ctypes call, more than the kernels take -- is what the captured training step does not pay)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            for _ in range(reps):
                fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); graph.replay(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / reps)
    return statistics.median(ts)


for shape in ((64, 128, 14, 14), (64, 128, 16, 16)):
    N, C, H, W = shape
    gen = torch.Generator(device="cuda").manual_seed(0)
    z = torch.relu(torch.randn(*shape, device="cuda", generator=gen))
    g_c8 = ops.nchw_to_c8(1e-5 * torch.randn(*shape, device="cuda", generator=gen))
    for mode, n in ((ops.MODE_CHANNEL, C), (ops.MODE_SPATIAL, H * W)):
        sums = torch.randn(N, n, device="cuda", dtype=torch.float64, generator=gen)
        rng = ops.NativeRNG(0)
        k = int(0.3 * n)
        fused = timed(lambda: ops.saliency_sums_mask_apply(sums, z, mode, k, soft=True, rng=rng))

        def chain():
            g = ops.c8_to_nchw(g_c8)
            zt, _, _, _ = ops.saliency_mask_apply(g, z, mode, k, soft=True, rng=rng)
            return ops.nchw_to_c8(zt)
        print("%s mode %d: fused tail %.1f us, materialised chain %.1f us" % (shape, mode, fused, timed(chain)))
