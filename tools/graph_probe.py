"""Isolates what breaks under CUDA-graph capture/replay: each case runs in its own process.
usage: python tools/graph_probe.py            (runs all cases)      python tools/graph_probe.py CASE"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CASES = ["hardtrain"]


def case(name):
    import torch
    import cooperative_training_and_latent_space_data_augmentation_b200 as pkg
    ops = pkg.ops
    pkg.conv_blocks.set_precision("kernel")
    st = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    if name == "masking":
        z = torch.rand(8, 128, 14, 14, device="cuda"); gr = torch.randn(8, 128, 14, 14, device="cuda")
        rng = ops.NativeRNG(0)
        fn = lambda: ops.saliency_mask_apply(gr, z, ops.MODE_CHANNEL, 30, soft=True, rng=rng)[0]
    elif name == "conv":
        x = ops.nchw_to_c8(torch.randn(4, 16, 64, 64, device="cuda"))
        wp = ops.pack_conv_weight(torch.randn(16, 16, 3, 3, device="cuda"))
        fn = lambda: ops.conv2d_c8(x, wp, 16, 9)
    elif name in ("decoder_fwd", "decoder_fwd_bwd"):
        import torch.nn as nn
        dec = pkg.networks.MyDecoder(128, 4, feature_reduce=4, norm=nn.BatchNorm2d, up_type='NN').cuda().train()
        z = torch.randn(4, 128, 4, 4, device="cuda", requires_grad=True)
        if name == "decoder_fwd":
            def fn():
                with torch.no_grad():
                    return dec(z)
        else:
            def fn():
                y = dec(z)
                y.float().square().mean().backward()
                return y.detach()
    elif name in ("coop_repack", "coop_pool", "coop_two", "coop_instream", "coop_two_nopool", "coop_gc", "coop_empty", "coop_ptrs"):
        from cooperative_training_and_latent_space_data_augmentation_b200 import training, model_util, fastpath
        solver = pkg.AdvancedTripletReconSegmentationModel('FCN_16_standard', num_classes=4)
        solver.set_optimizers(capturable=True)
        cfgs = [{"loss_name": "mse", "mask_type": "channel", "max_threshold": 0.5, "random_threshold": True, "if_soft": True},
                {"loss_name": "ce", "mask_type": "spatial", "max_threshold": 0.5, "random_threshold": True, "if_soft": True}]
        tr = pkg.CooperativeTrainer(solver, 4, image_cfg=cfgs[0], seg_cfg=cfgs[1])
        img = torch.rand(4, 1, 64, 64, device="cuda"); lab = torch.randint(0, 4, (4, 64, 64), device="cuda")
        noise = 0.05 * torch.randn_like(img)

        def fn(opt=True):
            with model_util.recording_step_params(ops.StepParams("cuda")):
                return training.cooperative_step(solver, img, lab, cfgs[0], cfgs[1], noise=noise, optimize=opt)['loss']
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            for _ in range(2):
                fn()
        torch.cuda.synchronize()
        pool = torch.cuda.graph_pool_handle() if name in ("coop_pool", "coop_two") else None
        if name == "coop_repack":
            fastpath.weights_changed()
        g2 = torch.cuda.CUDAGraph()
        calls = []
        if name == "coop_ptrs":
            lib = pkg._lib.load()
            for fname in pkg._lib.KERNELS_PER_CALL:
                orig = getattr(lib, fname)
                setattr(lib, fname, (lambda o, f: (lambda *a: (calls.append((f, a)), o(*a))[1]))(orig, fname))
        if name == "coop_instream":
            with torch.cuda.stream(st):
                with torch.cuda.graph(g, pool=pool, stream=st):
                    out = fn(False)
        else:
            with torch.cuda.graph(g, pool=pool, stream=st):
                out = fn(False)
        if name in ("coop_two", "coop_two_nopool"):
            with torch.cuda.graph(g2, pool=pool, stream=st):
                solver.optimize_all_params()
        print("captured", flush=True)
        if name == "coop_ptrs":
            snap = torch.cuda.memory_snapshot()
            blocks = []
            for seg in snap:
                addr = seg["address"]
                for b in seg["blocks"]:
                    blocks.append((addr, addr + b["size"], b["state"], seg.get("segment_pool_id"), b["size"]))
                    addr += b["size"]
            import collections
            bad = collections.Counter()
            for f, a in calls:
                for i, v in enumerate(a):
                    if isinstance(v, int) and v > (1 << 32):
                        for lo, hi, state, pool, size in blocks:
                            if lo <= v < hi:
                                if state != "active_allocated" and tuple(pool) == (0, 0):
                                    bad[(f, i, state, str(pool), size)] += 1
                                break
                        else:
                            bad[(f, i, "unmapped", "", 0)] += 1
            for k, v in sorted(bad.items()):
                print("PTR", k, v, flush=True)
            print("calls", len(calls), flush=True)
        if name == "coop_gc":
            import gc
            gc.collect()
        if name == "coop_empty":
            torch.cuda.empty_cache()
        g.replay()
        torch.cuda.synchronize()
        print("replayed 1", flush=True)
        if name in ("coop_two", "coop_two_nopool"):
            g2.replay()
        torch.cuda.synchronize()
        print("replayed; finite:", bool(torch.isfinite(out.float()).all()), flush=True)
        return
    elif name in ("standard", "standard_nobwd", "hardgen", "hardtrain", "coop_noopt", "opt_only", "zero_only"):
        from cooperative_training_and_latent_space_data_augmentation_b200 import training, model_util
        solver = pkg.AdvancedTripletReconSegmentationModel('FCN_16_standard', num_classes=4)
        solver.set_optimizers(capturable=True)
        cfgs = [{"loss_name": "mse", "mask_type": "channel", "max_threshold": 0.5, "random_threshold": True, "if_soft": True},
                {"loss_name": "ce", "mask_type": "spatial", "max_threshold": 0.5, "random_threshold": True, "if_soft": True}]
        tr = pkg.CooperativeTrainer(solver, 4, image_cfg=cfgs[0], seg_cfg=cfgs[1])
        img = torch.rand(4, 1, 64, 64, device="cuda"); lab = torch.randint(0, 4, (4, 64, 64), device="cuda")
        noise = 0.05 * torch.randn_like(img)
        sp = ops.StepParams("cuda")

        def fn():
            solver.train()
            if name == "zero_only":
                solver.reset_all_optimizers()
                return img
            if name == "opt_only":
                solver.optimize_all_params()
                return img
            if name == "coop_noopt":
                with model_util.recording_step_params(ops.StepParams("cuda")):
                    return training.cooperative_step(solver, img, lab, cfgs[0], cfgs[1], noise=noise, optimize=False)['loss']
            a, b, c, d = solver.standard_training(img, lab, perturbed_image=torch.clamp(img + noise, 0, 1))
            loss = a + b + c + d
            if name == "standard":
                loss.backward()
            if name in ("hardgen", "hardtrain"):
                with model_util.recording_step_params(ops.StepParams("cuda")):
                    pi, ps = solver.hard_example_generation(img, lab, corrupted_image_DA_config=cfgs[0],
                                                            corrupted_seg_DA_config=cfgs[1])
                if name == "hardtrain":
                    e = solver.hard_example_training(pi, img, ps, lab)
                    (loss + sum(e)).backward()
                return pi.detach()
            return loss.detach()
    elif name == "trainer":
        solver = pkg.AdvancedTripletReconSegmentationModel('FCN_16_standard', num_classes=4)
        cfgs = [{"loss_name": "mse", "mask_type": "channel", "max_threshold": 0.5, "random_threshold": True, "if_soft": True},
                {"loss_name": "ce", "mask_type": "spatial", "max_threshold": 0.5, "random_threshold": True, "if_soft": True}]
        tr = pkg.GraphedCooperativeTrainer(solver, 4, image_cfg=cfgs[0], seg_cfg=cfgs[1], eager_steps=2)
        img = torch.rand(4, 1, 64, 64, device="cuda"); lab = torch.randint(0, 4, (4, 64, 64), device="cuda")
        for i in range(5):
            out = tr.step(img, lab)
            torch.cuda.synchronize()
            print("step", i, float(out['loss']), flush=True)
        return
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        for _ in range(2):
            want = fn()
    torch.cuda.synchronize()
    torch.cuda.memory._record_memory_history(max_entries=400000, stacks='python')
    with torch.cuda.graph(g, stream=st):
        out = fn()
    snap = torch.cuda.memory._snapshot()
    torch.cuda.memory._record_memory_history(enabled=None)
    segs = [(sg["address"], sg["address"] + sg["total_size"], tuple(sg.get("segment_pool_id", (0, 0)))) for sg in snap["segments"]]
    seen = 0
    for tr in snap["device_traces"][0]:
        if tr["action"] != "alloc":
            continue
        pool = None
        for lo, hi, pid in segs:
            if lo <= tr["addr"] < hi:
                pool = pid
                break
        if pool == (0, 0) or pool is None:
            seen += 1
            frames = [f for f in tr.get("frames", []) if "site-packages/torch" not in f["filename"]][:6]
            print("REGULAR-POOL ALLOC size", tr["size"], "stream", tr.get("stream"), "pool", pool, "::",
                  " < ".join("%s:%d %s" % (os.path.basename(f["filename"]), f["line"], f["name"]) for f in frames), flush=True)
    print("regular-pool allocations during capture:", seen, "of", sum(1 for t in snap["device_traces"][0] if t["action"] == "alloc"), flush=True)
    print("captured", flush=True)
    torch.cuda.empty_cache()
    g.replay()
    torch.cuda.synchronize()
    print("replayed; finite:", bool(torch.isfinite(out.float()).all()), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        case(sys.argv[1])
    else:
        for pdl in ("1",):
            for c in CASES:
                env = dict(os.environ, CTL_PDL=pdl)
                r = subprocess.run([sys.executable, os.path.abspath(__file__), c], env=env, stdout=subprocess.PIPE,
                                   stderr=subprocess.STDOUT, text=True, timeout=300)
                tail = [l for l in r.stdout.strip().splitlines() if l.strip()][-40:]
                print("== PDL=%s %s rc=%d :: %s" % (pdl, c, r.returncode, "\n   ".join(t[:200] for t in tail)), flush=True)
