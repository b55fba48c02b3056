"""One cooperative training step (the reference's loop body,
medseg/train_adv_supervised_segmentation_triplet.py:171-237) and its batch-sharded data-parallel
form over NCCL (the reference has no distributed code).

Step = clean pass (standard_training) -> hard-example generation (latent masking, K1/K2) ->
corrupted-image and corrupted-shape passes (hard_example_training) -> backward -> gradient
all-reduce (the flat fp32 gradient buffer of optim.FlatAdam, 2.53 M elements) -> ONE multi-tensor Adam launch that also
takes the 1/world average.

Differences from the reference loop that do not change results: losses stay on the device (the
reference calls .item() nine times per step), no gc.collect()/empty_cache().

Data parallel contract (8e):
  * the batch is split in contiguous per-rank slices; no collective inside the masking kernels
  * host RNG draws (mask type via python `random`, percentile via numpy) must be identical on all
    ranks -> `seed_host_rng` seeds both the same way everywhere
  * device RNG uses the native Philox mode keyed by the GLOBAL sample index, so a shard draws what
    the full batch would
  * BatchNorm uses per-rank batch statistics (equals the reference at the per-rank batch size;
    no SyncBN); running statistics are therefore per-rank and rank 0's are the ones checkpointed
  * gradients are averaged over ranks (each rank's losses are means over its own slice)
"""
import random as _pyrandom

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, model_util, ops, trainpath

LOSS_KEYS = ('loss/standard/total', 'loss/standard/seg', 'loss/standard/image', 'loss/standard/shape',
             'loss/standard/gt_shape', 'loss/hard/total', 'loss/hard/seg', 'loss/hard/image', 'loss/hard/shape')

DEFAULT_IMAGE_CFG = {"loss_name": "mse", "mask_type": "random", "max_threshold": 0.5, "random_threshold": True,
                     "if_soft": True}
DEFAULT_SEG_CFG = {"loss_name": "ce", "mask_type": "random", "max_threshold": 0.5, "random_threshold": True,
                   "if_soft": True}


def seed_host_rng(seed):
    """python `random` + numpy global generator: the two host streams the hot path consumes."""
    _pyrandom.seed(seed)
    np.random.seed(seed)


def latent_da_configs(experiment_opt):
    """Reads the `latent_DA` block of config/ACDC/cooperative_training.json verbatim
    (train...triplet.py:125-142).  Returns (gen_image, image_cfg, gen_seg, seg_cfg)."""
    if not experiment_opt.get('learning', {}).get('latent_DA', False):
        return False, None, False, None
    block = experiment_opt['latent_DA']
    scope = block['mask_scope']
    gi, gs = 'image code' in scope, 'shape code' in scope
    return gi, block['image code'] if gi else None, gs, block['shape code'] if gs else None


def cooperative_step(solver, clean_image_l, label_l, corrupted_image_DA_config=None, corrupted_seg_DA_config=None,
                     gen_corrupted_image=True, gen_corrupted_seg=True, latent_DA=True, separate_training=False,
                     noise=None, grad_sync=None, optimize=True, hard_examples=None):
    """Runs the loop body once.  Returns a dict of 0-d DEVICE tensors keyed like the reference's loss_dict
    plus 'loss' (nothing is synchronised; call .item() on what you want to log).
    hard_examples: optional (perturbed_image, perturbed_seg) to train on instead of generating them (replaying a
    recorded step; the parity tests use it to compare two numerics modes on IDENTICAL hard examples)."""
    icfg = corrupted_image_DA_config or DEFAULT_IMAGE_CFG
    scfg = corrupted_seg_DA_config or DEFAULT_SEG_CFG
    solver.train()
    solver.reset_all_optimizers()
    if noise is None:
        noise = 0.05 * torch.randn_like(clean_image_l)
    image_l = torch.clamp(clean_image_l + noise, 0, 1)

    seg_loss, image_recon_loss, gt_recon_loss, shape_recon_loss = solver.standard_training(
        clean_image_l, label_l, perturbed_image=image_l, separate_training=separate_training)
    standard_loss = seg_loss + image_recon_loss + shape_recon_loss + gt_recon_loss
    out = {'loss/standard/total': standard_loss.detach(), 'loss/standard/seg': seg_loss.detach(),
           'loss/standard/image': image_recon_loss.detach(), 'loss/standard/shape': shape_recon_loss.detach(),
           'loss/standard/gt_shape': gt_recon_loss.detach()}

    if latent_DA:
        if hard_examples is not None:
            p_img, p_seg = hard_examples
        else:
            p_img, p_seg = solver.hard_example_generation(
                clean_image_l.detach(), label_l.detach(), gen_corrupted_seg=gen_corrupted_seg,
                gen_corrupted_image=gen_corrupted_image, corrupted_image_DA_config=icfg, corrupted_seg_DA_config=scfg)
        h_seg, h_img, h_shape2, h_cshape = solver.hard_example_training(
            perturbed_image=p_img, perturbed_seg=p_seg, clean_image_l=clean_image_l, label_l=label_l,
            separate_training=separate_training)
        hard_loss = h_seg + h_img + h_shape2 + h_cshape
        out.update({'loss/hard/total': hard_loss.detach(), 'loss/hard/seg': h_seg.detach(),
                    'loss/hard/image': h_img.detach(), 'loss/hard/shape': (h_shape2 + h_cshape).detach(),
                    'perturbed_image': p_img, 'perturbed_seg': p_seg})
    else:
        hard_loss = torch.zeros((), device=clean_image_l.device)
        out.update({'loss/hard/total': hard_loss, 'loss/hard/seg': hard_loss, 'loss/hard/image': hard_loss,
                    'loss/hard/shape': hard_loss})

    loss = standard_loss + hard_loss
    solver.reset_all_optimizers()
    # parameters whose .grad already exists (the trainers' flat bucket, or zero_grad(set_to_none=False)) receive their
    # gradients straight from the backward kernels instead of one autograd add per parameter and pass
    with trainpath.accumulate_into_grads():
        loss.backward()
    if grad_sync is not None:
        grad_sync()
    if optimize:
        solver.optimize_all_params()
    out['loss'] = loss.detach()
    return out


def shard_bounds(global_batch, world_size, rank):
    """Contiguous N/world slices (8e).  The global batch must divide evenly."""
    if global_batch % world_size:
        raise ValueError("global batch %d is not divisible by world size %d" % (global_batch, world_size))
    per = global_batch // world_size
    return rank * per, (rank + 1) * per


def broadcast_module_state(modules, src=0, group=None):
    """Parameters and buffers of rank `src` to everyone (identical initial weights)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for m in modules:
        for t in list(m.parameters()) + list(m.buffers()):
            dist.broadcast(t.data, src=src, group=group)
    # the write went through .data: neither Tensor._version nor the optimizer hook saw it, and ranks != src may hold
    # packed bf16 copies of their pre-broadcast weights (a forward that ran before the trainer was built)
    from . import fastpath
    fastpath.weights_changed()


class _GradExchange:
    """The data-parallel exchange of one step: ONE all-reduce (sum) of the solver's flat gradient buffer; the 1/world
    average is applied inside the Adam kernel (FlatAdam.grad_scale), so no separate division pass runs."""

    def __init__(self, flat_adam, group, world):
        self.flat_adam, self.group, self.world = flat_adam, group, world
        self.flat = flat_adam.flat_grads
        flat_adam.grad_scale = 1.0 / world

    def attached(self):
        return self.flat_adam.attached()

    def reattach(self):
        self.flat_adam.reattach()

    def all_reduce_sum(self):
        if self.world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)

    def mean_gradients(self):
        """The averaged gradient as the optimizer sees it (a copy; the buffer itself holds the sum)."""
        return self.flat * (1.0 / self.world)


class CooperativeTrainer:
    """Batch-sharded cooperative training: one process per GPU, one NCCL all-reduce of the flat gradient buffer."""

    def __init__(self, solver, global_batch, seed=0, image_cfg=None, seg_cfg=None, group=None):
        self.solver = solver
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.global_batch = global_batch
        self.lo, self.hi = shard_bounds(global_batch, self.world, self.rank)
        self.image_cfg = image_cfg or DEFAULT_IMAGE_CFG
        self.seg_cfg = seg_cfg or DEFAULT_SEG_CFG
        self.seed = seed
        self.step_index = 0
        if solver.optimizers is None:
            solver.set_optimizers()
        if not solver.flat_adam.attached():
            solver.flat_adam.reattach()
        broadcast_module_state(solver.model.values(), 0, group)
        self.bucket = _GradExchange(solver.flat_adam, group, self.world)
        seed_host_rng(seed)                         # identical host draws on every rank
        model_util.set_rng_mode("philox", seed=seed, first_sample=self.lo)

    def local_slice(self, global_tensor):
        return global_tensor[self.lo:self.hi]

    def params_in_sync(self):
        """True when every rank holds bit-identical parameters (checksum MIN == MAX over ranks of the flat parameter
        buffer's sum and absolute sum, in fp64)."""
        flat = self.solver.flat_adam.flat_params.double()
        probe = torch.stack([flat.sum(), flat.abs().sum(), (flat * flat).sum()])
        if self.world == 1:
            return True
        lo, hi = probe.clone(), probe.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.group)
        return bool(torch.equal(lo, hi))

    def step(self, clean_local, label_local, noise_local=None):
        """`clean_local` / `label_local` are this rank's slice of the global batch."""
        # Philox stream: advance first_sample by the global batch per step so no (sample, step) pair repeats
        model_util.native_rng().first_sample = self.step_index * self.global_batch + self.lo
        out = cooperative_step(self.solver, clean_local, label_local, self.image_cfg, self.seg_cfg,
                               noise=noise_local, grad_sync=self.bucket.all_reduce_sum)
        self.step_index += 1
        return out


def draw_host_params(cfg):
    """The host draws one `perturb_latent_code` call makes, in the reference's order: python `random.shuffle` for a
    'random' mask type (advanced...model.py:325-328), then numpy's global generator for a random percentile
    (model_util.py:224-225, :285-286; dropout uses the threshold as is).  Returns (perturb_type, percentile)."""
    kind = cfg["mask_type"]
    if kind == 'random':
        candidates = ['dropout', 'spatial', 'channel']
        _pyrandom.shuffle(candidates)
        kind = candidates[0]
    p = cfg["max_threshold"]
    if kind != 'dropout' and cfg["random_threshold"]:
        p = np.random.rand() * p
    return kind, p


class _CapturedStep:
    """The CUDA graph of one (image mask type, shape mask type) combination and what its replay needs."""

    def __init__(self, device):
        self.graph = torch.cuda.CUDAGraph()
        self.optimizers = None      # second graph, only when the all-reduce is NOT captured
        self.params = ops.StepParams(device)
        self.out = None
        self.kernels = 0            # kernels of libctl_b200.so recorded in the graph(s)


class GraphedCooperativeTrainer(CooperativeTrainer):
    """CooperativeTrainer whose step is replayed from ONE CUDA graph (the step is >1000 short launches: issued one by
    one from Python it is host-bound on a B200).

    * the graph = zero grads + clean pass + hard-example generation + corrupted passes + backward + the NCCL all-reduce
      of the flat gradient buffer + the multi-tensor Adam launch (1/world folded in): nothing of a step is issued from
      the host but the inputs' copy, one 192-byte parameter upload and the graph launch.
      `capture_collective=False` (or env CTL_NO_CAPTURED_COLLECTIVE=1) keeps the all-reduce an ordinary call between a
      forward/backward graph and an optimizer graph
    * inputs are copied into static buffers (so `step` takes host-pinned or device tensors alike); the returned losses
      and perturbed examples are static tensors that the NEXT step overwrites
    * the per-step host draws (mask type, percentile -> k) are made here exactly as the eager path makes them
      (`draw_host_params`), uploaded as device-resident step parameters (ops.StepParams) and read by the *_dyn
      kernels; one graph is captured lazily per (image mask type, shape mask type) combination, all in one memory pool
    * the first `eager_steps` calls run eagerly on the capture stream (library / allocator / communicator warm-up)
    * ONE batch shape per trainer; BatchNorm running statistics are per rank (rank 0's are the ones checkpointed)
    * optimizer state (moments, step counts) lives in solver.flat_adam and is never recreated here: build the trainer
      after `solver.load_snapshots(...)` to resume, or call load_snapshots later -- it copies in place, so captured
      graphs stay valid
    """

    def __init__(self, solver, global_batch, seed=0, image_cfg=None, seg_cfg=None, group=None, eager_steps=3,
                 capture_collective=True):
        super().__init__(solver, global_batch, seed, image_cfg, seg_cfg, group)
        import os
        if os.environ.get("CTL_NO_CAPTURED_COLLECTIVE", "0") == "1":
            capture_collective = False
        self.capture_collective = bool(capture_collective) or self.world == 1
        self.eager_steps = eager_steps
        self.stream = torch.cuda.Stream()
        self.copy_stream = torch.cuda.Stream()      # input prefetch (prefetch()): H2D under the previous step's compute
        self._stage, self._prefetched = None, None
        self._stage_consumed = torch.cuda.Event()
        self._stage_consumed.record()
        self.pool = torch.cuda.graph_pool_handle()
        self.captured = {}
        self.static = None

    # -------------------------------------------------------------------------------------------- helpers
    def _stage_inputs(self, clean, label, noise):
        if self.static is None:
            dev = next(self.solver.parameters()).device
            self.static = {"clean": torch.empty(clean.shape, device=dev, dtype=torch.float32),
                           "label": torch.empty(label.shape, device=dev, dtype=torch.int64),
                           "noise": None}
        st = self.static
        if tuple(clean.shape) != tuple(st["clean"].shape):
            raise ValueError("a graphed trainer replays ONE batch shape: got %s, captured %s"
                             % (tuple(clean.shape), tuple(st["clean"].shape)))
        pf = self._prefetched
        if pf is not None and pf["clean"] is clean and pf["label"] is label:
            # the host -> device copy of this batch already ran on the copy stream, under the previous step's compute
            torch.cuda.current_stream().wait_event(pf["ready"])
            st["clean"].copy_(pf["dev_clean"], non_blocking=True)
            st["label"].copy_(pf["dev_label"], non_blocking=True)
            self._stage_consumed.record()
            self._prefetched = None
        else:
            st["clean"].copy_(clean, non_blocking=True)
            st["label"].copy_(label, non_blocking=True)
        if noise is not None:
            if st["noise"] is None:
                st["noise"] = torch.empty_like(st["clean"])
            st["noise"].copy_(noise, non_blocking=True)
        return st["clean"], st["label"], (st["noise"] if noise is not None else None)

    def prefetch(self, clean, label):
        """Starts the host -> device copy of the NEXT step's batch (pinned host tensors) on a side stream, so that it
        overlaps the step that is currently running; `step(clean, label)` with the SAME tensor objects then only does a
        device-to-device copy into the graph's static inputs.  Input pipelining: call it right after `step` returns and
        before reading that step's results.  The host tensors must not be modified until the next `step` has returned."""
        if self.static is None or not (clean.device.type == "cpu" and clean.is_pinned() and label.is_pinned()):
            return                                      # first step not staged yet / nothing to overlap
        if self._stage is None:
            self._stage = {"clean": torch.empty_like(self.static["clean"]), "label": torch.empty_like(self.static["label"])}
        if tuple(clean.shape) != tuple(self._stage["clean"].shape) or tuple(label.shape) != tuple(self._stage["label"].shape):
            return
        self.copy_stream.wait_event(self._stage_consumed)   # the previous batch has left the staging buffers
        with torch.cuda.stream(self.copy_stream):
            self._stage["clean"].copy_(clean, non_blocking=True)
            self._stage["label"].copy_(label, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record()
        self._prefetched = {"clean": clean, "label": label, "ready": ready, "dev_clean": self._stage["clean"],
                            "dev_label": self._stage["label"]}

    def _eager(self, clean, label, noise):
        return cooperative_step(self.solver, clean, label, self.image_cfg, self.seg_cfg, noise=noise,
                                grad_sync=self.bucket.all_reduce_sum)

    def _capture(self, key, clean, label, noise):
        from . import fastpath
        cs = _CapturedStep(clean.device)
        if not self.bucket.attached():
            self.bucket.reattach()                  # every parameter / .grad is a view of the flat buffers before recording
        fastpath.prepare_packing()                  # job table of the batched weight packing (host -> device copy)
        fastpath.weights_changed()                  # every packed weight is rebuilt INSIDE the graph
        n0 = _lib.LAUNCHES["count"]
        with model_util.recording_step_params(cs.params):
            with torch.cuda.graph(cs.graph, pool=self.pool, stream=self.stream):
                cs.out = cooperative_step(self.solver, clean, label, self.image_cfg, self.seg_cfg, noise=noise,
                                          grad_sync=self.bucket.all_reduce_sum if self.capture_collective else None,
                                          optimize=self.capture_collective)
        if not self.capture_collective:
            cs.optimizers = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cs.optimizers, pool=self.pool, stream=self.stream):
                self.solver.optimize_all_params()
        if not self.bucket.attached():
            raise RuntimeError("a parameter or gradient left the flat buffers during capture: the recorded step would "
                               "update memory the optimizer does not read")
        cs.kernels = _lib.LAUNCHES["count"] - n0
        _lib.LAUNCHES["count"] = n0                 # nothing ran yet: replays add the count
        self.captured[key] = cs
        return cs

    def close(self):
        """Destroys the captured graphs.  Call before torch.distributed.destroy_process_group(): a communicator whose
        collectives are still referenced by live CUDA graphs cannot be torn down (the destroy blocks)."""
        torch.cuda.synchronize()
        for cs in self.captured.values():
            cs.graph.reset()
            if cs.optimizers is not None:
                cs.optimizers.reset()
        self.captured.clear()
        torch.cuda.synchronize()

    # -------------------------------------------------------------------------------------------- the step
    def step(self, clean_local, label_local, noise_local=None):
        from . import fastpath
        rng = model_util.native_rng()
        rng.first_sample = self.step_index * self.global_batch + self.lo
        cur = torch.cuda.current_stream()
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            clean, label, noise = self._stage_inputs(clean_local, label_local, noise_local)
            if self.step_index < self.eager_steps:
                out = self._eager(clean, label, noise)
            else:
                host_state = (_pyrandom.getstate(), np.random.get_state())
                draws = [draw_host_params(self.image_cfg), draw_host_params(self.seg_cfg)]
                key = (draws[0][0], draws[1][0], noise is not None)
                cs = self.captured.get(key)
                if cs is None:
                    # the capture pass runs the ordinary host code, which makes the same draws again
                    _pyrandom.setstate(host_state[0])
                    np.random.set_state(host_state[1])
                    offset0 = rng.offset
                    cs = self._capture(key, clean, label, noise)
                    rng.offset = offset0
                cs.params.begin()
                rows = iter(enumerate(cs.params.rows))
                for kind, p in draws:
                    i, row = next(rows)
                    if row["kind"] != ("dropout" if kind == "dropout" else "mask"):
                        raise RuntimeError("captured step does not match this step's draws")
                    k = int(row["n"] * p) if kind != "dropout" else 0
                    if k >= row["n"] or k < -row["n"]:
                        raise IndexError("index {} is out of bounds for dimension 1 with size {}".format(k, row["n"]))
                    cs.params.fill(i, k % row["n"] if kind != "dropout" else 0, rng if row["draws"] else None)
                cs.params.upload()
                cs.graph.replay()
                if cs.optimizers is not None:
                    self.bucket.all_reduce_sum()
                    cs.optimizers.replay()
                _lib.LAUNCHES["count"] += cs.kernels
                fastpath.weights_changed()          # the graph stepped the weights behind autograd's back
                out = cs.out
        cur.wait_stream(self.stream)
        self.step_index += 1
        return out
