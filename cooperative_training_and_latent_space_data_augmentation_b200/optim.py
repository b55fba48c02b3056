"""The five Adam optimizers of the reference (medseg/models/advanced_triplet_recon_segmentation_model.py:774-785:
one `optim.Adam(model.parameters(), lr)` per sub-network) as ONE multi-tensor kernel over flat buffers
(csrc/optim_metrics.cu `ctl_adam_flat`).

`FlatAdam` re-points every parameter of the given sub-networks into one flat fp32 buffer (module tree, names and
`state_dict()` untouched), keeps gradients and both Adam moments in buffers of the same layout and steps everything --
or any subset of the sub-networks -- with one launch.  The same flat gradient buffer is what the data-parallel
all-reduce exchanges (training.py), its 1/world average is folded into the Adam pass (`grad_scale`), and so is the
clearing of the gradients for the next step.

`FlatAdam.view(name)` is what `solver.optimizers[name]` holds: `.step()`, `.zero_grad()`, `.state_dict()` /
`.load_state_dict()` in torch.optim.Adam's own format (so reference checkpoints written with save_optimizers=True load,
advanced...model.py:676-677, :731-732) and a `param_groups` list whose 'lr' is honoured.

Semantics kept from torch.optim.Adam(lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False).  One
difference that cannot change results: gradients are never None (they are views of the flat buffer, zero when unused);
a parameter whose gradient is identically zero receives a zero update (exp_avg = exp_avg_sq = 0 -> 0 / (0 + eps)).
"""
import torch

from . import ops

_ALIGN_SEGMENT = 64       # elements: every sub-network starts on a 256-byte boundary of the flat buffers


class FlatAdam:
    def __init__(self, named_modules, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.names = list(named_modules)
        if not 1 <= len(self.names) <= 8:
            raise ValueError("FlatAdam takes 1..8 sub-networks")
        self.modules = dict(named_modules)
        self.params, self.offsets, self.bounds, self.segment_of = [], [], [], []
        off = 0
        for s, name in enumerate(self.names):
            off = (off + _ALIGN_SEGMENT - 1) // _ALIGN_SEGMENT * _ALIGN_SEGMENT
            begin = off
            for p in self.modules[name].parameters():
                if p.dim() > 1:
                    off = (off + 3) // 4 * 4          # weight tensors on 16-byte boundaries (16-byte reductions of K3w)
                self.params.append(p)
                self.offsets.append(off)
                self.segment_of.append(s)
                off += p.numel()
            self.bounds.append((begin, off))
        if not self.params:
            raise ValueError("no parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        if dt != torch.float32 or any(p.dtype != dt or p.device != dev for p in self.params):
            raise ValueError("FlatAdam needs fp32 parameters on one device")
        self.numel = (off + 3) // 4 * 4
        self.flat_params = torch.zeros(self.numel, device=dev, dtype=dt)
        self.flat_grads = torch.zeros(self.numel, device=dev, dtype=dt)
        self.exp_avg = torch.zeros(self.numel, device=dev, dtype=dt)
        self.exp_avg_sq = torch.zeros(self.numel, device=dev, dtype=dt)
        self.steps = torch.zeros(len(self.names), device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                view = self.flat_params[o:o + p.numel()].view_as(p)
                view.copy_(p)
                old_grad = p.grad
                p.data = view
                g = self.flat_grads[o:o + p.numel()].view_as(p)
                if old_grad is not None:
                    g.copy_(old_grad)
                p.grad = g
        self.hyper = [{"lr": float(lr), "betas": (float(betas[0]), float(betas[1])), "eps": float(eps),
                       "weight_decay": float(weight_decay)} for _ in self.names]
        self.grad_scale = 1.0            # data-parallel trainers set 1/world: the average is taken inside the Adam pass
        self.clear_grads_in_step = False
        self._views = {name: _AdamView(self, s) for s, name in enumerate(self.names)}
        for s, name in enumerate(self.names):
            self.modules[name].__dict__['_ctl_flat_segment'] = (self, s)

    # ------------------------------------------------------------------------------------------------ bookkeeping
    def view(self, name):
        return self._views[name]

    def attached(self):
        """True while every parameter and gradient is still its view of the flat buffers (module.to(), a foreign
        optimizer or zero_grad(set_to_none=True) would break that)."""
        es = self.flat_params.element_size()
        for p, o in zip(self.params, self.offsets):
            if p.data_ptr() != self.flat_params.data_ptr() + o * es:
                return False
            if p.grad is None or p.grad.data_ptr() != self.flat_grads.data_ptr() + o * es:
                return False
        return True

    def reattach(self):
        """Folds parameters / gradients that were replaced (p.grad = None, p.data = ...) back into the flat buffers."""
        with torch.no_grad():
            es = self.flat_params.element_size()
            for p, o in zip(self.params, self.offsets):
                if p.data_ptr() != self.flat_params.data_ptr() + o * es:
                    view = self.flat_params[o:o + p.numel()].view_as(p)
                    view.copy_(p)
                    p.data = view
                g = self.flat_grads[o:o + p.numel()].view_as(p)
                if p.grad is None:
                    g.zero_()
                    p.grad = g
                elif p.grad.data_ptr() != g.data_ptr():
                    g.copy_(p.grad)
                    p.grad = g

    def segment_slice(self, tensor, s):
        b, e = self.bounds[s]
        return tensor[b:e]

    def zero_grad(self, segment=None):
        if segment is None:
            self.flat_grads.zero_()
        else:
            self.segment_slice(self.flat_grads, segment).zero_()

    # ------------------------------------------------------------------------------------------------ the step
    def step(self, segments=None):
        """One launch for all sub-networks (segments=None) or the listed ones.  Segments with different
        hyper-parameters (someone edited a view's param_groups) are stepped in groups of equal settings."""
        todo = list(range(len(self.names))) if segments is None else list(segments)
        groups = {}
        for s in todo:
            h = self.hyper[s]
            groups.setdefault((h["lr"], h["betas"], h["eps"], h["weight_decay"]), []).append(s)
        for (lr, betas, eps, wd), segs in groups.items():
            mask = 0
            for s in segs:
                mask |= 1 << s
            ops.adam_flat(self.flat_params, self.flat_grads, self.exp_avg, self.exp_avg_sq, self.bounds, self.steps,
                          lr, betas, eps, wd, grad_scale=self.grad_scale, zero_grad=self.clear_grads_in_step,
                          seg_mask=mask)
        from . import fastpath
        fastpath.weights_changed()            # packed bf16 copies of the conv weights are stale now

    # ------------------------------------------------------------------------------------------------ checkpoints
    def _segment_params(self, s):
        return [(i, p, o) for i, (p, o) in enumerate(zip(self.params, self.offsets)) if self.segment_of[i] == s]

    def state_dict(self, s):
        """torch.optim.Adam.state_dict() layout for sub-network s (per-parameter 'step', 'exp_avg', 'exp_avg_sq')."""
        h = self.hyper[s]
        entries = self._segment_params(s)
        step = self.steps[s].detach().clone()
        state = {}
        if float(step.item()) > 0:
            for j, (_, p, o) in enumerate(entries):
                state[j] = {"step": step.clone(),
                            "exp_avg": self.exp_avg[o:o + p.numel()].view_as(p).clone(),
                            "exp_avg_sq": self.exp_avg_sq[o:o + p.numel()].view_as(p).clone()}
        group = {"lr": h["lr"], "betas": h["betas"], "eps": h["eps"], "weight_decay": h["weight_decay"],
                 "amsgrad": False, "maximize": False, "foreach": None, "capturable": True, "differentiable": False,
                 "fused": True, "params": list(range(len(entries)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, s, sd):
        """Copies a torch.optim.Adam state (this class's own, or one written by the reference) IN PLACE into the flat
        moment buffers -- safe after a CUDA graph of the step was captured."""
        entries = self._segment_params(s)
        groups = sd.get("param_groups", [])
        order = [i for g in groups for i in g["params"]]
        if groups and len(order) != len(entries):
            raise ValueError("optimizer state has %d parameters, the sub-network has %d" % (len(order), len(entries)))
        state = sd.get("state", {})
        step = 0.0
        with torch.no_grad():
            for j, (_, p, o) in enumerate(entries):
                key = order[j] if order else j
                st = state.get(key, state.get(str(key)))
                m = self.exp_avg[o:o + p.numel()].view_as(p)
                v = self.exp_avg_sq[o:o + p.numel()].view_as(p)
                if st is None:
                    m.zero_()
                    v.zero_()
                    continue
                m.copy_(st["exp_avg"])
                v.copy_(st["exp_avg_sq"])
                step = max(step, float(st["step"]))
            self.steps[s] = step
        if groups:
            g = groups[0]
            self.hyper[s].update(lr=float(g["lr"]), betas=(float(g["betas"][0]), float(g["betas"][1])),
                                 eps=float(g["eps"]), weight_decay=float(g.get("weight_decay", 0.0)))
            if g.get("amsgrad") or g.get("maximize"):
                raise NotImplementedError("amsgrad / maximize Adam states are not supported")


class _AdamView:
    """What `solver.optimizers[name]` holds: the torch.optim.Optimizer surface the reference uses."""

    def __init__(self, owner, segment):
        self.owner, self.segment = owner, segment

    @property
    def param_groups(self):
        h = self.owner.hyper[self.segment]             # a live dict: `group['lr'] = x` takes effect on the next step
        h.setdefault("params", [p for _, p, _ in self.owner._segment_params(self.segment)])
        return [h]

    def step(self, closure=None):
        loss = closure() if closure is not None else None
        self.owner.step([self.segment])
        return loss

    def zero_grad(self, set_to_none=False):
        # gradients stay views of the flat buffer (the kernels accumulate into them, NCCL reduces them): always in place
        self.owner.zero_grad(self.segment)

    def state_dict(self):
        return self.owner.state_dict(self.segment)

    def load_state_dict(self, sd):
        self.owner.load_state_dict(self.segment, sd)
