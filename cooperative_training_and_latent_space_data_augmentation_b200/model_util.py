"""Drop-in for the hot-path functions of the reference's `medseg/models/model_util.py`.

Same names, positional order, defaults, return tuples and exceptions as the reference
(SURVEY.md section 8b); underneath, everything after `autograd.grad` runs in the sm_100a kernels
of libctl_b200.so (K1 saliency reduce, K2 top-p select + mask build + 128-bit apply).

    mask_latent_code_channel_wise   <- medseg/models/model_util.py:180-255
    mask_latent_code_spatial_wise   <- medseg/models/model_util.py:258-318
    makeVariable                    <- medseg/models/model_util.py:603-618
    make_one_hot                    <- medseg/models/model_util.py:168-177
    cross_entropy_2D                <- medseg/models/model_util.py:104-135
    set_grad                        <- medseg/models/model_util.py:163-165
    _disable_tracking_bn_stats      <- medseg/models/model_util.py:414-451

Random numbers.  The reference consumes two generators here: numpy's global one for the random
percentile and torch's device generator for the soft-mask values.  Both are kept, in the same
order, in the default `rng_mode == "torch"`, so a run seeded like the reference produces the
reference's masks.  `set_rng_mode("philox", seed, first_sample)` switches the device draw to the
in-kernel counter-based generator (no [N,n] rand tensor round trip, shard-invariant under data
parallelism); the numpy draw stays on the host because every rank must see the same percentile.
"""
import contextlib

import numpy as np
import torch
import torch.nn.functional as F

from . import ops

_RNG_MODE = "torch"
_NATIVE_RNG = ops.NativeRNG(0)


def set_rng_mode(mode, seed=0, first_sample=0):
    """'torch': soft-mask / dropout draws come from torch's device generator (reference-compatible).
    'philox': drawn inside the kernels from Philox4x32-10(seed, call counter, global element index)."""
    global _RNG_MODE, _NATIVE_RNG
    if mode not in ("torch", "philox"):
        raise ValueError("rng mode must be 'torch' or 'philox'")
    _RNG_MODE = mode
    _NATIVE_RNG = ops.NativeRNG(seed, first_sample)


def get_rng_mode():
    return _RNG_MODE


def native_rng():
    return _NATIVE_RNG


_STEP_PARAMS = None
_REUSE_FORWARD = True       # the saliency pass reuses the clean pass's decoder forward when it is the same computation
_FUSED_SALIENCY = True      # tests switch it off to compare with the materialised-gradient chain (K1 -> select -> K2)


def set_reuse_forward(flag):
    """True (default): when mask_latent_code_* is asked to differentiate decoder(code) and that exact forward (same
    storage, weights and BatchNorm mode) is the decoder's most recent one -- the case inside hard_example_generation --
    only the backward runs; False: the forward is recomputed as the reference does."""
    global _REUSE_FORWARD
    _REUSE_FORWARD = bool(flag)


def set_fused_saliency(flag):
    """True (default): on the kernel training route the latent saliency is reduced in the epilogue of the decoder's last
    input-gradient convolution and dL/dz is never materialised; False: autograd.grad + K1 + select + K2."""
    global _FUSED_SALIENCY
    _FUSED_SALIENCY = bool(flag)



@contextlib.contextmanager
def recording_step_params(step_params):
    """Inside the context (the CUDA-graph capture pass of training.GraphedCooperativeTrainer) every masking /
    dropout call launches the device-parameter form of its kernel on the next row of `step_params`."""
    global _STEP_PARAMS
    prev, _STEP_PARAMS = _STEP_PARAMS, step_params
    try:
        yield step_params
    finally:
        _STEP_PARAMS = prev


def step_params():
    return _STEP_PARAMS


# ------------------------------------------------------------------------------------------------
def makeVariable(tensor, use_gpu=True, type='long', requires_grad=True):
    """Detached leaf of the requested type.  The build is CUDA-only: `use_gpu=False` keeps the
    tensor where it is (the masking entry points then refuse CPU tensors)."""
    tensor = tensor.detach()
    if type == 'long':
        tensor = tensor.long()
    elif type == 'float':
        tensor = tensor.float()
    else:
        raise NotImplementedError
    if use_gpu:
        tensor = tensor.cuda()
    if requires_grad and tensor.is_floating_point():
        # .float()/.cuda() may have returned the caller's own storage: never flip its flag in place
        tensor = tensor.detach().requires_grad_(True)
    return tensor


def set_grad(module, requires_grad=False):
    for p in module.parameters():
        p.requires_grad = requires_grad


def make_one_hot(y, num_classes=4):
    """[N,H,W] integer label map -> [N,num_classes,H,W] float32 (a permuted view, like the reference)."""
    n, h, w = y.size(0), y.size(1), y.size(2)
    onehot = torch.zeros(n * h * w, num_classes, dtype=torch.float32, device=y.device)
    onehot.scatter_(1, y.reshape(n * h * w, 1), 1)
    return onehot.view(n, h, w, num_classes).permute(0, 3, 1, 2)


def cross_entropy_2D(input, target, weight=None, size_average=True):
    """Saliency cross entropy: label-map targets -> sum NLL / (numel + 1e-10); 4-D targets are
    treated as logits (softmax) and give -mean(q * log p)."""
    n, c, h, w = input.size()
    if weight is None and ops.ce2d_supported(input, target):
        return ops.cross_entropy_2d(input, target, 1.0 / float(target.numel() + 1e-10) if size_average else 1.0)
    log_p = F.log_softmax(input, dim=1)
    if target.dim() == 3:
        if weight is not None:
            weight = torch.softmax(weight, dim=0) * c
        loss = F.nll_loss(log_p, target, weight=weight, reduction='sum')
        if size_average:
            loss = loss / float(target.numel() + 1e-10)
        return loss
    if target.dim() == 4:
        q = F.softmax(target, dim=1)
        prod = q * log_p
        if weight is None:
            return -1 * torch.mean(torch.mean(prod.permute(0, 2, 3, 1).reshape(-1, c), dim=1))
        weight = torch.softmax(weight, dim=0) * c
        flat = prod.permute(0, 2, 3, 1).reshape(-1, c)
        total = 0.
        for i in range(c):
            total = total + torch.mean(flat[:, i] * weight[i])
        return -1 * total
    raise NotImplementedError


@contextlib.contextmanager
def _disable_tracking_bn_stats(model):
    """Inside the context every BatchNorm2d of `model` normalises with batch statistics but leaves
    running_mean / running_var / num_batches_tracked untouched, and its gamma/beta do not require
    grad; on exit the previous track_running_stats flags are restored (and, like the reference,
    gamma/beta.requires_grad is set to that flag)."""
    saved = []
    for module in model.modules():
        if isinstance(module, torch.nn.BatchNorm2d):
            saved.append((module, module.track_running_stats))
            module.track_running_stats = False
            if getattr(module, 'weight', None) is not None:
                module.weight.requires_grad_(False)
            if getattr(module, 'bias', None) is not None:
                module.bias.requires_grad_(False)
    try:
        yield
    finally:
        for module, flag in saved:
            module.track_running_stats = flag
            if getattr(module, 'weight', None) is not None:
                module.weight.requires_grad_(flag)
            if getattr(module, 'bias', None) is not None:
                module.bias.requires_grad_(flag)


# ------------------------------------------------------------------------------------------------
class _MaskedCode(torch.autograd.Function):
    """z * mask with the product computed by K2.  Used when `if_detach=False`, where the reference
    multiplies the *original* latent (graph attached) by the mask: backward is grad * mask."""

    @staticmethod
    def forward(ctx, latent, masked_value, mask_all):
        ctx.save_for_backward(mask_all)
        return masked_value.view_as(latent)

    @staticmethod
    def backward(ctx, grad_out):
        (mask_all,) = ctx.saved_tensors
        return grad_out * mask_all, None, None


def _latent_gradient(latent_code, decoder_function, label, num_classes, loss_type, fused_mode=None):
    """Returns (code leaf, dL/dcode, saliency request or None).  fused_mode (ops.MODE_*) asks for the on-chip saliency
    reduction: when this build's own decoder serves it, the request comes back `served` with the per-sample sums and the
    returned gradient tensor is an uninitialised placeholder (dL/dz was never stored)."""
    if not latent_code.is_cuda:
        raise RuntimeError("this build runs the latent masking on CUDA only (sm_100a kernels, no CPU fallback); "
                           "got a latent code on %s" % latent_code.device)
    from . import trainpath
    code = makeVariable(latent_code, use_gpu=True, type='float', requires_grad=True)
    request = None
    if fused_mode is not None and loss_type in ('mse', 'ce') and trainpath.saliency_fusable(decoder_function, code):
        N, C, H, W = code.shape
        request = trainpath.SaliencyRequest(fused_mode, N, C if fused_mode == ops.MODE_CHANNEL else H * W, code.device)
    if request is not None and _REUSE_FORWARD:
        # `code` is the latent the clean pass has just decoded (same storage, same weights, same BatchNorm mode): the
        # forward the reference recomputes here is already on tape -- only its backward w.r.t. the code runs, and the
        # skipped forward's running-statistics update is replayed
        hit = trainpath.cached_forward(decoder_function, code)
        if hit is not None:
            out = hit["out"]
            if loss_type == 'mse' and label.dim() == code.dim() and ops.sse_supported(out, label):
                dout = ops.sse_grad(out, label, 1.0 / out.numel())
            elif loss_type == 'ce' and ops.ce2d_supported(out, label):
                dout = ops.ce2d_grad(out, label, 1.0 / float(label.numel() + 1e-10))
            else:
                dout = None
            if dout is not None and trainpath.saliency_from_tape(decoder_function, hit, dout, request):
                if hit["tracked"]:
                    trainpath.replay_bn_tracking(decoder_function, hit["tape"])
                return code, None, request
    gt_y = make_one_hot(label, num_classes) if label.dim() < code.dim() else label
    with (request if request is not None else contextlib.nullcontext()):
        if loss_type == 'corr':
            loss = torch.mean(decoder_function(code) * gt_y)
        elif loss_type == 'mse':
            pred = decoder_function(code)
            if ops.sse_supported(pred, gt_y):
                loss = ops.squared_error(pred, gt_y, 1.0 / pred.numel())   # torch.mean((pred - gt)**2), one kernel each way
            else:
                loss = torch.mean((pred - gt_y) ** 2)
        elif loss_type == 'ce':
            loss = torch.mean(cross_entropy_2D(input=decoder_function(code), target=label, weight=None,
                                               size_average=True))
        else:
            # the reference falls through to an unbound `loss` here -> UnboundLocalError
            raise UnboundLocalError("loss_type %r is not one of 'corr', 'mse', 'ce'" % (loss_type,))
        gradient = torch.autograd.grad(loss, [code])[0]
    return code, gradient, request


def _mask_latent_code(mode, latent_code, decoder_function, label, num_classes, percentile, random, loss_type,
                      if_detach, if_soft):
    code, gradient, request = _latent_gradient(latent_code, decoder_function, label, num_classes, loss_type,
                                               fused_mode=mode if _FUSED_SALIENCY else None)
    N, C, H, W = code.shape
    n = C if mode == ops.MODE_CHANNEL else H * W
    if random:
        percentile = np.random.rand() * percentile          # numpy GLOBAL generator, as in the reference
    k = int(n * percentile)
    if k >= n or k < -n:
        # same exception, same message shape, raised before any device random number is drawn
        raise IndexError("index {} is out of bounds for dimension 1 with size {}".format(k, n))
    if k < 0:
        k += n                                               # python-style negative index, like tensor[:, k]
    rand = rng = None
    if if_soft:
        if _RNG_MODE == "torch":
            rand = torch.rand((N, n), device=code.device, dtype=torch.float32)   # == torch.rand_like(s)
        else:
            rng = _NATIVE_RNG
    if request is not None and request.served:
        # dL/dz was reduced inside the decoder's last input-gradient convolution; one kernel per call finishes the job
        # and also writes the masked code in the decoder's own blocked layout
        masked, mask, _, _, c8 = ops.saliency_sums_mask_apply(request.sums, code.detach(), mode, k, soft=if_soft,
                                                              rand=rand, rng=rng, step_params=_STEP_PARAMS)
        ops.register_c8_twin(masked, c8)
    else:
        masked, mask, _, _ = ops.saliency_mask_apply(gradient, code.detach(), mode, k, soft=if_soft, rand=rand, rng=rng,
                                                     out_dtype=torch.float32, step_params=_STEP_PARAMS)
    mask_all = mask.view(N, C, 1, 1) if mode == ops.MODE_CHANNEL else mask.view(N, 1, H, W)
    if not if_detach:
        # graph stays attached to the caller's latent (model_util.py:246-247)
        if latent_code.dtype == torch.float32 and latent_code.requires_grad:
            masked_latent_code = _MaskedCode.apply(latent_code, masked, mask_all)
        elif latent_code.dtype == torch.float32:
            masked_latent_code = masked
        else:
            masked_latent_code = latent_code * mask_all      # type promotion exactly as torch would do it
    else:
        # the reference returns `code * mask_all`: a non-leaf hanging off the fp32 leaf `code`
        masked_latent_code = _MaskedCode.apply(code, masked, mask_all)
    # model_util.py:251-254.  The decoder's parameter gradients are cleared IN PLACE (nn.Module.zero_grad as of the
    # reference's pinned torch 1.9, set_to_none=False): freeing them here would leave a captured CUDA graph of the
    # step zeroing / accumulating into gradient buffers that no longer exist, and would detach them from the
    # data-parallel gradient bucket.  Callables without that signature get the reference's bare call.
    flat = getattr(decoder_function, '_ctl_flat_segment', None)
    if flat is not None and flat[0].attached():
        flat[0].zero_grad(flat[1])                          # one memset of the decoder's slice of the flat gradients
    else:
        try:
            try:
                decoder_function.zero_grad(set_to_none=False)
            except TypeError:
                decoder_function.zero_grad()
        except Exception:  # noqa: BLE001  (the reference swallows everything here)
            pass
    return masked_latent_code, mask_all


def mask_latent_code_channel_wise(latent_code, decoder_function, label, num_classes=2, percentile=1 / 3.0,
                                  random=False, loss_type='corr', if_detach=True, if_soft=False):
    """Masks the top `percentile` channels of `latent_code` ranked by the mean over space of
    d loss / d code (loss between decoder_function(code) and `label`).  Returns
    (masked_latent_code [N,C,H,W] fp32, mask_all [N,C,1,1] fp32)."""
    return _mask_latent_code(ops.MODE_CHANNEL, latent_code, decoder_function, label, num_classes, percentile, random,
                             loss_type, if_detach, if_soft)


def mask_latent_code_spatial_wise(latent_code, decoder_function, label, num_classes, percentile=1 / 3.0,
                                  random=False, loss_type='corr', if_detach=True, if_soft=False):
    """Masks the top `percentile` spatial positions ranked by the mean over channels of
    d loss / d code.  Returns (masked_latent_code [N,C,H,W] fp32, mask_all [N,1,H,W] fp32)."""
    return _mask_latent_code(ops.MODE_SPATIAL, latent_code, decoder_function, label, num_classes, percentile, random,
                             loss_type, if_detach, if_soft)
