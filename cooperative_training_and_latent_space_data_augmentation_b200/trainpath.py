"""Forward AND backward of the reference-shaped FTN/STN modules (networks.py) on this build's kernels.

One torch.autograd.Function per sub-network (MyEncoder / Dual_Branch_Encoder / MyDecoder).  Its forward is the kernel
sequence of fastpath.py in train-mode BatchNorm ('track' = batch statistics + running-stat update, 'batch' = batch
statistics only, as inside _disable_tracking_bn_stats) and keeps the blocked C8 activations; its backward is written
out by hand on the backward kernels:

  head / stem      ctl_head_bwd_c8, ctl_stem_wgrad_c8, ctl_stem_dgrad_c8 (softmax(x/T) of construct_input fused)
  BN + LReLU/ReLU  ctl_bn_bwd_reduce_c8 + ctl_bn_bwd_apply_c8
  conv dgrad       ctl_conv2d_c8_bf16 with transposed + flipped weights (K3 itself)
  conv wgrad       ctl_conv_wgrad_c8_bf16 (K3w, tcgen05 with MN-major operands)
  resampling       ctl_downsample2x_sum_c8 (nearest x2), ctl_zero_stuff2x_c8 (stride-2 conv),
                   ctl_split_parity2x2_c8 (ConvTranspose2d k2 s2)

torch autograd only connects the sub-networks (losses, detach points), exactly where the reference's graph has them
(medseg/models/advanced_triplet_recon_segmentation_model.py:414-467, :525-559), so `loss.backward()` and the five Adam
optimizers are unchanged.  Activations and activation gradients are bf16 (C8); parameters and their gradients fp32.

Gradient of a convolution bias that feeds a train-mode BatchNorm is identically zero (the mean subtraction removes it;
the reference accumulates ~1e-9 rounding noise there): no gradient tensor is produced for those biases, so the
optimizer leaves them alone -- the network function does not depend on them.
"""
import contextlib

import torch
import torch.nn as nn

from . import ops
from .fastpath import _act_code, _packed

LRELU = ops.ACT_LRELU


# ------------------------------------------------------------------------------------------------ helpers
class _Fwd:
    """Per-forward scratch of one sub-network: ONE zeroed float64 arena for the BatchNorm sums the conv epilogues
    accumulate, and the num_batches_tracked counters to bump (one foreach add at the end instead of one per layer)."""

    def __init__(self, module, device):
        n = getattr(module, '_ctl_bn_channels', None)
        if n is None:
            n = sum(m.num_features for m in module.modules() if isinstance(m, nn.BatchNorm2d))
            module._ctl_bn_channels = n
        self.arena = torch.zeros(2 * n, device=device, dtype=torch.float64)
        self.used = 0
        self.tracked = []

    def sums(self, channels):
        out = self.arena[self.used:self.used + 2 * channels].view(2, channels)
        self.used += 2 * channels
        return out

    def finish(self):
        if self.tracked:
            torch._foreach_add_(self.tracked, 1)
            self.tracked = []


def _bn_mode_stats(fw, bn, a):
    """Train-mode BatchNorm statistics of the raw conv output `a`; honours track_running_stats like nn.BatchNorm2d."""
    track = bn.track_running_stats and bn.running_mean is not None
    scale, shift, mean, var = ops.bn_batch_affine_c8(
        a, bn.weight, bn.bias, bn.eps, bn.running_mean if track else None, bn.running_var if track else None,
        bn.momentum if bn.momentum is not None else 0.1, want_stats=True)
    if track:
        fw.tracked.append(bn.num_batches_tracked)
    return scale, shift, mean, var


def _conv_bn(fw, conv, bn, x):
    """Raw conv output (+bias) and the train-mode BatchNorm statistics of it.  When the conv's N tile allows it the
    statistics are accumulated by the conv epilogue itself (no second pass over the tensor)."""
    k = conv.kernel_size[0]
    if conv.stride[0] != 1 or not ops.conv_stats_fusable(conv.in_channels, conv.out_channels, k * k):
        a = _conv_raw(conv, x)
        return (a,) + _bn_mode_stats(fw, bn, a)
    sums = fw.sums(conv.out_channels)
    wp = _packed(conv.weight, ops.pack_conv_weight)
    a = ops.conv2d_c8(x, wp, conv.out_channels, k * k, shift=conv.bias, stats=sums)
    track = bn.track_running_stats and bn.running_mean is not None
    stats = ops.bn_affine_from_sums(sums, a.shape[0] * a.shape[2] * a.shape[3], bn.weight, bn.bias, bn.eps,
                                    bn.running_mean if track else None, bn.running_var if track else None,
                                    bn.momentum if bn.momentum is not None else 0.1)
    if track:
        fw.tracked.append(bn.num_batches_tracked)
    return (a,) + stats


def _conv_bn_apply(fw, conv, bn, x, act, low=None, packed=None):
    """h = act(BatchNorm_train(conv(x)) [+ up2(low)]) -> (a, h, scale, shift, mean, var).  With the statistics accumulated
    by the conv epilogue the BatchNorm finalisation runs in the prologue of the apply kernel: conv + ONE streaming launch."""
    k = conv.kernel_size[0]
    if conv.stride[0] != 1 or conv.out_channels > 256 or \
            not ops.conv_stats_fusable(conv.in_channels, conv.out_channels, k * k):
        a, scale, shift, mean, var = _conv_bn(fw, conv, bn, x)
        h = ops.scale_shift_act_c8(a, scale, shift, act) if low is None else \
            ops.scale_shift_upadd_act_c8(a, scale, shift, low, act)
        return a, h, scale, shift, mean, var
    sums = fw.sums(conv.out_channels)
    wp = packed if packed is not None else _packed(conv.weight, ops.pack_conv_weight)
    a = ops.conv2d_c8(x, wp, conv.out_channels, k * k, shift=conv.bias, stats=sums)
    track = bn.track_running_stats and bn.running_mean is not None
    h, scale, shift, mean, var = ops.bn_apply_from_sums_c8(
        a, sums, bn.weight, bn.bias, bn.eps, act, bn.running_mean if track else None,
        bn.running_var if track else None, bn.momentum if bn.momentum is not None else 0.1, low=low)
    if track:
        fw.tracked.append(bn.num_batches_tracked)
    return a, h, scale, shift, mean, var


def _pack_stem_weight(weight):
    return ops.pack_conv_weight(ops.pad_stem_weight(weight))


def _conv_raw(conv, x):
    k = conv.kernel_size[0]
    wp = _packed(conv.weight, ops.pack_conv_weight_s2 if (conv.stride[0] == 2 and k == 3) else ops.pack_conv_weight)
    return ops.conv2d_c8(x, wp, conv.out_channels, k * k, subsample=conv.stride[0], shift=conv.bias)


def _dgrad(conv, dy, res=None):
    """Input gradient of a stride-1 conv (dy at the conv's output resolution)."""
    k = conv.kernel_size[0]
    wp = _packed(conv.weight, ops.pack_conv_weight_dgrad, tag='dgrad')
    return ops.conv2d_c8(dy, wp, conv.in_channels, k * k, res=res)


def _dgrad_into_bn(grads, conv, dy, bn_saved, act):
    """Input gradient of the stride-1 3x3 `conv` whose result flows into the backward of act(BatchNorm(a)) of the layer
    in front of it (`bn_saved` = that layer's forward record).  Where the kernel allows it the BatchNorm-backward sums
    are accumulated by the convolution's epilogue (no reduction pass over dy and a): returns (dh, totals or None)."""
    a, scale, shift = bn_saved[1], bn_saved[5], bn_saved[6]
    if (conv.kernel_size[0] == 3 and act in (ops.ACT_LRELU, ops.ACT_RELU)
            and ops.conv_bnbwd_fusable(conv.out_channels, conv.in_channels)):
        totals = grads.bn_totals(conv.in_channels)
        if totals is not None:
            wp = _packed(conv.weight, ops.pack_conv_weight_dgrad, tag='dgrad')
            return ops.conv2d_c8_bnbwd(dy, wp, conv.in_channels, a, scale, shift, act, totals), totals
    return _dgrad(conv, dy), None


def _wgrad(grads, conv, x, dy):
    """K3w straight into nn.Conv2d's [Cout][Cin][k][k] layout, accumulated into a slice of the backward's zero arena."""
    k = conv.kernel_size[0]
    return ops.conv_wgrad_c8(x, dy, k * k, out=grads.buffer(conv.weight), layout='conv')


class SaliencyRequest:
    """While active (a `with` block around the saliency pass of model_util._mask_latent_code), the backward of the next
    decoder Function does not materialise dL/dz: its last input-gradient convolution accumulates the per-sample saliency
    sums in its epilogue (ops.conv2d_c8_saliency) into `sums` and `served` turns True."""
    active = None             # process-global on purpose: backward nodes run on autograd's device thread

    def __init__(self, mode, N, n, device):
        self.mode, self.N, self.n = mode, N, n
        self.sums = torch.zeros((N, n), device=device, dtype=torch.float64)
        self.served = False

    def __enter__(self):
        self.prev, SaliencyRequest.active = SaliencyRequest.active, self
        return self

    def __exit__(self, *exc):
        SaliencyRequest.active = self.prev
        return False


def saliency_fusable(decoder, code):
    """The fused form covers this build's own decoders on the kernel training route with a 128-channel CUDA fp32 code."""
    from . import networks
    return (isinstance(decoder, networks.MyDecoder) and networks._kernel_route(decoder, code) == 'train'
            and code.dim() == 4 and code.dtype == torch.float32 and code.shape[1] % 8 == 0
            and decoder.up1.conv_input.in_channels == code.shape[1])


# decoder -> its most recent training-route forward: {"z": input tensor, "version", "epoch", "out", "tape", "tracked"}.
# hard_example_generation differentiates decoder(code) w.r.t. code where `code` IS the latent the clean pass has just
# decoded with the same weights in the same BatchNorm mode (advanced...model.py:440-447 then :497-523): the forward the
# reference computes a second time is bit-for-bit the one already on tape.  model_util reuses it for the saliency pass
# (only the backward runs; the BatchNorm running-stat side effect of the skipped forward is replayed).
import weakref as _weakref

_FORWARD_CACHE = _weakref.WeakKeyDictionary()      # keyed by the decoder module itself: a dead model's tape goes with it


def forget_forwards():
    _FORWARD_CACHE.clear()


def cached_forward(dec, code):
    """The tape of decoder(code) when that exact forward is the decoder's most recent one, else None."""
    from . import fastpath
    hit = _FORWARD_CACHE.get(dec)
    if hit is None:
        return None
    z = hit["z"]
    tracked = bool(dec.training and all(m.track_running_stats for m in dec.modules() if isinstance(m, nn.BatchNorm2d)))
    if (z.data_ptr() != code.data_ptr() or tuple(z.shape) != tuple(code.shape) or z._version != hit["version"]
            or code._version != hit["version"] or hit["epoch"] != fastpath._WEIGHTS_EPOCH[0] or hit["tracked"] != tracked
            or not dec.training):
        return None
    return hit


def replay_bn_tracking(dec, tape):
    """The running-statistics update a train-mode forward performs, from the batch statistics on `tape` (the forward
    itself is not recomputed): running = (1 - m) * running + m * batch (unbiased variance), num_batches_tracked += 1."""
    rm, rv, means, variances, tracked, keep, m_mean, m_var = [], [], [], [], [], [], [], []
    blocks = (dec.up1, dec.up2, dec.up3, dec.up4)
    for blk, (x, s) in zip(blocks, tape[:4]):
        s1, a2, mean2, var2, out = s
        for bn, a, mean, var in ((blk.conv[1], s1[1], s1[3], s1[4]), (blk.conv[4], a2, mean2, var2)):
            if not (bn.track_running_stats and bn.running_mean is not None):
                continue
            count = a.shape[0] * a.shape[2] * a.shape[3]
            m = bn.momentum if bn.momentum is not None else 0.1
            rm.append(bn.running_mean); rv.append(bn.running_var)
            means.append(mean); variances.append(var)
            tracked.append(bn.num_batches_tracked)
            keep.append(1.0 - m); m_mean.append(m); m_var.append(m * count / max(1, count - 1))
    if rm:                                   # seven multi-tensor launches for the decoder's eight BatchNorm layers
        torch._foreach_mul_(rm, keep)
        torch._foreach_add_(rm, torch._foreach_mul(means, m_mean))
        torch._foreach_mul_(rv, keep)
        torch._foreach_add_(rv, torch._foreach_mul(variances, m_var))
        torch._foreach_add_(tracked, 1)


def saliency_from_tape(dec, hit, dout, request):
    """Backward of the cached decoder forward w.r.t. its latent input only, with the saliency sums fused into its last
    convolution (`request`); the tape is left intact for the training backward that follows."""
    params = tuple(dec.parameters())
    grads = _Grads(params, [False] * len(params))
    with request:
        decoder_bwd(dec, hit["tape"], dout.contiguous(), grads, True)
    return request.served


_DIRECT = {"on": False}      # process-global on purpose: backward nodes run on autograd's device thread


@contextlib.contextmanager
def accumulate_into_grads():
    """While active, a backward of the Functions below ADDS the gradient of every parameter that already owns a
    `.grad` tensor into that tensor itself -- the weight-gradient kernels accumulate atomically into it, the small
    per-channel vectors go through one multi-tensor add per sub-network -- and returns None for it, instead of handing
    ~280 tensors per pass to autograd's AccumulateGrad (one `add` launch each, per pass).  Same values, same place;
    only valid around a plain `loss.backward()` (never around torch.autograd.grad w.r.t. parameters)."""
    prev = _DIRECT["on"]
    _DIRECT["on"] = True
    try:
        yield
    finally:
        _DIRECT["on"] = prev


class _Grads(dict):
    """parameter -> gradient for the parameters autograd asked for (requires_grad AT FORWARD TIME, which is what the
    reference's graph records -- e.g. BatchNorm affine parameters are frozen inside _disable_tracking_bn_stats)."""

    def __init__(self, params, needs):
        super().__init__()
        wanted = [p for p, n in zip(params, needs) if n]
        self.need = {id(p) for p in wanted}
        self.direct = bool(wanted) and _DIRECT["on"] and all(
            p.grad is not None and p.grad.is_contiguous() and p.grad.dtype == torch.float32 for p in wanted)
        self.small_dst, self.small_src = [], []
        self.arena, self.used = None, 0
        # ONE zeroed float64 arena for the BatchNorm-backward totals of this backward pass (2 * C values per layer;
        # the deepest sub-network has 13 BatchNorm layers of <= 128 channels)
        self.bn_arena = torch.zeros(8192, device=params[0].device, dtype=torch.float64)
        self.bn_used = 0
        if not self.direct:
            # ONE zeroed fp32 arena for every accumulated (atomic) parameter gradient of this backward
            total = sum((p.numel() + 3) // 4 * 4 for p in wanted if p.dim() > 1)
            self.arena = torch.zeros(total + 256, device=params[0].device, dtype=torch.float32) if total else None

    def buffer(self, p, extra=0):
        """Zero-initialised (or, in direct mode, the live `.grad`) fp32 buffer shaped like parameter `p` that a kernel
        accumulates into; `extra` more elements follow it in the arena (head: weight | bias)."""
        if self.direct:
            return p.grad
        n = p.numel() + extra
        out = self.arena[self.used:self.used + n]
        self.used += (n + 3) // 4 * 4                      # keep slices 16-byte aligned
        return out.view(p.shape) if not extra else out

    def bn_totals(self, channels):
        n = 2 * channels
        if self.bn_used + n > self.bn_arena.numel():
            return None                                     # falls back to the three-launch form
        out = self.bn_arena[self.bn_used:self.bn_used + n]
        self.bn_used += n
        return out

    def wants(self, p):
        return p is not None and id(p) in self.need

    def add(self, p, g):
        if g is None or not self.wants(p):
            return
        if self.direct:
            if g.data_ptr() == p.grad.data_ptr():          # a kernel accumulated into .grad itself: done
                return
            if any(d is p.grad for d in self.small_dst):   # second contribution in one backward: not in one foreach
                p.grad.add_(g.view_as(p.grad))
            else:
                self.small_dst.append(p.grad)
                self.small_src.append(g.view_as(p.grad))
            return
        key = id(p)
        self[key] = g if key not in self else self[key] + g

    def results(self, params):
        if self.direct:
            if self.small_dst:
                torch._foreach_add_(self.small_dst, self.small_src)
            return (None,) * len(params)
        return tuple(self.get(id(p)) for p in params)


# ------------------------------------------------------------------------------------------------ conv + BN + act
def conv_bn_act_fwd(fw, conv, bn, x, act):
    a, h, scale, shift, mean, var = _conv_bn_apply(fw, conv, bn, x, act)
    return h, (x, a, h, mean, var, scale, shift)


def _bn_act_bwd(bn, act, saved, dh, grads, totals=None):
    """da, with the BatchNorm parameter gradients recorded; `totals`: the sums already accumulated by the convolution
    that produced dh (then only the apply pass runs)."""
    x, a, h, mean, var, scale, shift = saved
    if totals is not None:
        da, dg, db = ops.bn_bwd_apply_totals_c8(dh, a, act, mean, var, bn.eps, bn.weight, totals, (scale, shift))
    else:
        # act'(h) from sign(a*scale + shift): the backward never reads h (it stays alive only as the next layer's input)
        da, dg, db, _ = ops.bn_act_bwd_c8(dh, h, a, act, mean, var, bn.eps, bn.weight, act_affine=(scale, shift),
                                          totals=grads.bn_totals(a.shape[1] * 8))
    grads.add(bn.weight, dg)
    grads.add(bn.bias, db)
    return da


def conv_bn_act_bwd(conv, bn, act, saved, dh, grads, need_dx=True, totals=None, next_bn=None):
    """next_bn = (forward record, activation) of the conv-BN-act layer in FRONT of this one: its BatchNorm-backward sums
    are then accumulated by this layer's input-gradient convolution; returns (dx, totals for that layer) in that case."""
    x = saved[0]
    da = _bn_act_bwd(bn, act, saved, dh, grads, totals)
    if grads.wants(conv.weight):
        grads.add(conv.weight, _wgrad(grads, conv, x, da))
    if next_bn is not None:
        return _dgrad_into_bn(grads, conv, da, next_bn[0], next_bn[1]) if need_dx else (None, None)
    return _dgrad(conv, da) if need_dx else None


# ------------------------------------------------------------------------------------------------ residual body
def residual_fwd(fw, block, xr, x_low=None):
    """out = LReLU(conv_input(xr) + BN2(conv2(LReLU(BN1(conv1(xr))))))   (encoder_decoder.py:54-57, :334-337).
    x_low: the block input BEFORE a nearest x2 up-sampling (xr = up(x_low)).  A 1x1 convolution commutes with nearest
    up-sampling, so the shortcut is then taken at the low resolution (a quarter of the pixels) and added, up-sampled on
    the fly, by the kernel that applies BN2 + LReLU."""
    seq = block.conv
    h1, s1 = conv_bn_act_fwd(fw, seq[0], seq[1], xr, LRELU)
    ci = block.conv_input
    wp = _packed(ci.weight, ops.pack_conv_weight)
    if x_low is not None:
        c = ops.conv2d_c8(x_low, wp, ci.out_channels, 1, shift=ci.bias)
        a2, out, scale2, shift2, mean2, var2 = _conv_bn_apply(fw, seq[3], seq[4], h1, LRELU, low=c)
    else:
        a2, scale2, shift2, mean2, var2 = _conv_bn(fw, seq[3], seq[4], h1)
        out = ops.conv2d_c8(xr, wp, ci.out_channels, 1, shift=ci.bias, res=a2, res_scale=scale2, res_shift=shift2,
                            act=LRELU)
    return out, (s1, a2, mean2, var2, out)


def residual_bwd(block, saved, dout, grads, x_low=None):
    """Returns (gradient w.r.t. xr, dc): dc is None unless x_low was given -- then the shortcut's share of the input
    gradient still has to be added at the LOW resolution by the caller: dgrad(conv_input, dc)."""
    s1, a2, mean2, var2, out = saved
    xr, h1 = s1[0], s1[2]
    seq, ci = block.conv, block.conv_input
    bn2 = seq[4]
    # tail: d(pre) = dout * LReLU'(out) feeds both the 1x1 branch and BN2
    da2, dg2, db2, dpre = ops.bn_act_bwd_c8(dout, out, a2, LRELU, mean2, var2, bn2.eps, bn2.weight, want_dv=True,
                                            totals=grads.bn_totals(a2.shape[1] * 8))
    grads.add(bn2.weight, dg2)
    grads.add(bn2.bias, db2)
    dc = dxr = None
    if x_low is not None:
        dc = ops.downsample2x_sum_c8(dpre)                  # backward of the on-the-fly nearest up-sampling
        if grads.wants(ci.weight):
            grads.add(ci.weight, _wgrad(grads, ci, x_low, dc))
            grads.add(ci.bias, db2)
    else:
        if grads.wants(ci.weight):
            grads.add(ci.weight, _wgrad(grads, ci, xr, dpre))
            grads.add(ci.bias, db2)                             # sum over pixels of dpre == dbeta of BN2
        dxr = _dgrad(ci, dpre)
    # conv2
    if grads.wants(seq[3].weight):
        grads.add(seq[3].weight, _wgrad(grads, seq[3], h1, da2))
    dh1, totals1 = _dgrad_into_bn(grads, seq[3], da2, s1, LRELU)
    # BN1 + LReLU + conv1
    da1 = _bn_act_bwd(seq[1], LRELU, s1, dh1, grads, totals1)
    if grads.wants(seq[0].weight):
        grads.add(seq[0].weight, _wgrad(grads, seq[0], xr, da1))
    return _dgrad(seq[0], da1, res=dxr), dc


# ------------------------------------------------------------------------------------------------ down / up blocks
def down_fwd(fw, block, x):
    xd = _conv_raw(block.down, x)                          # 3x3 stride 2 (no norm / activation)
    out, s = residual_fwd(fw, block, xd)
    return out, (x, s)


def down_bwd(block, saved, dout, grads, need_dx=True, next_bn=None):
    """next_bn: see conv_bn_act_bwd (the block's input is the output of a conv-BN-act layer); returns (dx, totals) then."""
    x, s = saved
    dxd, _ = residual_bwd(block, s, dout, grads)
    down = block.down
    want_w = grads.wants(down.weight)
    if not (want_w or need_dx):
        return (None, None) if next_bn is not None else None
    dyz = ops.zero_stuff2x_c8(dxd)                          # the stride-2 output gradient seen at full resolution
    if want_w:
        grads.add(down.weight, _wgrad(grads, down, x, dyz))
        grads.add(down.bias, ops.channel_sum_c8(dxd))
    if next_bn is not None:
        return _dgrad_into_bn(grads, down, dyz, next_bn[0], next_bn[1]) if need_dx else (None, None)
    return _dgrad(down, dyz) if need_dx else None


def _convT_tap_weight(up, d):
    return up.weight.detach()[:, :, d // 2, d % 2].reshape(up.in_channels, up.out_channels, 1, 1)


def up_fwd(fw, block, x):
    if block.up_type == 'NN':
        xu = ops.upsample2x_c8(x)
        out, s = residual_fwd(fw, block, xu, x_low=x)
    else:
        up = block.up
        wp = _packed(up.weight, ops.pack_convtranspose2x2_weight)
        xu = ops.conv2d_c8(x, wp, 4 * up.out_channels, 1, up2x=True, shift=up.bias.detach().repeat(4))
        out, s = residual_fwd(fw, block, xu)
    return out, (x, s)


def _dgrad_saliency(sal, weight, fn, tag, dy, cout, res):
    """Last input-gradient convolution of a decoder with the saliency sums fused into its epilogue; dL/dz is not stored."""
    wp = _packed(weight, fn, tag=tag)
    ops.conv2d_c8_saliency(dy, wp, cout, res, sal.sums, sal.mode, store_out=False)
    sal.served = True
    return None


def up_bwd(block, saved, dout, grads, need_dx=True, sal=None):
    x, s = saved
    if sal is not None and (sal.sums.shape[0] != x.shape[0] or
                            sal.n != (x.shape[1] * 8 if sal.mode == ops.MODE_CHANNEL else x.shape[2] * x.shape[3])):
        sal = None                                              # a request for another tensor shape: ignore it
    if block.up_type == 'NN':
        dxu, dc = residual_bwd(block, s, dout, grads, x_low=x)
        if not need_dx:
            return None
        # main branch back through the up-sampling (2x2 sums) + the shortcut's input gradient, both at low resolution
        ci = block.conv_input
        if sal is not None and ci.out_channels in (64, 128):
            return _dgrad_saliency(sal, ci.weight, ops.pack_conv_weight_dgrad, 'dgrad', dc, ci.in_channels,
                                   ops.downsample2x_sum_c8(dxu))
        return _dgrad(block.conv_input, dc, res=ops.downsample2x_sum_c8(dxu))
    dxu, _ = residual_bwd(block, s, dout, grads)
    up = block.up
    want_w = grads.wants(up.weight)
    if not (want_w or need_dx):
        return None
    parts = ops.split_parity2x2_c8(dxu)                     # [4][N, C/8, H, W, 8]: dy per kernel tap
    if want_w:
        dW = grads.buffer(up.weight)                       # [ci][co][2][2], tap d written with stride 4
        for d in range(4):
            ops.conv_wgrad_c8(x, parts[d], 1, out=dW, layout=('convT', d))
        grads.add(up.weight, dW)
        grads.add(up.bias, ops.channel_sum_c8(dxu))
    dx = None
    if need_dx:
        for d in range(4):
            fn = lambda w, d=d: ops.pack_conv_weight(_convT_tap_weight(up, d))      # noqa: E731
            if d == 3 and sal is not None and up.out_channels in (64, 128):
                return _dgrad_saliency(sal, up.weight, fn, 'dgradT3', parts[3], up.in_channels, dx)
            wp = _packed(up.weight, fn, tag='dgradT%d' % d)
            dx = ops.conv2d_c8(parts[d], wp, up.in_channels, 1, res=dx)
    return dx


# ------------------------------------------------------------------------------------------------ encoder
def encoder_fwd(enc, x, in_mode, temperature):
    """MyEncoder: planar input (fp32 image / logits, or int64 label map with in_mode 2) -> C8 latent."""
    inc = enc.inc
    fw = _Fwd(enc, inc[0].weight.device)
    # stem on the tensor core: the input (image / softmax(logits/T) / one-hot labels) becomes a 16-channel C8 tensor of
    # bf16 (hi | lo | hi) channel groups, the weight [16,16,3,3] = (w_hi | w_hi | w_lo | 0): fp32-like products on K3,
    # which then also accumulates the BatchNorm statistics
    xin = ops.stem_input_c8(x, inc[0].in_channels, in_mode, temperature)
    wp = _packed(inc[0].weight, _pack_stem_weight, tag='stem')
    sums = fw.sums(inc[0].out_channels)
    a0 = ops.conv2d_c8(xin, wp, inc[0].out_channels, 9, shift=inc[0].bias, stats=sums)
    bn0 = inc[1]
    track = bn0.track_running_stats and bn0.running_mean is not None
    h0, scale0, shift0, mean0, var0 = ops.bn_apply_from_sums_c8(
        a0, sums, bn0.weight, bn0.bias, bn0.eps, LRELU, bn0.running_mean if track else None,
        bn0.running_var if track else None, bn0.momentum if bn0.momentum is not None else 0.1)
    if track:
        fw.tracked.append(bn0.num_batches_tracked)
    h, s_inc = conv_bn_act_fwd(fw, inc[3], inc[4], h0, LRELU)   # BN then F.leaky_relu (encoder_decoder.py:405)
    tape = [(a0, h0, mean0, var0, scale0, shift0, xin), s_inc]
    for blk in (enc.down1, enc.down2, enc.down3, enc.down4):
        h, s = down_fwd(fw, blk, h)
        tape.append(s)
    z, s_fin = conv_bn_act_fwd(fw, enc.final_conv[0], enc.final_conv[1], h, _act_code(enc.act))
    tape.append(s_fin)
    fw.finish()
    return z, tape


def encoder_bwd(enc, tape, dz, grads, x, in_mode, temperature, need_dx):
    inc = enc.inc
    d = conv_bn_act_bwd(enc.final_conv[0], enc.final_conv[1], _act_code(enc.act), tape[6], dz, grads)
    for i, blk in zip((5, 4, 3), (enc.down4, enc.down3, enc.down2)):
        d = down_bwd(blk, tape[i], d, grads)
    # the 224^2 level: every input-gradient convolution accumulates the BatchNorm-backward sums of the layer in front
    d, t4 = down_bwd(enc.down1, tape[2], d, grads, next_bn=(tape[1], LRELU))
    a0, h0, mean0, var0, scale0, shift0, xin = tape[0]
    rec0 = (xin, a0, h0, mean0, var0, scale0, shift0)
    dh0, t1 = conv_bn_act_bwd(inc[3], inc[4], LRELU, tape[1], d, grads, totals=t4, next_bn=(rec0, LRELU))
    da0 = _bn_act_bwd(inc[1], LRELU, rec0, dh0, grads, t1)
    if grads.wants(inc[0].weight):
        # K3w on the 16-channel stem input; the gradient of the real input channels is the leading slice
        dW16 = ops.conv_wgrad_c8(xin, da0, 9, layout='conv')
        grads.add(inc[0].weight, ops.stem_weight_grad(dW16, inc[0].in_channels))
    if need_dx:
        return ops.stem_dgrad_c8(da0, x, inc[0].weight, in_mode, temperature)
    return None


def decoupler_fwd(seq, z):
    fw = _Fwd(seq, z.device)
    h, s1 = conv_bn_act_fwd(fw, seq[0], seq[1], z, LRELU)
    out, s2 = conv_bn_act_fwd(fw, seq[3], seq[4], h, _act_code(seq[5]))
    fw.finish()
    return out, (s1, s2)


def decoupler_bwd(seq, saved, dout, grads):
    s1, s2 = saved
    dh = conv_bn_act_bwd(seq[3], seq[4], _act_code(seq[5]), s2, dout, grads)
    return conv_bn_act_bwd(seq[0], seq[1], LRELU, s1, dh, grads)


# ------------------------------------------------------------------------------------------------ decoder
def decoder_fwd(dec, z_c8):
    tape = []
    y = z_c8
    fw = _Fwd(dec, z_c8.device)
    for blk in (dec.up1, dec.up2, dec.up3, dec.up4):
        y, s = up_fwd(fw, blk, y)
        tape.append(s)
    fw.finish()
    act = _act_code(dec.last_act)
    out = ops.head_conv_c8(y, dec.final_conv.weight, dec.final_conv.bias, act)
    tape.append((y, out if act != ops.ACT_NONE else None))
    return out, tape


def decoder_bwd(dec, tape, dout, grads, need_dz):
    y, out = tape[4]
    fc = dec.final_conv
    want_fc = grads.wants(fc.weight) and grads.wants(fc.bias)
    if want_fc and grads.direct:
        gbuf = (fc.weight.grad, fc.bias.grad)
    else:
        gbuf = grads.buffer(fc.weight, extra=fc.out_channels) if (want_fc and grads.arena is not None) else None
    dy, dW, db = ops.head_bwd_c8(dout, out, y, fc.weight, _act_code(dec.last_act), out=gbuf)
    grads.add(fc.weight, dW)
    grads.add(fc.bias, db)
    d = dy
    blocks = (dec.up1, dec.up2, dec.up3, dec.up4)
    sal = SaliencyRequest.active if need_dz else None
    for i in (3, 2, 1, 0):
        d = up_bwd(blocks[i], tape[i], d, grads, need_dx=(i > 0 or need_dz), sal=sal if i == 0 else None)
    return d


# ------------------------------------------------------------------------------------------------ autograd glue
def _param_grads(params, grads):
    return grads.results(params)


class _DecoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dec, z, *params):
        z_c8 = ops.c8_twin(z)               # written by the masking kernel itself when z is a freshly masked code
        out, tape = decoder_fwd(dec, z_c8 if z_c8 is not None else ops.nchw_to_c8(z))
        ctx.dec, ctx.tape, ctx.params = dec, tape, params
        ctx.z_shape = tuple(z.shape)
        from . import fastpath
        _FORWARD_CACHE[dec] = {
            "z": z, "version": z._version, "epoch": fastpath._WEIGHTS_EPOCH[0], "out": out, "tape": tape,
            "tracked": all(m.track_running_stats for m in dec.modules() if isinstance(m, nn.BatchNorm2d))}
        return out

    @staticmethod
    def backward(ctx, dout):
        if ctx.tape is None:
            raise RuntimeError("this sub-network's activations were released by its first backward pass "
                               "(retain_graph / a second backward through the kernel path is not supported)")
        grads = _Grads(ctx.params, ctx.needs_input_grad[2:])
        sal = SaliencyRequest.active
        served0 = sal.served if sal is not None else False
        dz = decoder_bwd(ctx.dec, ctx.tape, dout.contiguous(), grads, ctx.needs_input_grad[1])
        ctx.tape = None
        if dz is None and ctx.needs_input_grad[1] and sal is not None and sal.served and not served0:
            # the saliency pass: dL/dz was reduced on-chip and never stored; autograd still wants a tensor of the
            # right shape for the (ignored) input gradient -- an allocation, no kernel
            return (None, torch.empty(ctx.z_shape, device=dout.device, dtype=torch.float32)) + _param_grads(ctx.params, grads)
        dz = ops.c8_to_nchw(dz) if dz is not None else None
        return (None, dz) + _param_grads(ctx.params, grads)


class _EncoderFn(torch.autograd.Function):
    """MyEncoder, optionally followed by the code decoupler of Dual_Branch_Encoder (two outputs then)."""

    @staticmethod
    def forward(ctx, enc, decoupler, in_mode, temperature, x, *params):
        z, tape = encoder_fwd(enc, x, in_mode, temperature)
        ctx.enc, ctx.decoupler, ctx.tape, ctx.params = enc, decoupler, tape, params
        ctx.in_mode, ctx.temperature, ctx.x = in_mode, temperature, x
        z_i = ops.c8_to_nchw(z)
        if decoupler is None:
            ctx.tape2 = None
            return z_i
        zs, ctx.tape2 = decoupler_fwd(decoupler, z)
        return z_i, ops.c8_to_nchw(zs)

    @staticmethod
    def backward(ctx, dz_i, dz_s=None):
        if ctx.tape is None:
            raise RuntimeError("this sub-network's activations were released by its first backward pass "
                               "(retain_graph / a second backward through the kernel path is not supported)")
        grads = _Grads(ctx.params, ctx.needs_input_grad[5:])
        dz = ops.nchw_to_c8(dz_i) if dz_i is not None else None
        if ctx.decoupler is not None and dz_s is not None:
            d2 = decoupler_bwd(ctx.decoupler, ctx.tape2, ops.nchw_to_c8(dz_s), grads)
            dz = d2 if dz is None else ops.nchw_to_c8(dz_i + ops.c8_to_nchw(d2))
        need_dx = ctx.needs_input_grad[4] and ctx.in_mode != 2
        dx = encoder_bwd(ctx.enc, ctx.tape, dz, grads, ctx.x, ctx.in_mode, ctx.temperature, need_dx)
        ctx.tape = ctx.tape2 = ctx.x = None
        return (None, None, None, None, dx) + _param_grads(ctx.params, grads)


class _DecouplerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, seq, z, *params):
        out, ctx.tape = decoupler_fwd(seq, ops.nchw_to_c8(z))
        ctx.seq, ctx.params = seq, params
        return ops.c8_to_nchw(out)

    @staticmethod
    def backward(ctx, dout):
        if ctx.tape is None:
            raise RuntimeError("this sub-network's activations were released by its first backward pass "
                               "(retain_graph / a second backward through the kernel path is not supported)")
        grads = _Grads(ctx.params, ctx.needs_input_grad[2:])
        dz = decoupler_bwd(ctx.seq, ctx.tape, ops.nchw_to_c8(dout), grads)
        ctx.tape = None
        return (None, ops.c8_to_nchw(dz)) + _param_grads(ctx.params, grads)


def _trainable_kernel_path(module, x):
    """The kernel path covers CUDA tensors in train mode (batch statistics).  Eval-mode forward with autograd enabled
    is not on the reference's hot path (predict / decoder_inference(eval=True) run under no_grad)."""
    return x.is_cuda and module.training


def decoder_apply(dec, z):
    params = tuple(dec.parameters())
    return _DecoderFn.apply(dec, z.float(), *params)


def encoder_apply(enc, x, in_mode=0, temperature=1.0):
    params = tuple(enc.parameters())
    if in_mode != 2:
        x = x.float()
    return _EncoderFn.apply(enc, None, in_mode, float(temperature), x, *params)


def dual_encoder_apply(dual, x):
    params = tuple(dual.general_encoder.parameters()) + tuple(dual.code_decoupler.parameters())
    return _EncoderFn.apply(dual.general_encoder, dual.code_decoupler, 0, 1.0, x.float(), *params)


def decoupler_apply(dual, z):
    return _DecouplerFn.apply(dual.code_decoupler, z.float(), *tuple(dual.code_decoupler.parameters()))
