"""No-grad forward of the reference-shaped FTN/STN modules (networks.py) entirely in this build's kernels:
K3 (tcgen05 implicit-GEMM conv, conv_tc.cu) + the C8 streaming kernels (c8_ops.cu).  Activations stay in
the blocked C8 layout between layers; only module inputs/outputs are planar NCHW (the image, the logits,
and the latent codes that the masking API exposes).

BatchNorm handling (`bn` argument):
  'eval'   running statistics folded into the conv epilogue (inference; reference: module.eval())
  'batch'  batch statistics, running stats untouched (reference: train mode inside _disable_tracking_bn_stats,
           e.g. decoder_inference(..., disable_track_bn_stats=True), advanced...model.py:396-412)
  'track'  batch statistics + momentum update of running_mean/var and num_batches_tracked (plain train mode)

Kernel sequence of one residual block on the resampled input x' (encoder_decoder.py:54-57, :334-337):
  eval : y1 = K3(x', W1, BN1-folded, LReLU); y2 = K3(y1, W2, BN2-folded); out = K3_1x1(x', Win, +y2, LReLU)
  batch: y1 = K3(x', W1, +b1) -> stats -> scale/shift+LReLU (in place) ; y2 = K3(y1, W2, +b2) -> stats ->
         out = K3_1x1(x', Win, + BN2(y2) as the epilogue's residual affine, LReLU)
"""
import torch
import torch.nn as nn

from . import ops

# Bumped after EVERY optimizer step (global post-step hook below).  Fused / foreach optimizers update parameters without
# bumping Tensor._version, so the version counter alone would leave the packed bf16 copies stale -- and training on
# stale forward weights silently stops learning.
_WEIGHTS_EPOCH = [0]


def weights_changed():
    """Call after modifying parameters through a path autograd's version counter does not see."""
    _WEIGHTS_EPOCH[0] += 1


def _after_optimizer_step(optimizer, args, kwargs):
    _WEIGHTS_EPOCH[0] += 1


from torch.optim.optimizer import register_optimizer_step_post_hook as _register_post_hook  # noqa: E402

_register_post_hook(_after_optimizer_step)


class _PackRegistry:
    """Every (conv weight, forward | input-gradient) packing the networks have asked for, with a PERSISTENT bf16 output
    buffer each, so that after an optimizer step ONE launch (ctl_pack_conv_weights_batched, job table in device memory)
    refreshes all of them instead of ~160 three-microsecond launches.  The first request for a weight packs it alone and
    adds it to the table; a weight whose storage moved is re-registered."""

    def __init__(self):
        self.entries = {}           # (id(param), tag) -> dict(param=weakref, ptr, out, transposed, epoch)
        self.table = None           # device int64 [n, 8]
        self.table_keys = []
        self.max_elements = 0
        self.retired = []
        self.sources = []
        self.dirty = False          # entries were added since the table was built

    def _rebuild_table(self):
        rows, keys, sources, biggest = [], [], [], 0
        for key, e in list(self.entries.items()):
            p = e["param"]()
            if p is None:
                del self.entries[key]                       # a dead model's weight: its packed copy goes with it
                continue
            src = p.detach()
            rows.append(ops.pack_job(src, e["transposed"], e["out"], e["tap_major"]))
            keys.append(key)
            sources.append(src)
            e["table_ptr"] = src.data_ptr()
            biggest = max(biggest, p.numel())
        if self.table is not None:
            # a captured CUDA graph may still read the old table AND the weights it points at: both stay alive
            self.retired.append((self.table, self.sources))
            del self.retired[:-8]
        self.sources = sources                              # the rows hold raw pointers: pin the storages they name
        if not rows:
            self.table, self.table_keys, self.max_elements, self.dirty = None, [], 0, False
            return
        dev = sources[0].device
        self.table = torch.tensor(rows, dtype=torch.int64).to(dev)
        self.table_keys, self.max_elements = keys, biggest
        self.dirty = False

    def _table_is_current(self):
        """False when a registered weight died or its storage moved since the table was built (the rows are raw pointers)."""
        for k in self.table_keys:
            e = self.entries.get(k)
            p = e["param"]() if e is not None else None
            if p is None or p.data_ptr() != e.get("table_ptr"):
                return False
        return True

    def get(self, param, transposed, tag, tap_major=None):
        import weakref
        epoch = _WEIGHTS_EPOCH[0]
        key = (id(param), tag)
        e = self.entries.get(key)
        if e is not None and (e["param"]() is not param or e["ptr"] != param.data_ptr()):
            e = None                                        # a dead parameter's id, or the storage moved
        if e is None:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("a conv weight was packed for the first time inside a CUDA-graph capture: run the "
                                   "step eagerly once before capturing")
            out = ops._pack_kernel(param, transposed, tap_major)
            self.entries[key] = {"param": weakref.ref(param), "ptr": param.data_ptr(), "out": out,
                                 "transposed": transposed, "tap_major": tap_major, "epoch": epoch,
                                 "version": param._version}
            self.dirty = True
            return out
        if e["epoch"] != epoch or e["version"] != param._version:
            capturing = torch.cuda.is_current_stream_capturing()
            if self.table is None or self.dirty or (not capturing and not self._table_is_current()):
                if capturing:
                    raise RuntimeError("the weight-packing table must be built before a CUDA-graph capture "
                                       "(fastpath.prepare_packing())")
                self._rebuild_table()
            # refresh EVERY registered packing at once (the optimizers stepped all networks)
            ops.pack_conv_weights_batched(self.table, len(self.table_keys), self.max_elements)
            for k in self.table_keys:
                ent = self.entries[k]
                p = ent["param"]()
                ent["epoch"] = epoch
                if p is not None:
                    ent["version"] = p._version
        return e["out"]


_REGISTRY = _PackRegistry()


def prepare_packing():
    """Builds the device job table of the batched weight packing now (host -> device copy): call before capturing a
    CUDA graph of a step, after at least one eager step has registered the networks' weights."""
    if _REGISTRY.entries and (_REGISTRY.table is None or _REGISTRY.dirty or not _REGISTRY._table_is_current()):
        _REGISTRY._rebuild_table()


def _packed(param, fn, tag='fwd'):
    """Packed bf16 copy of a weight, rebuilt when the parameter is modified in place, its storage is replaced, or any
    optimizer has stepped since.  `tag` distinguishes the packings of one parameter (forward, input-gradient, per-tap
    transposed-conv slices).  Plain Conv2d packings (forward / input-gradient) of fp32 contiguous weights go through the
    batched registry above; derived ones (transposed-conv slices) keep a per-parameter cache: it lives ON the parameter
    object -- a global table keyed by id() alone would hand a new parameter the packed weights of a dead one."""
    if (fn is ops.pack_conv_weight or fn is ops.pack_conv_weight_dgrad or fn is ops.pack_conv_weight_s2) and param.is_cuda \
            and param.dtype == torch.float32 and param.is_contiguous():
        if fn is ops.pack_conv_weight_s2 and tag == 'fwd':
            tag = 'fwd_s2'
        return _REGISTRY.get(param, fn is ops.pack_conv_weight_dgrad, tag,
                             tap_major=True if fn is ops.pack_conv_weight_s2 else None)
    cache = param.__dict__.get('_ctl_packed')
    if cache is None:
        cache = param.__dict__['_ctl_packed'] = {}
    ver = (param.data_ptr(), param._version, _WEIGHTS_EPOCH[0])
    hit = cache.get(tag)
    if hit is None or hit[0] != ver:
        hit = (ver, fn(param))
        cache[tag] = hit
    return hit[1]


def _pack_stem_weight(weight):
    return ops.pack_conv_weight(ops.pad_stem_weight(weight))


def _act_code(module):
    if module is None:
        return ops.ACT_NONE
    if isinstance(module, nn.LeakyReLU):
        return ops.ACT_LRELU
    if isinstance(module, nn.ReLU):
        return ops.ACT_RELU
    if isinstance(module, nn.Sigmoid):
        return ops.ACT_SIGMOID
    raise NotImplementedError("activation %r has no fused epilogue" % (module,))


_FROZEN_AFFINES = [None]       # dict (id(conv), id(bn)) -> (scale, shift) while a frozen_eval_affines() context is active


class frozen_eval_affines:
    """Inference on FIXED weights (inference.GraphedPredictor): inside the context the folded eval-mode BatchNorm affine
    of every (conv, BatchNorm) pair is computed once and reused, instead of ~5 tiny launches per pair and forward.  The
    dict is returned by __enter__ so the caller can keep it alive / reuse it (`frozen_eval_affines(cache)`)."""

    def __init__(self, cache=None):
        self.cache = {} if cache is None else cache

    def __enter__(self):
        self.prev, _FROZEN_AFFINES[0] = _FROZEN_AFFINES[0], self.cache
        return self.cache

    def __exit__(self, *exc):
        _FROZEN_AFFINES[0] = self.prev
        return False


def _fold_eval(conv, bn):
    """conv bias + eval-mode BN -> per-channel (scale, shift)."""
    cache = _FROZEN_AFFINES[0]
    key = (id(conv), id(bn))
    if cache is not None and key in cache:
        return cache[key]
    scale = bn.weight.detach() * torch.rsqrt(bn.running_var + bn.eps)
    shift = bn.bias.detach() + (conv.bias.detach() - bn.running_mean) * scale
    if cache is not None:
        cache[key] = (scale, shift)
    return scale, shift


def _bn_batch(bn, y_raw, mode):
    track = mode == 'track' and bn.track_running_stats
    scale, shift = ops.bn_batch_affine_c8(y_raw, bn.weight, bn.bias, bn.eps,
                                          bn.running_mean if track else None, bn.running_var if track else None,
                                          bn.momentum if bn.momentum is not None else 0.1)
    if track:
        bn.num_batches_tracked += 1
    return scale, shift


def _conv(conv, x, **kw):
    k = conv.kernel_size[0]
    sub = conv.stride[0]
    wp = _packed(conv.weight, ops.pack_conv_weight_s2 if (sub == 2 and k == 3) else ops.pack_conv_weight)
    return ops.conv2d_c8(x, wp, conv.out_channels, k * k, subsample=sub, **kw)


def conv_bn_act(conv, bn, x, act, mode):
    """K3 conv + BatchNorm + activation on a C8 tensor."""
    if mode == 'eval':
        scale, shift = _fold_eval(conv, bn)
        return _conv(conv, x, scale=scale, shift=shift, act=act)
    y = _conv(conv, x, shift=conv.bias)
    scale, shift = _bn_batch(bn, y, mode)
    return ops.scale_shift_act_c8(y, scale, shift, act, inplace=True)


def double_conv(seq, x, mode, final_act=ops.ACT_NONE):
    y = conv_bn_act(seq[0], seq[1], x, ops.ACT_LRELU, mode)
    return conv_bn_act(seq[3], seq[4], y, final_act, mode)


def residual_block(block, xr, mode, x_low=None):
    """xr: resampled input (C8).  out = LReLU(conv_input(xr) + BN2(conv2(LReLU(BN1(conv1(xr)))))).
    x_low: the input before a nearest x2 up-sampling -- the 1x1 shortcut is then taken at the low resolution
    (conv1x1(up(x)) == up(conv1x1(x))) and added, up-sampled on the fly, together with BN2 + LReLU."""
    seq = block.conv
    ci = block.conv_input
    y1 = conv_bn_act(seq[0], seq[1], xr, ops.ACT_LRELU, mode)
    if mode == 'eval':
        s2, t2 = _fold_eval(seq[3], seq[4])
        y2 = _conv(seq[3], y1, scale=s2, shift=t2)
        if x_low is not None:
            return ops.scale_shift_upadd_act_c8(y2, None, None, _conv(ci, x_low, shift=ci.bias), ops.ACT_LRELU)
        return _conv(ci, xr, shift=ci.bias, res=y2, act=ops.ACT_LRELU)
    y2 = _conv(seq[3], y1, shift=seq[3].bias)
    s2, t2 = _bn_batch(seq[4], y2, mode)
    if x_low is not None:
        return ops.scale_shift_upadd_act_c8(y2, s2, t2, _conv(ci, x_low, shift=ci.bias), ops.ACT_LRELU)
    return _conv(ci, xr, shift=ci.bias, res=y2, res_scale=s2, res_shift=t2, act=ops.ACT_LRELU)


def down_block(block, x, mode):
    xd = _conv(block.down, x, shift=block.down.bias)            # 3x3 stride 2: subsample inferred from conv.stride
    return residual_block(block, xd, mode)


def up_block(block, x, mode):
    if block.up_type == 'NN':
        return residual_block(block, ops.upsample2x_c8(x), mode, x_low=x)
    up = block.up
    wp = _packed(up.weight, ops.pack_convtranspose2x2_weight)
    xu = ops.conv2d_c8(x, wp, 4 * up.out_channels, 1, up2x=True, shift=up.bias.detach().repeat(4))
    return residual_block(block, xu, mode)


def encoder_forward(enc, x, mode, in_mode=0, temperature=1.0):
    """MyEncoder on planar input (fp32 image / logits, or an int64 label map with in_mode=2) -> C8 latent."""
    inc = enc.inc
    # stem on the tensor core: 16-channel C8 input of bf16 (hi | lo | hi) groups, weight (w_hi | w_hi | w_lo | 0)
    xin = ops.stem_input_c8(x, inc[0].in_channels, in_mode, temperature)
    wp = _packed(inc[0].weight, _pack_stem_weight, tag='stem')
    if mode == 'eval':
        s0, t0 = _fold_eval(inc[0], inc[1])
        y = ops.conv2d_c8(xin, wp, inc[0].out_channels, 9, scale=s0, shift=t0, act=ops.ACT_LRELU)
    else:
        y = ops.conv2d_c8(xin, wp, inc[0].out_channels, 9, shift=inc[0].bias)
        s0, t0 = _bn_batch(inc[1], y, mode)
        y = ops.scale_shift_act_c8(y, s0, t0, ops.ACT_LRELU, inplace=True)
    y = conv_bn_act(inc[3], inc[4], y, ops.ACT_LRELU, mode)     # BN then F.leaky_relu (encoder_decoder.py:405)
    for blk in (enc.down1, enc.down2, enc.down3, enc.down4):
        y = down_block(blk, y, mode)
    return conv_bn_act(enc.final_conv[0], enc.final_conv[1], y, _act_code(enc.act), mode)


def filter_code(dual, z_c8, mode):
    seq = dual.code_decoupler
    return double_conv(seq, z_c8, mode, final_act=_act_code(seq[5]))


def decoder_forward(dec, z_c8, mode):
    """MyDecoder on a C8 latent -> planar fp32 [N,Cout,H,W] (logits, or the sigmoid image)."""
    y = z_c8
    for blk in (dec.up1, dec.up2, dec.up3, dec.up4):
        y = up_block(blk, y, mode)
    return ops.head_conv_c8(y, dec.final_conv.weight, dec.final_conv.bias, _act_code(dec.last_act))


def decoder_from_nchw(dec, z, mode):
    with torch.no_grad():
        return decoder_forward(dec, ops.nchw_to_c8(z), mode)


def ftn_forward(dual, seg_dec, image, mode):
    """Dual_Branch_Encoder + segmentation decoder: image -> (z_i, z_s) as fp32 NCHW, logits."""
    with torch.no_grad():
        z_i = encoder_forward(dual.general_encoder, image, mode)
        z_s = filter_code(dual, z_i, mode)
        logits = decoder_forward(seg_dec, z_s, mode)
        return ops.c8_to_nchw(z_i), ops.c8_to_nchw(z_s), logits


def stn_forward(shape_enc, shape_dec, seg, mode, is_label_map=False, temperature=2.0):
    """recon_shape: construct_input fused into the stem (softmax(logit/T) or one-hot) -> Es -> Dsh -> logits."""
    with torch.no_grad():
        z = encoder_forward(shape_enc, seg, mode, in_mode=2 if is_label_map else 1, temperature=temperature)
        return decoder_forward(shape_dec, z, mode)
