"""Tensor-level wrappers over the C ABI.  torch is used for device memory and streams only:
every function here hands raw device pointers + the current CUDA stream to libctl_b200.so."""
import numpy as np
import torch

from . import _lib
from ._lib import MODE_CHANNEL, MODE_SPATIAL  # noqa: F401  (re-exported)

_DTYPES = {torch.float32: _lib.CTL_F32, torch.bfloat16: _lib.CTL_BF16}


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.CtlError("ctl_b200 ops run on CUDA tensors only (got a %s tensor); there is no CPU fallback"
                                % t.device)


def _dtype(t):
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise TypeError("unsupported dtype %s (float32 / bfloat16 only)" % t.dtype) from None


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _nchw(t):
    if t.dim() != 4:
        raise ValueError("expected a 4-D NCHW tensor, got shape %s" % (tuple(t.shape),))
    N, C, H, W = t.shape
    return N, C, H * W


class NativeRNG:
    """Counter-based generator state of the native (Philox) random mode: `seed` fixes the stream,
    `offset` advances by one per draw call, `first_sample` is the global index of this rank's first
    sample so that a batch shard draws exactly what the full batch would."""

    def __init__(self, seed=0, first_sample=0):
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.offset = 0
        self.first_sample = int(first_sample)

    def next_offset(self):
        o = self.offset
        self.offset += 1
        return o


def saliency_reduce(g, mode):
    """K1: fp32 [N,n] mean of g over space (channel mode) or channels (spatial mode)."""
    _need_cuda(g)
    g = g.contiguous()
    N, C, HW = _nchw(g)
    s = torch.empty((N, C if mode == MODE_CHANNEL else HW), device=g.device, dtype=torch.float32)
    with torch.cuda.device(g.device):
        _lib.check(_lib.load().ctl_saliency_reduce(g.data_ptr(), _dtype(g), N, C, HW, mode, s.data_ptr(), _stream()))
    return s


def topp_mask_apply(s, z, mode, k, soft=False, rand=None, rng=None, out_dtype=torch.float32, want_thr=False):
    """K2 on a given saliency s.  Returns (z_masked, mask[N,n], thr or None)."""
    _need_cuda(s, z, rand)
    z = z.contiguous()
    s = s.contiguous()
    N, C, HW = _nchw(z)
    n = C if mode == MODE_CHANNEL else HW
    if tuple(s.shape) != (N, n) or s.dtype != torch.float32:
        raise ValueError("s must be float32 [%d,%d]" % (N, n))
    if rand is not None:
        rand = rand.contiguous()
        if tuple(rand.shape) != (N, n) or rand.dtype != torch.float32:
            raise ValueError("rand must be float32 [%d,%d]" % (N, n))
    seed = offset = first = 0
    if soft and rand is None:
        if rng is None:
            raise ValueError("soft masking needs either `rand` (torch-compatible mode) or `rng` (native mode)")
        seed, offset, first = rng.seed, rng.next_offset(), rng.first_sample
    z_out = torch.empty(z.shape, device=z.device, dtype=out_dtype)
    mask = torch.empty((N, n), device=z.device, dtype=torch.float32)
    thr = torch.empty((N,), device=z.device, dtype=torch.float32) if want_thr else None
    with torch.cuda.device(z.device):
        _lib.check(_lib.load().ctl_topp_mask_apply(
            s.data_ptr(), z.data_ptr(), _dtype(z), N, C, HW, mode, int(k), int(bool(soft)), _ptr(rand), seed, offset,
            first, mask.data_ptr(), _ptr(thr), z_out.data_ptr(), _DTYPES[out_dtype], _stream()))
    return z_out, mask, thr


def saliency_mask_apply(g, z, mode, k, soft=False, rand=None, rng=None, out_dtype=torch.float32, want_thr=False):
    """K1+K2 fused entry point.  Returns (z_masked, mask[N,n], s[N,n], thr or None)."""
    _need_cuda(g, z, rand)
    g = g.contiguous()
    z = z.contiguous()
    N, C, HW = _nchw(z)
    if tuple(g.shape) != tuple(z.shape):
        raise ValueError("g and z must have the same shape")
    n = C if mode == MODE_CHANNEL else HW
    if rand is not None:
        rand = rand.contiguous()
        if tuple(rand.shape) != (N, n) or rand.dtype != torch.float32:
            raise ValueError("rand must be float32 [%d,%d]" % (N, n))
    seed = offset = first = 0
    if soft and rand is None:
        if rng is None:
            raise ValueError("soft masking needs either `rand` (torch-compatible mode) or `rng` (native mode)")
        seed, offset, first = rng.seed, rng.next_offset(), rng.first_sample
    s = torch.empty((N, n), device=z.device, dtype=torch.float32)
    z_out = torch.empty(z.shape, device=z.device, dtype=out_dtype)
    mask = torch.empty((N, n), device=z.device, dtype=torch.float32)
    thr = torch.empty((N,), device=z.device, dtype=torch.float32) if want_thr else None
    with torch.cuda.device(z.device):
        _lib.check(_lib.load().ctl_saliency_mask_apply(
            g.data_ptr(), _dtype(g), z.data_ptr(), _dtype(z), N, C, HW, mode, int(k), int(bool(soft)), _ptr(rand),
            seed, offset, first, s.data_ptr(), mask.data_ptr(), _ptr(thr), z_out.data_ptr(), _DTYPES[out_dtype],
            _stream()))
    return z_out, mask, s, thr


def dropout_scale(p, dtype=torch.float32):
    """fp32(1/(1-p)) exactly as ATen's feature_dropout computes it (noise.div_(1 - p))."""
    if p >= 1.0:
        return 0.0
    return float(np.float32(1.0) / np.float32(1.0 - p))


def channel_dropout(z, p, keep=None, rng=None, want_mask=True, want_keep=False, out_dtype=None):
    """Random channel dropout + the reference's full-size `masked == z` mask.
    keep: float32 [N,C] of 0/1 (torch-compatible mode) or None with rng (native Philox mode)."""
    _need_cuda(z, keep)
    if p < 0.0 or p > 1.0:
        raise ValueError("dropout probability has to be between 0 and 1, but got {}".format(p))
    z = z.contiguous()
    N, C, HW = _nchw(z)
    out_dtype = out_dtype or z.dtype
    seed = offset = first = 0
    if keep is None:
        if rng is None:
            raise ValueError("channel_dropout needs either `keep` (torch-compatible mode) or `rng` (native mode)")
        seed, offset, first = rng.seed, rng.next_offset(), rng.first_sample
    else:
        keep = keep.reshape(N, C).to(torch.float32).contiguous()
    z_out = torch.empty(z.shape, device=z.device, dtype=out_dtype)
    mask = torch.empty(z.shape, device=z.device, dtype=torch.float32) if want_mask else None
    keep_out = torch.empty((N, C), device=z.device, dtype=torch.float32) if want_keep else None
    with torch.cuda.device(z.device):
        _lib.check(_lib.load().ctl_channel_dropout(
            z.data_ptr(), _dtype(z), N, C, HW, float(p), dropout_scale(p), _ptr(keep), seed, offset, first,
            z_out.data_ptr(), _DTYPES[out_dtype], _ptr(mask), _ptr(keep_out), _stream()))
    return z_out, mask, keep_out


def philox_uniform(seed, offset, first_index, count, device="cuda"):
    out = torch.empty((count,), device=device, dtype=torch.float32)
    with torch.cuda.device(out.device):
        _lib.check(_lib.load().ctl_philox_uniform(int(seed), int(offset), int(first_index), count, out.data_ptr(),
                                                  _stream()))
    return out


# ------------------------------------------------------------------------------------------------ K3: conv blocks
ACT_NONE, ACT_LRELU, ACT_RELU, ACT_SIGMOID = _lib.ACT_NONE, _lib.ACT_LRELU, _lib.ACT_RELU, _lib.ACT_SIGMOID


def conv_supported(cin, cout, kernel_size):
    """True when ctl_conv2d_nhwc_bf16 has a tensor-core kernel for this layer class."""
    taps = kernel_size * kernel_size
    return taps in (1, 9) and _lib.load().ctl_conv2d_n_tile(int(cin), int(cout), taps) > 0


def pack_conv_weight(weight):
    """[Cout,Cin,k,k] (k = 1 or 3) -> bf16 [Cout/NT][taps][Cin/8][NT][8], the K-major core-matrix order the
    UMMA descriptors of conv_tc.cu expect (include/ctl_b200.h).  Tiny tensors: plain torch ops."""
    cout, cin, kh, kw = weight.shape
    taps = kh * kw
    nt = _lib.load().ctl_conv2d_n_tile(cin, cout, taps)
    if kh != kw or nt <= 0:
        raise NotImplementedError("no tcgen05 conv kernel for weight shape %s" % (tuple(weight.shape),))
    w = weight.detach().to(torch.bfloat16).permute(2, 3, 1, 0).reshape(taps, cin // 8, 8, cout // nt, nt)
    return w.permute(3, 0, 1, 4, 2).contiguous()


def _channels_last_bf16(x):
    if x.dtype != torch.bfloat16:
        x = x.to(torch.bfloat16)
    return x.contiguous(memory_format=torch.channels_last)


def conv2d_bf16(x, w_packed, cout, taps, subsample=1, scale=None, shift=None, res=None, res_scale=None, res_shift=None,
                act=ACT_NONE):
    """out = act(conv(x) * scale + shift + res * res_scale + res_shift) on the tcgen05 kernel.
    x / res / out: logical [N,C,H,W] bf16 tensors in channels_last memory format (NHWC in HBM)."""
    _need_cuda(x, w_packed, scale, shift, res, res_scale, res_shift)
    x = _channels_last_bf16(x)
    N, cin, H, W = x.shape
    Ho, Wo = H // subsample, W // subsample
    out = torch.empty((N, cout, Ho, Wo), device=x.device, dtype=torch.bfloat16, memory_format=torch.channels_last)
    if res is not None:
        res = _channels_last_bf16(res)
        if tuple(res.shape) != tuple(out.shape):
            raise ValueError("res must have the output shape %s" % (tuple(out.shape),))
    vecs = []
    for v in (scale, shift, res_scale, res_shift):
        if v is not None:
            v = v.to(torch.float32).contiguous()
            if v.numel() != cout:
                raise ValueError("per-channel vectors must have Cout=%d elements" % cout)
        vecs.append(v)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().ctl_conv2d_nhwc_bf16(
            x.data_ptr(), N, H, W, cin, w_packed.data_ptr(), cout, taps, subsample, _ptr(vecs[0]), _ptr(vecs[1]),
            _ptr(res), _ptr(vecs[2]), _ptr(vecs[3]), act, out.data_ptr(), _stream()))
    return out
