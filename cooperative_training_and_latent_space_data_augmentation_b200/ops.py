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


class StepParams:
    """Per-step scalar draws kept in DEVICE memory so that a CUDA graph of the cooperative step can be replayed with
    new values (include/ctl_b200.h: ctl_saliency_mask_apply_dyn / ctl_channel_dropout_dyn).  One row of three int64
    per masking call of the step: {k, philox offset, global index of the first local sample}.

    While `recording` (the capture pass) every masking call takes the next row, notes what it will need at replay
    time and launches the *_dyn kernel on that row.  At replay the owner refills the rows on the host (`fill`) and
    `upload()` enqueues ONE small pinned H2D copy in front of the graph launch; a ring of pinned buffers guarded by
    events keeps a copy that has not executed yet from being overwritten by the next step's draws."""

    ROWS, RING = 8, 4

    def __init__(self, device):
        self.dev = torch.zeros((self.ROWS, 3), dtype=torch.int64, device=device)
        self.host = [torch.zeros((self.ROWS, 3), dtype=torch.int64).pin_memory() for _ in range(self.RING)]
        self.events = [None] * self.RING
        self.turn = 0
        self.rows = []              # per row: {"kind": "mask"|"dropout", "n": int, "draws": bool}
        self.recording = False
        self._stage = self.host[0]

    def take(self, kind, n, k, rng):
        """Capture pass: claims the next row for a call that runs now with (k, rng state) and returns its device
        view.  `rng` is the NativeRNG the call draws from, or None when it consumes no Philox offset."""
        if len(self.rows) >= self.ROWS:
            raise RuntimeError("more than %d masking calls in one captured step" % self.ROWS)
        i = len(self.rows)
        self.rows.append({"kind": kind, "n": int(n), "draws": rng is not None})
        self.fill(i, k, rng)
        if not torch.cuda.is_current_stream_capturing():
            self.dev[i].copy_(self._stage[i], non_blocking=True)     # eager use: the launch below really runs
        return self.dev[i]

    def begin(self):
        """Replay pass: picks the pinned staging buffer of this step (waits if its last copy is still pending)."""
        ev = self.events[self.turn]
        if ev is not None:
            ev.synchronize()
        self._stage = self.host[self.turn]

    def fill(self, i, k, rng):
        row = self._stage[i]
        row[0] = int(k)
        row[1] = rng.next_offset() if rng is not None else 0
        row[2] = rng.first_sample if rng is not None else 0

    def upload(self):
        self.dev.copy_(self._stage, non_blocking=True)
        ev = self.events[self.turn] or torch.cuda.Event()
        ev.record()
        self.events[self.turn] = ev
        self.turn = (self.turn + 1) % self.RING


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


def saliency_mask_apply(g, z, mode, k, soft=False, rand=None, rng=None, out_dtype=torch.float32, want_thr=False,
                        step_params=None):
    """K1+K2 fused entry point.  Returns (z_masked, mask[N,n], s[N,n], thr or None).
    step_params: a recording StepParams -> the launch reads (k, offset, first_sample) from device memory."""
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
    native = soft and rand is None
    if native and rng is None:
        raise ValueError("soft masking needs either `rand` (torch-compatible mode) or `rng` (native mode)")
    s = torch.empty((N, n), device=z.device, dtype=torch.float32)
    z_out = torch.empty(z.shape, device=z.device, dtype=out_dtype)
    mask = torch.empty((N, n), device=z.device, dtype=torch.float32)
    thr = torch.empty((N,), device=z.device, dtype=torch.float32) if want_thr else None
    if step_params is not None:
        if not 0 <= int(k) < n:
            raise IndexError("index {} is out of bounds for dimension 1 with size {}".format(int(k), n))
        row = step_params.take("mask", n, k, rng if native else None)
        with torch.cuda.device(z.device):
            _lib.check(_lib.load().ctl_saliency_mask_apply_dyn(
                g.data_ptr(), _dtype(g), z.data_ptr(), _dtype(z), N, C, HW, mode, int(bool(soft)), _ptr(rand),
                rng.seed if native else 0, row.data_ptr(), s.data_ptr(), mask.data_ptr(), _ptr(thr), z_out.data_ptr(),
                _DTYPES[out_dtype], _stream()))
        return z_out, mask, s, thr
    if native:
        seed, offset, first = rng.seed, rng.next_offset(), rng.first_sample
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


def channel_dropout(z, p, keep=None, rng=None, want_mask=True, want_keep=False, out_dtype=None, step_params=None):
    """Random channel dropout + the reference's full-size `masked == z` mask.
    keep: float32 [N,C] of 0/1 (torch-compatible mode) or None with rng (native Philox mode)."""
    _need_cuda(z, keep)
    if p < 0.0 or p > 1.0:
        raise ValueError("dropout probability has to be between 0 and 1, but got {}".format(p))
    z = z.contiguous()
    N, C, HW = _nchw(z)
    out_dtype = out_dtype or z.dtype
    seed = offset = first = 0
    if keep is None and rng is None:
        raise ValueError("channel_dropout needs either `keep` (torch-compatible mode) or `rng` (native mode)")
    if keep is not None:
        keep = keep.reshape(N, C).to(torch.float32).contiguous()
    z_out = torch.empty(z.shape, device=z.device, dtype=out_dtype)
    mask = torch.empty(z.shape, device=z.device, dtype=torch.float32) if want_mask else None
    keep_out = torch.empty((N, C), device=z.device, dtype=torch.float32) if want_keep else None
    if keep is None and step_params is not None:
        row = step_params.take("dropout", C, 0, rng)
        with torch.cuda.device(z.device):
            _lib.check(_lib.load().ctl_channel_dropout_dyn(
                z.data_ptr(), _dtype(z), N, C, HW, float(p), dropout_scale(p), rng.seed, row.data_ptr(),
                z_out.data_ptr(), _DTYPES[out_dtype], _ptr(mask), _ptr(keep_out), _stream()))
        return z_out, mask, keep_out
    if keep is None:
        seed, offset, first = rng.seed, rng.next_offset(), rng.first_sample
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

# Blocked activation layout "C8": a bf16 tensor of shape [N, C/8, H, W, 8] (contiguous).  See DESIGN.md section 3.


def _c8_dims(x):
    if x.dim() != 5 or x.shape[-1] != 8 or x.dtype != torch.bfloat16 or not x.is_contiguous():
        raise ValueError("expected a contiguous bf16 C8 tensor [N, C/8, H, W, 8], got %s %s" % (tuple(x.shape), x.dtype))
    N, C8, H, W, _ = x.shape
    return N, C8 * 8, H, W


def _vec(v, n):
    if v is None:
        return None
    v = v.detach().to(torch.float32).contiguous()
    if v.numel() != n:
        raise ValueError("per-channel vector must have %d elements, got %d" % (n, v.numel()))
    return v


def conv_supported(cin, cout, kernel_size):
    """True when ctl_conv2d_c8_bf16 has a tensor-core kernel for this layer class."""
    taps = kernel_size * kernel_size
    return taps in (1, 9) and _lib.load().ctl_conv2d_n_tile(int(cin), int(cout), taps) > 0


def _tap_major(cout_p, cin_p, taps, stride=1):
    """Weight layout the kernel of this layer class reads (packed-view channel counts)."""
    return not _lib.load().ctl_conv2d_vpacked(int(cin_p), int(cout_p), int(taps), int(stride))


def pack_conv_weight_torch(weight, tap_major=None):
    """[Cout,Cin,k,k] (k = 1 or 3) -> the packed bf16 weight of conv_tc.cu (include/ctl_b200.h), stated in torch ops
    (used for derived weights -- transposed-conv slices -- and by the tests).  1x1 and tap_major 3x3 (the stride-2
    kernel): [Cout/NT][taps][Cin/8][NT][8]; 3x3 stride 1: [Cout/NT][3 (s)][Cin/8][3*NT (r, n)][8]."""
    cout, cin, kh, kw = weight.shape
    taps = kh * kw
    nt = _lib.load().ctl_conv2d_n_tile(cin, cout, taps)
    if kh != kw or nt <= 0:
        raise NotImplementedError("no tcgen05 conv kernel for weight shape %s" % (tuple(weight.shape),))
    wb = weight.detach().to(torch.bfloat16)
    if tap_major is None:
        tap_major = _tap_major(cout, cin, taps)
    if taps == 1 or tap_major:
        w = wb.permute(2, 3, 1, 0).reshape(taps, cin // 8, 8, cout // nt, nt)
        return w.permute(3, 0, 1, 4, 2).contiguous()
    w = wb.reshape(cout // nt, nt, cin // 8, 8, 3, 3)            # [t][n][q][j][r][s]
    return w.permute(0, 5, 2, 4, 1, 3).contiguous()              # [t][s][q][r][n][j]


def _pack_kernel(weight, transposed, tap_major=None):
    cout, cin, kh, kw = weight.shape
    taps = kh * kw
    co_p, ci_p = (cin, cout) if transposed else (cout, cin)
    if tap_major is None:
        tap_major = _tap_major(co_p, ci_p, taps)
    if kh != kw or taps not in (1, 9) or _lib.load().ctl_conv2d_n_tile(ci_p, co_p, taps) <= 0:
        raise NotImplementedError("no tcgen05 conv kernel for weight shape %s" % (tuple(weight.shape),))
    w = weight.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.to(torch.float32).contiguous()
    _need_cuda(w)
    out = torch.empty(cout * cin * taps, device=w.device, dtype=torch.bfloat16)
    with torch.cuda.device(w.device):
        _lib.check(_lib.load().ctl_pack_conv_weight(w.data_ptr(), cout, cin, taps, int(transposed), int(bool(tap_major)),
                                                    out.data_ptr(), _stream()))
    return out


def pack_job(weight, transposed, out, tap_major=None):
    """One row of the ctl_pack_conv_weights_batched job table for a contiguous fp32 [Cout,Cin,k,k] weight."""
    cout, cin, kh, kw = weight.shape
    taps = kh * kw
    co_p, ci_p = (cin, cout) if transposed else (cout, cin)
    if tap_major is None:
        tap_major = _tap_major(co_p, ci_p, taps)
    nt = _lib.load().ctl_conv2d_n_tile(ci_p, co_p, taps)
    if kh != kw or taps not in (1, 9) or nt <= 0 or weight.dtype != torch.float32 or not weight.is_contiguous():
        raise NotImplementedError("no batched packing for weight %s %s" % (tuple(weight.shape), weight.dtype))
    return [weight.data_ptr(), out.data_ptr(), cout, cin, taps, nt, int(transposed), int(bool(tap_major))]


def pack_conv_weights_batched(table, n_jobs, max_elements):
    """table: DEVICE int64 [n_jobs, 8] built from pack_job rows; packs every listed weight in one launch."""
    _need_cuda(table)
    with torch.cuda.device(table.device):
        _lib.check(_lib.load().ctl_pack_conv_weights_batched(table.data_ptr(), int(n_jobs), int(max_elements), _stream()))


def pack_conv_weight(weight):
    """Packed bf16 forward weight of a stride-1 Conv2d (one kernel launch; see pack_conv_weight_torch for the layout)."""
    return _pack_kernel(weight, False)


def pack_conv_weight_s2(weight):
    """Packed bf16 forward weight of the 3x3 STRIDE-2 Conv2d (`down`): tap-major layout of the unpacked-tap kernel."""
    return _pack_kernel(weight, False, tap_major=True)


def pack_convtranspose2x2_weight(weight):
    """ConvTranspose2d(k=2, s=2) weight [Cin, Cout, 2, 2] -> the 1x1 GEMM weight [4*Cout, Cin, 1, 1] with rows
    ordered (dy*2+dx)*Cout + co, packed for ctl_conv2d_c8_bf16(up2x=1)."""
    cin, cout, kh, kw = weight.shape
    if (kh, kw) != (2, 2):
        raise NotImplementedError("only kernel 2 / stride 2 transposed convolutions are on the hot path")
    w = weight.detach().permute(2, 3, 1, 0).reshape(4 * cout, cin, 1, 1)
    return pack_conv_weight_torch(w)


def conv_stats_fusable(cin, cout, taps):
    """True when the conv epilogue can accumulate the BatchNorm statistics of its own output (N tile <= 32)."""
    return 0 < _lib.load().ctl_conv2d_n_tile(int(cin), int(cout), int(taps)) <= 32


def conv2d_c8(x, w_packed, cout, taps, subsample=1, up2x=False, scale=None, shift=None, res=None, res_scale=None,
              res_shift=None, act=ACT_NONE, stats=None):
    """out = act(conv(x) * scale + shift + res * res_scale + res_shift) on the tcgen05 kernel; C8 in, C8 out.
    `cout` is the number of GEMM columns (4 * out_channels when up2x).  stats: zeroed float64 [2, cout] that receives
    the per-channel sum / sum of squares of the stored outputs (see conv_stats_fusable)."""
    _need_cuda(x, w_packed, scale, shift, res, res_scale, res_shift)
    N, cin, H, W = _c8_dims(x)
    if up2x:
        out = torch.empty((N, cout // 32, 2 * H, 2 * W, 8), device=x.device, dtype=torch.bfloat16)
    else:
        out = torch.empty((N, cout // 8, H // subsample, W // subsample, 8), device=x.device, dtype=torch.bfloat16)
    if res is not None and (tuple(res.shape) != tuple(out.shape) or res.dtype != torch.bfloat16 or not res.is_contiguous()):
        raise ValueError("res must be a contiguous bf16 C8 tensor of the output shape %s" % (tuple(out.shape),))
    sc, sh, rs, rb = _vec(scale, cout), _vec(shift, cout), _vec(res_scale, cout), _vec(res_shift, cout)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().ctl_conv2d_c8_bf16(
            x.data_ptr(), N, H, W, cin, w_packed.data_ptr(), cout, taps, subsample, int(bool(up2x)), _ptr(sc), _ptr(sh),
            _ptr(res), _ptr(rs), _ptr(rb), act, out.data_ptr(), _ptr(stats), _stream()))
    return out


def bn_affine_from_sums(sums, count, gamma, beta, eps, running_mean=None, running_var=None, momentum=0.1):
    """(scale, shift, mean, var) from the [2, C] float64 sums a conv epilogue accumulated; running stats updated in place."""
    _need_cuda(sums, gamma, beta, running_mean, running_var)
    C = sums.shape[1]
    out = torch.empty((4, C), device=sums.device, dtype=torch.float32)
    g, b = _vec(gamma, C), _vec(beta, C)
    with torch.cuda.device(sums.device):
        _lib.check(_lib.load().ctl_bn_affine_from_sums(
            sums.data_ptr(), C, int(count), _ptr(g), _ptr(b), float(eps), out[0].data_ptr(), out[1].data_ptr(),
            out[2].data_ptr(), out[3].data_ptr(), _ptr(running_mean), _ptr(running_var), float(momentum), _stream()))
    return out[0], out[1], out[2], out[3]


def nchw_to_c8(x):
    _need_cuda(x)
    x = x.contiguous()
    N, C, H, W = x.shape
    y = torch.empty((N, C // 8, H, W, 8), device=x.device, dtype=torch.bfloat16)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().ctl_nchw_to_c8(x.data_ptr(), _dtype(x), N, C, H, W, y.data_ptr(), _stream()))
    return y


def c8_to_nchw(x, dtype=torch.float32):
    _need_cuda(x)
    N, C, H, W = _c8_dims(x)
    y = torch.empty((N, C, H, W), device=x.device, dtype=dtype)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().ctl_c8_to_nchw(x.data_ptr(), N, C, H, W, y.data_ptr(), _DTYPES[dtype], _stream()))
    return y


def stem_conv_c8(x, weight, scale=None, shift=None, act=ACT_NONE, in_mode=0, temperature=1.0):
    """3x3 pad-1 stem from planar fp32 [N,Cin,H,W] (in_mode 0/1) or an int64 label map [N,H,W] (in_mode 2)."""
    _need_cuda(x, weight)
    cout, cin = weight.shape[0], weight.shape[1]
    if in_mode == 2:
        lab = x.contiguous()
        if lab.dtype != torch.int64 or lab.dim() != 3:
            raise ValueError("in_mode 2 expects an int64 label map [N,H,W]")
        N, H, W = lab.shape
        xp, lp = 0, lab.data_ptr()
    else:
        xf = x.to(torch.float32).contiguous()
        N, c, H, W = xf.shape
        if c != cin:
            raise ValueError("input has %d channels, weight expects %d" % (c, cin))
        xp, lp = xf.data_ptr(), 0
    w = weight.detach().to(torch.float32).contiguous()
    y = torch.empty((N, cout // 8, H, W, 8), device=weight.device, dtype=torch.bfloat16)
    sc, sh = _vec(scale, cout), _vec(shift, cout)
    with torch.cuda.device(weight.device):
        _lib.check(_lib.load().ctl_stem_conv3x3_c8(xp, lp, in_mode, float(temperature), N, cin, H, W, w.data_ptr(), cout,
                                                   _ptr(sc), _ptr(sh), act, y.data_ptr(), _stream()))
    return y


def stem_input_c8(x, cin, in_mode=0, temperature=1.0):
    """The stem's input as a 16-channel C8 tensor: image (in_mode 0), softmax(x / T) (1) or the one-hot of an int64
    label map (2), each channel as a bf16 (hi | lo | hi) triple in channel groups of `cin` (see pad_stem_weight) -- the
    operand of the tensor-core stem (conv2d_c8 / conv_wgrad_c8 with Cin 16)."""
    _need_cuda(x)
    if in_mode == 2:
        lab = x.contiguous()
        if lab.dtype != torch.int64 or lab.dim() != 3:
            raise ValueError("in_mode 2 expects an int64 label map [N,H,W]")
        N, H, W = lab.shape
        xp, lp = 0, lab.data_ptr()
    else:
        xf = x.to(torch.float32).contiguous()
        N, c, H, W = xf.shape
        if c != cin:
            raise ValueError("input has %d channels, expected %d" % (c, cin))
        xp, lp = xf.data_ptr(), 0
    out = torch.empty((N, 2, H, W, 8), device=x.device, dtype=torch.bfloat16)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().ctl_stem_input_c8(xp, lp, in_mode, float(temperature), N, cin, H, W, out.data_ptr(), _stream()))
    return out


def pad_stem_weight(weight):
    """[16, cin, 3, 3] -> [16, 16, 3, 3] fp32 for the tensor-core stem: input-channel groups (w_hi | w_hi | w_lo | 0) with
    w_hi = bf16(w), w_lo = w - w_hi, matching stem_input_c8's (x_hi | x_lo | x_hi | 0): x*w ~ x_hi*w_hi + x_lo*w_hi +
    x_hi*w_lo keeps ~16 mantissa bits of both operands (the CUDA-core stem it replaces was fp32)."""
    cout, cin = weight.shape[0], weight.shape[1]
    w = weight.detach().to(torch.float32)
    hi = w.to(torch.bfloat16).to(torch.float32)
    out = torch.zeros((cout, 16, 3, 3), device=weight.device, dtype=torch.float32)
    out[:, :cin] = hi
    out[:, cin:2 * cin] = hi
    out[:, 2 * cin:3 * cin] = w - hi
    return out


def stem_weight_grad(dW16, cin):
    """Weight gradient [16, cin, 3, 3] from K3w's 16-channel result: the (x_hi | x_lo) channel groups carry sum dy * x."""
    return dW16[:, :cin] + dW16[:, cin:2 * cin]


def head_conv_c8(x, weight, bias=None, act=ACT_NONE):
    """1x1 head from a 16-channel C8 tensor to planar fp32 [N,Cout,H,W] (Cout <= 4)."""
    _need_cuda(x, weight, bias)
    N, cin, H, W = _c8_dims(x)
    cout = weight.shape[0]
    w = weight.detach().to(torch.float32).reshape(cout, cin).contiguous()
    b = _vec(bias, cout)
    y = torch.empty((N, cout, H, W), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().ctl_head_conv1x1_c8(x.data_ptr(), N, cin, H, W, w.data_ptr(), _ptr(b), cout, act,
                                                   y.data_ptr(), _stream()))
    return y


def upsample2x_c8(x):
    _need_cuda(x)
    N, C, H, W = _c8_dims(x)
    y = torch.empty((N, C // 8, 2 * H, 2 * W, 8), device=x.device, dtype=torch.bfloat16)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().ctl_upsample2x_c8(x.data_ptr(), N, C, H, W, y.data_ptr(), _stream()))
    return y


def bn_batch_affine_c8(x, gamma, beta, eps, running_mean=None, running_var=None, momentum=0.1, want_stats=False):
    """Batch statistics of a C8 tensor folded into (scale, shift); updates the running statistics in place when
    they are given (pass None to reproduce _disable_tracking_bn_stats).  want_stats: also return the batch mean and
    biased variance (what the backward needs)."""
    _need_cuda(x, gamma, beta, running_mean, running_var)
    N, C, H, W = _c8_dims(x)
    ws = torch.empty(_lib.load().ctl_bn_workspace_bytes(N, C), device=x.device, dtype=torch.uint8)
    out = torch.empty((4 if want_stats else 2, C), device=x.device, dtype=torch.float32)
    g, b = _vec(gamma, C), _vec(beta, C)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().ctl_bn_batch_affine_c8(
            x.data_ptr(), N, C, H, W, _ptr(g), _ptr(b), float(eps), ws.data_ptr(), out[0].data_ptr(), out[1].data_ptr(),
            out[2].data_ptr() if want_stats else 0, out[3].data_ptr() if want_stats else 0,
            _ptr(running_mean), _ptr(running_var), float(momentum), _stream()))
    if want_stats:
        return out[0], out[1], out[2], out[3]
    return out[0], out[1]


def scale_shift_act_c8(x, scale, shift, act=ACT_NONE, inplace=False):
    _need_cuda(x, scale, shift)
    N, C, H, W = _c8_dims(x)
    y = x if inplace else torch.empty_like(x)
    sc, sh = _vec(scale, C), _vec(shift, C)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().ctl_scale_shift_act_c8(x.data_ptr(), N, C, H, W, sc.data_ptr(), sh.data_ptr(), act,
                                                      y.data_ptr(), _stream()))
    return y


def scale_shift_upadd_act_c8(x, scale, shift, low, act=ACT_NONE):
    """act(x*scale + shift + nearest_up2(low)): the tail of a nearest-x2 residual up block (shortcut taken at low
    resolution).  x: C8 [N,C/8,H,W,8]; low: C8 [N,C/8,H/2,W/2,8]; scale / shift: [C] or None (1 / 0)."""
    _need_cuda(x, low, scale, shift)
    N, C, H, W = _c8_dims(x)
    if tuple(low.shape) != (N, C // 8, H // 2, W // 2, 8) or H % 2 or W % 2:
        raise ValueError("low must be C8 %s for x %s" % ((N, C // 8, H // 2, W // 2, 8), tuple(x.shape)))
    y = torch.empty_like(x)
    sc = _vec(scale, C) if scale is not None else None
    sh = _vec(shift, C) if shift is not None else None
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().ctl_scale_shift_upadd_act_c8(x.data_ptr(), N, C, H, W, _ptr(sc), _ptr(sh), low.data_ptr(),
                                                            act, y.data_ptr(), _stream()))
    return y


# ------------------------------------------------------------------------------------------------ backward kernels
def pack_conv_weight_dgrad(weight):
    """Packed weights of the INPUT-gradient convolution of a stride-1 (or zero-stuffed stride-2) Conv2d:
    dx = conv(dy, w') with w'[ci][co][r][s] = w[co][ci][k-1-r][k-1-s]."""
    return _pack_kernel(weight, True)


def conv_wgrad_c8(x, dy, taps, out=None, layout='kernel'):
    """K3w (tcgen05, both operands straight from C8): sum_p x[p + tap - pad] (x) dy[p], accumulated into `out` (zeroed
    fp32; allocated when None).  layout 'kernel': [taps][Cin][Cout]; 'conv': nn.Conv2d's [Cout][Cin][k][k];
    ('convT', d): tap d of a ConvTranspose2d weight [Cin][Cout][2][2] (taps must be 1)."""
    _need_cuda(x, dy, out)
    N, cin, H, W = _c8_dims(x)
    N2, cout, H2, W2 = _c8_dims(dy)
    if (N2, H2, W2) != (N, H, W):
        raise ValueError("x %s and dy %s must share N, H, W" % (tuple(x.shape), tuple(dy.shape)))
    offset = 0
    if layout == 'kernel':
        shape, strides = (taps, cin, cout), (1, cout, cin * cout)
    elif layout == 'conv':
        k = 3 if taps == 9 else 1
        shape, strides = (cout, cin, k, k), (cin * taps, taps, 1)
    else:
        shape, strides, offset = (cin, cout, 2, 2), (4, 4 * cout, 1), int(layout[1])
    numel = shape[0] * shape[1] * shape[2] * (shape[3] if len(shape) > 3 else 1)
    if out is None:
        out = torch.zeros(shape, device=x.device, dtype=torch.float32)
    elif out.dtype != torch.float32 or not out.is_contiguous() or out.numel() != numel:
        raise ValueError("out must be a contiguous float32 tensor of %d elements" % numel)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().ctl_conv_wgrad_c8_bf16(x.data_ptr(), dy.data_ptr(), N, H, W, cin, cout, taps,
                                                      out.data_ptr() + 4 * offset, strides[0], strides[1], strides[2],
                                                      _stream()))
    return out


def wgrad_to_conv_weight(dW, k):
    """[taps][Cin][Cout] -> nn.Conv2d weight layout [Cout][Cin][k][k]."""
    taps, cin, cout = dW.shape
    return dW.permute(2, 1, 0).reshape(cout, cin, k, k)


def _reduce_ws(N, C, device):
    return torch.empty(_lib.load().ctl_reduce_workspace_bytes(N, C), device=device, dtype=torch.uint8)


def channel_sum_c8(x):
    _need_cuda(x)
    N, C, H, W = _c8_dims(x)
    out = torch.empty(C, device=x.device, dtype=torch.float32)
    ws = _reduce_ws(N, C, x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().ctl_channel_sums_c8(x.data_ptr(), N, C, H, W, ws.data_ptr(), out.data_ptr(), 0, _stream()))
    return out


def bn_act_bwd_c8(dy, h, a, act, mean, var, eps, gamma, want_dv=False, want_param_grads=True, act_affine=None,
                  totals=None):
    """Backward of h = act(BatchNorm_train(a)).  Returns (da, dgamma, dbeta, dv): dv = dy*act'(h) is materialised only
    when want_dv (h must then be given); dgamma/dbeta are None unless want_param_grads.
    act_affine = (scale, shift) of the forward (h = act(a*scale + shift), LReLU / ReLU): h is then not read -- the sign
    act' needs is recomputed from `a` -- unless dv is wanted."""
    _need_cuda(dy, h, a, mean, var, gamma)
    N, C, H, W = _c8_dims(a)
    if tuple(dy.shape) != tuple(a.shape) or (h is not None and tuple(h.shape) != tuple(a.shape)):
        raise ValueError("dy, h and a must have the same C8 shape")
    _c8_dims(dy)
    lib = _lib.load()
    ws = _reduce_ws(N, C, a.device)
    coef = torch.empty((3, C), device=a.device, dtype=torch.float32)
    pg = torch.empty((2, C), device=a.device, dtype=torch.float32) if want_param_grads else None
    sc = sh = None
    if act_affine is not None and not want_dv and act in (ACT_LRELU, ACT_RELU):
        sc, sh = _vec(act_affine[0], C), _vec(act_affine[1], C)
        h = None
    dv = torch.empty_like(a) if (want_dv and h is not None) else None
    g = _vec(gamma, C)
    da = torch.empty_like(a)
    if totals is not None and C <= 256:
        # zeroed float64 [2, C] scratch from the caller's arena: reduction -> totals -> apply, no finalisation launch
        with torch.cuda.device(a.device):
            _lib.check(lib.ctl_bn_bwd_c8(dy.data_ptr(), _ptr(h), a.data_ptr(), N, C, H, W, act, mean.data_ptr(),
                                         var.data_ptr(), float(eps), _ptr(g), totals.data_ptr(), _ptr(dv), da.data_ptr(),
                                         pg[0].data_ptr() if pg is not None else 0,
                                         pg[1].data_ptr() if pg is not None else 0, _ptr(sc), _ptr(sh), _stream()))
        if want_dv and dv is None:
            dv = dy
        return da, (pg[0] if pg is not None else None), (pg[1] if pg is not None else None), dv
    with torch.cuda.device(a.device):
        _lib.check(lib.ctl_bn_bwd_reduce_c8(dy.data_ptr(), _ptr(h), a.data_ptr(), N, C, H, W, act, mean.data_ptr(),
                                            var.data_ptr(), float(eps), _ptr(g), ws.data_ptr(), _ptr(dv), coef.data_ptr(),
                                            pg[0].data_ptr() if pg is not None else 0,
                                            pg[1].data_ptr() if pg is not None else 0, _ptr(sc), _ptr(sh), _stream()))
        if dv is not None:
            _lib.check(lib.ctl_bn_bwd_apply_c8(dv.data_ptr(), 0, a.data_ptr(), N, C, H, W, act, coef.data_ptr(),
                                               da.data_ptr(), 0, 0, _stream()))
        else:
            _lib.check(lib.ctl_bn_bwd_apply_c8(dy.data_ptr(), _ptr(h), a.data_ptr(), N, C, H, W, act, coef.data_ptr(),
                                               da.data_ptr(), _ptr(sc), _ptr(sh), _stream()))
    if want_dv and dv is None:
        dv = dy
    return da, (pg[0] if pg is not None else None), (pg[1] if pg is not None else None), dv


def act_bwd_c8(dy, h, act):
    _need_cuda(dy, h)
    N, C, H, W = _c8_dims(dy)
    dv = torch.empty_like(dy)
    with torch.cuda.device(dy.device):
        _lib.check(_lib.load().ctl_act_bwd_c8(dy.data_ptr(), h.data_ptr(), N, C, H, W, act, dv.data_ptr(), _stream()))
    return dv


def downsample2x_sum_c8(dy):
    _need_cuda(dy)
    N, C, H2, W2 = _c8_dims(dy)
    dx = torch.empty((N, C // 8, H2 // 2, W2 // 2, 8), device=dy.device, dtype=torch.bfloat16)
    with torch.cuda.device(dy.device):
        _lib.check(_lib.load().ctl_downsample2x_sum_c8(dy.data_ptr(), N, C, H2 // 2, W2 // 2, dx.data_ptr(), _stream()))
    return dx


def zero_stuff2x_c8(x):
    _need_cuda(x)
    N, C, H, W = _c8_dims(x)
    y = torch.empty((N, C // 8, 2 * H, 2 * W, 8), device=x.device, dtype=torch.bfloat16)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().ctl_zero_stuff2x_c8(x.data_ptr(), N, C, H, W, y.data_ptr(), _stream()))
    return y


def split_parity2x2_c8(x):
    """C8 [N,C/8,2H,2W,8] -> [4, N, C/8, H, W, 8]; slice d holds pixels (2i + d//2, 2j + d%2)."""
    _need_cuda(x)
    N, C, H2, W2 = _c8_dims(x)
    y = torch.empty((4, N, C // 8, H2 // 2, W2 // 2, 8), device=x.device, dtype=torch.bfloat16)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().ctl_split_parity2x2_c8(x.data_ptr(), N, C, H2 // 2, W2 // 2, y.data_ptr(), _stream()))
    return y


def head_bwd_c8(dy, y, x, weight, act=ACT_NONE, out=None):
    """Backward of head_conv_c8.  Returns (dx C8, dW [Cout,16,1,1] fp32, db [Cout] fp32); `out`: zeroed fp32 buffer of
    Cout*16 + Cout elements to accumulate the parameter gradients into (allocated when None)."""
    _need_cuda(dy, y, x, weight)
    N, cin, H, W = _c8_dims(x)
    cout = weight.shape[0]
    dy = dy.to(torch.float32).contiguous()
    if tuple(dy.shape) != (N, cout, H, W):
        raise ValueError("dy must be [%d,%d,%d,%d]" % (N, cout, H, W))
    w = weight.detach().to(torch.float32).reshape(cout, cin).contiguous()
    dx = torch.empty_like(x)
    if isinstance(out, tuple):            # (dW, db): two live accumulation targets (e.g. the parameters' .grad tensors)
        gw, gb = out
        if gw.dtype != torch.float32 or gb.dtype != torch.float32 or gw.numel() != cout * cin or gb.numel() != cout \
                or not (gw.is_contiguous() and gb.is_contiguous()):
            raise ValueError("out=(dW, db) must be contiguous float32 tensors of %d and %d elements" % (cout * cin, cout))
    else:
        grads = out if out is not None else torch.zeros(cout * cin + cout, device=x.device, dtype=torch.float32)
        gw, gb = grads[:cout * cin].view(cout, cin, 1, 1), grads[cout * cin:]
    yp = 0
    if act != ACT_NONE:
        y = y.to(torch.float32).contiguous()
        yp = y.data_ptr()
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().ctl_head_bwd_c8(dy.data_ptr(), yp, x.data_ptr(), N, cin, H, W, w.data_ptr(), cout, act,
                                               dx.data_ptr(), gw.data_ptr(), gb.data_ptr(), _stream()))
    return dx, gw, gb


def stem_wgrad_c8(dy, x, cin, in_mode=0, temperature=1.0, out=None):
    """Weight gradient [16,cin,3,3] of stem_conv_c8 (dy: gradient of the raw stem output, C8 with 16 channels),
    accumulated into `out` (zeroed fp32; allocated when None)."""
    _need_cuda(dy, x)
    N, cout, H, W = _c8_dims(dy)
    if in_mode == 2:
        lab = x.contiguous()
        xp, lp = 0, lab.data_ptr()
    else:
        xf = x.detach().to(torch.float32).contiguous()
        xp, lp = xf.data_ptr(), 0
    dW = out if out is not None else torch.zeros((cout, cin, 3, 3), device=dy.device, dtype=torch.float32)
    with torch.cuda.device(dy.device):
        _lib.check(_lib.load().ctl_stem_wgrad_c8(dy.data_ptr(), xp, lp, in_mode, float(temperature), N, cin, H, W,
                                                 dW.data_ptr(), _stream()))
    return dW


def stem_dgrad_c8(dy, x, weight, in_mode=0, temperature=1.0):
    """Input gradient (planar fp32 [N,cin,H,W]) of stem_conv_c8; in_mode 1 chains through softmax(x / temperature)."""
    _need_cuda(dy, x, weight)
    N, cout, H, W = _c8_dims(dy)
    cin = weight.shape[1]
    xf = x.detach().to(torch.float32).contiguous() if x is not None else None
    w = weight.detach().to(torch.float32).contiguous()
    dx = torch.empty((N, cin, H, W), device=dy.device, dtype=torch.float32)
    with torch.cuda.device(dy.device):
        _lib.check(_lib.load().ctl_stem_dgrad_c8(dy.data_ptr(), _ptr(xf), in_mode, float(temperature), N, cin, H, W,
                                                 w.data_ptr(), dx.data_ptr(), _stream()))
    return dx


# ------------------------------------------------------------------------------------------------ fused cross entropy
_CE_WS = {}


def _ce_workspace(device):
    """16 zeroed bytes per (device, stream) (fp64 sum + CTA ticket); ctl_ce2d_fwd leaves them zeroed, so one buffer
    serves every call of that stream -- also inside a captured CUDA graph (persistent address).  Per stream because two
    streams running the kernel concurrently must not share the accumulator."""
    key = (device.type, device.index, _stream())
    ws = _CE_WS.get(key)
    if ws is None:
        ws = _CE_WS[key] = torch.zeros(2, device=device, dtype=torch.float64)
    return ws


def ce2d_supported(logits, target):
    """Shapes / dtypes the fused kernels take: CUDA fp32 contiguous [N,C,H,W] logits with C in {2,3,4,8}, int64
    contiguous [N,H,W] label map."""
    return (logits.is_cuda and logits.dtype == torch.float32 and logits.dim() == 4 and logits.shape[1] in (2, 3, 4, 8)
            and target.is_cuda and target.dtype == torch.int64 and target.dim() == 3
            and tuple(target.shape) == (logits.shape[0], logits.shape[2], logits.shape[3]))


class _CrossEntropy2D(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, scale):
        x = logits.detach().contiguous()
        t = target.contiguous()
        N, C, H, W = x.shape
        out = torch.empty(1, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().ctl_ce2d_fwd(x.data_ptr(), t.data_ptr(), N, C, H, W, float(scale),
                                                _ce_workspace(x.device).data_ptr(), out.data_ptr(), _stream()))
        ctx.save_for_backward(x, t)
        ctx.scale = float(scale)
        return out.view(())

    @staticmethod
    def backward(ctx, gout):
        x, t = ctx.saved_tensors
        N, C, H, W = x.shape
        g = gout.detach().to(torch.float32).contiguous()
        dx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().ctl_ce2d_bwd(x.data_ptr(), t.data_ptr(), N, C, H, W, ctx.scale, g.data_ptr(),
                                                dx.data_ptr(), _stream()))
        return dx, None, None


def cross_entropy_2d(logits, target, scale=1.0):
    """scale * sum over pixels of -log softmax(logits)[target]  (0-d fp32 tensor, differentiable w.r.t. logits)."""
    if not ce2d_supported(logits, target):
        raise ValueError("cross_entropy_2d: unsupported logits %s %s / target %s %s"
                         % (tuple(logits.shape), logits.dtype, tuple(target.shape), target.dtype))
    return _CrossEntropy2D.apply(logits, target, scale)


# ------------------------------------------------------------------------------------------------ fused squared error
class _SquaredError(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, scale):
        x = pred.detach().contiguous()
        t = target.detach().contiguous()
        out = torch.empty(1, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().ctl_sse_fwd(x.data_ptr(), t.data_ptr(), x.numel(), float(scale),
                                               _ce_workspace(x.device).data_ptr(), out.data_ptr(), _stream()))
        ctx.save_for_backward(x, t)
        ctx.scale = float(scale)
        return out.view(())

    @staticmethod
    def backward(ctx, gout):
        x, t = ctx.saved_tensors
        g = gout.detach().to(torch.float32).contiguous()
        dx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().ctl_sse_bwd(x.data_ptr(), t.data_ptr(), x.numel(), ctx.scale, g.data_ptr(),
                                               dx.data_ptr(), _stream()))
        return dx, None, None


def sse_supported(pred, target):
    return (pred.is_cuda and target.is_cuda and pred.dtype == torch.float32 and target.dtype == torch.float32
            and tuple(pred.shape) == tuple(target.shape) and pred.numel() > 0 and not target.requires_grad)


def squared_error(pred, target, scale):
    """scale * sum((pred - target)**2) as a 0-d fp32 tensor, differentiable w.r.t. pred (one kernel each way):
    mean squared error with scale = 1/numel, the reference's 0.5 * MSELoss with scale = 0.5/numel."""
    if not sse_supported(pred, target):
        raise ValueError("squared_error: CUDA fp32 tensors of one shape expected, got %s %s / %s %s"
                         % (tuple(pred.shape), pred.dtype, tuple(target.shape), target.dtype))
    return _SquaredError.apply(pred, target, scale)


# ------------------------------------------------------------------------------------------------ multi-tensor Adam
def adam_flat(params, grads, exp_avg, exp_avg_sq, bounds, steps, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
              grad_scale=1.0, zero_grad=False, seg_mask=None):
    """One Adam step over flat fp32 buffers (ctl_adam_flat).  bounds: [(begin, end), ...] one element range per
    optimizer segment; steps: device fp32 [len(bounds)] step counters (incremented by the call)."""
    _need_cuda(params, grads, exp_avg, exp_avg_sq, steps)
    import ctypes
    n = len(bounds)
    arr = (ctypes.c_int64 * (2 * n))(*[int(v) for b in bounds for v in b])
    mask = (1 << n) - 1 if seg_mask is None else int(seg_mask)
    with torch.cuda.device(params.device):
        _lib.check(_lib.load().ctl_adam_flat(params.data_ptr(), grads.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(),
                                             arr, n, mask, steps.data_ptr(), float(lr), float(betas[0]), float(betas[1]),
                                             float(eps), float(weight_decay), float(grad_scale), int(bool(zero_grad)),
                                             _stream()))


# ------------------------------------------------------------------------------------------------ evaluation metric
def confusion_update(hist, gt, logits=None, pred_labels=None, want_labels=False):
    """hist (uint64-as-int64 [C,C], row = gt, column = prediction) += confusion matrix of argmax(logits) (or of
    pred_labels) against gt; returns the uint8 label map [N,H,W] when want_labels."""
    src = logits if logits is not None else pred_labels
    _need_cuda(src, gt, hist)
    if logits is not None:
        logits = logits.detach()
        if logits.dtype != torch.float32 or not logits.is_contiguous():
            logits = logits.to(torch.float32).contiguous()
        N, C = logits.shape[0], logits.shape[1]
        HW = logits[0, 0].numel()
        shape = (N,) + tuple(logits.shape[2:])
    else:
        pred_labels = pred_labels.to(torch.int64).contiguous()
        N, HW = pred_labels.shape[0], pred_labels[0].numel()
        C = hist.shape[0]
        shape = tuple(pred_labels.shape)
    if gt is not None:
        gt = gt.to(torch.int64).contiguous()
        if gt.numel() != N * HW:
            raise ValueError("gt has %d elements, predictions have %d" % (gt.numel(), N * HW))
        if hist.dtype != torch.int64 or tuple(hist.shape) != (C, C) or not hist.is_contiguous():
            raise ValueError("hist must be a contiguous int64 [%d,%d] tensor" % (C, C))
    labels = torch.empty(shape, device=src.device, dtype=torch.uint8) if want_labels else None
    with torch.cuda.device(src.device):
        _lib.check(_lib.load().ctl_confusion_update(_ptr(logits), _ptr(pred_labels), _ptr(gt), N, C, HW,
                                                    _ptr(hist) if gt is not None else 0, _ptr(labels), _stream()))
    return labels


def argmax_labels(logits):
    """uint8 label map [N,H,W] = argmax over the class dimension of planar fp32 logits (first maximum)."""
    return confusion_update(None, None, logits=logits, want_labels=True)


def confusion_scores(hist):
    """Device fp64 [4 + C]: overall acc, mean acc, frequency-weighted acc, mean IoU, IoU per class."""
    _need_cuda(hist)
    C = hist.shape[0]
    out = torch.empty(4 + C, device=hist.device, dtype=torch.float64)
    with torch.cuda.device(hist.device):
        _lib.check(_lib.load().ctl_confusion_scores(hist.data_ptr(), C, out.data_ptr(), _stream()))
    return out


# ------------------------------------------------------------------------------------------------ fused latent saliency
def conv2d_c8_saliency(x, w_packed, cout, res, sal_sums, mode, store_out=False):
    """1x1 convolution out = conv(x) + res (the decoder's last input-gradient convolution) whose epilogue accumulates the
    per-sample saliency sums of its bf16-rounded output into `sal_sums` (fp64 [N, cout] channel mode / [N, H*W] spatial
    mode, zeroed by the caller).  store_out=False: the output -- dL/dz -- is not written at all; returns None then."""
    _need_cuda(x, w_packed, res, sal_sums)
    N, cin, H, W = _c8_dims(x)
    n = cout if mode == MODE_CHANNEL else H * W
    if sal_sums.dtype != torch.float64 or tuple(sal_sums.shape) != (N, n) or not sal_sums.is_contiguous():
        raise ValueError("sal_sums must be a contiguous float64 [%d,%d] tensor" % (N, n))
    if res is None or tuple(res.shape) != (N, cout // 8, H, W, 8):
        raise ValueError("res must be the C8 tensor the convolution output is added to")
    out = torch.empty((N, cout // 8, H, W, 8), device=x.device, dtype=torch.bfloat16) if store_out else None
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().ctl_conv2d_c8_bf16_saliency(x.data_ptr(), N, H, W, cin, w_packed.data_ptr(), cout,
                                                           res.data_ptr(), _ptr(out), sal_sums.data_ptr(), mode,
                                                           int(bool(store_out)), _stream()))
    return out


def saliency_sums_mask_apply(sums, z, mode, k, soft=False, rand=None, rng=None, want_thr=False, step_params=None,
                             want_c8=True):
    """The masking tail on saliency SUMS (see conv2d_c8_saliency).  Returns (z_masked NCHW fp32, mask [N,n], s [N,n],
    thr or None, z_masked as C8 bf16 or None)."""
    _need_cuda(sums, z, rand)
    z = z.contiguous()
    N, C, HW = _nchw(z)
    n = C if mode == MODE_CHANNEL else HW
    if z.dtype != torch.float32 or sums.dtype != torch.float64 or tuple(sums.shape) != (N, n):
        raise ValueError("expected fp32 z and float64 sums [%d,%d]" % (N, n))
    if rand is not None:
        rand = rand.contiguous()
        if tuple(rand.shape) != (N, n) or rand.dtype != torch.float32:
            raise ValueError("rand must be float32 [%d,%d]" % (N, n))
    native = soft and rand is None
    if native and rng is None:
        raise ValueError("soft masking needs either `rand` (torch-compatible mode) or `rng` (native mode)")
    s = torch.empty((N, n), device=z.device, dtype=torch.float32)
    z_out = torch.empty_like(z)
    mask = torch.empty((N, n), device=z.device, dtype=torch.float32)
    thr = torch.empty((N,), device=z.device, dtype=torch.float32) if want_thr else None
    c8 = torch.empty((N, C // 8, z.shape[2], z.shape[3], 8), device=z.device, dtype=torch.bfloat16) if want_c8 else None
    seed = offset = first = 0
    row = None
    if step_params is not None:
        if not 0 <= int(k) < n:
            raise IndexError("index {} is out of bounds for dimension 1 with size {}".format(int(k), n))
        row = step_params.take("mask", n, k, rng if native else None)
        seed = rng.seed if native else 0
    elif native:
        seed, offset, first = rng.seed, rng.next_offset(), rng.first_sample
    with torch.cuda.device(z.device):
        _lib.check(_lib.load().ctl_saliency_sums_mask_apply(
            sums.data_ptr(), z.data_ptr(), N, C, HW, mode, int(k), int(bool(soft)), _ptr(rand), seed, offset, first,
            _ptr(row), s.data_ptr(), mask.data_ptr(), _ptr(thr), z_out.data_ptr(), _ptr(c8), _stream()))
    return z_out, mask, s, thr, c8


# NCHW fp32 latent code -> its blocked bf16 twin written by the same kernel (saliency_sums_mask_apply): the decoder's
# forward picks the twin up instead of converting the layout again.  The entry keeps the NCHW tensor alive, so its
# address cannot be reused by another tensor while the entry exists; two entries (image code, shape code) are kept.
_C8_TWINS = []


def register_c8_twin(nchw, c8):
    _C8_TWINS.append((nchw, nchw._version, c8))
    del _C8_TWINS[:-2]


def c8_twin(nchw):
    for t, version, c8 in _C8_TWINS:
        if t.data_ptr() == nchw.data_ptr() and tuple(t.shape) == tuple(nchw.shape) and t._version == version \
                and nchw._version == version and nchw.dtype == t.dtype and nchw.is_contiguous():
            return c8
    return None


# ------------------------------------------------------------------------------------------------ loss gradients alone
def ce2d_grad(logits, target, scale):
    """d/dlogits of scale * sum_p -log softmax(logits)[target] (what autograd.grad of the scalar loss returns)."""
    x, t = logits.detach().contiguous(), target.contiguous()
    N, C, H, W = x.shape
    dx = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().ctl_ce2d_bwd(x.data_ptr(), t.data_ptr(), N, C, H, W, float(scale), 0, dx.data_ptr(), _stream()))
    return dx


def sse_grad(pred, target, scale):
    """d/dpred of scale * sum (pred - target)**2."""
    x, t = pred.detach().contiguous(), target.detach().contiguous()
    dx = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().ctl_sse_bwd(x.data_ptr(), t.data_ptr(), x.numel(), float(scale), 0, dx.data_ptr(), _stream()))
    return dx


# ------------------------------------------------------------------------------------------------ BN-backward reduction in the dgrad
def conv_bnbwd_fusable(cin, cout):
    """True when the 3x3 stride-1 convolution cin -> cout (packed-view channels of an input-gradient convolution) can
    accumulate the BatchNorm-backward sums of the layer its output flows into (ctl_conv2d_c8_bf16_bnbwd)."""
    lib = _lib.load()
    return (cin <= 64 and not lib.ctl_conv2d_vpacked(int(cin), int(cout), 9, 1)
            and 0 < lib.ctl_conv2d_n_tile(int(cin), int(cout), 9) <= 32)


def conv2d_c8_bnbwd(x, w_packed, cout, bn_a, bn_scale, bn_shift, bn_act, totals):
    """out = conv3x3(x) (an activation gradient) + the BatchNorm-backward sums of h = act(BN(bn_a)) accumulated into
    `totals` (zeroed float64 [2*cout]) by the epilogue."""
    _need_cuda(x, w_packed, bn_a, bn_scale, bn_shift, totals)
    N, cin, H, W = _c8_dims(x)
    if tuple(bn_a.shape) != (N, cout // 8, H, W, 8):
        raise ValueError("bn_a must be the C8 tensor of the output's shape")
    sc, sh = _vec(bn_scale, cout), _vec(bn_shift, cout)
    out = torch.empty((N, cout // 8, H, W, 8), device=x.device, dtype=torch.bfloat16)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().ctl_conv2d_c8_bf16_bnbwd(x.data_ptr(), N, H, W, cin, w_packed.data_ptr(), cout, bn_a.data_ptr(),
                                                        sc.data_ptr(), sh.data_ptr(), int(bn_act), out.data_ptr(),
                                                        totals.data_ptr(), _stream()))
    return out


def bn_bwd_apply_totals_c8(dy, a, act, mean, var, eps, gamma, totals, act_affine, want_param_grads=True):
    """The apply pass of the BatchNorm + activation backward from per-channel totals (sum dv | sum dv*a) that a
    convolution epilogue (conv2d_c8_bnbwd) or the stand-alone reduction left in `totals`.  Returns (da, dgamma, dbeta)."""
    _need_cuda(dy, a, mean, var, gamma, totals)
    N, C, H, W = _c8_dims(a)
    sc, sh = _vec(act_affine[0], C), _vec(act_affine[1], C)
    pg = torch.empty((2, C), device=a.device, dtype=torch.float32) if want_param_grads else None
    g = _vec(gamma, C)
    da = torch.empty_like(a)
    with torch.cuda.device(a.device):
        _lib.check(_lib.load().ctl_bn_bwd_apply_totals_c8(
            dy.data_ptr(), a.data_ptr(), N, C, H, W, act, mean.data_ptr(), var.data_ptr(), float(eps), _ptr(g),
            totals.data_ptr(), da.data_ptr(), pg[0].data_ptr() if pg is not None else 0,
            pg[1].data_ptr() if pg is not None else 0, sc.data_ptr(), sh.data_ptr(), _stream()))
    return da, (pg[0] if pg is not None else None), (pg[1] if pg is not None else None)


# ------------------------------------------------------------------------------------------------ BN finalisation in the apply pass
def bn_apply_from_sums_c8(x, sums, gamma, beta, eps, act, running_mean=None, running_var=None, momentum=0.1, low=None):
    """h = act(BatchNorm_train(x) [+ nearest_up2(low)]) from the per-channel sums (float64 [2, C]) the convolution that
    produced x accumulated: the finalisation (scale / shift / batch mean / variance, running-statistics update) runs in
    the prologue of the apply kernel.  Returns (h, scale, shift, mean, var)."""
    _need_cuda(x, sums, gamma, beta, running_mean, running_var, low)
    N, C, H, W = _c8_dims(x)
    if low is not None and (tuple(low.shape) != (N, C // 8, H // 2, W // 2, 8) or H % 2 or W % 2):
        raise ValueError("low must be C8 %s for x %s" % ((N, C // 8, H // 2, W // 2, 8), tuple(x.shape)))
    out = torch.empty((4, C), device=x.device, dtype=torch.float32)
    y = torch.empty_like(x)
    g, b = _vec(gamma, C), _vec(beta, C)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().ctl_bn_apply_from_sums_c8(
            x.data_ptr(), N, C, H, W, sums.data_ptr(), _ptr(g), _ptr(b), float(eps), _ptr(low), act, y.data_ptr(),
            out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), out[3].data_ptr(), _ptr(running_mean),
            _ptr(running_var), float(momentum), _stream()))
    return y, out[0], out[1], out[2], out[3]
