"""Numerics-mode switch of the FTN/STN conv blocks.

The product computes the conv blocks on this build's sm_100a kernels only -- `'kernel'`, the default: whole
sub-networks run forward AND backward through trainpath.py / fastpath.py (bf16 C8 activations, fp32 master weights and
statistics).  There is no library (cuDNN) path in this package and no dispatch between back ends.

For parity testing and for the eager-GPU baseline of bench.py a *yardstick* -- torch / cuDNN restatements of the same
blocks, `yardstick/torch_modes.py` at the repository root, beside the oracle -- can be installed with
`install_yardstick(module)`; only then are `set_precision('fp32')` (the reference's own fp32 op sequence) and
`set_precision('bf16')` (cuDNN bf16) accepted, and only then do inputs the kernels do not cover (eval-mode BatchNorm
with autograd enabled) have somewhere to go.  Without a yardstick those cases raise.
"""
_PRECISION = "kernel"
_YARDSTICK = None
LRELU_SLOPE = 0.2


def install_yardstick(module):
    """Registers the torch-ops restatement of the blocks (test / baseline infrastructure, never imported by the package)."""
    global _YARDSTICK
    _YARDSTICK = module


def yardstick_installed():
    return _YARDSTICK is not None


def set_precision(mode):
    global _PRECISION
    if mode not in ("fp32", "bf16", "kernel"):
        raise ValueError("precision must be 'kernel' (product) or, with a yardstick installed, 'fp32' / 'bf16'")
    if mode != "kernel" and _YARDSTICK is None:
        raise RuntimeError("precision %r is a library-convolution yardstick mode; the product package has no such path. "
                           "Tests and baselines install one first: `import yardstick; yardstick.install()`" % (mode,))
    _PRECISION = mode
    if _YARDSTICK is not None:
        _YARDSTICK.set_mode(mode)


def get_precision():
    return _PRECISION


def _yardstick(what):
    if _YARDSTICK is None:
        raise NotImplementedError(
            "%s is not covered by the sm_100a kernels of this build (they take CUDA tensors, train-mode BatchNorm, or "
            "eval-mode BatchNorm under torch.no_grad(), H and W multiples of 16) and the product has no library "
            "fallback" % what)
    return _YARDSTICK


def resample_down(down, x):
    return _yardstick("this stride-2 convolution").resample_down(down, x)


def resample_up(up, up_type, x):
    return _yardstick("this up-sampling block").resample_up(up, up_type, x)


def double_conv(seq, x, final_act=None):
    return _yardstick("this double convolution").double_conv(seq, x, final_act)


def residual_block(block, x):
    return _yardstick("this residual block").residual_block(block, x)


def stem(inc, x):
    return _yardstick("this encoder stem").stem(inc, x)


def conv_bn_act(conv, bn, x, act):
    return _yardstick("this convolution").conv_bn_act(conv, bn, x, act)


def head(conv, x, last_act):
    return _yardstick("this output head").head(conv, x, last_act)
