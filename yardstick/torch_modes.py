"""Library-convolution yardsticks for the FTN/STN conv blocks -- TEST AND BASELINE INFRASTRUCTURE, not product code.

The product package (cooperative_training_and_latent_space_data_augmentation_b200/) computes the conv blocks only on
its own sm_100a kernels ('kernel' precision) and has no library path.  This module restates the blocks with torch /
cuDNN ops on the reference-shaped nn.Modules, in two modes:

  'fp32'  the exact op sequence of the reference in true fp32 (TF32 off) -- what the reference itself runs on a GPU
          (medseg/models/ebm/encoder_decoder.py:19-68, :285-348, :351-415, :418-453, :456-503); the parity tests compare
          the kernel path and the CPU oracle against it, and bench.py times it as `eager_gpu_baseline`
  'bf16'  bf16 NHWC activations through cuDNN (what a recompiled library path would give)

`install()` registers this module with the product's conv_blocks hook; only tests/conftest.py, bench.py's
eager_gpu_baseline leg, tools/ and __graft_entry__.smoke() do that.
"""
import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.2
_PRECISION = "fp32"
_SAVED_TF32 = None


def set_mode(mode):
    """Called by conv_blocks.set_precision: 'fp32' switches TF32 off for cuDNN / matmul (true fp32 parity), anything
    else restores what the process had before."""
    global _PRECISION, _SAVED_TF32
    if mode == "fp32":
        if _SAVED_TF32 is None:
            _SAVED_TF32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    elif _SAVED_TF32 is not None:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = _SAVED_TF32
        _SAVED_TF32 = None
    _PRECISION = mode if mode in ("fp32", "bf16") else "bf16"      # 'kernel': eval-mode BN under autograd -> bf16 ops


def install():
    import importlib
    import sys
    cb = importlib.import_module("cooperative_training_and_latent_space_data_augmentation_b200.conv_blocks")
    cb.install_yardstick(sys.modules[__name__])


def _prep(x):
    """bf16 mode keeps activations NHWC bf16 end to end; fp32 mode leaves the tensor alone."""
    if _PRECISION != "fp32":
        if x.dtype != torch.bfloat16:
            x = x.to(torch.bfloat16)
        return x.contiguous(memory_format=torch.channels_last)
    return x


def _conv(conv, x):
    if _PRECISION != "fp32":
        return F.conv2d(x, conv.weight.to(torch.bfloat16), conv.bias.to(torch.bfloat16) if conv.bias is not None
                        else None, conv.stride, conv.padding)
    return conv(x)


def _bn(bn, x):
    # nn.BatchNorm2d.forward handles training / eval / track_running_stats=False (batch stats, no update);
    # cuDNN computes the statistics in fp32 for bf16 inputs.
    return bn(x)


def resample_down(down, x):
    return _conv(down, _prep(x))


def resample_up(up, up_type, x):
    x = _prep(x)
    if up_type == 'NN':
        return F.interpolate(x, scale_factor=2, mode='nearest')
    if _PRECISION != "fp32":
        return F.conv_transpose2d(x, up.weight.to(torch.bfloat16), up.bias.to(torch.bfloat16), stride=2)
    return up(x)


def double_conv(seq, x, final_act=None):
    """conv3x3 - BN - LReLU(0.2) - conv3x3 - BN [- act]   (nn.Sequential indices 0,1,2,3,4[,5])."""
    x = _prep(x)
    y = F.leaky_relu(_bn(seq[1], _conv(seq[0], x)), LRELU_SLOPE)
    y = _bn(seq[4], _conv(seq[3], y))
    return final_act(y) if final_act is not None else y


def residual_block(block, x):
    """LReLU(conv1x1(x) + double_conv(x)); x is already resampled."""
    return F.leaky_relu(_conv(block.conv_input, x) + double_conv(block.conv, x), LRELU_SLOPE)


def stem(inc, x):
    return F.leaky_relu(double_conv(inc, x), LRELU_SLOPE)


def conv_bn_act(conv, bn, x, act):
    y = _bn(bn, _conv(conv, _prep(x)))
    return act(y) if act is not None else y


def head(conv, x, last_act):
    y = _conv(conv, _prep(x))
    return last_act(y) if last_act is not None else y
