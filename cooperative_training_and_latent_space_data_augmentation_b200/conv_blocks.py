"""Arithmetic of the FTN/STN conv blocks.

Three numerics modes, selected with `set_precision`:
  'fp32'    parity mode: the exact op sequence of the reference in fp32 (what the reference itself
            runs on a GPU); used by the parity tests that compare against the CPU oracle at 1e-4.
  'kernel'  product mode: whole sub-networks run forward AND backward on this build's sm_100a kernels
            (trainpath.py / fastpath.py; bf16 C8 activations, fp32 master weights and statistics).  The
            functions below are then only reached for the one case the kernels do not cover (eval-mode
            BatchNorm with autograd enabled), where they behave like 'bf16'.
  'bf16'    library comparison mode: bf16 activations in NHWC (channels_last) through cuDNN.

Every function takes the reference-shaped nn.Module that owns the parameters, so BatchNorm
bookkeeping (training flag, track_running_stats, momentum, num_batches_tracked) follows the
module state exactly as in the reference -- including inside `_disable_tracking_bn_stats`.
"""
import torch
import torch.nn.functional as F

_PRECISION = "fp32"
LRELU_SLOPE = 0.2


def set_precision(mode):
    global _PRECISION
    if mode not in ("fp32", "bf16", "kernel"):
        raise ValueError("precision must be 'fp32', 'bf16' or 'kernel'")
    _PRECISION = mode
    _apply_tf32_policy()


def _apply_tf32_policy():
    # parity mode must be true fp32: torch lets cuDNN use TF32 for fp32 convolutions by default
    torch.backends.cudnn.allow_tf32 = _PRECISION != "fp32"
    torch.backends.cuda.matmul.allow_tf32 = _PRECISION != "fp32"


def get_precision():
    return _PRECISION


_apply_tf32_policy()


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
