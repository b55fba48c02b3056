"""Losses and STN-input construction on the hot path.

    basic_loss_fn('cross entropy')  <- medseg/models/custom_loss.py:8-19 -> cross_entropy_2D :706-741
    construct_input                 <- medseg/common_utils/basic_operations.py:110-158

The reference's training cross entropy builds a `ones_like` mask, transposes the log-probabilities
to NHWC and calls `float(torch.sum(mask[:, 0]))` -- a device->host sync on every call (7 per
cooperative step).  With mask=None that divisor is the constant N*H*W,
so it is computed on the host from the shape and the loss never leaves the device.
"""
import torch
import torch.nn.functional as F

from . import ops


def cross_entropy_2D(input, target, weight=None, size_average=True, mask=None, is_gt=False):
    """sum over pixels of NLL (or of -q*log p for 4-D soft targets) / number of (unmasked) pixels."""
    n, c, h, w = input.size()
    if mask is None and weight is None and ops.ce2d_supported(input, target):
        # one fused pass (csrc/loss.cu): no log-probability tensor, no NHWC transpose, divisor folded into the kernel
        return ops.cross_entropy_2d(input, target, 1.0 / float(n * h * w) if size_average else 1.0)
    log_p = F.log_softmax(input.float(), dim=1)
    if mask is not None:
        mask = (mask != 0).to(log_p.dtype)
        region = mask[:, 0].sum()                 # stays on the device
    else:
        region = float(n * h * w)
    if target.dim() == 3:
        if weight is not None:
            weight = weight / weight.sum() * c
        nll = F.nll_loss(log_p, target, weight=weight, reduction='none')
        if mask is not None:
            nll = nll * mask[:, 0]
        loss = nll.sum()
        return loss / region if size_average else loss
    if target.dim() == 4:
        q = target.float() if is_gt else F.softmax(target.float(), dim=1)
        prod = q * log_p
        if mask is not None:
            prod = prod * mask
        if weight is not None:
            wt = torch.as_tensor(weight, dtype=prod.dtype, device=prod.device)
            wt = wt / wt.sum() * c
            prod = prod * wt.view(1, c, 1, 1)
        total = prod.sum()
        if size_average:
            total = total / region
        return -1 * total
    raise NotImplementedError


def half_mse_loss(pred, target):
    """0.5 * nn.MSELoss()(pred, target) (advanced...model.py:443-447) -- one fused kernel each way on CUDA fp32."""
    pred = pred.float()
    if ops.sse_supported(pred, target):
        return ops.squared_error(pred, target, 0.5 / pred.numel())
    return 0.5 * F.mse_loss(pred, target, reduction='mean')


def basic_loss_fn(pred, target, loss_type='cross_entropy', class_weights=None, use_gpu=True):
    """Only the losses the ACDC configs select are built ('cross entropy', 'weighted cross entropy')."""
    num_classes = pred.size(1)
    if loss_type == 'cross entropy':
        return cross_entropy_2D(pred, target)
    if loss_type == 'weighted cross entropy':
        if class_weights is None:
            class_weights = num_classes * [1. / num_classes]
        assert len(class_weights) == num_classes
        return cross_entropy_2D(pred, target, torch.tensor(class_weights, dtype=torch.float32, device=pred.device))
    raise NotImplementedError("loss_type %r is not on the cooperative-training path" % (loss_type,))


def construct_input(segmentation, image=None, num_classes=None, apply_softmax=True, temperature=2,
                    is_labelmap=False, smooth_label=False, shuffle=False, use_gpu=True):
    """STN input: softmax(logit / T) for predictions, one-hot for label maps (optionally concatenated with
    an image, as the reference allows)."""
    assert (apply_softmax and is_labelmap) is False
    if not is_labelmap:
        if apply_softmax:
            assert segmentation.dim() == 4
            segmentation = torch.softmax(segmentation.float() / temperature, dim=1)
    else:
        assert num_classes is not None, 'please specify num_classes'
        onehot = F.one_hot(segmentation.long(), num_classes).permute(0, 3, 1, 2).to(torch.float32)
        if smooth_label:
            alpha = torch.rand(1, device=onehot.device) * 0.1
            onehot = (1 - alpha) * onehot + alpha / num_classes
        segmentation = onehot
    if shuffle and image is not None:
        image = torch.roll(image, shifts=-1, dims=0)
    if image is not None:
        return torch.cat([segmentation, image], dim=1)
    return segmentation
