"""FTN / STN networks of 'FCN_16_standard' with the reference's module tree, so that
`state_dict()` keys and shapes are interchangeable with reference checkpoints
(e.g. `general_encoder.down1.conv.0.weight [32,16,3,3]`, `up1.up.weight [128,128,2,2]`).

    res_convdown        <- medseg/models/ebm/encoder_decoder.py:19-68
    res_up_family       <- medseg/models/ebm/encoder_decoder.py:285-348
    MyEncoder           <- medseg/models/ebm/encoder_decoder.py:351-415
    MyDecoder           <- medseg/models/ebm/encoder_decoder.py:418-453
    Dual_Branch_Encoder <- medseg/models/ebm/encoder_decoder.py:456-503

Only the variants the ACDC configs instantiate are built (BatchNorm2d, no spectral norm, no
dropout, 'NN' / 'Conv2' up-sampling); anything else raises NotImplementedError.

The arithmetic of every block goes through `conv_blocks` (this package), which owns the choice
between the fp32 parity path and the bf16 NHWC tensor-core path.
"""
import torch
import torch.nn as nn

from . import conv_blocks as cb
from . import fastpath, ops, trainpath

LRELU_SLOPE = 0.2


def _kernel_route(module, x):
    """'train': forward+backward on the kernels (trainpath); 'eval': BN-folded no-grad forward (fastpath);
    None: the torch path of conv_blocks (parity / library modes, CPU tensors, eval-mode BN under autograd)."""
    if cb.get_precision() != 'kernel' or not x.is_cuda:
        return None
    if module.training:
        return 'train'
    return None if torch.is_grad_enabled() else 'eval'


def _only_supported(if_SN, dropout, norm):
    if if_SN:
        raise NotImplementedError("spectral norm variants are not part of the ACDC hot path")
    if dropout is not None:
        raise NotImplementedError("encoder/decoder dropout is None in every shipped config")
    if norm is not nn.BatchNorm2d:
        raise NotImplementedError("the solver always passes norm=nn.BatchNorm2d")


def _conv_bn_lrelu_conv_bn(in_ch, out_ch, norm):
    return nn.Sequential(
        nn.Conv2d(in_ch, out_ch, 3, padding=1, bias=True), norm(out_ch), nn.LeakyReLU(LRELU_SLOPE),
        nn.Conv2d(out_ch, out_ch, 3, padding=1, bias=True), norm(out_ch))


class res_convdown(nn.Module):
    """x' = conv3x3/s2(x);  out = LReLU(conv1x1(x') + BN(conv3(LReLU(BN(conv3(x'))))))."""

    def __init__(self, in_ch, out_ch, norm=nn.BatchNorm2d, if_SN=False, bias=True, dropout=None):
        super().__init__()
        _only_supported(if_SN, dropout, norm)
        self.down = nn.Conv2d(in_ch, in_ch, 3, stride=2, padding=1, bias=bias)
        self.conv = _conv_bn_lrelu_conv_bn(in_ch, out_ch, norm)
        self.conv_input = nn.Conv2d(in_ch, out_ch, kernel_size=1, stride=1, padding=0, bias=True)
        self.last_act = nn.LeakyReLU(LRELU_SLOPE)

    def forward(self, x):
        return cb.residual_block(self, cb.resample_down(self.down, x))


class res_up_family(nn.Module):
    """x' = up(x) ('NN': nearest x2, 'Conv2': ConvTranspose2d k2 s2);  same residual body."""

    def __init__(self, in_ch, out_ch, norm=nn.BatchNorm2d, if_SN=False, bias=True, dropout=None, up_type='bilinear'):
        super().__init__()
        _only_supported(if_SN, dropout, norm)
        if up_type == 'NN':
            self.up = nn.Sequential(nn.UpsamplingNearest2d(scale_factor=2))
        elif up_type == 'Conv2':
            self.up = nn.ConvTranspose2d(in_ch, in_ch, kernel_size=2, stride=2)
        else:
            raise NotImplementedError("up_type %r is not used by FCN_16_standard" % (up_type,))
        self.up_type = up_type
        self.conv = _conv_bn_lrelu_conv_bn(in_ch, out_ch, norm)
        self.conv_input = nn.Conv2d(in_ch, out_ch, kernel_size=1, stride=1, padding=0, bias=True)
        self.last_act = nn.LeakyReLU(LRELU_SLOPE)

    def forward(self, x):
        return cb.residual_block(self, cb.resample_up(self.up, self.up_type, x))


class MyEncoder(nn.Module):
    def __init__(self, input_channel, output_channel=None, feature_reduce=1, encoder_dropout=None,
                 norm=nn.BatchNorm2d, if_SN=False, act=torch.nn.Sigmoid()):
        super().__init__()
        _only_supported(if_SN, encoder_dropout, norm)
        w = [64 // feature_reduce, 128 // feature_reduce, 256 // feature_reduce, 512 // feature_reduce]
        self.inc = _conv_bn_lrelu_conv_bn(input_channel, w[0], norm)
        self.down1 = res_convdown(w[0], w[1], norm=norm)
        self.down2 = res_convdown(w[1], w[2], norm=norm)
        self.down3 = res_convdown(w[2], w[3], norm=norm)
        self.down4 = res_convdown(w[3], w[3], norm=norm)
        self.final_conv = nn.Sequential(nn.Conv2d(w[3], w[3], kernel_size=1, stride=1, padding=0), norm(w[3]))
        self.act = act

    def forward(self, x):
        route = _kernel_route(self, x)
        if route == 'train':
            return trainpath.encoder_apply(self, x)
        if route == 'eval':
            return ops.c8_to_nchw(fastpath.encoder_forward(self, x, 'eval'))
        x = cb.stem(self.inc, x)
        x = self.down4(self.down3(self.down2(self.down1(x))))
        return cb.conv_bn_act(self.final_conv[0], self.final_conv[1], x, self.act)

    def forward_from_segmentation(self, segmentation, is_label_map=False, temperature=2):
        """STN entry (construct_input fused into the stem kernel): logits -> softmax(x / T), label map -> one-hot.
        Only valid on the kernel routes; the caller falls back to construct_input + forward otherwise."""
        route = _kernel_route(self, segmentation)
        in_mode = 2 if is_label_map else 1
        if route == 'train':
            return trainpath.encoder_apply(self, segmentation, in_mode=in_mode, temperature=temperature)
        if route == 'eval':
            return ops.c8_to_nchw(fastpath.encoder_forward(self, segmentation, 'eval', in_mode=in_mode,
                                                           temperature=temperature))
        return None


class MyDecoder(nn.Module):
    def __init__(self, input_channel, output_channel, feature_reduce=1, decoder_dropout=None, norm=nn.BatchNorm2d,
                 up_type='bilinear', if_SN=False, last_act=None):
        super().__init__()
        _only_supported(if_SN, decoder_dropout, norm)
        w = [256 // feature_reduce, 128 // feature_reduce, 64 // feature_reduce]
        self.up1 = res_up_family(input_channel, w[0], norm=norm, up_type=up_type)
        self.up2 = res_up_family(w[0], w[1], norm=norm, up_type=up_type)
        self.up3 = res_up_family(w[1], w[2], norm=norm, up_type=up_type)
        self.up4 = res_up_family(w[2], w[2], norm=norm, up_type=up_type)
        self.final_conv = nn.Conv2d(w[2], output_channel, kernel_size=1, stride=1, padding=0)
        # the reference runs normal_init over the decoder's DIRECT children (encoder_decoder.py:441-442, :13-16): of those
        # only final_conv is a convolution -- its bias starts at zero (the weight is re-initialised by kaiming later)
        nn.init.zeros_(self.final_conv.bias)
        self.last_act = last_act

    def forward(self, x):
        route = _kernel_route(self, x)
        if route == 'train':
            return trainpath.decoder_apply(self, x)
        if route == 'eval':
            return fastpath.decoder_from_nchw(self, x, 'eval')
        x = self.up4(self.up3(self.up2(self.up1(x))))
        return cb.head(self.final_conv, x, self.last_act)


class Dual_Branch_Encoder(nn.Module):
    """FTN encoder: z_i = general_encoder(x), z_s = code_decoupler(z_i)."""

    def __init__(self, input_channel, z_level_1_channel=None, z_level_2_channel=None, feature_reduce=1,
                 encoder_dropout=None, norm=nn.BatchNorm2d, if_SN=False):
        super().__init__()
        _only_supported(if_SN, encoder_dropout, norm)
        self.general_encoder = MyEncoder(input_channel, output_channel=z_level_1_channel, feature_reduce=feature_reduce,
                                         encoder_dropout=encoder_dropout, norm=norm, act=torch.nn.ReLU())
        c1, c2 = z_level_1_channel, z_level_2_channel
        self.code_decoupler = nn.Sequential(
            nn.Conv2d(c1, c2, 3, padding=1, bias=True), norm(c2), nn.LeakyReLU(LRELU_SLOPE),
            nn.Conv2d(c2, c2, 3, padding=1, bias=True), norm(c2), nn.ReLU())

    def filter_code(self, z):
        route = _kernel_route(self, z)
        if route == 'train':
            return trainpath.decoupler_apply(self, z)
        if route == 'eval':
            return ops.c8_to_nchw(fastpath.filter_code(self, ops.nchw_to_c8(z), 'eval'))
        return cb.double_conv(self.code_decoupler, z, final_act=self.code_decoupler[5])

    def forward(self, x):
        route = _kernel_route(self, x)
        if route == 'train':
            return trainpath.dual_encoder_apply(self, x)
        if route == 'eval':
            z_i = fastpath.encoder_forward(self.general_encoder, x, 'eval')
            return ops.c8_to_nchw(z_i), ops.c8_to_nchw(fastpath.filter_code(self, z_i, 'eval'))
        z_i = self.general_encoder(x)
        return z_i, self.filter_code(z_i)


def init_weights_kaiming(net):
    """init_weights(net, 'kaiming') (medseg/models/init_weight.py:30-39): kaiming-normal fan_in on every
    nn.Conv2d, BN gamma ~ N(1, 0.02), beta = 0; ConvTranspose2d keeps torch's default."""
    for m in net.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight.data, a=0, mode='fan_in')
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.normal_(m.weight.data, 1.0, 0.02)
            nn.init.constant_(m.bias.data, 0.0)
    return net
