"""TEST INFRASTRUCTURE ONLY -- CPU (torch fp32) oracle for the FTN/STN networks, the losses
and the three-pass cooperative step.

Restates, from scratch, the arithmetic of
  medseg/models/ebm/encoder_decoder.py:19-68 (res_convdown), :285-348 (res_up_family),
      :351-415 (MyEncoder), :418-453 (MyDecoder), :456-503 (Dual_Branch_Encoder)
  medseg/models/model_util.py:104-135 (cross_entropy_2D used by the saliency 'ce' loss),
      :168-177 (make_one_hot), :414-451 (_disable_tracking_bn_stats)
  medseg/models/custom_loss.py:8-19, :706-741 (training cross entropy)
  medseg/common_utils/basic_operations.py:110-158 (construct_input)
  medseg/models/advanced_triplet_recon_segmentation_model.py:300-350, :396-601 (solver)
  medseg/train_adv_supervised_segmentation_triplet.py:171-237 (one cooperative step)
with state_dict keys identical to the reference modules so that weights move both ways.

Parity status: pinned against the imported reference (tests/golden/model_*.npz and
tests/test_oracle_vs_reference.py, which runs the two side by side when /root/reference is
present).  The reference itself holds no golden vectors.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this file.  The product package never does.
"""
import random as _pyrandom

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import masking_oracle as mo

LRELU = 0.2
WIDTHS = (16, 32, 64, 128)          # reduce_factor = 4 (advanced...model.py:88-90)
LATENT_CH = 128


class TrackableBN(nn.BatchNorm2d):
    """BatchNorm2d whose forward takes `track`.  track=False restates the effect of the
    reference's _disable_tracking_bn_stats (model_util.py:414-451): batch statistics are
    used, running stats / num_batches_tracked do not move, and gamma/beta are constants
    (requires_grad switched off for the pass => no gradient reaches them)."""

    def forward(self, x, track=True):
        if not self.training:
            return F.batch_norm(x, self.running_mean, self.running_var, self.weight, self.bias,
                                False, 0.0, self.eps)
        if track:
            self.num_batches_tracked.add_(1)
            return F.batch_norm(x, self.running_mean, self.running_var, self.weight, self.bias,
                                True, self.momentum, self.eps)
        return F.batch_norm(x, None, None, self.weight.detach(), self.bias.detach(),
                            True, 0.0, self.eps)


def _double_conv(cin, cout):
    # indices 0,1,(2),3,4 mirror the reference's nn.Sequential numbering (2 = LeakyReLU)
    return nn.ModuleList([nn.Conv2d(cin, cout, 3, padding=1), TrackableBN(cout), nn.Identity(),
                          nn.Conv2d(cout, cout, 3, padding=1), TrackableBN(cout)])


def _run_double_conv(seq, x, track):
    y = F.leaky_relu(seq[1](seq[0](x), track), LRELU)
    return seq[4](seq[3](y), track)


class ResBlock(nn.Module):
    """out = LReLU(conv1x1(x') + BN(conv3(LReLU(BN(conv3(x'))))));  x' = resample(x).
    resample: 'down' (3x3 stride-2 conv), 'NN' (nearest x2), 'Conv2' (ConvTranspose 2x2 s2)."""

    def __init__(self, cin, cout, resample):
        super().__init__()
        self.resample = resample
        if resample == "down":
            self.down = nn.Conv2d(cin, cin, 3, stride=2, padding=1)
        elif resample == "Conv2":
            self.up = nn.ConvTranspose2d(cin, cin, kernel_size=2, stride=2)
        elif resample != "NN":
            raise NotImplementedError(resample)
        self.conv = _double_conv(cin, cout)
        self.conv_input = nn.Conv2d(cin, cout, 1)

    def forward(self, x, track=True):
        if self.resample == "down":
            x = self.down(x)
        elif self.resample == "Conv2":
            x = self.up(x)
        else:
            x = F.interpolate(x, scale_factor=2, mode="nearest")
        return F.leaky_relu(self.conv_input(x) + _run_double_conv(self.conv, x, track), LRELU)


class Encoder(nn.Module):
    def __init__(self, in_ch):
        super().__init__()
        w = WIDTHS
        self.inc = _double_conv(in_ch, w[0])
        # NB: first conv of inc is in_ch -> w0, second w0 -> w0 (handled by _double_conv)
        self.down1 = ResBlock(w[0], w[1], "down")
        self.down2 = ResBlock(w[1], w[2], "down")
        self.down3 = ResBlock(w[2], w[3], "down")
        self.down4 = ResBlock(w[3], w[3], "down")
        self.final_conv = nn.ModuleList([nn.Conv2d(w[3], w[3], 1), TrackableBN(w[3])])

    def forward(self, x, track=True):
        x = F.leaky_relu(_run_double_conv(self.inc, x, track), LRELU)
        for blk in (self.down1, self.down2, self.down3, self.down4):
            x = blk(x, track)
        x = self.final_conv[1](self.final_conv[0](x), track)
        return F.relu(x)


class DualBranchEncoder(nn.Module):
    def __init__(self, in_ch):
        super().__init__()
        self.general_encoder = Encoder(in_ch)
        c = LATENT_CH
        self.code_decoupler = nn.ModuleList([nn.Conv2d(c, c, 3, padding=1), TrackableBN(c), nn.Identity(),
                                             nn.Conv2d(c, c, 3, padding=1), TrackableBN(c)])

    def filter_code(self, z, track=True):
        return F.relu(_run_double_conv(self.code_decoupler, z, track))

    def forward(self, x, track=True):
        z_i = self.general_encoder(x, track)
        return z_i, self.filter_code(z_i, track)


class Decoder(nn.Module):
    def __init__(self, out_ch, up_type, sigmoid=False):
        super().__init__()
        w = WIDTHS
        self.up1 = ResBlock(LATENT_CH, w[2], up_type)
        self.up2 = ResBlock(w[2], w[1], up_type)
        self.up3 = ResBlock(w[1], w[0], up_type)
        self.up4 = ResBlock(w[0], w[0], up_type)
        self.final_conv = nn.Conv2d(w[0], out_ch, 1)
        self.sigmoid = sigmoid

    def forward(self, x, track=True):
        for blk in (self.up1, self.up2, self.up3, self.up4):
            x = blk(x, track)
        x = self.final_conv(x)
        return torch.sigmoid(x) if self.sigmoid else x


def build_networks(image_ch=1, num_classes=4, seed=None):
    """The five sub-networks of 'FCN_16_standard' (advanced...model.py:92-106)."""
    if seed is not None:
        torch.manual_seed(seed)
    nets = {
        "image_encoder": DualBranchEncoder(image_ch),
        "segmentation_decoder": Decoder(num_classes, "NN"),
        "shape_encoder": Encoder(num_classes),
        "shape_decoder": Decoder(num_classes, "NN"),
        "image_decoder": Decoder(image_ch, "Conv2", sigmoid=True),
    }
    for net in nets.values():       # init_weights(..., 'kaiming') (init_weight.py:30-39)
        for m in net.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, a=0, mode="fan_in")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.normal_(m.weight, 1.0, 0.02)
                nn.init.zeros_(m.bias)
    return nets


# --------------------------------------------------------------------------- losses / inputs
def one_hot(label, num_classes):
    """model_util.py:168-177 / basic_operations.py:133-141."""
    return F.one_hot(label.long(), num_classes).permute(0, 3, 1, 2).to(torch.float32)


def stn_input(seg, is_label_map, num_classes=4, temperature=2):
    """construct_input (basic_operations.py:110-158) as the solver calls it
    (advanced...model.py:240-241): one-hot for label maps, softmax(logit/T) otherwise."""
    if is_label_map:
        return one_hot(seg, num_classes)
    return torch.softmax(seg / temperature, dim=1)


def ce_training(logit, label):
    """custom_loss.py:706-741 with mask=None, weight=None: sum NLL / (N*H*W)."""
    n, c, h, w = logit.shape
    logp = F.log_softmax(logit, dim=1)
    picked = torch.gather(logp, 1, label.long().view(n, 1, h, w))
    return -(picked.sum()) / float(n * h * w)


def ce_saliency(logit, label):
    """model_util.py:104-135 with a 3-D target: sum NLL / (numel + 1e-10)."""
    n, c, h, w = logit.shape
    logp = F.log_softmax(logit, dim=1)
    picked = torch.gather(logp, 1, label.long().view(n, 1, h, w))
    return -(picked.sum()) / float(label.numel() + 1e-10)


def saliency_loss(decoder, code, label, loss_type, num_classes):
    """model_util.py:205-221.  `decoder` is called in whatever mode it is in: BN training
    with tracking ON."""
    if loss_type == "ce":
        return ce_saliency(decoder(code), label)
    gt = one_hot(label, num_classes) if label.dim() < code.dim() else label
    out = decoder(code)
    if loss_type == "mse":
        return torch.mean((out - gt) ** 2)
    if loss_type == "corr":
        return torch.mean(out * gt)
    raise AssertionError("not implemented loss")


def latent_gradient(decoder, latent, label, loss_type, num_classes):
    code = latent.detach().to(torch.float32).clone().requires_grad_(True)
    loss = saliency_loss(decoder, code, label, loss_type, num_classes)
    (g,) = torch.autograd.grad(loss, [code])
    return code, g


def mask_latent_code(latent, decoder, label, mode, num_classes=2, percentile=1 / 3.0, random=False,
                     loss_type="corr", if_detach=True, if_soft=False, saliency="reference"):
    """mask_latent_code_{channel,spatial}_wise (model_util.py:180-318).  RNG: numpy global for
    p, torch global for the soft values -- the same streams, in the same order, as the reference.
    saliency='reference' reduces with torch.mean (bit-identical to the reference on this device);
    saliency='f64' uses the order-independent definition the CUDA kernel implements."""
    code, g = latent_gradient(decoder, latent, label, loss_type, num_classes)
    N, C, H, W = code.shape
    n = C if mode == mo.MODE_CHANNEL else H * W
    if saliency == "reference":
        s = torch.from_numpy(mo.saliency_reduce_reference_order(g.numpy(), mode).copy())
    else:
        s = torch.from_numpy(mo.saliency_reduce(g.numpy(), mode))
    k, _ = mo.threshold_index(n, percentile, random, np.random)
    thr = mo.topp_threshold(s.numpy(), k)
    rand = torch.rand_like(s).numpy() if if_soft else None
    vec = mo.build_mask(s.numpy(), thr, if_soft, rand)
    shape = (N, C, 1, 1) if mode == mo.MODE_CHANNEL else (N, 1, H, W)
    mask_all = torch.from_numpy(vec).view(shape)
    masked = (code if if_detach else latent) * mask_all
    return masked, mask_all, g


class OracleSolver:
    """The slice of AdvancedTripletReconSegmentationModel on the hot path, CPU fp32."""

    def __init__(self, num_classes=4, image_ch=1, learning_rate=1e-4, seed=None):
        self.num_classes = num_classes
        self.model = build_networks(image_ch, num_classes, seed)
        self.optimizers = {k: torch.optim.Adam(m.parameters(), lr=learning_rate)
                           for k, m in self.model.items()}      # advanced...model.py:774-781
        self.training = True
        self.z_i = self.z_s = None

    # -- mode / state --------------------------------------------------------
    def train(self):
        self.training = True
        for m in self.model.values():
            m.train()
            for p in m.parameters():
                p.requires_grad_(True)

    def eval(self):
        self.training = False
        for m in self.model.values():
            m.eval()

    def load_state_dicts(self, dicts):
        for k, m in self.model.items():
            m.load_state_dict(dicts[k])

    def zero_grad(self):
        for opt in self.optimizers.values():
            opt.zero_grad()

    # -- forward pieces ------------------------------------------------------
    def fast_predict(self, x, track=True):
        if not self.training:
            with torch.no_grad():
                z_i, z_s = self.model["image_encoder"](x)
                return (z_i, z_s), self.model["segmentation_decoder"](z_s)
        z_i, z_s = self.model["image_encoder"](x, track)
        return (z_i, z_s), self.model["segmentation_decoder"](z_s, track)

    def recon_shape(self, seg, is_label_map=False, track=True):
        inp = stn_input(seg, is_label_map, self.num_classes)
        return self.model["shape_decoder"](self.model["shape_encoder"](inp, track), track)

    def standard_training(self, clean, label, perturbed, compute_gt_recon=True, update_latent=True,
                          track=True):
        """advanced...model.py:414-467 (separate_training=False)."""
        (z_i, z_s), y0 = self.fast_predict(perturbed, track)
        if update_latent:
            self.z_i, self.z_s = z_i, z_s
        seg = ce_training(y0, label)
        rec = 0.5 * F.mse_loss(self.model["image_decoder"](z_i), clean)   # image decoder always tracks
        gt = ce_training(self.recon_shape(label.detach().clone(), True), label) if compute_gt_recon \
            else torch.tensor(0.0)
        shp = ce_training(self.recon_shape(y0, False, track), label)
        return seg, rec, gt, shp

    def perturb_latent_code(self, latent, decoder, label_y=None, perturb_type="random", threshold=0.5,
                            if_soft=False, random_threshold=False, loss_type="mse", if_detach=False):
        """advanced...model.py:300-350."""
        assert perturb_type in ["random", "dropout", "spatial", "channel"], "invalid method name"
        if perturb_type == "random":
            cands = ["dropout", "spatial", "channel"]
            _pyrandom.shuffle(cands)
            perturb_type = cands[0]
        if perturb_type == "dropout":
            masked = F.dropout2d(latent, p=threshold)
            mask = torch.where(masked == latent, torch.ones_like(masked), torch.zeros_like(masked))
        else:
            assert loss_type in ["mse", "ce", "corr"], "not implemented loss"
            mode = mo.MODE_SPATIAL if perturb_type == "spatial" else mo.MODE_CHANNEL
            masked, mask, _ = mask_latent_code(latent, decoder, label_y, mode, self.num_classes, threshold,
                                               random_threshold, loss_type, if_detach, if_soft)
        if if_detach:
            masked = masked.detach().clone()
        return masked, mask, perturb_type

    def hard_example_generation(self, clean, label, gen_corrupted_seg=True, gen_corrupted_image=True,
                                corrupted_image_DA_config=None, corrupted_seg_DA_config=None):
        """advanced...model.py:469-523."""
        dflt = {"mask_type": "random", "max_threshold": 0.5, "random_threshold": True, "if_soft": True}
        icfg = corrupted_image_DA_config or dict(dflt, loss_name="mse")
        scfg = corrupted_seg_DA_config or dict(dflt, loss_name="ce")
        frozen = [self.model["segmentation_decoder"], self.model["image_decoder"]]
        for m in frozen:
            for p in m.parameters():
                p.requires_grad_(False)
        out_img = out_seg = None
        self.last_types = [None, None]
        if gen_corrupted_image:
            self.zero_grad()
            dec = self.model["image_decoder"]
            zt, _, self.last_types[0] = self.perturb_latent_code(
                self.z_i, dec, clean, icfg["mask_type"], icfg["max_threshold"], icfg["if_soft"],
                icfg["random_threshold"], icfg["loss_name"], True)
            out_img = dec(zt, False)            # decoder_inference(..., disable_track_bn_stats=True)
        if gen_corrupted_seg:
            self.zero_grad()
            dec = self.model["segmentation_decoder"]
            zt, _, self.last_types[1] = self.perturb_latent_code(
                self.z_s, dec, label, scfg["mask_type"], scfg["max_threshold"], scfg["if_soft"],
                scfg["random_threshold"], scfg["loss_name"], True)
            out_seg = dec(zt, False)
        for m in frozen:
            for p in m.parameters():
                p.requires_grad_(True)
        return out_img, out_seg

    def hard_example_training(self, perturbed_image, clean, perturbed_seg, label):
        """advanced...model.py:525-559."""
        zero = torch.tensor(0.0)
        seg = rec = shp = pshp = zero
        if perturbed_image is not None:
            seg, rec, _, shp = self.standard_training(clean, label, perturbed_image.detach().clone(),
                                                      compute_gt_recon=False, update_latent=False, track=False)
        if perturbed_seg is not None:
            pshp = ce_training(self.recon_shape(perturbed_seg, False, False), label)
        return seg, rec, shp, pshp

    def predict(self, x, n_iter=2):
        """advanced...model.py:375-394 + :608-641.  The reference's refinement loop re-encodes
        the ORIGINAL logits each iteration, so the result equals one STN pass."""
        self.eval()
        with torch.no_grad():
            _, pred = self.fast_predict(x)
            if n_iter > 1:
                pred = self.recon_shape(pred.detach().clone())
        return pred

    # -- one cooperative step --------------------------------------------------
    def cooperative_step(self, clean, label, image_cfg=None, seg_cfg=None, noise=None, optimize=True):
        """train...triplet.py:171-237.  `noise` (the 0.05*randn draw) may be supplied so a
        CUDA run and the oracle see the same noisy input.  Returns a dict of the 9 losses."""
        self.train()
        self.zero_grad()
        if noise is None:
            noise = 0.05 * torch.randn_like(clean)
        noisy = torch.clamp(clean + noise, 0, 1)
        s_seg, s_img, s_gt, s_shp = self.standard_training(clean, label, noisy)
        standard = s_seg + s_img + s_shp + s_gt
        self.zero_grad()
        p_img, p_seg = self.hard_example_generation(clean.detach().clone(), label.detach().clone(),
                                                    corrupted_image_DA_config=image_cfg,
                                                    corrupted_seg_DA_config=seg_cfg)
        h_seg, h_img, h_shp, h_pshp = self.hard_example_training(p_img, clean, p_seg, label)
        hard = h_seg + h_img + h_shp + h_pshp
        loss = standard + hard
        self.zero_grad()
        loss.backward()
        if optimize:
            for opt in self.optimizers.values():
                opt.step()
        return {"loss": loss.detach(), "standard/seg": s_seg.detach(), "standard/image": s_img.detach(),
                "standard/shape": s_shp.detach(), "standard/gt_shape": s_gt.detach(),
                "hard/seg": h_seg.detach(), "hard/image": h_img.detach(), "hard/shape": h_shp.detach(),
                "hard/perturbed_shape": h_pshp.detach(), "perturbed_image": p_img.detach(),
                "perturbed_seg": p_seg.detach()}
