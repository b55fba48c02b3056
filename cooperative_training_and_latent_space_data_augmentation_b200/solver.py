"""Drop-in for the hot-path surface of the reference's
`medseg/models/advanced_triplet_recon_segmentation_model.py::AdvancedTripletReconSegmentationModel`
(class at :24).  Method names, argument order, defaults and return tuples follow the reference:

    get_network :76           init_model :157          encode_image :191     decode_image :218
    encode_shape :228         decode_shape :245        recon_shape :259      perturb_latent_code :300
    predict :375              decoder_inference :396   standard_training :414
    hard_example_generation :469   hard_example_training :525   fast_predict :561
    slow_refinement :608      save_model :666          save_snapshots :680   load_snapshots :703
    train/eval :740-753       optimizer helpers :755-797

What is different underneath (B200-first, same results):
  * latent masking / dropout run in libctl_b200.so (K1/K2), see model_util.py and ops.py
  * no gc.collect() / torch.cuda.empty_cache() on the step path (those force
    allocator flushes every iteration), no host syncs inside the losses
  * CUDA only: use_gpu=False raises -- there is no CPU path in this build
Evaluation/plotting helpers (running metric, PNG dumps) are out of scope.
"""
import os
import random
from os.path import join

import torch
import torch.nn as nn

from . import conv_blocks, model_util, ops
from .losses import basic_loss_fn, construct_input, half_mse_loss
from .model_util import (_disable_tracking_bn_stats, makeVariable, mask_latent_code_channel_wise,
                         mask_latent_code_spatial_wise, set_grad)
from .optim import FlatAdam
from .networks import Dual_Branch_Encoder, MyDecoder, MyEncoder, init_weights_kaiming


class _ChannelDropout(torch.autograd.Function):
    """F.dropout2d replacement backed by ctl_channel_dropout; backward = grad * noise."""

    @staticmethod
    def forward(ctx, latent, p, keep, rng):
        out, mask, keep_used = ops.channel_dropout(latent.detach(), p, keep=keep, rng=rng, want_mask=True,
                                                   want_keep=True, step_params=model_util.step_params())
        N, C = latent.shape[:2]
        noise = (keep_used * ops.dropout_scale(p)).view(N, C, 1, 1).to(latent.dtype)
        ctx.save_for_backward(noise)
        ctx.mark_non_differentiable(mask)
        return out, mask

    @staticmethod
    def backward(ctx, grad_out, _grad_mask):
        (noise,) = ctx.saved_tensors
        return grad_out * noise, None, None, None


class AdvancedTripletReconSegmentationModel(nn.Module):
    def __init__(self, network_type='FCN_16_standard', image_ch=1,
                 learning_rate=1e-4,
                 encoder_dropout=None,
                 decoder_dropout=None,
                 num_classes=4, n_iter=1,
                 checkpoint_dir=None, use_gpu=True, debug=False
                 ):
        super().__init__()
        if not use_gpu:
            raise RuntimeError("this build is CUDA (sm_100a) only: use_gpu=False is not supported, there is no "
                               "CPU fallback")
        self.network_type = network_type
        self.image_ch = image_ch
        self.checkpoint_dir = checkpoint_dir
        self.num_classes = num_classes
        self.encoder_dropout = encoder_dropout
        self.decoder_dropout = decoder_dropout
        self.learning_rate = learning_rate
        self.n_iter = n_iter
        self.use_gpu = use_gpu
        self.debug = debug
        self.model = self.get_network(checkpoint_dir=checkpoint_dir)
        self.optimizers = None
        self.reset_all_optimizers()
        self.latent_code = {'image': None, 'segmentation': None, 'shape': None}
        self.z_i = None
        self.z_s = None
        self.training = True
        self.loss = 0.

    # ------------------------------------------------------------------ construction
    def get_network(self, checkpoint_dir=None):
        network_type = self.network_type
        if network_type not in ['FCN_16_standard', 'FCN_16_standard_w_o_filter', 'FCN_16_standard_share_code']:
            raise NotImplementedError
        r = 4   # '16' in the name -> reduce_factor 4: widths 16/32/64/128, 128-channel latent at H/16
        nets = {
            'image_encoder': Dual_Branch_Encoder(input_channel=self.image_ch, z_level_1_channel=512 // r,
                                                 z_level_2_channel=512 // r, feature_reduce=r,
                                                 encoder_dropout=self.encoder_dropout, norm=nn.BatchNorm2d),
            'segmentation_decoder': MyDecoder(input_channel=512 // r, up_type='NN', output_channel=self.num_classes,
                                              feature_reduce=r, decoder_dropout=self.decoder_dropout,
                                              norm=nn.BatchNorm2d),
            'shape_encoder': MyEncoder(input_channel=self.num_classes, output_channel=512 // r, feature_reduce=r,
                                       encoder_dropout=self.encoder_dropout, norm=nn.BatchNorm2d, act=nn.ReLU()),
            'shape_decoder': MyDecoder(input_channel=512 // r, up_type='NN', output_channel=self.num_classes,
                                       feature_reduce=r, decoder_dropout=self.decoder_dropout, norm=nn.BatchNorm2d),
            'image_decoder': MyDecoder(input_channel=512 // r, up_type='Conv2', output_channel=self.image_ch,
                                       feature_reduce=r, decoder_dropout=self.decoder_dropout, norm=nn.BatchNorm2d,
                                       last_act=nn.Sigmoid()),
        }
        for name in nets:
            path = None
            if checkpoint_dir is not None and not checkpoint_dir == "":
                path = join(checkpoint_dir, name + '.pth')
            nets[name] = self.init_model(nets[name], resume_path=path).to('cuda')
        return nets

    def init_model(self, model, resume_path=None):
        if resume_path is None:
            return init_weights_kaiming(model)
        if resume_path == '':
            return model
        assert os.path.exists(resume_path), 'path: {} must exist'.format(resume_path)
        state = torch.load(resume_path, map_location='cpu')
        try:
            model.load_state_dict(state)
        except Exception:  # noqa: BLE001  -- historical checkpoints wrap the dict (advanced...model.py:166-169)
            model.load_state_dict(state['model_state'], strict=False)
        return model

    def parameters(self):
        for module in self.model.values():
            yield from module.parameters()

    def named_parameters(self):
        for name, module in self.model.items():
            for k, p in module.named_parameters():
                yield name + '.' + k, p

    # ------------------------------------------------------------------ forward pieces
    def run(self, input):
        zi, zs = self.encode_image(input)
        recon_image = self.decode_image(zi)
        init_predict = self.model['segmentation_decoder'](zs)
        return recon_image, init_predict, self.recon_shape(init_predict)

    def _maybe_untracked(self, module, disable_track_bn_stats, fn):
        if disable_track_bn_stats:
            with _disable_tracking_bn_stats(module):
                return fn()
        return fn()

    def encode_image(self, input, disable_track_bn_stats=False):
        encoder = self.model['image_encoder']
        z_i, z_s = self._maybe_untracked(encoder, disable_track_bn_stats, lambda: encoder(input))
        if 'share_code' in self.network_type:
            z_i = z_s
        elif 'w_o_filter' in self.network_type:
            z_s = z_i
        self.latent_code['image'], self.latent_code['segmentation'] = z_i, z_s
        return z_i, z_s

    def decode_segmentation_from_image_code(self, latent_code_i, disable_track_bn_stats=False):
        encoder, decoder = self.model['image_encoder'], self.model['segmentation_decoder']
        z_s = self._maybe_untracked(encoder, disable_track_bn_stats, lambda: encoder.filter_code(latent_code_i))
        return self._maybe_untracked(decoder, disable_track_bn_stats, lambda: decoder(z_s))

    def decode_image(self, latent_code, disable_track_bn_stats=False):
        dec = self.model['image_decoder']
        return self._maybe_untracked(dec, disable_track_bn_stats, lambda: dec(latent_code))

    def encode_shape(self, segmentation, is_label_map=False, disable_track_bn_stats=False, temperature=2):
        enc = self.model['shape_encoder']

        def run():
            # kernel routes fuse construct_input (softmax(logit/T) | one-hot) into the first STN convolution
            code = enc.forward_from_segmentation(segmentation, is_label_map=is_label_map, temperature=temperature)
            if code is not None:
                return code
            prediction_map = construct_input(segmentation, image=None, num_classes=self.num_classes,
                                             apply_softmax=not is_label_map, is_labelmap=is_label_map,
                                             temperature=temperature, use_gpu=self.use_gpu, smooth_label=False)
            return enc(prediction_map)

        shape_code = self._maybe_untracked(enc, disable_track_bn_stats, run)
        self.latent_code['shape'] = shape_code
        return shape_code

    def decode_shape(self, latent_code, disable_track_bn_stats=False):
        dec = self.model['shape_decoder']
        return self._maybe_untracked(dec, disable_track_bn_stats, lambda: dec(latent_code))

    def recon_shape(self, segmentation_logit, is_label_map=False, disable_track_bn_stats=False):
        return self.decode_shape(self.encode_shape(segmentation_logit, is_label_map, disable_track_bn_stats),
                                 disable_track_bn_stats)

    def recon_image(self, image, disable_track_bn_stats=False):
        z_i, _ = self.encode_image(image, disable_track_bn_stats=disable_track_bn_stats)
        return self.decode_image(z_i, disable_track_bn_stats=disable_track_bn_stats)

    def forward(self, input):
        _, predict = self.fast_predict(input)
        return predict

    def fast_predict(self, input, disable_track_bn_stats=False):
        encoder, decoder = self.model['image_encoder'], self.model['segmentation_decoder']

        def both():
            z_i, z_s = encoder(input)
            if 'share_code' in self.network_type:
                z_i = z_s
            elif 'w_o_filter' in self.network_type:
                z_s = z_i
            return z_i, z_s

        if not self.training:
            with torch.no_grad():
                z_i, z_s = both()
                y_0 = decoder(z_s)
        else:
            z_i, z_s = self._maybe_untracked(encoder, disable_track_bn_stats, both)
            y_0 = self._maybe_untracked(decoder, disable_track_bn_stats, lambda: decoder(z_s))
        return (z_i, z_s), y_0

    # ------------------------------------------------------------------ latent-space data augmentation
    def perturb_latent_code(self, latent_code, decoder_function, label_y=None,
                            perturb_type='random', threshold=0.5,
                            if_soft=False, random_threshold=False,
                            loss_type='mse', if_detach=False):
        assert perturb_type in ['random', 'dropout',
                                'spatial', 'channel'], 'invalid method name'
        if perturb_type == 'random':
            candidates = ['dropout', 'spatial', 'channel']
            random.shuffle(candidates)                  # python's global generator, as in the reference
            perturb_type = candidates[0]
        self.last_perturb_type = perturb_type

        if perturb_type == 'dropout':
            if not latent_code.is_cuda:
                raise RuntimeError("CUDA only: latent code is on %s" % latent_code.device)
            N, C = latent_code.shape[:2]
            keep = rng = None
            if threshold <= 0.0:                        # feature_dropout returns its input: nothing is drawn
                keep = torch.ones((N, C), device=latent_code.device)
            elif threshold >= 1.0:                      # input * zeros: nothing is drawn either
                keep = torch.zeros((N, C), device=latent_code.device)
            elif model_util.get_rng_mode() == "torch":  # the [N,C,1,1] Bernoulli(1-p) draw of feature_dropout
                keep = torch.empty((N, C, 1, 1), device=latent_code.device,
                                   dtype=latent_code.dtype).bernoulli_(1 - threshold).view(N, C)
            else:
                rng = model_util.native_rng()
            masked_latent_code, mask = _ChannelDropout.apply(latent_code, float(threshold), keep, rng)
        else:
            assert loss_type in ['mse', 'ce', 'corr'], 'not implemented loss'
            fn = mask_latent_code_spatial_wise if perturb_type == 'spatial' else mask_latent_code_channel_wise
            masked_latent_code, mask = fn(latent_code, num_classes=self.num_classes,
                                          decoder_function=decoder_function, label=label_y, percentile=threshold,
                                          random=random_threshold, loss_type=loss_type, if_detach=if_detach,
                                          if_soft=if_soft)
        if if_detach:
            masked_latent_code = masked_latent_code.detach()    # already a fresh buffer: no clone needed
        return masked_latent_code, mask

    def decoder_inference(self, decoder, latent_code, eval=False, disable_track_bn_stats=False):
        decoder_state = decoder.training
        if eval:
            decoder.eval()
            with torch.no_grad():
                logit = decoder(latent_code)
        else:
            logit = self._maybe_untracked(decoder, disable_track_bn_stats, lambda: decoder(latent_code))
        decoder.train(mode=decoder_state)
        return logit

    def hard_example_generation(self,
                                clean_image_l,
                                label_l,
                                gen_corrupted_seg=True,
                                gen_corrupted_image=True,
                                corrupted_image_DA_config={"loss_name": "mse",
                                                           "mask_type": "random",
                                                           "max_threshold": 0.5,
                                                           "random_threshold": True,
                                                           "if_soft": True},
                                corrupted_seg_DA_config={"loss_name": "ce",
                                                         "mask_type": "random",
                                                         "max_threshold": 0.5,
                                                         "random_threshold": True,
                                                         "if_soft": True}):
        seg_dec, img_dec = self.model['segmentation_decoder'], self.model['image_decoder']
        set_grad(seg_dec, requires_grad=False)
        set_grad(img_dec, requires_grad=False)
        perturbed_image_0, perturbed_y_0 = None, None
        try:
            if gen_corrupted_image:
                self.reset_all_optimizers()
                cfg = corrupted_image_DA_config
                z, _ = self.perturb_latent_code(latent_code=self.z_i, label_y=clean_image_l,
                                                perturb_type=cfg["mask_type"], decoder_function=img_dec,
                                                loss_type=cfg["loss_name"], threshold=cfg["max_threshold"],
                                                random_threshold=cfg["random_threshold"], if_detach=True,
                                                if_soft=cfg["if_soft"])
                perturbed_image_0 = self.decoder_inference(decoder=img_dec, latent_code=z, eval=False,
                                                           disable_track_bn_stats=True)
            if gen_corrupted_seg:
                self.reset_all_optimizers()
                cfg = corrupted_seg_DA_config
                z, _ = self.perturb_latent_code(latent_code=self.z_s, label_y=label_l,
                                                perturb_type=cfg["mask_type"], decoder_function=seg_dec,
                                                loss_type=cfg["loss_name"], threshold=cfg["max_threshold"],
                                                random_threshold=cfg["random_threshold"], if_detach=True,
                                                if_soft=cfg["if_soft"])
                perturbed_y_0 = self.decoder_inference(decoder=seg_dec, latent_code=z, eval=False,
                                                       disable_track_bn_stats=True)
        finally:
            set_grad(seg_dec, requires_grad=True)
            set_grad(img_dec, requires_grad=True)
        return perturbed_image_0, perturbed_y_0

    # ------------------------------------------------------------------ training passes
    def standard_training(self, clean_image_l, label_l, perturbed_image, separate_training=False,
                          compute_gt_recon=True, update_latent=True, disable_track_bn_stats=False):
        zero = torch.zeros((), device=clean_image_l.device)     # a fill kernel: legal under stream capture
        (z_i, z_s), y_0 = self.fast_predict(perturbed_image, disable_track_bn_stats=disable_track_bn_stats)
        if update_latent:
            self.z_i, self.z_s = z_i, z_s
        standard_supervised_loss = basic_loss_fn(pred=y_0, target=label_l.detach(), loss_type='cross entropy')
        image_recon = self.decode_image(z_i)
        image_recon_loss = half_mse_loss(image_recon, clean_image_l)      # 0.5 * MSELoss (advanced...model.py:443-447)
        if compute_gt_recon:
            gt_recon = self.recon_shape(label_l.detach(), is_label_map=True)
            gt_shape_recon_loss = basic_loss_fn(pred=gt_recon, target=label_l, loss_type='cross entropy')
        else:
            gt_shape_recon_loss = zero
        y_0_new = y_0.detach() if separate_training else y_0
        p_recon = self.recon_shape(y_0_new, is_label_map=False, disable_track_bn_stats=disable_track_bn_stats)
        pred_shape_recon_loss = basic_loss_fn(pred=p_recon, target=label_l, loss_type='cross entropy')
        return standard_supervised_loss, image_recon_loss, gt_shape_recon_loss, pred_shape_recon_loss

    def hard_example_training(self, perturbed_image, clean_image_l, perturbed_seg, label_l, separate_training=False,
                              use_gpu=True):
        zero = torch.zeros((), device=clean_image_l.device)     # a fill kernel: legal under stream capture
        seg_loss, recon_loss, shape_loss, perturbed_p_recon_loss = zero, zero, zero, zero
        if perturbed_image is not None:
            seg_loss, recon_loss, _, shape_loss = self.standard_training(
                clean_image_l=clean_image_l, label_l=label_l, perturbed_image=perturbed_image.detach(),
                compute_gt_recon=False, separate_training=separate_training, update_latent=False,
                disable_track_bn_stats=True)
        if perturbed_seg is not None:
            if separate_training:
                perturbed_seg = perturbed_seg.detach()
            perturbed_p_recon = self.recon_shape(perturbed_seg, is_label_map=False, disable_track_bn_stats=True)
            perturbed_p_recon_loss = basic_loss_fn(pred=perturbed_p_recon, target=label_l, loss_type='cross entropy')
        return seg_loss, recon_loss, shape_loss, perturbed_p_recon_loss

    # ------------------------------------------------------------------ inference
    def predict(self, input, softmax=False, n_iter=None):
        self.eval()
        n_iter = self.n_iter if n_iter is None else n_iter
        with torch.no_grad():
            _, pred = self.fast_predict(input)
            for _ in range(max(0, n_iter - 1)):
                pred, _ = self.slow_refinement(pred_logit=pred, n_steps=n_iter, save_internal_predicts=False)
        if softmax:
            pred = torch.softmax(pred, dim=1)
        return pred

    def slow_refinement(self, pred_logit, n_steps=1, auto_stop=False, save_internal_predicts=False):
        """The reference re-encodes the ORIGINAL logits in every iteration (advanced...model.py:627-629), so
        all n_steps passes are identical; one STN pass is computed and reused unless auto_stop / the internal
        trace need the per-step values (which are then that same tensor)."""
        if n_steps is None:
            n_steps = self.n_iter
        internal_predicts = {0: [pred_logit]}
        s_t = pred_logit
        if n_steps >= 1:
            refined = self.recon_shape(pred_logit.detach())
            for i in range(n_steps):
                prev, s_t = s_t, refined
                stop = bool(auto_stop and torch.sqrt(torch.mean((prev - s_t) ** 2)) < 1e-4)
                if stop:
                    s_t = prev
                if save_internal_predicts:
                    internal_predicts[i] = [s_t]        # recorded before the break, as the reference does (:636-640)
                if stop:
                    break
        return s_t, internal_predicts

    # ------------------------------------------------------------------ state
    def train(self, if_testing=False):
        self.training = True
        for v in self.model.values():
            if not if_testing:
                v.train()
                set_grad(v, requires_grad=True)
            else:
                v.eval()

    def eval(self):
        """Modules to eval mode, no-grad inference.  Documented deviation: the reference's eval() calls
        train(if_testing=True), which sets `self.training` back to True (advanced...model.py:740-753), so a
        fast_predict() issued after eval() but OUTSIDE predict() keeps autograd enabled on eval-mode modules there.
        Here `self.training` stays False and fast_predict runs under torch.no_grad(): eval-mode BatchNorm with
        autograd enabled is not on the hot path and has no kernel route (predict / evaluate, which wrap everything
        in no_grad in the reference too, are unaffected)."""
        self.training = False
        self.train(if_testing=True)
        self.training = False

    # ------------------------------------------------------------------ evaluation metric (on the device)
    def set_running_metric(self):
        from .metrics import runningScore
        return runningScore(n_classes=self.num_classes)

    def evaluate(self, input, targets_npy, n_iter=None):
        """advanced...model.py:643-664: predict, arg-max, running-metric update -- the arg-max and the confusion
        matrix stay on the device (metrics.runningScore); `targets_npy` may be a numpy array or a tensor [N,H,W].
        Returns the logits; `self.cur_eval_predicts` holds the uint8 label map (device)."""
        if getattr(self, 'running_metric', None) is None:
            self.running_metric = self.set_running_metric()
        n_iter = self.n_iter if n_iter is None else n_iter
        self.train(if_testing=True)
        pred = self.predict(input, n_iter=n_iter)
        self.cur_eval_predicts = self.running_metric.update_from_logits(targets_npy, pred, want_labels=True)
        self.cur_eval_images, self.cur_eval_gts = input, targets_npy
        return pred

    def reset_all_optimizers(self):
        if self.optimizers is None:
            self.set_optimizers()
        self.flat_adam.zero_grad()                  # one memset of the flat gradient buffer (all five networks)

    def get_optimizer(self, model_name=None):
        assert self.optimizers, 'please set optimizers first before fetching'
        return self.optimizers if model_name is None else self.optimizers[model_name]

    def set_optimizers(self, capturable=None):
        """One Adam optimizer per sub-network (advanced...model.py:774-785), all five backed by ONE multi-tensor kernel
        over flat parameter / gradient / moment buffers (optim.FlatAdam).  `solver.optimizers[name]` keeps the
        torch.optim surface the reference uses (step / zero_grad / state_dict / load_state_dict / param_groups).
        Step counters live on the device, so `optimize_all_params` can always be recorded into a CUDA graph
        (`capturable` is accepted for compatibility and ignored).  Calling it again keeps the existing optimizer
        state (moments, step counts): a trainer built after load_snapshots resumes where the checkpoint stopped."""
        assert self.model
        flat = getattr(self, 'flat_adam', None)
        if flat is None:
            flat = self.flat_adam = FlatAdam(self.model, lr=self.learning_rate)
        elif not flat.attached():
            flat.reattach()
        self.optimizers = {name: flat.view(name) for name in self.model}

    def optimize_all_params(self):
        self.flat_adam.step()                       # one launch for the five networks

    def optimize_params(self, model_name):
        self.optimizers[model_name].step()

    def reset_optimizer(self, model_name):
        self.optimizers[model_name].zero_grad()

    def save_model(self, save_dir, epoch_iter, model_prefix=None, save_optimizers=False):
        epoch_path = join(save_dir, *[str(epoch_iter), 'checkpoints'])
        os.makedirs(epoch_path, exist_ok=True)
        for model_name, model in self.model.items():
            torch.save(model.state_dict(), join(epoch_path, '{}.pth'.format(model_name)))
        if save_optimizers:
            for model_name, optimizer in self.optimizers.items():
                torch.save(optimizer.state_dict(), join(epoch_path, '{}_optim.pth'.format(model_name)))

    def save_snapshots(self, save_dir, epoch, model_prefix='interrupted'):
        epoch_path = join(save_dir, *['interrupted', 'checkpoints'])
        os.makedirs(epoch_path, exist_ok=True)
        save_path = join(epoch_path, self.network_type + '.pkl')
        torch.save({'network_type': self.network_type, 'epoch': epoch,
                    'model_state': {k: m.state_dict() for k, m in self.model.items()},
                    'optimizer_state': {k: o.state_dict() for k, o in self.optimizers.items()}}, save_path)
        return save_path

    def load_snapshots(self, file_path):
        if file_path is None or file_path == '' or not os.path.exists(file_path):
            return 0
        checkpoint = torch.load(file_path, map_location='cuda')
        for k, v in self.model.items():
            v.load_state_dict(checkpoint['model_state'][k])
        for k, v in self.optimizers.items():
            v.load_state_dict(checkpoint['optimizer_state'][k])
        return checkpoint['epoch']
