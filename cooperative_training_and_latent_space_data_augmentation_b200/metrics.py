"""Drop-in for `medseg/common_utils/metrics.py::runningScore` (:12-57) with the confusion matrix kept ON THE DEVICE
(SURVEY.md section 8 row f4).

The reference moves every prediction to the host (`pred.max(1)[1].cpu().numpy()`, advanced...model.py:656-659) and
histograms it with numpy (`_fast_hist`, metrics.py:18-23).  Here `update` / `update_from_logits` launch ONE kernel
(csrc/optim_metrics.cu `ctl_confusion_update`: argmax over classes fused with the n x n histogram) and nothing leaves
the device until `get_scores()` / `confusion_matrix` is read (one 8-value / n*n-value copy).

Same surface: `runningScore(n_classes)`, `.update(label_trues, label_preds)`, `.get_scores()` -> (dict with the
reference's keys, per-class IoU dict), `.reset()`, `.confusion_matrix` (float64 numpy array).
"""
import numpy as np
import torch

from . import ops


class runningScore(object):

    def __init__(self, n_classes, device="cuda"):
        self.n_classes = n_classes
        self.device = torch.device(device)
        self._hist = None

    def _matrix(self):
        if self._hist is None:
            self._hist = torch.zeros((self.n_classes, self.n_classes), dtype=torch.int64, device=self.device)
        return self._hist

    def _dev(self, a):
        if isinstance(a, np.ndarray):
            a = torch.from_numpy(np.ascontiguousarray(a))
        return a.to(self.device, non_blocking=True)

    @property
    def confusion_matrix(self):
        return self._matrix().cpu().numpy().astype(np.float64)

    def update(self, label_trues, label_preds):
        """label_trues / label_preds: integer label maps of one shape (numpy arrays or tensors, any device).  Pixels
        whose true label is outside [0, n_classes) are ignored, as `_fast_hist`'s mask does."""
        lt, lp = self._dev(label_trues), self._dev(label_preds)
        if lt.numel() != lp.numel():
            raise ValueError("label_trues and label_preds differ in size")
        ops.confusion_update(self._matrix(), lt.reshape(1, -1), pred_labels=lp.reshape(1, -1))

    def update_from_logits(self, label_trues, logits, want_labels=False):
        """Same as update(label_trues, logits.max(1)[1]) without materialising the int64 argmax; returns the uint8
        label map when want_labels."""
        return ops.confusion_update(self._matrix(), self._dev(label_trues), logits=logits, want_labels=want_labels)

    def get_scores(self):
        s = ops.confusion_scores(self._matrix()).cpu().numpy()
        cls_iu = dict(zip(range(self.n_classes), s[4:4 + self.n_classes]))
        return {'Overall Acc: \t': s[0],
                'Mean Acc : \t': s[1],
                'FreqW Acc : \t': s[2],
                'Mean IoU : \t': s[3], }, cls_iu

    def reset(self):
        if self._hist is not None:
            self._hist.zero_()
