"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's evaluation metric and of torch.optim.Adam's update.

    fast_hist / RunningScoreOracle  <- medseg/common_utils/metrics.py:12-57 (runningScore)
    adam_step                       <- torch.optim.Adam (the reference's pinned dependency, requirements.txt:2;
                                       amsgrad=False, maximize=False, weight_decay=0 as advanced...model.py:778 builds it)

Pinned against the unmodified reference by tests/golden/metrics_scores.npz (oracle/make_golden.py gen_metrics) and,
for Adam, against torch.optim.Adam itself in tests/test_optim_metrics_cpu.py.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline leg may import this module.
"""
import numpy as np


def fast_hist(label_true, label_pred, n_class):
    """metrics.py:18-23"""
    label_true = np.asarray(label_true).reshape(-1)
    label_pred = np.asarray(label_pred).reshape(-1)
    mask = (label_true >= 0) & (label_true < n_class)
    return np.bincount(n_class * label_true[mask].astype(int) + label_pred[mask].astype(int),
                       minlength=n_class ** 2).reshape(n_class, n_class)


class RunningScoreOracle:
    def __init__(self, n_classes):
        self.n_classes = n_classes
        self.confusion_matrix = np.zeros((n_classes, n_classes))

    def update(self, label_trues, label_preds):
        """metrics.py:25-28"""
        for lt, lp in zip(label_trues, label_preds):
            self.confusion_matrix += fast_hist(lt.flatten(), lp.flatten(), self.n_classes)

    def get_scores(self):
        """metrics.py:30-53"""
        hist = self.confusion_matrix
        with np.errstate(divide="ignore", invalid="ignore"):
            acc = np.diag(hist).sum() / hist.sum()
            acc_cls = np.nanmean(np.diag(hist) / hist.sum(axis=1))
            iu = np.diag(hist) / (hist.sum(axis=1) + hist.sum(axis=0) - np.diag(hist))
            mean_iu = np.nanmean(iu)
            freq = hist.sum(axis=1) / hist.sum()
            fwavacc = (freq[freq > 0] * iu[freq > 0]).sum()
        return np.array([acc, acc_cls, fwavacc, mean_iu]), iu


def adam_step(p, g, m, v, t, lr=1e-4, b1=0.9, b2=0.999, eps=1e-8, weight_decay=0.0, grad_scale=1.0):
    """One Adam update in float64 on numpy arrays; t = step number after the increment (>= 1).  Returns (p, m, v)."""
    p, g, m, v = (np.asarray(a, np.float64) for a in (p, g, m, v))
    g = g * grad_scale + weight_decay * p
    m = m + (g - m) * (1 - b1)
    v = b2 * v + (1 - b2) * g * g
    step_size = lr / (1 - b1 ** t)
    denom = np.sqrt(v) / np.sqrt(1 - b2 ** t) + eps
    return p - step_size * m / denom, m, v
