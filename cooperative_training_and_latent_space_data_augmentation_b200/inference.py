"""Inference-only FTN + STN refinement on 3-D stacks (BASELINE.json configs[4]).

Restates the reference's evaluation loop -- `TestSegmentationNetwork.evaluate`
(medseg/test_basic_segmentation_solver.py:85-114: a patient's slices in chunks of <= 10, `predict` per chunk,
`predict_a.max(1)[1].cpu().numpy()`, running metric updated patient by patient) over
`AdvancedTripletReconSegmentationModel.predict` (advanced...model.py:375-394) -- as ONE CUDA graph per chunk shape:

    static input [B,1,H,W] -> FTN (BN folded into the conv epilogues) -> logits -> (n_iter - 1) x STN refinement
    -> argmax label map (uint8) + confusion-matrix update, all on the device

so a chunk costs one host->device copy, one graph launch and one device->host copy of the uint8 label map.  The redundant
STN repeats of the reference (it re-encodes the ORIGINAL logits in every iteration, :627-629) are computed once
(solver.slow_refinement).  Replicas only under multi-GPU: stacks are independent, no collective (section 8e).
"""
import torch

from . import fastpath, ops
from .metrics import runningScore


class GraphedPredictor:
    def __init__(self, solver, chunk_shape, n_iter=None, with_metric=True, max_chunk=10):
        """chunk_shape: (B, 1, H, W) with B <= max_chunk, the shape every full chunk has; shorter tail chunks run
        through a second graph captured on demand."""
        self.solver = solver
        self.n_iter = solver.n_iter if n_iter is None else n_iter
        self.max_chunk = max_chunk
        self.metric = runningScore(solver.num_classes) if with_metric else None
        self.affines = {}
        self.stream = torch.cuda.Stream()
        self.copy_stream = torch.cuda.Stream()
        self.pool = torch.cuda.graph_pool_handle()
        self.graphs = {}
        self.chunk_shape = tuple(chunk_shape)
        solver.eval()

    # ------------------------------------------------------------------------------------------------ one chunk
    def _forward(self, image, label):
        logits = self.solver.predict(image, softmax=False, n_iter=self.n_iter)
        if label is not None and self.metric is not None:
            labels = self.metric.update_from_logits(label, logits, want_labels=True)
        else:
            labels = ops.argmax_labels(logits)
        return logits, labels

    def _graph_for(self, shape, with_label):
        key = (tuple(shape), with_label)
        g = self.graphs.get(key)
        if g is not None:
            return g
        dev = next(self.solver.parameters()).device
        st = {"image": torch.zeros(shape, device=dev, dtype=torch.float32),
              "label": torch.zeros((shape[0],) + tuple(shape[2:]), device=dev, dtype=torch.int64) if with_label else None}
        hist0 = self.metric._matrix().clone() if self.metric is not None else None
        with torch.cuda.stream(self.stream), fastpath.frozen_eval_affines(self.affines):
            for _ in range(2):                              # warm-up: packs weights, fills the folded-affine cache
                self._forward(st["image"], st["label"])
            self.stream.synchronize()
            fastpath.prepare_packing()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, pool=self.pool, stream=self.stream):
                st["logits"], st["labels"] = self._forward(st["image"], st["label"])
        if hist0 is not None:
            self.metric._matrix().copy_(hist0)              # the warm-up passes counted zeros against zeros
        st["graph"] = graph
        self.graphs[key] = st
        return st

    def predict_chunk(self, image, label=None, want_logits=False):
        """image: [B,1,H,W] fp32 (pinned host or device), label: optional int64 [B,H,W].  Returns the uint8 label map
        (a static DEVICE tensor, overwritten by the next call with the same shape) and, when asked, the logits."""
        st = self._graph_for(image.shape, label is not None)
        cur = torch.cuda.current_stream()
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            st["image"].copy_(image, non_blocking=True)
            if label is not None:
                st["label"].copy_(label, non_blocking=True)
            st["graph"].replay()
        cur.wait_stream(self.stream)
        return (st["labels"], st["logits"]) if want_logits else st["labels"]

    # ------------------------------------------------------------------------------------------------ one patient
    def predict_stack(self, image, label=None, out=None):
        """A whole stack [S,1,H,W] in chunks of <= max_chunk slices (test_basic_segmentation_solver.py:97-110).
        Returns the uint8 label maps [S,H,W] on the host (pinned `out` is filled asynchronously when given: synchronise
        the current stream before reading it)."""
        S = image.shape[0]
        if out is None:
            out = torch.empty((S,) + tuple(image.shape[2:]), dtype=torch.uint8).pin_memory()
        for a in range(0, S, self.max_chunk):
            b = min(S, a + self.max_chunk)
            labels = self.predict_chunk(image[a:b], None if label is None else label[a:b])
            out[a:b].copy_(labels, non_blocking=True)
        return out

    def scores(self):
        return self.metric.get_scores() if self.metric is not None else None
