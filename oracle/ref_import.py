"""TEST INFRASTRUCTURE ONLY -- imports the upstream reference (read-only, /root/reference).

Only usable in the build container: the GPU box has no /root/reference.  Used by
`oracle/make_golden.py` (fixture generation) and by the `not gpu` tests that pin the
oracle restatement against the real reference when it is present.

The reference's hot path is pure torch, but its modules import non-arithmetic packages at
module top (SimpleITK, medpy, IPython, matplotlib, skimage, seaborn, scipy.misc and the
numpy-1.x-only `numpy.lib.function_base`; see medseg/common_utils/basic_operations.py:5,16,
metrics.py:5,7, save.py:2-13).  We register empty stub modules for those names before import;
none of them is touched by the functions on the path.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CTL_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "medseg"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave like a package so `import a.b` works
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent:
        setattr(_stub(parent), child, m)
    return m


def install_stubs():
    import numpy as np
    if "numpy.lib.function_base" not in sys.modules:
        fb = types.ModuleType("numpy.lib.function_base")
        fb.copy = np.copy
        sys.modules["numpy.lib.function_base"] = fb
    _stub("SimpleITK", sitkLinear=1, sitkNearestNeighbor=0)
    _stub("medpy")
    _stub("medpy.metric")
    _stub("medpy.metric.binary", dc=lambda a, b: 0.0)
    _stub("IPython")
    _stub("IPython.display", display=lambda *a, **k: None, HTML=lambda *a, **k: None)
    try:
        import matplotlib  # noqa: F401
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        _stub("matplotlib", use=lambda *a, **k: None)
        _stub("matplotlib.pyplot")
    _stub("skimage")
    _stub("skimage.transform", resize=lambda *a, **k: None)
    _stub("seaborn")
    try:
        import scipy.misc  # noqa: F401
    except Exception:
        _stub("scipy.misc")


def import_reference():
    """Returns (model_util module, solver class) of the unmodified reference."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import medseg.models.model_util as ref_model_util
    from medseg.models.advanced_triplet_recon_segmentation_model import (
        AdvancedTripletReconSegmentationModel as RefSolver)
    return ref_model_util, RefSolver
