"""B200-native (sm_100a) cooperative-training hot path: drop-in for the reference's
`mask_latent_code_*`, `perturb_latent_code`, `hard_example_generation` and the three-pass
cooperative step, backed by libctl_b200.so (C ABI in include/ctl_b200.h)."""
from . import _lib, conv_blocks, inference, losses, metrics, model_util, networks, ops, optim  # noqa: F401
from .inference import GraphedPredictor  # noqa: F401
from .metrics import runningScore  # noqa: F401
from .model_util import (mask_latent_code_channel_wise, mask_latent_code_spatial_wise,  # noqa: F401
                         set_rng_mode)
from .solver import AdvancedTripletReconSegmentationModel  # noqa: F401
from .training import CooperativeTrainer, GraphedCooperativeTrainer, cooperative_step  # noqa: F401

__version__ = "0.2.0"
