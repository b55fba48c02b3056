"""Torch / cuDNN restatements of the conv blocks used as parity yardsticks and as the eager-GPU baseline (not product)."""
from .torch_modes import install  # noqa: F401
