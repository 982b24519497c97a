"""njf_b200 - B200-native volumetric-rendering hot path of neural-jacobian-field.

Host-side mirror (Python/PyTorch) of the reference's ``Model`` / decoder-registry surface;
all rendering arithmetic runs in ``libnjf_b200.so`` (hand-written sm_100a CUDA, C-ABI in
``include/njf_b200.h``).
"""
from . import _lib  # noqa: F401
