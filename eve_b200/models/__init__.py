"""Drop-in replacement of the reference's ``src/models`` package (same module names, class
names, signatures and state_dict keys) running on hand-written sm_100a kernels."""
from .eve import EVE
from .eye_net import EyeNet
from .refine_net import RefineNet

__all__ = ('EVE', 'EyeNet', 'RefineNet')
