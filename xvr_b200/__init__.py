"""xvr_b200: B200-native (sm_100a) DRR rendering and pose-optimisation hot path behind xvr's DiffDRR API.

The package mirrors the slice of ``diffdrr`` that xvr's training and registration loops call
(SURVEY.md section 2.3): ``drr.DRR``, ``pose.{RigidTransform, convert, make_matrix}``,
``renderers.{Trilinear, Siddon}``, ``metrics.*``, ``registration.Registration``,
``data.transform_hu_to_density``.  All bandwidth-bound work runs in hand-written CUDA kernels reached through
the C-ABI of ``include/xvr_b200.h`` (``libxvr_b200.so``); there is no CPU or PyTorch fallback.
"""

from . import data, drr, metrics, pose, registration, renderers, utils, visualization  # noqa: F401
from .data import read, transform_hu_to_density  # noqa: F401
from .drr import DRR, Detector  # noqa: F401
from .pose import RigidTransform, convert, make_matrix  # noqa: F401
from .registration import N_ANGULAR_COMPONENTS, Registration  # noqa: F401
from .renderers import Siddon, Trilinear  # noqa: F401

__version__ = "0.1.0"
