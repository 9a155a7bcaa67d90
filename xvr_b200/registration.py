"""``Registration``: pose parameters as ``nn.Parameter``s in front of a DRR (``diffdrr.registration``).

Pinned by /root/reference/src/xvr/registrar/base.py:168-169,224-225,249: ``Registration(drr, rot, xyz,
parameterization, convention)``, ``.rotation`` / ``.translation`` handed to Adam, ``.pose``, ``.drr`` and
``reg() -> drr(reg.pose)``.
"""

import torch

from .pose import N_ANGULAR_COMPONENTS, convert

__all__ = ["Registration", "N_ANGULAR_COMPONENTS"]


class Registration(torch.nn.Module):
    def __init__(self, drr, rotation, translation, parameterization, convention=None):
        super().__init__()
        self.drr = drr
        self.rotation = torch.nn.Parameter(rotation.detach().clone())
        self.translation = torch.nn.Parameter(translation.detach().clone())
        self.parameterization = parameterization
        self.convention = convention

    @property
    def pose(self):
        return convert(self.rotation, self.translation, parameterization=self.parameterization,
                       convention=self.convention)

    def forward(self, **kwargs):
        # same image as drr(self.pose); handing over the parameters lets DRR.forward take its one-launch
        # Euler -> camera path (csrc/regstep.cu) instead of the convert/compose chain
        return self.drr(self.rotation, self.translation, parameterization=self.parameterization,
                        convention=self.convention, **kwargs)
