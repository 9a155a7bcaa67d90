"""Alias package: ``import diffdrr`` resolves to xvr_b200.

Put this directory's parent on ``sys.path`` ahead of any real DiffDRR install
(``sys.path.insert(0, os.path.dirname(xvr_b200.compat.__file__))`` or ``PYTHONPATH=.../xvr_b200/compat``) and the
import statements of xvr's sources (/root/reference/src/xvr/renderer/load.py:1-2, model/loss.py:2,
registrar/base.py:7-11, model/sampler.py:2, ...) bind to the B200 kernels without touching xvr.
"""

import sys

import xvr_b200
from xvr_b200 import data, drr, metrics, pose, registration, renderers, utils, visualization

__version__ = "0.6.0+xvr_b200." + xvr_b200.__version__

for _name, _mod in dict(data=data, drr=drr, metrics=metrics, pose=pose, registration=registration,
                        renderers=renderers, utils=utils, visualization=visualization).items():
    sys.modules[f"{__name__}.{_name}"] = _mod
