"""2D/3D registration error metrics (all in mm): the host-side mirror of xvr's ``Evaluator``
(/root/reference/src/xvr/metrics/evaluator.py:7-43, driven by scripts/evaluate.py and the registrars' logging).

Pure pose algebra on a handful of fiducials -- O(B * n_fiducials) -- so it stays in PyTorch; it only needs the
``DRR`` surface rebuilt in ``xvr_b200.drr`` (``perspective_projection``, ``inverse_projection``, ``detector``).
"""

import torch

from .metrics import DoubleGeodesicSE3

__all__ = ["Evaluator"]


class Evaluator:
    """``Evaluator(drr, fiducials)(true_pose, pred_pose) -> [mPE, mRPE, mTRE, dGeo]`` (per pose when B > 1).

    * mPE  -- mean projection error: fiducials projected with both poses, distance on the detector (pixels * delx)
    * mRPE -- mean reprojection error: those projections lifted back onto each detector plane, distance in 3-D
    * mTRE -- mean target registration error: fiducials moved by both poses, distance in 3-D
    * dGeo -- double geodesic distance between the poses (DoubleGeodesicSE3 with eps = 0)
    """

    def __init__(self, drr, fiducials):
        self.drr = drr
        self.fiducials = fiducials
        self.geodesic = DoubleGeodesicSE3(drr.detector.sdd, eps=0.0)

    def __call__(self, true_pose, pred_pose):
        x = self.drr.perspective_projection(pred_pose, self.fiducials)
        y = self.drr.perspective_projection(true_pose, self.fiducials)
        mpe = (self.drr.detector.delx * (x - y)).norm(dim=-1).mean(dim=-1)

        x = self.drr.inverse_projection(pred_pose, x)
        y = self.drr.inverse_projection(true_pose, y)
        mrpe = (x - y).norm(dim=-1).mean(dim=-1)

        x = pred_pose(self.fiducials)
        y = true_pose(self.fiducials)
        mtre = (x - y).norm(dim=-1).mean(dim=-1)

        *_, dgeo = self.geodesic(true_pose, pred_pose)
        return torch.stack([mpe, mrpe, mtre, dgeo], dim=-1).squeeze().cpu().tolist()
