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
        self.drr, self.fiducials = drr, fiducials
        # eps = 0: identical poses must score exactly zero (the loss keeps an eps for its square root's gradient)
        self._pose_distance = DoubleGeodesicSE3(drr.detector.sdd, eps=0.0)

    @staticmethod
    def _mean_distance(a, b):
        """Mean Euclidean distance between corresponding points, (B, n, d) x (B, n, d) -> (B,)."""
        return torch.linalg.vector_norm(a - b, dim=-1).mean(dim=-1)

    def __call__(self, true_pose, pred_pose):
        poses = {"true": true_pose, "pred": pred_pose}
        on_detector = {k: self.drr.perspective_projection(p, self.fiducials) for k, p in poses.items()}
        lifted = {k: self.drr.inverse_projection(p, on_detector[k]) for k, p in poses.items()}
        moved = {k: p(self.fiducials) for k, p in poses.items()}
        errors = [
            self.drr.detector.delx * self._mean_distance(on_detector["pred"], on_detector["true"]),  # mPE
            self._mean_distance(lifted["pred"], lifted["true"]),                                     # mRPE
            self._mean_distance(moved["pred"], moved["true"]),                                       # mTRE
            self._pose_distance(true_pose, pred_pose)[2],                                            # dGeo
        ]
        return torch.stack(errors, dim=-1).squeeze().cpu().tolist()
