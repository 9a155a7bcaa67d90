"""Pose-regressor checkpoints and the initial-pose prediction that seeds a model-based registration.

Host-side mirror of the xvr functions that sit directly in front of the registration hot loop:
``load_model`` (/root/reference/src/xvr/model/network.py:57-78), ``predict_pose`` / ``_resample_xray`` /
``_correct_pose`` / ``_construct_antipode`` (model/inference.py:9-55) and the ``*.pth`` schema written by
``Trainer._checkpoint`` (model/trainer.py:318-332).  None of this is bandwidth-bound: one CNN forward per X-ray.

Checkpoint compatibility: the dictionary keys are the reference's, and for the ResNet family the backbone's
parameter names under torchvision (used here, timm is not installable offline) are the ones timm uses
(``conv1``, ``bn1``, ``layerL.B.convK`` / ``bnK`` / ``downsample.{0,1}``; ``num_classes=0`` leaves no ``fc``
parameters), so a ``model_state_dict`` written by xvr loads as is.  UNPINNED for lack of timm here.
"""

from datetime import datetime

import torch

from .pose import RigidTransform, convert
from .preprocess import XrayTransforms
from .trainer import PoseRegressor
from .utils import resample

__all__ = ["save_checkpoint", "load_model", "predict_pose", "correct_pose", "construct_antipode"]

_REQUIRED_CONFIG = ("model_name", "parameterization", "convention", "norm_layer", "height")


def save_checkpoint(path, model, optimizer, scheduler, itr, model_number, config):
    """Write ``path`` with the keys of ``Trainer._checkpoint`` (trainer.py:318-332); returns ``path``."""
    torch.save({"model_state_dict": model.state_dict(), "optimizer_state_dict": optimizer.state_dict(),
                "scheduler_state_dict": scheduler.state_dict(), "itr": itr, "model_number": model_number,
                "date": datetime.now(), "config": dict(config)}, path)
    return path


def load_model(ckptpath, meta=False, device="cuda"):
    """Rebuild the ``PoseRegressor`` a checkpoint describes, in eval mode on ``device`` (the reference: ``.cuda()``).

    Returns ``(model, config)`` or, with ``meta=True``, ``(model, config, date)``.  Like the reference, a checkpoint
    without ``unit_conversion_factor`` gets 1.0 (models trained before the metre -> millimetre switch)."""
    ckpt = torch.load(ckptpath, weights_only=False, map_location="cpu")
    config = ckpt["config"]
    missing = [k for k in _REQUIRED_CONFIG if k not in config]
    if missing:
        raise KeyError(f"checkpoint config lacks {missing}")
    model = PoseRegressor(model_name=config["model_name"], parameterization=config["parameterization"],
                          convention=config["convention"], norm_layer=config["norm_layer"], height=config["height"],
                          unit_conversion_factor=config.get("unit_conversion_factor", 1.0))
    model.load_state_dict(ckpt["model_state_dict"])
    model = model.to(device).eval()
    return (model, config, ckpt.get("date")) if meta else (model, config)


def _resample_xray(img, sdd, delx, dely, x0, y0, config):
    """Bring an X-ray to the intrinsics the model was trained with (inference.py:26-39): its focal length, a pixel
    size that makes the shorter image side span ``config["height"]`` model pixels, principal point at the centre."""
    if delx != dely:
        raise AssertionError("Non-square pixels are not yet supported")
    height, width = img.shape[-2:]
    subsample = min(height, width) / config["height"]
    img = resample(img, sdd, delx, x0, y0, config["sdd"], config["delx"] / subsample, 0, 0)
    return img, height, width


def predict_pose(model, config, img, sdd, delx, dely, x0, y0):
    """(B,1,H,W) X-ray + its intrinsics -> (initial pose, the preprocessed image the model saw)."""
    img, height, width = _resample_xray(img, sdd, delx, dely, x0, y0, config)
    side = min(height, width)
    top, left = int(round((height - side) / 2.0)), int(round((width - side) / 2.0))  # torchvision center_crop
    img = img[..., top:top + side, left:left + side]
    device = next(model.parameters()).device
    img = XrayTransforms(config["height"])(img).to(device)
    with torch.no_grad():
        init_pose = model(img)
    return init_pose, img


def correct_pose(pose, frame=None):
    """Re-express a predicted pose in another CT frame: ``pose.compose(frame)`` with ``frame`` the SE(3) transform
    (RigidTransform or (4,4) matrix) relating the CT to the template the model was trained in; ``None`` is a no-op.
    (The reference derives ``frame`` from an ANTs warp, inference.py:42-48 -- file IO, out of scope.)"""
    if frame is None:
        return pose
    if not isinstance(frame, RigidTransform):
        frame = RigidTransform(torch.as_tensor(frame, dtype=pose.matrix.dtype))
    return pose.compose(frame.to(pose.matrix.device))


def construct_antipode(pose):
    """The view from the opposite side of the patient (inference.py:51-55): negate the first two ZXY Euler angles
    and add pi to the first.  An involution up to 2 pi."""
    rot, xyz = pose.convert("euler_angles", "ZXY")
    flip = torch.tensor([-1.0, -1.0, 1.0], dtype=rot.dtype, device=rot.device)
    turn = torch.tensor([torch.pi, 0.0, 0.0], dtype=rot.dtype, device=rot.device)
    return convert(rot * flip + turn, xyz, parameterization="euler_angles", convention="ZXY")


def _correct_pose(pose, warp, volume=None, invert=False):
    """The reference's private helper with its signature (inference.py:42-48): ``warp`` is ``None`` (no-op) or --
    here -- the SE(3) transform itself (RigidTransform / (4,4) matrix), inverted first when ``invert``.  The
    reference obtains that matrix from an ANTs warp *file* and the CT (``get_4x4(warp, volume, invert)``): file IO
    through antspyx, out of scope, so a path raises instead of being silently ignored."""
    if warp is None:
        return pose
    if isinstance(warp, (str, bytes)) or hasattr(warp, "__fspath__"):
        raise NotImplementedError("reading an ANTs warp file needs antspyx (not part of the hot path); pass the 4x4 "
                                  "transform it encodes instead")
    frame = warp if isinstance(warp, RigidTransform) else RigidTransform(torch.as_tensor(warp, dtype=torch.float32))
    return correct_pose(pose, frame.inverse() if invert else frame)


# the reference's other private name, for scripts that import it
_construct_antipode = construct_antipode
