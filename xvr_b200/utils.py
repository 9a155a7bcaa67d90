"""``diffdrr.utils.resample`` -- import compatibility for xvr (model/inference.py:37, ``_resample_xray``: bring a real
X-ray to the intrinsics the pose regressor was trained with).  One call per registration, outside the hot path
(SURVEY.md 2.3 marks it OUT OF SCOPE), so it is plain PyTorch.

UNPINNED: restated from the documented behaviour of DiffDRR's function (shift the principal point, crop to change
the focal length, pad to change the pixel size, resize back to the input shape each time); DiffDRR builds it from
kornia's ``translate`` / ``center_crop`` / ``resize``, here ``affine_grid`` / slicing / ``interpolate`` stand in.
"""

import torch
import torch.nn.functional as F

__all__ = ["resample"]


def resample(img, focal_len, delx, x0=0.0, y0=0.0, new_focal_len=None, new_delx=None, new_x0=None, new_y0=None):
    """Resample ``img`` (B,C,H,W), taken with (focal_len, delx, x0, y0), to the new intrinsics; same output shape."""
    new_focal_len = focal_len if new_focal_len is None else new_focal_len
    new_delx = delx if new_delx is None else new_delx
    new_x0 = x0 if new_x0 is None else new_x0
    new_y0 = y0 if new_y0 is None else new_y0
    x = img.clone()
    B, _, height, width = x.shape

    # translate by the change of principal point, in pixels (bilinear, zero padding)
    tx, ty = (new_x0 - x0) / delx, (new_y0 - y0) / delx
    if tx != 0.0 or ty != 0.0:
        theta = torch.tensor([[1.0, 0.0, -2.0 * tx / width], [0.0, 1.0, -2.0 * ty / height]]).to(x)
        grid = F.affine_grid(theta[None].expand(B, -1, -1), x.shape, align_corners=False)
        x = F.grid_sample(x, grid, mode="bilinear", padding_mode="zeros", align_corners=False)

    # a longer focal length sees a smaller field of view: centre-crop, then resize back
    focal_scaling = new_focal_len / focal_len
    ch, cw = int(height / focal_scaling), int(width / focal_scaling)
    if (ch, cw) != (height, width):
        if ch <= height and cw <= width:
            top, left = (height - ch) // 2, (width - cw) // 2
            x = x[..., top:top + ch, left:left + cw]
        else:
            ph, pw = (ch - height) // 2, (cw - width) // 2
            x = F.pad(x, (pw, cw - width - pw, ph, ch - height - ph))
        x = F.interpolate(x, size=(height, width), mode="bilinear", align_corners=False, antialias=True)

    # larger pixels cover more of the scene: pad, then resize back
    pixel_scaling = new_delx / delx
    ph, pw = int(height * (pixel_scaling - 1) / 2), int(width * (pixel_scaling - 1) / 2)
    if ph != 0 or pw != 0:
        if ph >= 0 and pw >= 0:
            x = F.pad(x, (pw, pw, ph, ph))
        else:
            x = x[..., -ph:height + ph, -pw:width + pw]
        x = F.interpolate(x, size=(height, width), mode="bilinear", align_corners=False, antialias=True)
    return x
