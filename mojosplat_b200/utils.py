"""Camera container of the forward path (mirrors the reference's mojosplat/utils.py:5-31)."""
from dataclasses import dataclass

import torch


@dataclass
class Camera:
    """Pinhole camera, world->camera extrinsics.

    Field-for-field compatible with the reference dataclass (utils.py:5-19): positional order
    ``R, T, H, W, fx, fy, cx, cy, near, far`` and the two derived tensors ``view_matrix``
    (4x4 ``[R|T; 0 0 0 1]``, utils.py:27-29) and ``Ks`` (3x3, utils.py:31), both created on
    ``R``'s device/dtype when not supplied.
    """

    R: torch.Tensor  # (3, 3) world-to-camera rotation
    T: torch.Tensor  # (3,) world-to-camera translation
    H: int
    W: int
    fx: float
    fy: float
    cx: float
    cy: float
    near: float = 0.1
    far: float = 100.0
    view_matrix: torch.Tensor = None
    Ks: torch.Tensor = None

    def __post_init__(self):
        if self.view_matrix is None:
            vm = torch.eye(4, device=self.R.device, dtype=self.R.dtype)
            vm[:3, :3] = self.R
            vm[:3, 3] = self.T
            self.view_matrix = vm
        if self.Ks is None:
            self.Ks = torch.tensor(
                [[self.fx, 0, self.cx], [0, self.fy, self.cy], [0, 0, 1]],
                device=self.R.device, dtype=self.R.dtype)

    def to(self, device) -> "Camera":
        """Same camera with its tensors on ``device`` (additive helper, not in the reference)."""
        return Camera(R=self.R.to(device), T=self.T.to(device), H=self.H, W=self.W,
                      fx=self.fx, fy=self.fy, cx=self.cx, cy=self.cy, near=self.near,
                      far=self.far, view_matrix=self.view_matrix.to(device),
                      Ks=self.Ks.to(device))
