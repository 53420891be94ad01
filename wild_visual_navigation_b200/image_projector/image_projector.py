"""ImageProjector (reference: wild_visual_navigation/image_projector/image_projector.py:16-200).

Pinhole projection of the robot's footprint polygons into the camera images and their rasterisation into supervision
masks.  Same constructor / ``project`` / ``project_and_render`` / ``resize_image`` / ``scaled_camera_matrix`` surface;
projection, validity test and the scan-line polygon fill run in one kernel (csrc/footprint_kernels.cu) instead of the
reference's chain of kornia tensor expressions, and ``update_supervision_masks`` fuses the
``fmin(supervision_masks, mask * traversability)`` of ``TraversabilityEstimator.add_supervision_node``
(traversability_estimator.py:281-284) into the same pass.  No kornia / liegroups dependency.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

from .. import ops


class ImageProjector:
    def __init__(self, K: torch.Tensor, h, w, new_h: int = None, new_w: int = None):
        """K (B,4,4) camera matrices; h, w: image size (ints or 0-dim tensors); new_h / new_w: size of the resized
        image the masks are rendered for (image_projector.py:17-86)."""
        h, w = int(h), int(w)
        self.K, self.height, self.width = K, h, w
        new_h = h if new_h is None else int(new_h)
        sy = new_h / h
        sx = (new_w / w) if new_w is not None else sy
        self._square = new_w is None or new_w == new_h
        self._new_h, self._new_w = new_h, (new_h if new_w is None else int(new_w))
        sK = K.clone().float()
        if self._square:   # resize the short side + center crop: both focal lengths / centres follow the height scale
            sK[:, 0, 0] = K[:, 1, 1] * sy
            sK[:, 0, 2] = K[:, 1, 2] * sy
            sK[:, 1, 1] = K[:, 1, 1] * sy
            sK[:, 1, 2] = K[:, 1, 2] * sy
        else:
            sK[:, 0, 0] = K[:, 0, 0] * sx
            sK[:, 0, 2] = K[:, 0, 2] * sx
            sK[:, 1, 1] = K[:, 1, 1] * sy
            sK[:, 1, 2] = K[:, 1, 2] * sy
        self._sK = sK.contiguous()
        B = K.shape[0]
        # the attributes callers read off the reference's kornia PinholeCamera
        self.camera = SimpleNamespace(intrinsics=self._sK, camera_matrix=self._sK[:, :3, :3], batch_size=B,
                                      height=torch.tensor([self._new_h], dtype=torch.int32, device=K.device),
                                      width=torch.tensor([self._new_w], dtype=torch.int32, device=K.device),
                                      extrinsics=torch.eye(4, device=K.device).expand(B, 4, 4))
        self.masks = None

    @property
    def scaled_camera_matrix(self):
        return self.camera.intrinsics.clone()[:3, :3]

    def change_device(self, device):
        self.K, self._sK = self.K.to(device), self._sK.to(device)
        c = self.camera
        self.camera = SimpleNamespace(intrinsics=self._sK, camera_matrix=self._sK[:, :3, :3], batch_size=c.batch_size,
                                      height=c.height.to(device), width=c.width.to(device),
                                      extrinsics=c.extrinsics.to(device))

    def project(self, pose_camera_in_world: torch.Tensor, points_W: torch.Tensor):
        """-> projected (B,N,2), valid (B,N) bool, valid_z (B,N) bool (image_projector.py:126-150; points behind the
        camera come back as NaN, which is what project_and_render turns them into anyway)."""
        proj, valid = ops.project_and_render(self._sK, pose_camera_in_world, points_W, None, self._new_h, self._new_w,
                                             render=False)[1:]
        return proj, valid, ~torch.isnan(proj[..., 0])

    def project_and_render(self, pose_camera_in_world: torch.Tensor, points: torch.Tensor, colors: torch.Tensor,
                           image: torch.Tensor = None):
        """-> masks (B,3,H,W) with NaN outside the footprint, image overlay (or None), projected points, valid points
        (image_projector.py:152-197)."""
        self.masks, proj, valid = ops.project_and_render(self._sK, pose_camera_in_world, points, colors, self._new_h,
                                                         self._new_w)
        overlay = image
        if image is not None:   # visualisation only: the same fill drawn onto the camera image
            if image.dim() != 4:
                image = image[None]
            ones = torch.ones(3, device=points.device)
            inside = ~torch.isnan(ops.project_and_render(self._sK, pose_camera_in_world, points, ones, self._new_h,
                                                         self._new_w)[0][:, :1])
            col = colors.to(image.device, image.dtype)
            col = col if col.dim() == 2 else col.expand(image.shape[0], 3)
            overlay = torch.where(inside, col[:, :, None, None], image)
        return self.masks, overlay, proj, valid

    def update_supervision_masks(self, supervision_masks: torch.Tensor, pose_camera_in_world: torch.Tensor,
                                 points: torch.Tensor, colors: torch.Tensor, traversability: torch.Tensor):
        """In place: supervision_masks = fmin(supervision_masks, project_and_render(...)[0] * traversability)
        (traversability_estimator.py:276-284) without materialising the rendered masks."""
        ops.project_and_render(self._sK, pose_camera_in_world, points, colors, self._new_h, self._new_w, render=False,
                               supervision=supervision_masks, traversability=traversability)
        return supervision_masks

    def resize_image(self, image: torch.Tensor):
        """T.Resize(new_h, NEAREST) + CenterCrop(new_h), or T.Resize([new_h, new_w], NEAREST) (image_projector.py:55-59,
        199-200) as one gather.  The feature extractor does not need this: its patch loader folds the same index map in."""
        H, W = image.shape[-2:]
        if self._square:
            s = self._new_h
            if H <= W:
                rh, rw = s, int(s * W / H)
            else:
                rh, rw = int(s * H / W), s
            top, left = int(round((rh - s) / 2.0)), int(round((rw - s) / 2.0))
            oh = ow = s
        else:
            rh, rw, top, left, oh, ow = self._new_h, self._new_w, 0, 0, self._new_h, self._new_w
        dev = image.device
        # F.interpolate(mode="nearest"): src = floor(dst * in / out), computed in float32
        yi = ((torch.arange(top, top + oh, device=dev, dtype=torch.float32) * (H / rh)).floor().long()).clamp_(max=H - 1)
        xi = ((torch.arange(left, left + ow, device=dev, dtype=torch.float32) * (W / rw)).floor().long()).clamp_(max=W - 1)
        return image[..., yi[:, None], xi[None, :]]
