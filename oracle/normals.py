"""TEST INFRASTRUCTURE ONLY — torch restatement (CPU or GPU, float32 or float64) of the depth -> normal map of
train_st_normals.py:113-123:

    normals = kornia.geometry.depth.depth_to_normals(depth[None,None], K, normalize_points=False)
    image_normals = (normals - amin) / (amax - amin + 1e-6)

kornia is an un-vendored dependency of the reference with NO pinned version (environment.yml lists none) and it
is not installed here: **parity unpinned** for this term.  Its published algorithm, restated op for op:
  kornia/geometry/depth.py    depth_to_3d:  points_2d = create_meshgrid(H, W, normalized_coordinates=False);
                              unproject_points: xyz = ((u - cx)/fx, (v - cy)/fy, 1) * depth
                              depth_to_normals: gradients = spatial_gradient(xyz); a, b = gradients[:, :, 0], [:, :, 1];
                                                normals = cross(a, b, dim=1); F.normalize(normals, dim=1, p=2)
  kornia/filters/sobel.py     spatial_gradient(mode="sobel", order=1, normalized=True): 3x3 kernels
                              [[-1,0,1],[-2,0,2],[-1,0,1]] and its transpose, divided by 8 (normalize_kernel2d: sum of
                              absolute values), replicate padding, F.conv3d (cross-correlation).
Autograd gives the backward (incl. torch's amin / amax rule: gradient shared evenly among ties)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def depth_to_normals(depth: torch.Tensor, fx, fy, cx, cy) -> torch.Tensor:
    """Unit normals [3,H,W] from depth [H,W]."""
    H, W = depth.shape
    dt, dev = depth.dtype, depth.device
    ys, xs = torch.meshgrid(torch.arange(H, dtype=dt, device=dev), torch.arange(W, dtype=dt, device=dev), indexing="ij")
    x = (xs - cx) / fx
    y = (ys - cy) / fy
    xyz = torch.stack([x, y, torch.ones_like(x)], 0) * depth[None]            # [3,H,W]
    pad = F.pad(xyz[None], (1, 1, 1, 1), mode="replicate")[0]                  # [3,H+2,W+2]
    kx = torch.tensor([[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]], dtype=dt, device=dev) / 8.0
    ky = kx.t().contiguous()
    ker = torch.stack([kx, ky])[:, None]                                       # [2,1,3,3]
    g = F.conv2d(pad[:, None], ker)                                            # [3,2,H,W]
    a, b = g[:, 0], g[:, 1]
    n = torch.cross(a, b, dim=0)
    return F.normalize(n, dim=0, p=2)


def depth_to_normals01(depth: torch.Tensor, fx, fy, cx, cy) -> torch.Tensor:
    n = depth_to_normals(depth, fx, fy, cx, cy)
    mins = torch.amin(n, dim=(0, 1, 2), keepdim=True)
    maxs = torch.amax(n, dim=(0, 1, 2), keepdim=True)
    return (n - mins) / (maxs - mins + 1e-6)
