"""Scene-side pieces of the hot path: cameras with the reference's matrix conventions, a seeded
synthetic scene generator, and the slice of GaussianModel the style-transfer
step touches (scene/gaussian_model.py:26-41 activations, :124-147 create_from_pcd incl. the
distCUDA2 scale initialisation, :149-167 training_setup).  Everything else of the reference's
scene/ package (dataset readers, densification, PLY I/O) is out of scope.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
from torch import nn

from .sh import RGB2SH


# ----------------------------------------------------------------------------- cameras
def world_to_view(R: np.ndarray, t: np.ndarray) -> np.ndarray:
    """getWorld2View2 with zero translate / unit scale (utils/graphics_utils.py:38-49)."""
    Rt = np.zeros((4, 4), dtype=np.float64)
    Rt[:3, :3] = R.T
    Rt[:3, 3] = t
    Rt[3, 3] = 1.0
    return Rt.astype(np.float32)


def projection_matrix(znear: float, zfar: float, fovX: float, fovY: float) -> torch.Tensor:
    """getProjectionMatrix (utils/graphics_utils.py:51-71)."""
    ty, tx = math.tan(fovY / 2), math.tan(fovX / 2)
    top, right = ty * znear, tx * znear
    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * znear / (2 * right)
    P[1, 1] = 2.0 * znear / (2 * top)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


class Camera:
    """The attributes render() reads from scene/cameras.py:17-59 (no image payload)."""

    def __init__(self, R, T, FoVx, FoVy, width, height, device="cuda", uid=0):
        self.uid = uid
        self.R, self.T = np.asarray(R, dtype=np.float64), np.asarray(T, dtype=np.float64)
        self.FoVx, self.FoVy = float(FoVx), float(FoVy)
        self.image_width, self.image_height = int(width), int(height)
        self.znear, self.zfar = 0.01, 100.0
        wv = torch.tensor(world_to_view(self.R, self.T)).transpose(0, 1)
        proj = projection_matrix(self.znear, self.zfar, self.FoVx, self.FoVy).transpose(0, 1)
        full = wv.unsqueeze(0).bmm(proj.unsqueeze(0)).squeeze(0)
        center = wv.inverse()[3, :3]
        self.world_view_transform = wv.contiguous().to(device)
        self.projection_matrix = proj.contiguous().to(device)
        self.full_proj_transform = full.contiguous().to(device)
        self.camera_center = center.contiguous().to(device)

    def to(self, device):
        c = Camera.__new__(Camera)
        c.__dict__.update(self.__dict__)
        for k in ("world_view_transform", "projection_matrix", "full_proj_transform", "camera_center"):
            setattr(c, k, getattr(self, k).to(device))
        return c


def look_at(eye, target=(0.0, 0.0, 0.0), up=(0.0, 0.0, 1.0)):
    """(R, T) in the reference's convention: R is camera-to-world rotation (columns = camera
    axes, +z forward, +y down like COLMAP), T = world-to-camera translation."""
    eye, target, up = (np.asarray(v, dtype=np.float64) for v in (eye, target, up))
    f = target - eye
    f /= np.linalg.norm(f)
    r = np.cross(f, up)
    if np.linalg.norm(r) < 1e-8:
        r = np.cross(f, np.array([1.0, 0.0, 0.0]))
    r /= np.linalg.norm(r)
    d = np.cross(f, r)  # camera +y (down)
    R = np.stack([r, d, f], axis=1)
    T = -R.T @ eye
    return R, T


def fov_y_from_x(fovx: float, width: int, height: int) -> float:
    focal = width / (2 * math.tan(fovx / 2))
    return 2 * math.atan(height / (2 * focal))


def orbit_cameras(n: int, radius: float, height: float, fovx: float, width: int, img_h: int,
                  device="cuda", sphere=False):
    cams = []
    for k in range(n):
        a = 2 * math.pi * k / n
        if sphere:  # poses on a sphere of `radius`, alternating elevation
            el = math.radians(20.0 + 25.0 * (k % 3))
            eye = (radius * math.cos(el) * math.cos(a), radius * math.cos(el) * math.sin(a),
                   radius * math.sin(el))
        else:
            eye = (radius * math.cos(a), radius * math.sin(a), height)
        R, T = look_at(eye)
        cams.append(Camera(R, T, fovx, fov_y_from_x(fovx, width, img_h), width, img_h, device, uid=k))
    return cams


# ----------------------------------------------------------------------------- scenes
@dataclass
class SceneSpec:
    name: str
    P: int
    width: int
    height: int
    fovx: float
    cam_radius: float
    cam_height: float
    sphere_cams: bool
    log_scale_mu: float
    garden: bool


CONFIGS = {
    # BASELINE.json configs[1]: lego-like, 300k Gaussians, 800x800, SH degree 3
    "c2": SceneSpec("c2", 300_000, 800, 800, 0.6911, 4.03, 0.0, True, -4.6, False),
    # configs[2]: garden-scale, 3M Gaussians, 1297x840
    "c3": SceneSpec("c3", 3_000_000, 1297, 840, 1.19, 6.0, 2.0, False, -4.0, True),
    # configs[4]: depth/normal-loss variant, 1920x1080, 3M
    "c5": SceneSpec("c5", 3_000_000, 1920, 1080, 1.19, 6.0, 2.0, False, -4.0, True),
}


def synthetic_gaussians(P: int, seed: int = 0, garden: bool = False, log_scale_mu: float = -4.6,
                        sh_degree: int = 3) -> dict:
    """Seeded raw (pre-activation) Gaussian parameters as float32 numpy arrays.

    xyz: mixture of 64 anisotropic blobs on a shell (object scenes, extent ~ +-1.3) or a ground
    disc of radius 8 plus a central object (garden scenes); log-scales N(mu, 0.8^2) clipped to
    [-9, -1.7]; quaternions N(0, I) (un-normalised, normalised by the model like the reference);
    opacity logits N(0, 2^2); f_dc N(0, 1); f_rest N(0, 0.1^2).
    """
    rng = np.random.default_rng(seed)
    K = 64
    centers = rng.normal(size=(K, 3))
    centers /= np.linalg.norm(centers, axis=1, keepdims=True)
    centers *= rng.uniform(0.6, 1.0, size=(K, 1))
    axes = rng.uniform(0.03, 0.25, size=(K, 3))
    which = rng.integers(0, K, size=P)
    xyz = centers[which] + rng.normal(size=(P, 3)) * axes[which]
    xyz = np.clip(xyz, -1.3, 1.3)
    if garden:
        n_ground = int(P * 0.6)
        r = 8.0 * np.sqrt(rng.uniform(size=n_ground))
        th = rng.uniform(0, 2 * np.pi, size=n_ground)
        ground = np.stack([r * np.cos(th), r * np.sin(th), rng.normal(scale=0.05, size=n_ground) - 1.0], 1)
        xyz = xyz * 1.5
        xyz[:n_ground] = ground
        xyz = xyz[rng.permutation(P)]
    M = (sh_degree + 1) ** 2
    out = {
        "xyz": xyz.astype(np.float32),
        "log_scales": np.clip(rng.normal(log_scale_mu, 0.8, size=(P, 3)), -9.0, -1.7).astype(np.float32),
        "rotations": rng.normal(size=(P, 4)).astype(np.float32),
        "opacity_logits": rng.normal(0.0, 2.0, size=(P, 1)).astype(np.float32),
        "f_dc": rng.normal(0.0, 1.0, size=(P, 1, 3)).astype(np.float32),
        "f_rest": rng.normal(0.0, 0.1, size=(P, M - 1, 3)).astype(np.float32),
    }
    return out


def scene_cameras(spec: SceneSpec, n: int = 8, device="cuda"):
    return orbit_cameras(n, spec.cam_radius, spec.cam_height, spec.fovx, spec.width, spec.height,
                         device, sphere=spec.sphere_cams)


# ----------------------------------------------------------------------------- model
class PipelineParams:
    """arguments/__init__.py PipelineParams: the three switches render() reads."""

    def __init__(self, convert_SHs_python=False, compute_cov3D_python=False, debug=False,
                 fused_activations=True):
        self.convert_SHs_python = convert_SHs_python
        self.compute_cov3D_python = compute_cov3D_python
        self.debug = debug
        # extension: fold the GaussianModel activations into the kernels (model_render.py)
        self.fused_activations = fused_activations


@dataclass
class OptimizationParams:
    """Learning rates of arguments/__init__.py OptimizationParams used by training_setup."""
    position_lr_init: float = 0.00016
    position_lr_final: float = 0.0000016
    position_lr_delay_mult: float = 0.01
    position_lr_max_steps: int = 30_000
    feature_lr: float = 0.0025
    opacity_lr: float = 0.05
    scaling_lr: float = 0.005
    rotation_lr: float = 0.001
    percent_dense: float = 0.01


def inverse_sigmoid(x):
    return torch.log(x / (1 - x))


def build_rotation(r: torch.Tensor) -> torch.Tensor:
    """utils/general_utils.py:78-100 (normalises the quaternion first)."""
    q = r / r.norm(dim=1, keepdim=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], dim=1)
    return R.view(-1, 3, 3)


def covariance_from_scaling_rotation(scaling, scaling_modifier, rotation) -> torch.Tensor:
    """scene/gaussian_model.py:27-31: L = R S, Sigma = L L^T, six upper-triangular entries."""
    L = build_rotation(rotation) * (scaling_modifier * scaling).unsqueeze(1)
    S = L @ L.transpose(1, 2)
    return torch.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], 1)


class GaussianModel:
    """Parameter container with the reference's names, activations and Adam groups."""

    def __init__(self, sh_degree: int):
        self.active_sh_degree = 0
        self.max_sh_degree = sh_degree
        e = torch.empty(0)
        self._xyz = self._features_dc = self._features_rest = e
        self._scaling = self._rotation = self._opacity = e
        self.optimizer = None
        self.spatial_lr_scale = 0.0

    # getters (scene/gaussian_model.py:95-120)
    @property
    def get_scaling(self):
        return torch.exp(self._scaling)

    @property
    def get_rotation(self):
        return torch.nn.functional.normalize(self._rotation)

    @property
    def get_xyz(self):
        return self._xyz

    @property
    def get_features(self):
        # a view-parallel optimizer may still be exchanging the features on its side stream
        # (peer.PeerShardedAdam late_params): order the current stream behind it before reading them
        sync = getattr(getattr(self, "optimizer", None), "sync", None)
        if callable(sync):
            sync()
        return torch.cat((self._features_dc, self._features_rest), dim=1)

    @property
    def get_opacity(self):
        return torch.sigmoid(self._opacity)

    def get_covariance(self, scaling_modifier=1):
        return covariance_from_scaling_rotation(self.get_scaling, scaling_modifier, self._rotation)

    def oneupSHdegree(self):
        if self.active_sh_degree < self.max_sh_degree:
            self.active_sh_degree += 1

    def parameters(self):
        return [self._xyz, self._features_dc, self._features_rest, self._opacity, self._scaling,
                self._rotation]

    @classmethod
    def from_arrays(cls, arrs: dict, sh_degree: int = 3, device="cuda", requires_grad=True):
        m = cls(sh_degree)
        m.active_sh_degree = sh_degree
        t = lambda k: nn.Parameter(torch.as_tensor(arrs[k]).float().to(device).contiguous(),
                                   requires_grad=requires_grad)
        m._xyz, m._features_dc, m._features_rest = t("xyz"), t("f_dc"), t("f_rest")
        m._scaling, m._rotation, m._opacity = t("log_scales"), t("rotations"), t("opacity_logits")
        return m

    def create_from_pcd(self, points, colors, spatial_lr_scale: float, device="cuda"):
        """scene/gaussian_model.py:124-147; `points`/`colors` are [P,3] arrays (BasicPointCloud)."""
        from .simple_knn._C import distCUDA2

        self.spatial_lr_scale = spatial_lr_scale
        pts = torch.as_tensor(np.asarray(points)).float().to(device)
        fused_color = RGB2SH(torch.as_tensor(np.asarray(colors)).float().to(device))
        K = (self.max_sh_degree + 1) ** 2
        features = torch.zeros((pts.shape[0], 3, K), dtype=torch.float32, device=device)
        features[:, :3, 0] = fused_color
        dist2 = torch.clamp_min(distCUDA2(pts), 0.0000001)
        scales = torch.log(torch.sqrt(dist2))[..., None].repeat(1, 3)
        rots = torch.zeros((pts.shape[0], 4), device=device)
        rots[:, 0] = 1
        opacities = inverse_sigmoid(0.1 * torch.ones((pts.shape[0], 1), dtype=torch.float, device=device))
        self._xyz = nn.Parameter(pts.requires_grad_(True))
        self._features_dc = nn.Parameter(features[:, :, 0:1].transpose(1, 2).contiguous().requires_grad_(True))
        self._features_rest = nn.Parameter(features[:, :, 1:].transpose(1, 2).contiguous().requires_grad_(True))
        self._scaling = nn.Parameter(scales.requires_grad_(True))
        self._rotation = nn.Parameter(rots.requires_grad_(True))
        self._opacity = nn.Parameter(opacities.requires_grad_(True))

    def training_setup(self, training_args: OptimizationParams = OptimizationParams(), fused: bool = False,
                       peer: bool = False, group=None, average: bool = True, in_backward: bool = False,
                       overlap_features: bool = False, feature_records: bool = False):
        """Adam with the reference's six groups (scene/gaussian_model.py:154-163).
        `fused=True` swaps torch.optim.Adam for the library's fused Adam kernel (same maths);
        `peer=True` for the view-parallel single-kernel optimizer over NVLink peer memory
        (peer.PeerShardedAdam; parameters move into its arena, render() writes gradients there);
        with `overlap_features=True` the SH features (81% of the bytes) are exchanged on a side stream
        while the next render() projects, sorts and bins — render() orders its colour kernel behind them;
        `in_backward=True` (single GPU) applies the update inside the rasteriser's backward kernel
        (optim.BackwardFusedAdam): no gradient tensors, optimizer.step() launches nothing."""
        a = training_args
        groups = [
            {"params": [self._xyz], "lr": a.position_lr_init * self.spatial_lr_scale, "name": "xyz"},
            {"params": [self._features_dc], "lr": a.feature_lr, "name": "f_dc"},
            {"params": [self._features_rest], "lr": a.feature_lr / 20.0, "name": "f_rest"},
            {"params": [self._opacity], "lr": a.opacity_lr, "name": "opacity"},
            {"params": [self._scaling], "lr": a.scaling_lr, "name": "scaling"},
            {"params": [self._rotation], "lr": a.rotation_lr, "name": "rotation"},
        ]
        self.grad_sink = None
        if peer and feature_records:
            # optional (peer_records.py; measured equal to the feature exchange, profiles/r02_scaling.md): features updated
            # from per-view colour records, never exchanged
            from .peer_records import PeerRecordAdam
            self.optimizer = PeerRecordAdam(groups, self._xyz, self._features_dc, self._features_rest,
                                            lambda: self.active_sh_degree, lr=0.0, eps=1e-15, group=group,
                                            average=average)
            self.grad_sink = self.optimizer.grad_sink
        elif peer:
            from .peer import PeerShardedAdam
            late = [self._features_dc, self._features_rest] if overlap_features else None
            self.optimizer = PeerShardedAdam(groups, lr=0.0, eps=1e-15, group=group, average=average,
                                             late_params=late)
            self.grad_sink = self.optimizer.grad_sink
        elif in_backward:
            from .optim import BackwardFusedAdam
            self.optimizer = BackwardFusedAdam(groups, lr=0.0, eps=1e-15)
            self.grad_sink = self.optimizer.grad_sink
        elif fused:
            from .optim import FusedAdam
            self.optimizer = FusedAdam(groups, lr=0.0, eps=1e-15)
        else:
            self.optimizer = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
        self._xyz_lr = (a.position_lr_init * self.spatial_lr_scale, a.position_lr_final * self.spatial_lr_scale,
                        a.position_lr_delay_mult, a.position_lr_max_steps)
        return self.optimizer

    def update_learning_rate(self, iteration):
        """Per-step learning rate of the positions (scene/gaussian_model.py:169-176 with the schedule of
        utils/general_utils.py:29-62): log-linear interpolation from position_lr_init to position_lr_final over
        position_lr_max_steps (no delay phase, as the reference configures it: lr_delay_steps = 0).  Returns the rate."""
        lr0, lr1, _delay_mult, max_steps = self._xyz_lr
        if iteration < 0 or (lr0 == 0.0 and lr1 == 0.0):
            lr = 0.0
        else:
            t = min(max(iteration / max_steps, 0.0), 1.0)
            lr = float(math.exp(math.log(lr0) * (1.0 - t) + math.log(lr1) * t))
        for group in self.optimizer.param_groups:
            if group["name"] == "xyz":
                group["lr"] = lr
                return lr
