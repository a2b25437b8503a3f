"""Synthetic DTU-shaped scenes and weights (no dataset is available offline).

Everything here is seeded CPU-generator randomness so that the golden fixtures
(tests/golden, produced by oracle/make_golden.py from the *reference* code), the
parity tests and bench.py all see bit-identical inputs without shipping megabytes
of weights.  The recipe follows SURVEY.md section 8(d):

  * weights: the reference's geometric initialisation
    (implicit_differentiable_renderer.py:53-68) restated on an explicit
    torch.Generator, optionally perturbed so the surface is not a perfect sphere
    and the positional-encoding columns carry signal;
  * cameras: a ring of radius 3 around the origin at +-20 degrees elevation
    looking at the origin, DTU-like focal length f = 2892.33 * (W / 1600);
  * MVS cameras: [2,4,4] = [world->cam extrinsic ; K] (my_utils.py:98-110 convention)
    with size = 2, center = 0 so that world == normalised coordinates;
  * feature maps: box-filtered Gaussian noise, [views, 32, H/2, W/2] (NCHW like
    scene_dataset.py:149).
"""
from __future__ import annotations

import hashlib
import math
from typing import Dict, List, Optional

import numpy as np
import torch

PE_POINTS = 6   # multires      (confs/mvsdf_dtu.conf:29)
PE_VIEWS = 4    # multires_view (confs/mvsdf_dtu.conf:38)
FEATURE_SIZE = 256


def sdf_layer_dims(width: int, n_hidden: int = 8, feature_size: int = FEATURE_SIZE) -> List[int]:
    d0 = 3 + 6 * PE_POINTS
    return [d0] + [width] * n_hidden + [1 + 1 + feature_size]


def render_layer_dims(width: int, n_hidden: int = 4, feature_size: int = FEATURE_SIZE) -> List[int]:
    d0 = 9 + feature_size + (3 + 6 * PE_VIEWS - 3)
    return [d0] + [width] * n_hidden + [3]


def make_state_dict(width: int = 512, render_width: Optional[int] = None, seed: int = 0,
                    perturb: float = 0.0, pe_noise: float = 0.0, bias: float = 0.6,
                    skip_in=(4,)) -> Dict[str, torch.Tensor]:
    """state_dict with the reference's key names
    (implicit_network.lin{l}.{bias,weight_g,weight_v}, rendering_network.lin{l}.*).

    perturb  : relative N(0, (perturb*std)^2) noise on weight_v of SDF lin1..lin7 and a
               +-perturb relative jitter of weight_g (so that weight-norm folding matters);
    pe_noise : absolute N(0, pe_noise^2) noise on the positional-encoding columns of lin0
               and the skip columns of lin4 (zero under geometric init).
    """
    g = torch.Generator().manual_seed(seed)
    rw = width if render_width is None else render_width
    sd: Dict[str, torch.Tensor] = {}
    dims = sdf_layer_dims(width)
    n_lin = len(dims) - 1
    for l in range(n_lin):
        out_dim = dims[l + 1] - dims[0] if (l + 1) in skip_in else dims[l + 1]
        in_dim = dims[l]
        if l == n_lin - 1:
            w = torch.empty(out_dim, in_dim).normal_(math.sqrt(math.pi) / math.sqrt(in_dim), 1e-4, generator=g)
            b = torch.full((out_dim,), -bias)
        elif l == 0:
            w = torch.zeros(out_dim, in_dim)
            w[:, :3] = torch.empty(out_dim, 3).normal_(0.0, math.sqrt(2) / math.sqrt(out_dim), generator=g)
            b = torch.zeros(out_dim)
        elif l in skip_in:
            w = torch.empty(out_dim, in_dim).normal_(0.0, math.sqrt(2) / math.sqrt(out_dim), generator=g)
            w[:, -(dims[0] - 3):] = 0.0
            b = torch.zeros(out_dim)
        else:
            w = torch.empty(out_dim, in_dim).normal_(0.0, math.sqrt(2) / math.sqrt(out_dim), generator=g)
            b = torch.zeros(out_dim)
        sd[f"implicit_network.lin{l}.bias"] = b
        sd[f"implicit_network.lin{l}.weight_v"] = w
    rdims = render_layer_dims(rw)
    for l in range(len(rdims) - 1):
        bound = 1.0 / math.sqrt(rdims[l])
        sd[f"rendering_network.lin{l}.weight_v"] = (torch.rand(rdims[l + 1], rdims[l], generator=g) * 2 - 1) * bound
        sd[f"rendering_network.lin{l}.bias"] = (torch.rand(rdims[l + 1], generator=g) * 2 - 1) * bound

    if perturb > 0 or pe_noise > 0:
        gp = torch.Generator().manual_seed(seed + 1000003)
        for l in range(n_lin - 1):
            w = sd[f"implicit_network.lin{l}.weight_v"]
            if perturb > 0 and l >= 1:
                w.add_(torch.randn(w.shape, generator=gp) * (perturb * float(w.std())))
            if pe_noise > 0 and l == 0:
                w[:, 3:].add_(torch.randn(w[:, 3:].shape, generator=gp) * pe_noise)
            if pe_noise > 0 and l in skip_in:
                k = dims[0] - 3
                w[:, -k:].add_(torch.randn(w[:, -k:].shape, generator=gp) * pe_noise)
        if perturb > 0:
            for l in range(len(rdims) - 1):
                b = sd[f"rendering_network.lin{l}.bias"]
                b.add_(torch.randn(b.shape, generator=gp) * 0.05)

    # weight_norm(dim=0): g = ||v|| per output row so that the folded weight equals v ...
    for key in [k for k in sd if k.endswith("weight_v")]:
        v = sd[key]
        gk = key[:-1] + "g"
        sd[gk] = v.norm(dim=1, keepdim=True).clone()
    # ... then jitter g so that folding is actually exercised
    if perturb > 0:
        gp2 = torch.Generator().manual_seed(seed + 2000003)
        for key in [k for k in sd if k.endswith("weight_g")]:
            if key.startswith("implicit_network.lin8"):
                continue
            sd[key].mul_(1.0 + perturb * 0.25 * (torch.rand(sd[key].shape, generator=gp2) * 2 - 1))
    return {k: sd[k].contiguous() for k in sorted(sd)}


def state_dict_checksum(sd: Dict[str, torch.Tensor]) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()[:16]


def _look_at_pose(cam_pos: np.ndarray) -> np.ndarray:
    """cam->world 4x4 with the camera looking at the origin (+z forward, +y down)."""
    fwd = -cam_pos / np.linalg.norm(cam_pos)
    up_world = np.array([0.0, 0.0, 1.0])
    right = np.cross(fwd, up_world)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    pose = np.eye(4, dtype=np.float64)
    pose[:3, 0] = right
    pose[:3, 1] = down
    pose[:3, 2] = fwd
    pose[:3, 3] = cam_pos
    return pose


def make_cameras(n_views: int, H: int, W: int, radius: float = 3.0, seed: int = 0):
    """Returns pose [n,4,4] (cam->world), intrinsics [n,4,4], mvs_cam [n,2,4,4] as float32 tensors."""
    f = 2892.33 * (W / 1600.0)
    poses, intr, cams = [], [], []
    for i in range(n_views):
        az = 2 * math.pi * i / max(n_views, 1) * 0.35 + 0.3 + 0.01 * seed   # views bunch on one side like an MVS rig
        el = math.radians(20.0) * (1 if i % 2 == 0 else -1)
        c = radius * np.array([math.cos(az) * math.cos(el), math.sin(az) * math.cos(el), math.sin(el)])
        pose = _look_at_pose(c)
        K = np.eye(4, dtype=np.float64)
        K[0, 0] = f
        K[1, 1] = f
        K[0, 2] = W / 2.0
        K[1, 2] = H / 2.0
        cam = np.zeros((2, 4, 4), dtype=np.float64)
        cam[0] = np.linalg.inv(pose)
        cam[1] = np.eye(4)
        cam[1, :3, :3] = K[:3, :3]
        poses.append(pose)
        intr.append(K)
        cams.append(cam)
    t = lambda a: torch.from_numpy(np.stack(a).astype(np.float32))
    return t(poses), t(intr), t(cams)


def make_feature_maps(pose: torch.Tensor, intrinsics: torch.Tensor, h: int, w: int, scale: int = 2,
                      channels: int = 32, seed: int = 2, radius: float = 0.6, noise: float = 0.15) -> torch.Tensor:
    """View-consistent synthetic CNN features, [views, C, h, w] (NCHW like scene_dataset.py:149).

    Each pixel of the (1/scale)-resolution map looks along its camera ray, hits an analytic
    sphere of radius `radius` (closest approach when it misses) and evaluates a smooth 3-D
    field F(x) = sin(x A + phi) there, plus box-filtered noise.  Projections of one surface
    point into different views therefore see correlated features, which is what makes the
    consistency loss (loss.py:115-165) non-trivial."""
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(3, channels, generator=g) * 2.0
    phi = torch.rand(channels, generator=g) * 6.2831853
    n_views = pose.shape[0]
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    px = (xs + 0.5) * scale          # pixel centre in full-resolution image coordinates
    py = (ys + 0.5) * scale
    out = torch.empty(n_views, channels, h, w)
    kbox = torch.ones(channels, 1, 5, 5) / 25.0
    for v in range(n_views):
        K = intrinsics[v]
        d_cam = torch.stack([(px - K[0, 2]) / K[0, 0], (py - K[1, 2]) / K[1, 1], torch.ones_like(px)], dim=-1)
        d = d_cam @ pose[v, :3, :3].T
        d = d / d.norm(dim=-1, keepdim=True)
        c = pose[v, :3, 3]
        b = (d * c).sum(-1)
        disc = b * b - (c.dot(c) - radius * radius)
        t = torch.where(disc > 0, -b - torch.sqrt(disc.clamp_min(0)), -b)
        x = c + t.unsqueeze(-1) * d
        f = torch.sin(x @ A + phi)                                     # [h,w,C]
        nz = torch.randn(1, channels, h, w, generator=g)
        nz = torch.nn.functional.conv2d(nz, kbox, padding=2, groups=channels)[0] * (5.0 * noise)
        out[v] = f.permute(2, 0, 1) + nz
    return out.contiguous()


def make_depth_maps(pose: torch.Tensor, intrinsics: torch.Tensor, h: int, w: int, scale: int = 2,
                    radius: float = 0.6):
    """Synthetic MVS depth maps of the analytic sphere (no RNG): z-depth at the (1/scale)-resolution pixel centres,
    0 where the ray misses, and the matching depth cameras [n,1,2,4,4] = (world->cam extrinsic, K / scale) -- the layout
    of ground_truth['depths'] / ['depth_cams'] (datasets/scene_dataset.py:105-116, :205-206)."""
    n = pose.shape[0]
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float64), torch.arange(w, dtype=torch.float64), indexing="ij")
    depths = torch.zeros(n, 1, 1, h, w, dtype=torch.float32)
    cams = torch.zeros(n, 1, 2, 4, 4, dtype=torch.float32)
    for i in range(n):
        K = intrinsics[i].double().clone()
        K[:2, :3] /= scale
        P = pose[i].double()
        d_cam = torch.stack([(xs + 0.5 - K[0, 2]) / K[0, 0], (ys + 0.5 - K[1, 2]) / K[1, 1], torch.ones_like(xs)], dim=-1)
        d_w = d_cam @ P[:3, :3].T
        c = P[:3, 3]
        a = (d_w * d_w).sum(-1)
        b = (d_w * c).sum(-1)
        disc = b * b - a * (c.dot(c) - radius * radius)
        z = torch.where(disc > 0, (-b - torch.sqrt(disc.clamp_min(0))) / a, torch.zeros_like(a))
        depths[i, 0, 0] = z.float()
        cams[i, 0, 0] = torch.linalg.inv(P).float()
        cams[i, 0, 1] = torch.eye(4)
        cams[i, 0, 1, :3, :3] = K[:3, :3].float()
    return depths, cams


def make_scene(H: int, W: int, n_images: int = 1, n_src: int = 1, n_rays: Optional[int] = None,
               seed: int = 0, mask_mode: str = "ones") -> Dict[str, torch.Tensor]:
    """A mini-batch in the layout IDRNetwork.forward / IDRLoss.forward consume
    (SURVEY.md section 8(a) rows a12 and a15). n_rays=None -> full pixel grid."""
    n_views = n_images + n_src
    pose, intr, cams = make_cameras(n_views, H, W, seed=seed)
    h, w = H // 2, W // 2
    feats = make_feature_maps(pose, intr, h, w, seed=seed + 2)
    vv, uu = np.mgrid[0:H, 0:W]
    uv_full = torch.from_numpy(np.stack([uu.reshape(-1), vv.reshape(-1)], axis=1).astype(np.float32))
    g = torch.Generator().manual_seed(seed + 7)
    if n_rays is None or n_rays >= H * W:
        sel = torch.arange(H * W)
    else:
        sel = torch.randperm(H * W, generator=g)[:n_rays]
    uv = uv_full[sel]
    N = uv.shape[0]
    B = n_images
    if mask_mode == "ones":
        obj = torch.ones(B, N, dtype=torch.bool)
    else:
        cx, cy = W / 2.0, H / 2.0
        rad = 0.42 * min(H, W)
        obj = (((uv[:, 0] - cx) ** 2 + (uv[:, 1] - cy) ** 2) < rad * rad).unsqueeze(0).repeat(B, 1)
    src_idx = [[(i + 1 + s) % n_views for s in range(n_src)] for i in range(B)]
    scene = {
        "uv": uv.unsqueeze(0).repeat(B, 1, 1).contiguous(),
        "pose": pose[:B].contiguous(),
        "intrinsics": intr[:B].contiguous(),
        "object_mask": obj.contiguous(),
        "cam": cams[:B].contiguous(),
        "src_cams": torch.stack([cams[idx] for idx in src_idx]).contiguous(),
        "feat": feats[:B].contiguous(),
        "feat_src": torch.stack([feats[idx] for idx in src_idx]).contiguous(),
        "size": torch.full((B,), 2.0),
        "center": torch.zeros(B, 3),
        "rgb": (torch.rand(B, N, 3, generator=g) * 2 - 1).contiguous(),
    }
    scene["depths"], scene["depth_cams"] = make_depth_maps(pose[:B], intr[:B], h, w)
    return scene
