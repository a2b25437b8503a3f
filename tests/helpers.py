"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np
import torch

from mvsdf_b200 import synth

WEIGHT_PRESETS = {
    "w256": dict(width=256, seed=1, perturb=0.05, pe_noise=0.003, bias=0.6),
    "w256_geo": dict(width=256, seed=1, perturb=0.0, pe_noise=0.0, bias=0.6),
    "w512": dict(width=512, seed=0, perturb=0.05, pe_noise=0.003, bias=0.75),
}


# stress presets for the tracer's prefilter: stronger perturbations than any golden fixture uses
STRESS_PRESETS = {
    "w256_p10": dict(width=256, seed=1, perturb=0.1, pe_noise=0.003, bias=0.6),
    "w256_pe8": dict(width=256, seed=1, perturb=0.05, pe_noise=0.008, bias=0.6),      # (pe_noise 0.02 destroys the surface)
    "w512_p10": dict(width=512, seed=0, perturb=0.1, pe_noise=0.003, bias=0.75),
}


def trained_like_state_dict(width=256, steps=200, seed=3, lr=1e-3, n_pts=2048):
    """A "trained-like" network: `steps` Adam steps of the CPU oracle's SDF MLP (fp32 autograd) from the geometric init
    towards a bumpy, non-convex target SDF (sphere of radius 0.55 with 0.06-amplitude sinusoidal bumps of spatial frequency
    9) plus the eikonal term -- weights leave the initialisation regime (larger row norms, signal in the positional-
    encoding columns), which is what the screening-precision kernel has never seen in the golden fixtures."""
    import torch
    from oracle import mvsdf_oracle as O
    sd = synth.make_state_dict(width=width, seed=seed, perturb=0.05, pe_noise=0.003, bias=0.6)
    names = [k for k in sd if k.startswith("implicit_network")]
    params = {k: sd[k].clone().requires_grad_(True) for k in names}
    opt = torch.optim.Adam(list(params.values()), lr=lr)
    g = torch.Generator().manual_seed(seed + 17)
    torch.set_num_threads(max(1, min(16, torch.get_num_threads())))
    for _ in range(steps):
        x = (torch.rand(n_pts, 3, generator=g) * 2 - 1)
        target = x.norm(dim=1) - 0.55 - 0.06 * torch.sin(9 * x[:, 0]) * torch.sin(9 * x[:, 1]) * torch.sin(9 * x[:, 2])
        cur = dict(sd)
        cur.update(params)
        w = O.sdf_weights(cur)
        xr = x.clone().requires_grad_(True)
        f = O.sdf_mlp(xr, w)[:, 0]
        (gx,) = torch.autograd.grad(f.sum(), xr, create_graph=True)
        loss = (f - target).abs().mean() + 0.1 * ((gx.norm(dim=1) - 1) ** 2).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
    out = dict(sd)
    out.update({k: v.detach().clone() for k, v in params.items()})
    return out


def preset_state_dict(name, expect_sha=None):
    sd = synth.make_state_dict(**WEIGHT_PRESETS[name])
    if expect_sha is not None:
        got = synth.state_dict_checksum(sd)
        assert got == str(expect_sha), (
            f"synthetic weights for preset {name} drifted from the golden fixture ({got} != {expect_sha}); "
            "regenerate with python -m oracle.make_golden in the build container")
    return sd


def scene_from_meta(g):
    H, W, n_images, n_src, n_rays, seed = [int(v) for v in g["meta_scene"]]
    mode = str(g["meta_mask_mode"]) if "meta_mask_mode" in g else "ones"
    return synth.make_scene(H, W, n_images=n_images, n_src=n_src, n_rays=None if n_rays < 0 else n_rays,
                            seed=seed, mask_mode=mode)


# ---- parity gates ---------------------------------------------------------------------------------------------------
# Every tolerance of the GPU parity tests goes through gate(): it asserts value <= limit and records (test, name, value,
# limit); conftest.py dumps the records to gpurun_out/gate_report.json at the end of the session, which is how the
# limits were set (~3x the measured error on the B200, VERDICT r1 "weak #2") and how a regression shows up as a number.
GATE_LOG = []


def gate(name, value, limit, note=""):
    import os
    value = float(value)
    GATE_LOG.append({"test": os.environ.get("PYTEST_CURRENT_TEST", "").split(" ")[0], "gate": name, "value": value,
                     "limit": float(limit), "note": note})
    assert value <= limit, f"gate {name}: {value:.4e} > {limit:.4e} {note}"
    return value


def t(a):
    return torch.from_numpy(np.asarray(a))


def rel_err(a, b, floor=1e-6):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return ((a - b).abs() / b.abs().clamp_min(floor)).max().item() if a.numel() else 0.0
