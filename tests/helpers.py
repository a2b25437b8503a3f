"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np
import torch

from mvsdf_b200 import synth

WEIGHT_PRESETS = {
    "w256": dict(width=256, seed=1, perturb=0.05, pe_noise=0.003, bias=0.6),
    "w256_geo": dict(width=256, seed=1, perturb=0.0, pe_noise=0.0, bias=0.6),
    "w512": dict(width=512, seed=0, perturb=0.05, pe_noise=0.003, bias=0.75),
}


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


def t(a):
    return torch.from_numpy(np.asarray(a))


def rel_err(a, b, floor=1e-6):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return ((a - b).abs() / b.abs().clamp_min(floor)).max().item() if a.numel() else 0.0
