"""Edge cases of the hot path on the GPU, checked against the CPU oracle on the same seeded inputs:
rays that miss the bounding sphere, no surface at all, ragged multi-image batches, the IDR_RENDER tracer
switch, training mode with masked-out pixels and with minimal_sdf_points skipped."""
import os

import pytest
import torch

from mvsdf_b200 import synth
from oracle import mvsdf_oracle as O
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
IN = ["uv", "pose", "intrinsics", "object_mask"]
GT = ["rgb", "feat", "cam", "feat_src", "src_cams", "size", "center"]


def _model(sd, width, dev):
    from mvsdf_b200.network import B200IDRNetwork, default_conf
    m = B200IDRNetwork(default_conf(width)).to(dev)
    m.load_state_dict(sd)
    return m


def _compare(out, ref, scene, max_flips=2, depth_rtol=1e-4):
    nm, rm = out["network_object_mask"].cpu(), ref["network_object_mask"]
    assert int((nm != rm).sum()) <= max_flips
    both = nm & rm
    if both.any():
        cam = scene["pose"][:, :3, 3].unsqueeze(1).repeat(1, scene["uv"].shape[1], 1).reshape(-1, 3)
        d_new = (out["points"].cpu() - cam).norm(dim=1)[both]
        d_ref = (ref["points"] - cam).norm(dim=1)[both]
        rel = (d_new - d_ref).abs() / d_ref.clamp_min(1e-6)
        assert (rel > depth_rtol).float().mean().item() <= 0.02, rel.max().item()
        err = (out["rgb_values"].cpu() - ref["rgb_values"]).abs().max(dim=1).values[both]
        assert err.median().item() < 2e-4
    # rays nobody hit keep rgb = 1 (implicit_differentiable_renderer.py:302)
    miss = ~nm
    if miss.any():
        assert torch.all(out["rgb_values"].cpu()[miss] == 1.0)
    return int((nm != rm).sum())


def test_rays_missing_the_sphere_and_ragged_batch():
    dev = torch.device("cuda:0")
    sd = synth.make_state_dict(width=256, seed=1, perturb=0.05, pe_noise=0.003, bias=0.6)
    scene = synth.make_scene(30, 26, n_images=3, n_src=2, n_rays=333, seed=11)
    scene["intrinsics"][:, 0, 0] *= 0.3          # wide field of view: most rays miss the unit sphere
    scene["intrinsics"][:, 1, 1] *= 0.3
    scene["cam"][:, 1, 0, 0] *= 0.3
    scene["cam"][:, 1, 1, 1] *= 0.3
    scene["src_cams"][:, :, 1, 0, 0] *= 0.3
    scene["src_cams"][:, :, 1, 1, 1] *= 0.3
    dirs, cam = O.camera_rays(scene["uv"], scene["pose"], scene["intrinsics"])
    _, hit = O.sphere_intersection(cam, dirs)
    assert 0.05 < hit.float().mean().item() < 0.8
    model = _model(sd, 256, dev).eval()
    out = model({k: scene[k].to(dev) for k in IN})
    ref = O.idr_forward(O.sdf_weights(sd), O.render_weights(sd), scene, None, False)
    flips = _compare(out, ref, scene)
    from mvsdf_b200.loss import B200IDRLoss
    losses = B200IDRLoss().hot_path_losses(out, {k: scene[k].to(dev) for k in GT}, 0.5)
    ref_l = O.hot_path_losses(ref, scene, 0.5)
    if flips == 0:
        assert rel_err(losses["rgb_loss"].cpu(), ref_l["rgb_loss"]) < 1e-3
        assert rel_err(losses["feat_loss"].cpu(), ref_l["feat_loss"], floor=1e-4) < 3e-2


def test_no_surface_at_all():
    dev = torch.device("cuda:0")
    sd = synth.make_state_dict(width=256, seed=1, perturb=0.0, pe_noise=0.0, bias=-0.3)     # SDF > 0 everywhere
    scene = synth.make_scene(16, 16, n_images=2, n_src=1, seed=3)
    model = _model(sd, 256, dev).eval()
    out = model({k: scene[k].to(dev) for k in IN})
    assert int(out["network_object_mask"].sum()) == 0
    assert out["diff_surf_pts"].shape == (0, 3)
    assert torch.all(out["rgb_values"] == 1.0)
    from mvsdf_b200.loss import B200IDRLoss
    losses = B200IDRLoss().hot_path_losses(out, {k: scene[k].to(dev) for k in GT}, 0.5)
    assert float(losses["rgb_loss"]) == 0.0 and float(losses["feat_loss"]) == 0.0


def test_idr_render_switch(monkeypatch):
    """IDR_USE_ENV=1 IDR_RENDER=1: dist_clip 0.05 and 40 sphere-tracing iterations (ray_tracing.py:127-129)."""
    dev = torch.device("cuda:0")
    sd = synth.make_state_dict(width=256, seed=1, perturb=0.05, pe_noise=0.003, bias=0.6)
    scene = synth.make_scene(20, 20, n_images=1, n_src=1, seed=5)
    monkeypatch.setenv("IDR_USE_ENV", "1")
    monkeypatch.setenv("IDR_RENDER", "1")
    model = _model(sd, 256, dev).eval()
    out = model({k: scene[k].to(dev) for k in IN})
    prm = O.TracerParams(dist_clip=0.05, sphere_tracing_iters=40)
    ref = O.idr_forward(O.sdf_weights(sd), O.render_weights(sd), scene, None, False, tracer=prm)
    _compare(out, ref, scene)


@pytest.mark.parametrize("skip_min_sdf", [False, True])
def test_training_forward_with_masked_pixels(skip_min_sdf):
    dev = torch.device("cuda:0")
    sd = synth.make_state_dict(width=256, seed=1, perturb=0.05, pe_noise=0.003, bias=0.6)
    scene = synth.make_scene(22, 22, n_images=2, n_src=2, n_rays=300, seed=8, mask_mode="disc")
    g = torch.Generator().manual_seed(77)
    steps = torch.rand(100, generator=g)
    eik = torch.rand(300, 3, generator=g) * 2 - 1
    model = _model(sd, 256, dev).train()
    model.skip_min_sdf = skip_min_sdf
    out = model({k: scene[k].to(dev) for k in IN}, 0.7, steps01=steps, eik_points=eik)
    ref = O.idr_forward(O.sdf_weights(sd), O.render_weights(sd), scene, 0.7, True, steps01=steps, eik_points=eik,
                        skip_min_sdf=skip_min_sdf)
    flips = _compare(out, ref, scene)
    if flips == 0:
        assert out["grad_theta"].shape == ref["grad_theta"].shape
        assert (out["grad_theta"].cpu() - ref["grad_theta"].detach()).abs().max().item() < 5e-3
        assert (out["eikonal_output"].cpu() - ref["eikonal_output"].detach()).abs().max().item() < 1e-4
        assert (out["surf_indicator_output"].cpu() - ref["surf_indicator_output"].detach()).abs().max().item() < 1e-4
        # `points` of non-hit rays come from minimal_sdf_points / closest approach (ray_tracing.py:73-94)
        if not skip_min_sdf:
            assert (out["points"].cpu() - ref["points"]).abs().max().item() < 2e-3


def test_quaternion_pose_forward_equals_matrix_pose_forward():
    """rend_util.get_camera_params accepts [B,7] poses (quaternion + centre, :49-54); so does the drop-in.  The 4x4 built
    from the quaternion by the oracle must give the same outputs as passing the quaternion itself."""
    from tests.test_oracle import _pose7_from_matrix
    dev = torch.device("cuda:0")
    sd = synth.make_state_dict(width=256, seed=1, perturb=0.05, pe_noise=0.003, bias=0.6)
    model = _model(sd, 256, dev)
    model.eval()
    scene = synth.make_scene(20, 20, n_images=2, n_src=1, seed=6)
    pose7 = _pose7_from_matrix(scene["pose"])
    mat = O.quaternion_pose_matrix(pose7)
    a = model({"uv": scene["uv"].to(dev), "pose": pose7.to(dev), "intrinsics": scene["intrinsics"].to(dev),
               "object_mask": scene["object_mask"].to(dev)})
    b = model({"uv": scene["uv"].to(dev), "pose": mat.to(dev), "intrinsics": scene["intrinsics"].to(dev),
               "object_mask": scene["object_mask"].to(dev)})
    assert int(a["network_object_mask"].sum()) > 0
    hit = a["network_object_mask"] & b["network_object_mask"]
    assert (a["network_object_mask"] != b["network_object_mask"]).float().mean().item() < 0.01
    assert (a["points"][hit] - b["points"][hit]).abs().max().item() < 1e-4
    assert (a["rgb_values"][hit] - b["rgb_values"][hit]).abs().max().item() < 2e-3


def test_fp16_range_monitors_raise():
    """Weights / activations are fp16 hi/lo pairs of 64*x (csrc/netplan.h): |W| >= 1023.5 cannot be packed and an
    activation >= 1023 turns non-finite.  Both must be reported loudly, never silently."""
    from mvsdf_b200 import _lib, ops, synth
    from mvsdf_b200.network import B200IDRNetwork, default_conf
    dev = torch.device("cuda:0")
    sd = synth.make_state_dict(width=256, seed=1, perturb=0.05, pe_noise=0.003, bias=0.6)
    scene = synth.make_scene(16, 16, n_images=1, n_src=1, seed=0)
    inp = {k: scene[k].to(dev) for k in ["uv", "pose", "intrinsics", "object_mask"]}
    # (1) healthy network: monitors stay at zero
    model = B200IDRNetwork(default_conf(256)).to(dev)
    model.load_state_dict(sd)
    model.eval()
    model(inp)
    assert int(model.implicit_network._net.status()[:2].abs().sum()) == 0
    assert int(model.rendering_network._net.status()[:2].abs().sum()) == 0
    # (2) a folded weight beyond the fp16 range: the packer counts it, forward() raises
    bad = {k: v.clone() for k, v in sd.items()}
    bad["rendering_network.lin1.weight_g"][3] *= 5.0e4
    model.load_state_dict(bad)
    with pytest.raises(_lib.MvsdfError, match="beyond the fp16 range"):
        model(inp)
    # (3) weights in range but activations that leave it: a large bias on a ReLU layer of the rendering net
    bad = {k: v.clone() for k, v in sd.items()}
    bad["rendering_network.lin0.bias"] += 2000.0
    model.load_state_dict(bad)
    with pytest.raises(_lib.MvsdfError, match="non-finite"):
        model(inp)
    # (4) the same monitors through the plain op wrappers
    net = ops.PackedNet("sdf", 256, 8).pack_state_dict(sd, "implicit_network", dev)
    ops.sdf_forward(net, torch.zeros(10, 3, device=dev), ops.HEAD_SDF_ONLY)
    net.check_status()
    ops.sdf_forward(net, torch.full((10, 3), float("nan"), device=dev), ops.HEAD_SDF_ONLY)
    with pytest.raises(_lib.MvsdfError, match="non-finite"):
        net.check_status()
