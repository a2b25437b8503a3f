"""BASELINE.json's full sizes (cfg2: 1200x1600 = 1.92 M rays, 8x512 SDF net, 4 source views; cfg5-shaped 1080p) are far
beyond what the CPU oracle finishes in seconds, so parity at these sizes is checked through size-independent properties
of the hot path (the small-size cases are compared value by value in test_gpu_pipeline.py):
  * the result of a ray does not depend on where it sits in the batch (tile, CTA, sampler batch): tracing the rays in
    reverse order gives the reversed outputs bit for bit;
  * the tracer's prefilter does not change a single bit;
  * points = cam_loc + dists * ray_dirs (implicit_differentiable_renderer.py:200);
  * hit rays end on the surface: the SDF re-evaluated at the returned points is ~0 (ray_tracing.py:143-151 threshold
    5e-5 for sphere-traced rays, 8 secant steps for the others);
  * the loss partials of the two halves of the image add up to those of the whole image (row e: ray sharding)."""
import pytest
import torch

from mvsdf_b200 import synth

pytestmark = pytest.mark.gpu

CFG2_WEIGHTS = dict(width=512, seed=0, perturb=0.05, pe_noise=0.003, bias=0.75)
IN_KEYS = ["uv", "pose", "intrinsics", "object_mask"]
GT_KEYS = ["rgb", "feat", "cam", "feat_src", "src_cams", "size", "center"]


def _same(a, b):
    if a.is_floating_point():
        return bool(((a == b) | (torch.isnan(a) & torch.isnan(b))).all())
    return bool((a == b).all())


@pytest.fixture(scope="module")
def cfg2():
    from mvsdf_b200.network import B200IDRNetwork, default_conf
    dev = torch.device("cuda:0")
    scene = synth.make_scene(1200, 1600, n_images=1, n_src=4, seed=0)
    model = B200IDRNetwork(default_conf(512)).to(dev)
    model.load_state_dict(synth.make_state_dict(**CFG2_WEIGHTS))
    model.eval()
    return model, scene, dev


def test_cfg2_tracer_order_and_prefilter_invariance(cfg2):
    model, scene, dev = cfg2
    sdf_net = model.implicit_network.packed()
    uv, pose, K = scene["uv"].to(dev), scene["pose"].to(dev), scene["intrinsics"].to(dev)
    R = uv.shape[0] * uv.shape[1]
    assert R == 1920000
    obj = torch.ones(R, dtype=torch.uint8, device=dev)
    tau = model.prefilter_tau
    assert tau > 0.0
    dirs, cam, dists, nm, pts = [t.clone() for t in model.trace(sdf_net, uv, pose, K, obj, False)]
    cnt = model.last_trace_counters.cpu()
    assert int(cnt[255]) == 0, "screening guard tripped on the benchmark weights"
    assert 0 < int(cnt[254]) < 20 * R
    # (1) prefilter off: identical bits
    model.prefilter_tau = 0.0
    try:
        d0, c0, t0, m0, p0 = model.trace(sdf_net, uv, pose, K, obj, False)
        assert _same(t0, dists) and _same(m0, nm) and _same(p0, pts)
        assert torch.equal(model.last_trace_counters.cpu()[:251], cnt[:251])
    finally:
        model.prefilter_tau = tau
    # (2) reversed ray order: reversed outputs
    d1, c1, t1, m1, p1 = model.trace(sdf_net, uv.flip(1).contiguous(), pose, K, obj, False)
    assert _same(d1.flip(0), dirs) and _same(t1.flip(0), dists) and _same(m1.flip(0), nm) and _same(p1.flip(0), pts)
    # (3) points = cam + dists * dirs
    assert (pts - (cam[0] + dists.unsqueeze(1) * dirs)).abs().max().item() < 1e-6
    # (4) hit rays lie on the zero level set
    from mvsdf_b200 import ops
    hit = nm.bool()
    frac = hit.float().mean().item()
    assert 0.3 < frac < 0.9, frac
    s = ops.sdf_forward(sdf_net, pts[hit].contiguous(), ops.HEAD_SDF_ONLY).abs()
    assert (s < 1e-3).float().mean().item() > 0.995, (s < 1e-3).float().mean().item()
    assert s.median().item() < 5e-5


def test_cfg2_forward_loss_partials_add_up_over_image_halves(cfg2):
    from mvsdf_b200.loss import B200IDRLoss
    model, scene, dev = cfg2
    gt = {k: scene[k].to(dev) for k in GT_KEYS}
    full_loss = B200IDRLoss()
    out = model({k: scene[k].to(dev) for k in IN_KEYS})
    full = full_loss.hot_path_losses(out, gt, 0.5)
    assert torch.isfinite(full["rgb_loss"]) and torch.isfinite(full["feat_loss"])
    p_rgb, p_feat = full_loss.last_partials["rgb"].clone(), full_loss.last_partials["feat"].clone()
    N = scene["uv"].shape[1]
    acc_rgb, acc_feat = torch.zeros_like(p_rgb), torch.zeros_like(p_feat)
    for sl in (slice(0, N // 2), slice(N // 2, N)):
        part_in = {"uv": scene["uv"][:, sl].contiguous().to(dev), "pose": scene["pose"].to(dev),
                   "intrinsics": scene["intrinsics"].to(dev), "object_mask": scene["object_mask"][:, sl].contiguous().to(dev)}
        part_gt = dict(gt)
        part_gt["rgb"] = scene["rgb"][:, sl].contiguous().to(dev)
        l2 = B200IDRLoss()
        l2.hot_path_losses(model(part_in), part_gt, 0.5)
        acc_rgb += l2.last_partials["rgb"]
        acc_feat += l2.last_partials["feat"]
    assert torch.allclose(acc_rgb, p_rgb, rtol=1e-9, atol=1e-9)
    assert torch.allclose(acc_feat, p_feat, rtol=1e-7, atol=1e-9)


def test_cfg5_shaped_1080p_12_views_runs_and_is_order_invariant():
    """Tanks&Temples-shaped config 5 (1080p, 12 source views, 8x512): one rank's share of the 8-GPU sweep is the whole
    image here; reversed ray order must reproduce the same losses."""
    from mvsdf_b200.loss import B200IDRLoss
    from mvsdf_b200.network import B200IDRNetwork, default_conf
    dev = torch.device("cuda:0")
    scene = synth.make_scene(1080, 1920, n_images=1, n_src=12, seed=1, n_rays=400000)
    model = B200IDRNetwork(default_conf(512)).to(dev)
    model.load_state_dict(synth.make_state_dict(**CFG2_WEIGHTS))
    model.eval()
    gt = {k: scene[k].to(dev) for k in GT_KEYS}
    a = B200IDRLoss().hot_path_losses(model({k: scene[k].to(dev) for k in IN_KEYS}), gt, 0.5)
    rev_in = {"uv": scene["uv"].flip(1).contiguous().to(dev), "pose": scene["pose"].to(dev),
              "intrinsics": scene["intrinsics"].to(dev), "object_mask": scene["object_mask"].flip(1).contiguous().to(dev)}
    rev_gt = dict(gt)
    rev_gt["rgb"] = scene["rgb"].flip(1).contiguous().to(dev)
    b = B200IDRLoss().hot_path_losses(model(rev_in), rev_gt, 0.5)
    for k in ("rgb_loss", "feat_loss"):
        assert torch.isfinite(a[k])
        assert abs(float(a[k]) - float(b[k])) <= 1e-6 * max(1.0, abs(float(a[k]))), k


# ---- value parity at the full-size configurations -----------------------------------------------------------------------
# Rays are independent, so a random sub-sample of the full image is a legitimate unit for the CPU oracle: the GPU path is
# run on the WHOLE image (1.92 M rays / 1080p), the sub-sample's rows are shown to be bit-identical to a GPU run over the
# sub-sample alone, and that run is compared value by value with oracle.idr_forward + hot_path_losses on exactly those rays
# (same cameras, same 600x800 / 540x960 feature maps, same weights).
def _subsample_parity(model, scene, dev, n_sub, seed, full_out=None):
    from mvsdf_b200.loss import B200IDRLoss
    from oracle import mvsdf_oracle as O
    from tests.helpers import gate
    N = scene["uv"].shape[1]
    idx = torch.randperm(N, generator=torch.Generator().manual_seed(seed))[:n_sub].sort().values
    sub = dict(scene)
    for k in ("uv", "object_mask", "rgb"):
        sub[k] = scene[k][:, idx].contiguous()
    out = model({k: sub[k].to(dev) for k in IN_KEYS})
    losses = B200IDRLoss().hot_path_losses(out, {k: sub[k].to(dev) for k in GT_KEYS}, 0.5)
    if full_out is not None:       # the sub-sample run IS the full-size run restricted to these rays
        for k in ("points", "rgb_values", "network_object_mask", "diff_surf_pts"):
            want = full_out[k][idx.to(dev)] if k != "diff_surf_pts" else full_out["points"][idx.to(dev)][out["network_object_mask"]]
            assert _same(want, out[k]), f"{k}: sub-sample run differs from the full-image run"
        # sdf_output is the one launch with a host-known count: 4 096 points take the single-CTA scheduling of the tile
        # core (mlp_kernel.cuh), 1.92 M the CTA-pair kernel, whose softplus is formulated differently and whose one-row head
        # is an fp32 dot product in the last hidden layer's epilogue instead of an UMMA over fp16 hi/lo activations -- same
        # value to fp32 rounding of a 512-term sum (measured 4.5e-6; both are within 2e-5 of the fp64 oracle), not the same bits
        gate("sdf_output_small_vs_pair_kernel", (full_out["sdf_output"][idx.to(dev)] - out["sdf_output"]).abs().max().item(), 1.5e-5)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    with torch.no_grad():
        ref = O.idr_forward(O.sdf_weights(sd), O.render_weights(sd), sub, None, False)
        ref_l = O.hot_path_losses(ref, sub, 0.5)
    nm, rm = out["network_object_mask"].cpu(), ref["network_object_mask"]
    flips = int((nm != rm).sum())
    print(f"sub-sample of {n_sub} rays: {flips} hit-mask flips, {int(rm.sum())} hits")
    gate("hit_mask_flip_frac", flips / n_sub, 1e-3)
    both = nm & rm
    cam = scene["pose"][0, :3, 3]
    d_ref = (ref["points"] - cam).norm(dim=1)[both]
    d_new = (out["points"].cpu() - cam).norm(dim=1)[both]
    rel = (d_new - d_ref).abs() / d_ref
    gate("depth_rel_frac_above_1e-4", (rel > 1e-4).float().mean().item(), 0.002, f"(max {rel.max():.2e})")
    gate("depth_rel_median", rel.median().item(), 2e-5)
    rgb_err = (out["rgb_values"].cpu() - ref["rgb_values"]).abs().max(dim=1).values[both]
    gate("rgb_abs_median", rgb_err.median().item(), 5e-7)
    gate("rgb_abs_frac_above_1e-5", (rgb_err > 1e-5).float().mean().item(), 0.002, f"(max {rgb_err.max():.2e})")
    if flips == 0:
        # a few grazing rays sit on a sampler / secant branch boundary (they are the rays counted by depth_rel_frac above):
        # bound the bulk tightly and the worst ray loosely
        dsp = (out["diff_surf_pts"].cpu() - ref["diff_surf_pts"]).norm(dim=1)
        gate("diff_surf_pts_q998", torch.quantile(dsp, 0.998).item(), 3e-4)
        gate("diff_surf_pts_abs_max", dsp.max().item(), 5e-3)
        gate("rgb_loss_rel", abs(float(losses["rgb_loss"]) - float(ref_l["rgb_loss"])) / abs(float(ref_l["rgb_loss"])), 2e-6)
        # (the grazing rays above move their feature-loss terms: measured 1.1e-3 on the 4 096-ray sample, 4e-5 where none occur)
        gate("feat_loss_rel", abs(float(losses["feat_loss"]) - float(ref_l["feat_loss"])) / abs(float(ref_l["feat_loss"])), 3e-3)
    else:
        gate("rgb_loss_abs", abs(float(losses["rgb_loss"]) - float(ref_l["rgb_loss"])), 5e-3)


def test_cfg2_random_subsample_matches_cpu_oracle(cfg2):
    model, scene, dev = cfg2
    full_out = model({k: scene[k].to(dev) for k in IN_KEYS})
    _subsample_parity(model, scene, dev, 4096, seed=11, full_out=full_out)


def test_cfg5_random_subsample_matches_cpu_oracle():
    """1080p, 12 source views (BASELINE.json configs[4]): 2 048 random rays of the full image against the CPU oracle."""
    from mvsdf_b200.network import B200IDRNetwork, default_conf
    dev = torch.device("cuda:0")
    scene = synth.make_scene(1080, 1920, n_images=1, n_src=12, seed=1)
    model = B200IDRNetwork(default_conf(512)).to(dev)
    model.load_state_dict(synth.make_state_dict(**CFG2_WEIGHTS))
    model.eval()
    _subsample_parity(model, scene, dev, 2048, seed=12)
