"""GPU parity of the full hot path (tracer -> shading -> losses) against the golden outputs of the
reference and the CPU oracle.  Everything goes through B200IDRNetwork / B200IDRLoss -> C ABI."""
import pytest
import torch

from oracle import mvsdf_oracle as O
from tests.helpers import WEIGHT_PRESETS, gate, preset_state_dict, rel_err, scene_from_meta, t

pytestmark = pytest.mark.gpu

DEPTH_RTOL = 1e-4          # BASELINE.md parity gate: 1e-4 relative (fp32) on rays whose discrete decisions agree


def _model(preset, device):
    from mvsdf_b200.network import B200IDRNetwork, default_conf
    sd = preset_state_dict(preset)
    m = B200IDRNetwork(default_conf(WEIGHT_PRESETS[preset]["width"])).to(device)
    m.load_state_dict(sd)
    return m, sd


def _to(d, keys, dev):
    return {k: d[k].to(dev) for k in keys}


def _check_forward(out, g, scene, training):
    """Limits = ~3x the errors measured on the B200 (gpurun_out/gate_report.json of round 2), never looser than the
    north-star gate of 1e-4 relative on rays whose discrete decisions agree."""
    nm = out["network_object_mask"].cpu()
    ref_nm = t(g["network_object_mask"])
    flips = int((nm != ref_nm).sum())
    gate("hit_mask_flips", flips, G_FLIPS, f"of {nm.numel()} rays")
    both = nm & ref_nm
    cam = scene["pose"][:, :3, 3].unsqueeze(1).repeat(1, scene["uv"].shape[1], 1).reshape(-1, 3)
    d_ref = (t(g["points"]) - cam).norm(dim=1)
    d_new = (out["points"].cpu() - cam).norm(dim=1)
    rel = ((d_new - d_ref).abs() / d_ref.clamp_min(1e-6))[both]
    # continuous parity on rays whose decisions agree; a grazing ray may sit on a sampler/secant branch boundary
    # (SURVEY 7.3-1): their fraction and their error are bounded
    gate("depth_rel_frac_above_1e-4", (rel > DEPTH_RTOL).float().mean().item(), G_DEPTH_FRAC, f"(max {rel.max():.2e})")
    gate("depth_rel_max", rel.max().item(), G_DEPTH_MAX)
    gate("depth_rel_median", rel.median().item(), G_DEPTH_MEDIAN)
    rgb_err = (out["rgb_values"].cpu() - t(g["rgb_values"])).abs().max(dim=1).values[both]
    gate("rgb_abs_max", rgb_err.max().item(), G_RGB_MAX)
    gate("rgb_abs_median", rgb_err.median().item(), G_RGB_MEDIAN)
    sdf_err = (out["sdf_output"].cpu() - t(g["sdf_output"])).abs()[both]
    gate("sdf_output_abs_max", sdf_err.max().item(), G_SDF_MAX)
    return flips


# gate limits (see _check_forward)
# measured on the B200 (profiles/r02/gate_report.json), worst case over the fixtures -> limit
G_FLIPS = 1                 # 0 flips on every fixture
G_DEPTH_FRAC = 0.002        # 0 rays above 1e-4
G_DEPTH_MAX = 1.5e-4        # 5.1e-5
G_DEPTH_MEDIAN = 2e-5       # 5.6e-6
G_RGB_MAX = 1e-5            # 1.6e-6
G_RGB_MEDIAN = 5e-7         # 1.2e-7
G_SDF_MAX = 1e-4            # 3.6e-5
G_SURF_PTS = 3e-4           # 1.0e-4
G_RGB_LOSS_REL = 2e-6       # 1.1e-7
G_FEAT_LOSS_REL = 1.5e-4    # 4.2e-5
G_GRAD_THETA = 2e-3         # 6.6e-4
G_EIK_LOSS_REL = 1.5e-4     # 4.1e-5
G_SURF_LOSS_REL = 1.5e-5    # 4.2e-6


@pytest.mark.parametrize("name", ["cfg1_eval_w256", "small_eval_w512", "cfg2_shape_eval_w512"])
def test_eval_forward_vs_reference_golden(golden, name):
    from mvsdf_b200.loss import B200IDRLoss
    g = golden(name)
    dev = torch.device("cuda:0")
    model, sd = _model(str(g["meta_preset"]), dev)
    scene = scene_from_meta(g)
    model.eval()
    out = model(_to(scene, ["uv", "pose", "intrinsics", "object_mask"], dev))
    flips = _check_forward(out, g, scene, False)
    gt = _to(scene, ["rgb", "feat", "cam", "feat_src", "src_cams", "size", "center"], dev)
    losses = B200IDRLoss().hot_path_losses(out, gt, 0.5)
    if flips == 0:
        assert out["diff_surf_pts"].shape == tuple(g["diff_surf_pts"].shape)
        gate("diff_surf_pts_abs_max", (out["diff_surf_pts"].cpu() - t(g["diff_surf_pts"])).abs().max().item(), G_SURF_PTS)
        gate("rgb_loss_rel", rel_err(losses["rgb_loss"].cpu(), g["rgb_loss"]), G_RGB_LOSS_REL)
        gate("feat_loss_rel", rel_err(losses["feat_loss"].cpu(), g["feat_loss"]), G_FEAT_LOSS_REL)
    else:
        gate("rgb_loss_abs_with_flips", abs(float(losses["rgb_loss"]) - float(g["rgb_loss"])), 0.01)


@pytest.mark.parametrize("name", ["cfg1_train_w256", "cfg3_shape_train_w512"])
def test_train_forward_vs_reference_golden(golden, name):
    from mvsdf_b200.loss import B200IDRLoss
    g = golden(name)
    dev = torch.device("cuda:0")
    model, sd = _model(str(g["meta_preset"]), dev)
    scene = scene_from_meta(g)
    model.train()
    steps = t(g["steps01"]) if "steps01" in g else None
    out = model(_to(scene, ["uv", "pose", "intrinsics", "object_mask"], dev), 0.5, steps01=steps, eik_points=t(g["eik_points"]))
    flips = _check_forward(out, g, scene, True)
    gt = _to(scene, ["rgb", "feat", "cam", "feat_src", "src_cams", "size", "center"], dev)
    losses = B200IDRLoss().hot_path_losses(out, gt, 0.5)
    if flips == 0:
        gate("grad_theta_abs_max", (out["grad_theta"].cpu() - t(g["grad_theta"])).abs().max().item(), G_GRAD_THETA)
        gate("eikonal_loss_rel", rel_err(losses["eikonal_loss"].cpu(), g["eikonal_loss"]), G_EIK_LOSS_REL)
        gate("surf_loss_rel", rel_err(losses["surf_loss"].cpu(), g["surf_loss"]), G_SURF_LOSS_REL)
        gate("rgb_loss_rel", rel_err(losses["rgb_loss"].cpu(), g["rgb_loss"]), G_RGB_LOSS_REL)
        gate("feat_loss_rel", rel_err(losses["feat_loss"].cpu(), g["feat_loss"]), G_FEAT_LOSS_REL)


@pytest.mark.parametrize("name", ["cfg1_train_w256", "train_phase0_w256"])
def test_depth_carving_loss_vs_reference_golden(golden, name):
    """IDRLoss.get_depth_loss (loss.py:37-63, carving_t2) on the reference's own eikonal set."""
    from mvsdf_b200.loss import B200IDRLoss
    g = golden(name)
    dev = torch.device("cuda:0")
    scene = scene_from_meta(g)
    tp = float(g["meta_tp"])
    loss = B200IDRLoss()
    dl = loss.get_depth_loss(t(g["eikonal_points_hom_all"]).to(dev), t(g["eikonal_output"]).to(dev), scene["depths"].to(dev),
                             scene["depth_cams"].to(dev), scene["size"][:1].to(dev), scene["center"][:1].to(dev),
                             train_progress=tp)
    assert rel_err(dl.cpu(), g["depth_loss"]) < 1e-5


def test_full_loss_dict_vs_reference_golden(golden):
    """B200IDRLoss.forward returns the reference's dict (loss.py:212-219) with its schedule-dependent weights."""
    from mvsdf_b200.loss import B200IDRLoss
    from mvsdf_b200 import conf
    g = golden("cfg1_train_w256")
    dev = torch.device("cuda:0")
    model, sd = _model(str(g["meta_preset"]), dev)
    scene = scene_from_meta(g)
    model.train()
    tp = float(g["meta_tp"])
    with torch.no_grad():
        out = model(_to(scene, ["uv", "pose", "intrinsics", "object_mask"], dev), tp, steps01=t(g["steps01"]),
                    eik_points=t(g["eik_points"]))
    gt = _to(scene, ["rgb", "feat", "cam", "feat_src", "src_cams", "size", "center", "depths", "depth_cams"], dev)
    res = B200IDRLoss()(out, gt, tp, 2)
    assert set(res) == {"loss", "rgb_loss", "eikonal_loss", "depth_loss", "feat_loss", "surf_loss"}
    if int((out["network_object_mask"].cpu() != t(g["network_object_mask"])).sum()) == 0:
        gate("depth_loss_rel", rel_err(res["depth_loss"].cpu(), g["depth_loss"]), 2e-4)
        want = (0.5 * float(g["rgb_loss"]) + conf.eikonal_weight * float(g["eikonal_loss"]) + conf.surf_weight * float(g["surf_loss"])
                + conf.feat_weight(tp) * float(g["feat_loss"]) + float(g["depth_loss"]))
        gate("total_loss_rel", abs(float(res["loss"]) - want) / abs(want), 2e-4)


def test_train_phase0_forward_vs_reference_golden(golden):
    """train_progress < 1/6: depth-surface samples (on the MVS depth surface and jittered around it) join the eikonal
    set (implicit_differentiable_renderer.py:226-251); the reference's rand_like / np.random.choice draws are replayed."""
    g = golden("train_phase0_w256")
    dev = torch.device("cuda:0")
    model, sd = _model(str(g["meta_preset"]), dev)
    scene = scene_from_meta(g)
    model.train()
    keys = ["uv", "pose", "intrinsics", "object_mask", "depths", "depth_cams", "center", "size"]
    rnd = dict(jitter01=t(g["dsurf_jitter01"]), idx_on=g["dsurf_idx_on"], idx_jitter=g["dsurf_idx_jitter"])
    out = model(_to(scene, keys, dev), float(g["meta_tp"]), steps01=t(g["steps01"]), eik_points=t(g["eik_points"]),
                dsurf_rand=rnd)
    flips = _check_forward(out, g, scene, True)
    n_hit = int(t(g["network_object_mask"]).sum())
    # the depth-surface rows are independent of the tracer's decisions: compare them even if a ray flipped
    hom = out["eikonal_points_hom"].cpu().reshape(-1, 4)
    ref_hom = t(g["eikonal_points_hom"]).reshape(-1, 4)
    n_extra = ref_hom.shape[0] - n_hit
    assert (hom[-n_extra:] - ref_hom[-n_extra:]).abs().max().item() < 2e-6
    assert (out["eikonal_output"].cpu().reshape(-1)[-n_extra:] - t(g["eikonal_output"]).reshape(-1)[-n_extra:]).abs().max().item() < 5e-5
    gate("grad_theta_extra_abs_max", (out["grad_theta"].cpu()[-n_extra:] - t(g["grad_theta"])[-n_extra:]).abs().max().item(), G_GRAD_THETA)
    if flips == 0:
        assert out["grad_theta"].shape == tuple(g["grad_theta"].shape)
        gate("grad_theta_abs_max", (out["grad_theta"].cpu() - t(g["grad_theta"])).abs().max().item(), G_GRAD_THETA)
        assert (out["surf_indicator_output"].cpu() - t(g["surf_indicator_output"])).abs().max().item() < 1e-4


@pytest.mark.parametrize("name", ["tracer_eval_w256", "tracer_train_w256", "tracer_eval_w256_geo"])
def test_tracer_vs_reference_golden(golden, name):
    g = golden(name)
    dev = torch.device("cuda:0")
    model, sd = _model(str(g["meta_preset"]), dev)
    scene = scene_from_meta(g)
    training = bool(int(g["meta_training"]))
    sdf_net = model.implicit_network.packed()
    steps = t(g["steps01"]) if "steps01" in g else None
    obj = scene["object_mask"].reshape(-1).to(torch.uint8).to(dev)
    dirs, cam, dists, nm, pts = model.trace(sdf_net, scene["uv"].to(dev), scene["pose"].to(dev), scene["intrinsics"].to(dev),
                                            obj, training, steps)
    assert (dirs.cpu() - t(g["ray_dirs"]).reshape(-1, 3)).abs().max().item() < 2e-6
    nm = nm.bool().cpu()
    ref_nm = t(g["network_object_mask"])
    gate("hit_mask_flips", int((nm != ref_nm).sum()), G_FLIPS)
    both = nm & ref_nm
    d_ref = t(g["dists"])
    rel = ((dists.cpu() - d_ref).abs() / d_ref.abs().clamp_min(1e-6))[both]
    gate("depth_rel_frac_above_1e-4", (rel > DEPTH_RTOL).float().mean().item(), G_DEPTH_FRAC, f"(max {rel.max():.2e})")
    gate("depth_rel_max", rel.max().item(), G_DEPTH_MAX)
    # E_trace (the numerator of bench.py's roofline): the request counters must be the evaluation count of the REFERENCE
    # algorithm -- here the oracle's counter on the same rays -- whatever the prefilter skips or repeats
    from mvsdf_b200 import _lib
    cnt = model.last_trace_counters.cpu()
    e_trace = int(cnt[:_lib.CTR_SCREENED].sum())
    oc = O.TraceCounters()
    o_dirs, o_cam = O.camera_rays(scene["uv"], scene["pose"], scene["intrinsics"])
    sw = O.sdf_weights(sd)
    with torch.no_grad():
        O.trace_rays(lambda x: O.sdf_mlp(x, sw)[:, 0], o_cam, scene["object_mask"].reshape(-1), o_dirs, training=training,
                     steps01=steps, counters=oc)
    # one ray that enters (or stays out of) the 100-sample sampler on a rounding difference moves the count by ~100
    gate("e_trace_abs_diff", abs(e_trace - oc.total), max(150, 1e-3 * oc.total), f"({e_trace} vs {oc.total})")
    if model.prefilter_tau > 0:
        assert int(cnt[_lib.CTR_SCREENED]) + int(cnt[_lib.CTR_REFINED]) < oc.sampler + oc.min_sdf or oc.sampler + oc.min_sdf == 0


def test_shard_invariance_of_loss_partials():
    """Rows (e): rays shard across GPUs with one all-reduce(SUM) of the loss partials.  Splitting the rays of
    each image in two on ONE device and summing the partials must reproduce the unsplit losses."""
    from mvsdf_b200.loss import B200IDRLoss
    from mvsdf_b200 import synth
    dev = torch.device("cuda:0")
    model, sd = _model("w256", dev)
    model.eval()
    scene = synth.make_scene(24, 24, n_images=2, n_src=2, seed=4)
    keys_in = ["uv", "pose", "intrinsics", "object_mask"]
    gt_keys = ["rgb", "feat", "cam", "feat_src", "src_cams", "size", "center"]
    loss = B200IDRLoss()
    full = loss.hot_path_losses(model(_to(scene, keys_in, dev)), _to(scene, gt_keys, dev), 0.5)
    p_rgb, p_feat = loss.last_partials["rgb"].clone(), loss.last_partials["feat"].clone()
    N = scene["uv"].shape[1]
    acc_rgb, acc_feat = torch.zeros_like(p_rgb), torch.zeros_like(p_feat)
    for sl in (slice(0, N // 2), slice(N // 2, N)):
        part = dict(scene)
        part["uv"] = scene["uv"][:, sl].contiguous()
        part["object_mask"] = scene["object_mask"][:, sl].contiguous()
        part["rgb"] = scene["rgb"][:, sl].contiguous()
        l2 = B200IDRLoss()
        l2.hot_path_losses(model(_to(part, keys_in, dev)), _to(part, gt_keys, dev), 0.5)
        acc_rgb += l2.last_partials["rgb"]
        acc_feat += l2.last_partials["feat"]
    assert torch.allclose(acc_rgb, p_rgb, rtol=1e-12, atol=1e-12)
    assert torch.allclose(acc_feat, p_feat, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("with_grad", [False, True])
def test_custom_schedule_switches_vs_oracle(with_grad):
    """A schedule module other than the shipped one (the reference's IDR_USE_ENV=1 + IDR_CONF=<module> override, :15-17) with the
    eight point-set switches of model/conf.py:4-14 in a mixed combination: eikonal_output / eikonal_points_hom / grad_theta
    are assembled from the selected sets only (:259-286), on the native and on the autograd path; the oracle's restatement of
    the switches is pinned to the live reference in tests/test_oracle.py."""
    import types
    import numpy as np
    from mvsdf_b200 import conf as shipped
    from mvsdf_b200.network import B200IDRNetwork, default_conf
    dev = torch.device("cuda:0")
    sd = preset_state_dict("w256")
    tp = 0.3
    sched = types.SimpleNamespace(**{k: getattr(shipped, k) for k in dir(shipped) if not k.startswith("_")})
    sched.d_use_eik = lambda t: False
    sched.d_use_dsurf_on = lambda t: True
    sched.eik_use_rt_surf = lambda t: False
    sched.eik_use_dsurf_jitter = lambda t: True
    scene = synth_scene = __import__("mvsdf_b200.synth", fromlist=["x"]).make_scene(48, 48, n_images=2, n_src=1, n_rays=128, seed=9)
    g = torch.Generator().manual_seed(5)
    steps = torch.rand(100, generator=g)
    eik = torch.rand(128, 3, generator=g) * 2 - 1
    ds = O.depth_surface_points(scene["depths"], scene["depth_cams"], scene["center"][:1], scene["size"][:1])
    torch.manual_seed(11)
    np.random.seed(11)
    _, _, rnd = O.depth_surface_samples(ds, 128, 1.0)
    with torch.no_grad():
        ref = O.idr_forward(O.sdf_weights(sd), O.render_weights(sd), scene, tp, True, steps01=steps, eik_points=eik, dsurf_rand=rnd,
                            schedule=sched)
    model = B200IDRNetwork(default_conf(256), schedule=sched).to(dev)
    model.load_state_dict(sd)
    model.train()
    keys = ["uv", "pose", "intrinsics", "object_mask", "depths", "depth_cams", "center", "size"]
    rnd_dev = {k: (v.to(dev) if k == "jitter01" else v.cpu().numpy()) for k, v in rnd.items()}
    ctx = torch.enable_grad() if with_grad else torch.no_grad()
    with ctx:
        out = model(_to(scene, keys, dev), tp, steps01=steps, eik_points=eik, dsurf_rand=rnd_dev)
    assert bool(out["rgb_values"].requires_grad) == with_grad
    if int((out["network_object_mask"].cpu() != ref["network_object_mask"]).sum()) != 0:
        pytest.skip("a discrete tracer decision flipped on this input")
    n_hit = int(ref["network_object_mask"].sum())
    assert out["eikonal_output"].shape == ref["eikonal_output"].shape == (1, n_hit + 128)
    assert out["grad_theta"].shape == ref["grad_theta"].shape == (256, 3)
    gate("eikonal_output_abs_max", (out["eikonal_output"].detach().cpu() - ref["eikonal_output"]).abs().max().item(), G_SDF_MAX)
    gate("eikonal_points_abs_max", (out["eikonal_points_hom"].detach().cpu() - ref["eikonal_points_hom"]).abs().max().item(), G_SURF_PTS)
    gate("grad_theta_abs_max", (out["grad_theta"].detach().cpu() - ref["grad_theta"]).abs().max().item(), G_GRAD_THETA)
