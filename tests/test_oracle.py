"""Pins oracle/mvsdf_oracle.py (the CPU restatement) against (a) the golden outputs of the
unmodified reference (tests/golden, everywhere) and (b) the live reference when
/root/reference is present (build container only)."""
import numpy as np
import pytest
import torch

from oracle import mvsdf_oracle as O
from oracle import ref_shim
from tests.helpers import preset_state_dict, rel_err, scene_from_meta, t

torch.set_num_threads(4)


@pytest.mark.parametrize("name", ["mlp_w256", "mlp_w512"])
def test_mlp_golden(golden, name):
    g = golden(name)
    sd = preset_state_dict(str(g["meta_preset"]), g["meta_weights_sha"])
    sw, rw = O.sdf_weights(sd), O.render_weights(sd)
    x = t(g["x"])
    full = O.sdf_mlp(x, sw)
    assert torch.allclose(full, t(g["sdf_full"]), rtol=1e-5, atol=2e-6)
    grad = O.sdf_gradient(x, sw)
    assert torch.allclose(grad, t(g["grad"]), rtol=1e-4, atol=2e-6)
    rgb = O.render_mlp(x, t(g["grad"]), t(g["view"]), t(g["sdf_full"])[:, 2:], rw)
    assert torch.allclose(rgb, t(g["rgb"]), rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("name", ["tracer_eval_w256", "tracer_train_w256", "tracer_eval_w256_geo"])
def test_tracer_golden(golden, name):
    g = golden(name)
    sd = preset_state_dict(str(g["meta_preset"]), g["meta_weights_sha"])
    sw = O.sdf_weights(sd)
    scene = scene_from_meta(g)
    dirs, cam = O.camera_rays(scene["uv"], scene["pose"], scene["intrinsics"])
    assert torch.allclose(dirs, t(g["ray_dirs"]), atol=1e-6)
    tnf, hit = O.sphere_intersection(cam, dirs)
    assert torch.equal(hit, t(g["hit_sphere"]))
    assert torch.allclose(tnf, t(g["t_near_far"]), atol=1e-5)
    training = bool(int(g["meta_training"]))
    steps = t(g["steps01"]) if "steps01" in g else None
    cnt = O.TraceCounters()
    with torch.no_grad():
        pts, nm, dists = O.trace_rays(lambda x: O.sdf_mlp(x, sw)[:, 0], cam, scene["object_mask"].reshape(-1),
                                      dirs, training=training, steps01=steps, counters=cnt)
    ref_nm = t(g["network_object_mask"])
    assert int((nm != ref_nm).sum()) == 0
    assert torch.allclose(dists, t(g["dists"]), rtol=1e-5, atol=1e-5)
    assert torch.allclose(pts, t(g["points"]), rtol=1e-5, atol=1e-5)
    assert cnt.total > 0


@pytest.mark.parametrize("name", ["cfg1_eval_w256", "cfg1_train_w256", "small_eval_w512", "train_phase0_w256",
                                  "cfg2_shape_eval_w512", "cfg3_shape_train_w512"])
def test_forward_golden(golden, name):
    g = golden(name)
    sd = preset_state_dict(str(g["meta_preset"]), g["meta_weights_sha"])
    sw, rw = O.sdf_weights(sd), O.render_weights(sd)
    scene = scene_from_meta(g)
    training = bool(int(g["meta_training"]))
    tp = float(g["meta_tp"]) if float(g["meta_tp"]) >= 0 else None
    kw = {}
    if training:
        kw = dict(steps01=t(g["steps01"]) if "steps01" in g else None, eik_points=t(g["eik_points"]))
        if "dsurf_jitter01" in g:       # phase 0: the reference's rand_like / np.random.choice draws
            kw["dsurf_rand"] = dict(jitter01=t(g["dsurf_jitter01"]), idx_on=g["dsurf_idx_on"], idx_jitter=g["dsurf_idx_jitter"])
    out = O.idr_forward(sw, rw, scene, tp, training, **kw)
    nm = out["network_object_mask"]
    assert torch.equal(nm, t(g["network_object_mask"]))
    assert torch.allclose(out["points"], t(g["points"]), rtol=1e-5, atol=1e-5)
    assert torch.allclose(out["rgb_values"], t(g["rgb_values"]), rtol=1e-4, atol=1e-5)
    assert torch.allclose(out["sdf_output"], t(g["sdf_output"]), rtol=1e-4, atol=2e-6)
    assert torch.allclose(out["diff_surf_pts"], t(g["diff_surf_pts"]), rtol=1e-5, atol=1e-5)
    losses = O.hot_path_losses(out, scene, 0.5 if tp is None else tp)
    assert rel_err(losses["rgb_loss"], g["rgb_loss"]) < 1e-5
    if tp is not None and tp < O.PHASE[0]:     # loss.py:196-199 switches the term off in phase 0; the fixture still holds its value
        losses["feat_loss"] = O.feat_loss_corr(out["diff_surf_pts"], scene["feat"], scene["cam"], scene["feat_src"],
                                               scene["src_cams"], scene["size"][:1], scene["center"][:1], nm, out["object_mask"])
    assert rel_err(losses["feat_loss"], g["feat_loss"]) < 1e-4
    if training:
        assert torch.allclose(out["grad_theta"], t(g["grad_theta"]), rtol=1e-4, atol=1e-5)
        assert rel_err(losses["eikonal_loss"], g["eikonal_loss"]) < 1e-4
        assert rel_err(losses["surf_loss"], g["surf_loss"]) < 1e-5
        if "depth_loss" in g:     # depth-carving term (loss.py:37-63) on the reference's own eikonal set
            dl = O.depth_loss(t(g["eikonal_points_hom_all"]), t(g["eikonal_output"]), scene["depths"], scene["depth_cams"],
                              scene["size"][:1], scene["center"][:1], tp)
            assert rel_err(dl, g["depth_loss"]) < 1e-5
            assert rel_err(losses["depth_loss"], g["depth_loss"]) < 1e-4
        if "eikonal_points_hom" in g:
            assert torch.allclose(out["eikonal_points_hom"], t(g["eikonal_points_hom"]), rtol=1e-5, atol=1e-6)
            assert torch.allclose(out["eikonal_output"], t(g["eikonal_output"]), rtol=1e-4, atol=2e-6)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree only exists in the build container")
def test_live_reference_train_step_matches_oracle():
    """Runs the unmodified reference and the restatement side by side on a fresh seed,
    including gradients of the hot-path losses w.r.t. the parameters."""
    ref = ref_shim.load()
    from mvsdf_b200 import synth
    sd = synth.make_state_dict(width=64, seed=5, perturb=0.05, pe_noise=0.003, bias=0.6)
    model = ref.idr.IDRNetwork(ref_shim.DictConf(ref_shim.model_conf(64)))
    model.load_state_dict(sd)
    scene = synth.make_scene(16, 16, n_images=2, n_src=2, n_rays=128, seed=9)
    model.train()
    torch.manual_seed(7)
    with ref_shim.quiet():
        r = model({k: scene[k].clone() for k in ["uv", "pose", "intrinsics", "object_mask"]}, 0.3)
        lm = ref.loss.IDRLoss()
        r_feat = lm.get_feat_loss_corr(r["diff_surf_pts"], None, scene["feat"], scene["cam"], scene["feat_src"],
                                       scene["src_cams"], scene["size"][:1], scene["center"][:1],
                                       r["network_object_mask"], r["object_mask"])
        r_rgb = lm.get_rgb_loss(r["rgb_values"], scene["rgb"], r["network_object_mask"], r["object_mask"])
        r_eik = lm.get_eikonal_loss(r["grad_theta"])
    (r_rgb + r_feat + r_eik).backward()
    ref_grads = {k: p.grad.clone() for k, p in model.named_parameters()}

    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    sw, rw = O.sdf_weights(params), O.render_weights(params)
    torch.manual_seed(7)
    o = O.idr_forward(sw, rw, scene, 0.3, True)
    losses = O.hot_path_losses(o, scene, 0.3)
    assert torch.equal(o["network_object_mask"], r["network_object_mask"])
    assert rel_err(losses["rgb_loss"], r_rgb) < 1e-5
    assert rel_err(losses["feat_loss"], r_feat) < 1e-5
    assert rel_err(losses["eikonal_loss"], r_eik) < 1e-5
    (losses["rgb_loss"] + losses["feat_loss"] + losses["eikonal_loss"]).backward()
    for k, gr in ref_grads.items():
        assert torch.allclose(params[k].grad, gr, rtol=1e-3, atol=1e-6), k


def _pose7_from_matrix(pose):
    """Test helper: cam->world [B,4,4] -> [B,7] (quaternion r,i,j,k ; centre), valid for trace(R) > -1."""
    R = pose[:, :3, :3].double()
    qr = torch.sqrt(1.0 + R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2]) / 2
    q = torch.stack([qr, (R[:, 2, 1] - R[:, 1, 2]) / (4 * qr), (R[:, 0, 2] - R[:, 2, 0]) / (4 * qr),
                     (R[:, 1, 0] - R[:, 0, 1]) / (4 * qr)], dim=1)
    return torch.cat([q, pose[:, :3, 3].double()], dim=1).float()


def test_quaternion_pose_is_the_matrix_pose():
    """Row a1, quaternion branch (rend_util.py:49-54): a [B,7] pose gives the rays of the equivalent 4x4 pose."""
    from mvsdf_b200 import synth
    scene = synth.make_scene(12, 12, n_images=3, n_src=1, seed=4)
    pose7 = _pose7_from_matrix(scene["pose"])
    d7, c7 = O.camera_rays(scene["uv"], pose7, scene["intrinsics"])
    d4, c4 = O.camera_rays(scene["uv"], scene["pose"], scene["intrinsics"])
    assert torch.allclose(c7, c4, atol=1e-6) and torch.allclose(d7, d4, atol=2e-6)
    if ref_shim.available():
        ref = ref_shim.load()
        rd, rc = ref.rend_util.get_camera_params(scene["uv"], pose7 * 1.7 * torch.tensor([1, 1, 1, 1, 1 / 1.7, 1 / 1.7, 1 / 1.7]),
                                                 scene["intrinsics"])          # un-normalised quaternion: normalised inside
        od, oc = O.camera_rays(scene["uv"], pose7 * 1.7 * torch.tensor([1, 1, 1, 1, 1 / 1.7, 1 / 1.7, 1 / 1.7]), scene["intrinsics"])
        assert torch.allclose(od, rd, atol=1e-6) and torch.allclose(oc, rc, atol=1e-6)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("training", [False, True])
def test_e_trace_counts_are_the_reference_tracers_own_evaluations(training):
    """bench.py's roofline counts algorithmic FLOPs as E_trace x FLOPs per SDF evaluation.  E_trace
    must be what the REFERENCE tracer evaluates: run the unmodified RayTracing.forward with a counting SDF closure and
    compare with the oracle's TraceCounters on the same rays (the GPU test then ties the kernels' counters to the oracle)."""
    ref = ref_shim.load()
    from mvsdf_b200 import synth
    sd = preset_state_dict("w256")
    sw = O.sdf_weights(sd)
    scene = synth.make_scene(20, 20, n_images=2, n_src=1, seed=4, mask_mode="disc" if training else "ones")
    dirs, cam = O.camera_rays(scene["uv"], scene["pose"], scene["intrinsics"])
    rt_conf = ref_shim.model_conf(256)["ray_tracer"]
    tracer = ref.ray_tracing.RayTracing(**rt_conf)
    tracer.train(training)
    n_evals = [0]

    def counting_sdf(x):
        n_evals[0] += x.shape[0]
        return O.sdf_mlp(x, sw)[:, 0]

    torch.manual_seed(3)
    with torch.no_grad(), ref_shim.quiet():
        r_pts, r_mask, r_dists = tracer(sdf=counting_sdf, cam_loc=cam, object_mask=scene["object_mask"].reshape(-1),
                                        ray_directions=dirs)
    torch.manual_seed(3)
    oc = O.TraceCounters()
    with torch.no_grad():
        o_pts, o_mask, o_dists = O.trace_rays(lambda x: O.sdf_mlp(x, sw)[:, 0], cam, scene["object_mask"].reshape(-1), dirs,
                                              training=training, counters=oc)
    assert torch.equal(o_mask, r_mask)
    assert torch.allclose(o_dists, r_dists, rtol=1e-6, atol=1e-6)
    # the reference evaluates the 100 sampler points of EVERY ray slot it gathered, in 100 000-point chunks: same count
    assert oc.total == n_evals[0], (oc, n_evals[0])


def test_featext_oracle_matches_live_reference():
    """Row f4: oracle/featext_oracle.py against the unmodified FeatExt (my_utils.py:693-708) with the checkpoint the reference
    ships (utils/vismvsnet.pt) -- only where /root/reference exists."""
    import os
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present")
    from oracle import featext_oracle as FO
    ref = ref_shim.load()
    cwd, orig_load = os.getcwd(), torch.load
    try:
        os.chdir(ref_shim.REFERENCE_CODE)
        torch.load = lambda *a, **k: orig_load(*a, **{**k, "map_location": "cpu", "weights_only": False})
        fe = ref.my_utils.FeatExt()
    finally:
        torch.load = orig_load
        os.chdir(cwd)
    fe.eval()
    x = torch.randn(1, 3, 40, 56, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        want = fe(x)
        got = FO.featext_forward(fe.state_dict(), x)
    for a, b in zip(got, want):
        assert (a - b).abs().max().item() <= 1e-5 * max(1.0, b.abs().max().item())


def test_live_reference_pose_gradients_match_oracle():
    """train_cameras=True (utils/rend_util.py:49-57, model/sample_network.py:15-19): gradients of the hot-path losses w.r.t. a
    [B,7] quaternion pose, reference vs restatement -- the pin for B200IDRNetwork's pose-gradient path (tests/test_gpu_autograd.py)."""
    ref = ref_shim.load()
    from mvsdf_b200 import synth
    sd = synth.make_state_dict(width=64, seed=5, perturb=0.05, pe_noise=0.003, bias=0.6)
    model = ref.idr.IDRNetwork(ref_shim.DictConf(ref_shim.model_conf(64)))
    model.load_state_dict(sd)
    scene = synth.make_scene(16, 16, n_images=2, n_src=2, n_rays=128, seed=9)
    pose7 = _pose7_from_matrix(scene["pose"])
    model.train()
    torch.manual_seed(7)
    p_ref = pose7.clone().requires_grad_(True)
    with ref_shim.quiet():
        r = model({"uv": scene["uv"].clone(), "pose": p_ref, "intrinsics": scene["intrinsics"].clone(),
                   "object_mask": scene["object_mask"].clone()}, 0.3)
        lm = ref.loss.IDRLoss()
        r_feat = lm.get_feat_loss_corr(r["diff_surf_pts"], None, scene["feat"], scene["cam"], scene["feat_src"],
                                       scene["src_cams"], scene["size"][:1], scene["center"][:1],
                                       r["network_object_mask"], r["object_mask"])
        r_rgb = lm.get_rgb_loss(r["rgb_values"], scene["rgb"], r["network_object_mask"], r["object_mask"])
        r_eik = lm.get_eikonal_loss(r["grad_theta"])
    (r_rgb + r_feat + r_eik).backward()
    torch.manual_seed(7)
    p_o = pose7.clone().requires_grad_(True)
    inp = dict(scene)
    inp["pose"] = p_o
    o = O.idr_forward(O.sdf_weights(sd), O.render_weights(sd), inp, 0.3, True)
    losses = O.hot_path_losses(o, scene, 0.3)
    (losses["rgb_loss"] + losses["feat_loss"] + losses["eikonal_loss"]).backward()
    assert p_ref.grad is not None and float(p_ref.grad.abs().max()) > 0
    assert torch.allclose(p_o.grad, p_ref.grad, rtol=2e-3, atol=1e-6 * float(p_ref.grad.abs().max())), (p_o.grad, p_ref.grad)


def test_schedule_switches_match_live_reference():
    """model/conf.py's eight point-set switches (d_use_* / eik_use_*, implicit_differentiable_renderer.py:259-286) in a
    combination the shipped schedule never produces: the oracle's restatement against the reference with its conf module
    patched (the reference's own override mechanism is IDR_USE_ENV=1 + IDR_CONF=<module>, :15-17)."""
    import types
    import numpy as np
    ref = ref_shim.load()
    from mvsdf_b200 import conf as shipped, synth
    sd = synth.make_state_dict(width=64, seed=5, perturb=0.05, pe_noise=0.003, bias=0.6)
    model = ref.idr.IDRNetwork(ref_shim.DictConf(ref_shim.model_conf(64)))
    model.load_state_dict(sd)
    scene = synth.make_scene(32, 32, n_images=2, n_src=1, n_rays=96, seed=9)
    tp = 0.3
    override = dict(d_use_eik=lambda t: False, d_use_dsurf_on=lambda t: True, eik_use_rt_surf=lambda t: False,
                    eik_use_dsurf_jitter=lambda t: True)
    sched = types.SimpleNamespace(**{k: getattr(shipped, k) for k in dir(shipped) if not k.startswith("_")})
    for k, v in override.items():
        setattr(sched, k, v)
    saved = {k: getattr(ref.idr.conf, k) for k in override}
    model.train()
    try:
        for k, v in override.items():
            setattr(ref.idr.conf, k, v)
        torch.manual_seed(7)
        np.random.seed(7)
        with ref_shim.quiet():
            r = model({k: scene[k].clone() for k in ["uv", "pose", "intrinsics", "object_mask", "depths", "depth_cams", "center", "size"]}, tp)
    finally:
        for k, v in saved.items():
            setattr(ref.idr.conf, k, v)
    torch.manual_seed(7)
    np.random.seed(7)
    o = O.idr_forward(O.sdf_weights(sd), O.render_weights(sd), scene, tp, True, schedule=sched)
    assert torch.equal(o["network_object_mask"], r["network_object_mask"])
    for k in ("eikonal_output", "eikonal_points_hom", "grad_theta", "surf_indicator_output"):
        assert o[k].shape == r[k].shape, (k, o[k].shape, r[k].shape)
        assert torch.allclose(o[k], r[k], rtol=1e-4, atol=1e-5), k
    n_hit = int((r["network_object_mask"] & r["object_mask"]).sum())
    n_eik = 96
    assert r["eikonal_output"].shape[1] == n_hit + n_eik                   # rt-surface + dsurf-on
    assert r["grad_theta"].shape[0] == n_eik + n_eik                      # eikonal samples + dsurf-jitter
