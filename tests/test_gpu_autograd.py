"""GPU: a training step through B200IDRNetwork / B200IDRLoss with autograd enabled -- loss values from the native
kernels, parameter gradients compared with PyTorch autograd through the oracle restatement of the reference
(oracle/mvsdf_oracle.py, itself pinned to the reference's gradients by tests/test_oracle.py::test_live_reference_*)."""
import pytest
import torch

from mvsdf_b200 import synth
from mvsdf_b200.loss import B200IDRLoss
from mvsdf_b200.network import B200IDRNetwork, default_conf
from oracle import mvsdf_oracle as O
from tests.helpers import gate

pytestmark = pytest.mark.gpu

# limits = ~3x the errors measured on the B200 (gpurun_out/gate_report.json)
G_LOSS_REL = 2e-5           # measured 4.0e-6
G_PARAM_GRAD = 5e-4         # measured 1.45e-4 (fraction of the tensor's max |gradient|)

IN = ["uv", "pose", "intrinsics", "object_mask", "depths", "depth_cams", "center", "size"]
GT = ["rgb", "feat", "cam", "feat_src", "src_cams", "size", "center", "depths", "depth_cams"]


@pytest.mark.parametrize("tp", [0.3, 0.1])
def test_parameter_gradients_match_oracle_autograd(tp):
    dev = torch.device("cuda:0")
    sd = synth.make_state_dict(width=256, seed=1, perturb=0.05, pe_noise=0.003, bias=0.6)
    scene = synth.make_scene(64, 64, n_images=2, n_src=2, n_rays=160, seed=6)
    g = torch.Generator().manual_seed(5)
    steps = torch.rand(100, generator=g)
    eik = torch.rand(160, 3, generator=g) * 2 - 1

    # --- oracle (CPU, fp32 autograd incl. the second-order terms)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    rnd = None
    if tp < O.PHASE[0]:
        ds = O.depth_surface_points(scene["depths"], scene["depth_cams"], scene["center"][:1], scene["size"][:1])
        torch.manual_seed(11)
        import numpy as np
        np.random.seed(11)
        _, _, rnd = O.depth_surface_samples(ds, 160, 1.0)
    ref = O.idr_forward(O.sdf_weights(params), O.render_weights(params), scene, tp, True, steps01=steps, eik_points=eik,
                        dsurf_rand=rnd)
    rl = O.hot_path_losses(ref, scene, tp)
    # the reference's total (loss.py:206-210); the surface-indicator term is switched off in phase 0 (:201-204)
    surf_w = 0.01 if tp >= O.PHASE[0] else 0.0
    ref_total = (0.5 * rl["rgb_loss"] + 0.1 * rl["eikonal_loss"] + surf_w * rl["surf_loss"]
                 + O.feat_weight(tp) * rl["feat_loss"].sum() + rl["depth_loss"])
    ref_total.backward()

    # --- product path
    model = B200IDRNetwork(default_conf(256)).to(dev)
    model.load_state_dict(sd)
    model.train()
    out = model({k: scene[k].to(dev) for k in IN}, tp, steps01=steps, eik_points=eik,
                dsurf_rand=None if rnd is None else {k: (v.to(dev) if k == "jitter01" else v.cpu().numpy()) for k, v in rnd.items()})
    assert out["rgb_values"].requires_grad and out["grad_theta"].requires_grad
    if int((out["network_object_mask"].cpu() != ref["network_object_mask"]).sum()) != 0:
        pytest.skip("a discrete tracer decision flipped on this input; gradient comparison is not meaningful")
    ls = B200IDRLoss()(out, {k: scene[k].to(dev) for k in GT}, tp, 2)          # the call of idr_train.py:269
    for k in ("rgb_loss", "eikonal_loss", "depth_loss"):
        gate(k + "_rel", abs(float(ls[k]) - float(rl[k])) / max(1.0, abs(float(rl[k]))), G_LOSS_REL)
    gate("total_loss_rel", abs(float(ls["loss"]) - float(ref_total)) / abs(float(ref_total)), G_LOSS_REL)
    ls["loss"].sum().backward()
    worst = 0.0
    for name, p in model.named_parameters():
        gr = params[name].grad
        assert p.grad is not None, name
        scale = gr.abs().max().item() + 1e-8
        err = (p.grad.cpu() - gr).abs().max().item() / scale
        worst = max(worst, err)
    gate("param_grad_rel_of_max", worst, G_PARAM_GRAD)


def test_no_grad_training_forward_keeps_native_path():
    dev = torch.device("cuda:0")
    sd = synth.make_state_dict(width=256, seed=1, perturb=0.05, pe_noise=0.003, bias=0.6)
    scene = synth.make_scene(24, 24, n_images=1, n_src=1, n_rays=200, seed=2)
    model = B200IDRNetwork(default_conf(256)).to(dev)
    model.load_state_dict(sd)
    model.train()
    g = torch.Generator().manual_seed(5)
    steps, eik = torch.rand(100, generator=g), torch.rand(100, 3, generator=g) * 2 - 1
    inp = {k: scene[k].to(dev) for k in IN}
    with torch.no_grad():
        a = model(inp, 0.5, steps01=steps, eik_points=eik)
    b = model(inp, 0.5, steps01=steps, eik_points=eik)
    assert not a["rgb_values"].requires_grad and b["rgb_values"].requires_grad
    # same kernels, same values: the graph-building path evaluates the surface points with the same fused kernel
    assert torch.equal(a["network_object_mask"], b["network_object_mask"])
    assert (a["rgb_values"] - b["rgb_values"]).abs().max().item() < 1e-5
    assert (a["grad_theta"] - b["grad_theta"]).abs().max().item() < 1e-5
    assert (a["diff_surf_pts"] - b["diff_surf_pts"]).abs().max().item() < 1e-6


@pytest.mark.parametrize("n_src,hw,seed", [(1, 32, 3), (4, 48, 8), (8, 40, 9)])
def test_native_feat_loss_backward_vs_oracle_autograd(n_src, hw, seed):
    """mvsdf_feat_loss_backward (d loss / d diff_surf_pts: projections with the eps-guarded divisions, bilinear tap
    derivatives with zero padding, cosine similarity, the in-range / <0.5 masks) against PyTorch autograd through the
    oracle's get_feat_loss_corr restatement, in fp64 so that the comparison is not limited by the reference's own rounding.
    Points are random inside the unit ball: many project outside some views, some terms are dropped by the 0.5 gate."""
    dev = torch.device("cuda:0")
    B = 2
    scene = synth.make_scene(hw, hw, n_images=B, n_src=n_src, seed=seed)
    g = torch.Generator().manual_seed(seed)
    counts = [137, 91]
    pts = torch.randn(sum(counts), 3, generator=g)
    pts = 0.6 * pts / pts.norm(dim=1, keepdim=True) * torch.rand(sum(counts), 1, generator=g) ** (1 / 3)
    N = scene["uv"].shape[1]
    mask = torch.zeros(B, N, dtype=torch.bool)
    for i, c in enumerate(counts):
        mask[i, :c] = True
    mask = mask.reshape(-1)
    # oracle, fp64 autograd
    p64 = pts.double().requires_grad_(True)
    ref = O.feat_loss_corr(p64, scene["feat"].double(), scene["cam"].double(), scene["feat_src"].double(),
                           scene["src_cams"].double(), scene["size"][:1].double(), scene["center"][:1].double(), mask, mask)
    ref = ref.sum() if ref.dim() else ref
    (g_ref,) = torch.autograd.grad(ref * 3.0, p64)
    # product path: forward + native backward through the autograd Function
    loss_mod = B200IDRLoss()
    pd = pts.to(dev).requires_grad_(True)
    offs = torch.tensor([0, counts[0], sum(counts)], dtype=torch.int32, device=dev)
    out = loss_mod.get_feat_loss_corr(pd, None, scene["feat"].to(dev), scene["cam"].to(dev), scene["feat_src"].to(dev),
                                      scene["src_cams"].to(dev), scene["size"][:1].to(dev), scene["center"][:1].to(dev),
                                      mask.to(dev), mask.to(dev), hit_offsets=offs)
    assert abs(float(out) - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))
    (out * 3.0).backward()
    got = pd.grad.cpu().double()
    scale = g_ref.abs().max().item()
    assert scale > 0.0
    row_err = (got - g_ref).abs().max(dim=1).values / scale
    # a point exactly on a gate (in-range edge, |1-corr| = 0.5) may be kept in fp64 and dropped in fp32: allow one
    bad = row_err > 2e-4
    assert int(bad.sum()) <= 1, f"{int(bad.sum())} rows differ, worst {row_err.max().item():.3e}"
    nz = (g_ref.abs().sum(dim=1) > 0)
    assert int(((got.abs().sum(dim=1) > 0) != nz).sum()) <= 1, "kept / dropped terms differ from the oracle"


def test_pose_gradients_match_oracle_autograd():
    """train_cameras=True (training/idr_train.py:121-127: pose_vecs = nn.Embedding(n_images, 7)): the gradient of the training
    loss w.r.t. a [B,7] quaternion pose -- through x_s / x_diff (SdfEval's dx), the view direction (RenderEval's d_view) and the
    feature-warp backward -- against autograd through the oracle, which tests/test_oracle.py pins to the live reference."""
    dev = torch.device("cuda:0")
    tp = 0.3
    sd = synth.make_state_dict(width=256, seed=1, perturb=0.05, pe_noise=0.003, bias=0.6)
    scene = synth.make_scene(64, 64, n_images=2, n_src=2, n_rays=160, seed=6)
    P = scene["pose"]
    R = P[:, :3, :3].double()
    qr = torch.sqrt(1.0 + R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2]) / 2
    q = torch.stack([qr, (R[:, 2, 1] - R[:, 1, 2]) / (4 * qr), (R[:, 0, 2] - R[:, 2, 0]) / (4 * qr), (R[:, 1, 0] - R[:, 0, 1]) / (4 * qr)], dim=1)
    pose7 = torch.cat([q, P[:, :3, 3].double()], dim=1).float()
    g = torch.Generator().manual_seed(5)
    steps = torch.rand(100, generator=g)
    eik = torch.rand(160, 3, generator=g) * 2 - 1
    # oracle
    p_o = pose7.clone().requires_grad_(True)
    inp = dict(scene)
    inp["pose"] = p_o
    ref = O.idr_forward(O.sdf_weights(sd), O.render_weights(sd), inp, tp, True, steps01=steps, eik_points=eik)
    rl = O.hot_path_losses(ref, scene, tp)
    ref_total = (0.5 * rl["rgb_loss"] + 0.1 * rl["eikonal_loss"] + 0.01 * rl["surf_loss"] + O.feat_weight(tp) * rl["feat_loss"].sum()
                 + rl["depth_loss"])
    ref_total.backward()
    # product path
    model = B200IDRNetwork(default_conf(256)).to(dev)
    model.load_state_dict(sd)
    model.train()
    p_g = pose7.clone().to(dev).requires_grad_(True)
    din = {k: scene[k].to(dev) for k in IN}
    din["pose"] = p_g
    out = model(din, tp, steps01=steps, eik_points=eik)
    if int((out["network_object_mask"].cpu() != ref["network_object_mask"]).sum()) != 0:
        pytest.skip("a discrete tracer decision flipped on this input; gradient comparison is not meaningful")
    ls = B200IDRLoss()(out, {k: scene[k].to(dev) for k in GT}, tp, 2)
    gate("total_loss_rel", abs(float(ls["loss"]) - float(ref_total)) / abs(float(ref_total)), G_LOSS_REL)
    ls["loss"].sum().backward()
    assert p_g.grad is not None
    scale = p_o.grad.abs().max().item()
    assert scale > 0
    gate("pose_grad_rel_of_max", (p_g.grad.cpu() - p_o.grad).abs().max().item() / scale, 2e-3)
    # the parameters still receive their gradients alongside
    assert all(p.grad is not None for p in model.parameters())
