"""GPU: the native backward of the two MLPs (csrc/mlp_bwd_kernel.cuh, csrc/train_abi.cu) against
the explicit reverse chain of oracle/backward_spec.py in fp64 -- itself pinned to autograd through the oracle restatement
of the reference (tests/test_backward_spec.py) -- and the fused Adam step against torch.optim.Adam + clip_grad_norm_.
Everything goes through the C ABI (mvsdf_sdf_forward_train / _backward, mvsdf_render_*, mvsdf_weight_grads, mvsdf_adam_step)."""
import pytest
import torch

from mvsdf_b200 import ops, synth
from oracle import backward_spec as S
from tests.helpers import gate

pytestmark = pytest.mark.gpu

# limits: fraction of the tensor's max |gradient|; fp16 hi/lo operands (22 bits), fp32 accumulation
G_DX = 2e-4
G_PARAM = 2e-4


def _sdf_lists(sd, n_lin=9, dtype=torch.float64, device="cpu"):
    vs = [sd[f"implicit_network.lin{l}.weight_v"].to(device=device, dtype=dtype) for l in range(n_lin)]
    gs = [sd[f"implicit_network.lin{l}.weight_g"].to(device=device, dtype=dtype) for l in range(n_lin)]
    bs = [sd[f"implicit_network.lin{l}.bias"].to(device=device, dtype=dtype) for l in range(n_lin)]
    return vs, gs, bs


def _rel(a, b):
    return (a.double().cpu() - b).abs().max().item() / (b.abs().max().item() + 1e-30)


# n >= 4736 (296 tiles of 16 points = two per SM): the CTA-pair sweep (mlp_bwd_pair_kernel.cuh); 5003 points = an odd tile count
@pytest.mark.parametrize("width,n,use_full,use_grad", [(256, 300, True, True), (512, 1000, True, True), (512, 17, True, False),
                                                        (256, 4097, False, True), (512, 4760, True, True), (256, 5003, True, True)])
def test_sdf_backward_vs_explicit_chain(width, n, use_full, use_grad):
    dev = torch.device("cuda:0")
    sd = synth.make_state_dict(width=width, seed=1, perturb=0.05, pe_noise=0.003, bias=0.6)
    net = ops.PackedNet("sdf", width, 8).pack_state_dict(sd, "implicit_network", dev)
    g = torch.Generator().manual_seed(n)
    x = torch.rand(n, 3, generator=g) * 1.6 - 0.8
    F = 256
    g_full = torch.randn(n, F + 2, generator=g) * 1e-3 if use_full else None
    g_grad = torch.randn(n, 3, generator=g) * 1e-2 if use_grad else None
    vs, gs, bs = _sdf_lists(sd)
    zf = torch.zeros(n, F + 2, dtype=torch.float64)
    zg = torch.zeros(n, 3, dtype=torch.float64)
    dx_ref, dv_ref, dg_ref, db_ref = S.sdf_value_grad_backward(x.double(), vs, gs, bs, (4,), 6,
                                                               g_full.double() if use_full else zf,
                                                               g_grad.double() if use_grad else zg)
    full, grad, save = ops.sdf_forward_train(net, x.to(dev))
    full0, grad0 = ops.sdf_value_grad(net, x.to(dev), ops.HEAD_FULL)
    gate("forward_train_vs_forward_full", (full - full0).abs().max().item(), 5e-6)
    gate("forward_train_vs_forward_grad", (grad - grad0).abs().max().item(), 5e-5)
    dx, dw, db = ops.sdf_backward(net, x.to(dev), save, None if g_full is None else g_full.to(dev),
                                  None if g_grad is None else g_grad.to(dev), need_dx=True)
    gate("dx_rel_of_max", _rel(dx, dx_ref), G_DX)
    v32, g32, _ = _sdf_lists(sd, dtype=torch.float32, device=dev)
    dvs, dgs, dbs = ops.weight_grads(net, dw, db, v32, g32)
    worst = 0.0
    for l in range(9):
        for name, got, ref in (("dv", dvs[l], dv_ref[l]), ("dg", dgs[l], dg_ref[l]), ("db", dbs[l], db_ref[l])):
            if ref.abs().max().item() == 0.0:
                assert got.abs().max().item() == 0.0, (l, name)
                continue
            e = _rel(got.reshape(ref.shape), ref)
            worst = max(worst, e)
            assert e < 5 * G_PARAM, f"layer {l} {name}: {e:.3e}"
    gate("param_grad_rel_of_max", worst, G_PARAM)
    # the saved activations can be swept more than once (the surface set gets two sweeps per step)
    dx2, dw2, db2 = ops.sdf_backward(net, x.to(dev), save, None if g_full is None else g_full.to(dev),
                                     None if g_grad is None else g_grad.to(dev), need_dx=True)
    # (dx and dW are accumulated with floating-point atomics: equal up to summation order)
    assert (dx - dx2).abs().max().item() <= 1e-5 * dx.abs().max().item()
    assert (dw - dw2).abs().max().item() <= 1e-5 * dw.abs().max().item()


# n >= 18944 (296 tiles of 64 points): the CTA-pair sweep; 19001 points = 297 tiles
@pytest.mark.parametrize("width,n", [(256, 200), (512, 1000), (512, 65), (256, 19001)])
def test_render_backward_vs_explicit_chain(width, n):
    dev = torch.device("cuda:0")
    sd = synth.make_state_dict(width=width, seed=2, perturb=0.05, pe_noise=0.003, bias=0.6)
    net = ops.PackedNet("render", width, 4, n_freqs=4).pack_state_dict(sd, "rendering_network", dev)
    g = torch.Generator().manual_seed(n + 1)
    pts = torch.rand(n, 3, generator=g) - 0.5
    nrm = torch.randn(n, 3, generator=g)
    view = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=1)
    feats = torch.randn(n, 256, generator=g) * 0.5
    g_rgb = torch.randn(n, 3, generator=g) * 1e-3
    vs = [sd[f"rendering_network.lin{l}.weight_v"].double() for l in range(5)]
    gs = [sd[f"rendering_network.lin{l}.weight_g"].double() for l in range(5)]
    bs = [sd[f"rendering_network.lin{l}.bias"].double() for l in range(5)]
    dp_ref, dn_ref, dview_ref, df_ref, dv_ref, dg_ref, db_ref = S.render_backward(pts.double(), nrm.double(), view.double(), feats.double(),
                                                                         vs, gs, bs, 4, g_rgb.double())
    rgb, save = ops.render_forward_train(net, pts.to(dev), view.to(dev), nrm.to(dev), feats.to(dev))
    rgb0 = ops.render_forward(net, pts.to(dev), view.to(dev), nrm.to(dev), feats.to(dev))
    gate("render_forward_train_vs_forward", (rgb - rgb0).abs().max().item(), 5e-6)
    d_points, d_normals, d_feats, d_view, dw, db = ops.render_backward(net, save, rgb, g_rgb.to(dev), view.to(dev))
    # ReLU kinks: a pre-activation within fp32 rounding of zero (a few per million units) has a different sign in the fp64
    # chain than in the fp32 forward, which switches one hidden unit of one point on / off -- an O(1/sqrt(width)) change of
    # that point's input gradients.  Per-point errors are therefore gated on the 99th percentile, the outliers are counted.
    flipped = 0
    for name, got, ref in (("d_points", d_points, dp_ref), ("d_normals", d_normals, dn_ref), ("d_feats", d_feats, df_ref),
                           ("d_view", d_view, dview_ref)):
        row = (got.double().cpu() - ref).abs().max(dim=1).values / ref.abs().max().item()
        gate(name + "_rel_q99", torch.quantile(row, 0.99).item(), G_DX)
        gate(name + "_rows_above_gate", int((row > G_DX).sum()), max(1, n // 100))
        flipped = max(flipped, int((row > G_DX).sum()))
    dvs, dgs, dbs = ops.weight_grads(net, dw, db, [v.float().to(dev) for v in vs], [x.float().to(dev) for x in gs])
    # parameter gradients are sums over the points: a ReLU-kink flip (see above) moves the entries of the affected units by
    # one point's contribution.  Without flips (small n) every entry is gated; with them the bulk (median) and the energy of
    # the difference (Frobenius norm) are.
    worst, worst_med, worst_fro = 0.0, 0.0, 0.0
    for l in range(5):
        for got, ref in ((dvs[l], dv_ref[l]), (dgs[l], dg_ref[l]), (dbs[l], db_ref[l])):
            diff = (got.reshape(ref.shape).double().cpu() - ref)
            e = diff.abs().flatten() / ref.abs().max().item()
            worst = max(worst, e.max().item())
            worst_med = max(worst_med, e.median().item())
            worst_fro = max(worst_fro, (diff.norm() / ref.norm()).item())
    if n < 300:
        gate("param_grad_rel_of_max", worst, G_PARAM)
    # a flipped unit of one point changes that point's contribution to EVERY entry of the layers below it (measured: 4-5
    # flipped points of 19 001 move the median to 1e-4 of max); the kernel itself is pinned at this size by
    # test_pair_sweep_matches_single_cta_sweep (input gradients bit-identical to the single-CTA sweep)
    gate("param_grad_rel_of_max_median", worst_med, G_PARAM if flipped > 1 else G_PARAM / 4, note=f"{flipped} points with a ReLU-kink flip")
    gate("param_grad_rel_frobenius", worst_fro, 2e-2)


@pytest.mark.parametrize("kind,n", [("sdf", 5003), ("sdf", 4736), ("render", 19001)])
def test_pair_sweep_matches_single_cta_sweep(kind, n, monkeypatch):
    """The two reverse-sweep kernels (single CTA per 64-column tile / CTA pair per 128 columns) run the same arithmetic on the
    same saved activations; only the order of the floating-point atomics (dx, bias gradients, split-K dW) differs."""
    dev = torch.device("cuda:0")
    width = 512
    g = torch.Generator().manual_seed(n)
    if kind == "sdf":
        sd = synth.make_state_dict(width=width, seed=1, perturb=0.05, pe_noise=0.003, bias=0.6)
        net = ops.PackedNet("sdf", width, 8).pack_state_dict(sd, "implicit_network", dev)
        x = (torch.rand(n, 3, generator=g) * 1.6 - 0.8).to(dev)
        g_full = (torch.randn(n, 258, generator=g) * 1e-3).to(dev)
        g_grad = (torch.randn(n, 3, generator=g) * 1e-2).to(dev)
        _, _, save = ops.sdf_forward_train(net, x)
        run = lambda: ops.sdf_backward(net, x, save, g_full, g_grad, need_dx=True)
    else:
        sd = synth.make_state_dict(width=width, seed=2, perturb=0.05, pe_noise=0.003, bias=0.6)
        net = ops.PackedNet("render", width, 4, n_freqs=4).pack_state_dict(sd, "rendering_network", dev)
        pts = (torch.rand(n, 3, generator=g) - 0.5).to(dev)
        nrm = torch.randn(n, 3, generator=g).to(dev)
        view = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=1).to(dev)
        feats = (torch.randn(n, 256, generator=g) * 0.5).to(dev)
        g_rgb = (torch.randn(n, 3, generator=g) * 1e-3).to(dev)
        rgb, save = ops.render_forward_train(net, pts, view, nrm, feats)
        run = lambda: ops.render_backward(net, save, rgb, g_rgb, view)
    monkeypatch.setenv("MVSDF_PAIR_SWEEP", "0")
    single = [t.clone() for t in run() if t is not None]
    monkeypatch.setenv("MVSDF_PAIR_SWEEP", "1")
    pair = [t.clone() for t in run() if t is not None]
    assert len(single) == len(pair)
    for i, (a, b) in enumerate(zip(single, pair)):
        assert torch.isfinite(b).all()
        gate(f"pair_vs_single_sweep[{kind},{i}]", (a - b).abs().max().item() / (a.abs().max().item() + 1e-30), 2e-5)


def test_fused_adam_matches_torch_adam_with_clipping():
    from mvsdf_b200.optim import B200Adam
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    shapes = [(512, 39), (512, 1), (512,), (473, 512), (3, 512), (3,)]
    base = [torch.randn(*s, generator=g) for s in shapes]
    ours = [torch.nn.Parameter(b.clone().to(dev)) for b in base]
    theirs = [torch.nn.Parameter(b.clone().to(dev)) for b in base]
    opt_a = B200Adam(ours, lr=2e-4 * 8)
    opt_b = torch.optim.Adam(theirs, lr=2e-4 * 8)
    for step in range(4):
        grads = [torch.randn(*s, generator=g).to(dev) * (3.0 if step % 2 else 0.01) for s in shapes]
        for p, q, gr in zip(ours, theirs, grads):
            p.grad = gr.clone()
            q.grad = gr.clone()
        total = torch.cat([gr.flatten() for gr in grads]).norm()
        torch.nn.utils.clip_grad_norm_(theirs, 2.0)
        opt_b.step()
        opt_a.step(max_grad_norm=2.0)
        assert abs(float(opt_a.grad_norm) - float(total)) < 1e-4 * float(total)
        for p, q in zip(ours, theirs):
            assert torch.allclose(p, q, rtol=1e-5, atol=1e-6), step
