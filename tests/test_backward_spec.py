"""oracle/backward_spec.py (the explicit second-order reverse sweep a fused f1 kernel has to execute) against PyTorch
autograd through the oracle restatement of ImplicitNetwork.forward + .gradient (create_graph=True, exactly how the
reference gets these gradients, implicit_differentiable_renderer.py:96-107).  fp64, so that agreement is to rounding."""
import pytest
import torch

from mvsdf_b200 import synth
from oracle import backward_spec as B
from oracle import mvsdf_oracle as O


@pytest.mark.parametrize("width,seed", [(64, 0), (96, 1)])
def test_explicit_reverse_sweep_matches_autograd(width, seed):
    torch.manual_seed(seed)
    sd = synth.make_state_dict(width=width, seed=seed, perturb=0.05, pe_noise=0.003, bias=0.6)
    sd = {k: v.double() for k, v in sd.items()}
    n_lin = 9
    names = [f"implicit_network.lin{l}" for l in range(n_lin)]
    params = {k: sd[k].clone().requires_grad_(True) for n in names for k in (n + ".weight_v", n + ".weight_g", n + ".bias")}
    P = 37
    x = (torch.rand(P, 3, dtype=torch.float64) * 2 - 1).requires_grad_(True)
    # points on both sides of the softplus knee and beyond its linear threshold are all present at beta = 100
    w = O.weights_from_state_dict(params, "implicit_network", skip_in=(4,), n_freqs=6)
    full = O.sdf_mlp(x, w)
    grad = O.sdf_gradient(x, w, create_graph=True) if not x.requires_grad else \
        torch.autograd.grad(full[:, 0].sum(), x, create_graph=True)[0]
    g_full = torch.randn_like(full)
    g_grad = torch.randn(P, 3, dtype=torch.float64)
    loss = (full * g_full).sum() + (grad * g_grad).sum()
    wanted = [x] + [params[n + s] for n in names for s in (".weight_v", ".weight_g", ".bias")]
    ref = torch.autograd.grad(loss, wanted)

    with torch.no_grad():
        dx, dv, dg, db = B.sdf_value_grad_backward(
            x.detach(), [sd[n + ".weight_v"] for n in names], [sd[n + ".weight_g"] for n in names],
            [sd[n + ".bias"] for n in names], skip_in=(4,), n_freqs=6, g_full=g_full, g_grad=g_grad)

    def close(a, b, what):
        scale = b.abs().max().item() + 1e-30
        err = (a - b).abs().max().item() / scale
        assert err < 1e-9, f"{what}: relative error {err:.2e}"

    close(dx, ref[0], "dx")
    for l in range(n_lin):
        close(dv[l], ref[1 + 3 * l], f"lin{l}.weight_v")
        close(dg[l], ref[2 + 3 * l], f"lin{l}.weight_g")
        close(db[l], ref[3 + 3 * l], f"lin{l}.bias")


def test_render_reverse_sweep_matches_autograd():
    torch.manual_seed(2)
    sd = {k: v.double() for k, v in synth.make_state_dict(width=64, seed=3, perturb=0.05, pe_noise=0.003, bias=0.6).items()}
    names = [f"rendering_network.lin{l}" for l in range(5)]
    params = {k: sd[k].clone().requires_grad_(True) for n in names for k in (n + ".weight_v", n + ".weight_g", n + ".bias")}
    P = 29
    ins = [torch.randn(P, d, dtype=torch.float64, requires_grad=True) for d in (3, 3, 3, 256)]     # points, normals, view, feats
    w = O.weights_from_state_dict(params, "rendering_network", skip_in=(), n_freqs=4)
    rgb = O.render_mlp(ins[0], ins[1], ins[2], ins[3], w)
    g_rgb = torch.randn_like(rgb)
    wanted = ins + [params[n + s] for n in names for s in (".weight_v", ".weight_g", ".bias")]
    ref = torch.autograd.grad((rgb * g_rgb).sum(), wanted)
    with torch.no_grad():
        d_pts, d_nrm, d_view, d_feat, dv, dg, db = B.render_backward(
            ins[0].detach(), ins[1].detach(), ins[2].detach(), ins[3].detach(), [sd[n + ".weight_v"] for n in names],
            [sd[n + ".weight_g"] for n in names], [sd[n + ".bias"] for n in names], 4, g_rgb)

    def close(a, b, what):
        err = (a - b).abs().max().item() / (b.abs().max().item() + 1e-30)
        assert err < 1e-9, f"{what}: relative error {err:.2e}"

    for got, want, what in ((d_pts, ref[0], "points"), (d_nrm, ref[1], "normals"), (d_view, ref[2], "view"), (d_feat, ref[3], "feats")):
        close(got, want, what)
    for l in range(5):
        close(dv[l], ref[4 + 3 * l], f"lin{l}.weight_v")
        close(dg[l], ref[5 + 3 * l], f"lin{l}.weight_g")
        close(db[l], ref[6 + 3 * l], f"lin{l}.bias")
