"""GPU parity of the fused tcgen05 MLP tile kernels against the CPU oracle and the golden
outputs of the reference (rows a3/a4/a5/a14 of SURVEY.md section 8).  All calls go through the C ABI."""
import pytest
import torch

from oracle import mvsdf_oracle as O
from mvsdf_b200 import synth
from tests.helpers import preset_state_dict, t

pytestmark = pytest.mark.gpu

TOL_SDF = 5e-5      # absolute, on O(1) SDF values (tensor-core fp32 accumulation truncates): inside the 1e-4 relative depth gate
TOL_GRAD = 1e-4
TOL_RGB = 1e-4


def _nets(preset, device):
    from mvsdf_b200 import ops
    from tests.helpers import WEIGHT_PRESETS
    sd = preset_state_dict(preset)
    w = WEIGHT_PRESETS[preset]["width"]
    sdf = ops.PackedNet("sdf", w, 8).pack_state_dict(sd, "implicit_network", device)
    rend = ops.PackedNet("render", w, 4, n_freqs=4).pack_state_dict(sd, "rendering_network", device)
    return sd, sdf, rend


@pytest.mark.parametrize("name", ["mlp_w256", "mlp_w512"])
def test_mlp_vs_reference_golden(golden, name):
    from mvsdf_b200 import ops
    g = golden(name)
    dev = torch.device("cuda:0")
    sd, sdf, rend = _nets(str(g["meta_preset"]), dev)
    x = t(g["x"]).to(dev)
    full = ops.sdf_forward(sdf, x, ops.HEAD_FULL).cpu()
    ref = t(g["sdf_full"])
    assert (full - ref).abs().max().item() < TOL_SDF, (full - ref).abs().max().item()
    s = ops.sdf_forward(sdf, x, ops.HEAD_SDF_ONLY).cpu()
    assert (s - ref[:, 0]).abs().max().item() < TOL_SDF
    full2, grad = ops.sdf_value_grad(sdf, x, ops.HEAD_FULL)
    assert (full2.cpu() - ref).abs().max().item() < TOL_SDF
    assert (grad.cpu() - t(g["grad"])).abs().max().item() < TOL_GRAD
    s3, grad3 = ops.sdf_value_grad(sdf, x, ops.HEAD_SDF_ONLY)
    assert (s3.cpu() - ref[:, 0]).abs().max().item() < TOL_SDF
    assert (grad3.cpu() - t(g["grad"])).abs().max().item() < TOL_GRAD
    rgb = ops.render_forward(rend, x, t(g["view"]).to(dev), t(g["grad"]).to(dev), ref[:, 2:].contiguous().to(dev))
    assert (rgb.cpu() - t(g["rgb"])).abs().max().item() < TOL_RGB


@pytest.mark.parametrize("preset,n", [("w256", 1), ("w256", 63), ("w256", 65), ("w512", 5000), ("w512", 100003)])
def test_mlp_vs_oracle_ragged_sizes(preset, n):
    """Empty / ragged / multi-wave point counts against the oracle (fp32) with an fp64 tie-breaker."""
    from mvsdf_b200 import ops
    dev = torch.device("cuda:0")
    sd, sdf, rend = _nets(preset, dev)
    gen = torch.Generator().manual_seed(n)
    x = (torch.rand(n, 3, generator=gen) * 2 - 1)
    m = min(n, 4096)                      # the CPU oracle checks a bounded sample
    idx = torch.randperm(n, generator=gen)[:m]
    w64 = O.sdf_weights(sd, dtype=torch.float64)
    with torch.no_grad():
        ref64 = O.sdf_mlp(x[idx].double(), w64)
    gref = O.sdf_gradient(x[idx].double(), w64)
    full, grad = ops.sdf_value_grad(sdf, x.to(dev), ops.HEAD_FULL)
    assert (full.cpu()[idx].double() - ref64).abs().max().item() < TOL_SDF
    assert (grad.cpu()[idx].double() - gref).abs().max().item() < TOL_GRAD
    s = ops.sdf_forward(sdf, x.to(dev), ops.HEAD_SDF_ONLY)
    assert (s.cpu()[idx].double() - ref64[:, 0]).abs().max().item() < TOL_SDF
    torch.cuda.synchronize()


def test_mlp_empty_batch():
    from mvsdf_b200 import ops
    dev = torch.device("cuda:0")
    sd, sdf, rend = _nets("w256", dev)
    out = ops.sdf_forward(sdf, torch.zeros(0, 3, device=dev), ops.HEAD_SDF_ONLY)
    assert out.shape == (0,)


def test_mlp_bitwise_deterministic():
    """The tile kernels synchronise through mbarrier / tcgen05.commit chains that compute-sanitizer's racecheck cannot
    model; a data race would show up as run-to-run differences. Same inputs -> bit-identical outputs (pair and
    single-CTA scheduling both exercised: 100003 and 3000 points)."""
    from mvsdf_b200 import ops
    dev = torch.device("cuda:0")
    sd, sdf, rend = _nets("w512", dev)
    for n in (100003, 3000):
        x = (torch.rand(n, 3, generator=torch.Generator().manual_seed(7)) * 2 - 1).to(dev)
        a = ops.sdf_forward(sdf, x, ops.HEAD_SDF_ONLY).clone()
        fa, ga = ops.sdf_value_grad(sdf, x, ops.HEAD_FULL)
        fa, ga = fa.clone(), ga.clone()
        for _ in range(3):
            b = ops.sdf_forward(sdf, x, ops.HEAD_SDF_ONLY)
            fb, gb = ops.sdf_value_grad(sdf, x, ops.HEAD_FULL)
            assert torch.equal(a, b) and torch.equal(fa, fb) and torch.equal(ga, gb)



@pytest.mark.parametrize("preset,n", [("w512", 18944 + 37), ("w256", 40001), ("w512", 64 * 148 * 2)])
def test_fused_head_matches_umma_head_and_is_position_independent(preset, n, monkeypatch):
    """Exact SDF-only evaluations on the CTA-pair kernel compute the one-row head as an fp32 dot product in the last hidden
    layer's epilogue (mlp_pair2_kernel.cuh, MlpArgs::fuse_head); MVSDF_FUSE_HEAD=0 keeps it an UMMA layer.  Both against the
    fp64 oracle, against each other, and the fused path for independence of where a point sits (tile, column, CTA)."""
    from mvsdf_b200 import ops
    from tests.helpers import gate
    dev = torch.device("cuda:0")
    sd, sdf, rend = _nets(preset, dev)
    gen = torch.Generator().manual_seed(n)
    x = (torch.rand(n, 3, generator=gen) * 2 - 1)
    xd = x.to(dev)
    monkeypatch.setenv("MVSDF_FUSE_HEAD", "0")
    umma = ops.sdf_forward(sdf, xd, ops.HEAD_SDF_ONLY).clone()
    monkeypatch.setenv("MVSDF_FUSE_HEAD", "1")
    fused = ops.sdf_forward(sdf, xd, ops.HEAD_SDF_ONLY).clone()
    assert torch.isfinite(fused).all()
    idx = torch.randperm(n, generator=gen)[:4096]
    with torch.no_grad():
        ref64 = O.sdf_mlp(x[idx].double(), O.sdf_weights(sd, dtype=torch.float64))[:, 0]
    gate(f"fused_head_vs_fp64[{preset}]", (fused.cpu()[idx].double() - ref64).abs().max().item(), TOL_SDF)
    gate(f"umma_head_vs_fp64[{preset}]", (umma.cpu()[idx].double() - ref64).abs().max().item(), TOL_SDF)
    gate(f"fused_vs_umma_head[{preset}]", (fused - umma).abs().max().item(), 1.5e-5)
    # reversed and rotated point order: every point lands in another tile / column / CTA -- same bits
    rev = ops.sdf_forward(sdf, xd.flip(0).contiguous(), ops.HEAD_SDF_ONLY).flip(0)
    assert torch.equal(rev, fused), "fused head: a point's value depends on its position"
    rot = ops.sdf_forward(sdf, torch.roll(xd, 77, 0).contiguous(), ops.HEAD_SDF_ONLY)
    assert torch.equal(torch.roll(rot, -77, 0), fused)
    # a device-side count (the tracer's request lists) takes the same path
    n_dev = torch.tensor([n - 5], dtype=torch.int32, device=dev)
    out = torch.full((n,), float("nan"), device=dev)
    from mvsdf_b200 import _lib
    _lib.check(_lib.lib().mvsdf_sdf_forward(sdf.handle, _lib.ptr(sdf.blob), _lib.ptr(xd), n, _lib.ptr(n_dev), ops.HEAD_SDF_ONLY,
                                            _lib.ptr(out), None, ops._stream(dev)))
    torch.cuda.synchronize()
    assert torch.equal(out[:n - 5], fused[:n - 5]) and torch.isnan(out[n - 5:]).all()


def test_dense_grid_matches_oracle_on_a_subsample():
    """Row f3: the 3-D grid plots.py feeds to marching cubes, here 160^3 = 4.1 M points in one SDF-only launch;
    checked on a strided subsample against the fp64 oracle and for the grid's point ordering."""
    from mvsdf_b200.network import B200IDRNetwork, default_conf
    dev = torch.device("cuda:0")
    sd = synth.make_state_dict(width=256, seed=1, perturb=0.05, pe_noise=0.003, bias=0.6)
    model = B200IDRNetwork(default_conf(256)).to(dev)
    model.load_state_dict(sd)
    res = 160
    vol = model.implicit_network.sdf_grid(res)
    assert vol.shape == (res, res, res)
    ax = torch.linspace(-1.0, 1.0, res)
    idx = torch.arange(0, res, 13)
    iy, ix, iz = torch.meshgrid(idx, idx, idx, indexing="ij")
    pts = torch.stack([ax[ix.reshape(-1)], ax[iy.reshape(-1)], ax[iz.reshape(-1)]], dim=1)
    w64 = O.sdf_weights(sd, dtype=torch.float64)
    with torch.no_grad():
        ref = O.sdf_mlp(pts.double(), w64)[:, 0]
    got = vol[iy.reshape(-1), ix.reshape(-1), iz.reshape(-1)].cpu().double()
    assert (got - ref).abs().max().item() < TOL_SDF


@pytest.mark.parametrize("width,n", [(128, 3000), (384, 100), (384, 40000), (128, 40000)])
def test_other_widths_forward_render_and_backward(width, n):
    """Hidden widths 128 and 384 (the architecture family is "multiples of 128 up to 512"): 384 has an odd number of 128-row
    output tiles (a half-empty CTA-pair tile, a 256 + 128 split of the K dimension), 128 a single one.  Small counts take
    the single-CTA kernel, large ones the CTA-pair kernel; the native backward runs on top."""
    from mvsdf_b200 import ops
    from oracle import backward_spec as S
    dev = torch.device("cuda:0")
    sd = synth.make_state_dict(width=width, seed=3, perturb=0.05, pe_noise=0.003, bias=0.6)
    sdf = ops.PackedNet("sdf", width, 8).pack_state_dict(sd, "implicit_network", dev)
    rend = ops.PackedNet("render", width, 4, n_freqs=4).pack_state_dict(sd, "rendering_network", dev)
    gen = torch.Generator().manual_seed(n + width)
    x = torch.rand(n, 3, generator=gen) * 1.6 - 0.8
    m = min(n, 1024)
    idx = torch.randperm(n, generator=gen)[:m]
    w64 = O.sdf_weights(sd, dtype=torch.float64)
    with torch.no_grad():
        ref64 = O.sdf_mlp(x[idx].double(), w64)
    gref = O.sdf_gradient(x[idx].double(), w64)
    full, grad = ops.sdf_value_grad(sdf, x.to(dev), ops.HEAD_FULL)
    assert (full.cpu()[idx].double() - ref64).abs().max().item() < TOL_SDF
    assert (grad.cpu()[idx].double() - gref).abs().max().item() < TOL_GRAD
    s = ops.sdf_forward(sdf, x.to(dev), ops.HEAD_SDF_ONLY)
    assert (s.cpu()[idx].double() - ref64[:, 0]).abs().max().item() < TOL_SDF
    scr = ops.sdf_forward(sdf, x.to(dev), ops.HEAD_SDF_SCREEN)
    assert (scr - s).abs().max().item() < 2e-3
    # rendering net
    view = torch.nn.functional.normalize(torch.randn(n, 3, generator=gen), dim=1)
    rw64 = O.render_weights(sd, dtype=torch.float64)
    with torch.no_grad():
        rgb_ref = O.render_mlp(x[idx].double(), gref, view[idx].double(), ref64[:, 2:], rw64)
    rgb = ops.render_forward(rend, x.to(dev), view.to(dev), grad, full[:, 2:].contiguous())
    assert (rgb.cpu()[idx].double() - rgb_ref).abs().max().item() < TOL_RGB
    # native backward on a sub-batch against the explicit fp64 chain
    nb = min(n, 600)
    xb = x[:nb]
    g_full = torch.randn(nb, 258, generator=gen) * 1e-3
    g_grad = torch.randn(nb, 3, generator=gen) * 1e-2
    vs = [sd[f"implicit_network.lin{l}.weight_v"].double() for l in range(9)]
    gs = [sd[f"implicit_network.lin{l}.weight_g"].double() for l in range(9)]
    bs = [sd[f"implicit_network.lin{l}.bias"].double() for l in range(9)]
    dx_ref, dv_ref, dg_ref, db_ref = S.sdf_value_grad_backward(xb.double(), vs, gs, bs, (4,), 6, g_full.double(), g_grad.double())
    _, _, save = ops.sdf_forward_train(sdf, xb.to(dev))
    dx, dw, db = ops.sdf_backward(sdf, xb.to(dev), save, g_full.to(dev), g_grad.to(dev), need_dx=True)
    assert (dx.cpu().double() - dx_ref).abs().max().item() <= 2e-4 * dx_ref.abs().max().item()
    dvs, dgs, dbs = ops.weight_grads(sdf, dw, db, [v.float().to(dev) for v in vs], [g_.float().to(dev) for g_ in gs])
    for l in range(9):
        for got, ref in ((dvs[l], dv_ref[l]), (dgs[l], dg_ref[l]), (dbs[l], db_ref[l])):
            assert (got.reshape(ref.shape).cpu().double() - ref).abs().max().item() <= 2e-4 * ref.abs().max().item(), (width, l)
