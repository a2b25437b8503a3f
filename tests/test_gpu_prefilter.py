"""The tracer's prefilter (include/mvsdf_b200.h: mvsdf_tracer_params.prefilter_tau): the 100-sample stages of
ray_sampler (ray_tracing.py:198-258) and minimal_sdf_points (:280-308) evaluate every sample with the screening kernel
(single fp16 product) and only the samples the selection logic can depend on with the exact kernel.  The claim tested
here is bit-identity with the prefilter off, not a tolerance."""
import pytest
import torch

from mvsdf_b200 import synth
from tests.helpers import WEIGHT_PRESETS, preset_state_dict

pytestmark = pytest.mark.gpu

SCREEN_TOL = 2e-3       # absolute error bound of the screening head used for the default tau (measured max: 9.4e-4)


def _model(preset, dev):
    from mvsdf_b200.network import B200IDRNetwork, default_conf
    m = B200IDRNetwork(default_conf(WEIGHT_PRESETS[preset]["width"])).to(dev)
    m.load_state_dict(preset_state_dict(preset))
    return m


@pytest.mark.parametrize("preset,n", [("w256", 18945), ("w256_geo", 50000), ("w512", 100003), ("w512", 111 * 256)])
def test_screening_head_error_bound(preset, n):
    """Ragged point counts through the 256-column screening tiles; every point is written exactly once."""
    from mvsdf_b200 import ops
    dev = torch.device("cuda:0")
    sd = preset_state_dict(preset)
    sdf = ops.PackedNet("sdf", WEIGHT_PRESETS[preset]["width"], 8).pack_state_dict(sd, "implicit_network", dev)
    x = (torch.rand(n, 3, generator=torch.Generator().manual_seed(n)) * 2 - 1).to(dev)
    exact = ops.sdf_forward(sdf, x, ops.HEAD_SDF_ONLY)
    lp = torch.full((n + 300,), float("nan"), device=dev)
    out = lp[:n]
    from mvsdf_b200 import _lib
    _lib.check(_lib.lib().mvsdf_sdf_forward(sdf.handle, _lib.ptr(sdf.blob), _lib.ptr(x), n, None, ops.HEAD_SDF_SCREEN,
                                            _lib.ptr(out), None, ops._stream(dev)))
    torch.cuda.synchronize()
    assert not torch.isnan(out).any()
    assert torch.isnan(lp[n:]).all(), "screening kernel wrote past the last point"
    err = (out - exact).abs().max().item()
    assert 0.0 < err < SCREEN_TOL, err
    again = ops.sdf_forward(sdf, x, ops.HEAD_SDF_SCREEN)
    assert torch.equal(again, out)


@pytest.mark.parametrize("preset,hw,training", [("w256", 64, False), ("w256", 64, True), ("w512", 96, False),
                                                 ("w512", 80, True), ("w256_geo", 64, True)])
def test_tracer_prefilter_is_bit_identical(preset, hw, training):
    dev = torch.device("cuda:0")
    model = _model(preset, dev)
    model.train(training)
    scene = synth.make_scene(hw, hw, n_images=2, n_src=1, seed=5, mask_mode="disk" if training else "ones")
    uv, pose, K = scene["uv"].to(dev), scene["pose"].to(dev), scene["intrinsics"].to(dev)
    obj = scene["object_mask"].reshape(-1).to(dev).to(torch.uint8).contiguous()
    steps = torch.rand(100, generator=torch.Generator().manual_seed(3))
    sdf_net = model.implicit_network.packed()
    res = {}
    for tau in (0.0, 2e-3):
        model.prefilter_tau = tau
        dirs, cam, dists, nm, pts = model.trace(sdf_net, uv, pose, K, obj, training, steps)
        res[tau] = (dists.clone(), nm.clone(), pts.clone(), model.last_trace_counters.cpu().clone())
    a, b = res[0.0], res[2e-3]
    R = a[0].numel()
    assert torch.equal(a[1], b[1])
    for i in (0, 2):
        assert bool(((a[i] == b[i]) | (torch.isnan(a[i]) & torch.isnan(b[i]))).all())
    assert torch.equal(a[3][:251], b[3][:251]) and torch.equal(a[3][252:254], b[3][252:254]), "E_trace accounting must not depend on the prefilter"
    assert int(a[3][251]) == 0 and 0 < int(b[3][251]) <= 100 * (int(b[3][252]) + int(b[3][253]))
    assert int(a[3][254]) == 0 and int(b[3][255]) == 0
    n_sampled = int(b[3][252]) + int(b[3][253])
    assert n_sampled > 0 and 0 < int(b[3][254]) < 40 * n_sampled, (int(b[3][254]), n_sampled, R)


def test_forward_guard_falls_back_to_exact():
    """A tau below the screening error trips the guard counter; forward() must then repeat the step exactly."""
    dev = torch.device("cuda:0")
    model = _model("w256", dev)
    model.eval()
    scene = synth.make_scene(48, 48, n_images=1, n_src=1, seed=2)
    inp = {k: scene[k].to(dev) for k in ["uv", "pose", "intrinsics", "object_mask"]}
    model.prefilter_tau = 0.0
    ref = model(inp)
    model.prefilter_tau = 1e-5
    out = model(inp)
    assert model.prefilter_fallbacks == 1 and model.prefilter_tau == 2e-5      # widened for the next call
    for k in ("points", "rgb_values", "sdf_output", "network_object_mask"):
        assert torch.equal(ref[k], out[k]), k
    # the widening converges: after a few steps the guard is quiet and the outputs are still the exact ones
    for _ in range(12):
        out = model(inp)
    assert 1e-3 <= model.prefilter_tau <= 2.5e-2 and int(model.last_trace_counters[255]) == 0
    for k in ("points", "rgb_values", "sdf_output", "network_object_mask"):
        assert torch.equal(ref[k], out[k]), k
