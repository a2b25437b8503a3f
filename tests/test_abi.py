"""CPU-side checks of the drop-in boundary: the C-ABI library builds (nvcc cross-compiles sm_100a without a
GPU), loads, exports every symbol include/mvsdf_b200.h declares, validates arguments, and fails loudly --
never silently falls back -- when no CUDA device is present.  No compute calls here."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from mvsdf_b200 import _lib
    return _lib.lib()


def test_every_declared_symbol_is_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "mvsdf_b200.h")).read()
    names = sorted(set(re.findall(r"\b(mvsdf_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mvsdf_b200.h but not exported"


def test_plans_and_argument_validation(lib):
    assert lib.mvsdf_abi_version() == 1
    h = lib.mvsdf_sdf_net_create(512, 8, 4, 6, 256)
    assert h
    assert lib.mvsdf_net_num_layers(h) == 9
    # 8 hidden layers + two heads, fp16 hi/lo tiles: a bit over 2 x 2 bytes per padded weight
    assert 8_000_000 < lib.mvsdf_net_packed_bytes(h) < 10_000_000
    lib.mvsdf_net_destroy(h)
    assert not lib.mvsdf_sdf_net_create(100, 8, 4, 6, 256)          # width must be a multiple of 128 (whole output tiles)
    assert not lib.mvsdf_sdf_net_create(320, 8, 4, 6, 256)
    assert b"unsupported" in lib.mvsdf_last_error()
    assert not lib.mvsdf_render_net_create(512, 4, 3, 256)
    r = lib.mvsdf_render_net_create(256, 4, 4, 256)
    assert r and lib.mvsdf_net_num_layers(r) == 5
    lib.mvsdf_net_destroy(r)
    assert lib.mvsdf_trace_workspace_bytes(1024, 1) > 1024 * 60
    assert lib.mvsdf_shade_workspace_bytes(1024, 256) > 1024 * 24


def test_state_dict_names_match_reference_checkpoints():
    from mvsdf_b200 import synth
    from mvsdf_b200.network import B200IDRNetwork, default_conf
    m = B200IDRNetwork(default_conf(256))
    keys = set(m.state_dict().keys())
    assert keys == set(synth.make_state_dict(256).keys())
    assert "implicit_network.lin3.weight_v" in keys and m.implicit_network.lin3.weight_v.shape == (256 - 39, 256)
    assert m.rendering_network.lin0.weight_v.shape == (256, 289)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    from mvsdf_b200 import _lib, ops
    from mvsdf_b200.network import B200IDRNetwork, default_conf
    m = B200IDRNetwork(default_conf(256))
    with pytest.raises((_lib.MvsdfError, AssertionError)):
        m.implicit_network(torch.zeros(4, 3))
    net = ops.PackedNet("sdf", 256, 8)
    rc = lib.mvsdf_sdf_forward(net.handle, ctypes.c_void_p(16), ctypes.c_void_p(16), 4, None, 0, ctypes.c_void_p(16), None, None)
    assert rc < 0 and b"no CUDA device" in lib.mvsdf_last_error()


def test_python_constants_match_the_header():
    """The ctypes side indexes out_counters / the profile arrays with literals; they must be the header's."""
    from mvsdf_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "mvsdf_b200.h")).read()
    macro = lambda name: int(re.search(r"#define\s+" + name + r"\s+(\d+)", hdr).group(1))
    assert _lib.CTR_SCREENED == macro("MVSDF_CTR_SCREENED")
    assert _lib.CTR_SAMPLER_RAYS == macro("MVSDF_CTR_SAMPLER_RAYS")
    assert _lib.CTR_MINSDF_RAYS == macro("MVSDF_CTR_MINSDF_RAYS")
    assert _lib.CTR_REFINED == macro("MVSDF_CTR_REFINED")
    assert _lib.CTR_VIOLATIONS == macro("MVSDF_CTR_VIOLATIONS")
    assert _lib.PROFILE_KINDS == macro("MVSDF_PROFILE_KINDS")
    assert macro("MVSDF_NUM_TRACE_COUNTERS") == 256
    # the tracer parameter block: same field order and size as the C struct (10 x 4 bytes)
    fields = [f for f, _ in _lib.lib().TracerParams._fields_]
    c_fields = re.search(r"typedef struct mvsdf_tracer_params \{(.*?)\} mvsdf_tracer_params;", hdr, re.S).group(1)
    c_names = re.findall(r"^\s*(?:float|int)\s+(\w+);", c_fields, re.M)
    assert fields == c_names, (fields, c_names)
    assert ctypes.sizeof(_lib.lib().TracerParams) == 4 * len(c_names)


def test_argument_validation_of_the_tracer_and_loss_entry_points(lib):
    """Every entry point rejects bad arguments with a negative status and a message BEFORE touching the device
    (so these run without a GPU): null pointers, wrong sample count, short workspaces, wrong channel count."""
    from mvsdf_b200 import ops
    P = ctypes.c_void_p
    net = ops.PackedNet("sdf", 256, 8)
    prm = lib.TracerParams(1.0, 5e-5, 0.5, 0.5, 3, 10, 100, 8, 0, 2e-3)
    d = P(256)            # a non-null dummy "device" address; validation must fail before it is dereferenced
    ws_need = lib.mvsdf_trace_workspace_bytes(1024, 1)
    args = lambda **kw: [net.handle, kw.get("packed", d), d, d, d, None, ctypes.byref(kw.get("prm", prm)), 1, 1024, 0, d, None,
                         ctypes.c_size_t(kw.get("ws", ws_need)), d, d, d, d, d, d, d, None]
    assert lib.mvsdf_trace(*args(packed=None)) == -1 and b"null" in lib.mvsdf_last_error()
    bad = lib.TracerParams(1.0, 5e-5, 0.5, 0.5, 3, 10, 64, 8, 0, 0.0)
    assert lib.mvsdf_trace(*args(prm=bad)) == -1 and b"n_steps" in lib.mvsdf_last_error()
    bad = lib.TracerParams(1.0, 5e-5, 0.5, 0.5, 3, 100, 100, 8, 0, 0.0)
    assert lib.mvsdf_trace(*args(prm=bad)) == -1 and b"iteration" in lib.mvsdf_last_error()
    assert lib.mvsdf_trace(*args(ws=ws_need - 4096)) == -3 and b"workspace" in lib.mvsdf_last_error()
    # training mode needs the caller's CPU-generator samples (ray_tracing.py:287)
    a = args()
    a[9] = 1
    assert lib.mvsdf_trace(*a) == -1 and b"steps01" in lib.mvsdf_last_error()
    # workspace grows with the ray count and covers the prefilter's lists
    assert lib.mvsdf_trace_workspace_bytes(1 << 20, 1) > lib.mvsdf_trace_workspace_bytes(1 << 18, 1) > ws_need
    # feature-consistency forward / backward: 32 channels, >= 2 views, no null pointers
    assert lib.mvsdf_feat_loss_partials(d, d, d, d, 1, 2, 8, 8, 16, d, d, d, None) == -1 and b"32" in lib.mvsdf_last_error()
    assert lib.mvsdf_feat_loss_partials(d, d, d, d, 1, 1, 8, 8, 32, d, d, d, None) == -1
    assert lib.mvsdf_feat_loss_backward(d, d, d, d, 1, 2, 8, 8, 32, d, d, d, None, d, None) == -1
    assert b"null" in lib.mvsdf_last_error()
    assert lib.mvsdf_feat_loss_backward(d, d, d, d, 1, 2, 8, 8, 16, d, d, d, d, d, None) == -1
    assert lib.mvsdf_rgb_l1_partials(d, d, d, 0, d, None) == -1
    assert lib.mvsdf_depth_loss_partials(d, 2, d, 10, d, d, 1, 8, 8, d, d, 0.5, 1.0, 1.0, 1.0, 1.0, d, d, d, None) == -1


def test_argument_validation_of_the_round2_entry_points(lib):
    """Training step, FeatExt, scene-store and counter-budget checks: bad arguments are rejected with a negative status and a
    message before the device is touched (no GPU needed)."""
    from mvsdf_b200 import ops
    P = ctypes.c_void_p
    d = P(256)
    sdf = ops.PackedNet("sdf", 256, 8)
    rend = ops.PackedNet("render", 256, 4, n_freqs=4)
    # sizes reported by the planning calls are consistent
    assert lib.mvsdf_train_packed_t_bytes(sdf.handle) > 2_000_000
    assert lib.mvsdf_train_dw_floats(sdf.handle) >= 256 * 64 + 7 * 256 * 256 + 384 * 256
    assert lib.mvsdf_train_db_floats(sdf.handle) == 8 * 256 + 384
    # 16 points x 4 columns = one 64-column tile; buffers hold an even number of tiles (the CTA-pair forward writes pair tiles)
    s16, s33 = lib.mvsdf_train_save_bytes(sdf.handle, 16, 1), lib.mvsdf_train_save_bytes(sdf.handle, 33, 1)
    assert s33 > s16 > 0 and lib.mvsdf_train_save_bytes(sdf.handle, 64, 0) == s16 == lib.mvsdf_train_save_bytes(sdf.handle, 32, 1)
    assert lib.mvsdf_train_workspace_bytes(rend.handle, 1000, 0) > 0
    # wrong plan kind / null pointers / short buffers
    assert lib.mvsdf_sdf_forward_train(rend.handle, d, d, 10, 1 << 30, d, d, d, None) == -1 and b"SDF net" in lib.mvsdf_last_error()
    assert lib.mvsdf_sdf_forward_train(sdf.handle, d, d, 10, 16, d, d, d, None) == -3 and b"save buffer" in lib.mvsdf_last_error()
    assert lib.mvsdf_sdf_forward_train(sdf.handle, d, d, 10, 1 << 30, None, d, d, None) == -1
    assert lib.mvsdf_sdf_backward(sdf.handle, d, d, 0, d, d, d, 1 << 30, d, None, d, d, None) == -1 and b"empty" in lib.mvsdf_last_error()
    assert lib.mvsdf_sdf_backward(rend.handle, d, d, 10, d, d, d, 1 << 30, d, None, d, d, None) == -1
    assert lib.mvsdf_render_backward(rend.handle, d, 10, d, d, None, None, 1 << 30, d, d, d, d, None, d, d, None) == -1
    assert lib.mvsdf_render_backward(rend.handle, d, 10, d, d, d, None, 1 << 30, d, d, d, d, d, d, d, None) == -1 and b"view_dirs" in lib.mvsdf_last_error()
    assert lib.mvsdf_weight_grads(sdf.handle, None, d, None, None, None, None, None, None) == -1
    assert lib.mvsdf_adam_step(0, None, None, None, None, None, 1e-3, 0.9, 0.999, 1e-8, 1, 0.0, d, None, None) == -1
    assert lib.mvsdf_adam_step(1, None, None, None, None, None, 1e-3, 0.9, 0.999, 1e-8, 0, 0.0, d, None, None) == -1
    # FeatExt: 27 convolutions, image sides multiples of 8, workspace
    assert lib.mvsdf_featext_num_convs() == 27 and lib.mvsdf_featext_packed_floats() > 1_000_000
    need = lib.mvsdf_featext_workspace_bytes(2, 48, 64)
    assert need > 2 * 48 * 64 * 3 * 4
    assert lib.mvsdf_featext_forward(d, d, 2, 50, 64, need, d, None, None, d, None) == -1 and b"multiples of 8" in lib.mvsdf_last_error()
    assert lib.mvsdf_featext_forward(d, d, 2, 48, 64, need - 64, d, None, None, d, None) == -3
    assert lib.mvsdf_featext_forward(d, d, 2, 48, 64, need, d, None, None, None, None) == -1
    # scene-store variant of the feature loss validates like the tensor one
    assert lib.mvsdf_feat_loss_partials_indexed(d, d, d, d, d, 1, 2, 8, 8, 16, d, d, d, None) == -1 and b"32" in lib.mvsdf_last_error()
    # the tracer refuses phase counts that would run out of request counters, before launching anything
    net = ops.PackedNet("sdf", 256, 8)
    prm = lib.TracerParams(1.0, 5e-5, 0.5, 0.5, 8, 64, 100, 8, 0, 0.0)          # 1 + 64 * 9 phases > 251 counters
    ws_need = lib.mvsdf_trace_workspace_bytes(1024, 1)
    rc = lib.mvsdf_trace(net.handle, d, d, d, d, None, ctypes.byref(prm), 1, 1024, 0, d, None, ctypes.c_size_t(ws_need), d, d, d, d, d, d, d, None)
    assert rc == -1 and b"request counters" in lib.mvsdf_last_error()
    # widths outside the whole-tile family are refused for both nets
    assert not lib.mvsdf_render_net_create(320, 4, 4, 256) and not lib.mvsdf_render_net_create(64, 4, 4, 256)
