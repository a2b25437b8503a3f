"""Prints error statistics of the CUDA MLP kernels vs the fp64 oracle (diagnostic, not a test)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from mvsdf_b200 import ops, synth
from oracle import mvsdf_oracle as O
from tests.helpers import WEIGHT_PRESETS

dev = torch.device("cuda:0")
for preset in ["w256", "w512"]:
    kw = WEIGHT_PRESETS[preset]
    sd = synth.make_state_dict(**kw)
    sdf = ops.PackedNet("sdf", kw["width"], 8).pack_state_dict(sd, "implicit_network", dev)
    rend = ops.PackedNet("render", kw["width"], 4, n_freqs=4).pack_state_dict(sd, "rendering_network", dev)
    torch.cuda.synchronize()
    g = torch.Generator().manual_seed(5)
    x = torch.rand(1000, 3, generator=g) * 2 - 1
    w64 = O.sdf_weights(sd, dtype=torch.float64)
    with torch.no_grad():
        ref = O.sdf_mlp(x.double(), w64)
    gref = O.sdf_gradient(x.double(), w64)
    s = ops.sdf_forward(sdf, x.to(dev), ops.HEAD_SDF_ONLY).cpu().double()
    print(preset, "sdf-only head: max abs err", (s - ref[:, 0]).abs().max().item(), " first:", s[:4].tolist(), ref[:4, 0].tolist())
    full = ops.sdf_forward(sdf, x.to(dev), ops.HEAD_FULL).cpu().double()
    print(preset, "full head: sdf err", (full[:, 0] - ref[:, 0]).abs().max().item(), "indicator err", (full[:, 1] - ref[:, 1]).abs().max().item(),
          "feat err", (full[:, 2:] - ref[:, 2:]).abs().max().item())
    full2, grad = ops.sdf_value_grad(sdf, x.to(dev), ops.HEAD_FULL)
    print(preset, "value+grad: value err", (full2.cpu().double() - ref).abs().max().item(), "grad err", (grad.cpu().double() - gref).abs().max().item(),
          " grad first:", grad[0].tolist(), gref[0].tolist())
    view = torch.nn.functional.normalize(torch.randn(1000, 3, generator=g), dim=1)
    rw = O.render_weights(sd, dtype=torch.float64)
    with torch.no_grad():
        rgb_ref = O.render_mlp(x.double(), gref, view.double(), ref[:, 2:], rw)
    rgb = ops.render_forward(rend, x.to(dev), view.to(dev), gref.float().to(dev), ref[:, 2:].float().contiguous().to(dev))
    print(preset, "render: rgb err", (rgb.cpu().double() - rgb_ref).abs().max().item())
    # throughput of the SDF-only head
    n = 148 * 64 * 40
    xx = (torch.rand(n, 3, device=dev) * 2 - 1)
    for _ in range(2):
        ops.sdf_forward(sdf, xx, ops.HEAD_SDF_ONLY)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.sdf_forward(sdf, xx, ops.HEAD_SDF_ONLY)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    W = kw["width"]
    macs = 39 * W + 6 * W * W + (W - 39) * W + W   # sdf-only head
    print(preset, f"sdf-only: {n} pts in {ms:.3f} ms -> {n / ms / 1e3:.2f} Mpts/s, {2 * macs * n / ms / 1e9:.1f} algorithmic TFLOP/s")
