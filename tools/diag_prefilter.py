"""Screening-precision SDF kernel and tracer prefilter: error statistics, speed, bit-exactness vs tau (diagnostic)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvsdf_b200 import ops, synth
from mvsdf_b200.network import B200IDRNetwork, default_conf
from tests.helpers import WEIGHT_PRESETS

dev = torch.device("cuda:0")


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for preset in ["w256", "w256_geo", "w512"]:
    kw = WEIGHT_PRESETS[preset]
    sd = synth.make_state_dict(**kw)
    sdf = ops.PackedNet("sdf", kw["width"], 8).pack_state_dict(sd, "implicit_network", dev)
    g = torch.Generator().manual_seed(7)
    n = 148 * 64 * 40
    x = (torch.rand(n, 3, generator=g) * 2 - 1).to(dev)
    exact = ops.sdf_forward(sdf, x, ops.HEAD_SDF_ONLY)
    lp = ops.sdf_forward(sdf, x, ops.HEAD_SDF_SCREEN)
    err = (lp - exact).abs()
    near = exact.abs() < 0.05
    q = torch.quantile(err[:200000], torch.tensor([0.5, 0.99, 0.999], device=dev)).tolist()
    print(f"{preset}: screening error max {err.max().item():.3e} (near surface {err[near].max().item():.3e}), "
          f"median/99/99.9% {q[0]:.2e}/{q[1]:.2e}/{q[2]:.2e}; nan {int(torch.isnan(lp).sum())}")
    t_e = timed(lambda: ops.sdf_forward(sdf, x, ops.HEAD_SDF_ONLY))
    t_l = timed(lambda: ops.sdf_forward(sdf, x, ops.HEAD_SDF_SCREEN))
    print(f"{preset}: {n} pts exact {t_e:.3f} ms, screening {t_l:.3f} ms ({t_e / t_l:.2f}x)")

for (preset, H, W, training) in [("w256", 96, 96, False), ("w256", 96, 96, True), ("w512", 400, 400, False), ("w512", 300, 300, True)]:
    kw = WEIGHT_PRESETS[preset]
    sd = synth.make_state_dict(**kw)
    scene = synth.make_scene(H, W, n_images=1, n_src=1, seed=0)
    model = B200IDRNetwork(default_conf(kw["width"])).to(dev)
    model.load_state_dict(sd)
    model.train(training)
    sdf_net = model.implicit_network.packed()
    uv, pose, K = scene["uv"].to(dev), scene["pose"].to(dev), scene["intrinsics"].to(dev)
    obj = torch.ones(uv.shape[0] * uv.shape[1], dtype=torch.uint8, device=dev)
    steps = torch.rand(100, generator=torch.Generator().manual_seed(3))
    base = None
    for tau in [0.0, 1e-3, 2e-3, 4e-3, 8e-3, 1.6e-2]:
        model.prefilter_tau = tau
        run = lambda: model.trace(sdf_net, uv, pose, K, obj, training, steps)
        ms = timed(run, 3)
        dirs, cam, dists, nm, pts = run()
        cnt = model.last_trace_counters.cpu()
        R = dists.numel()
        if base is None:
            base = (dists.clone(), nm.clone(), pts.clone())
            same = "-"
        else:
            eq = lambda a, b: bool(((a == b) | (torch.isnan(a) & torch.isnan(b))).all()) if a.is_floating_point() else bool((a == b).all())
            same = f"dists {eq(dists, base[0])} mask {eq(nm, base[1])} points {eq(pts, base[2])} (max |d dists| {(dists - base[0]).abs().nan_to_num().max().item():.2e})"
        print(f"{preset} {H}x{W} train={training} tau={tau:g}: {ms:.2f} ms, evals/ray {int(cnt[:251].sum()) / R:.1f}, screened/ray {int(cnt[251]) / R:.1f}, "
              f"sampler rays {int(cnt[252]) / R:.3f}, minsdf rays {int(cnt[253]) / R:.3f}, refined/ray {int(cnt[254]) / R:.2f}, "
              f"violations {int(cnt[255])}; bit-identical: {same}")
