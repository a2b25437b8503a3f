"""Diagnostic: eager launch sequence vs CUDA-graph replay of the no-grad forward at training sizes (cfg3-shaped)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvsdf_b200 import synth
from mvsdf_b200.network import B200IDRNetwork, default_conf

dev = torch.device("cuda:0")
for (H, W, width, n_images, n_rays, training) in [(1200, 1600, 512, 2, 4096, True), (32, 32, 256, 1, None, False)]:
    scene = synth.make_scene(H, W, n_images=n_images, n_src=1, n_rays=n_rays, seed=0)
    sd = synth.make_state_dict(width=width, seed=0, perturb=0.05, pe_noise=0.003, bias=0.75 if width == 512 else 0.6)
    model = B200IDRNetwork(default_conf(width)).to(dev)
    model.load_state_dict(sd)
    model.train(training)
    inp = {k: scene[k].to(dev) for k in ["uv", "pose", "intrinsics", "object_mask"]}
    R = scene["uv"].shape[0] * scene["uv"].shape[1]
    g = torch.Generator().manual_seed(1)
    steps01, eik = torch.rand(100, generator=g), torch.rand(R // 2, 3, generator=g) * 2 - 1
    kw = dict(steps01=steps01, eik_points=eik) if training else {}
    for use in (False, True):
        model.use_graphs = use
        with torch.no_grad():
            for _ in range(3):
                model(inp, 0.5 if training else None, **kw)
            torch.cuda.synchronize()
            ts = []
            for _ in range(12):
                t0 = time.perf_counter()
                out = model(inp, 0.5 if training else None, **kw)
                torch.cuda.synchronize()
                ts.append((time.perf_counter() - t0) * 1e3)
        print(f"R={R} width={width} training={training} graphs={use}: per-step ms {['%.2f' % t for t in ts]}  "
              f"graphs cached {len(model._graphs)} replays {model.graph_replays} fallbacks {model.prefilter_fallbacks}")
