"""Diagnostic: what do the MLP kernels return when an activation leaves the fp16 hi/lo range?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvsdf_b200 import ops, synth

dev = torch.device("cuda:0")
sd = synth.make_state_dict(width=256, seed=1, perturb=0.05, pe_noise=0.003, bias=0.6)
for n in (100, 40000):
    for add in (0.0, 500.0, 2000.0, 1e6):
        bad = {k: v.clone() for k, v in sd.items()}
        bad["rendering_network.lin0.bias"] += add
        net = ops.PackedNet("render", 256, 4, n_freqs=4).pack_state_dict(bad, "rendering_network", dev)
        g = torch.Generator().manual_seed(0)
        p, v, nrm = (torch.randn(n, 3, generator=g).to(dev) for _ in range(3))
        f = torch.randn(n, 256, generator=g).to(dev)
        rgb = ops.render_forward(net, p, v, nrm, f)
        torch.cuda.synchronize()
        print(n, add, "rgb finite:", bool(torch.isfinite(rgb).all()), "min/max", float(rgb.min()), float(rgb.max()),
              "status", net.status().cpu().tolist())
