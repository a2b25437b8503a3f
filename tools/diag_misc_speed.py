"""Timings of the two one-off-per-scene rows: the dense SDF grid for mesh extraction (utils/plots.py:113-163) at
512^3 and FeatExt over the 49 images of a DTU scene at 1200x1600 (datasets/scene_dataset.py:138-149, row f4)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvsdf_b200 import synth
from mvsdf_b200.featext import B200FeatExt
from mvsdf_b200.loss import FeatureStore
from mvsdf_b200.network import B200IDRNetwork, default_conf

dev = torch.device("cuda:0")
model = B200IDRNetwork(default_conf(512)).to(dev)
model.load_state_dict(synth.make_state_dict(width=512, seed=0, perturb=0.05, pe_noise=0.003, bias=0.75))
model.eval()
for res in (256, 512):
    model.implicit_network.sdf_grid(64)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    vol = model.implicit_network.sdf_grid(res)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"sdf_grid {res}^3 = {res**3 / 1e6:.1f} M points: {dt * 1e3:.1f} ms = {res**3 / dt / 1e6:.1f} M points/s "
          f"({res**3 * 3.671e6 / dt / 1e12:.0f} algorithmic TFLOP/s), inside fraction {(vol < 0).float().mean().item():.3f}")
    del vol
fe = B200FeatExt(seed=1).to(dev)
store = FeatureStore()
imgs = torch.randn(49, 3, 1200, 1600)
fe.forward_nhwc(imgs[:2].to(dev))
torch.cuda.synchronize()
t0 = time.perf_counter()
maps = fe.fill_store(store, imgs, batch=7, device=dev)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
x = imgs[:7].to(dev)
e0.record()
fe.forward_nhwc(x)
e1.record()
torch.cuda.synchronize()
print(f"FeatExt 49 x 1200x1600 -> store {tuple(maps.shape)} ({maps.numel() * 4 / 1e9:.2f} GB): {dt:.2f} s incl. the host->device copies; "
      f"device time {e0.elapsed_time(e1) / 7:.1f} ms per image")
