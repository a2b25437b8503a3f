"""Times the SDF-only MLP kernel on a fixed batch (W=512). Env: MVSDF_CLUSTER, MVSDF_DEBUG_FLAGS, SCREEN=1 (screening kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from mvsdf_b200 import ops, synth
from oracle import mvsdf_oracle as O
dev = torch.device("cuda:0")
HEAD = ops.HEAD_SDF_SCREEN if os.environ.get('SCREEN', '0') == '1' else ops.HEAD_SDF_ONLY
kw = dict(width=512, seed=0, perturb=0.05, pe_noise=0.003, bias=0.75)
sd = synth.make_state_dict(**kw)
sdf = ops.PackedNet("sdf", 512, 8).pack_state_dict(sd, "implicit_network", dev)
n = 148 * 64 * 40
g = torch.Generator().manual_seed(3)
xx = (torch.rand(n, 3, generator=g) * 2 - 1).to(dev)
for _ in range(2):
    out = ops.sdf_forward(sdf, xx, HEAD)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    out = ops.sdf_forward(sdf, xx, HEAD)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
macs = 39 * 512 + 6 * 512 * 512 + (512 - 39) * 512 + 512
err = float("nan")
if int(os.environ.get("MVSDF_DEBUG_FLAGS", "0")) == 0:
    idx = torch.arange(0, n, n // 2000)[:2000]
    w64 = O.sdf_weights(sd, dtype=torch.float64)
    with torch.no_grad():
        ref = O.sdf_mlp(xx[idx].cpu().double(), w64)[:, 0]
    err = (out[idx].cpu().double() - ref).abs().max().item()
print(f"screen={os.environ.get('SCREEN','0')} cluster={os.environ.get('MVSDF_CLUSTER','4')} debug={os.environ.get('MVSDF_DEBUG_FLAGS','0')}: {n} pts {ms:.3f} ms "
      f"-> {n/ms/1e3:.1f} Mpts/s, {2*macs*n/ms/1e9:.1f} alg TFLOP/s, max|err| {err:.2e}")
