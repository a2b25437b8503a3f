"""Times the native backward kernels of the SDF net on a fixed batch (W=512): saving forward, reverse sweep, dW GEMM.
Used for the ncu captures of profiles/r02 (mlp_bwd_sweep_kernel / mlp_bwd_dw_kernel)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvsdf_b200 import _lib, ops, synth

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 16 * 16          # 16 tiles of 16 points per SM
sd = synth.make_state_dict(width=512, seed=0, perturb=0.05, pe_noise=0.003, bias=0.75)
net = ops.PackedNet("sdf", 512, 8).pack_state_dict(sd, "implicit_network", dev)
g = torch.Generator().manual_seed(3)
x = (torch.rand(n, 3, generator=g) * 1.6 - 0.8).to(dev)
g_full = (torch.randn(n, 258, generator=g) * 1e-3).to(dev)
g_grad = (torch.randn(n, 3, generator=g) * 1e-2).to(dev)
L = _lib.lib()
full, grad, save = ops.sdf_forward_train(net, x)
for _ in range(2):
    ops.sdf_backward(net, x, save, g_full, g_grad, need_dx=True)
torch.cuda.synchronize()
L.mvsdf_profile_enable(1)
reps = 5
for _ in range(reps):
    full, grad, save = ops.sdf_forward_train(net, x)
    ops.sdf_backward(net, x, save, g_full, g_grad, need_dx=True)
torch.cuda.synchronize()
ms = (ctypes.c_float * 8)()
cnt = (ctypes.c_int * 8)()
_lib.check(L.mvsdf_profile_collect(ms, cnt))
L.mvsdf_profile_enable(0)
full_flop = 3.934e6
fwd, sweep, dw = ms[2] / reps, ms[5] / reps, ms[6] / reps
print(f"{n} points x 4 columns, W=512: saving forward {fwd:.3f} ms ({n * 4 * full_flop / fwd / 1e9:.0f} alg TFLOP/s), "
      f"reverse sweep {sweep:.3f} ms ({n * 4 * full_flop / sweep / 1e9:.0f}), dW GEMM {dw:.3f} ms ({n * 4 * full_flop / dw / 1e9:.0f}); "
      f"saved activations {save.numel() / 1e6:.0f} MB")
