"""Timeline of CTA pair 0 of the pair2 MLP kernel (clock64 stamps written by the kernel when a trace buffer is set).
Prints, per layer of a steady-state tile, when the issuer / epilogue warp 0 reached each point (cycles relative to the
tile's first event).  Debug tool: needs a library built with -DMVSDF_TRACE, e.g.
  nvcc -DMVSDF_TRACE <flags of __graft_entry__.NVCC_FLAGS> -o tools/_exp/lib_trace.so mvsdf_b200/csrc/{mlp_abi,tracer,render}.cu
  MVSDF_LIB_PATH=$PWD/tools/_exp/lib_trace.so python tools/diag_trace.py
The production build compiles the stamps out."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvsdf_b200 import ops, synth, _lib
dev = torch.device("cuda:0")
sd = synth.make_state_dict(width=512, seed=0, perturb=0.05, pe_noise=0.003, bias=0.75)
sdf = ops.PackedNet("sdf", 512, 8).pack_state_dict(sd, "implicit_network", dev)
n = 148 * 64 * 40
xx = (torch.rand(n, 3, generator=torch.Generator().manual_seed(3)) * 2 - 1).to(dev)
L = _lib.lib()
HEAD = ops.HEAD_SDF_SCREEN if os.environ.get('SCREEN', '0') == '1' else ops.HEAD_SDF_ONLY     # SCREEN=1: screening kernel
out = ops.sdf_forward(sdf, xx, HEAD)
torch.cuda.synchronize()
buf = torch.zeros(4 * 16384, dtype=torch.int64, device=dev)
L.mvsdf_debug_set_trace.argtypes = [ctypes.c_void_p]
L.mvsdf_debug_set_trace.restype = None
L.mvsdf_debug_set_trace(ctypes.c_void_p(buf.data_ptr()))
out = ops.sdf_forward(sdf, xx, HEAD)
torch.cuda.synchronize()
L.mvsdf_debug_set_trace(ctypes.c_void_p(0))
t = buf.cpu().view(4, 16384)
n_run = 9
tile = int(os.environ.get("TILE", "5"))
iss = t[0, tile * n_run * 8:(tile + 1) * n_run * 8].view(n_run, 8)
ep0 = t[1, tile * (n_run + 1) * 8:(tile + 1) * (n_run + 1) * 8].view(n_run + 1, 8)
ep1 = t[2, tile * (n_run + 1) * 8:(tile + 1) * (n_run + 1) * 8].view(n_run + 1, 8)
base = int(ep0[0, 0])
def rel(v, b=base):
    return [int(x) - b if int(x) else -1 for x in v]
print("tile", tile, "cycles relative to CTA0 prologue start; next tile prologue starts at",
      int(t[1, (tile + 1) * (n_run + 1) * 8]) - base)
print("CTA0 epilogue warp0: prologue start/done", rel(ep0[0, :2]))
print("layer | issuer: start x0ok p2reach p1reach x1ok p3reach d1ok issued | epi0: waitD0 D0rdy drained E0done waitD1 D1rdy drained E1done")
for l in range(n_run):
    print(l, rel(iss[l]), rel(ep0[l + 1]))
b1 = int(ep1[0, 0])
print("CTA1 epilogue warp0 (own clock, relative to its prologue start):")
for l in range(n_run):
    print(l, rel(ep1[l + 1], b1))

