"""Diagnostic for the native backward (csrc/mlp_bwd_kernel.cuh): decodes the saved layer inputs and the dumped [dZ | dS]
images and compares them, layer by layer, with the explicit chain of oracle/backward_spec.py (fp64).  Prints one line per
layer so that a wrong layout / scale / epilogue shows up where it happens.  Test infrastructure (imports oracle/)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from mvsdf_b200 import _lib, ops, synth
from oracle import backward_spec as S


def decode(img: torch.Tensor, kc: int, n_tiles: int, scale: float) -> torch.Tensor:
    """K-sliced image (uint8) -> [n_tiles * 64 columns, kc * 8 features] float64."""
    h = img[: n_tiles * kc * 2048].view(torch.float16).view(n_tiles, 4, kc, 2, 2, 8, 8)      # tile, slice, fb, hi/lo, cb, row, col
    v = (h[:, :, :, 0].double() + h[:, :, :, 1].double()) / scale                               # tile, slice, fb, cb, row, col
    v = v.permute(0, 1, 3, 5, 2, 4)                                                             # tile, slice, cb, col, fb, row
    return v.reshape(n_tiles * 64, kc * 8)


def main(width=256, n=300, seed=0):
    dev = torch.device("cuda:0")
    L = _lib.lib()
    sd = synth.make_state_dict(width=width, seed=1, perturb=0.05, pe_noise=0.003, bias=0.6)
    net = ops.PackedNet("sdf", width, 8).pack_state_dict(sd, "implicit_network", dev)
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(n, 3, generator=g) * 1.6 - 0.8)
    F = 256
    g_full = torch.randn(n, F + 2, generator=g) * 1e-3
    g_grad = torch.randn(n, 3, generator=g) * 1e-2
    vs = [sd[f"implicit_network.lin{l}.weight_v"].double() for l in range(9)]
    gs = [sd[f"implicit_network.lin{l}.weight_g"].double() for l in range(9)]
    bs = [sd[f"implicit_network.lin{l}.bias"].double() for l in range(9)]
    tr = {}
    dx_ref, dv_ref, dg_ref, db_ref = S.sdf_value_grad_backward(x.double(), vs, gs, bs, (4,), 6, g_full.double(), g_grad.double(), trace=tr)

    full, grad, save = ops.sdf_forward_train(net, x.to(dev))
    full0, grad0 = ops.sdf_value_grad(net, x.to(dev), ops.HEAD_FULL)
    print("forward_train vs value_grad: full", float((full - full0).abs().max()), "grad", float((grad - grad0).abs().max()))
    n_tiles = (n + 15) // 16
    n_alloc = (n_tiles + 1) // 2 * 2          # buffers hold an even number of tiles (train_abi.cu alloc_tiles)
    # saved layer inputs
    kcs = [8] + [width // 8] * 8
    off = 0
    for l in range(9):
        img = save[off: off + n_tiles * kcs[l] * 2048]
        off += n_alloc * kcs[l] * 2048
        dec = decode(img, kcs[l], n_tiles, 64.0).cpu()[: n * 4]                        # [n*4, feat]
        Hl, Tl = tr["H"][l], tr["T"][l]                                                  # [n, feat], [n, feat, 3]
        ref = torch.cat([Hl.unsqueeze(1), Tl.permute(0, 2, 1)], dim=1).reshape(n * 4, -1)   # column = pt*4 + j
        if l == 4:
            ref = ref * 2 ** 0.5          # the kernels keep cat([h, PE]) unscaled; 1/sqrt 2 is folded into the packed weights
        d = (dec[:, : ref.shape[1]] - ref).abs().max().item()
        print(f"saved H_{l}: kc {kcs[l]}, max |err| {d:.3e} (ref max {ref.abs().max().item():.3e}), pad max {dec[:, ref.shape[1]:].abs().max().item() if dec.shape[1] > ref.shape[1] else 0:.3e}")

    dx, dw, db = ops.sdf_backward(net, x.to(dev), save, g_full.to(dev), g_grad.to(dev), need_dx=True)
    torch.cuda.synchronize()
    ws = ops._POOL["sdf_bwd"]
    gscale = ws[:8].view(torch.float32).cpu()
    Sg = float(gscale[0])
    print("gscale", Sg, float(gscale[1]), "max |g|", float(max(g_full.abs().max(), g_grad.abs().max())))
    m_tiles = [width // 128] * 8 + [3]
    off = 256
    order = None
    dw_off = 0
    db_off = 0
    for l in range(9):
        kc = m_tiles[l] * 16
        img = ws[off: off + n_tiles * kc * 2048]
        off += n_alloc * kc * 2048
        dec = decode(img, kc, n_tiles, Sg).cpu()[: n * 4]
        dz, ds = tr["DZ"][l], tr["DS"][l]                                               # [n, out], [n, out, 3]
        ref = torch.cat([dz.unsqueeze(1), ds.permute(0, 2, 1)], dim=1).reshape(n * 4, -1)
        if l == 8:                                                                      # head rows: features first, then sdf, indicator
            ref = torch.cat([ref[:, 2:], ref[:, :2]], dim=1)
        d = (dec[:, : ref.shape[1]] - ref).abs().max().item()
        print(f"dumped dZ_{l}: max |err| {d:.3e} (ref max {ref.abs().max().item():.3e})")
        # dW in plan coordinates
        in_pad = kcs[l] * 8
        rows = m_tiles[l] * 128
        dwl = dw[dw_off: dw_off + rows * in_pad].view(rows, in_pad).cpu().double()
        dw_off += rows * in_pad
        ref_w = tr["DW"][l]
        if l == 8:
            ref_w = torch.cat([ref_w[2:], ref_w[:2]], dim=0)
        cs = (1.0 / 2 ** 0.5) if l == 4 else 1.0
        d = (dwl[: ref_w.shape[0], : ref_w.shape[1]] * cs - ref_w).abs().max().item()
        print(f"   dW_{l}: max |err| {d:.3e} (ref max {ref_w.abs().max().item():.3e})")
        dbl = db[db_off: db_off + rows].cpu().double()
        db_off += rows
        ref_b = tr["DB"][l]
        if l == 8:
            ref_b = torch.cat([ref_b[2:], ref_b[:2]])
        print(f"   db_{l}: max |err| {(dbl[: ref_b.shape[0]] - ref_b).abs().max().item():.3e} (ref max {ref_b.abs().max().item():.3e})")
    print("dx: max |err|", float((dx.cpu().double() - dx_ref).abs().max()), "ref max", float(dx_ref.abs().max()))
    vs32 = [sd[f"implicit_network.lin{l}.weight_v"].to(dev) for l in range(9)]
    gs32 = [sd[f"implicit_network.lin{l}.weight_g"].to(dev) for l in range(9)]
    dvs, dgs, dbs = ops.weight_grads(net, dw, db, vs32, gs32)
    for l in range(9):
        print(f"layer {l}: dv err {(dvs[l].cpu().double() - dv_ref[l]).abs().max().item():.3e} / {dv_ref[l].abs().max().item():.3e}   "
              f"dg err {(dgs[l].cpu().double() - dg_ref[l]).abs().max().item():.3e} / {dg_ref[l].abs().max().item():.3e}   "
              f"db err {(dbs[l].cpu().double() - db_ref[l]).abs().max().item():.3e} / {db_ref[l].abs().max().item():.3e}")


if __name__ == "__main__":
    main(width=int(sys.argv[1]) if len(sys.argv) > 1 else 256, n=int(sys.argv[2]) if len(sys.argv) > 2 else 300)
