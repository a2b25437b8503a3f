"""Explicit reverse sweep through ImplicitNetwork.forward + .gradient -- the math of SURVEY.md section 8 row f1.

TEST INFRASTRUCTURE (like everything under oracle/): a checker, never the product.

The reference obtains parameter gradients of the eikonal / normal terms by differentiating autograd's own backward
(``create_graph=True``, code/model/implicit_differentiable_renderer.py:96-107).  A fused kernel cannot do that; it has to
run the second-order chain explicitly.  This module writes that chain out, layer by layer, in the form the tile core can
execute (GEMMs with W, W^T and outer products over the points), and tests/test_backward_spec.py pins it against autograd
through the oracle restatement of the reference -- so that the round-2 kernel has a CPU statement to be compared with.

Forward, per point, with 1 + 3 columns (value, d/dx, d/dy, d/dz) -- exactly what mlp_pair2_kernel<NET_SDF, value+grad> does:
    H_0 = PE(x)                          T_0[:, j] = dPE/dx_j
    Z_l = W_l H_l + b_l                  S_l[:, j] = W_l T_l[:, j]
    H_{l+1} = sp(Z_l)                    T_{l+1}[:, j] = sp'(Z_l) * S_l[:, j]          (sp = softplus, beta = 100)
    at the skip layer:  H <- cat(H, PE) / sqrt 2,  T <- cat(T, dPE) / sqrt 2          (:86-87)
    outputs:  full = Z_L,   grad = S_L[0, :]
Reverse, given G_full = dLoss/dfull and G_grad = dLoss/dgrad:
    dZ_L = G_full,  dS_L[0, j] = G_grad[j]
    dW_l = dZ_l H_l^T + sum_j dS_l[:, j] T_l[:, j]^T        db_l = dZ_l
    dH_l = W_l^T dZ_l                                        dT_l[:, j] = W_l^T dS_l[:, j]
    dS_{l-1}[:, j] = sp'(Z_{l-1}) * dT_l[:, j]
    dZ_{l-1} = sp'(Z_{l-1}) * dH_l + sp''(Z_{l-1}) * sum_j S_{l-1}[:, j] * dT_l[:, j]
    dx = J_PE^T dH_0 + sum_j (dJ_PE/dx_j)^T dT_0[:, j]       (only the diagonal of PE's second derivative is non-zero)
Weight norm (W = g v / |v|_row, :70-71) is folded at the end:  dg = <dW, v> / |v|,  dv = g/|v| (dW - <dW, v^> v^).
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import torch

BETA = 100.0
THRESHOLD = 20.0          # nn.Softplus(beta=100) switches to the identity above beta z = 20


def _sp(z):
    return torch.nn.functional.softplus(z, beta=BETA, threshold=THRESHOLD)


def _sp1(z):
    s = torch.sigmoid(BETA * z)
    return torch.where(BETA * z > THRESHOLD, torch.ones_like(z), s)


def _sp2(z):
    s = torch.sigmoid(BETA * z)
    return torch.where(BETA * z > THRESHOLD, torch.zeros_like(z), BETA * s * (1 - s))


def _pe_all(x: torch.Tensor, n_freqs: int):
    """PE [P, D], its Jacobian columns dPE/dx_j [P, D, 3] and the second derivatives d2PE/dx_j^2 [P, D, 3]
    (model/embedder.py:5-50: [x, sin(2^k x), cos(2^k x)]_k; every feature depends on one coordinate only)."""
    P = x.shape[0]
    feats, d1, d2 = [x], [torch.eye(3, dtype=x.dtype).expand(P, 3, 3)], [torch.zeros(P, 3, 3, dtype=x.dtype)]
    for k in range(n_freqs):
        f = float(2 ** k)
        s, c = torch.sin(x * f), torch.cos(x * f)
        feats += [s, c]
        d1 += [torch.diag_embed(f * c), torch.diag_embed(-f * s)]
        d2 += [torch.diag_embed(-f * f * s), torch.diag_embed(-f * f * c)]
    return torch.cat(feats, dim=1), torch.cat(d1, dim=1), torch.cat(d2, dim=1)


def fold(v: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    return v * (g / v.norm(dim=1, keepdim=True))


def sdf_value_grad_backward(x: torch.Tensor, vs: Sequence[torch.Tensor], gs: Sequence[torch.Tensor],
                            bs: Sequence[torch.Tensor], skip_in: Sequence[int], n_freqs: int, g_full: torch.Tensor,
                            g_grad: torch.Tensor, trace: dict = None) -> Tuple[torch.Tensor, List[torch.Tensor], List[torch.Tensor], List[torch.Tensor]]:
    """x [P,3]; vs/gs/bs = weight_v [out,in] / weight_g [out,1] / bias [out] per layer; g_full [P, 2+F]; g_grad [P,3].
    Returns (dx [P,3], [dv_l], [dg_l], [db_l]).  `trace` (optional dict) receives the per-layer intermediates
    H, T (layer inputs), DZ, DS (gradients w.r.t. the pre-activations) and DW (w.r.t. the folded weights) -- what the
    kernels of csrc/mlp_bwd_kernel.cuh keep in their saved / dumped images (tests/diag/diag_backward.py compares them)."""
    n = len(vs)
    W = [fold(v, g) for v, g in zip(vs, gs)]
    pe, dpe, d2pe = _pe_all(x, n_freqs)
    inv_sqrt2 = 1.0 / math.sqrt(2.0)
    # ---------------- forward, keeping what the reverse sweep reads
    H, T, Z, S = [], [], [], []
    h, t = pe, dpe                                            # [P,D], [P,D,3]
    for l in range(n):
        if l in skip_in:
            h = torch.cat([h, pe], dim=1) * inv_sqrt2
            t = torch.cat([t, dpe], dim=1) * inv_sqrt2
        H.append(h)
        T.append(t)
        z = h @ W[l].T + bs[l]
        s = torch.einsum("oi,pij->poj", W[l], t)
        Z.append(z)
        S.append(s)
        if l < n - 1:
            h = _sp(z)
            t = _sp1(z).unsqueeze(-1) * s
    # ---------------- reverse
    dW = [None] * n
    db = [None] * n
    dz = g_full.clone()
    ds = torch.zeros_like(S[-1])
    ds[:, 0, :] = g_grad
    d_pe = torch.zeros_like(pe)                               # gradient reaching the positional encoding (value path)
    d_dpe = torch.zeros_like(dpe)                             # ... and its Jacobian columns (tangent path)
    DZ, DS = [None] * n, [None] * n
    for l in range(n - 1, -1, -1):
        DZ[l], DS[l] = dz, ds
        dW[l] = dz.T @ H[l] + torch.einsum("poj,pij->oi", ds, T[l])
        db[l] = dz.sum(dim=0)
        dh = dz @ W[l]
        dt = torch.einsum("oi,poj->pij", W[l], ds)
        if l in skip_in:
            k = H[l].shape[1] - pe.shape[1]
            d_pe = d_pe + dh[:, k:] * inv_sqrt2
            d_dpe = d_dpe + dt[:, k:] * inv_sqrt2
            dh, dt = dh[:, :k] * inv_sqrt2, dt[:, :k] * inv_sqrt2
        if l == 0:
            d_pe = d_pe + dh
            d_dpe = d_dpe + dt
        else:
            zp = Z[l - 1]
            ds = _sp1(zp).unsqueeze(-1) * dt
            dz = _sp1(zp) * dh + _sp2(zp) * (S[l - 1] * dt).sum(dim=-1)
    # PE: feature d depends on coordinate c(d) only, so J^T and the second derivative are sums over features
    dx = torch.einsum("pd,pdj->pj", d_pe, dpe) + torch.einsum("pdj,pdj->pj", d_dpe, d2pe)
    if trace is not None:
        trace.update(H=H, T=T, DZ=DZ, DS=DS, DW=dW, DB=db)
    # ---------------- weight norm
    dv, dg = [], []
    for l in range(n):
        nv = vs[l].norm(dim=1, keepdim=True)
        vhat = vs[l] / nv
        dot = (dW[l] * vhat).sum(dim=1, keepdim=True)
        dg.append(dot)
        dv.append((gs[l] / nv) * (dW[l] - dot * vhat))
    return dx, dv, dg, db


def render_backward(points: torch.Tensor, normals: torch.Tensor, view: torch.Tensor, feats: torch.Tensor,
                    vs: Sequence[torch.Tensor], gs: Sequence[torch.Tensor], bs: Sequence[torch.Tensor], n_freqs_view: int,
                    g_rgb: torch.Tensor):
    """Reverse sweep through RenderingNetwork.forward, mode 'idr' (:145-167): input cat[points, PE(view), normals, feats],
    ReLU hidden layers, tanh output.  Returns (d_points, d_normals, d_view, d_feats, [dv_l], [dg_l], [db_l]); d_normals
    is what enters sdf_value_grad_backward as g_grad, d_feats as g_full[:, 2:] (get_rbg_value, :324-338)."""
    n = len(vs)
    W = [fold(v, g) for v, g in zip(vs, gs)]
    pe, dpe, _ = _pe_all(view, n_freqs_view)
    h = torch.cat([points, pe, normals, feats], dim=1)
    H, Z = [], []
    for l in range(n):
        H.append(h)
        z = h @ W[l].T + bs[l]
        Z.append(z)
        h = torch.relu(z) if l < n - 1 else torch.tanh(z)
    dz = g_rgb * (1 - h * h)
    dW, db = [None] * n, [None] * n
    for l in range(n - 1, -1, -1):
        dW[l] = dz.T @ H[l]
        db[l] = dz.sum(dim=0)
        dh = dz @ W[l]
        if l > 0:
            dz = dh * (Z[l - 1] > 0).to(dh.dtype)
    d_pe = dh[:, 3:3 + pe.shape[1]]
    d_view = torch.einsum("pd,pdj->pj", d_pe, dpe)
    o = 3 + pe.shape[1]
    dv, dg = [], []
    for l in range(n):
        nv = vs[l].norm(dim=1, keepdim=True)
        vhat = vs[l] / nv
        dot = (dW[l] * vhat).sum(dim=1, keepdim=True)
        dg.append(dot)
        dv.append((gs[l] / nv) * (dW[l] - dot * vhat))
    return dh[:, :3], dh[:, o:o + 3], d_view, dh[:, o + 3:], dv, dg, db
