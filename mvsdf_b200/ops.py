"""Torch-facing wrappers over the C ABI: device memory and streams come from PyTorch
(plumbing), every computation is a call into libmvsdf_b200.so."""
from __future__ import annotations

from ctypes import c_void_p
from typing import Dict, List, Optional

import torch

from . import _lib

HEAD_SDF_ONLY = 0
HEAD_FULL = 1
HEAD_SDF_SCREEN = 2     # SDF column at screening precision (the tracer's prefilter; include/mvsdf_b200.h)


def _stream(device) -> c_void_p:
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _f32(t: torch.Tensor) -> torch.Tensor:
    assert t.is_cuda, "mvsdf_b200 operates on CUDA tensors only (no CPU fallback)"
    return t.detach().to(torch.float32).contiguous()


class PackedNet:
    """A layout plan (host) plus its packed fp16 hi/lo weight blob (device)."""

    def __init__(self, kind: str, width: int, n_hidden: int, feature_size: int = 256, skip_layer: int = 4,
                 n_freqs: int = 6):
        L = _lib.lib()
        self.kind = kind
        self.feature_size = feature_size
        if kind == "sdf":
            self.handle = L.mvsdf_sdf_net_create(width, n_hidden, skip_layer, n_freqs, feature_size)
        elif kind == "render":
            self.handle = L.mvsdf_render_net_create(width, n_hidden, n_freqs, feature_size)
        else:
            raise ValueError(kind)
        if not self.handle:
            raise _lib.MvsdfError(L.mvsdf_last_error().decode())
        self.n_layers = L.mvsdf_net_num_layers(self.handle)
        self.nbytes = L.mvsdf_net_packed_bytes(self.handle)
        self.status_off = L.mvsdf_net_status_offset(self.handle)
        self.blob: Optional[torch.Tensor] = None
        self.blob_t: Optional[torch.Tensor] = None      # transposed blob of the native backward (packed lazily, per pack())
        self._t_valid = False

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.lib().mvsdf_net_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def pack(self, weight_v: List[torch.Tensor], weight_g: List[Optional[torch.Tensor]], bias: List[torch.Tensor]):
        """mvsdf_pack_weights: fold weight-norm and write the tiled fp16 hi/lo blob."""
        assert len(weight_v) == self.n_layers == len(bias)
        dev = weight_v[0].device
        if self.blob is None or self.blob.device != dev:
            self.blob = torch.empty(self.nbytes, dtype=torch.uint8, device=dev)
        vs = [_f32(w) for w in weight_v]
        gs = [None if g is None else _f32(g) for g in weight_g]
        bs = [_f32(b) for b in bias]
        self._keep = (vs, gs, bs)   # keep alive until the stream has consumed them
        self._t_valid = False
        _lib.check(_lib.lib().mvsdf_pack_weights(self.handle, _lib.ptr_array(vs), _lib.ptr_array(gs),
                                                 _lib.ptr_array(bs), _lib.ptr(self.blob), _stream(dev)))
        return self

    def pack_t(self):
        """mvsdf_pack_weights_t: the transposed fp16 hi/lo tiles the reverse sweep streams (W^T, same fold / scales);
        built on first use after every pack()."""
        if self._t_valid:
            return self
        L = _lib.lib()
        vs, gs, _ = self._keep
        dev = vs[0].device
        if self.blob_t is None or self.blob_t.device != dev:
            self.blob_t = torch.zeros(L.mvsdf_train_packed_t_bytes(self.handle), dtype=torch.uint8, device=dev)
        _lib.check(L.mvsdf_pack_weights_t(self.handle, _lib.ptr_array(vs), _lib.ptr_array(gs), _lib.ptr(self.blob_t), _stream(dev)))
        self._t_valid = True
        return self

    def status(self) -> torch.Tensor:
        """int32 view [4] of the blob's range monitors (include/mvsdf_b200.h MVSDF_STATUS_*); device tensor, no sync."""
        return self.blob[self.status_off:self.status_off + 16].view(torch.int32)

    def check_status(self):
        """Host read (synchronises) of the range monitors; raises when the fp16 hi/lo representation was exceeded."""
        st = self.status().cpu()
        if int(st[0]) != 0:
            raise _lib.MvsdfError(f"{self.kind} net: {int(st[0])} packed weight elements have |64*W| beyond the fp16 range "
                                  "(|W| >= 1023.5 or non-finite): this network cannot be evaluated with the fp16 hi/lo split")
        if int(st[1]) != 0:
            raise _lib.MvsdfError(f"{self.kind} net: {int(st[1])} non-finite outputs since the last pack -- an activation "
                                  "left the fp16 hi/lo range (|x| >= 1023) or the inputs were non-finite")

    def pack_state_dict(self, sd: Dict[str, torch.Tensor], prefix: str, device):
        vs, gs, bs = [], [], []
        for l in range(self.n_layers):
            if f"{prefix}.lin{l}.weight_v" in sd:
                vs.append(sd[f"{prefix}.lin{l}.weight_v"].to(device))
                gs.append(sd[f"{prefix}.lin{l}.weight_g"].to(device))
            else:
                vs.append(sd[f"{prefix}.lin{l}.weight"].to(device))
                gs.append(None)
            bs.append(sd[f"{prefix}.lin{l}.bias"].to(device))
        return self.pack(vs, gs, bs)


def sdf_forward(net: PackedNet, x: torch.Tensor, head: int = HEAD_FULL):
    """ImplicitNetwork.forward: [n,3] -> sdf [n] (HEAD_SDF_ONLY) or full [n, 2+F] (HEAD_FULL)."""
    x = _f32(x)
    n = x.shape[0]
    if head in (HEAD_SDF_ONLY, HEAD_SDF_SCREEN):
        out = torch.empty(n, dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().mvsdf_sdf_forward(net.handle, _lib.ptr(net.blob), _lib.ptr(x), n, None, head,
                                                _lib.ptr(out), None, _stream(x.device)))
        return out
    out = torch.empty(n, net.feature_size + 2, dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().mvsdf_sdf_forward(net.handle, _lib.ptr(net.blob), _lib.ptr(x), n, None, head, None,
                                            _lib.ptr(out), _stream(x.device)))
    return out


def sdf_value_grad(net: PackedNet, x: torch.Tensor, head: int = HEAD_FULL):
    """ImplicitNetwork.forward + .gradient fused: returns (sdf [n] or full [n,2+F], grad [n,3])."""
    x = _f32(x)
    n = x.shape[0]
    grad = torch.empty(n, 3, dtype=torch.float32, device=x.device)
    if head == HEAD_SDF_ONLY:
        out = torch.empty(n, dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().mvsdf_sdf_value_grad(net.handle, _lib.ptr(net.blob), _lib.ptr(x), n, None, head,
                                                   _lib.ptr(out), None, _lib.ptr(grad), _stream(x.device)))
    else:
        out = torch.empty(n, net.feature_size + 2, dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().mvsdf_sdf_value_grad(net.handle, _lib.ptr(net.blob), _lib.ptr(x), n, None, head, None,
                                                   _lib.ptr(out), _lib.ptr(grad), _stream(x.device)))
    return out, grad


def render_forward(net: PackedNet, points, view_dirs, normals, features):
    """RenderingNetwork.forward: -> rgb [n,3] in [-1,1]."""
    points, view_dirs, normals, features = _f32(points), _f32(view_dirs), _f32(normals), _f32(features)
    n = points.shape[0]
    rgb = torch.empty(n, 3, dtype=torch.float32, device=points.device)
    _lib.check(_lib.lib().mvsdf_render_forward(net.handle, _lib.ptr(net.blob), _lib.ptr(points), _lib.ptr(view_dirs),
                                               _lib.ptr(normals), _lib.ptr(features), n, None, _lib.ptr(rgb),
                                               _stream(points.device)))
    return rgb


# ---- training step: saving forward, native backward, weight-norm chain (csrc/train_abi.cu, csrc/mlp_bwd_kernel.cuh) ----
_POOL: Dict[str, torch.Tensor] = {}


def _scratch(name: str, nbytes: int, device) -> torch.Tensor:
    """Reusable byte buffer (the backward workspaces are hundreds of MB: allocate once, grow on demand)."""
    t = _POOL.get(name)
    if t is None or t.numel() < nbytes or t.device != device:
        t = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=device)
        _POOL[name] = t
    return t


def sdf_forward_train(net: PackedNet, x: torch.Tensor):
    """ImplicitNetwork.forward + .gradient with the layer inputs saved for the backward: (full [n,2+F], grad [n,3], save)."""
    L = _lib.lib()
    x = _f32(x)
    n = x.shape[0]
    full = torch.empty(n, net.feature_size + 2, dtype=torch.float32, device=x.device)
    grad = torch.empty(n, 3, dtype=torch.float32, device=x.device)
    nbytes = L.mvsdf_train_save_bytes(net.handle, n, 1)
    save = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    if n > 0:
        _lib.check(L.mvsdf_sdf_forward_train(net.handle, _lib.ptr(net.blob), _lib.ptr(x), n, nbytes, _lib.ptr(save), _lib.ptr(full),
                                             _lib.ptr(grad), _stream(x.device)))
    return full, grad, save


def sdf_backward(net: PackedNet, x: torch.Tensor, save: torch.Tensor, g_full, g_grad, need_dx: bool, need_dw: bool = True):
    """Reverse sweep + dW GEMM of the SDF net: returns (dx [n,3] or None, dw [plan floats], db [plan floats]);
    need_dw=False: dx-only sweep (dw = db = None)."""
    L = _lib.lib()
    x = _f32(x)
    n = x.shape[0]
    dev = x.device
    net.pack_t()
    dw = torch.empty(L.mvsdf_train_dw_floats(net.handle), dtype=torch.float32, device=dev) if need_dw else None
    db = torch.empty(L.mvsdf_train_db_floats(net.handle), dtype=torch.float32, device=dev) if need_dw else None
    dx = torch.empty(n, 3, dtype=torch.float32, device=dev) if need_dx else None
    ws_bytes = L.mvsdf_train_workspace_bytes(net.handle, n if need_dw else 0, 1)      # n = 0: header + sweep scratch only
    ws = _scratch("sdf_bwd", ws_bytes, dev)
    g_full = None if g_full is None else _f32(g_full)
    g_grad = None if g_grad is None else _f32(g_grad)
    _lib.check(L.mvsdf_sdf_backward(net.handle, _lib.ptr(net.blob_t), _lib.ptr(x), n, _lib.ptr(save), _lib.ptr(g_full), _lib.ptr(g_grad),
                                    ws.numel(), _lib.ptr(ws), _lib.ptr(dx), _lib.ptr(dw), _lib.ptr(db), _stream(dev)))
    return dx, dw, db


def render_forward_train(net: PackedNet, points, view_dirs, normals, features):
    L = _lib.lib()
    points, view_dirs, normals, features = _f32(points), _f32(view_dirs), _f32(normals), _f32(features)
    n = points.shape[0]
    rgb = torch.empty(n, 3, dtype=torch.float32, device=points.device)
    nbytes = L.mvsdf_train_save_bytes(net.handle, n, 0)
    save = torch.empty(nbytes, dtype=torch.uint8, device=points.device)
    if n > 0:
        _lib.check(L.mvsdf_render_forward_train(net.handle, _lib.ptr(net.blob), _lib.ptr(points), _lib.ptr(view_dirs), _lib.ptr(normals),
                                                _lib.ptr(features), n, nbytes, _lib.ptr(save), _lib.ptr(rgb), _stream(points.device)))
    return rgb, save


def render_backward(net: PackedNet, save: torch.Tensor, rgb: torch.Tensor, g_rgb: torch.Tensor, view_dirs: Optional[torch.Tensor] = None):
    """Reverse sweep + dW GEMM of the rendering net: (d_points, d_normals, d_feats, d_view or None, dw, db); d_view is
    produced when view_dirs is given (trained camera poses)."""
    L = _lib.lib()
    rgb, g_rgb = _f32(rgb), _f32(g_rgb)
    n = rgb.shape[0]
    dev = rgb.device
    net.pack_t()
    F = net.feature_size
    dw = torch.empty(L.mvsdf_train_dw_floats(net.handle), dtype=torch.float32, device=dev)
    db = torch.empty(L.mvsdf_train_db_floats(net.handle), dtype=torch.float32, device=dev)
    d_points = torch.empty(n, 3, dtype=torch.float32, device=dev)
    d_normals = torch.empty(n, 3, dtype=torch.float32, device=dev)
    d_feats = torch.empty(n, F, dtype=torch.float32, device=dev)
    view = None if view_dirs is None else _f32(view_dirs)
    d_view = None if view is None else torch.empty(n, 3, dtype=torch.float32, device=dev)
    ws_bytes = L.mvsdf_train_workspace_bytes(net.handle, n, 0)
    ws = _scratch("render_bwd", ws_bytes, dev)
    _lib.check(L.mvsdf_render_backward(net.handle, _lib.ptr(net.blob_t), n, _lib.ptr(save), _lib.ptr(rgb), _lib.ptr(g_rgb), _lib.ptr(view),
                                       ws.numel(), _lib.ptr(ws), _lib.ptr(d_points), _lib.ptr(d_normals), _lib.ptr(d_feats),
                                       _lib.ptr(d_view), _lib.ptr(dw), _lib.ptr(db), _stream(dev)))
    return d_points, d_normals, d_feats, d_view, dw, db


def weight_grads(net: PackedNet, dw: torch.Tensor, db: torch.Tensor, vs: List[torch.Tensor], gs: List[Optional[torch.Tensor]]):
    """Weight-norm chain: plan-coordinate dW / db -> per-layer (dv [out,in], dg [out,1], dbias [out]) lists."""
    L = _lib.lib()
    dev = dw.device
    vs32 = [_f32(v) for v in vs]
    gs32 = [None if g is None else _f32(g) for g in gs]
    dvs = [torch.empty_like(v) for v in vs32]
    dgs = [None if g is None else torch.empty_like(g) for g in gs32]
    dbs = [torch.empty(v.shape[0], dtype=torch.float32, device=dev) for v in vs32]
    _lib.check(L.mvsdf_weight_grads(net.handle, _lib.ptr(dw), _lib.ptr(db), _lib.ptr_array(vs32), _lib.ptr_array(gs32),
                                    _lib.ptr_array(dvs), _lib.ptr_array(dgs), _lib.ptr_array(dbs), _stream(dev)))
    return dvs, dgs, dbs
