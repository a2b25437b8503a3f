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
        _lib.check(_lib.lib().mvsdf_pack_weights(self.handle, _lib.ptr_array(vs), _lib.ptr_array(gs),
                                                 _lib.ptr_array(bs), _lib.ptr(self.blob), _stream(dev)))
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
