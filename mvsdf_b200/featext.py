"""Host-side mirror of the reference's feature extractor (code/utils/my_utils.py:693-708 FeatExt; UNet :595-690) and the writer of
the scene feature store.

``B200FeatExt`` holds parameters / buffers under the reference's own names (``init_conv.0.weight``,
``unet.enc_blocks.2d2_0.0.conv1.weight``, ``unet.dec_blocks.2d16_3.0.weight``, ``final_conv_3.weight`` ...), so the slice of
``utils/vismvsnet.pt`` the reference loads (keys ``module.feat_ext.*``, :702-703) goes straight into ``load_state_dict``.  The
forward runs in libmvsdf_b200.so (csrc/featext.cu): BatchNorm folded once, channels-last activations, and the finest map comes
out as [image, h, w, 32] -- the layout the feature-warp kernel gathers from -- so ``FeatureStore.set_scene_maps`` adopts it
without a transpose.  Eval mode only (the reference freezes it: scene_dataset.py:139-142)."""
from __future__ import annotations

from ctypes import c_void_p
from typing import List, Optional, Tuple

import torch
from torch import nn

from . import _lib, ops

# (reference module path, kind) in the order csrc/featext.cu expects (kFe[]): conv weight [+ BatchNorm]
_BLOCKS = [("unet.enc_blocks.2d2_0", 16, 32, True), ("unet.enc_blocks.2d4_1", 32, 64, True), ("unet.enc_blocks.2d8_2", 64, 128, True)]


def _conv_specs() -> List[Tuple[str, Optional[str], Tuple[int, ...]]]:
    """[(conv weight key, bn prefix or None, weight shape)] for the 27 convolutions."""
    specs = [("init_conv.0.weight", "init_conv.1", (16, 3, 5, 5))]
    for prefix, cin, cout, _ in _BLOCKS:
        specs += [(f"{prefix}.0.conv1.weight", f"{prefix}.0.bn1", (cout, cin, 3, 3)),
                  (f"{prefix}.0.conv2.weight", f"{prefix}.0.bn2", (cout, cout, 3, 3)),
                  (f"{prefix}.0.downsample.0.weight", f"{prefix}.0.downsample.1", (cout, cin, 1, 1)),
                  (f"{prefix}.1.conv1.weight", f"{prefix}.1.bn1", (cout, cout, 3, 3)),
                  (f"{prefix}.1.conv2.weight", f"{prefix}.1.bn2", (cout, cout, 3, 3))]
    for prefix, cin, cout in (("unet.dec_blocks.2d16_3", 128, 64), ("unet.dec_blocks.2d8_4", 64, 32)):
        specs += [(f"{prefix}.0.weight", None, (cin, cout, 3, 3)),                 # ConvTranspose2d: [in, out, k, k]
                  (f"{prefix}.1.weight", None, (cout, 2 * cout, 3, 3)),
                  (f"{prefix}.2.0.conv1.weight", f"{prefix}.2.0.bn1", (cout, cout, 3, 3)),
                  (f"{prefix}.2.0.conv2.weight", f"{prefix}.2.0.bn2", (cout, cout, 3, 3))]
    specs += [("final_conv_1.weight", None, (32, 128, 3, 3)), ("final_conv_2.weight", None, (32, 64, 3, 3)),
              ("final_conv_3.weight", None, (32, 32, 3, 3))]
    return specs


class _Holder(nn.Module):
    """Plain container: gives parameters / buffers the dotted names of the reference's module tree."""

    def set(self, dotted: str, tensor: torch.Tensor, buffer: bool = False):
        head, _, rest = dotted.partition(".")
        if rest:
            if head not in self._modules:
                self.add_module(head, _Holder())
            self._modules[head].set(rest, tensor, buffer)
        elif buffer:
            self.register_buffer(head, tensor)
        else:
            self.register_parameter(head, nn.Parameter(tensor, requires_grad=False))


class B200FeatExt(_Holder):
    BN_EPS = 1e-5

    def __init__(self, seed: Optional[int] = None):
        """seed = None: zero-initialised holders (load a checkpoint); an int: seeded random weights / BatchNorm statistics
        (tests and benchmarks -- utils/vismvsnet.pt cannot travel)."""
        super().__init__()
        g = torch.Generator().manual_seed(seed) if seed is not None else None
        self._specs = _conv_specs()
        for wkey, bn, shape in self._specs:
            fan_in = shape[1] * shape[2] * shape[3]
            w = torch.randn(shape, generator=g) * (2.0 / fan_in) ** 0.5 if g is not None else torch.zeros(shape)
            self.set(wkey, w)
            if bn is not None:
                c = shape[0]
                self.set(bn + ".weight", 0.5 + torch.rand(c, generator=g) if g is not None else torch.ones(c))
                self.set(bn + ".bias", 0.2 * torch.randn(c, generator=g) if g is not None else torch.zeros(c))
                self.set(bn + ".running_mean", 0.2 * torch.randn(c, generator=g) if g is not None else torch.zeros(c), buffer=True)
                self.set(bn + ".running_var", 0.5 + torch.rand(c, generator=g) if g is not None else torch.ones(c), buffer=True)
                self.set(bn + ".num_batches_tracked", torch.zeros((), dtype=torch.long), buffer=True)
        self._packed: Optional[torch.Tensor] = None
        self._packed_key = None
        self.eval()

    def _tensor(self, dotted: str) -> torch.Tensor:
        obj = self
        for part in dotted.split("."):
            obj = obj._modules[part] if part in obj._modules else (obj._parameters.get(part) if part in obj._parameters else obj._buffers[part])
        return obj

    def _pack(self) -> torch.Tensor:
        L = _lib.lib()
        ws, bw, bb, bm, bv = [], [], [], [], []
        for wkey, bn, _ in self._specs:
            ws.append(ops._f32(self._tensor(wkey)))
            for lst, suffix in ((bw, ".weight"), (bb, ".bias"), (bm, ".running_mean"), (bv, ".running_var")):
                lst.append(None if bn is None else ops._f32(self._tensor(bn + suffix)))
        dev = ws[0].device
        key = tuple((t.data_ptr(), t._version) for t in ws + [t for t in bw + bb + bm + bv if t is not None])
        if self._packed is not None and self._packed_key == key and self._packed.device == dev:
            return self._packed
        assert L.mvsdf_featext_num_convs() == len(self._specs)
        packed = torch.empty(L.mvsdf_featext_packed_floats(), dtype=torch.float32, device=dev)
        self._keep = (ws, bw, bb, bm, bv)
        _lib.check(L.mvsdf_featext_pack(_lib.ptr_array(ws), _lib.ptr_array(bw), _lib.ptr_array(bb), _lib.ptr_array(bm), _lib.ptr_array(bv),
                                        float(self.BN_EPS), _lib.ptr(packed), c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        self._packed, self._packed_key = packed, key
        return packed

    @torch.no_grad()
    def forward_nhwc(self, x: torch.Tensor, all_scales: bool = False):
        """x [n,3,H,W] (normalised rgb, scene_dataset.py:118-131) -> finest feature map [n, H/2, W/2, 32] channels-last;
        all_scales: also the 1/8 and 1/4 maps ([n,H/8,W/8,32], [n,H/4,W/4,32], finest)."""
        if self.training:
            raise _lib.MvsdfError("B200FeatExt runs in eval mode only (frozen BatchNorm statistics, scene_dataset.py:139-142)")
        L = _lib.lib()
        x = ops._f32(x)
        n, c, H, W = x.shape
        assert c == 3
        dev = x.device
        packed = self._pack()
        f = dict(dtype=torch.float32, device=dev)
        out2 = torch.empty(n, H // 2, W // 2, 32, **f)
        out4 = torch.empty(n, H // 4, W // 4, 32, **f) if all_scales else None
        out8 = torch.empty(n, H // 8, W // 8, 32, **f) if all_scales else None
        ws_bytes = L.mvsdf_featext_workspace_bytes(n, H, W)
        ws = ops._scratch("featext", ws_bytes, dev)
        _lib.check(L.mvsdf_featext_forward(_lib.ptr(packed), _lib.ptr(x), n, H, W, ws.numel(), _lib.ptr(ws), _lib.ptr(out8), _lib.ptr(out4),
                                           _lib.ptr(out2), c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        return (out8, out4, out2) if all_scales else out2

    def forward(self, x: torch.Tensor):
        """The reference's signature: three NCHW maps (1/8, 1/4, 1/2 resolution; views of the channels-last results)."""
        o8, o4, o2 = self.forward_nhwc(x, all_scales=True)
        return o8.permute(0, 3, 1, 2), o4.permute(0, 3, 1, 2), o2.permute(0, 3, 1, 2)

    @torch.no_grad()
    def fill_store(self, store, images: torch.Tensor, batch: int = 20, device=None) -> torch.Tensor:
        """scene_dataset.py:143-149 (`feat_ext(eval_batch.cuda())[2]` in batches of 20, then torch.cat on the host): computes
        the finest map of every image of the scene straight into the resident channels-last store."""
        device = torch.device(device) if device is not None else next(self.parameters()).device
        n, _, H, W = images.shape
        maps = torch.empty(n, H // 2, W // 2, 32, dtype=torch.float32, device=device)
        for s in range(0, n, batch):
            maps[s:s + batch] = self.forward_nhwc(images[s:s + batch].to(device))
        return store.set_scene_maps(maps)
