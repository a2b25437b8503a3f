"""Host-side mirror of IDRLoss for the terms on the north-star path (code/model/loss.py):
get_rgb_loss (:21-28), get_feat_loss_corr (:115-165, forward and backward) and the depth-carving term
get_depth_loss (:37-63, my_utils.carving_t2) run in libmvsdf_b200.so; the eikonal and surface-indicator
terms are tiny reductions on tensors the model already produced."""
from __future__ import annotations

from ctypes import c_void_p
from typing import Dict, Optional

import torch
from torch import nn

from . import _lib, conf as default_schedule, ops


class FeatureStore:
    """Channels-last, device-resident copies of the (constant) CNN feature maps -- the store the warp kernel reads.
    The reference keeps NCHW maps on the host (self.feats, scene_dataset.py:138-149), uploads feats[idx] / feats[src_idxs]
    for every item of every step (:205-207) and pays 32 sectors per bilinear tap.  Two ways in:

    * ``get(feat, feat_src)``: the reference's per-batch tensors ([B,32,h,w], [B,S,32,h,w], host or device).  Restacked to
      [B,V,h,w,32] with mvsdf_feat_nchw_to_nhwc on first sight and kept in a small LRU keyed on the IDENTITY of the
      caller's tensors (the originals are pinned so their storage cannot be recycled under the key): a caller that
      passes the same tensors again -- eval over one image in chunks, bench.py's steps -- neither re-uploads nor
      re-transposes anything.
    * ``register_scene(feats)``: all maps of a scene once ([n,32,h,w] -> [n,h,w,32]); batches then name their maps by
      index (ground_truth["feat_index"] [B], ["src_index"] [B,S]) and the kernels read the store through a [B,V] index
      table (mvsdf_feat_loss_partials_indexed): nothing is copied per step.  FeatExt (mvsdf_b200/featext.py) writes
      this layout directly."""

    MAX_ENTRIES = 4

    def __init__(self):
        self._lru = []            # [(key, originals, maps)]
        self.scene_maps: Optional[torch.Tensor] = None
        self.restacks = 0         # number of NCHW -> channels-last conversions performed (tests / bench read it)

    @staticmethod
    def _key(t: torch.Tensor):
        return (id(t), t.data_ptr(), tuple(t.shape), t._version, str(t.device), t.dtype)

    def _restack(self, src: torch.Tensor, dst: torch.Tensor):
        """src [n,C,h,w] (device fp32 contiguous) -> dst [n,h,w,C]."""
        n, C, h, w = src.shape
        stream = c_void_p(torch.cuda.current_stream(src.device).cuda_stream)
        _lib.check(_lib.lib().mvsdf_feat_nchw_to_nhwc(c_void_p(src.data_ptr()), n, C, h, w, c_void_p(dst.data_ptr()), stream))
        self.restacks += 1

    def get(self, feat: torch.Tensor, feat_src: torch.Tensor, device) -> torch.Tensor:
        key = (self._key(feat), self._key(feat_src))
        for i, (k, _, maps) in enumerate(self._lru):
            if k == key and maps.device == device:
                self._lru.append(self._lru.pop(i))
                return maps
        B, C, h, w = feat.shape
        S = feat_src.shape[1]
        f_dev = ops._f32(feat.to(device, non_blocking=True))
        fs_dev = ops._f32(feat_src.to(device, non_blocking=True))
        maps = torch.empty(B, 1 + S, h, w, C, dtype=torch.float32, device=device)
        for b in range(B):
            self._restack(f_dev[b:b + 1], maps[b, 0:1])
            self._restack(fs_dev[b], maps[b, 1:])
        # the key tensors themselves are held: their ids / addresses cannot be handed to another batch while cached
        self._lru.append((key, (feat, feat_src), maps))
        if len(self._lru) > self.MAX_ENTRIES:
            self._lru.pop(0)
        return maps

    def register_scene(self, feats: torch.Tensor, device=None) -> torch.Tensor:
        """feats [n,32,h,w] (what FeatExt returns at its finest scale for the n images of a scene) -> resident [n,h,w,32]."""
        device = torch.device(device) if device is not None else feats.device
        src = ops._f32(feats.to(device))
        n, C, h, w = src.shape
        self.scene_maps = torch.empty(n, h, w, C, dtype=torch.float32, device=device)
        self._restack(src, self.scene_maps)
        return self.scene_maps

    def set_scene_maps(self, maps_nhwc: torch.Tensor):
        """Adopt maps that already are channels-last [n,h,w,32] (FeatExt's native writer)."""
        assert maps_nhwc.dim() == 4 and maps_nhwc.is_cuda
        self.scene_maps = ops._f32(maps_nhwc)
        return self.scene_maps


class B200IDRLoss(nn.Module):
    def __init__(self, schedule=None):
        super().__init__()
        self.schedule = schedule if schedule is not None else default_schedule
        self.store = FeatureStore()
        self.last_partials: Dict[str, torch.Tensor] = {}

    # ---- loss.py:21-28
    def get_rgb_loss(self, rgb_values, rgb_gt, network_object_mask, object_mask, reduce_fn=None):
        mask = network_object_mask & object_mask
        if torch.is_grad_enabled() and rgb_values.requires_grad:
            from .autograd import RgbL1
            return RgbL1.apply(self, rgb_values, rgb_gt.to(rgb_values.device), mask, reduce_fn)
        return self._rgb_loss_native(rgb_values, rgb_gt, mask, reduce_fn)

    @torch.no_grad()
    def _rgb_loss_native(self, rgb_values, rgb_gt, mask, reduce_fn=None):
        L = _lib.lib()
        dev = rgb_values.device
        rgb_values = ops._f32(rgb_values)
        rgb_gt = ops._f32(rgb_gt.to(dev)).reshape(-1, 3)
        mask = mask.to(torch.uint8).contiguous()
        R = mask.shape[0]
        partial = torch.empty(2, dtype=torch.float64, device=dev)
        out = torch.empty((), dtype=torch.float32, device=dev)
        stream = c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(L.mvsdf_rgb_l1_partials(_lib.ptr(rgb_values), _lib.ptr(rgb_gt), _lib.ptr(mask), R, _lib.ptr(partial), stream))
        if reduce_fn is not None:
            reduce_fn(partial)          # multi-GPU: one all-reduce(SUM) of the partials
        _lib.check(L.mvsdf_rgb_l1_finalize(_lib.ptr(partial), _lib.ptr(out), stream))
        self.last_partials["rgb"] = partial
        return out

    # ---- loss.py:115-165 (uncerts is never produced by the reference: uncert_network is not instantiated)
    def get_feat_loss_corr(self, diff_surf_pts, uncerts, feat, cam, feat_src, src_cams, size, center,
                           network_object_mask, object_mask, hit_offsets: Optional[torch.Tensor] = None, reduce_fn=None,
                           feat_index: Optional[torch.Tensor] = None, src_index: Optional[torch.Tensor] = None):
        """feat / feat_src: the reference's tensors (host or device; cached channels-last in self.store), or None when the
        scene store is used (self.store.register_scene + feat_index [B], src_index [B,S])."""
        if uncerts is not None:
            raise NotImplementedError("the uncertainty branch (loss.py:156-159) is dead code in the reference")
        dev = diff_surf_pts.device
        B = cam.shape[0]
        if hit_offsets is None:
            m = (network_object_mask & object_mask).view(B, -1).sum(-1)
            hit_offsets = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), m.cumsum(0)]).to(torch.int32)
        hit_offsets = hit_offsets.to(device=dev, dtype=torch.int32).contiguous()
        if feat is None:
            if self.store.scene_maps is None or feat_index is None or src_index is None:
                raise _lib.MvsdfError("get_feat_loss_corr: pass feat / feat_src tensors, or register_scene() + feat_index / src_index")
            maps = self.store.scene_maps
            map_index = torch.cat([feat_index.reshape(B, 1), src_index.reshape(B, -1)], dim=1).to(device=dev, dtype=torch.int32).contiguous()
        else:
            b4 = feat.dim() == 4
            maps = self.store.get(feat if b4 else feat.unsqueeze(0), feat_src if b4 else feat_src.unsqueeze(0), dev)
            map_index = None
        args = (diff_surf_pts, hit_offsets, maps, map_index, cam.to(dev), src_cams.to(dev), size.to(dev), center.to(dev), reduce_fn)
        if torch.is_grad_enabled() and diff_surf_pts.requires_grad:
            from .autograd import FeatConsistency
            return FeatConsistency.apply(self, *args)
        return self._feat_loss_native(*args)

    @torch.no_grad()
    def _feat_loss_native(self, diff_surf_pts, hit_offsets, maps, map_index, cam, src_cams, size, center, reduce_fn=None):
        L = _lib.lib()
        dev = diff_surf_pts.device
        B = cam.shape[0]
        if map_index is None:
            _, V, h, w, C = maps.shape
        else:
            V = map_index.shape[1]
            _, h, w, C = maps.shape
        cams = torch.cat([cam.to(dev).unsqueeze(1), src_cams.to(dev)], dim=1).to(torch.float32).contiguous()   # [B,V,2,4,4]
        pts = ops._f32(diff_surf_pts)
        if pts.numel() == 0:            # no surface point at all: the kernel reads M = 0 from hit_offsets and touches nothing
            pts = torch.zeros(1, 3, dtype=torch.float32, device=dev)
        size = ops._f32(size.to(dev)).reshape(-1)[:1].contiguous()
        center = ops._f32(center.to(dev)).reshape(-1)[:3].contiguous()
        partial = torch.empty(B, 2, dtype=torch.float64, device=dev)
        out = torch.empty((), dtype=torch.float32, device=dev)
        stream = c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(L.mvsdf_feat_loss_partials_indexed(_lib.ptr(pts), _lib.ptr(hit_offsets), _lib.ptr(cams), _lib.ptr(maps),
                                                      _lib.ptr(map_index), B, V, h, w, C, _lib.ptr(size), _lib.ptr(center),
                                                      _lib.ptr(partial), stream))
        if reduce_fn is not None:
            reduce_fn(partial)
        _lib.check(L.mvsdf_feat_loss_finalize(_lib.ptr(partial), B, _lib.ptr(out), stream))
        self.last_partials["feat"] = partial
        self._feat_ctx = (pts, hit_offsets, cams, maps, size, center)      # operands of the native backward
        self._feat_map_index = map_index
        return out

    @torch.no_grad()
    def _feat_loss_backward_native(self, ctx_tensors, partial, upstream, map_index=None):
        """d loss / d diff_surf_pts through mvsdf_feat_loss_backward (autograd of loss.py:132-155 in the reference)."""
        pts, hit_offsets, cams, maps, size, center = ctx_tensors
        L = _lib.lib()
        dev = pts.device
        B, V = cams.shape[0], cams.shape[1]
        h, w, C = maps.shape[-3:]
        grad = torch.empty_like(pts)
        up = upstream.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        stream = c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(L.mvsdf_feat_loss_backward_indexed(_lib.ptr(pts), _lib.ptr(hit_offsets), _lib.ptr(cams), _lib.ptr(maps),
                                                      _lib.ptr(map_index), B, V, h, w, C, _lib.ptr(size), _lib.ptr(center),
                                                      _lib.ptr(partial), _lib.ptr(up), _lib.ptr(grad), stream))
        return grad

    # ---- loss.py:37-63 (+ carving_t2 / RunningTopK, utils/my_utils.py:168-201, :269-331)
    def get_depth_loss(self, eikonal_points_hom, eikonal_output, depths, cams, size, center, far_thresh=None, far_att=None,
                       near_thresh=None, near_att=None, smooth=None, train_progress=None, reduce_fn=None):
        """Same positional signature as the reference; the attenuation arguments default to the schedule's values at
        `train_progress`.  Unlike the reference, eikonal_points_hom is NOT rewritten in place (loss.py:42 does that through
        detach(), a side effect nothing reads)."""
        conf = self.schedule
        if conf.use_invalid:
            raise NotImplementedError("use_invalid=True (carving_t, my_utils.py:204-266) is off in model/conf.py:16")
        if smooth is not None:
            raise NotImplementedError("the SmoothL1 variant (loss.py:57-58) is never selected: conf.smooth(tp) is None")
        tp = 1.0 if train_progress is None else train_progress
        att = (conf.far_thresh if far_thresh is None else far_thresh, conf.far_att(tp) if far_att is None else far_att,
               conf.near_thresh if near_thresh is None else near_thresh, conf.near_att(tp) if near_att is None else near_att)
        dev = eikonal_output.device
        args = (eikonal_points_hom, eikonal_output, depths.to(dev), cams.to(dev), size.to(dev), center.to(dev), att, reduce_fn)
        if torch.is_grad_enabled() and eikonal_output.requires_grad:
            from .autograd import DepthL1
            return DepthL1.apply(self, *args)
        return self._depth_loss_native(*args)[0]

    @torch.no_grad()
    def _depth_loss_native(self, eikonal_points_hom, eikonal_output, depths, cams, size, center, att, reduce_fn=None):
        L = _lib.lib()
        dev = eikonal_output.device
        pts = ops._f32(eikonal_points_hom).reshape(-1, 4)
        f = ops._f32(eikonal_output).reshape(-1)
        E = f.shape[0]
        d = ops._f32(depths).reshape(-1, *depths.shape[-2:])                     # nv1hw -> [V,h,w] (v = 1 per image)
        c = ops._f32(cams).reshape(-1, 2, 4, 4)
        V, h, w = d.shape
        size = ops._f32(size).reshape(-1)[:1].contiguous()
        center = ops._f32(center).reshape(-1)[:3].contiguous()
        target = torch.empty(E, dtype=torch.float32, device=dev)
        weight = torch.empty(E, dtype=torch.float32, device=dev)
        partial = torch.empty(2, dtype=torch.float64, device=dev)
        out = torch.empty((), dtype=torch.float32, device=dev)
        stream = c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(L.mvsdf_depth_loss_partials(_lib.ptr(pts), 4, _lib.ptr(f), E, _lib.ptr(d), _lib.ptr(c), V, h, w, _lib.ptr(size),
                                               _lib.ptr(center), float(self.schedule.out_thresh_perc), float(att[0]), float(att[1]),
                                               float(att[2]), float(att[3]), _lib.ptr(target), _lib.ptr(weight), _lib.ptr(partial),
                                               stream))
        if reduce_fn is not None:
            reduce_fn(partial)
        _lib.check(L.mvsdf_rgb_l1_finalize(_lib.ptr(partial), _lib.ptr(out), stream))
        self.last_partials["depth"] = partial
        return out, target, weight

    # ---- loss.py:30-35, :167-174 (elementwise reductions on small tensors)
    def get_eikonal_loss(self, grad_theta, reduce_fn=None):
        """loss.py:30-35.  With reduce_fn (multi-GPU) the mean runs over the eikonal points of ALL ranks: (sum, count)
        partials through the same all-reduce as the other terms (parallel.mean_from_partials)."""
        from .parallel import mean_from_partials
        if grad_theta.shape[0] == 0 and reduce_fn is None:
            return torch.tensor(0.0, device=grad_theta.device)
        sq = (grad_theta.norm(2, dim=1) - 1) ** 2
        return mean_from_partials(sq.sum(), grad_theta.shape[0], reduce_fn)

    def get_surf_loss(self, surf_indicator_output, network_object_mask, object_mask_true, reduce_fn=None):
        """loss.py:167-174: BCE-with-logits, ones for the surface hits inside the true mask, zeros for the eikonal samples;
        mean over all entries (of all ranks with reduce_fn)."""
        from .parallel import mean_from_partials
        n = int((network_object_mask & object_mask_true).sum())
        gt = torch.cat([torch.ones(n), torch.zeros(surf_indicator_output.shape[0] - n)]).to(surf_indicator_output)
        if reduce_fn is None:
            return nn.functional.binary_cross_entropy_with_logits(surf_indicator_output, gt)
        s = nn.functional.binary_cross_entropy_with_logits(surf_indicator_output, gt, reduction="sum")
        return mean_from_partials(s, surf_indicator_output.shape[0], reduce_fn)

    def hot_path_losses(self, model_outputs, ground_truth, train_progress, reduce_fn=None):
        """rgb L1 + feature consistency (+ eikonal / surface indicator when the forward ran in training mode)."""
        conf = self.schedule
        nm, om = model_outputs["network_object_mask"], model_outputs["object_mask"]
        dev = model_outputs["rgb_values"].device
        res = {}
        if conf.enable_rgb:
            res["rgb_loss"] = self.get_rgb_loss(model_outputs["rgb_values"], ground_truth["rgb"], nm, om, reduce_fn=reduce_fn)
        else:
            res["rgb_loss"] = torch.zeros(1, device=dev)
        if conf.phase[0] <= train_progress and conf.enable_feat:
            res["feat_loss"] = self.get_feat_loss_corr(
                model_outputs["diff_surf_pts"], model_outputs.get("uncerts"), ground_truth.get("feat"), ground_truth["cam"],
                ground_truth.get("feat_src"), ground_truth["src_cams"], ground_truth["size"][:1], ground_truth["center"][:1],
                nm, om, hit_offsets=model_outputs.get("hit_offsets"), reduce_fn=reduce_fn,
                feat_index=ground_truth.get("feat_index"), src_index=ground_truth.get("src_index"))
        else:
            res["feat_loss"] = torch.zeros(1, device=dev)
        if model_outputs.get("grad_theta") is not None:
            res["eikonal_loss"] = self.get_eikonal_loss(model_outputs["grad_theta"], reduce_fn=reduce_fn)
            res["surf_loss"] = self.get_surf_loss(model_outputs["surf_indicator_output"], nm, model_outputs["object_mask_true"],
                                                  reduce_fn=reduce_fn)
        return res

    def forward(self, model_outputs, ground_truth, train_progress, n_img=None, reduce_fn=None):
        """IDRLoss.forward (loss.py:176-219): the five terms, their schedule-dependent weights and the reference's dict."""
        conf = self.schedule
        dev = model_outputs["rgb_values"].device
        zero = lambda: torch.zeros(1, dtype=torch.float32, device=dev)
        part = self.hot_path_losses(model_outputs, ground_truth, train_progress, reduce_fn=reduce_fn)
        rgb_loss, feat_loss = part["rgb_loss"], part["feat_loss"]
        eikonal_loss = part["eikonal_loss"]
        depth_loss = self.get_depth_loss(model_outputs["eikonal_points_hom"], model_outputs["eikonal_output"],
                                         ground_truth["depths"], ground_truth["depth_cams"], ground_truth["size"],
                                         ground_truth["center"], train_progress=train_progress, reduce_fn=reduce_fn)
        surf_loss = part["surf_loss"] if conf.phase[0] <= train_progress else zero()      # :201-204
        loss = (rgb_loss * conf.rgb_weight(train_progress) + eikonal_loss * conf.eikonal_weight + surf_loss * conf.surf_weight
                + feat_loss * conf.feat_weight(train_progress) + depth_loss * conf.depth_weight(train_progress))
        return {"loss": loss, "rgb_loss": rgb_loss, "eikonal_loss": eikonal_loss, "depth_loss": depth_loss,
                "feat_loss": feat_loss, "surf_loss": surf_loss}
