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
    """Channels-last copy [B, V, h, w, 32] of the (constant) CNN feature maps of a mini-batch, built once
    per (feat, feat_src) pair with mvsdf_feat_nchw_to_nhwc.  The reference keeps NCHW maps
    (scene_dataset.py:149) and pays 32 sectors per bilinear tap."""

    def __init__(self):
        self._key = None
        self._maps = None

    def get(self, feat: torch.Tensor, feat_src: torch.Tensor) -> torch.Tensor:
        key = (feat.data_ptr(), feat_src.data_ptr(), tuple(feat.shape), tuple(feat_src.shape), feat._version,
               feat_src._version)
        if key == self._key:
            return self._maps
        L = _lib.lib()
        B, C, h, w = feat.shape
        S = feat_src.shape[1]
        V = 1 + S
        feat = ops._f32(feat)
        feat_src = ops._f32(feat_src)
        maps = torch.empty(B, V, h, w, C, dtype=torch.float32, device=feat.device)
        stream = c_void_p(torch.cuda.current_stream(feat.device).cuda_stream)
        for b in range(B):
            _lib.check(L.mvsdf_feat_nchw_to_nhwc(c_void_p(feat[b].data_ptr()), 1, C, h, w,
                                                 c_void_p(maps[b, 0].data_ptr()), stream))
            _lib.check(L.mvsdf_feat_nchw_to_nhwc(c_void_p(feat_src[b].data_ptr()), S, C, h, w,
                                                 c_void_p(maps[b, 1].data_ptr()), stream))
        self._key, self._maps = key, maps
        self._keep = (feat, feat_src)
        return maps


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
                           network_object_mask, object_mask, hit_offsets: Optional[torch.Tensor] = None, reduce_fn=None):
        if uncerts is not None:
            raise NotImplementedError("the uncertainty branch (loss.py:156-159) is dead code in the reference")
        dev = diff_surf_pts.device
        B = feat.shape[0]
        if hit_offsets is None:
            m = (network_object_mask & object_mask).view(B, -1).sum(-1)
            hit_offsets = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), m.cumsum(0)]).to(torch.int32)
        hit_offsets = hit_offsets.to(device=dev, dtype=torch.int32).contiguous()
        args = (diff_surf_pts, hit_offsets, feat.to(dev), cam.to(dev), feat_src.to(dev), src_cams.to(dev), size.to(dev),
                center.to(dev), reduce_fn)
        if torch.is_grad_enabled() and diff_surf_pts.requires_grad:
            from .autograd import FeatConsistency
            return FeatConsistency.apply(self, *args)
        return self._feat_loss_native(*args)

    @torch.no_grad()
    def _feat_loss_native(self, diff_surf_pts, hit_offsets, feat, cam, feat_src, src_cams, size, center, reduce_fn=None):
        L = _lib.lib()
        dev = diff_surf_pts.device
        B = feat.shape[0]
        maps = self.store.get(feat.to(dev), feat_src.to(dev))
        _, V, h, w, C = maps.shape
        cams = torch.cat([cam.to(dev).unsqueeze(1), src_cams.to(dev)], dim=1).to(torch.float32).contiguous()   # [B,V,2,4,4]
        pts = ops._f32(diff_surf_pts)
        if pts.numel() == 0:            # no surface point at all: the kernel reads M = 0 from hit_offsets and touches nothing
            pts = torch.zeros(1, 3, dtype=torch.float32, device=dev)
        size = ops._f32(size.to(dev)).reshape(-1)[:1].contiguous()
        center = ops._f32(center.to(dev)).reshape(-1)[:3].contiguous()
        partial = torch.empty(B, 2, dtype=torch.float64, device=dev)
        out = torch.empty((), dtype=torch.float32, device=dev)
        stream = c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(L.mvsdf_feat_loss_partials(_lib.ptr(pts), _lib.ptr(hit_offsets), _lib.ptr(cams), _lib.ptr(maps), B, V, h, w, C,
                                              _lib.ptr(size), _lib.ptr(center), _lib.ptr(partial), stream))
        if reduce_fn is not None:
            reduce_fn(partial)
        _lib.check(L.mvsdf_feat_loss_finalize(_lib.ptr(partial), B, _lib.ptr(out), stream))
        self.last_partials["feat"] = partial
        self._feat_ctx = (pts, hit_offsets, cams, maps, size, center)      # operands of the native backward
        return out

    @torch.no_grad()
    def _feat_loss_backward_native(self, ctx_tensors, partial, upstream):
        """d loss / d diff_surf_pts through mvsdf_feat_loss_backward (autograd of loss.py:132-155 in the reference)."""
        pts, hit_offsets, cams, maps, size, center = ctx_tensors
        L = _lib.lib()
        dev = pts.device
        B, V, h, w, C = maps.shape
        grad = torch.empty_like(pts)
        up = upstream.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        stream = c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(L.mvsdf_feat_loss_backward(_lib.ptr(pts), _lib.ptr(hit_offsets), _lib.ptr(cams), _lib.ptr(maps), B, V, h, w, C,
                                              _lib.ptr(size), _lib.ptr(center), _lib.ptr(partial), _lib.ptr(up), _lib.ptr(grad),
                                              stream))
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
    def get_eikonal_loss(self, grad_theta):
        if grad_theta.shape[0] == 0:
            return torch.tensor(0.0, device=grad_theta.device)
        return ((grad_theta.norm(2, dim=1) - 1) ** 2).mean()

    def get_surf_loss(self, surf_indicator_output, network_object_mask, object_mask_true):
        n = int((network_object_mask & object_mask_true).sum())
        gt = torch.cat([torch.ones(n), torch.zeros(surf_indicator_output.shape[0] - n)]).to(surf_indicator_output)
        return nn.functional.binary_cross_entropy_with_logits(surf_indicator_output, gt)

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
                model_outputs["diff_surf_pts"], model_outputs.get("uncerts"), ground_truth["feat"], ground_truth["cam"],
                ground_truth["feat_src"], ground_truth["src_cams"], ground_truth["size"][:1], ground_truth["center"][:1],
                nm, om, hit_offsets=model_outputs.get("hit_offsets"), reduce_fn=reduce_fn)
        else:
            res["feat_loss"] = torch.zeros(1, device=dev)
        if model_outputs.get("grad_theta") is not None:
            res["eikonal_loss"] = self.get_eikonal_loss(model_outputs["grad_theta"])
            res["surf_loss"] = self.get_surf_loss(model_outputs["surf_indicator_output"], nm, model_outputs["object_mask_true"])
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
