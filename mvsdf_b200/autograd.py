"""Autograd bridge of the hot path (SURVEY.md section 8(b), row "Autograd"; row f1).

Forward values AND gradients come from libmvsdf_b200.so.  This module only supplies the ``torch.autograd.Function``
wrappers that let ``loss.backward()`` of the reference's training loop (training/idr_train.py:287) reach the parameters:

* ``SdfEval``      -- ImplicitNetwork.forward + .gradient (implicit_differentiable_renderer.py:77-107): outputs
                      ``full [P, 2+F]`` and ``grad [P, 3]``; differentiable w.r.t. the points and the weight-norm
                      parameters, *including* the second-order terms the eikonal / normal paths need (the reference gets
                      them from ``create_graph=True``, :104).  Forward = the fused value+tangent kernel with the layer
                      inputs saved (mvsdf_sdf_forward_train); backward = the tcgen05 reverse sweep + the dW GEMM over the
                      points (mvsdf_sdf_backward, csrc/mlp_bwd_kernel.cuh) + the weight-norm chain (mvsdf_weight_grads).
                      The explicit chain is written out and pinned against autograd in oracle/backward_spec.py.
* ``RenderEval``   -- RenderingNetwork.forward (:145-167), same structure (mvsdf_render_forward_train / _backward).
* ``RgbL1`` / ``FeatConsistency`` / ``DepthL1`` -- IDRLoss.get_rgb_loss / get_feat_loss_corr / get_depth_loss
                      (model/loss.py:21-28, :115-165, :37-63).  RgbL1 and DepthL1 have closed-form backwards,
                      FeatConsistency's backward is the native kernel mvsdf_feat_loss_backward.

No PyTorch / cuBLAS GEMM runs in a training step: tests/test_gpu_autograd.py compares the parameter gradients with autograd
through the oracle restatement of the reference.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


def _param_lists(params):
    """(v0, g0, b0, v1, g1, b1, ...) -> ([v], [g], [b])."""
    return list(params[0::3]), list(params[1::3]), list(params[2::3])


def _interleave(dvs, dgs, dbs, needs):
    out = []
    for i, (dv, dg, db) in enumerate(zip(dvs, dgs, dbs)):
        out += [dv if needs[3 * i] else None,
                (dg.view(-1, 1) if dg is not None and needs[3 * i + 1] else None),
                db if needs[3 * i + 2] else None]
    return out


# ----------------------------------------------------------------------------------------------------------------------
class SdfEval(torch.autograd.Function):
    """(full, grad) = SdfEval.apply(net, shared, x, *params)   with params = (v0, g0, b0, v1, g1, b1, ...).

    ``shared``: None, or a dict shared by several evaluations at numerically IDENTICAL points (IDRNetwork.forward
    evaluates the surface points three times: :202, :325, :326 -- x_diff of sample_network.py equals x_0 in value).  The
    first call runs the kernel and leaves (full, grad, save) in the dict, later calls reuse them: one forward pass and one
    set of saved activations, while every call keeps its own node (and its own backward sweep) in the autograd graph."""

    @staticmethod
    def forward(ctx, net: ops.PackedNet, shared: Optional[dict], x, *params):
        if shared is not None and "save" in shared:
            full, grad, save = shared["full"].clone(), shared["grad"].clone(), shared["save"]
        else:
            full, grad, save = ops.sdf_forward_train(net, x)
            if shared is not None:
                shared.update(full=full, grad=grad, save=save)
        ctx.net = net
        ctx.save = save
        ctx.set_materialize_grads(False)         # an unused output arrives as None instead of a zero tensor
        ctx.save_for_backward(x, *params)
        return full, grad

    @staticmethod
    def backward(ctx, g_full, g_grad):
        x, *params = ctx.saved_tensors
        net = ctx.net
        n = x.shape[0]
        needs = ctx.needs_input_grad[2:]           # [x, v0, g0, b0, ...]
        if n == 0 or (g_full is None and g_grad is None):
            zeros = [torch.zeros_like(p) if nd else None for p, nd in zip(params, needs[1:])]
            return (None, None, torch.zeros_like(x) if needs[0] else None, *zeros)
        dx, dw, db = ops.sdf_backward(net, x, ctx.save, g_full, g_grad, need_dx=bool(needs[0]))
        vs, gs, _ = _param_lists(params)
        dvs, dgs, dbs = ops.weight_grads(net, dw, db, vs, gs)
        return (None, None, dx, *_interleave(dvs, dgs, dbs, needs[1:]))


class SurfaceEval(torch.autograd.Function):
    """(full_s, g_s, x_diff, full_d, n_d) = SurfaceEval.apply(net, x_s, t_s, c_s, d_s, *params) -- everything IDRNetwork.forward
    evaluates at the surface points in one node (fixed cameras):

      full_s, g_s   ImplicitNetwork(x_s) and its gradient at the traced points x_s = c + t d (:202, :275),
      x_diff        the implicit-differentiation point  c + (t - (f(x_s; theta) - f_0) / (g_0 . d)) d  (sample_network.py:10-20),
      full_d, n_d   ImplicitNetwork(x_diff) and its gradient: features and normals of get_rbg_value (:324-338).

    In value x_diff = x_s, so ONE forward pass serves all five outputs.  The backward exploits that the reverse sweep is
    linear in its upstream gradients and that both evaluations share the saved activations: a dx-only sweep of (G_full_d,
    G_n_d) gives dL/d x_diff, which enters f(x_s; theta)'s upstream as  -(dL/d x_diff . d) / (g_0 . d); ONE full sweep
    (with the dW GEMM) of the summed upstreams then yields all parameter gradients -- instead of two full backward passes
    over the surface set (two autograd nodes), which is what the generic SdfEval path does with trained cameras."""

    @staticmethod
    def forward(ctx, net: ops.PackedNet, x_s, t_s, c_s, d_s, *params):
        full, grad, save = ops.sdf_forward_train(net, x_s)
        ctx.net = net
        ctx.save = save
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(x_s, d_s, grad, *params)
        x_diff = c_s + t_s * d_s
        return full, grad, x_diff, full.clone(), grad.clone()

    @staticmethod
    def backward(ctx, g_full_s, g_g_s, g_xdiff, g_full_d, g_n_d):
        x_s, d_s, grad0, *params = ctx.saved_tensors
        net = ctx.net
        needs = ctx.needs_input_grad[5:]
        n = x_s.shape[0]
        if n == 0 or all(g is None for g in (g_full_s, g_g_s, g_xdiff, g_full_d, g_n_d)):
            return (None, None, None, None, None, *[torch.zeros_like(p) if nd else None for p, nd in zip(params, needs)])
        dxt = g_xdiff
        if g_full_d is not None or g_n_d is not None:
            dx_d, _, _ = ops.sdf_backward(net, x_s, ctx.save, g_full_d, g_n_d, need_dx=True, need_dw=False)
            dxt = dx_d if dxt is None else dxt + dx_d
        g_full = None
        for g in (g_full_s, g_full_d):
            if g is not None:
                g_full = g.clone() if g_full is None else g_full + g
        if dxt is not None:
            alpha = -(dxt * d_s).sum(-1) / (grad0 * d_s).sum(-1)
            if g_full is None:
                g_full = torch.zeros(n, net.feature_size + 2, dtype=torch.float32, device=x_s.device)
            g_full[:, 0] += alpha
        g_grad = None
        for g in (g_g_s, g_n_d):
            if g is not None:
                g_grad = g if g_grad is None else g_grad + g
        _, dw, db = ops.sdf_backward(net, x_s, ctx.save, g_full, g_grad, need_dx=False)
        vs, gs, _ = _param_lists(params)
        dvs, dgs, dbs = ops.weight_grads(net, dw, db, vs, gs)
        return (None, None, None, None, None, *_interleave(dvs, dgs, dbs, needs))


class RenderEval(torch.autograd.Function):
    """rgb = RenderEval.apply(net, points, normals, view, feats, *params).  The gradient w.r.t. the view direction is only
    produced when it is asked for (trained camera poses; otherwise the view direction is a constant of the ray)."""

    @staticmethod
    def forward(ctx, net: ops.PackedNet, points, normals, view, feats, *params):
        rgb, save = ops.render_forward_train(net, points, view, normals, feats)
        ctx.net = net
        ctx.save = save
        ctx.save_for_backward(rgb, view, *params)
        return rgb

    @staticmethod
    def backward(ctx, g_rgb):
        rgb, view, *params = ctx.saved_tensors
        net = ctx.net
        needs = ctx.needs_input_grad
        if rgb.shape[0] == 0:
            zin = [torch.zeros_like(t) if nd else None for t, nd in zip((rgb, rgb, view), needs[1:4])]
            return (None, *zin, None, *[torch.zeros_like(p) if nd else None for p, nd in zip(params, needs[5:])])
        d_points, d_normals, d_feats, d_view, dw, db = ops.render_backward(net, ctx.save, rgb, g_rgb, view if needs[3] else None)
        vs, gs, _ = _param_lists(params)
        dvs, dgs, dbs = ops.weight_grads(net, dw, db, vs, gs)
        return (None, d_points if needs[1] else None, d_normals if needs[2] else None, d_view if needs[3] else None,
                d_feats if needs[4] else None, *_interleave(dvs, dgs, dbs, needs[5:]))


# ----------------------------------------------------------------------------------------------------------------------
class RgbL1(torch.autograd.Function):
    """loss = RgbL1.apply(loss_module, rgb_values, rgb_gt, mask_u8_bool, reduce_fn): forward = mvsdf_rgb_l1_* kernels."""

    @staticmethod
    def forward(ctx, module, rgb_values, rgb_gt, mask, reduce_fn):
        out = module._rgb_loss_native(rgb_values, rgb_gt, mask, reduce_fn)
        ctx.save_for_backward(rgb_values, rgb_gt, mask, module.last_partials["rgb"])
        return out

    @staticmethod
    def backward(ctx, g):
        rgb_values, rgb_gt, mask, partial = ctx.saved_tensors          # partial = (sum, n_rays), global after the all-reduce
        d = torch.sign(rgb_values - rgb_gt.reshape(-1, 3)) * mask.unsqueeze(-1).to(rgb_values.dtype)
        return None, d * (g / partial[1].to(rgb_values.dtype)), None, None, None


class DepthL1(torch.autograd.Function):
    """loss = DepthL1.apply(loss_module, eik_points_hom, eik_output, depths, cams, size, center, tp, reduce_fn).
    Forward = mvsdf_depth_loss_partials; the kernel also leaves the per-point target / weight, so the backward is the
    closed form  d loss / d eik_output = weight * sign(eik_output - target) / n_points  (the points are detached, loss.py:38)."""

    @staticmethod
    def forward(ctx, module, pts_hom, eik_output, depths, cams, size, center, tp, reduce_fn):
        out, target, weight = module._depth_loss_native(pts_hom, eik_output, depths, cams, size, center, tp, reduce_fn)
        ctx.save_for_backward(eik_output, target, weight, module.last_partials["depth"])
        return out

    @staticmethod
    def backward(ctx, g):
        eik_output, target, weight, partial = ctx.saved_tensors
        d = torch.sign(eik_output.reshape(-1) - target) * weight * (g / partial[1].to(weight.dtype))
        return None, None, d.view_as(eik_output), None, None, None, None, None, None


class FeatConsistency(torch.autograd.Function):
    """loss = FeatConsistency.apply(loss_module, pts, hit_offsets, maps, map_index, cam, src_cams, size, center, reduce_fn).
    Forward = mvsdf_feat_loss_partials / _finalize; backward = mvsdf_feat_loss_backward (native: projections, bilinear
    tap derivatives and the cosine-similarity chain in one kernel; tests/test_gpu_autograd.py checks it against fp64
    autograd through the oracle)."""

    @staticmethod
    def forward(ctx, module, pts, hit_offsets, maps, map_index, cam, src_cams, size, center, reduce_fn):
        out = module._feat_loss_native(pts, hit_offsets, maps, map_index, cam, src_cams, size, center, reduce_fn)
        ctx.module = module
        ctx.map_index = map_index
        ctx.n_pts = pts.shape[0]
        ctx.pts_shape = pts.shape
        ctx.save_for_backward(module.last_partials["feat"], *module._feat_ctx)
        return out

    @staticmethod
    def backward(ctx, g):
        partial, *operands = ctx.saved_tensors
        if ctx.n_pts == 0:
            return None, g.new_zeros(ctx.pts_shape), None, None, None, None, None, None, None, None
        gp = ctx.module._feat_loss_backward_native(operands, partial, g, ctx.map_index)
        return None, gp, None, None, None, None, None, None, None, None
