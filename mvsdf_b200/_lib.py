"""ctypes binding of libmvsdf_b200.so (the C ABI declared in include/mvsdf_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails the
error is raised -- the product path never routes through PyTorch ops or the oracle.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MVSDF_LIB_PATH") or os.path.join(_HERE, "libmvsdf_b200.so")   # override: kernel experiments only

_lib = None

# fixed slots of mvsdf_trace's out_counters (include/mvsdf_b200.h, MVSDF_CTR_*); slots below CTR_SCREENED sum to E_trace
CTR_SCREENED, CTR_SAMPLER_RAYS, CTR_MINSDF_RAYS, CTR_REFINED, CTR_VIOLATIONS = 251, 252, 253, 254, 255
PROFILE_KINDS = 8


class MvsdfError(RuntimeError):
    pass


def _declare(lib):
    P = c_void_p
    lib.mvsdf_abi_version.restype = c_int
    lib.mvsdf_last_error.restype = c_char_p
    lib.mvsdf_launch_count.restype = ctypes.c_longlong
    lib.mvsdf_launch_count_add.restype = None
    lib.mvsdf_launch_count_add.argtypes = [ctypes.c_longlong]
    lib.mvsdf_profile_enable.restype = None
    lib.mvsdf_profile_enable.argtypes = [c_int]
    lib.mvsdf_profile_collect.restype = c_int
    lib.mvsdf_profile_collect.argtypes = [POINTER(c_float), POINTER(c_int)]
    lib.mvsdf_sdf_net_create.restype = P
    lib.mvsdf_sdf_net_create.argtypes = [c_int, c_int, c_int, c_int, c_int]
    lib.mvsdf_render_net_create.restype = P
    lib.mvsdf_render_net_create.argtypes = [c_int, c_int, c_int, c_int]
    lib.mvsdf_net_destroy.restype = None
    lib.mvsdf_net_destroy.argtypes = [P]
    lib.mvsdf_net_num_layers.restype = c_int
    lib.mvsdf_net_num_layers.argtypes = [P]
    lib.mvsdf_net_packed_bytes.restype = c_size_t
    lib.mvsdf_net_packed_bytes.argtypes = [P]
    lib.mvsdf_net_status_offset.restype = c_size_t
    lib.mvsdf_net_status_offset.argtypes = [P]
    lib.mvsdf_pack_weights.restype = c_int
    lib.mvsdf_pack_weights.argtypes = [P, POINTER(P), POINTER(P), POINTER(P), P, P]
    lib.mvsdf_sdf_forward.restype = c_int
    lib.mvsdf_sdf_forward.argtypes = [P, P, P, c_int64, P, c_int, P, P, P]
    lib.mvsdf_sdf_value_grad.restype = c_int
    lib.mvsdf_sdf_value_grad.argtypes = [P, P, P, c_int64, P, c_int, P, P, P, P]
    lib.mvsdf_render_forward.restype = c_int
    lib.mvsdf_render_forward.argtypes = [P, P, P, P, P, P, c_int64, P, P, P]
    # ---- training step (csrc/train_abi.cu)
    for name in ("mvsdf_train_packed_t_bytes", "mvsdf_train_dw_floats", "mvsdf_train_db_floats"):
        getattr(lib, name).restype = c_size_t
        getattr(lib, name).argtypes = [P]
    for name in ("mvsdf_train_save_bytes", "mvsdf_train_workspace_bytes"):
        getattr(lib, name).restype = c_size_t
        getattr(lib, name).argtypes = [P, c_int64, c_int]
    lib.mvsdf_pack_weights_t.restype = c_int
    lib.mvsdf_pack_weights_t.argtypes = [P, POINTER(P), POINTER(P), P, P]
    lib.mvsdf_sdf_forward_train.restype = c_int
    lib.mvsdf_sdf_forward_train.argtypes = [P, P, P, c_int64, c_size_t, P, P, P, P]
    lib.mvsdf_sdf_backward.restype = c_int
    lib.mvsdf_sdf_backward.argtypes = [P, P, P, c_int64, P, P, P, c_size_t, P, P, P, P, P]
    lib.mvsdf_render_forward_train.restype = c_int
    lib.mvsdf_render_forward_train.argtypes = [P, P, P, P, P, P, c_int64, c_size_t, P, P, P]
    lib.mvsdf_render_backward.restype = c_int
    lib.mvsdf_render_backward.argtypes = [P, P, c_int64, P, P, P, P, c_size_t, P, P, P, P, P, P, P, P]
    lib.mvsdf_weight_grads.restype = c_int
    lib.mvsdf_weight_grads.argtypes = [P, P, P, POINTER(P), POINTER(P), POINTER(P), POINTER(P), POINTER(P), P]
    lib.mvsdf_adam_step.restype = c_int
    lib.mvsdf_adam_step.argtypes = [c_int, POINTER(P), POINTER(P), POINTER(P), POINTER(P), POINTER(c_int64), c_float, c_float, c_float,
                                    c_float, c_int, c_float, P, P, P]
    # ---- FeatExt (csrc/featext.cu)
    lib.mvsdf_featext_num_convs.restype = c_int
    lib.mvsdf_featext_packed_floats.restype = c_size_t
    lib.mvsdf_featext_pack.restype = c_int
    lib.mvsdf_featext_pack.argtypes = [POINTER(P), POINTER(P), POINTER(P), POINTER(P), POINTER(P), c_float, P, P]
    lib.mvsdf_featext_workspace_bytes.restype = c_size_t
    lib.mvsdf_featext_workspace_bytes.argtypes = [c_int, c_int, c_int]
    lib.mvsdf_featext_forward.restype = c_int
    lib.mvsdf_featext_forward.argtypes = [P, P, c_int, c_int, c_int, c_size_t, P, P, P, P, P]
    class TracerParams(ctypes.Structure):
        _fields_ = [("object_bounding_sphere", c_float), ("sdf_threshold", c_float), ("line_search_step", c_float),
                    ("dist_clip", c_float), ("line_step_iters", c_int), ("sphere_tracing_iters", c_int),
                    ("n_steps", c_int), ("n_secant_steps", c_int), ("skip_min_sdf", c_int),
                    ("prefilter_tau", c_float), ("trace_screen_margin", c_float)]
    lib.TracerParams = TracerParams
    lib.mvsdf_trace_workspace_bytes.restype = c_size_t
    lib.mvsdf_trace_workspace_bytes.argtypes = [c_int64, c_int]
    lib.mvsdf_trace.restype = c_int
    lib.mvsdf_trace.argtypes = [P, P, P, P, P, P, POINTER(TracerParams), c_int, c_int, c_int, P, P, c_size_t, P,
                                P, P, P, P, P, P, P]
    lib.mvsdf_shade_workspace_bytes.restype = c_size_t
    lib.mvsdf_shade_workspace_bytes.argtypes = [c_int64, c_int]
    lib.mvsdf_shade_rays.restype = c_int
    lib.mvsdf_shade_rays.argtypes = [P, P, P, P, P, P, P, c_int, c_int, c_int, c_size_t, P, P, P, P, P, P, P, P, P]
    lib.mvsdf_depth_backproject.restype = c_int
    lib.mvsdf_depth_backproject.argtypes = [P, P, P, c_int, c_int, c_int, P, P, P, P, P]
    lib.mvsdf_feat_nchw_to_nhwc.restype = c_int
    lib.mvsdf_feat_nchw_to_nhwc.argtypes = [P, c_int, c_int, c_int, c_int, P, P]
    lib.mvsdf_feat_loss_partials.restype = c_int
    lib.mvsdf_feat_loss_partials.argtypes = [P, P, P, P, c_int, c_int, c_int, c_int, c_int, P, P, P, P]
    lib.mvsdf_feat_loss_backward.restype = c_int
    lib.mvsdf_feat_loss_backward.argtypes = [P, P, P, P, c_int, c_int, c_int, c_int, c_int, P, P, P, P, P, P]
    lib.mvsdf_feat_loss_partials_indexed.restype = c_int
    lib.mvsdf_feat_loss_partials_indexed.argtypes = [P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, P, P, P, P]
    lib.mvsdf_feat_loss_backward_indexed.restype = c_int
    lib.mvsdf_feat_loss_backward_indexed.argtypes = [P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, P, P, P, P, P, P]
    lib.mvsdf_feat_loss_finalize.restype = c_int
    lib.mvsdf_feat_loss_finalize.argtypes = [P, c_int, P, P]
    lib.mvsdf_depth_loss_partials.restype = c_int
    lib.mvsdf_depth_loss_partials.argtypes = [P, c_int, P, c_int64, P, P, c_int, c_int, c_int, P, P, c_float, c_float, c_float,
                                              c_float, c_float, P, P, P, P]
    lib.mvsdf_rgb_l1_partials.restype = c_int
    lib.mvsdf_rgb_l1_partials.argtypes = [P, P, P, c_int64, P, P]
    lib.mvsdf_rgb_l1_finalize.restype = c_int
    lib.mvsdf_rgb_l1_finalize.argtypes = [P, P, P]
    return lib


def lib():
    """Loads the shared library (built in-tree by ``__graft_entry__.build()``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MvsdfError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). mvsdf_b200 has no CPU / PyTorch fallback.")
        _lib = _declare(ctypes.CDLL(LIB_PATH))
    return _lib


def check(rc: int):
    if rc != 0:
        msg = lib().mvsdf_last_error()
        raise MvsdfError(f"mvsdf_b200 call failed ({rc}): {msg.decode() if msg else '?'}")


def ptr(t):
    """Device (or host) address of a torch tensor / None as c_void_p."""
    if t is None:
        return c_void_p(0)
    assert t.is_contiguous(), "mvsdf_b200 needs contiguous buffers"
    return c_void_p(t.data_ptr())


def ptr_array(tensors):
    arr = (c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr
