"""Schedules and flags read by the hot path -- a restatement of the reference's module-level
configuration (code/model/conf.py:1-33).  A different schedule module can be passed to
B200IDRNetwork / B200IDRLoss, mirroring the reference's IDR_USE_ENV=1 + IDR_CONF=<module> override
(implicit_differentiable_renderer.py:15-17)."""

feat_img_scale = 2
phase = (1 / 6, 1 / 2)


def _phase0_only(tp):
    return tp < phase[0]


def d_use_rt_surf(tp):
    return True


def d_use_eik(tp):
    return True


d_use_dsurf_on = _phase0_only
d_use_dsurf_jitter = _phase0_only


def eik_use_rt_surf(tp):
    return True


def eik_use_eik(tp):
    return True


eik_use_dsurf_on = _phase0_only
eik_use_dsurf_jitter = _phase0_only

disable_rgb_grad = False
use_invalid = False
use_mask = False
out_thresh_perc = 1 / 8
enable_feat = True
enable_rgb = True
far_thresh = 0.25
near_thresh = 0.1
surf_weight = 0.01
eikonal_weight = 0.1
enable_grad_cap = True


def far_att(tp):
    return 1


def near_att(tp):
    return 1 if tp < phase[0] else (0.1 if tp < phase[1] else 0.01)


def smooth(tp):
    return None


def rgb_weight(tp):
    return 0.5


def feat_weight(tp):
    return 0 if tp < phase[0] else (0.1 if tp < phase[1] else 0.01)


def depth_weight(tp):
    return 1


def grad_cap(tp):
    return 2 if tp < phase[1] else 0.5
