"""Generate tests/golden/*.npz by running the UNMODIFIED reference (CPU, shims) on seeded
synthetic inputs.  Run in the build container only:  python -m oracle.make_golden

TEST INFRASTRUCTURE ONLY.  The reference has no golden vectors of its own (SURVEY.md
section 4), so these files *are* the pin: outputs of the reference code itself
(/root/reference/code @ a5399816) on inputs that mvsdf_b200/synth.py regenerates
bit-identically from seeds.  Weights are not stored; each file carries the sha256 prefix
of the state_dict it was produced with and tests refuse to run on a mismatch.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from mvsdf_b200 import synth  # noqa: E402
from oracle import ref_shim  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# name -> (width, weight kwargs)
WEIGHT_PRESETS = {
    "w256": dict(width=256, seed=1, perturb=0.05, pe_noise=0.003, bias=0.6),
    "w256_geo": dict(width=256, seed=1, perturb=0.0, pe_noise=0.0, bias=0.6),
    "w512": dict(width=512, seed=0, perturb=0.05, pe_noise=0.003, bias=0.75),
}


def build_reference_model(ref, preset: str):
    kw = WEIGHT_PRESETS[preset]
    sd = synth.make_state_dict(**kw)
    model = ref.idr.IDRNetwork(ref_shim.DictConf(ref_shim.model_conf(kw["width"])))
    missing = model.load_state_dict(sd)
    assert not missing.missing_keys and not missing.unexpected_keys
    return model, sd


def npify(d):
    out = {}
    for k, v in d.items():
        if v is None:
            continue
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    return out


class UniformLog:
    """Records the CPU-generator draws the reference makes inside forward (fact 0.10): Tensor.uniform_ (tracer steps,
    eikonal points), torch.rand_like (depth-surface jitter, :240) and np.random.choice (depth-surface sub-sampling, :245)."""

    def __enter__(self):
        self.draws, self.rand_like, self.choices = [], [], []
        self._orig = torch.Tensor.uniform_
        self._orig_rl = torch.rand_like
        self._orig_ch = np.random.choice

        def logged(t, *a, **k):
            r = self._orig(t, *a, **k)
            self.draws.append(r.clone())
            return r

        def logged_rl(t, *a, **k):
            r = self._orig_rl(t, *a, **k)
            self.rand_like.append(r.clone())
            return r

        def logged_ch(*a, **k):
            r = self._orig_ch(*a, **k)
            self.choices.append(np.sort(np.array(r)))
            return r

        torch.Tensor.uniform_ = logged
        torch.rand_like = logged_rl
        np.random.choice = logged_ch
        return self

    def __exit__(self, *exc):
        torch.Tensor.uniform_ = self._orig
        torch.rand_like = self._orig_rl
        np.random.choice = self._orig_ch


def case_forward(ref, name, preset, H, W, n_images, n_src, n_rays, training, tp, seed, mask_mode="ones"):
    model, sd = build_reference_model(ref, preset)
    scene = synth.make_scene(H, W, n_images=n_images, n_src=n_src, n_rays=n_rays, seed=seed, mask_mode=mask_mode)
    model.train(training)
    inp = {k: scene[k].clone() for k in ["uv", "pose", "intrinsics", "object_mask"]}
    if training:      # ground-truth tensors idr_train.py:262-266 moves into model_input for the phase-0 depth-surface samples
        inp.update({k: scene[k].clone() for k in ["depths", "depth_cams", "center", "size"]})
    torch.manual_seed(4321 + seed)
    np.random.seed(1234 + seed)
    with ref_shim.quiet(), UniformLog() as log:
        out = model(inp, tp)
    loss_mod = ref.loss.IDRLoss()
    nm, om = out["network_object_mask"], out["object_mask"]
    with ref_shim.quiet():
        rgb_l = loss_mod.get_rgb_loss(out["rgb_values"], scene["rgb"], nm, om)
        feat_l = loss_mod.get_feat_loss_corr(out["diff_surf_pts"], None, scene["feat"], scene["cam"],
                                             scene["feat_src"], scene["src_cams"], scene["size"][:1],
                                             scene["center"][:1], nm, om)
    res = {
        "meta_preset": preset, "meta_weights_sha": synth.state_dict_checksum(sd),
        "meta_scene": np.array([H, W, n_images, n_src, -1 if n_rays is None else n_rays, seed]),
        "meta_mask_mode": mask_mode, "meta_training": int(training), "meta_tp": -1.0 if tp is None else tp,
        "points": out["points"], "dists": (out["points"] - scene["pose"][:, :3, 3].unsqueeze(1).repeat(
            1, scene["uv"].shape[1], 1).reshape(-1, 3)).norm(dim=1),
        "network_object_mask": nm, "rgb_values": out["rgb_values"], "sdf_output": out["sdf_output"],
        "diff_surf_pts": out["diff_surf_pts"], "rgb_loss": rgb_l, "feat_loss": feat_l,
    }
    if training:
        with ref_shim.quiet():
            res["eikonal_loss"] = loss_mod.get_eikonal_loss(out["grad_theta"])
            res["surf_loss"] = loss_mod.get_surf_loss(out["surf_indicator_output"], nm, out["object_mask_true"])
        res["grad_theta"] = out["grad_theta"]
        res["eikonal_output"] = out["eikonal_output"]
        res["eikonal_points_hom_all"] = out["eikonal_points_hom"].clone()
        c = ref.conf
        with ref_shim.quiet():      # get_depth_loss rewrites eikonal_points_hom in place (detach() shares storage): pass a copy
            res["depth_loss"] = loss_mod.get_depth_loss(out["eikonal_points_hom"].clone(), out["eikonal_output"], scene["depths"],
                                                        scene["depth_cams"], scene["size"][:1], scene["center"][:1],
                                                        far_thresh=c.far_thresh, far_att=c.far_att(tp), near_thresh=c.near_thresh,
                                                        near_att=c.near_att(tp), smooth=c.smooth(tp))
        res["surf_indicator_output"] = out["surf_indicator_output"]
        draws = log.draws
        # order of draws: [min-sdf steps (100) if that stage ran], eikonal points (n_eik x 3)
        if len(draws) == 2:
            res["steps01"] = draws[0]
            res["eik_points"] = draws[1]
        else:
            assert len(draws) == 1
            res["eik_points"] = draws[0]
        if log.rand_like:            # phase 0: depth-surface samples
            assert len(log.rand_like) == 1 and len(log.choices) == 2
            res["dsurf_jitter01"] = log.rand_like[0]
            res["dsurf_idx_on"] = log.choices[0]
            res["dsurf_idx_jitter"] = log.choices[1]
            res["eikonal_points_hom"] = out["eikonal_points_hom"]
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **npify(res))
    print(f"{name}: hits {int(nm.sum())}/{nm.numel()}  rgb_loss {float(rgb_l):.6f}  feat_loss {float(feat_l):.6f}")


def case_mlp(ref, name, preset, n_pts, seed):
    model, sd = build_reference_model(ref, preset)
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(n_pts, 3, generator=g) * 2 - 1) * 0.9
    model.train()
    full = model.implicit_network(x)
    grad = model.implicit_network.gradient(x.clone())[:, 0, :]
    view = torch.nn.functional.normalize(torch.randn(n_pts, 3, generator=g), dim=1)
    rgb = model.rendering_network(x, grad.detach(), view, full[:, 2:].detach())
    res = {
        "meta_preset": preset, "meta_weights_sha": synth.state_dict_checksum(sd),
        "x": x, "view": view, "sdf_full": full, "grad": grad, "rgb": rgb,
    }
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **npify(res))
    print(f"{name}: sdf range {float(full[:, 0].min()):.3f}..{float(full[:, 0].max()):.3f}")


def case_tracer(ref, name, preset, H, W, n_images, training, seed):
    """RayTracing.forward alone (row a6) incl. the intermediate ray set-up (rows a1, a2)."""
    model, sd = build_reference_model(ref, preset)
    scene = synth.make_scene(H, W, n_images=n_images, n_src=1, seed=seed)
    dirs, cam_loc = ref.rend_util.get_camera_params(scene["uv"], scene["pose"], scene["intrinsics"])
    t_nf, hit = ref.rend_util.get_sphere_intersection(cam_loc, dirs, r=1.0)
    model.ray_tracer.train(training)
    model.implicit_network.eval()
    torch.manual_seed(99 + seed)
    with torch.no_grad(), ref_shim.quiet(), UniformLog() as log:
        pts, nm, dists = model.ray_tracer(sdf=lambda x: model.implicit_network(x)[:, 0], cam_loc=cam_loc,
                                          object_mask=scene["object_mask"].reshape(-1), ray_directions=dirs)
    res = {
        "meta_preset": preset, "meta_weights_sha": synth.state_dict_checksum(sd),
        "meta_scene": np.array([H, W, n_images, 1, -1, seed]), "meta_training": int(training),
        "ray_dirs": dirs, "cam_loc": cam_loc, "t_near_far": t_nf, "hit_sphere": hit,
        "points": pts, "network_object_mask": nm, "dists": dists,
    }
    if log.draws:
        res["steps01"] = log.draws[0]
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **npify(res))
    print(f"{name}: hits {int(nm.sum())}/{nm.numel()}")


def case_featext(ref, name, seed, n, H, W):
    """FeatExt (utils/my_utils.py:693-708) with SEEDED RANDOM weights and BatchNorm statistics under the reference's own
    key names (mvsdf_b200.featext.B200FeatExt(seed) regenerates them; utils/vismvsnet.pt cannot travel to the GPU box).
    The reference class loads that checkpoint in its constructor from a cwd-relative path with CUDA storages: it is
    constructed under cwd = code/ with torch.load mapped to the CPU, then its state is overwritten."""
    from mvsdf_b200.featext import B200FeatExt
    mine = B200FeatExt(seed=seed)
    sd = mine.state_dict()
    cwd = os.getcwd()
    orig_load = torch.load
    try:
        os.chdir(ref_shim.REFERENCE_CODE)
        torch.load = lambda *a, **k: orig_load(*a, **{**k, "map_location": "cpu", "weights_only": False})
        fe = ref.my_utils.FeatExt()
    finally:
        torch.load = orig_load
        os.chdir(cwd)
    res = fe.load_state_dict(sd)
    assert not res.missing_keys and not res.unexpected_keys, res
    fe.eval()
    x = torch.randn(n, 3, H, W, generator=torch.Generator().manual_seed(seed + 100))
    with torch.no_grad():
        o8, o4, o2 = fe(x)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), meta_seed=seed, meta_shape=np.array([n, H, W]),
                        meta_weights_sha=synth.state_dict_checksum({k: v for k, v in sd.items() if v.dtype.is_floating_point}),
                        out_eighth=o8.numpy(), out_quarter=o4.numpy(), out_half=o2.numpy())
    print("wrote", name, [tuple(t.shape) for t in (o8, o4, o2)], "max", float(o2.abs().max()))


def main():
    """python -m oracle.make_golden [case-name ...]   (no names: every case)"""
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    ref = ref_shim.load()
    torch.set_num_threads(8)
    cases = [
        (case_mlp, ("mlp_w256", "w256", 192), dict(seed=11)),
        (case_mlp, ("mlp_w512", "w512", 128), dict(seed=12)),
        (case_tracer, ("tracer_eval_w256", "w256", 32, 32, 1, False), dict(seed=0)),
        (case_tracer, ("tracer_train_w256", "w256", 24, 24, 2, True), dict(seed=1)),
        (case_tracer, ("tracer_eval_w256_geo", "w256_geo", 24, 24, 1, False), dict(seed=2)),
        # BASELINE.json configs[0]: 32x32 rays, 8x256 SDF MLP, 1 source view
        (case_forward, ("cfg1_eval_w256", "w256", 32, 32, 1, 1, None, False, None), dict(seed=0)),
        (case_forward, ("cfg1_train_w256", "w256", 32, 32, 2, 1, 512, True, 0.5), dict(seed=0, mask_mode="disc")),
        (case_forward, ("small_eval_w512", "w512", 20, 20, 1, 2, None, False, None), dict(seed=3)),
        # phase 0 (train_progress < 1/6): depth-surface samples join the eikonal set (:226-251)
        (case_forward, ("train_phase0_w256", "w256", 64, 64, 2, 1, 256, True, 0.1), dict(seed=5, mask_mode="disc")),
        # the shapes of BASELINE.json configs[1] / configs[2] at the headline width, ray counts the CPU reference finishes
        # in seconds: eval with 4 source views; training (tp = 0.5) with 2 images x 8 source views
        (case_forward, ("cfg2_shape_eval_w512", "w512", 28, 28, 1, 4, None, False, None), dict(seed=8)),
        (case_forward, ("cfg3_shape_train_w512", "w512", 48, 48, 2, 8, 192, True, 0.5), dict(seed=9)),
        # row f4: the feature extractor with seeded random weights
        (case_featext, ("featext_seed5",), dict(seed=5, n=2, H=48, W=64)),
    ]
    only = set(sys.argv[1:])
    for fn, args, kw in cases:
        if not only or args[0] in only:
            fn(ref, *args, **kw)


if __name__ == "__main__":
    main()
