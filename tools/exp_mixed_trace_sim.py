"""CPU experiment for VERDICT r01 item 6 ("run sphere-tracing iterations whose |f| >> tau at screening precision"):
what would it do to parity?  The oracle's sphere tracer is run twice on the golden scenes -- once as is, once with the
SDF values of the sphere-tracing stage perturbed the way the screening kernel perturbs them (|err| <= 1e-3, measured max
9.6e-4, tests/test_gpu_prefilter.py) whenever the perturbed |f| exceeds a margin (the exact value is used inside the
margin, as the proposal says).  Sampler, secant and the final sdf_output stay exact.  Printed: hit-mask flips and the
distribution of |dists_mixed - dists_exact| over the rays both runs hit, against the 1e-4 depth gate.

Run here (CPU, needs no reference):  python tools/exp_mixed_trace_sim.py  > profiles/r02/exp_mixed_trace_sim.log"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mvsdf_oracle as O  # noqa: E402
from tests.helpers import preset_state_dict, scene_from_meta  # noqa: E402


def run(name, err, margin, seed=0):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"), allow_pickle=True)
    sd = preset_state_dict(str(g["meta_preset"]), g["meta_weights_sha"])
    sw = O.sdf_weights(sd)
    scene = scene_from_meta(g)
    dirs, cam = O.camera_rays(scene["uv"], scene["pose"], scene["intrinsics"])
    exact = lambda x: O.sdf_mlp(x, sw)[:, 0]
    gen = torch.Generator().manual_seed(seed)
    stage = {"sphere": False, "screened": 0, "total": 0}

    def mixed(x):
        f = exact(x)
        if not stage["sphere"]:
            return f
        fs = f + (torch.rand(f.shape, generator=gen) * 2 - 1) * err
        use = fs.abs() > margin
        stage["screened"] += int(use.sum())
        stage["total"] += f.numel()
        return torch.where(use, fs, f)

    orig = O._sphere_trace

    def patched(sdf, *a, **k):
        stage["sphere"] = True
        try:
            return orig(sdf, *a, **k)
        finally:
            stage["sphere"] = False

    om = scene["object_mask"].reshape(-1)
    with torch.no_grad():
        p0, m0, d0 = O.trace_rays(exact, cam, om, dirs)
        O._sphere_trace = patched
        try:
            p1, m1, d1 = O.trace_rays(mixed, cam, om, dirs)
        finally:
            O._sphere_trace = orig
    both = m0 & m1
    dd = (d1 - d0).abs()[both]
    q = lambda v: float(torch.quantile(dd, v)) if dd.numel() else 0.0
    print(f"{name:24s} err {err:.0e} margin {margin:.0e}: rays {m0.numel()}, hits {int(m0.sum())}, mask flips {int((m0 != m1).sum())}, "
          f"screened {stage['screened']}/{stage['total']} sphere-tracing evals; |d dists| median {q(0.5):.2e} p90 {q(0.9):.2e} "
          f"p99 {q(0.99):.2e} max {float(dd.max()) if dd.numel() else 0:.2e}; over 1e-4: {int((dd > 1e-4).sum())} "
          f"({100.0 * float((dd > 1e-4).float().mean()) if dd.numel() else 0:.2f} %), over 1e-5: {100.0 * float((dd > 1e-5).float().mean()) if dd.numel() else 0:.1f} %")


if __name__ == "__main__":
    torch.set_num_threads(8)
    for name in ["tracer_eval_w256", "tracer_eval_w256_geo", "cfg2_shape_eval_w512", "small_eval_w512"]:
        for err, margin in [(1e-3, 4e-3), (1e-3, 2e-2), (2.5e-4, 4e-3)]:
            run(name, err, margin)
