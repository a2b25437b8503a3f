"""world_size-2 gloo test of the multi-GPU host logic (ray sharding + the single all-reduce of loss
partials + finalisation).  The per-rank partials are produced by the CPU oracle here (no GPU in this
container); on the GPU box the same reduction is fed by mvsdf_feat_loss_partials / mvsdf_rgb_l1_partials
(tests/test_gpu_pipeline.py::test_shard_invariance_of_loss_partials)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mvsdf_b200 import parallel, synth
    from oracle import mvsdf_oracle as O
    sd = synth.make_state_dict(width=64, seed=5, perturb=0.05, pe_noise=0.003, bias=0.6)
    sw, rw = O.sdf_weights(sd), O.render_weights(sd)
    scene = synth.make_scene(20, 20, n_images=2, n_src=2, seed=6)
    part = parallel.shard_rays(scene, rank, world)
    with torch.no_grad():
        out = O.idr_forward(sw, rw, part, None, False)
    nm, om = out["network_object_mask"], out["object_mask"]
    m = nm & om
    rgb_p = torch.tensor([float((out["rgb_values"][m] - part["rgb"].reshape(-1, 3)[m]).abs().sum()), float(m.numel())],
                         dtype=torch.float64)
    _, parts = O.feat_loss_corr(out["diff_surf_pts"], part["feat"], part["cam"], part["feat_src"], part["src_cams"],
                                part["size"][:1], part["center"][:1], nm, om, return_parts=True)
    feat_p = torch.tensor(parts if parts else [(0.0, 0)] * 2, dtype=torch.float64)
    parallel.allreduce_partials(rgb_p, feat_p)
    if rank == 0:
        ret["rgb"] = float(parallel.finalize_rgb(rgb_p))
        ret["feat"] = float(parallel.finalize_feat(feat_p))
    dist.destroy_process_group()


def test_two_rank_loss_reduction_matches_unsharded():
    sys.path.insert(0, ROOT)
    from mvsdf_b200 import parallel, synth
    from oracle import mvsdf_oracle as O
    # shard bounds cover the range exactly once
    for n, w in [(10, 3), (4096, 8), (7, 8)]:
        covered = []
        for r in range(w):
            b, e = parallel.shard_bounds(n, r, w)
            covered += list(range(b, e))
        assert covered == list(range(n))
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    sd = synth.make_state_dict(width=64, seed=5, perturb=0.05, pe_noise=0.003, bias=0.6)
    scene = synth.make_scene(20, 20, n_images=2, n_src=2, seed=6)
    with torch.no_grad():
        out = O.idr_forward(O.sdf_weights(sd), O.render_weights(sd), scene, None, False)
    ref = O.hot_path_losses(out, scene, 0.5)
    assert abs(ret["rgb"] - float(ref["rgb_loss"])) < 1e-6
    assert abs(ret["feat"] - float(ref["feat_loss"])) < 1e-6


def _train_loss(out, rgb_gt, loss_mod, reduce_fn, n_total):
    """rgb (global denominator) + eikonal + surface-indicator terms of IDRLoss.forward with the reference's weights
    (loss.py:176-219); the eikonal / surface means go through B200IDRLoss's own partial reducers."""
    from mvsdf_b200 import conf
    m = out["network_object_mask"] & out["object_mask"]
    rgb = (out["rgb_values"][m] - rgb_gt.reshape(-1, 3)[m]).abs().sum() / n_total
    eik = loss_mod.get_eikonal_loss(out["grad_theta"], reduce_fn=reduce_fn)
    surf = loss_mod.get_surf_loss(out["surf_indicator_output"], out["network_object_mask"], out["object_mask_true"],
                                  reduce_fn=reduce_fn)
    return conf.rgb_weight(0.5) * rgb + conf.eikonal_weight * eik + conf.surf_weight * surf, eik, surf


def _grad_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mvsdf_b200 import parallel, synth
    from mvsdf_b200.loss import B200IDRLoss
    from oracle import mvsdf_oracle as O
    sd = synth.make_state_dict(width=64, seed=5, perturb=0.05, pe_noise=0.003, bias=0.6)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    scene = synth.make_scene(20, 20, n_images=2, n_src=2, seed=6)
    tile = 16                                             # round-robin ray tiles, like bench.py's strong-scaling split
    part = parallel.shard_rays(scene, rank, world, tile=tile)
    n_total = scene["uv"].shape[0] * scene["uv"].shape[1]
    eik_all = torch.rand(n_total // 2, 3, generator=torch.Generator().manual_seed(11)) * 2 - 1
    # eikonal samples split over the ranks in proportion to their rays: rank r draws R_r / 2 of them
    n_loc = [sum(int(parallel.shard_index(scene["uv"].shape[1], r, world, tile).numel()) for _ in range(scene["uv"].shape[0])) // 2
             for r in range(world)]
    b = sum(n_loc[:rank])
    e = b + n_loc[rank]
    out = O.idr_forward(O.sdf_weights(params), O.render_weights(params), part, 0.5, True, eik_points=eik_all[b:e],
                        skip_min_sdf=True)
    loss, eik, surf = _train_loss(out, part["rgb"], B200IDRLoss(), parallel.allreduce_partials, n_total)
    loss.backward()
    plist = list(params.values())
    parallel.allreduce_gradients(plist)
    if rank == 0:
        ret["grads"] = {k: (p.grad.clone() if p.grad is not None else None) for k, p in params.items()}
        ret["eik"], ret["surf"] = float(eik), float(surf)
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_matches_unsharded_backward():
    """Sharded training step (rgb + eikonal + surface-indicator terms, rays dealt in round-robin tiles, eikonal samples
    split) followed by allreduce_gradients == the unsharded backward: every term must be formed with GLOBAL denominators
    (local .mean()s of the eikonal / surface terms would be over-weighted by world_size)."""
    sys.path.insert(0, ROOT)
    from mvsdf_b200 import parallel, synth
    from mvsdf_b200.loss import B200IDRLoss
    from oracle import mvsdf_oracle as O
    # round-robin tiles cover every ray exactly once
    for n, w, tile in [(400, 2, 16), (4096, 8, 256), (1000, 3, 7)]:
        covered = sorted(int(i) for r in range(w) for i in parallel.shard_index(n, r, w, tile))
        assert covered == list(range(n))
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_grad_worker, args=(2, port, ret), nprocs=2, join=True)
    sd = synth.make_state_dict(width=64, seed=5, perturb=0.05, pe_noise=0.003, bias=0.6)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    scene = synth.make_scene(20, 20, n_images=2, n_src=2, seed=6)
    n_total = scene["uv"].shape[0] * scene["uv"].shape[1]
    eik_all = torch.rand(n_total // 2, 3, generator=torch.Generator().manual_seed(11)) * 2 - 1
    out = O.idr_forward(O.sdf_weights(params), O.render_weights(params), scene, 0.5, True, eik_points=eik_all,
                        skip_min_sdf=True)
    loss, eik, surf = _train_loss(out, scene["rgb"], B200IDRLoss(), None, n_total)
    loss.backward()
    assert abs(ret["eik"] - float(eik)) < 1e-6 * max(1.0, abs(float(eik)))
    assert abs(ret["surf"] - float(surf)) < 1e-6
    checked = 0
    for k, p in params.items():
        g = ret["grads"][k]
        if p.grad is None:
            assert g is None or float(g.abs().max()) == 0.0
            continue
        assert torch.allclose(g, p.grad, rtol=2e-4, atol=1e-7), (k, (g - p.grad).abs().max())
        checked += 1
    assert checked >= 20
