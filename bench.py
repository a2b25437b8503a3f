#!/usr/bin/env python
"""bench.py -- rays/sec of the MVSDF hot path (sphere-trace + render + feature loss + rgb L1).

Contract: `python bench.py --gpus N --steps K --warmup W [--impl reference]` prints ONE JSON line
(rank 0).  A "step" is one eval-mode pass of the hot path over one batch of synthetic rays:
IDRNetwork.forward + get_feat_loss_corr + get_rgb_loss.

Workloads (BASELINE.json `configs`):
  cfg2 (default, the configuration the metric is quoted on): DTU-shaped 1200x1600 full image =
       1.92 M rays, 4 source views, 8x512 SDF MLP + 4x512 rendering MLP, eval mode.  With --gpus N the rays of this ONE
       image are dealt to the ranks in round-robin tiles of --shard-tile rays (strong scaling, SURVEY.md section 8e;
       loss partials are all-reduced); --scaling weak renders one image per rank instead.
  cfg3: 2 x 4096 rays, 8 source views, train-mode forward (tp = 0.5).
  cfg1: 32x32 rays, 256-wide nets, 1 source view (the CPU-runnable case).

`--impl reference` times the reference algorithm's CPU implementation (the oracle port in
oracle/mvsdf_oracle.py, pinned against the unmodified reference by tests/golden) on the host cores,
each step a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from mvsdf_b200 import synth  # noqa: E402

METRIC = "rays/sec (sphere-trace+render+feat-loss) at DTU 1200x1600"

WORKLOADS = {
    "cfg2": dict(H=1200, W=1600, width=512, n_src=4, n_images=1, n_rays=None, training=False, shard_rays=True,
                 weights=dict(width=512, seed=0, perturb=0.05, pe_noise=0.003, bias=0.75)),
    "cfg3": dict(H=1200, W=1600, width=512, n_src=8, n_images=2, n_rays=4096, training=True,
                 weights=dict(width=512, seed=0, perturb=0.05, pe_noise=0.003, bias=0.75)),
    # the reference's training step (confs/mvsdf_dtu.conf:4 num_pixels = 4096, training/exp_runner.py:12 batch_size = 8,
    # scene_dataset.py:104 num_src = 2): forward + the five losses + backward + Adam, train_progress = 0.5
    "train32k": dict(H=1200, W=1600, width=512, n_src=2, n_images=8, n_rays=4096, training=True, train_step=True,
                     weights=dict(width=512, seed=0, perturb=0.05, pe_noise=0.003, bias=0.75)),
    "train8k": dict(H=1200, W=1600, width=512, n_src=2, n_images=2, n_rays=4096, training=True, train_step=True,
                    weights=dict(width=512, seed=0, perturb=0.05, pe_noise=0.003, bias=0.75)),
    "cfg1": dict(H=32, W=32, width=256, n_src=1, n_images=1, n_rays=None, training=False,
                 weights=dict(width=256, seed=1, perturb=0.05, pe_noise=0.003, bias=0.6)),
    # BASELINE.json configs[3]: ONE 1.2 M-ray image sharded across the ranks as contiguous ray ranges (strong scaling):
    # every rank holds the whole feature maps, the loss partials are all-reduced
    "cfg4": dict(H=1200, W=1600, width=512, n_src=4, n_images=1, n_rays=1200000, training=False, shard_rays=True,
                 weights=dict(width=512, seed=0, perturb=0.05, pe_noise=0.003, bias=0.75)),
    # configs[4]: Tanks&Temples-shaped 1080p, 12 source views; the 8-GPU throughput sweep shards ONE image like cfg2
    "cfg5": dict(H=1080, W=1920, width=512, n_src=12, n_images=1, n_rays=None, training=False, shard_rays=True,
                 weights=dict(width=512, seed=0, perturb=0.05, pe_noise=0.003, bias=0.75)),
}

# algorithmic FLOPs per SDF evaluation / shaded ray
FLOP = {512: dict(sdf_only=3.671e6, full=3.934e6, render=1.872e6), 256: dict(sdf_only=0.918e6, full=1.050e6, render=0.543e6)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_burst=d.get("bf16_tflops"), bf16_sustained=d.get("bf16_tflops_sustained"), hbm=d.get("hbm_gbs"),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


def ncu_traffic(workload, tau):
    """DRAM bytes per launch of the dominant kernel from the committed ncu pass (profiles/r02/dram_traffic_cfg2_r2.json:
    dram__bytes_read.sum + dram__bytes_write.sum summed over the exact + screening SDF launches / their number).
    It cannot be measured live; null for configurations that were not captured."""
    p = os.path.join(ROOT, "profiles", "r02", "dram_traffic_cfg2_r2.json")
    if workload != "cfg2" or abs(tau - 0.002) > 1e-9 or not os.path.exists(p):
        return None
    return json.load(open(p))["dram_bytes_per_launch"]


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md recipe).  Sampled in-process through NVML
    (nvidia_ml_py, initialised BEFORE the timed region): a fresh `nvidia-smi` process per sample initialises the driver's
    management library every time, which stalled kernel submission for 20-70 ms per sample (`value` 582 ms against 537 ms
    in the un-sampled e2e loop of the same run).  Falls back to one nvidia-smi process per sample if NVML is unavailable.
    One sample per second: an NVML query while hundreds of per-launch CUDA events are outstanding (the roofline's timing)
    costs ~9 ms on some boxes (measured: 572 ms at 4 Hz against 549 / 554 ms with either the sampler or the events off)."""

    def __init__(self, gpu_index: int, period: float = 1.0):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.period = period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self._nvml = self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: map the CUDA ordinal through CUDA_VISIBLE_DEVICES when it lists indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            phys = int(ids[gpu_index]) if (ids and all(v.isdigit() for v in ids) and gpu_index < len(ids)) else gpu_index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._nvml = pynvml
        except Exception:
            self._nvml = self._h = None

    def _sample_nvml(self):
        n = self._nvml
        self.samples.append(float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
        try:
            mask = n.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        except Exception:
            mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        for name, bit in (("hw_slowdown", n.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", n.nvmlClocksThrottleReasonHwThermalSlowdown),
                          ("sw_thermal_slowdown", n.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", n.nvmlClocksThrottleReasonSwPowerCap)):
            if mask & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                             capture_output=True, text=True, timeout=5).stdout.strip().split(",")
        self.samples.append(float(out[0]))
        self.max_mhz = float(out[1])
        for n, v in zip(names, out[2:]):
            if "Active" in v and "Not" not in v:
                self.reasons.add(n)

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "source": "nvml" if self._nvml is not None else "nvidia-smi"}


def make_inputs(cfg, rank, world=1, shard=None, tile=1024):
    """Host tensors of one step.  Weak scaling: every rank renders a different view of the same synthetic scene.
    Strong scaling (shard): all ranks hold the SAME image and rank r takes its share of the rays -- round-robin tiles of
    `tile` consecutive rays (tile 0: one contiguous range) -- with the feature maps replicated."""
    shard = bool(cfg.get("shard_rays")) if shard is None else shard
    seed = int(os.environ.get("MVSDF_BENCH_SEED", 0 if shard else rank))       # env: diagnostic (replay another rank's scene)
    scene = synth.make_scene(cfg["H"], cfg["W"], n_images=cfg["n_images"], n_src=cfg["n_src"], n_rays=cfg["n_rays"], seed=seed)
    if shard and world > 1:
        from mvsdf_b200 import parallel
        scene = parallel.shard_rays(scene, rank, world, tile=tile)
    sd = synth.make_state_dict(**cfg["weights"])
    return scene, sd


def pin(d):
    return {k: (v.pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}


IN_KEYS = ["uv", "pose", "intrinsics", "object_mask"]
GT_KEYS = ["rgb", "feat", "cam", "feat_src", "src_cams", "size", "center"]
# constants of the scene (scene_dataset.py:138-149 computes them once per scene): they live in the loss module's
# channels-last feature store (mvsdf_b200/loss.py FeatureStore) and are NOT part of a step's input
SCENE_KEYS = ["feat", "feat_src"]
STEP_KEYS = [k for k in IN_KEYS + GT_KEYS if k not in SCENE_KEYS]


def run_ours(args):
    import torch.distributed as dist
    from mvsdf_b200 import _lib, parallel
    from mvsdf_b200.loss import B200IDRLoss
    from mvsdf_b200.network import B200IDRNetwork, default_conf

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = WORKLOADS[args.workload]
    strong = bool(cfg.get("shard_rays")) and args.scaling == "strong"
    scene, sd = make_inputs(cfg, rank, world, shard=strong, tile=args.shard_tile)
    B, N = scene["uv"].shape[:2]
    R = B * N
    r_all = torch.tensor([R], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(r_all)
    R_total = int(r_all.item())                                  # rays of the whole job (strong: the one image)
    model = B200IDRNetwork(default_conf(cfg["width"])).to(dev)
    model.load_state_dict(sd)
    model.train(cfg["training"])
    model.skip_min_sdf = bool(args.skip_min_sdf)
    if args.prefilter_tau is not None:
        model.prefilter_tau = float(args.prefilter_tau)
    if args.trace_screen_margin is not None:      # experiment flag (off by default): not path-identical to the reference
        model.trace_screen_margin = float(args.trace_screen_margin)
    loss_mod = B200IDRLoss()
    L = _lib.lib()
    tp = 0.5
    g = torch.Generator().manual_seed(1234 + rank)
    steps01 = torch.rand(100, generator=g)
    eik = (torch.rand(R // 2, 3, generator=g) * 2 - 1)

    def reduce_fn(partial):
        if world > 1:
            parallel.allreduce_partials(partial)                # NCCL all-reduce(SUM) of the loss partials

    host = pin({k: scene[k] for k in IN_KEYS + GT_KEYS})
    resident = {k: host[k].to(dev) for k in IN_KEYS + GT_KEYS}

    @torch.no_grad()          # the metric is the forward path; the training step (backward + Adam) is --workload train32k
    def step(inputs):
        kw = dict(steps01=steps01, eik_points=eik) if cfg["training"] else {}
        out = model({k: inputs[k] for k in IN_KEYS}, tp if cfg["training"] else None, **kw)
        losses = loss_mod.hot_path_losses(out, {k: inputs[k] for k in GT_KEYS}, tp, reduce_fn=reduce_fn)
        return out, losses

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- device-resident timing (`value`)
    for _ in range(args.warmup):
        step(resident)
    barrier()
    # clocks / throttle reasons of the timed region: one nvidia-smi query per second on rank 0's GPU (every query takes a
    # driver-wide lock for tens of ms; with one sampler per rank at 5 Hz the ranks' launch threads stalled each other)
    clocks = ClockSampler(local, args.clock_period) if (rank == 0 and args.clock_period > 0) else None
    if clocks:
        clocks.start()
    launches0 = L.mvsdf_launch_count()
    L.mvsdf_profile_enable(0 if args.no_profile else 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out, losses = step(resident)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    ms_local = ms_total / args.steps
    clk = clocks.stop() if clocks else None
    launches = L.mvsdf_launch_count() - launches0
    import ctypes
    ms_kind = (ctypes.c_float * 8)()          # MVSDF_PROFILE_KINDS
    n_kind = (ctypes.c_int * 8)()
    _lib.check(L.mvsdf_profile_collect(ms_kind, n_kind))
    L.mvsdf_profile_enable(0)
    graphs_used = getattr(model, "graph_replays", 0) > 0
    if graphs_used and sum(n_kind) == 0:
        # small workloads replay a captured CUDA graph: no host code runs per launch, so the per-launch events of the
        # roofline leg are taken in a separate pass with the graph off (same kernels, same inputs; `value` stays the graph run)
        model.use_graphs = False
        step(resident)
        barrier()
        L.mvsdf_profile_enable(1)
        for _ in range(args.steps):
            step(resident)
        barrier()
        _lib.check(L.mvsdf_profile_collect(ms_kind, n_kind))
        L.mvsdf_profile_enable(0)
        model.use_graphs = True
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = R_total / (ms_per_step * 1e-3)

    # tracer evaluations of the last step (E_trace of SURVEY 8d: requests the reference algorithm issues)
    cnt = model.last_trace_counters.cpu()
    evals = int(cnt[:_lib.CTR_SCREENED].sum().item())             # include/mvsdf_b200.h: MVSDF_CTR_*
    screened, refined, violations = int(cnt[_lib.CTR_SCREENED]), int(cnt[_lib.CTR_REFINED]), int(cnt[_lib.CTR_VIOLATIONS])
    n_hit = int(out["hit_offsets"][-1].item())
    width = cfg["width"]
    fl = FLOP[width]
    sdf_only_evals = evals + R                                   # + sdf_output for every ray
    alg_flops_kernel = sdf_only_evals * fl["sdf_only"]            # per step, kernel kind 0
    ms_kernel = (ms_kind[0] + ms_kind[4]) / args.steps          # exact + screening launches of the SDF-only kernel
    peaks = load_peaks()
    # what the kernels actually executed (the prefilter proves most sampler evaluations irrelevant and skips them):
    # exact evaluations = requests outside the 100-sample stages + refined samples + sdf_output; 3 fp16 products each
    sampled = 100 * (int(cnt[_lib.CTR_SAMPLER_RAYS]) + (int(cnt[_lib.CTR_MINSDF_RAYS]) if cfg["training"] and not args.skip_min_sdf else 0))
    # sdf_output: the SDF-only launch covers the rays that are not surface rays; a surface ray's value comes out of the
    # value + normal + feature pass (mvsdf_shade_rays), so it is not an execution of the kernel whose roofline this is
    sdf_out_exec = R - n_hit
    if model.prefilter_tau > 0:
        exact_exec = evals - sampled + refined + sdf_out_exec
        screen_exec = screened
    else:
        exact_exec, screen_exec = evals + sdf_out_exec, 0
    exec_tflop = (3 * exact_exec + screen_exec) * fl["sdf_only"] / 1e12
    # roofline of the DOMINANT kernel = the exact SDF-only kernel (kind 0): algorithmic FLOPs of the evaluations its launches
    # really processed (the units one launch processes, not the evaluations the prefilter proved irrelevant)
    ms_exact = ms_kind[0] / args.steps
    achieved = exact_exec * fl["sdf_only"] / (ms_exact * 1e-3) / 1e12 if ms_exact > 0 else None
    credited = alg_flops_kernel / (ms_kernel * 1e-3) / 1e12 if ms_kernel > 0 else None
    total_alg_flops = evals * fl["sdf_only"] + R * fl["sdf_only"] + n_hit * (3 * fl["full"] + fl["render"])
    if cfg["training"]:
        total_alg_flops += (R // 2) * 4 * fl["full"]

    # ---------------- end-to-end through the public API with host buffers (`e2e`)
    def h2d_step():
        inputs = {k: host[k].to(dev, non_blocking=True) for k in STEP_KEYS}
        for k in SCENE_KEYS:                     # the same host tensors every step: served by the resident feature store
            inputs[k] = host[k]
        _, ls = step(inputs)
        vals = torch.stack([ls["rgb_loss"].reshape(()), ls["feat_loss"].reshape(())]).cpu()   # D2H of the step's result
        return vals

    h2d_bytes = sum(host[k].numel() * host[k].element_size() for k in STEP_KEYS)
    scene_bytes = sum(host[k].numel() * host[k].element_size() for k in SCENE_KEYS)
    restacks0 = None
    d2h_bytes = 8 + 4 * (B + 1)          # the two loss scalars + the hit-count read inside forward()
    for _ in range(max(1, args.warmup // 2)):
        h2d_step()
    barrier()
    restacks0 = loss_mod.store.restacks
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        vals = h2d_step()
    e1.record()
    barrier()
    assert loss_mod.store.restacks == restacks0, "the feature store re-uploaded the scene's maps inside the timed region"
    ms_e2e = e0.elapsed_time(e1)
    t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t.item()) / args.steps
    e2e_value = R_total / (ms_e2e * 1e-3)
    # per-rank load (strong scaling: is the split balanced?)
    per_rank = torch.tensor([ms_local, float(n_hit), float(R), float(evals)], dtype=torch.float64, device=dev)
    if world > 1:
        gathered = [torch.zeros_like(per_rank) for _ in range(world)]
        dist.all_gather(gathered, per_rank)
    else:
        gathered = [per_rank]
    per_rank = [dict(rank=i, ms_per_step=float(g[0]), rays=int(g[2]), hit_fraction=float(g[1] / g[2]),
                     tracer_evals_per_ray=float(g[3] / g[2])) for i, g in enumerate(gathered)]

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base = cpu_reference_sample(cfg, scene, sd, budget_s=args.cpu_budget)
        try:
            # the reference's own thread setting: torch.set_num_threads(1) (training/idr_train.py:21)
            one = cpu_reference_sample(cfg, scene, sd, budget_s=min(10.0, args.cpu_budget), threads=1)
            cpu_base["one_thread"] = {"value": one["value"], "unit": "rays/s", "cores": 1, "sample": one["sample"]}
        except Exception as e:
            cpu_base["one_thread"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
        try:
            cpu_base["torch_cuda_port"] = torch_cuda_port_sample(cfg, scene, sd, dev)
        except Exception as e:                      # a baseline, never a reason to lose the bench line
            cpu_base["torch_cuda_port"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
        try:
            cpu_base["torch_cuda_port_cfg3_train"] = torch_cuda_port_cfg3_train(dev)
        except Exception as e:
            cpu_base["torch_cuda_port_cfg3_train"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None,
            "dtype": "fp32-equivalent (fp16 hi/lo split operands, fp32 tensor-core accumulate)", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {cfg['H']}x{cfg['W']} image, {R} rays/GPU, {cfg['n_src']} src views, "
                                   f"8x{width} SDF MLP + 4x{width} render MLP, {'train' if cfg['training'] else 'eval'}-mode forward "
                                   f"+ feat loss + rgb L1", "rays_per_gpu": R, "hit_fraction": n_hit / R,
                       "tracer_evals_per_ray": evals / R,
                       "rays_total": R_total,
                       "parallelism": ((f"ONE image, rays dealt to {world} ranks in round-robin tiles of {args.shard_tile} rays"
                                        if args.shard_tile > 0 else f"ONE image, contiguous ray ranges over {world} ranks")
                                       if strong else f"one image per rank, dp{world}") + ", loss-partials all-reduce",
                       "per_rank": per_rank,
                       "feature_store": f"{scene_bytes} bytes of scene feature maps resident channels-last (uploaded once, not per step)",
                       "l2_policy": "inputs larger than L2 (>=400 MB of ray state + request lists per step)",
                       "skip_min_sdf": bool(args.skip_min_sdf),
                       "cuda_graph": bool(graphs_used),
                       "trace_screen_margin": model.trace_screen_margin,
                       "prefilter": {"tau": model.prefilter_tau, "screened_evals_per_ray": screened / R, "refined_evals_per_ray": refined / R,
                                     "sampler_rays_fraction": int(cnt[_lib.CTR_SAMPLER_RAYS]) / R,
                                     "minsdf_rays_fraction": int(cnt[_lib.CTR_MINSDF_RAYS]) / R,
                                     "guard_violations": violations, "exact_fallbacks": model.prefilter_fallbacks,
                                     "note": "100-sample stages: screening pass (1 fp16 product, 7 chunks of 2-30 samples, stops behind "
                                             "the first certainly negative sample) + exact pass (3 products) over the undecidable "
                                             "samples; outputs bit-identical to tau=0"}},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "rays/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "mlp_pair2_kernel<NET_SDF, plain, SDF-only head> (the exact launches)",
                         "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                         "frac": (achieved / peaks["bf16_sustained"]) if achieved else None,
                         "traffic": ncu_traffic(args.workload, model.prefilter_tau),
                         "peak_source": peaks["source"] + " bf16_tflops_sustained",
                         "algorithmic_flops_per_launch": exact_exec * fl["sdf_only"] / max(1, n_kind[0] / args.steps),
                         "launches_per_step": n_kind[0] / args.steps, "kernel_ms_per_step": ms_exact,
                         "kernel_share_of_step": ms_exact / ms_per_step,
                         "executed": {"exact_evals_per_ray": exact_exec / R, "screening_evals_per_ray": screen_exec / R,
                                      "fp16_tensor_tflops": exec_tflop / (ms_kernel * 1e-3) if ms_kernel > 0 else None,
                                      "frac_of_peak": (exec_tflop / (ms_kernel * 1e-3) / peaks["bf16_sustained"]) if ms_kernel > 0 else None,
                                      "kernel_ms_per_step_exact_plus_screening": ms_kernel,
                                      "note": "fp16 tensor FLOPs the exact + screening launches really issued (3 products per exact "
                                              "MAC, 1 per screening MAC) over their launch time"},
                         "credited_by_reference_count": {"tflops": credited, "frac": (credited / peaks["bf16_sustained"]) if credited else None,
                                                         "note": "SURVEY 8(d) contract: 2*MAC x the SDF evaluations the REFERENCE algorithm "
                                                                 "requests (E_trace + R, prefilter-independent) over the exact + screening "
                                                                 "launch time -- an algorithm-plus-kernel figure, not a kernel roofline"},
                         "note": "achieved = 2*MAC of the fp32 network x the evaluations the exact kernel's launches processed / their "
                                 "CUDA-event time; every MAC issues 3 fp16 UMMAs (hi*hi + lo*hi + hi*lo), so 1/3 is the ceiling of frac"},
            "step_algorithmic_tflop": total_alg_flops / 1e12,
            "losses": {"rgb": float(vals[0]), "feat": float(vals[1])},
            "mlp_ms_per_step_by_kind": {"sdf_only": ms_kind[0] / args.steps, "sdf_screen": ms_kind[4] / args.steps, "sdf_full": ms_kind[1] / args.steps,
                                        "value_grad": ms_kind[2] / args.steps, "render": ms_kind[3] / args.steps},
        }
        if cpu_base:
            line["cpu_baseline"] = cpu_base
        emit(line)
    if world > 1:
        dist.destroy_process_group()


TRAIN_GT_KEYS = GT_KEYS + ["depths", "depth_cams"]


def run_train(args):
    """`--workload train32k`: one optimisation step of the reference's training loop (idr_train.py:268-300) through the
    drop-in modules -- B200IDRNetwork.forward (training mode, autograd on), B200IDRLoss.forward (rgb + eikonal + surface
    indicator + feature consistency + depth carving), loss.backward() (native reverse sweep + dW GEMM, mlp_bwd_kernel.cuh),
    B200Adam.step with the reference's gradient-norm cap -- on 8 images x 4096 rays, train_progress = 0.5."""
    import ctypes
    from mvsdf_b200 import _lib, conf as schedule
    from mvsdf_b200.loss import B200IDRLoss
    from mvsdf_b200.network import B200IDRNetwork, default_conf
    from mvsdf_b200.optim import B200Adam

    import torch.distributed as dist
    from mvsdf_b200 import parallel
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # data-parallel training ("for end-to-end DP training add a gradient all-reduce"): every rank steps
        # on its OWN 8 images x 4096 rays, the gradients are averaged with one NCCL all-reduce of a flat fp32 bucket, every rank
        # applies the same Adam update.  Weak scaling: the global batch grows with N.
        dist.init_process_group("nccl", device_id=dev)
    cfg = WORKLOADS[args.workload]
    scene, sd = make_inputs(cfg, rank, world, shard=False)
    B, N = scene["uv"].shape[:2]
    R = B * N
    tp = 0.5
    model = B200IDRNetwork(default_conf(cfg["width"])).to(dev)
    model.load_state_dict(sd)
    model.train()
    model.skip_min_sdf = bool(args.skip_min_sdf)
    loss_mod = B200IDRLoss()
    params = [p for p in model.parameters() if p.requires_grad]
    opt = B200Adam(params, lr=2.0e-4 * B * world)                # idr_train.py:110-113: lr scaled by the (global) batch size
    L = _lib.lib()
    g = torch.Generator().manual_seed(1234 + rank)
    steps01 = torch.rand(100, generator=g)
    eik = torch.rand(R // 2, 3, generator=g) * 2 - 1
    host = pin({k: scene[k] for k in IN_KEYS + TRAIN_GT_KEYS})
    resident = {k: host[k].to(dev) for k in IN_KEYS + TRAIN_GT_KEYS}
    cap = schedule.grad_cap(tp) if (schedule.phase[0] <= tp and schedule.enable_grad_cap) else None

    def step(inputs):
        out = model({k: inputs[k] for k in IN_KEYS + ["depths", "depth_cams", "size", "center"]}, tp, steps01=steps01, eik_points=eik)
        ls = loss_mod(out, {k: inputs[k] for k in TRAIN_GT_KEYS}, tp, B)
        opt.zero_grad(set_to_none=True)
        (ls["loss"].sum() / world).backward()
        if world > 1:
            parallel.allreduce_gradients(params)       # SUM of the 1/world-scaled gradients = their mean
        opt.step(max_grad_norm=cap)
        return out, ls

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    for _ in range(args.warmup):
        step(resident)
    barrier()
    clocks = ClockSampler(local, args.clock_period) if (rank == 0 and args.clock_period > 0) else None
    if clocks:
        clocks.start()
    launches0 = L.mvsdf_launch_count()
    L.mvsdf_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out, ls = step(resident)
    e1.record()
    barrier()
    ms_per_step = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    clk = clocks.stop() if clocks else None
    launches = L.mvsdf_launch_count() - launches0
    ms_kind = (ctypes.c_float * 8)()
    n_kind = (ctypes.c_int * 8)()
    _lib.check(L.mvsdf_profile_collect(ms_kind, n_kind))
    L.mvsdf_profile_enable(0)
    n_hit = int(out["hit_offsets"][-1].item())
    cnt = model.last_trace_counters.cpu()
    evals = int(cnt[:_lib.CTR_SCREENED].sum().item())

    # e2e: the step's inputs come from pinned host memory, the loss scalar goes back
    step_keys = [k for k in IN_KEYS + TRAIN_GT_KEYS if k not in SCENE_KEYS]

    def h2d_step():
        inputs = {k: host[k].to(dev, non_blocking=True) for k in step_keys}
        for k in SCENE_KEYS:
            inputs[k] = host[k]
        _, l2 = step(inputs)
        return l2["loss"].detach().reshape(-1)[:1].cpu()

    h2d_bytes = sum(host[k].numel() * host[k].element_size() for k in step_keys)
    h2d_step()
    barrier()
    e0.record()
    for _ in range(args.steps):
        lv = h2d_step()
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return

    width = cfg["width"]
    fl = FLOP[width]
    # backward of the two MLPs, algorithmic: per swept SDF point 4 columns x (dX + dW) = 8 full-head passes; the surface set
    # is swept twice (x_diff and x_s nodes), the eikonal set once; rendering net: 2 passes per hit
    swept = 2 * n_hit + R // 2
    bwd_flops = swept * 8 * fl["full"] + n_hit * 2 * fl["render"]
    ms_bwd = (ms_kind[5] + ms_kind[6]) / args.steps
    peaks = load_peaks()
    achieved = bwd_flops / (ms_bwd * 1e-3) / 1e12 if ms_bwd > 0 else None
    base = None
    if not args.no_cpu_baseline and world == 1:
        try:
            base = torch_cuda_port_train_sample(cfg, scene, sd, dev, tp, steps01, eik)
        except Exception as e:
            base = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    line = {
        "metric": "rays/sec (training step: forward + 5 losses + backward + Adam) at 8 x 4096 rays, DTU-shaped",
        "value": world * R / (ms_per_step * 1e-3), "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32-equivalent (fp16 hi/lo split operands, fp32 tensor-core accumulate), fp32 Adam", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {B} images x {N} rays, {cfg['n_src']} src views, 8x{width} SDF MLP + 4x{width} render "
                               f"MLP, train_progress {tp}: IDRNetwork.forward(train) + IDRLoss.forward (rgb, eikonal, surf, feat, depth) "
                               f"+ backward + Adam(lr 2e-4 x {B}, grad cap {cap})",
                   "rays": R * world, "rays_per_gpu": R,
                   "parallelism": (f"dp{world}: every rank steps on its own {B} images, gradients averaged with one NCCL all-reduce of a flat "
                                   f"fp32 bucket, identical Adam update on every rank") if world > 1 else "single GPU",
                   "hit_fraction": n_hit / R, "tracer_evals_per_ray": evals / R, "skip_min_sdf": bool(args.skip_min_sdf),
                   "prefilter": {"tau": model.prefilter_tau, "exact_fallbacks": model.prefilter_fallbacks,
                                 "note": "a step whose screening guard trips is repeated without the prefilter and tau is doubled "
                                         "(B200IDRNetwork._redo_exact); such steps are inside the timed region"},
                   "swept_points": swept, "l2_policy": "inputs + saved activations (>= 2 GB per step) larger than L2"},
        "clocks": clk,
        "e2e": {"value": world * R / (ms_e2e * 1e-3), "unit": "rays/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": 4 + 4 * 7},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "mlp_bwd_sweep_pair_kernel + mlp_bwd_dw_kernel (reverse sweep and dW GEMM of both MLPs)",
                     "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                     "frac": (achieved / peaks["bf16_sustained"]) if achieved else None, "traffic": None,
                     "peak_source": peaks["source"] + " bf16_tflops_sustained",
                     "kernel_ms_per_step": ms_bwd, "sweep_ms_per_step": ms_kind[5] / args.steps, "dw_ms_per_step": ms_kind[6] / args.steps,
                     "kernel_share_of_step": ms_bwd / ms_per_step,
                     "note": "algorithmic FLOPs = 2*MAC x 8 full-head passes per swept point (4 value/tangent columns x (dX, dW)) + 2 "
                             "render passes per hit; every MAC issues 3 fp16 UMMAs (hi*hi + lo*hi + hi*lo)"},
        "mlp_ms_per_step_by_kind": {"sdf_only": ms_kind[0] / args.steps, "sdf_screen": ms_kind[4] / args.steps,
                                    "value_grad": ms_kind[2] / args.steps, "render": ms_kind[3] / args.steps,
                                    "bwd_sweep": ms_kind[5] / args.steps, "bwd_dw": ms_kind[6] / args.steps},
        "losses": {k: float(v.detach().reshape(-1)[0]) for k, v in ls.items()},
        "grad_norm": float(opt.grad_norm),
    }
    if base:
        line["cpu_baseline"] = {"value": base.get("value"), "unit": "rays/s", "cores": 0, "kind": "port",
                                "sample": base.get("sample", base.get("unavailable")), "torch_cuda_port_train": base}
    emit(line)


def torch_cuda_port_train_sample(cfg, scene, sd, dev, tp, steps01, eik):
    """The reference's training step as eager PyTorch-CUDA ops: the oracle port (unchanged) with its tensors on cuda:0,
    autograd backward (create_graph=True second-order terms), clip_grad_norm_, torch.optim.Adam -- with and without
    minimal_sdf_points."""
    from oracle import mvsdf_oracle as O
    to = lambda d: {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
    sc = to(scene)
    B, N = scene["uv"].shape[:2]
    R = B * N
    res = {}
    for skip in (False, True):
        params = {k: v.clone().to(dev).requires_grad_(True) for k, v in sd.items()}
        opt = torch.optim.Adam(list(params.values()), lr=2.0e-4 * B)

        def one():
            out = O.idr_forward(O.sdf_weights(params), O.render_weights(params), sc, tp, True, steps01=steps01.to(dev),
                                eik_points=eik.to(dev), skip_min_sdf=skip)
            rl = O.hot_path_losses(out, sc, tp)
            total = (0.5 * rl["rgb_loss"] + 0.1 * rl["eikonal_loss"] + 0.01 * rl["surf_loss"] + O.feat_weight(tp) * rl["feat_loss"].sum()
                     + rl["depth_loss"])
            opt.zero_grad(set_to_none=True)
            total.backward()
            torch.nn.utils.clip_grad_norm_(list(params.values()), 0.5)
            opt.step()
            return float(total)

        one()
        one()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        n_it = 3
        for _ in range(n_it):
            one()
        torch.cuda.synchronize(dev)
        dt = (time.perf_counter() - t0) / n_it
        res["skip_min_sdf" if skip else "with_min_sdf"] = {"rays_per_s": R / dt, "ms_per_step": dt * 1e3}
    return {"value": res["with_min_sdf"]["rays_per_s"], "unit": "rays/s", "device": torch.cuda.get_device_name(dev), **res,
            "sample": f"{R} rays ({B} x {N}), full training step (forward + 5 losses + autograd backward + clip + Adam), oracle port on "
                      "cuda:0 (eager PyTorch fp32, TF32 off), 3 timed steps after 2 warm-ups"}


def cpu_reference_sample(cfg, scene, sd, budget_s=20.0, threads=None):
    """The oracle port of the reference algorithm on the host cores, on a bounded sample of the same workload:
    a regular sub-grid of the image's rays (same cameras, weights, feature maps)."""
    from oracle import mvsdf_oracle as O
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    sw, rw = O.sdf_weights(sd), O.render_weights(sd)
    N = scene["uv"].shape[1]
    # calibrate: ~60 evals/ray * 2 * MACs at ~25 GFLOP/s/core-ish -> pick a sample, then rescale once
    n_sample = min(N, 128 if cfg["width"] >= 512 else 1024)
    warm = dict(scene)
    warm["uv"] = scene["uv"][:, :64].contiguous()
    warm["object_mask"] = scene["object_mask"][:, :64].contiguous()
    warm["rgb"] = scene["rgb"][:, :64].contiguous()
    O.idr_forward(sw, rw, warm, None, False)          # thread-pool / allocator warm-up, not timed
    while True:
        stride = max(1, N // n_sample)
        idx = torch.arange(0, N, stride)[:n_sample]
        sub = dict(scene)
        sub["uv"] = scene["uv"][:, idx].contiguous()
        sub["object_mask"] = scene["object_mask"][:, idx].contiguous()
        sub["rgb"] = scene["rgb"][:, idx].contiguous()
        t0 = time.perf_counter()
        out = O.idr_forward(sw, rw, sub, None, False)
        O.hot_path_losses(out, sub, 0.5)
        dt = time.perf_counter() - t0
        n_done = sub["uv"].shape[0] * sub["uv"].shape[1]
        if dt >= 0.4 * budget_s or n_sample >= N:
            break
        n_sample = int(min(N, n_sample * min(8.0, max(1.5, 0.8 * budget_s / max(dt, 1e-3)))))
    return {"value": n_done / dt, "unit": "rays/s", "cores": threads, "kind": "port",
            "sample": f"{n_done} rays (regular sub-grid of the {cfg['H']}x{cfg['W']} image, same weights/cameras/features), "
                      f"eval forward + feat loss + rgb L1 in {dt:.1f} s; oracle/mvsdf_oracle.py (PyTorch CPU restatement "
                      "pinned to the reference by tests/golden)"}


def torch_cuda_port_sample(cfg, scene, sd, dev, n_sample=240000):
    """Part of the baseline leg: the SAME oracle port, unchanged, with its tensors on cuda:0 -- i.e. the reference
    algorithm as eager PyTorch-CUDA ops (fp32 cuBLAS SGEMMs, boolean-mask gathers, host syncs), which is what
    BASELINE.json's north_star calls "the reference PyTorch-CUDA path".  Bounded sub-grid of the same workload."""
    from oracle import mvsdf_oracle as O
    N = scene["uv"].shape[1]
    to = lambda d: {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
    sw, rw = O.sdf_weights(sd).to(dev), O.render_weights(sd).to(dev)

    def sub_scene(n):
        stride = max(1, N // n)
        idx = torch.arange(0, N, stride)[:n]
        sub = dict(scene)
        for k in ("uv", "object_mask", "rgb"):
            sub[k] = scene[k][:, idx].contiguous()
        return to(sub)

    with torch.no_grad():
        warm = sub_scene(4096)
        O.hot_path_losses(O.idr_forward(sw, rw, warm, None, False), warm, 0.5)
        sub = sub_scene(min(N, n_sample))
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        out = O.idr_forward(sw, rw, sub, None, False)
        ls = O.hot_path_losses(out, sub, 0.5)
        float(ls["rgb_loss"])
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
    n_done = sub["uv"].shape[0] * sub["uv"].shape[1]
    return {"value": n_done / dt, "unit": "rays/s", "device": torch.cuda.get_device_name(dev),
            "sample": f"{n_done} rays (regular sub-grid), eval forward + feat loss + rgb L1 in {dt:.2f} s; oracle port on cuda:0 "
                      "(eager PyTorch fp32, TF32 off)"}


def torch_cuda_port_cfg3_train(dev):
    """the reference algorithm as eager PyTorch-CUDA ops (the oracle port on cuda:0) at config 3 --
    2 x 4096 rays, 8 source views, train-mode forward (tp = 0.5) + feat loss + rgb L1 -- with and without
    minimal_sdf_points (ray_tracing.py:280-308, 62 % of the reference's evaluations, read by no MVSDF loss)."""
    from oracle import mvsdf_oracle as O
    cfg = WORKLOADS["cfg3"]
    scene, sd = make_inputs(cfg, 0)
    sc = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in scene.items()}
    sw, rw = O.sdf_weights(sd).to(dev), O.render_weights(sd).to(dev)
    R = scene["uv"].shape[0] * scene["uv"].shape[1]
    g = torch.Generator().manual_seed(1234)
    steps01 = torch.rand(100, generator=g).to(dev)
    eik = (torch.rand(R // 2, 3, generator=g) * 2 - 1).to(dev)
    res = {}
    for skip in (False, True):
        def one():
            out = O.idr_forward(sw, rw, sc, 0.5, True, steps01=steps01, eik_points=eik, skip_min_sdf=skip)
            ls = O.hot_path_losses(out, sc, 0.5)
            return float(ls["rgb_loss"].detach())
        one()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(3):
            one()
        torch.cuda.synchronize(dev)
        dt = (time.perf_counter() - t0) / 3
        res["skip_min_sdf" if skip else "with_min_sdf"] = {"rays_per_s": R / dt, "ms_per_step": dt * 1e3}
    res["sample"] = f"{R} rays (cfg3), train-mode forward + feat loss + rgb L1, oracle port on cuda:0 (eager PyTorch fp32), 3 timed steps"
    return res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = WORKLOADS[args.workload]
    scene, sd = make_inputs(cfg, 0)
    R = scene["uv"].shape[0] * scene["uv"].shape[1]
    per_step_budget = max(5.0, min(30.0, 150.0 / max(1, args.steps + args.warmup)))
    res = None
    t_all = []
    for i in range(args.warmup + args.steps):
        res = cpu_reference_sample(cfg, scene, sd, budget_s=per_step_budget)
        if i >= args.warmup:
            t_all.append(res["value"])
    value = sum(t_all) / len(t_all)
    res["value"] = value
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": R / value * 1e3, "higher_is_better": True,
        "scaling": "strong" if (cfg.get("shard_rays") and args.scaling == "strong") else "weak", "vs_baseline": None,
        "dtype": "fp32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {cfg['H']}x{cfg['W']} image, {cfg['n_src']} src views, 8x{cfg['width']} SDF MLP + "
                               f"4x{cfg['width']} render MLP, eval-mode forward + feat loss + rgb L1; CPU, bounded sample per step",
                   "rays_per_gpu": R},
        "cpu_baseline": res,
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference is pure Python/PyTorch and cannot travel to the GPU box (no copying of reference sources); this arm "
                "times oracle/mvsdf_oracle.py, the restatement pinned to the unmodified reference by tests/golden and "
                "tests/test_oracle.py, on all host cores",
    }
    emit(line)


_REAL_STDOUT = None


def claim_stdout():
    """stdout carries exactly ONE JSON line: anything libraries print there while the bench runs (NCCL writes its version
    banner to stdout, the oracle port prints like the reference does) is sent to stderr instead."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: strong = shard ONE image's rays over the ranks (default), weak = one image per rank")
    ap.add_argument("--shard-tile", type=int, default=1024, help="strong scaling: rays per round-robin tile (0 = contiguous ranges)")
    ap.add_argument("--skip-min-sdf", type=int, default=0)
    ap.add_argument("--prefilter-tau", type=float, default=None, help="override B200IDRNetwork.prefilter_tau (0 = off)")
    ap.add_argument("--trace-screen-margin", type=float, default=None,
                    help="experiment: B200IDRNetwork.trace_screen_margin (default 0 = every sphere-tracing value exact)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--clock-period", type=float, default=1.0, help="seconds between nvidia-smi clock samples on rank 0 (0 = off)")
    ap.add_argument("--no-profile", action="store_true", help="diagnostic: no per-launch CUDA events in the timed region")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    elif WORKLOADS[args.workload].get("train_step"):
        run_train(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
