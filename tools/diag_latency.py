"""Small-batch (cfg3-sized: 2 x 4096 rays, training) latency split: host enqueue time vs device time of mvsdf_trace and
of the whole forward (diagnostic)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvsdf_b200 import synth
from mvsdf_b200.network import B200IDRNetwork, default_conf

dev = torch.device("cuda:0")
sd = synth.make_state_dict(width=512, seed=0, perturb=0.05, pe_noise=0.003, bias=0.75)
scene = synth.make_scene(1200, 1600, n_images=2, n_src=2, n_rays=int(os.environ.get("RAYS", "4096")), seed=0)
model = B200IDRNetwork(default_conf(512)).to(dev)
model.load_state_dict(sd)
training = os.environ.get("TRAIN", "1") == "1"
model.train(training)
uv, pose, K = scene["uv"].to(dev), scene["pose"].to(dev), scene["intrinsics"].to(dev)
obj = torch.ones(uv.shape[0] * uv.shape[1], dtype=torch.uint8, device=dev)
steps = torch.rand(100)
sdf_net = model.implicit_network.packed()
inp = {k: scene[k].to(dev) for k in ["uv", "pose", "intrinsics", "object_mask"]}
eik = torch.rand(uv.shape[0] * uv.shape[1] // 2, 3) * 2 - 1
for name, fn in [("trace", lambda: model.trace(sdf_net, uv, pose, K, obj, training, steps)),
                 ("pack", lambda: model.implicit_network.packed()),
                 ("forward", lambda: model(inp, 0.5, steps01=steps, eik_points=eik) if training else model(inp))]:
    with torch.no_grad():
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        n = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        host = 0.0
        t_all = time.perf_counter()
        e0.record()
        for _ in range(n):
            t0 = time.perf_counter()
            fn()
            host += time.perf_counter() - t0
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t_all
    print(f"{name}: host enqueue {host / n * 1e3:.2f} ms/call, device span {e0.elapsed_time(e1) / n:.2f} ms/call, wall {wall / n * 1e3:.2f} ms/call")
