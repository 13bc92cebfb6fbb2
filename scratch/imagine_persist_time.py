"""Times dv3_imagine_fwd (persistent vs stepwise) as CUDA-graph replays (no host in the timed region)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import importlib

pkg = importlib.import_module("dreamerv3-torch_b200")
import parity_cases as pc
import synth

dev = "cuda:0"
lib = pkg._lib.lib()


def timed(config, N, H, stepwise, reps=20):
    os.environ["DV3_IMAGINE_PERSISTENT"] = "0" if stepwise else "1"
    lib.dv3_reload_env()
    c = synth.CONFIGS[config]
    d = synth.dims_of(config)
    dist, layers = c["actor_dist"], c["actor_layers"]
    p = synth.rssm_params(d, 0)
    pa = synth.actor_params(config, 1)
    start, act_noise, u_state = synth.imagine_inputs(d, N, H, 0, dist)
    pd = pc.to_dev(p, dev)
    pad = pc.to_dev(pa, dev)
    spec = pkg.kernels.ActorSpec(layers, c["units"], dist, 0.1, 1.0, 0.01)
    args = (start["stoch"].argmax(-1).to(torch.int32).to(dev), start["deter"].to(dev),
            act_noise.to(dev), u_state.to(dev), None, H, pc.kdims(d), spec, pc.rssm_list(pkg, pd),
            pc.actor_list(pad, layers, dist))
    sl = start["logit"].to(dev)
    with torch.no_grad():
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                pkg.kernels.imagine(*args, start_logit=sl)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            out = pkg.kernels.imagine(*args, start_logit=sl)
        for _ in range(3):
            gr.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for config in sys.argv[1:] or ["dmc_proprio", "atari100k"]:
    a = timed(config, 1024, 15, True)
    b = timed(config, 1024, 15, False)
    print(f"{config}: imagine fwd as a graph replay: stepwise {a:.3f} ms, persistent {b:.3f} ms")
