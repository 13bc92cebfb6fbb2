"""A/B of the persistent imagination forward against the stepwise launches (same library, env knob),
plus timing of both.  usage: python scratch/imagine_persist_check.py [config ...]"""
import ctypes as C
import os
import sys
import time

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


def run(config, N, H, seed, stepwise, backward=True):
    os.environ["DV3_IMAGINE_PERSISTENT"] = "0" if stepwise else "1"
    lib.dv3_reload_env()
    c = synth.CONFIGS[config]
    d = synth.dims_of(config)
    dist, layers = c["actor_dist"], c["actor_layers"]
    p = synth.rssm_params(d, seed)
    pa = synth.actor_params(config, seed + 1)
    start, act_noise, u_state = synth.imagine_inputs(d, N, H, seed, dist)
    pd = pc.to_dev(p, dev)
    pad = pc.to_dev(pa, dev, grad=True)
    spec = pkg.kernels.ActorSpec(layers, c["units"], dist, 0.1, 1.0, 0.01)
    args = (start["stoch"].argmax(-1).to(torch.int32).to(dev), start["deter"].to(dev),
            act_noise.to(dev), u_state.to(dev), None, H, pc.kdims(d), spec, pc.rssm_list(pkg, pd),
            pc.actor_list(pad, layers, dist))
    feat, logit, action, idx = pkg.kernels.imagine(*args, start_logit=start["logit"].to(dev))
    out = dict(feat=feat.detach().clone(), logit=logit.detach().clone(), action=action.detach().clone(),
               idx=idx.clone())
    if backward:
        g = torch.Generator().manual_seed(seed + 9)
        SC = d.flat
        w_feat = torch.randn(H, N, SC + d.deter, generator=g).to(dev)
        w_log = (0.1 * torch.randn(H, N, d.stoch, d.classes, generator=g)).to(dev)
        w_act = (0.1 * torch.randn(H, N, d.actions, generator=g)).to(dev)
        ((feat * w_feat).sum() + (logit * w_log).sum() + (action * w_act).sum()).backward()
        for k in pa:
            out["d." + k] = pad[k].grad.detach().clone()
    # timing of the forward alone
    torch.cuda.synchronize()
    with torch.no_grad():
        for _ in range(3):
            pkg.kernels.imagine(*args, start_logit=start["logit"].to(dev))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            pkg.kernels.imagine(*args, start_logit=start["logit"].to(dev))
        e1.record()
        torch.cuda.synchronize()
    out["ms"] = e0.elapsed_time(e1) / 20
    return out


def main():
    configs = sys.argv[1:] or ["atari100k", "dmc_proprio"]
    ok = True
    for config in configs:
        for (N, H, seed) in [(1024, 15, 0), (256, 4, 3)]:
            a = run(config, N, H, seed, stepwise=True)
            b = run(config, N, H, seed, stepwise=False)
            print(f"== {config} N={N} H={H}: stepwise {a['ms']:.3f} ms, persistent {b['ms']:.3f} ms (eager, incl. host)")
            mism = int((a["idx"] != b["idx"]).sum())
            print(f"   idx mismatches {mism} of {a['idx'].numel()}")
            ok = ok and mism == 0
            for k in a:
                if k in ("idx", "ms"):
                    continue
                den = a[k].abs().max().item() + 1e-30
                err = (a[k] - b[k]).abs().max().item() / den
                bad = not (err < 1e-4)
                ok = ok and not bad
                print(f"   {k:24s} max|diff|/max|ref| = {err:.3e}{'   <-- BAD' if bad else ''}")
    if os.environ.get("DV3_IMAGINE_TIMING") == "1":
        H = 15
        buf = (C.c_ulonglong * (H * 16))()
        if lib.dv3_debug_imagine_timing(buf, H) == 0:
            names = ["start", "trunk", "wait_top", "head", "img_in", "gru", "out", "ims"]
            for k in (1, 7, 12):
                st = [buf[k * 16 + i] for i in range(8)]
                print(f"   step {k} phase ns:", {names[i + 1]: st[i + 1] - st[i] for i in range(7)},
                      "total", st[7] - st[0])
    print("RESULT", "OK" if ok else "MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
