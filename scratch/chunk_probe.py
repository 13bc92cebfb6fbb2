"""Probe: does running independent row chunks of the imagination rollout on parallel streams
(parallel branches of one CUDA graph) shorten it?  Rows never interact inside _imagine."""
import importlib, sys, os, torch
sys.path[:0] = ['/root/repo', '/root/repo/oracle', '/root/repo/tests']
import synth, parity_cases as pc
pkg = importlib.import_module('dreamerv3-torch_b200')
K = pkg.kernels
dev = 'cuda:0'
config = sys.argv[1] if len(sys.argv) > 1 else 'dmc_proprio'
c = synth.CONFIGS[config]; d = synth.dims_of(config)
N, H = 1024, 15
p = pc.to_dev(synth.rssm_params(d, 0), dev)
pa = pc.to_dev(synth.actor_params(config, 1), dev, grad=True)
start, an, us = synth.imagine_inputs(d, N, H, 0, c['actor_dist'])
sidx = start['stoch'].argmax(-1).to(torch.int32).to(dev); sdet = start['deter'].to(dev)
an, us = an.to(dev), us.to(dev)
spec = K.ActorSpec(c['actor_layers'], c['units'], c['actor_dist'], 0.1, 1.0, 0.01)
rl, al = pc.rssm_list(pkg, p), pc.actor_list(pa, c['actor_layers'], c['actor_dist'])
w_feat = torch.randn(H, N, d.flat + d.deter, device=dev)

def run(chunks, backward):
    n = N // chunks
    outs = []
    cur = torch.cuda.current_stream()
    streams = [torch.cuda.Stream() for _ in range(chunks)] if chunks > 1 else [cur]
    for i, s in enumerate(streams):
        if chunks > 1:
            s.wait_stream(cur)
        with torch.cuda.stream(s):
            r = slice(i * n, (i + 1) * n)
            feat, logit, action, idx = K.imagine(sidx[r].contiguous(), sdet[r].contiguous(), an[:, r].contiguous(),
                                                 us[:, r].contiguous(), None, H, pc.kdims(d), spec, rl, al)
            if backward:
                (feat * w_feat[:, r]).sum().backward()
            outs.append(feat)
    if chunks > 1:
        for s in streams:
            cur.wait_stream(s)
    return outs

for backward in (False, True):
    for chunks in (1, 2, 4):
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            for _ in range(2):
                for q in pa.values(): q.grad = None
                run(chunks, backward)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        for q in pa.values(): q.grad = None
        with torch.cuda.graph(g, stream=side):
            run(chunks, backward)
        for _ in range(3): g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): g.replay()
        e1.record(); torch.cuda.synchronize()
        print(f"{config} backward={backward} chunks={chunks}: {e0.elapsed_time(e1)/20:.3f} ms", flush=True)
        del g
