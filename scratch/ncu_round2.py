"""Small workload for `ncu --set full` of the round-2 kernels: the raw-A (TMEM-fed) GEMM on two
in-loop shapes and the persistent imagination forward (DV3_IMAGINE_PERSISTENT=1)."""
import importlib, os, sys, torch
os.environ["DV3_IMAGINE_PERSISTENT"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
pkg = importlib.import_module('dreamerv3-torch_b200')
import parity_cases as pc, synth
K = pkg.kernels; dev = 'cuda:0'
torch.manual_seed(0)
for (M, N, K1, K2) in [(1024, 512, 512, 0), (1024, 1536, 512, 512)]:
    a = torch.randn(M, K1, device=dev); a2 = torch.randn(M, K2, device=dev) if K2 else None
    w = K.split(torch.randn(N, K1 + K2, device=dev) / (K1 + K2) ** 0.5)
    for _ in range(3): K.gemm_tc_rawa(a, w, a2=a2)
torch.cuda.synchronize()
config, N, H = "dmc_proprio", 1024, 15
c = synth.CONFIGS[config]; d = synth.dims_of(config)
p = synth.rssm_params(d, 0); pa = synth.actor_params(config, 1)
start, act_noise, u_state = synth.imagine_inputs(d, N, H, 0, c["actor_dist"])
pd, pad = pc.to_dev(p, dev), pc.to_dev(pa, dev)
spec = K.ActorSpec(c["actor_layers"], c["units"], c["actor_dist"], 0.1, 1.0, 0.01)
with torch.no_grad():
    for _ in range(3):
        K.imagine(start["stoch"].argmax(-1).to(torch.int32).to(dev), start["deter"].to(dev), act_noise.to(dev),
                  u_state.to(dev), None, H, pc.kdims(d), spec, pc.rssm_list(pkg, pd),
                  pc.actor_list(pad, c["actor_layers"], c["actor_dist"]), start_logit=start["logit"].to(dev))
torch.cuda.synchronize()
print("ok")
