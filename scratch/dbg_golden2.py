import importlib, sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
pkg = importlib.import_module('dreamerv3-torch_b200')
import test_gpu_golden as T
import parity_cases as pc
device='cuda:0'; dist='normal'
g = T.load("train.pt")[dist]
c = g["cfg"]; cfgs = pkg.configs
over = dict(device=device, num_actions=c["num_actions"], dyn_stoch=8, dyn_discrete=8, dyn_deter=48,
            dyn_hidden=32, units=32, imag_horizon=4, imag_gradient=c["imag_gradient"],
            encoder=dict(mlp_units=40, mlp_layers=2), decoder=dict(mlp_units=40, mlp_layers=2))
cfg = cfgs.make_config("dmc_proprio", **over)
wm = pkg.models.WorldModel(cfgs.ObsSpace(cfgs.PROPRIO_SHAPES), None, 0, cfg)
beh = pkg.models.ImagBehavior(cfg, wm)
wm.load_state_dict(g["wm"], strict=True); beh.actor.load_state_dict(g["actor"], strict=True)
beh.value.load_state_dict(g["value"], strict=True); beh._slow_value.load_state_dict(g["value"], strict=True)
reward_fn = lambda f, s, a: wm.heads["reward"](wm.dynamics.get_feat(s)).mode()
st = g["steps"][0]
n = {k: v.to(device) for k, v in st["noise"].items()}
print({k: tuple(v.shape) for k, v in st["data"].items()})
post, _, m1 = wm._train(st["data"], noise=(n["u_prior"], n["u_post"]))
sd = wm.state_dict()
worst = sorted(((float((sd[k].cpu() - st["wm_after"][k]).abs().max()), k) for k in st["wm_after"]), reverse=True)[:8]
print("wm param diffs after step0:", worst)
_, _, _, _, m2 = beh._train(post, reward_fn, noise=(n["act_noise"], n["u_state"]))
m = {**m1, **m2}
for k in st["metrics"]:
    if k in m:
        r = pc.rel(torch.as_tensor(m[k]), st["metrics"][k])
        if r > 1e-4: print("metric", k, float(torch.as_tensor(m[k])), float(st["metrics"][k]), r)
