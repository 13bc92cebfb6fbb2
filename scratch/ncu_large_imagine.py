"""Workload for ncu: ONE ImagBehavior._imagine forward at BASELINE config 4 (1024 starts x H=15,
dyn_deter 4096, dyn_hidden / units 1024, 5-layer one-hot actor) between cudaProfilerStart/Stop.
    ncu --profile-from-start off --set full --clock-control none -k regex:umma ... python scratch/ncu_large_imagine.py"""
import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('dreamerv3-torch_b200')
cfgs = pkg.configs
dev = 'cuda:0'
torch.manual_seed(0)
cfg = cfgs.make_config("crafter", device=dev, device_metrics=True, encoder=dict(mlp_keys=".*", cnn_keys="$^"),
                       decoder=dict(mlp_keys=".*", cnn_keys="$^"))
wm = pkg.models.WorldModel(cfgs.ObsSpace(cfgs.PROPRIO_SHAPES), None, 0, cfg)
beh = pkg.models.ImagBehavior(cfg, wm)
B, T, S, C, D, H = 16, 64, cfg.dyn_stoch, cfg.dyn_discrete, cfg.dyn_deter, cfg.imag_horizon
g = torch.Generator(device=dev).manual_seed(1)
idx = torch.randint(0, C, (B, T, S), device=dev, generator=g)
start = dict(stoch=torch.nn.functional.one_hot(idx, C).float(),
             deter=torch.tanh(torch.randn(B, T, D, device=dev, generator=g)),
             logit=torch.randn(B, T, S, C, device=dev, generator=g))
with torch.no_grad():
    for _ in range(2):
        beh._imagine(start, beh.actor, H)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    beh._imagine(start, beh.actor, H)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("ok")
