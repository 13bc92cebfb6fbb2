import importlib, sys, os, time, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('dreamerv3-torch_b200')
cfgs = pkg.configs; dev = 'cuda:0'
torch.manual_seed(0)
cfg = cfgs.make_config('dmc_proprio', device=dev, device_metrics=True)
wm = pkg.models.WorldModel(cfgs.ObsSpace(cfgs.PROPRIO_SHAPES), None, 0, cfg)
beh = pkg.models.ImagBehavior(cfg, wm)
rs = np.random.RandomState(0); B, T, A = 16, 64, 6
host = {k: rs.randn(B, T, n).astype(np.float32) for k, n in (("orientations", 14), ("height", 1), ("velocity", 9))}
host["action"] = rs.uniform(-1, 1, size=(B, T, A)).astype(np.float32)
host["reward"] = rs.randn(B, T).astype(np.float32); host["discount"] = np.ones((B, T), np.float32)
host["is_terminal"] = np.zeros((B, T), np.float32); host["is_first"] = np.zeros((B, T), np.float32); host["is_first"][:, 0] = 1
pinned = {k: torch.from_numpy(v).pin_memory() for k, v in host.items()}
g = pkg.graphs.TrainStepGraph(wm, beh, warmup=3, device_metrics=True)
losses = []
for i in range(6):
    out = g(pinned); losses.append(float(out["wm_metrics"]["model_loss"]))
print("captured", g.captured, "model_loss per step", [round(x, 4) for x in losses])
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): out = g(pinned)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20 * 1e3
print("graph step ms", dt, "model_loss", float(out["wm_metrics"]["model_loss"]), "actor_loss", float(out["beh_metrics"]["actor_loss"]))
