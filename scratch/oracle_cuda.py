import sys, os, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import dv3_oracle as O, synth, train_step as TS
dev = "cuda:0"
suite = "dmc_proprio"
d = synth.dims_of(suite); c = synth.CONFIGS[suite]
P, Pa, Pv = synth.agent_params(suite, 0)
mv = lambda D: {k: v.to(dev) for k, v in D.items()}
cfg = TS.make_cfg(actor_layers=c["actor_layers"], actor_dist=c["actor_dist"], units=c["units"])
agent = TS.Agent(mv(P), mv(Pa), mv(Pv), cfg, d)
data = {k: torch.as_tensor(v).to(dev) for k, v in synth.replay_batch(d, 16, 64, 0).items()}
ts = []
for i in range(6):
    noise = {k: v.to(dev) for k, v in synth.train_noise(d, 16, 64, cfg.imag_horizon, i, c["actor_dist"]).items()}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    m = agent.train_step(data, noise)
    torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print("oracle on cuda: step times", [round(t, 3) for t in ts])
