'''Experiment: evaluate (tcgen05 path) time at small G under the current environment settings.'''
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import olfactory_navigation_b200 as onb
torch.cuda.set_device(0)
c = onb.synthetic.config('C2'); env = onb.Environment(**c['env_kwargs'])
agent = onb.agents.QMDP_Agent(env, thresholds=c['thresholds']); m = agent.model
S = m.state_count
rng = np.random.default_rng(0)
n = int(os.environ.get('N', 100000))
bs = onb.BeliefSet(m, n)
for _ in range(3):
    bs = bs.update(rng.integers(0, 4, n), (rng.random(n) < 0.15).astype(int))
out = {}
for G in [int(g) for g in os.environ.get('GS', '16,64,128').split(',')]:
    alphas = np.abs(rng.standard_normal((G, S))).astype(np.float32).astype(np.float64)
    vf = onb.ValueFunction(m, alphas, np.arange(G) % 4)
    for _ in range(3): vf.evaluate_device(bs, 2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): vf.evaluate_device(bs, 2)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 10
    out[G] = round(t, 3)
print(os.environ.get('TAG', ''), json.dumps(out), 'GB/s@first', round(4 * S * n / list(out.values())[0] / 1e6))
