'''Does the partial last wave of 128-row tiles cost a whole wave?  evaluate at G = 75 for row counts around multiples of 148 tiles.'''
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import olfactory_navigation_b200 as onb
torch.cuda.set_device(0)
c = onb.synthetic.config('C2'); env = onb.Environment(**c['env_kwargs'])
agent = onb.agents.QMDP_Agent(env, thresholds=c['thresholds']); m = agent.model
rng = np.random.default_rng(0)
G = int(sys.argv[1]) if len(sys.argv) > 1 else 75
vf = onb.ValueFunction(m, np.abs(rng.standard_normal((G, m.state_count))).astype(np.float32).astype(np.float64), np.arange(G) % 4)
full = onb.BeliefSet(m, 113664)
full = full.update(rng.integers(0, 4, 113664), np.zeros(113664, dtype=int), raise_on_impossible_belief=False)
def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for n in (5 * 148 * 128, 100000, 5 * 148 * 128 + 148 * 64, 6 * 148 * 128):
    bs = onb.BeliefSet(m, _buffers=(full._b[:n], full._scale[:n], n), device=0)
    t = timeit(lambda: vf.evaluate_device(bs, 2))
    print(json.dumps(dict(G=G, rows=n, tiles=(n + 127) // 128, waves=round((n + 127) // 128 / 148, 2), ms=round(t, 4), us_per_tile_wave=round(1e3 * t / -(-((n + 127) // 128) // 148), 1),
                          frac_hbm=round(4 * m.state_count * n / t / 1e6 / 6533.8, 3))), flush=True)
