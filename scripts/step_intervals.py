'''Per-step GPU timeline of the device-resident run_test loop: step-to-step interval vs the kernels inside it.'''
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import olfactory_navigation_b200 as onb
from olfactory_navigation_b200 import simulation as sim
torch.cuda.set_device(0)
c = onb.synthetic.config('C2'); env = onb.Environment(**c['env_kwargs'])
agent = onb.agents.FSVI_Agent(env, thresholds=c['thresholds'], rng=onb.synthetic.SEED)
agent.train(expansions=10, print_progress=False, print_stats=False)
n = int(os.environ.get('N', 100000)); steps = int(os.environ.get('STEPS', 12))
rng = np.random.default_rng(1)
starts = env.random_start_points(n, rng); tshift = rng.integers(0, env.timesteps, n)
common = dict(environment=env, time_loop=True, horizon=steps, initialization_values={}, reward_discount=0.99, distance_metric='l1', print_progress=False)
for rep in range(2):
    timing = []
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    sim._run_single(agent, n, starts, tshift, record=False, timing=timing, **common)
    t1.record(); torch.cuda.synchronize()
print('total ms', t0.elapsed_time(t1), 'per step', t0.elapsed_time(t1) / steps)
print('setup before first step ms', t0.elapsed_time(timing[0][1][0]))
for i, (live, evs, m) in enumerate(timing):
    nxt = timing[i + 1][1][0] if i + 1 < len(timing) else t1
    pol = evs[5].elapsed_time(evs[6]) if len(evs) > 6 else float('nan')
    print(f'step {i}: interval {evs[0].elapsed_time(nxt):6.3f}  in-call policy {evs[0].elapsed_time(evs[1]):5.3f}  env+plan {evs[1].elapsed_time(evs[2]):5.3f}  '
          f'update {evs[2].elapsed_time(evs[3]):5.3f}  finish {evs[3].elapsed_time(evs[4]):5.3f}  next policy {pol:5.3f}  tail {(evs[6].elapsed_time(nxt) if len(evs) > 6 else float("nan")):5.3f}')
