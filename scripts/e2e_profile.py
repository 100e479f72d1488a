'''
Where the end-to-end time of the bench's `run_test` call goes: the same call as bench.py's e2e leg (C2, 100 000 agents, trained FSVI
agent, 50 steps) under cProfile (host wall clock per function), plus the plain wall time and the device-resident loop for comparison.
Usage: python scripts/e2e_profile.py [agents] [steps]
'''
import os, sys, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import olfactory_navigation_b200 as onb
from olfactory_navigation_b200 import simulation as sim

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
c = onb.synthetic.config('C2')
env = onb.Environment(**c['env_kwargs'])
agent = onb.agents.FSVI_Agent(env, thresholds=c['thresholds'], rng=onb.synthetic.SEED)
agent.train(print_progress=False, print_stats=False, expansions=10)
rng = np.random.default_rng(onb.synthetic.SEED)
starts = env.random_start_points(n, rng)
tshift = rng.integers(0, env.timesteps, n)
kw = dict(n=n, start_points=starts, time_shift=tshift, print_progress=False, print_stats=False, batches=1)
onb.run_test(agent, horizon=3, **kw)
torch.cuda.synchronize()
for rep in range(2):
    t = time.perf_counter()
    hist = onb.run_test(agent, horizon=steps, **kw)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t
    print(f'run_test wall {wall*1e3:.1f} ms for {steps} steps = {wall*1e3/steps:.3f} ms/step, {n*steps/wall/1e6:.2f} M agent-steps/s (upper bound, no terminations counted)')
common = dict(environment=env, time_loop=True, horizon=steps, initialization_values={}, reward_discount=0.99, distance_metric='l1', print_progress=False)
t = time.perf_counter()
sim._run_single(agent, n, starts, tshift, record=False, **common)
torch.cuda.synchronize()
wall = time.perf_counter() - t
print(f'device-resident loop wall {wall*1e3:.1f} ms = {wall*1e3/steps:.3f} ms/step')
pr = cProfile.Profile()
pr.enable()
hist = onb.run_test(agent, horizon=steps, **kw)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(35)
print(s.getvalue()[:6000])
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(20)
print(s.getvalue()[:4000])
