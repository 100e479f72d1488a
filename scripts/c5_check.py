'''
BASELINE configs[4] shard: the 2x-resolution plume (166 x 742 = 123 172 states), 125 000 agents on one B200
(the per-GPU share of 10^6 agents on 8 GPUs): 2 x 61.6 GB of beliefs.  Runs run_test for a few steps and reports the step time.
'''
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import olfactory_navigation_b200 as onb
from olfactory_navigation_b200 import simulation as sim
torch.cuda.set_device(0)
c = onb.synthetic.config('C5'); env = onb.Environment(**c['env_kwargs'])
t0 = time.time()
agent = onb.agents.FSVI_Agent(env, thresholds=c['thresholds'], rng=onb.synthetic.SEED)
agent.train(expansions=int(os.environ.get('EXP', 6)), print_progress=False, print_stats=False)
m = agent.model
print('model', m.grid_shape, 'states', m.state_count, 'alphas', len(agent.value_function), f'train {time.time() - t0:.1f}s', flush=True)
n = int(os.environ.get('N', 125000)); steps = int(os.environ.get('STEPS', 10))
rng = np.random.default_rng(1)
starts = env.random_start_points(n, rng); tshift = rng.integers(0, env.timesteps, n)
common = dict(environment=env, time_loop=True, horizon=steps, initialization_values={}, reward_discount=0.99, distance_metric='l1', print_progress=False)
for rep in range(2):
    timing = []
    t1 = torch.cuda.Event(enable_timing=True)
    hist = sim._run_single(agent, n, starts, tshift, record=False, timing=timing, **common)
    t1.record(); torch.cuda.synchronize()
loop = timing[0][1][0].elapsed_time(t1)
upd = [e[2].elapsed_time(e[3]) for _, e, _ in timing]
pol = [e[5].elapsed_time(e[6]) for _, e, _ in timing if len(e) > 6]
S = m.state_count
out = dict(row='C5 shard run_test step', states=S, agents=n, alphas=len(agent.value_function), ms_per_step=loop / steps, agent_steps_per_s=hist._agent_steps / loop * 1e3,
           belief_step_ms=float(np.mean(upd)), belief_step_GBps=8.0 * S * np.mean([mm for _, _, mm in timing]) / np.mean(upd) / 1e6, policy_ms=float(np.mean(pol)),
           mem_GB=torch.cuda.max_memory_allocated() / 1e9)
print(json.dumps(out))
