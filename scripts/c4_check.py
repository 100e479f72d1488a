'''
BASELINE configs[3]: Infotaxis agent simulation on the 83x371-state environment, 100 000 agents, expected-entropy action selection.
Runs run_test for a few steps and reports the step time (policy = olfnav_infotaxis_choose, update = olfnav_belief_step).
Also: QMDP agent (G = 4) on the same shape, the cheapest policy.
'''
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import olfactory_navigation_b200 as onb
from olfactory_navigation_b200 import simulation as sim
torch.cuda.set_device(0)
c = onb.synthetic.config('C2'); env = onb.Environment(**c['env_kwargs'])
n = int(os.environ.get('N', 100000)); steps = int(os.environ.get('STEPS', 20))
rng = np.random.default_rng(1)
starts = env.random_start_points(n, rng); tshift = rng.integers(0, env.timesteps, n)
common = dict(environment=env, time_loop=True, horizon=steps, initialization_values={}, reward_discount=0.99, distance_metric='l1', print_progress=False)
pk = os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')
HBM = json.load(open(pk))['hbm_gbs'] if os.path.exists(pk) else 6533.8
for name, agent in (('infotaxis', onb.agents.Infotaxis_Agent(env, thresholds=c['thresholds'])), ('qmdp', onb.agents.QMDP_Agent(env, thresholds=c['thresholds']))):
    if name == 'qmdp':
        agent.train(print_progress=False, print_stats=False)
    S = agent.model.state_count
    for rep in range(2):
        timing = []
        t1 = torch.cuda.Event(enable_timing=True)
        hist = sim._run_single(agent, n, starts, tshift, record=False, timing=timing, **common)
        t1.record(); torch.cuda.synchronize()
    loop = timing[0][1][0].elapsed_time(t1)
    upd = [e[2].elapsed_time(e[3]) for _, e, _ in timing]
    pol = [e[5].elapsed_time(e[6]) for _, e, _ in timing if len(e) > 6]
    live = np.mean([mm for _, _, mm in timing])
    print(json.dumps(dict(row=f'C4-shape run_test step, {name} policy', states=S, agents=n, ms_per_step=loop / steps, agent_steps_per_s=hist._agent_steps / loop * 1e3,
                          belief_step_ms=float(np.mean(upd)), belief_step_frac_hbm=8.0 * S * live / np.mean(upd) / 1e6 / HBM,
                          policy_ms=float(np.mean(pol)), policy_frac_hbm=4.0 * S * live / np.mean(pol) / 1e6 / HBM)), flush=True)
