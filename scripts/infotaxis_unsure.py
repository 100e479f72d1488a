'''
Infotaxis choice at the C4 shape after h steps of run_test (h = 1 ... 50): time of olfnav_infotaxis_choose on the live agents and how many of
them went through the float64 re-evaluation (olfnav_infotaxis_last_unsure).  OLFNAV_INFOTAXIS_KAPPA=0 gives the FP32 path alone.
'''
import os, sys, json, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import olfactory_navigation_b200 as onb
from olfactory_navigation_b200 import simulation as sim
from olfactory_navigation_b200._lib import lib, check, stream
torch.cuda.set_device(0)
c = onb.synthetic.config('C2'); env = onb.Environment(**c['env_kwargs'])
n = int(os.environ.get('N', 100000))
rng = np.random.default_rng(1)
starts = env.random_start_points(n, rng); tshift = rng.integers(0, env.timesteps, n)
agent = onb.agents.Infotaxis_Agent(env, thresholds=c['thresholds'])
for h in (1, 5, 10, 20, 35, 50):
    common = dict(environment=env, time_loop=True, horizon=h, initialization_values={}, reward_discount=0.99, distance_metric='l1', print_progress=False)
    sim._run_single(agent, n, starts, tshift, record=False, **common)
    agent.choose_action_device()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        agent.choose_action_device()
    e1.record(); torch.cuda.synchronize()
    k = ctypes.c_int32(0)
    check(lib.olfnav_infotaxis_last_unsure(agent.model.handle(0), ctypes.byref(k), stream()))
    print(json.dumps(dict(steps=h, live=len(agent.belief), choose_ms=round(e0.elapsed_time(e1) / 3, 3), unsure=int(k.value), unsure_frac=round(k.value / max(len(agent.belief), 1), 4))), flush=True)
