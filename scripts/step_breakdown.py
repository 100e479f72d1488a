'Per-phase wall/GPU breakdown of the device-resident step loop (diagnostic).'
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import olfactory_navigation_b200 as onb
from olfactory_navigation_b200._lib import lib, ptr, stream, check
torch.cuda.set_device(0); dev = torch.device('cuda', 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
c = onb.synthetic.config('C2'); env = onb.Environment(**c['env_kwargs'])
agent = onb.agents.FSVI_Agent(env, thresholds=c['thresholds'], rng=1)
agent.train(expansions=int(sys.argv[2]) if len(sys.argv) > 2 else 10, print_progress=False, print_stats=False)
print('G =', len(agent.value_function))
rng = np.random.default_rng(0)
starts = env.random_start_points(n, rng); tshift = rng.integers(0, env.timesteps, n)
pos = torch.as_tensor(starts, dtype=torch.int32, device=dev).contiguous(); ts = torch.as_tensor(tshift, dtype=torch.int64, device=dev)
moves = torch.as_tensor(np.ascontiguousarray(agent.action_set), dtype=torch.int32, device=dev); thr = torch.as_tensor(agent.thresholds, dtype=torch.float64, device=dev)
odor = torch.empty(n, dtype=torch.float64, device=dev); oid = torch.empty(n, dtype=torch.int32, device=dev); rch = torch.empty(n, dtype=torch.uint8, device=dev)
env_h = env.device_handle(0); T = env.timesteps
agent.initialize_state(n); live = n
torch.cuda.synchronize()
names = ['evaluate', 'sort', 'env_step', 'update', 'filter+kill']
for i in range(14):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    h = [time.perf_counter()]
    ev[0].record()
    act = agent.choose_action_device(); ev[1].record(); h.append(time.perf_counter())
    t_eff = torch.sort((ts[:live] + i) % T).values.contiguous(); ev[2].record(); h.append(time.perf_counter())
    check(lib.olfnav_env_step(env_h, live, ptr(pos), ptr(act), ptr(moves), ptr(t_eff), 0, ptr(thr), 3, ptr(odor), ptr(oid), ptr(rch), stream())); ev[3].record(); h.append(time.perf_counter())
    reached = rch[:live].bool()
    ok = agent.update_state_device(oid[:live], reached); ev[4].record(); h.append(time.perf_counter())
    term = reached | ~ok; keep = ~term
    new_pos, new_ts = pos[:live][keep].contiguous(), ts[:live][keep].contiguous()
    agent.kill_device(term); ev[5].record(); h.append(time.perf_counter())
    nl = int(new_pos.shape[0]); pos[:nl], ts[:nl] = new_pos, new_ts; live = nl
    torch.cuda.synchronize(); h.append(time.perf_counter())
    print(i, live, 'GPU ms:', ' '.join(f'{nm}={a.elapsed_time(b):.2f}' for nm, a, b in zip(names, ev, ev[1:])),
          '| HOST ms:', ' '.join(f'{1e3*(b-a):.2f}' for a, b in zip(h, h[1:])))
