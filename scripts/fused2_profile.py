'''
One-pass update + policy kernel alone at the BASELINE shape: 100 000 agents (C2, 83 x 371), G alpha vectors, random actions.
Prints CUDA-event times of the whole call (plan + kernel + finish) and, with OLFNAV_F2_SPLIT=1, nothing else; meant to be run under
ncu (-k regex:fused2_kernel) for the kernel's own figures.  Usage: python scripts/fused2_profile.py [agents] [G] [reps]
'''
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import olfactory_navigation_b200 as onb
from olfactory_navigation_b200._lib import lib, ptr, stream, check

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 75
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
c = onb.synthetic.config('C2')
env = onb.Environment(**c['env_kwargs'])
agent = onb.agents.PBVI_Agent(env, thresholds=c['thresholds'])
m = agent.model
rng = np.random.default_rng(0)
alphas = np.abs(rng.standard_normal((G, m.state_count))).astype(np.float32).astype(np.float64)
vf = onb.ValueFunction(m, alphas, rng.integers(0, 4, G))
dev = torch.device('cuda', 0)
h, ah = m.handle(0), vf.handle(0)
assert lib.olfnav_can_fuse_policy(h, ah) == 1
bs = onb.BeliefSet(m, n)
for _ in range(2):
    bs = bs.update(rng.integers(0, 4, n), np.zeros(n, dtype=int))
cap = int(lib.olfnav_fused_out_rows(h, n))
out_b = torch.empty((cap, bs._b.shape[1]), dtype=torch.float32, device=dev)
out_s = torch.zeros(cap, dtype=torch.float32, device=dev)
ok = torch.empty(n, dtype=torch.uint8, device=dev)
action = torch.empty(n, dtype=torch.int32, device=dev)
row_out = torch.empty(n, dtype=torch.int32, device=dev)
S = m.state_count
for pattern in ('random', 'all_north', 'all_east'):
    act = torch.as_tensor(rng.integers(0, 4, n) if pattern == 'random' else np.full(n, 0 if pattern == 'all_north' else 1), dtype=torch.int32, device=dev)
    obs = torch.as_tensor((rng.random(n) < 0.1).astype(np.int32), device=dev)
    def call():
        check(lib.olfnav_belief_step_evaluate(h, ah, ptr(bs._b), bs._b.shape[0], ptr(out_b), cap, bs.ld, n, ptr(act), ptr(obs), None,
                                              ptr(bs._scale), ptr(out_s), ptr(ok), None, None, ptr(action), ptr(row_out), stream()))
    call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps(dict(pattern=pattern, n=n, G=G, ms=round(ms, 4), algorithmic_GBps=round(8 * S * n / ms / 1e6, 1), moved_GBps=round(8 * m.padded_states * n / ms / 1e6, 1),
                          agent_steps_per_s=round(n / ms * 1e3))), flush=True)
