'''Timeline experiment of the fused kernel: OLFNAV_FUSED_DBG=1024(+bits) python scripts/fused_timeline.py  (clock64 stamps on stderr).'''
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import olfactory_navigation_b200 as onb
from olfactory_navigation_b200._lib import lib, ptr, stream, check
torch.cuda.set_device(0); dev = torch.device('cuda', 0)
n = 100_000
c = onb.synthetic.config('C2'); env = onb.Environment(**c['env_kwargs'])
agent = onb.agents.QMDP_Agent(env, thresholds=c['thresholds']); m = agent.model
S, Sp = m.state_count, m.padded_states
rng = np.random.default_rng(0)
bs = onb.BeliefSet(m, n)
bs = bs.update(rng.integers(0, 4, n), np.zeros(n, dtype=int), raise_on_impossible_belief=False)
out_b, out_s = torch.empty_like(bs._b), torch.empty_like(bs._scale)
ok = torch.empty(n, dtype=torch.uint8, device=dev); obs = torch.zeros(n, dtype=torch.int32, device=dev); action = torch.empty(n, dtype=torch.int32, device=dev)
act = torch.as_tensor(rng.integers(0, 4, n), dtype=torch.int32, device=dev)
vf = onb.ValueFunction(m, np.abs(rng.standard_normal((75, S))).astype(np.float32).astype(np.float64), np.arange(75) % 4)
for _ in range(2):
    check(lib.olfnav_belief_step_evaluate(m.handle(0), vf.handle(0), ptr(bs._b), ptr(out_b), Sp, n, ptr(act), ptr(obs), None, ptr(bs._scale), ptr(out_s), ptr(ok), None, None, ptr(action), stream()))
torch.cuda.synchronize()
