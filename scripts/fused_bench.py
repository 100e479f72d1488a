'''
Fused belief update + evaluation (olfnav_belief_step_evaluate) against the two separate calls, C2 shape by default.
One JSON line per (G, action pattern).  Usage: python scripts/fused_bench.py [config] [agents] [G ...]
'''
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import olfactory_navigation_b200 as onb
from olfactory_navigation_b200._lib import lib, ptr, stream, check

torch.cuda.set_device(0); dev = torch.device('cuda', 0)
name = sys.argv[1] if len(sys.argv) > 1 else 'C2'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
Gs = [int(x) for x in sys.argv[3:]] or [75]
pk = os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')
HBM = json.load(open(pk))['hbm_gbs'] if os.path.exists(pk) else 6533.8

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

c = onb.synthetic.config(name); env = onb.Environment(**c['env_kwargs'])
agent = onb.agents.QMDP_Agent(env, thresholds=c['thresholds']); m = agent.model
S, Sp = m.state_count, m.padded_states
rng = np.random.default_rng(0)
bs = onb.BeliefSet(m, n)
for _ in range(3):
    bs = bs.update(rng.integers(0, 4, n), np.zeros(n, dtype=int), raise_on_impossible_belief=False)
out_b, out_s = torch.empty_like(bs._b), torch.empty_like(bs._scale)
ok = torch.empty(n, dtype=torch.uint8, device=dev)
obs = torch.zeros(n, dtype=torch.int32, device=dev)
action = torch.empty(n, dtype=torch.int32, device=dev)
h = m.handle(0)
pats = {'random': rng.integers(0, 4, n), 'all-N': np.zeros(n), 'all-W': np.full(n, 3), '90%W': np.where(rng.random(n) < 0.9, 3, rng.integers(0, 4, n))}
for G in Gs:
    alphas = np.abs(rng.standard_normal((G, S))).astype(np.float32).astype(np.float64)
    vf = onb.ValueFunction(m, alphas, np.arange(G) % 4); ah = vf.handle(0)
    for pname, a in pats.items():
        act = torch.as_tensor(a, dtype=torch.int32, device=dev)
        def separate():
            check(lib.olfnav_belief_step(h, ptr(bs._b), ptr(out_b), Sp, n, ptr(act), ptr(obs), None, None, ptr(bs._scale), ptr(out_s), ptr(ok), stream()))
            check(lib.olfnav_evaluate(h, ah, ptr(out_b), Sp, n, ptr(out_s), None, None, ptr(action), 0, stream()))
        def fused():
            check(lib.olfnav_belief_step_evaluate(h, ah, ptr(bs._b), ptr(out_b), Sp, n, ptr(act), ptr(obs), None, ptr(bs._scale), ptr(out_s), ptr(ok), None, None, ptr(action), stream()))
        ts, tf = timeit(separate), timeit(fused)
        print(json.dumps(dict(config=name, n=n, S=S, G=G, pattern=pname, separate_ms=round(ts, 4), fused_ms=round(tf, 4),
                              fused_frac_hbm_8S=round(8 * S * n / tf / 1e6 / HBM, 4), fused_agent_steps_per_s=round(n / tf * 1e3))), flush=True)
