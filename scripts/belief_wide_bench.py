'''
Belief step alone on a named synthetic config (default C5: 166 x 742 = 123 172 states), every action pattern, out of place and in place.
One JSON line per pattern.  Usage: [OLFNAV_BELIEF_THREADS=256|512] [OLFNAV_BELIEF_STAGGER=k] python scripts/belief_wide_bench.py [config] [agents]
'''
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import olfactory_navigation_b200 as onb
from olfactory_navigation_b200._lib import lib, ptr, stream, check

torch.cuda.set_device(0); dev = torch.device('cuda', 0)
name = sys.argv[1] if len(sys.argv) > 1 else 'C5'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25_000
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
h = m.handle(0)
tag = dict(config=name, n=n, S=S, threads=os.environ.get('OLFNAV_BELIEF_THREADS', 'auto'), stagger=os.environ.get('OLFNAV_BELIEF_STAGGER', 'default'))
pats = {'random': rng.integers(0, 4, n), 'all-N': np.zeros(n), 'all-E': np.ones(n), 'all-S': np.full(n, 2), 'all-W': np.full(n, 3),
        '90%W': np.where(rng.random(n) < 0.9, 3, rng.integers(0, 4, n))}
for pname, a in pats.items():
    act = torch.as_tensor(a, dtype=torch.int32, device=dev)
    t = timeit(lambda: check(lib.olfnav_belief_step(h, ptr(bs._b), ptr(out_b), Sp, n, ptr(act), ptr(obs), None, None, ptr(bs._scale), ptr(out_s), ptr(ok), stream())))
    print(json.dumps(dict(tag, pattern=pname, place='out', ms=round(t, 4), frac_hbm=round(8 * S * n / t / 1e6 / HBM, 4))), flush=True)
for pname in ('random', 'all-N', 'all-W'):
    act = torch.as_tensor(pats[pname], dtype=torch.int32, device=dev)
    t = timeit(lambda: check(lib.olfnav_belief_step(h, ptr(bs._b), ptr(bs._b), Sp, n, ptr(act), ptr(obs), None, None, ptr(bs._scale), ptr(bs._scale), ptr(ok), stream())))
    print(json.dumps(dict(tag, pattern=pname, place='in', ms=round(t, 4), frac_hbm=round(8 * S * n / t / 1e6 / HBM, 4))), flush=True)
