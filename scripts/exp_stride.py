'Experiment: belief-step bandwidth vs action pattern and row stride (ld).'
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import olfactory_navigation_b200 as onb
from olfactory_navigation_b200._lib import lib, ptr, stream, check
torch.cuda.set_device(0); dev = torch.device('cuda', 0)
c = onb.synthetic.config('C2'); env = onb.Environment(**c['env_kwargs'])
m = onb.agents.QMDP_Agent(env, thresholds=c['thresholds']).model
S, Sp = m.state_count, m.padded_states; h = m.handle(0)
n = 100_000; rng = np.random.default_rng(0)
def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / iters
pats = {'rand4': rng.integers(0, 4, n), 'allN': np.zeros(n, int), 'allE': np.ones(n, int), 'allS': np.full(n, 2), 'allW': np.full(n, 3),
        'NS': rng.integers(0, 2, n) * 2, 'EW': rng.integers(0, 2, n) * 2 + 1, 'NE': rng.integers(0, 2, n), 'blockNS': (np.arange(n) // 592 % 2) * 2}
for pad in [int(x) for x in (sys.argv[1:] or ['0'])]:
    ld = Sp + pad
    b = torch.zeros((n, ld), dtype=torch.float32, device=dev); o = torch.zeros((n, ld), dtype=torch.float32, device=dev)
    sc = torch.ones(n, dtype=torch.float32, device=dev); so = torch.empty(n, dtype=torch.float32, device=dev); ok = torch.empty(n, dtype=torch.uint8, device=dev)
    check(lib.olfnav_belief_init(h, ptr(b), ld, n, ptr(sc), stream()))
    obs = torch.zeros(n, dtype=torch.int32, device=dev)
    res = []
    for name, a in pats.items():
        act = torch.as_tensor(a, dtype=torch.int32, device=dev)
        t = timeit(lambda: check(lib.olfnav_belief_step(h, ptr(b), ptr(o), ld, n, ptr(act), ptr(obs), None, None, ptr(sc), ptr(so), ptr(ok), stream())))
        res.append(f'{name}={100 * 8 * S * n / t / 1e6 / 6533.8:.0f}%')
    print(f'ld=Sp+{pad}:', ' '.join(res), flush=True)
    del b, o
