'Diagnostic: tcgen05 evaluate path vs exact expectation using point-mass beliefs (V[i,g] = alpha[g, s_i]).'
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import olfactory_navigation_b200 as onb

torch.cuda.set_device(0)
cfg = sys.argv[1] if len(sys.argv) > 1 else 'T1'
c = onb.synthetic.config(cfg)
env = onb.Environment(**c['env_kwargs'])
agent = onb.agents.PBVI_Agent(env, thresholds=c['thresholds'])
m = agent.model
S = m.state_count
rng = np.random.default_rng(0)
for (n, G) in [(128, 16), (128, 40), (200, 8), (333, 130), (64, 300)]:
    B = np.zeros((n, S)); st = (np.arange(n) * 7) % S; B[np.arange(n), st] = 1.0
    alphas = (np.arange(G)[:, None] * 1000.0 + np.arange(S)[None, :]).astype(np.float64)   # alpha[g, s] = 1000 g + s
    vf = onb.ValueFunction(m, alphas, np.arange(G) % 4)
    bs = onb.BeliefSet(m, B)
    for path in (1, 2):
        try:
            val, arg, _ = vf.evaluate_device(bs, path)
            torch.cuda.synchronize()
        except Exception as e:
            print(f'n={n} G={G} path={path}: EXC {e}'); continue
        val, arg = val.cpu().numpy(), arg.cpu().numpy()
        want_v, want_g = 1000.0 * (G - 1) + st, np.full(n, G - 1)
        bad = np.flatnonzero((np.abs(val - want_v) > 0.01) | (arg != want_g))
        print(f'n={n} G={G} path={path}: {len(bad)} bad rows', '' if not len(bad) else f'first {bad[:6]} got v={val[bad[:6]]} g={arg[bad[:6]]} want v={want_v[bad[:6]]}')
    # random case: precision
    d = -np.log(1.0 - rng.random((n, S))); Br = (d / d.sum(1, keepdims=True)).astype(np.float32).astype(np.float64)
    al = rng.standard_normal((G, S)).astype(np.float32).astype(np.float64)
    vf = onb.ValueFunction(m, al, np.arange(G) % 4); bs = onb.BeliefSet(m, Br)
    sc = Br @ al.T
    for path in (1, 2):
        try:
            val, arg, _ = vf.evaluate_device(bs, path); torch.cuda.synchronize()
        except Exception as e:
            print(f'  random path={path}: EXC {e}'); continue
        err = np.abs(val.cpu().numpy() - sc.max(1)) / np.abs(sc).max()
        print(f'  random path={path}: max rel err {err.max():.3e}, argmax agree {(arg.cpu().numpy() == sc.argmax(1)).mean():.4f}')
