'''Backup at the C3 shape (B = 10 000 belief points, G alpha vectors): time of the selection (planes + GEMM + pick) and the assembly,
Gamma-free path against the three-call path with Gamma in memory.  Usage: python scripts/backup_breakdown.py [G ...]'''
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import olfactory_navigation_b200 as onb

Gs = [int(v) for v in sys.argv[1:]] or [512]
c = onb.synthetic.config('C2')
env = onb.Environment(**c['env_kwargs'])
agent = onb.agents.PBVI_Agent(env, thresholds=c['thresholds'])
m = agent.model
rng = np.random.default_rng(99)
B = 10_000
bs = onb.BeliefSet(m, B)
for _ in range(6):
    bs = bs.update(rng.integers(0, 4, B), (rng.random(B) < 0.15).astype(int))
dev = bs._b.device


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


for G in Gs:
    alphas = torch.as_tensor(np.abs(rng.standard_normal((G, m.padded_states))).astype(np.float32), device=dev)
    res = {}
    for mode in ('gamma_free', 'gamma_in_memory'):
        os.environ['OLFNAV_BACKUP_GAMMA'] = '0' if mode == 'gamma_free' else '1'
        t_sel, (ctx, bg, ba, bv) = timed(lambda: onb.solvers.backup_select_device(m, bs, alphas, 0.99, 2))
        t_asm, na = timed(lambda: onb.solvers.backup_assemble_device(m, ctx, G, bg, ba, 0))
        res[mode] = dict(select_ms=round(t_sel, 3), assemble_ms=round(t_asm, 3), total_ms=round(t_sel + t_asm, 3))
        res[mode + '_out'] = (na, ba, bg)
    same = torch.equal(res['gamma_free_out'][0], res['gamma_in_memory_out'][0]) and torch.equal(res['gamma_free_out'][2], res['gamma_in_memory_out'][2])
    print(json.dumps(dict(B=B, G=G, gamma_free=res['gamma_free'], gamma_in_memory=res['gamma_in_memory'], identical=bool(same))), flush=True)
