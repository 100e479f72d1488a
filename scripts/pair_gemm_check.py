'''Backup selection at the C3 shape (B = 10 000, G = 512) with the CTA-pair GEMM and with the single-CTA GEMM: CUDA-event time of
backup_select_device and agreement of the two selections.  Under ncu (-k regex:umma_f16) for the kernels' own durations.'''
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import olfactory_navigation_b200 as onb
G = int(sys.argv[1]) if len(sys.argv) > 1 else 512
B = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
c = onb.synthetic.config('C2')
env = onb.Environment(**c['env_kwargs'])
agent = onb.agents.PBVI_Agent(env, thresholds=c['thresholds'])
m = agent.model
rng = np.random.default_rng(99)
bs = onb.BeliefSet(m, B)
for _ in range(6):
    bs = bs.update(rng.integers(0, 4, B), (rng.random(B) < 0.15).astype(int))
alphas = torch.as_tensor(np.abs(rng.standard_normal((G, m.padded_states))).astype(np.float32), device=bs._b.device)
out = {}
for mode in ('0', '2', '0', '2'):
    os.environ['OLFNAV_GEMM_PAIR'] = mode
    onb.solvers.backup_select_device(m, bs, alphas, 0.99, 2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ctx, bg, ba, bv = onb.solvers.backup_select_device(m, bs, alphas, 0.99, 2)
    e1.record(); torch.cuda.synchronize()
    out[mode] = (bg, ba, bv)
    print(json.dumps(dict(pair=mode, select_ms=round(e0.elapsed_time(e1) / 3, 3))), flush=True)
print(json.dumps(dict(same_actions=float((out['0'][1] == out['2'][1]).float().mean()), same_alphas=float((out['0'][0] == out['2'][0]).float().mean()),
                      max_rel_value_diff=float(((out['0'][2] - out['2'][2]).abs() / out['0'][2].abs().clamp_min(1e-30)).max()))))
