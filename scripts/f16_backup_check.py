'''
FP16 2-way split backup GEMM (OLFNAV_GEMM_F16=1) against the FP32 SIMT path and the 3xTF32 path: parity on C1 (several shapes),
then time of the backup selection at the bench's size (C2 model, B = 10 000 belief points, G alpha vectors).
'''
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import olfactory_navigation_b200 as onb   # noqa: E402
from conftest import make                # noqa: E402


def select(m, bs, alphas, path, f16):
    os.environ['OLFNAV_GEMM_F16'] = '1' if f16 else '0'
    out = onb.solvers.backup_device(m, bs, alphas, 0.99, path)
    torch.cuda.synchronize()
    os.environ['OLFNAV_GEMM_F16'] = '0'
    return out


def parity():
    _, agent = make('C1')
    m = agent.model
    rng = np.random.default_rng(12)
    for B, G in ((896, 150), (300, 5), (1000, 64), (130, 33)):
        bs = onb.BeliefSet(m, B)
        for _ in range(4):
            bs = bs.update(rng.integers(0, 4, B), np.zeros(B, dtype=int), raise_on_impossible_belief=False)
        alphas = torch.as_tensor((100 * np.abs(rng.standard_normal((G, m.padded_states))) ** 3).astype(np.float32), device=bs._b.device)
        H, W = m.grid_shape
        alphas.view(G, H, -1)[:, :, W:] = 0
        na1, a1, v1, g1 = select(m, bs, alphas, 1, False)
        na3, a3, v3, g3 = select(m, bs, alphas, 2, False)
        na2, a2, v2, g2 = select(m, bs, alphas, 2, True)
        err = float(((v2 - v1).abs() / v1.abs()).max())
        err3 = float(((v3 - v1).abs() / v1.abs()).max())
        print(json.dumps({'B': B, 'G': G, 'value_rel_err_f16': err, 'value_rel_err_tf32': err3, 'best_a_equal': float((a2 == a1).float().mean()),
                          'best_g_equal': float((g2 == g1).float().mean()), 'best_g_equal_tf32': float((g3 == g1).float().mean())}), flush=True)


def timing():
    c = onb.synthetic.config('C2')
    env = onb.Environment(**c['env_kwargs'])
    agent = onb.agents.QMDP_Agent(env, thresholds=c['thresholds'])
    m = agent.model
    rng = np.random.default_rng(99)
    B = 10_000
    bs = onb.BeliefSet(m, B)
    for _ in range(6):
        bs = bs.update(rng.integers(0, m.action_count, B), (rng.random(B) < 0.15).astype(int))
    for G in (64, 512):
        alphas = torch.as_tensor(np.abs(rng.standard_normal((G, m.padded_states))).astype(np.float32), device=bs._b.device)
        res = {}
        for name, f16 in (('tf32x3', False), ('fp16x2', True)):
            for _ in range(2):
                out = select(m, bs, alphas, 2, f16)
            os.environ['OLFNAV_GEMM_F16'] = '1' if f16 else '0'
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                out = onb.solvers.backup_device(m, bs, alphas, 0.99, 2)
            e1.record()
            torch.cuda.synchronize()
            os.environ['OLFNAV_GEMM_F16'] = '0'
            res[name] = (e0.elapsed_time(e1) / 3, out)
        v_a, v_b = res['tf32x3'][1][2], res['fp16x2'][1][2]
        print(json.dumps({'G': G, 'B': len(bs), 'ms_tf32x3': res['tf32x3'][0], 'ms_fp16x2': res['fp16x2'][0],
                          'value_rel_diff': float(((v_a - v_b).abs() / v_a.abs()).max()),
                          'best_a_equal': float((res['tf32x3'][1][1] == res['fp16x2'][1][1]).float().mean()),
                          'best_g_equal': float((res['tf32x3'][1][3] == res['fp16x2'][1][3]).float().mean())}), flush=True)


if __name__ == '__main__':
    parity()
    if len(sys.argv) > 1 and sys.argv[1] == 'time':
        timing()
