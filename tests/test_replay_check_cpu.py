'''
The replay checker (tests/replay_check.py) on the reference's own golden runs, on the CPU: the reference run must pass it with zero
tie decisions (its decisions are float64 argmaxes of the stored value function), and a tampered history must be rejected.
This pins the checker — and oracle/sim.py's environment functions — against runs of the UNMODIFIED reference.
'''
import numpy as np
import pytest

from conftest import ALL_CFGS, golden
import replay_check


def _replay(g, cfg, acts=None, poss=None, done=None):
    return replay_check.replay(g['actions'].astype(int) if acts is None else acts, g['positions'].astype(int) if poss is None else poss,
                               g['observations'], g['done_at_step'] if done is None else done, g['reached_source'], cfg,
                               g['alphas'].astype(np.float64), g['alpha_actions'], g['start_points'].astype(int), g['time_shift'])


@pytest.mark.parametrize('cfg', ALL_CFGS)
@pytest.mark.parametrize('label', ['qmdp', 'fsvi'])
def test_reference_runs_replay_without_ties(cfg, label):
    g = golden(f'runtest_{label}_{cfg}')
    out = _replay(g, cfg)
    assert out['ties'] == 0 and out['exact'] == out['decisions'] > 500


def test_tampered_histories_are_rejected():
    g = golden('runtest_fsvi_T1')
    acts, poss = g['actions'].astype(int).copy(), g['positions'].astype(int).copy()
    # a wrong decision: agent 0 walks the other way at step 2 (position kept consistent with the move)
    k = 0
    assert g['done_at_step'][k] < 0 or g['done_at_step'][k] > 3
    bad = acts.copy()
    bad[2, k] = -bad[2, k]
    with pytest.raises(AssertionError, match='not argmax|positions differ'):
        _replay(g, 'T1', acts=bad)
    # a wrong position
    badp = poss.copy()
    badp[1, 3, 1] += 1
    with pytest.raises(AssertionError, match='positions differ'):
        _replay(g, 'T1', poss=badp)
    # a wrong termination
    done = g['done_at_step'].copy()
    live = np.flatnonzero(done > 2)[0]
    done[live] -= 1
    with pytest.raises(AssertionError, match='terminations differ'):
        _replay(g, 'T1', done=done)
