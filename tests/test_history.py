'''
The data format on the far side of run_test: SimulationHistory's per-run tables, analysis tables and CSV save/load
(reference simulation.py:259-386, :492-548, :621-997) against files written by the UNMODIFIED reference for the same fixed
add_step sequence (tests/golden/history_T0.npz, oracle/make_golden_history.py).  CPU only.
'''
import io

import numpy as np
import pandas as pd
import pytest

from conftest import golden, make
from history_feed import feed, STARTS_A, STARTS_B, T0_A, T0_B


def _both(parts: bool = False):
    import olfactory_navigation_b200 as onb
    env, _ = make('T0')
    agent = onb.Agent(env, thresholds=0.5)
    mk = lambda s, t, hz: onb.SimulationHistory(s, env, agent, t, horizon=hz, reward_discount=0.99, used_gpu=False)   # noqa: E731
    ha = feed(mk, STARTS_A, np.arange(len(STARTS_A)) * 3, 1, T0_A)
    hb = feed(mk, STARTS_B, np.arange(len(STARTS_B)) + 40, 2, T0_B)
    return (ha + hb, ha, hb) if parts else ha + hb


def test_history_compare_table():
    g = golden('history_T0')
    both, ha, hb = _both(parts=True)
    mine = both.compare(ha, hb)
    ref = pd.read_csv(io.StringIO(str(g['compare'][()])))
    assert list(mine.columns) == list(ref.columns) and len(mine) == 3
    pd.testing.assert_frame_equal(mine.astype(float), ref.astype(float), rtol=1e-12, atol=0)


def test_history_files_match_the_reference(tmp_path):
    g = golden('history_T0')
    both = _both()
    assert np.array_equal(both.done_at_step, g['done_at_step']) and np.array_equal(both.reached_source, g['reached_source'])
    assert both.summary == str(g['summary'][()])
    both.save(file='hist', folder=str(tmp_path))
    for key, name in (('csv', 'hist.csv'), ('runs_analysis', 'hist-runs_analysis.csv'), ('general_analysis', 'hist-general_analysis.csv')):
        mine = pd.read_csv(tmp_path / name)
        ref = pd.read_csv(io.StringIO(str(g[key][()])))
        assert list(mine.columns) == list(ref.columns), key
        pd.testing.assert_frame_equal(mine, ref, check_dtype=False, rtol=1e-12, atol=0)
    assert (tmp_path / 'hist.csv').read_text() == str(g['csv'][()])             # the simulations file is identical text


@pytest.mark.parametrize('source', ['reference', 'package'])
def test_history_load_round_trip(tmp_path, source):
    'load() rebuilds the history from a file written by the reference, and from one written here.'
    import olfactory_navigation_b200 as onb
    g = golden('history_T0')
    both = _both()
    path = tmp_path / 'hist.csv'
    if source == 'reference':
        path.write_text(str(g['csv'][()]))
    else:
        both.save(file='hist', folder=str(tmp_path), save_analysis=False)
    back = onb.SimulationHistory.load(str(path))
    assert back.n == both.n and back.horizon == both.horizon and back.reward_discount == both.reward_discount
    assert np.array_equal(back.start_points, both.start_points) and np.array_equal(back.time_shift, both.time_shift)
    assert np.array_equal(back.done_at_step, both.done_at_step) and np.array_equal(back.reached_source, both.reached_source)
    assert len(back) == len(both)
    for k in range(len(both)):
        assert np.array_equal(back.positions[k], both.positions[k])
        assert np.array_equal(back.actions[k], both.actions[k])
        assert np.array_equal(back.observations[k], both.observations[k])
    assert back.timestamps == both.timestamps                                 # the second batch crosses midnight
    assert np.array_equal(back.agent_thresholds, both.agent_thresholds)
    assert back.environment_shape == both.environment_shape and back.environment_layer_labels is None
    assert np.array_equal(back.environment_source_position, both.environment_source_position)
    pd.testing.assert_frame_equal(back.runs_analysis_df, both.runs_analysis_df, check_dtype=False)
    pd.testing.assert_frame_equal(back.general_analysis_df.astype(float), both.general_analysis_df.astype(float))
    env, _ = make('T0')
    assert onb.SimulationHistory.load_from_file(str(path), environment=env).environment is env


def test_history_too_short_for_its_properties(tmp_path):
    import olfactory_navigation_b200 as onb
    env, agent = make('T0')
    h = onb.SimulationHistory(np.array([[3, 5]]), env, agent, np.zeros(1, int), horizon=1)
    h.add_step(np.array([0]), np.array([[2, 5]]), np.array([.1]), np.array([True]), np.array([False]), action_is_id=True)
    with pytest.raises(AssertionError):
        h.save(file='x', folder=str(tmp_path))
