'''
The reference's test set-ups (test_setups.py: all start points, n per cell, shape and scale robustness) running on the device
through run_test, and the files they write.
'''
import os

import numpy as np
import pandas as pd
import pytest

from conftest import make

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def trained():
    import olfactory_navigation_b200 as onb
    c = onb.synthetic.config('T1')
    env = onb.Environment(**c['env_kwargs'])
    agent = onb.agents.QMDP_Agent(env, thresholds=c['thresholds'])
    agent.train(print_progress=False, print_stats=False)
    return env, agent


def test_all_starts_and_n_by_cell(trained):
    import olfactory_navigation_b200 as onb
    env, agent = trained
    hist = onb.test_setups.run_all_starts_test(agent, horizon=60, print_progress=False, print_stats=False)
    starts = np.argwhere(env.start_probabilities > 0)
    assert hist.n == len(starts) and np.array_equal(hist.start_points, starts)
    again = onb.run_test(agent, n=len(starts), start_points=starts, horizon=60, print_progress=False, print_stats=False)
    assert np.array_equal(hist.done_at_step, again.done_at_step) and np.array_equal(hist.reached_source, again.reached_source)

    width, per_cell = 5, 3
    hist = onb.test_setups.run_n_by_cell_test(agent, cell_width=width, n_by_cell=per_cell, rng=np.random.default_rng(4), horizon=40,
                                              print_progress=False, print_stats=False)
    H, W = env.shape
    p = env.start_probabilities[:H // width * width, :W // width * width]
    cells_with_mass = int(np.sum(p.reshape(H // width, width, W // width, width).sum(axis=(1, 3)) > 0))
    assert hist.n == cells_with_mass * per_cell
    assert np.all(env.start_probabilities[hist.start_points[:, 0], hist.start_points[:, 1]] > 0)
    cell_of = (hist.start_points // width)
    assert np.all(cell_of.reshape(cells_with_mass, per_cell, 2) == cell_of.reshape(cells_with_mass, per_cell, 2)[:, :1])
    with pytest.raises(AssertionError):
        onb.test_setups.run_all_starts_test(agent, environment=env.modify_scale(1.5), horizon=5, print_progress=False, print_stats=False)


def test_scale_and_shape_robustness_files(trained, tmp_path):
    import olfactory_navigation_b200 as onb
    env, agent = trained
    folder = str(tmp_path / 'scale')
    hists = onb.test_setups.test_scale_robustness(agent, multipliers=[120, 80, 100], horizon=50, print_progress=False, print_stats=False,
                                                  print_warning=False, save_folder=folder)
    assert len(hists) == 3
    assert sorted(os.listdir(folder)) == ['_analysis.csv', 'test_env_mult-100.csv', 'test_env_mult-120.csv', 'test_env_mult-80.csv']
    table = pd.read_csv(folder + '/_analysis.csv')
    assert table['multiplier'].tolist() == [80, 100, 120]
    assert np.allclose(table['converged_mean'], [h.success_count / h.n for h in hists])
    assert np.allclose(table['steps_taken_mean'], [np.mean(h.steps_taken) for h in hists])
    # 100 % is the agent's own environment: same outcome as the plain all-starts run
    plain = onb.test_setups.run_all_starts_test(agent, horizon=50, print_progress=False, print_stats=False)
    assert np.array_equal(hists[1].done_at_step, plain.done_at_step)
    back = onb.SimulationHistory.load(folder + '/test_env_mult-80.csv')
    # the file format cannot tell a run interrupted on the horizon step from one that was still running (same in the reference)
    assert np.array_equal(back.done_at_step, np.where((hists[0].done_at_step == 50) & ~hists[0].reached_source, -1, hists[0].done_at_step))
    assert np.array_equal(back.reached_source, hists[0].reached_source)
    assert back.environment_shape == tuple((np.array(env.shape) * 0.8).astype(int))
    for k in (0, len(back) - 1):
        assert np.array_equal(back.positions[k], hists[0].positions[k])

    folder = str(tmp_path / 'shape')
    hists = onb.test_setups.test_shape_robustness(agent, multipliers=[80, 100], horizon=30, print_progress=False, print_stats=False,
                                                  print_warning=False, save_folder=folder)
    table = pd.read_csv(folder + '/_analysis.csv')
    assert len(hists) == len(table) and set(table['y_multiplier']) <= {80, 100}
    assert all(os.path.exists(f'{folder}/test_env_y-{y}_x-{x}.csv') for y, x in zip(table['y_multiplier'], table['x_multiplier']))
    assert all(h.environment.shape == env.shape for h in hists)               # the plume is stretched inside the same grid


def test_run_test_prints_the_reference_summary(trained, capsys):
    import olfactory_navigation_b200 as onb
    env, agent = trained
    hist = onb.run_test(agent, n=50, horizon=80, print_progress=False, print_stats=True, batches=2)
    out = capsys.readouterr().out
    assert f'Simulations reached goal: {hist.success_count}/50' in out and 'Average discounted rewards (ADR):' in out
    assert sorted(hist.timestamps) == [0, 25]                                  # one timestamp list per batch, keyed by its first run
    runs = hist.runs_analysis_df
    assert np.array_equal(runs['steps_taken'], hist.steps_taken) and np.all(runs['time_s'] > 0)
    assert float(hist.general_analysis_df.loc['mean', 'delta_t1_s']) > 0


def test_agent_in_void(trained):
    'No odor anywhere: the single run starts from the farthest start cell and never observes anything above the threshold.'
    import olfactory_navigation_b200 as onb
    env, agent = trained
    hist = onb.test_setups.test_agent_in_void(agent, horizon=40, print_progress=False, print_stats=False, print_warning=False)
    starts = np.argwhere(env.start_probabilities > 0)
    far = starts[int(np.argmax(env.distance_to_source(starts)))]
    assert hist.n == 1 and np.array_equal(hist.start_points[0], far)
    obs = np.array([o[0] for o in hist.observations])
    assert np.all(obs[obs >= 0] == 0.0)
    assert np.all(np.asarray(env.data) >= 0) and np.any(np.asarray(env.data) > 0)      # the agent's own environment is untouched


def test_agents_by_environments_table(trained, tmp_path):
    import olfactory_navigation_b200 as onb
    env, agent = trained
    other = onb.agents.Infotaxis_Agent(env, thresholds=agent.thresholds[1:-1].tolist())
    untrained = onb.agents.QMDP_Agent(env, thresholds=agent.thresholds[1:-1].tolist())
    envs = [env, env.modify(source_radius=env.source_radius + 1)]
    table = onb.test_setups.test_agents(agent, other, untrained, environments=envs, horizon=40, print_stats=False, print_warning=False,
                                        save_histories_path=str(tmp_path))
    assert list(table.index.get_level_values('agent')) == ['agent_0', 'agent_0', 'agent_1', 'agent_1']      # the untrained one is skipped
    assert table['environment'].tolist() == ['environment_0', 'environment_1'] * 2
    plain = onb.test_setups.run_all_starts_test(agent, horizon=40, print_progress=False, print_stats=False)
    assert np.isclose(table.loc[('agent_0', 0), 'convergence'], plain.success_count / plain.n)
    assert np.isclose(table.loc[('agent_0', 0), 'steps_taken'], np.mean(plain.steps_taken))
    files = sorted(os.listdir(tmp_path))
    assert [f for f in files if f.startswith('Simualtions-')] == [f'Simualtions-agent_{a}-environment_{e}.csv' for a in (0, 1) for e in (0, 1)]
    assert any(f.startswith('Simulation_comparison-') for f in files)
