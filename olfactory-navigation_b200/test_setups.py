'''
Test set-ups built on run_test — the callers of the hot path the reference ships in `olfactory_navigation/test_setups.py`
(:24-115 all start points, :117-232 n per cell, :287-462 shape robustness, :517-687 scale robustness, :689-820 agents x
environments comparison, :1073-1166 void environment).  Same names, arguments
and outputs (list of SimulationHistory, CSV files, `_analysis.csv`); every simulation runs on the device through run_test.
'''
from __future__ import annotations

import os
from datetime import datetime

import numpy as np

from .agent import Agent
from .environment import Environment
from .simulation import SimulationHistory, run_test

_METRIC_COLS = ['converged', 'reached_horizon', 'steps_taken', 'discounted_rewards', 'extra_steps', 't_min_over_t']
_METRIC_ROWS = ['mean', 'standard_deviation', 'success_mean', 'success_standard_deviation']


def _resolve_environment(agent: Agent, environment: Environment):
    if environment is None:
        return agent.environment, None
    assert environment.shape == agent.environment.shape, "The provided environment's shape doesn't match the environment has been trained on..."
    return environment, environment


def run_all_starts_test(agent: Agent, environment: Environment = None, **run_test_kwargs) -> SimulationHistory:
    'One simulation from every cell with a non-zero start probability (test_setups.py:24-115).'
    env, passed = _resolve_environment(agent, environment)
    start_points = np.argwhere(env.start_probabilities > 0)
    return run_test(agent=agent, n=len(start_points), start_points=start_points, environment=passed, **run_test_kwargs)


def run_n_by_cell_test(agent: Agent, cell_width: int = 10, n_by_cell: int = 10, environment: Environment = None,
                       rng: np.random.Generator = None, **run_test_kwargs) -> SimulationHistory:
    '''
    The grid is cut in cell_width x cell_width cells; every cell that holds start probability gets n_by_cell start points drawn
    from the (renormalised) start probabilities inside it (test_setups.py:117-232).  `rng` (extra argument) makes the draw
    reproducible; the reference draws from NumPy's global state.
    '''
    env, passed = _resolve_environment(agent, environment)
    H, W = env.shape
    cells = np.arange(H * W).reshape(H, W)
    chosen = []
    for i in range(H // cell_width):
        for j in range(W // cell_width):
            window = (slice(i * cell_width, (i + 1) * cell_width), slice(j * cell_width, (j + 1) * cell_width))
            p = env.start_probabilities[window]
            if np.any(p > 0):
                p = (p / p.sum()).ravel()
                draw = rng.choice(cells[window].ravel(), size=n_by_cell, replace=True, p=p) if rng is not None else \
                    np.random.choice(cells[window].ravel(), size=n_by_cell, replace=True, p=p)
                chosen += draw.tolist()
    start_points = np.array(np.unravel_index(chosen, env.shape)).T
    return run_test(agent=agent, n=len(chosen), start_points=start_points, environment=passed, **run_test_kwargs)


def _default_multipliers(env: Environment, step_percentage: int, min_percentage: int, max_percentage: int):
    'Percent steps below 100 down to min_percentage and from 100 up to what the margins allow (test_setups.py:373-380).'
    with np.errstate(divide='ignore'):
        low = env.margins[:, 0] / env.data_source_position + 1
        high = 1 + env.margins[:, 1] / (np.array(env.data_shape) - env.data_source_position)
    max_mult = np.minimum(low, high)
    down = [100 - p for p in range(0, (100 - min_percentage) + step_percentage, step_percentage)[1:]]
    up = list(range(100, min(max_percentage, int(max(max_mult) * 100)), step_percentage))
    return sorted(down + up), max_mult


def _analysis_rows(all_histories, keys: list[dict]):
    import pandas as pd
    rows = []
    for hist, key in zip(all_histories, keys):
        df = hist.general_analysis_df
        if 'reached_horizon' not in df.columns:          # the reference reads a column its own table does not hold; fill it
            df['reached_horizon'] = [float(np.mean(hist.simulations_at_horizon)), None, None, None]
        row = dict(key)
        for col in _METRIC_COLS:
            for metric in _METRIC_ROWS:
                row[f'{col}_{metric}'] = df.loc[metric, col]
        rows.append(row)
    out = pd.DataFrame(rows)
    return out.drop(columns=[f'{c}_{m}' for c in ('converged', 'reached_horizon') for m in _METRIC_ROWS[2:]])


def analyse_shape_robustness(all_histories: list[SimulationHistory], multipliers: np.ndarray):
    return _analysis_rows(all_histories, [{'y_multiplier': int(m[0]), 'x_multiplier': int(m[1])} for m in np.asarray(multipliers)])


def analyse_scale_robustness(all_histories: list[SimulationHistory], multipliers: np.ndarray):
    return _analysis_rows(all_histories, [{'multiplier': int(m)} for m in multipliers])


def _robustness_sweep(kind: str, variants, agent, env, save, save_folder, save_analysis, print_progress, run_test_kwargs):
    'Runs every start point on every modified environment; `variants` yields (label text, file name, environment).'
    start_points = np.argwhere(env.start_probabilities > 0)
    if save or save_analysis:
        if save_folder is None:
            save_folder = f'./results/{datetime.now().strftime("%Y%m%d_%H%M%S")}_{kind}_robustness_test_' + env.name
        os.makedirs(save_folder, exist_ok=True)
        print(f'The results will be saved to: {save_folder}\n')
    histories = []
    for text, file_name, modified in variants:
        if print_progress:
            print(text)
        hist = run_test(agent=agent, n=len(start_points), start_points=start_points, environment=modified,
                        print_progress=False, **run_test_kwargs)
        histories.append(hist)
        if save:
            hist.save(file=file_name, folder=save_folder, save_analysis=False)
    return histories, save_folder


def test_shape_robustness(agent: Agent, environment: Environment = None, step_percentage: int = 20, min_percentage: int = 20,
                          max_percentage: int = 200, multipliers: list[int] = None, print_progress: bool = True, save: bool = True,
                          save_folder: str = None, save_analysis: bool = True, **run_test_kwargs) -> list[SimulationHistory]:
    '''
    Every start point on the plume stretched by each (height %, width %) pair (Environment.modify(multiplier=...)),
    test_setups.py:287-462.  The remaining keyword arguments go to run_test.
    '''
    env, _ = _resolve_environment(agent, environment)
    default, max_mult = _default_multipliers(env, step_percentage, min_percentage, max_percentage)
    multipliers = sorted(multipliers) if multipliers is not None else default
    pairs = np.array(np.meshgrid(multipliers, multipliers, indexing='xy')).T.reshape(-1, 2).astype(float) / 100
    pairs = pairs[np.all(pairs < max_mult, axis=1)]
    variants = ((f'Testing on environment with height {int(m[0] * 100)}% and width {int(m[1] * 100)}%',
                 f'test_env_y-{int(m[0] * 100)}_x-{int(m[1] * 100)}', env.modify(multiplier=m)) for m in pairs)
    histories, save_folder = _robustness_sweep('shape', variants, agent, env, save, save_folder, save_analysis, print_progress, run_test_kwargs)
    if save and save_analysis:
        analyse_shape_robustness(histories, pairs * 100).to_csv(save_folder + '/_analysis.csv', index=False)
        print(f'Shape robustness analysis saved to: {save_folder}/_analysis.csv')
    return histories


def test_scale_robustness(agent: Agent, environment: Environment = None, step_percentage: int = 20, min_percentage: int = 20,
                          max_percentage: int = 200, multipliers: list[int] = None, print_progress: bool = True, save: bool = True,
                          save_folder: str = None, save_analysis: bool = True, **run_test_kwargs) -> list[SimulationHistory]:
    'Every start point on the environment rescaled by each percentage (Environment.modify_scale), test_setups.py:517-687.'
    env, _ = _resolve_environment(agent, environment)
    multipliers = sorted(multipliers) if multipliers is not None else _default_multipliers(env, step_percentage, min_percentage, max_percentage)[0]
    variants = ((f'Testing on environment with scale modifier {m}%', f'test_env_mult-{m}', env.modify_scale(scale_factor=m / 100)) for m in multipliers)
    histories, save_folder = _robustness_sweep('scale', variants, agent, env, save, save_folder, save_analysis, print_progress, run_test_kwargs)
    if save and save_analysis:
        analyse_scale_robustness(histories, multipliers).to_csv(save_folder + '/_analysis.csv', index=False)
        print(f'Scale robustness analysis saved to: {save_folder}/_analysis.csv')
    return histories


def test_agent_in_void(agent: Agent, **run_test_kwargs) -> SimulationHistory:
    '''
    One trajectory in a copy of the agent's environment whose odor data is all zero, from the start cell farthest (manhattan) from
    the source (test_setups.py:1073-1166).  The remaining keyword arguments go to run_test.
    '''
    assert agent.trained, 'Agent must be trained...'
    void = agent.environment.modify()
    void._data = np.zeros_like(void.data)
    void.data_processed = True
    void._device_handles = {}                      # nothing uploaded yet; the device copy is made from the zeroed frames
    starts = np.argwhere(void.start_probabilities > 0)
    if len(starts) == 0:
        raise ValueError('No valid starting positions are available in the void environment.')
    far = starts[int(np.argmax(void.distance_to_source(starts, metric='manhattan')))]
    return run_test(agent=agent, n=1, start_points=far[None, :], environment=void, **run_test_kwargs)


def test_agents(*agents: Agent, environments: list[Environment], initialization_values: list[dict] = None, print_progress: bool = False,
                save_histories_path: str = None, save_result_table: bool = True, **run_test_kwargs):
    '''
    Every trained agent on every environment from every start cell; returns the comparison table (SimulationHistory.compare_all
    per agent, stacked under an `agent` index level) and optionally saves the histories and the table (test_setups.py:689-820).
    '''
    import pandas as pd
    initialization_values = [{}] * len(agents) if initialization_values is None else initialization_values
    tables, names = [], []
    for i_agent, (agent, init) in enumerate(zip(agents, initialization_values)):
        print(f'Testing Agent {i_agent}:')
        if not agent.trained:
            print(f'[Warning] Skipping agent {i_agent} due to it not being marked as trained...')
            continue
        histories = []
        for i_env, environment in enumerate(environments):
            print(f'- Environment {i_env}')
            hist = run_all_starts_test(agent=agent, environment=environment, initialization_values=init, print_progress=print_progress,
                                       **run_test_kwargs)
            histories.append(hist)
            if save_histories_path is not None:
                hist.save(file=f'Simualtions-agent_{i_agent}-environment_{i_env}', folder=save_histories_path, save_analysis=False)
        table = SimulationHistory.compare_all(*histories)
        table['environment'] = [f'environment_{i}' for i in range(len(environments))]
        tables.append(table)
        names.append(f'agent_{i_agent}')
        print('--------------------------------------')
    comparison = pd.concat(tables, keys=names, names=['agent'])
    if save_result_table:
        folder = './' if save_histories_path is None else save_histories_path
        comparison.to_csv(os.path.join(folder, 'Simulation_comparison-' + datetime.now().strftime('%Y%m%d_%H%M%S%f')))
    return comparison


# not collected as tests when the module is imported from a test file
test_agents.__test__ = False
test_agent_in_void.__test__ = False
test_shape_robustness.__test__ = False
test_scale_robustness.__test__ = False
