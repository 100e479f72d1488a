'''
run_test / SimulationHistory — device-resident version of the reference's simulation driver
(reference olfactory_navigation/simulation.py:1224-1519 and :34-227).

Same signature, same step order (act on b_t -> move -> observe at the new cell -> update to
b_{t+1} -> kill; simulation.py:1452-1497) and the same recorded quantities, but the whole step
stays on the GPU: positions, time shifts and beliefs never leave HBM; per step only the O(N) record
(action id, position, odor, flags) is copied to pinned host memory, asynchronously, on a side
stream.  Host <-> device boundary: start points / time shifts in (once), records out (per step).

The data format on the far side of the path — the per-run / general analysis tables and the CSV file a history is saved to
and loaded from (simulation.py:259-386, :492-548, :621-997) — is kept column for column, so files written here load in the
reference and vice versa.  Plots are not rebuilt.
'''
from __future__ import annotations

import math
from datetime import datetime, timedelta

import os
import time

import numpy as np
import torch

from ._lib import lib, ptr, stream, check
from .agent import Agent
from .environment import Environment


class SimulationHistory:
    '''
    Per-step record of n parallel simulations.  Mirrors the attributes of the reference's
    SimulationHistory that run_test fills (simulation.py:137-162,168-227): `actions`,
    `positions`, `observations` (lists over steps of dense (n, .) arrays with -1 for finished
    simulations — built lazily here), `reached_source`, `done_at_step` (1-based, -1 if never
    done), `timestamps`, and the summary statistics (:423-481).
    '''
    def __init__(self, start_points: np.ndarray, environment: Environment, agent: Agent, time_shift: np.ndarray,
                 horizon: int, reward_discount: float = 0.99, distance_metric: str = 'l1', used_gpu: bool = True) -> None:
        start_points = np.asarray(start_points)
        if start_points.ndim == 1:
            start_points = start_points[None, :]
        self.n = len(start_points)
        self.environment = environment
        self.agent = agent
        self.time_shift = np.asarray(time_shift)
        self.horizon = horizon
        self.reward_discount = reward_discount
        self.distance_metric = distance_metric
        self.used_gpu = used_gpu
        self.start_points = start_points
        self.timestamps = {0: [datetime.now()]}
        self.reached_source = np.zeros(self.n, dtype=bool)
        self.done_at_step = np.full(self.n, fill_value=-1)
        self._running_sims = np.arange(self.n)
        self._steps = []          # compact records: (sim ids, action ids, positions, observations)
        self._dense = None
        self._simulation_dfs = None
        self.environment_dimensions = environment.dimensions
        self.environment_shape = environment.shape
        self.environment_source_position = environment.source_position
        self.environment_source_radius = environment.source_radius
        self.environment_layer_labels = environment.layer_labels
        self.agent_thresholds = agent.thresholds

    def add_step(self, actions: np.ndarray, next_positions: np.ndarray, observations: np.ndarray,
                 reached_source: np.ndarray, interupt: np.ndarray, action_is_id: bool = False) -> None:
        'Same contract as the reference\'s add_step (simulation.py:168-227); arrays cover the running simulations.'
        self._dense = None
        self._simulation_dfs = None
        self.timestamps[0].append(datetime.now())
        sims = self._running_sims
        self._steps.append((sims, np.asarray(actions), np.asarray(next_positions), np.asarray(observations, dtype=float), action_is_id))
        done = np.asarray(reached_source, dtype=bool) | np.asarray(interupt, dtype=bool)
        self.done_at_step[sims[done]] = len(self._steps)
        self.reached_source[sims[np.asarray(reached_source, dtype=bool)]] = True
        self._running_sims = sims[~done]

    def _densify(self) -> None:
        if self._dense is not None:
            return
        acts, poss, obss = [], [], []
        action_set = np.asarray(self.agent.action_set)
        for sims, a, p, o, is_id in self._steps:
            av = action_set[a] if is_id else a
            A = np.full((self.n, action_set.shape[1]), -1)
            A[sims] = av
            P = np.full((self.n, self.environment_dimensions), -1)
            P[sims] = p
            O = np.full((self.n,), -1.0)
            O[sims] = o
            acts.append(A)
            poss.append(P)
            obss.append(O)
        self._dense = (acts, poss, obss)

    @property
    def actions(self) -> list:
        self._densify()
        return self._dense[0]

    @property
    def positions(self) -> list:
        self._densify()
        return self._dense[1]

    @property
    def observations(self) -> list:
        self._densify()
        return self._dense[2]

    def __len__(self) -> int:
        return len(self._steps)

    # -- statistics (simulation.py:388-481) -----------------------------------------------------
    @property
    def done_count(self) -> int:
        return self.n - len(self._running_sims)

    @property
    def successful_simulations(self) -> np.ndarray:
        return self.reached_source

    @property
    def success_count(self) -> int:
        return int(np.sum(self.reached_source))

    @property
    def simulations_at_horizon(self) -> np.ndarray:
        return self.done_at_step == -1

    @property
    def error_simulations(self) -> np.ndarray:
        return (self.done_at_step >= 0) & ~self.reached_source

    @property
    def error_count(self) -> int:
        return int(np.sum(self.error_simulations))

    def compute_distance_to_source(self) -> np.ndarray:
        d = self.environment_source_position[None, :] - self.start_points
        if self.distance_metric == 'l1':
            return np.sum(np.abs(d), axis=-1) - self.environment_source_radius
        if self.distance_metric == 'l2':
            return np.sqrt(np.sum(d ** 2, axis=-1)) - self.environment_source_radius
        raise NotImplementedError('This distance metric has not yet been implemented')

    @property
    def steps_taken(self) -> np.ndarray:
        return np.where(self.done_at_step >= 0, self.done_at_step, len(self._steps))

    @property
    def summary(self) -> str:
        'Success count and mean +- standard deviation of the run metrics, all runs and successful ones (simulation.py:438-481).'
        n_ok = self.success_count
        text = f'Simulations reached goal: {n_ok}/{self.n} ({(n_ok * 100) / self.n:.2f}% success)'
        if self.error_count > 0:
            text += f' - {self.error_count} AGENTS ENCOUNTERED ERRORS'
        if n_ok == 0:
            return text
        df = self.general_analysis_df
        for label, col in (('Average step count:', 'steps_taken'), ('Extra steps:', 'extra_steps'),
                           ('Average discounted rewards (ADR):', 'discounted_rewards'), ('Tmin/T:', 't_min_over_t'),
                           ('Average time to completion (s):', 'time_s'), ('Average time per iteration (s):', 'time_per_it_s')):
            v = [float(df.loc[row, col]) for row in ('mean', 'standard_deviation', 'success_mean', 'success_standard_deviation')]
            text += f"\n - {label:<35} {v[0]:.3f} +- {v[1]:.2f} (Successful only: {v[2]:.3f} +- {v[3]:.2f})"
        return text

    def __add__(self, other: 'SimulationHistory') -> 'SimulationHistory':
        'Concatenate two batches (simulation.py:550-618).'
        out = SimulationHistory(np.vstack([self.start_points, other.start_points]), self.environment, self.agent,
                                np.concatenate([self.time_shift, other.time_shift]), self.horizon, self.reward_discount,
                                self.distance_metric, self.used_gpu)
        out.reached_source = np.concatenate([self.reached_source, other.reached_source])
        out.done_at_step = np.concatenate([self.done_at_step, other.done_at_step])
        out._running_sims = np.concatenate([self._running_sims, other._running_sims + self.n])
        steps = []
        for k in range(max(len(self._steps), len(other._steps))):
            parts = []
            if k < len(self._steps):
                parts.append(self._steps[k])
            if k < len(other._steps):
                sims, a, p, o, is_id = other._steps[k]
                parts.append((sims + self.n, a, p, o, is_id))
            if len(parts) == 1:
                steps.append(parts[0])
            else:
                (s1, a1, p1, o1, i1), (s2, a2, p2, o2, i2) = parts
                assert i1 == i2
                steps.append((np.concatenate([s1, s2]), np.concatenate([a1, a2]), np.concatenate([p1, p2]), np.concatenate([o1, o2]), i1))
        out._steps = steps
        out.timestamps = {**self.timestamps, **{k + self.n: v for k, v in other.timestamps.items()}}    # keyed by the first run of each batch
        return out

    # -- analysis tables (simulation.py:259-386) -------------------------------------------------
    @property
    def _axes_labels(self) -> list:
        d = self.environment_dimensions
        return ['z', 'y', 'x'][-d:] if d <= 3 else [f'x{i}' for i in range(d)]

    def _batches(self):
        'Yields (first run, one past the last run, seconds since the batch started at the end of every step).'
        bounds = sorted(self.timestamps.keys()) + [self.n]
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            ts = self.timestamps[lo]
            yield lo, hi, np.array([(t - ts[0]).total_seconds() for t in ts[1:]])

    @property
    def start_time(self) -> datetime:
        return self.timestamps[0][0]

    @property
    def runs_analysis_df(self):
        'One row per run: start point, optimal step count, outcome, steps, discounted reward, extra steps, Tmin/T, wall time.'
        import pandas as pd
        run_times = np.zeros(self.n)
        for lo, hi, cum in self._batches():
            end = self.done_at_step[lo:hi]
            run_times[lo:hi] = cum[np.where(end > 0, end - 1, end)] if len(cum) else 0.0
        df = pd.DataFrame(self.start_points, columns=self._axes_labels)
        df['optimal_steps_count'] = self.compute_distance_to_source()
        df['converged'] = self.reached_source
        df['encountered_error'] = self.error_simulations
        df['steps_taken'] = self.steps_taken
        df['discounted_rewards'] = self.reward_discount ** df['steps_taken']
        df['extra_steps'] = df['steps_taken'] - df['optimal_steps_count']
        df['t_min_over_t'] = df['optimal_steps_count'] / df['steps_taken']
        df['time_s'] = run_times
        df['time_per_it_s'] = df['time_s'] / df['steps_taken']
        df.index = [f'run_{i}' for i in range(self.n)]
        return df

    @property
    def general_analysis_df(self):
        'mean / standard deviation over all runs and over the successful ones, run counts, and the two time-per-step figures.'
        import pandas as pd
        runs = self.runs_analysis_df
        cols = ['converged', 'encountered_error', 'steps_taken', 'discounted_rewards', 'extra_steps', 't_min_over_t', 'time_s', 'time_per_it_s']
        ok = runs.loc[runs['converged'], cols]
        out = pd.DataFrame([runs[cols].mean(), runs[cols].std(), ok.mean(), ok.std()],
                           index=['mean', 'standard_deviation', 'success_mean', 'success_standard_deviation'], columns=cols).astype(object)
        out.loc[['standard_deviation', 'success_mean', 'success_standard_deviation'], ['converged', 'encountered_error']] = None
        out.insert(0, 'count', [self.n, None, self.success_count, None])
        # seconds per agent-step until the first agent of each batch is done (t1) and until the last one is (t2);
        # both are divided by the t1 step count, like the reference does (simulation.py:377-378)
        time_t1 = time_t2 = 0.0
        steps_t1 = 0
        for lo, hi, cum in self._batches():
            end = self.done_at_step[lo:hi]
            end = np.where(end > 0, end, self.horizon)
            last = len(cum) - 1                               # (a history cut short of its horizon has fewer timestamps)
            time_t1 += cum[min(end.min() - 1, last)] if len(cum) else 0.0
            time_t2 += cum[min(end.max() - 1, last)] if len(cum) else 0.0
            steps_t1 += (hi - lo) * int(end.min())
        out['delta_t1_s'] = [time_t1 / steps_t1, None, None, None]
        out['delta_t2_s'] = [time_t2 / steps_t1, None, None, None]
        return out

    # -- per-run tables and the CSV format (simulation.py:492-548, :621-997) -----------------------
    @property
    def simulation_dfs(self) -> list:
        '''
        One table per run: row 0 is the start (time shift, start point, nothing else), row k the k-th step
        (time, position reached, action vector played, odor observed, 1 on the row where the source was reached).
        '''
        import pandas as pd
        if getattr(self, '_simulation_dfs', None) is not None:
            return self._simulation_dfs
        pos, act, obs = np.array(self.positions), np.array(self.actions), np.array(self.observations)
        axes = self._axes_labels
        layered = self.environment_layer_labels is not None
        dfs = []
        for i in range(self.n):
            length = int(self.done_at_step[i]) if self.done_at_step[i] >= 0 else len(pos)
            col = {'time': np.arange(length + 1) + self.time_shift[i]}
            for k, axis in enumerate(axes):
                col[axis] = np.hstack([self.start_points[i, k], pos[:length, i, k]]) if length else np.array([self.start_points[i, k]])
            if layered:
                col['layer'] = np.hstack([[None], act[:length, i, 0]]) if length else np.array([None])
            for k, axis in enumerate(axes):
                col['d' + axis] = np.hstack([[None], act[:length, i, k + (1 if layered else 0)]]) if length else np.array([None])
            col['o'] = np.hstack([[None], obs[:length, i]]) if length else np.array([None])
            col['reached_source'] = np.hstack([[None], np.where(np.arange(1, length + 1) == self.done_at_step[i], int(self.reached_source[i]), 0)])
            dfs.append(pd.DataFrame(col))
        self._simulation_dfs = dfs
        return dfs

    def save(self, file: str = None, folder: str = None, save_analysis: bool = True, save_components: bool = False) -> None:
        '''
        One CSV with every run stacked, the step timestamps of each batch, and a property / property_value column pair with
        what is needed to rebuild the history (same columns and property names as simulation.py:621-740).
        '''
        import pandas as pd
        assert (self.environment is not None) and (self.agent is not None), 'Function not available, the agent and/or the environment is not set.'
        if file is None:
            shape = '_'.join(str(v) for v in self.environment_shape)
            file = f'Simulations-s_{shape}-n_{self.n}-{self.start_time.strftime("%Y%m%d_%H%M%S")}-horizon_{self.horizon}.csv'
        if not file.endswith('.csv'):
            file += '.csv'
        folder = './' if folder is None else folder
        if '/' not in folder:
            folder = './' + folder
        if not os.path.exists(folder):
            os.mkdir(folder)
        if not folder.endswith('/'):
            folder += '/'
        if save_components:
            if (self.environment.saved_at is None) or (folder not in self.environment.saved_at):
                self.environment.save(folder=folder)
            if (self.agent.saved_at is None) or (folder not in self.agent.saved_at):
                self.agent.save(folder=folder)

        dfs = self.simulation_dfs
        table = pd.concat(dfs)
        first_row = np.concatenate([[0], np.cumsum([len(d) for d in dfs])])
        stamps = [None] * len(table)
        for run, ts in self.timestamps.items():
            for k, t in enumerate(ts):
                if first_row[run] + k < len(stamps):
                    stamps[first_row[run] + k] = t.strftime('%Y%m%d_%H%M%S%f') if k == 0 else t.strftime('%H%M%S%f')
        table['timestamps'] = stamps
        thr = np.asarray(self.agent_thresholds)
        thr_text = '&'.join('_'.join(str(v) for v in row) for row in thr[:, 1:-1]) if thr.ndim == 2 else '_'.join(str(v) for v in thr)
        props = {'horizon': str(self.horizon), 'reward_discount': str(self.reward_discount), 'distance_metric': self.distance_metric,
                 'used_gpu': str(self.used_gpu), 'environment_name': self.environment.name, 'environment_saved_at': self.environment.saved_at,
                 'environment_dimensions': str(self.environment_dimensions),
                 'environment_shape': '_'.join(str(v) for v in self.environment_shape),
                 'environment_source_position': '_'.join(str(v) for v in self.environment_source_position),
                 'environment_source_radius': str(self.environment_source_radius),
                 'environment_layer_labels': '' if self.environment_layer_labels is None else '&'.join(self.environment_layer_labels),
                 'agent_name': self.agent.name, 'agent_class_name': self.agent.class_name, 'agent_saved_at': self.agent.saved_at,
                 'agent_thresholds': thr_text}
        assert len(table) >= len(props), 'the history is too short to carry its property rows'
        pad = [None] * (len(table) - len(props))
        table['property'] = list(props.keys()) + pad
        table['property_value'] = list(props.values()) + pad
        table.to_csv(folder + file, index=False)
        print(f'Simulations saved to: {folder + file}')
        if save_analysis:
            self.runs_analysis_df.to_csv(folder + file.replace('.csv', '-runs_analysis.csv'))
            self.general_analysis_df.to_csv(folder + file.replace('.csv', '-general_analysis.csv'))
            print(f"Simulation's analysis saved to: {folder + file.replace('.csv', '-[runs|general]_analysis.csv')}")

    def compare(self, *other: 'SimulationHistory'):
        return self.__class__.compare_all(self, *other)

    @classmethod
    def compare_all(cls, *simulation_histories: 'SimulationHistory'):
        'One row per history: mean / std of the run metrics over all runs and over the successful ones (simulation.py:999-1047).'
        import pandas as pd
        rows = []
        for hist in simulation_histories:
            df = hist.general_analysis_df
            row = {'convergence': df.loc['mean', 'converged'], 'convergence_std': df.loc['standard_deviation', 'converged']}
            for col in ('steps_taken', 'discounted_rewards', 'extra_steps', 't_min_over_t'):
                for suffix, line in (('', 'mean'), ('_std', 'standard_deviation'), ('_success', 'success_mean'), ('_success_std', 'success_standard_deviation')):
                    row[col + suffix] = df.loc[line, col]
            rows.append(row)
        return pd.DataFrame(rows)

    @classmethod
    def load_from_file(cls, file: str, environment=False, agent=False) -> 'SimulationHistory':
        return cls.load(file, environment, agent)

    @classmethod
    def load(cls, file: str, environment=False, agent=False) -> 'SimulationHistory':
        '''
        Rebuilds a history from the CSV written by `save` (here or by the reference, simulation.py:752-997; the pre-property
        "legacy" layout is not read).  `environment` / `agent`: an instance, or True to load them from the recorded paths.
        '''
        import pandas as pd
        with open(file, 'r') as f:
            columns = f.readline().strip().split(',')
        dtypes = {c: float for c in columns}
        dtypes.update(time=int, timestamps=str, property=str, property_value=str)
        if 'layer' in columns:
            dtypes['layer'] = float            # empty on the start rows
        table = pd.read_csv(file, dtype=dtypes)
        props = {}
        for name, value in zip(table['property'], table['property_value']):
            if not isinstance(name, str):
                break
            props[name] = value
        horizon = int(props['horizon'])

        def recorded(path) -> bool:
            return isinstance(path, str) and len(path) > 0
        if environment is True:
            assert recorded(props.get('environment_saved_at')), \
                'Environment was not saved at the time of the saving of the simulation history. Input an environment to the environment parameter or toggle the parameter to False.'
            environment = Environment.load(props['environment_saved_at'])
        if agent is True:
            assert recorded(props.get('agent_saved_at')), \
                'Agent was not saved at the time of the saving of the simulation history. Input an agent to the agent parameter or toggle the parameter to False.'
            agent = Agent.load(props['agent_saved_at'])

        data_cols = [c for c in columns if c not in ('property', 'property_value')]
        layered = ((len(data_cols) - 4) % 2) == 1          # time, o, reached_source, timestamps + 2 per axis (+ layer)
        dims = (len(data_cols) - 4) // 2
        num_cols = [c for c in data_cols if c != 'timestamps']
        starts = np.flatnonzero(table['reached_source'].isnull().to_numpy())
        values = table[num_cols].to_numpy(dtype=float)
        runs = np.split(values, starts[1:])
        sizes = np.array([len(r) for r in runs])
        n, steps = len(runs), int(sizes.max()) - 1
        cube = np.full((steps + 1, n, len(num_cols)), -1.0)
        for i, r in enumerate(runs):
            cube[:len(r), i] = r
        a0 = 1 + dims
        a1 = a0 + (1 if layered else 0) + dims

        hist = cls.__new__(cls)
        hist.n = n
        hist.environment = environment if isinstance(environment, Environment) else None
        hist.agent = agent if isinstance(agent, Agent) else None
        hist.time_shift = cube[0, :, 0].astype(int)
        hist.horizon = horizon
        hist.reward_discount = float(props['reward_discount'])
        hist.distance_metric = props['distance_metric']
        hist.used_gpu = props['used_gpu'] == 'True'
        hist.start_points = cube[0, :, 1:a0].astype(int)
        hist.reached_source = np.array([r[-1, -1] == 1 for r in runs])
        hist.done_at_step = np.where((sizes - 1 < horizon) | hist.reached_source, sizes - 1, -1)    # a run that reached the source on the horizon step is done too
        hist._running_sims = np.flatnonzero(hist.done_at_step < 0)
        hist._steps = [None] * steps
        hist._dense = ([cube[k, :, a0:a1].astype(int) for k in range(1, steps + 1)],
                       [cube[k, :, 1:a0].astype(int) for k in range(1, steps + 1)],
                       [cube[k, :, a1] for k in range(1, steps + 1)])
        hist._simulation_dfs = [pd.DataFrame(r, columns=num_cols) for r in runs]

        # step timestamps: a full date on the first row of every batch, time of day afterwards (a smaller value = next day)
        hist.timestamps = {}
        stamp = table['timestamps']
        heads = [k for k, v in enumerate(stamp) if isinstance(v, str) and '_' in v]
        for lo, hi in zip(heads, heads[1:] + [len(stamp)]):
            group = [v for v in stamp[lo:hi] if isinstance(v, str)]
            ts = [datetime.strptime(group[0], '%Y%m%d_%H%M%S%f')]
            day = group[0][:9]
            shift = timedelta(days=0)
            for text in group[1:]:
                t = datetime.strptime(day + text, '%Y%m%d_%H%M%S%f') + shift
                if t < ts[-1]:
                    shift += timedelta(days=1)
                    t += timedelta(days=1)
                ts.append(t)
            hist.timestamps[int(np.flatnonzero(starts == lo)[0])] = ts
        if len(hist.timestamps) == 0:
            print('[Warning] No timestamps were found in the loaded simulation history...')

        hist.environment_dimensions = int(props['environment_dimensions'])
        hist.environment_shape = tuple(int(v) for v in str(props['environment_shape']).split('_'))
        hist.environment_source_position = np.array([float(v) for v in str(props['environment_source_position']).split('_')])
        hist.environment_source_radius = float(props['environment_source_radius'])
        labels = props.get('environment_layer_labels')
        hist.environment_layer_labels = labels.split('&') if isinstance(labels, str) and len(labels) > 0 else None
        text = props['agent_thresholds']
        hist.agent_thresholds = (np.array([np.array(row.split('_')).astype(float) for row in text.split('&')]) if '&' in text
                                 else np.array(text.split('_')).astype(float))
        return hist


def run_test(agent: Agent,
             n: int = None,
             start_points: np.ndarray = None,
             environment: Environment = None,
             time_shift: int | np.ndarray = 0,
             time_loop: bool = True,
             horizon: int = 1000,
             initialization_values: dict = {},
             reward_discount: float = 0.99,
             distance_metric: str = 'l1',
             print_progress: bool = True,
             print_stats: bool = True,
             print_warning: bool = True,
             use_gpu: bool = True,
             parallel_agent_simulation: bool = True,
             batches: int = -1) -> SimulationHistory:
    '''
    Run n simulations of `agent` in its environment (or `environment`), all on the device.
    Signature of the reference's run_test (simulation.py:1224-1240); `use_gpu` is accepted for
    compatibility — this implementation has no CPU path.
    '''
    if n is None:
        n = 1 if (start_points is None or np.asarray(start_points).ndim == 1) else len(start_points)
    if environment is not None:
        if print_warning:
            if environment.shape != agent.environment.shape:
                print("[Warning] The provided environment's shape doesn't match the environment has been trained on...")
            print('Using the provided environment, not the agent environment.')
    else:
        environment = agent.environment

    if start_points is not None:
        start_points = np.asarray(start_points).reshape(-1, environment.dimensions)
        assert start_points.shape == (n, environment.dimensions), f'The provided start_points are of the wrong shape (expected {environment.dimensions}; received {start_points.shape[1]})'
    else:
        start_points = environment.random_start_points(n, agent.rng)

    if isinstance(time_shift, (int, np.integer)):
        time_shift = np.ones(n) * time_shift
    else:
        time_shift = np.array(time_shift)
        assert time_shift.shape == (n,), f'time_shift array has a wrong shape (Given: {time_shift.shape}, expected ({n},))'
    time_shift = time_shift.astype(int)

    if not parallel_agent_simulation:
        batches = n

    common = dict(environment=environment, time_loop=time_loop, horizon=horizon, initialization_values=initialization_values,
                  reward_discount=reward_discount, distance_metric=distance_metric, print_progress=print_progress)
    if batches < 0:
        # grow the number of batches on out-of-memory (the reference keys on MemoryError, simulation.py:1346-1367)
        tries = 2 ** np.arange(max(int(np.log2(max(n, 1))), 0) + 1)
        hist = None
        for try_batches in tries:
            try:
                hist = _run_batches(agent, n, start_points, time_shift, int(try_batches), **common)
                break
            except (torch.cuda.OutOfMemoryError, MemoryError) as e:
                print(f'Memory full: {str(e)[:120]}')
                print('Increasing the amount of batches...')
                torch.cuda.empty_cache()
        if hist is None:
            raise MemoryError('run_test: not enough device memory even with one simulation per batch')
    else:
        hist = _run_batches(agent, n, start_points, time_shift, max(int(batches), 1), **common)

    if print_stats:
        total = (hist.timestamps[max(hist.timestamps)][-1] - hist.timestamps[0][0]).total_seconds()
        print(f'Simulations done in {total:.3f}s:')
        print(hist.summary)
    return hist


def _run_batches(agent, n, start_points, time_shift, batches, **kw) -> SimulationHistory:
    if batches <= 1:
        return _run_single(agent, n, start_points, time_shift, **kw)
    sizes = np.full(batches, n // batches)
    sizes[:n % batches] += 1
    combined, start = None, 0
    for b_n in sizes:
        if b_n == 0:
            continue
        h = _run_single(agent, int(b_n), start_points[start:start + b_n], time_shift[start:start + b_n], **kw)
        start += b_n
        combined = h if combined is None else combined + h
    return combined


_PINNED = {}
_DEBUG_STALL_MS = float(os.environ.get('OLFNAV_DEBUG_STALL_MS', '0') or 0)
_SLEEP_ON_EVENTS = 2 * int(os.environ.get('LOCAL_WORLD_SIZE', os.environ.get('WORLD_SIZE', '1'))) > (os.cpu_count() or 1)


_BUSY_STATE = {'snap': None, 'value': 0.0}


def _host_busy_fraction(window_s: float = 0.03) -> float:
    '''
    Fraction of the host cores that were busy recently (/proc/stat; 0 where that is not available).  The first call in a process
    measures over a short window; later calls use the interval since the previous call, so that a run_test call does not pay for it.
    '''
    def snap():
        with open('/proc/stat') as f:
            v = [float(x) for x in f.readline().split()[1:]]
        return sum(v), v[3] + (v[4] if len(v) > 4 else 0.0)
    try:
        now = snap()
        prev = _BUSY_STATE['snap']
        if prev is None:
            time.sleep(window_s)
            prev, now = now, snap()
        if now[0] - prev[0] >= 16:                       # a meaningful number of clock ticks
            _BUSY_STATE['value'] = max(0.0, 1.0 - (now[1] - prev[1]) / (now[0] - prev[0]))
            _BUSY_STATE['snap'] = now
        elif _BUSY_STATE['snap'] is None:
            _BUSY_STATE['snap'] = prev
        return _BUSY_STATE['value']
    except Exception:
        return 0.0


def _pinned(name: str, nbytes: int) -> torch.Tensor:
    '''
    Page-locked host staging buffers are kept across run_test calls (grown on demand): cudaHostAlloc of a few MB costs milliseconds,
    which is a visible fraction of a short run.  One set per process: run_test is not re-entrant on a device anyway.
    '''
    buf = _PINNED.get(name)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 16), dtype=torch.uint8, pin_memory=True)
        _PINNED[name] = buf
    return buf[:nbytes]


def _run_single(agent, n, start_points, time_shift, environment, time_loop, horizon, initialization_values,
                reward_discount, distance_metric, print_progress, record: bool = True, timing: list = None) -> SimulationHistory:
    '''
    One batch, device resident.  Per step: ONE C-ABI call (olfnav_sim_step: environment step -> compaction plan -> fused
    belief update), the policy of the next step (value-function evaluation / infotaxis choice) enqueued right behind it, and
    one 16-byte read-back of the survivor counts that the host waits for while that policy runs; the O(N) history record of
    the step is copied to pinned host memory on a side stream while the next step runs.
    '''
    from ._lib import SimStepArgs
    from .agents.pbvi_agent import BeliefSimulationMixin
    from .pomdp import current_device
    import ctypes
    # The device environment step bins the odor cue only: a space-aware agent (observation id = odor bin x position cell,
    # reference agent.py:422-434) goes through the reference's host-side protocol loop instead.
    if not isinstance(agent, BeliefSimulationMixin) or getattr(agent, 'space_aware', False):
        return _run_generic(agent, n, start_points, time_shift, environment, time_loop, horizon, initialization_values,
                            reward_discount, distance_metric, print_progress)
    device = current_device()
    dev = torch.device('cuda', device)
    env_h = environment.device_handle(device)
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)

    model = agent.model
    model_h = model.handle(device)
    vf = getattr(agent, 'value_function', None)
    is_infotaxis = not hasattr(agent, 'value_function')
    alphas_h = None if is_infotaxis else vf.handle(device)
    # One-pass step: the belief update of a step also evaluates the value function on the updated rows and delivers the next
    # step's actions (olfnav_belief_step_evaluate: 8 S instead of 12 S bytes per agent-step).  The kernel keeps the rows grouped by
    # action, so the loop carries a row map (brow) next to the beliefs and asks the agent for a few spare rows per buffer.
    # OLFNAV_FUSED_STEP=0 keeps the two kernels (belief step, then evaluate).
    # Default: on when this process is the only rank on its node AND the host is not saturated.  The kernel has an open fault
    # (csrc/fused_step2.cu store issuer): a hardware exception once in a few hundred run_test calls when all eight GPUs
    # of a box run it or every host core is busy, never on a quiet single-GPU box in 3 000+ calls, never with the two kernels in 1 900
    # eight-GPU runs.  Until that is understood, multi-rank jobs and busy hosts take the two kernels; OLFNAV_FUSED_STEP=1 forces the
    # one-pass step, OLFNAV_FUSED_STEP=0 forbids it.
    fused_env = os.environ.get('OLFNAV_FUSED_STEP', '')
    ranks_on_node = int(os.environ.get('LOCAL_WORLD_SIZE', os.environ.get('WORLD_SIZE', '1')) or 1)
    fused = not is_infotaxis and fused_env != '0' and lib.olfnav_can_fuse_policy(model_h, alphas_h) == 1
    if fused:
        quiet = ranks_on_node == 1 and _host_busy_fraction() < 0.5
        if fused_env != '1':
            fused = quiet
        elif not quiet:
            # forced onto a busy or multi-rank host: the kernel variant whose stores complete before a stage is released (no fault
            # seen, 12 % slower; the library reads the variable at every launch)
            os.environ.setdefault('OLFNAV_F2_STORE_MODE', '1')
    # In-place mode: ONE belief buffer, the survivors are updated in the rows they occupy and a row map carries where they are.
    # Chosen when two buffers do not fit the device (10^6 agents x 30 793 states = 127 GB of beliefs on a 180 GB part), or with
    # OLFNAV_INPLACE=1; it takes the two-kernel step (the one-pass kernel writes the rows somewhere else by construction) and the
    # policy runs over all rows, dead ones included, so it pays off only when memory forces it.
    row_bytes = model.padded_states * 4
    # free device memory + what torch's caching allocator holds without using it (the buffers of a previous run)
    free_bytes = torch.cuda.mem_get_info(dev)[0] + torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev)
    inplace_env = os.environ.get('OLFNAV_INPLACE', '')
    inplace = (inplace_env == '1') or (inplace_env != '0' and 2 * n * row_bytes > 0.92 * free_bytes >= n * row_bytes)
    if inplace:
        fused = False
    agent._row_capacity = int(lib.olfnav_fused_out_rows(model_h, n)) if fused else 0
    agent._single_buffer = inplace
    try:
        agent.initialize_state(n, **initialization_values)
    finally:
        agent._row_capacity = 0
        agent._single_buffer = False
    hist = SimulationHistory(start_points=start_points, environment=environment, agent=agent, time_shift=time_shift,
                             horizon=horizon, reward_discount=reward_discount, distance_metric=distance_metric, used_gpu=True)

    # host -> device, once
    cap = max(n, 1)
    pos = [torch.as_tensor(np.ascontiguousarray(start_points), dtype=torch.int32, device=dev).contiguous().reshape(cap, 2) if n else torch.zeros((1, 2), dtype=torch.int32, device=dev),
           torch.empty((cap, 2), dtype=torch.int32, device=dev)]
    ts = [torch.as_tensor(np.ascontiguousarray(time_shift), dtype=torch.int64, device=dev).contiguous() if n else torch.zeros(1, dtype=torch.int64, device=dev),
          torch.empty(cap, dtype=torch.int64, device=dev)]
    thresholds = torch.as_tensor(agent.thresholds, dtype=torch.float64, device=dev).contiguous()
    moves = torch.as_tensor(np.ascontiguousarray(agent.movement_set), dtype=torch.int32, device=dev).contiguous()
    layered = bool(environment.has_layers)
    act_layer = torch.as_tensor(np.ascontiguousarray(agent.action_layers), dtype=torch.int32, device=dev).contiguous() if layered else None
    hist_layer = torch.zeros(len(environment.layers) + 1, dtype=torch.int32, device=dev) if layered else None
    T = environment.timesteps

    def i32(k=cap):
        return torch.empty(k, dtype=torch.int32, device=dev)

    def u8(k=cap):
        return torch.empty(k, dtype=torch.uint8, device=dev)
    act, obs_id, src_rows, act_c, obs_c = i32(), i32(), i32(), i32(), i32()
    ok, at_end, term_c = u8(), u8(), u8()
    hist_t = i32(T + 1)
    eval_val, eval_idx = torch.empty(cap, dtype=torch.float32, device=dev), i32()
    counts = torch.zeros(4, dtype=torch.int32, device=dev)
    counts_host = _pinned('counts', 16).view(torch.int32)
    counts_host.zero_()
    # two record sets so that step i+1 can run while step i's record is still being copied out
    # Record sets: ONE contiguous device buffer each (odor f64 | pos 2 x i32 | act i32 | reached u8 | term u8, 22 bytes per agent)
    # so that a step's record leaves the GPU as a single copy; two device sets (step i + 1 runs while step i's record is copied
    # out) and three pinned host sets allocated once (the host consumes records two steps late).
    def record_views(buf):
        o = 0
        v = {}
        for name, dt, shape in (('odor', torch.float64, (cap,)), ('pos', torch.int32, (cap, 2)), ('act', torch.int32, (cap,)),
                                ('reached', torch.uint8, (cap,)), ('term', torch.uint8, (cap,))):
            nbytes = int(np.prod(shape)) * torch.empty(0, dtype=dt).element_size()
            v[name] = buf[o:o + nbytes].view(dt).view(shape)
            o += nbytes
        v['_raw'] = buf
        return v
    rec_bytes = 22 * cap
    rec_dev = [torch.empty(rec_bytes, dtype=torch.uint8, device=dev) for _ in range(2)]
    recs = [record_views(b) for b in rec_dev]
    rec_host = [_pinned(f'rec{k}', rec_bytes) for k in range(3)] if record else []
    rec_host_views = [record_views(b) for b in rec_host]
    rec_free = [None, None]            # event: the record set has been copied out
    pending = []

    def flush(upto: int) -> None:
        # One copy of the step's 22-byte records out of the pinned staging set (it is reused two steps later); the history keeps
        # typed views of that copy (int32 ids / positions, uint8 flags) and widens them lazily when a table is asked for.  Five
        # astype / copy passes over 100 000 rows per step and rank used to sit here, on host cores that 8 ranks share.
        while len(pending) > upto:
            ev, h, m_rec = pending.pop(0)
            ev.synchronize()
            own = record_views(h['_raw'].clone())
            hist.add_step(actions=own['act'].numpy()[:m_rec], next_positions=own['pos'].numpy()[:m_rec], observations=own['odor'].numpy()[:m_rec],
                          reached_source=own['reached'].numpy()[:m_rec].view(np.bool_), interupt=own['term'].numpy()[:m_rec].view(np.bool_),
                          action_is_id=True)

    def run_policy(b, scale, rows: int, act_out, row_map=None) -> None:
        'argmax_g b.alpha_g -> action id (agents/pbvi_agent.py:741) or the infotaxis choice (agents/infotaxis_agent.py:244-301).'
        if inplace:
            # one buffer: the policy runs over every row the run started with (dead rows included: their results are never read),
            # then the actions of the live agents are gathered through the row map
            run_policy_rows(b, scale, n, act_by_row)
            if row_map is None:
                act_out[:rows].copy_(act_by_row[:rows])
            else:
                torch.index_select(act_by_row, 0, row_map[:rows].long(), out=act_out[:rows])
            return
        run_policy_rows(b, scale, rows, act_out)

    def run_policy_rows(b, scale, rows: int, act_out) -> None:
        with torch.cuda.device(device):
            if is_infotaxis:
                check(lib.olfnav_infotaxis_choose(model_h, b.data_ptr(), b.stride(0), rows, scale.data_ptr(), act_out.data_ptr(), None, stream()))
            else:
                check(lib.olfnav_evaluate(model_h, alphas_h, b.data_ptr(), b.stride(0), rows, scale.data_ptr(), None, None, act_out.data_ptr(),
                                          int(getattr(agent, 'gemm_path', 0)), stream()))

    args = SimStepArgs()
    hist._step_mode = 'in_place' if inplace else ('one_pass' if fused else 'two_kernel')
    brow = [torch.empty(cap, dtype=torch.int32, device=dev) for _ in range(2)] if (fused or inplace) else None
    act_by_row = torch.empty(cap, dtype=torch.int32, device=dev) if inplace else None       # in-place mode: policy output per belief ROW
    act_next_buf = [torch.empty(cap, dtype=torch.int32, device=dev) for _ in range(2)] if fused else None

    # Running ahead.  The size of step i + 1 (the survivors of step i) is final as soon as step i's compaction plan has run, i.e. BEFORE
    # its belief update starts: the library records an event there (ev_plan_done), a side stream copies counts[0] out behind it, and the
    # host enqueues step i + 1 while step i is still running — the GPU never waits for the host between two steps.  What is not known
    # yet at that point is whether step i will flag failed updates (normaliser 0) or agents at the end of the recording: step i + 1 is
    # launched on the assumption that it does not, and DISCARDED when it does (it only writes buffers that are two steps old: the
    # state after step i is intact), the explicit compaction runs and the step is launched again.  Not in in-place mode (a step
    # overwrites its own input) and not with time_loop = False (agents reach the end of the recording at every step near the end).
    # The per-step flags that a discarded step would overwrite are double-buffered (counts, term_c).
    ahead = os.environ.get('OLFNAV_RUN_AHEAD', '1') != '0' and bool(time_loop) and not inplace
    counts2 = [counts, torch.zeros(4, dtype=torch.int32, device=dev)]
    counts_host2 = [counts_host, _pinned('counts1', 16).view(torch.int32)]
    term_c2 = [term_c, u8()]
    early_host = [_pinned(f'early{k}', 16).view(torch.int32) for k in range(2)]
    side = torch.cuda.Stream(device=dev) if ahead else None
    plan_events = []
    for _ in range(2 if ahead else 0):
        e = torch.cuda.Event()
        e.record(main)                                             # creates the underlying cudaEvent_t handle
        plan_events.append(e)

    S = dict(cur=0, brow_valid=False, n_live=n)     # brow_valid False: agent i's belief is row i (start of the run, or after an explicit compaction)

    def launch(i: int, n_in: int, have_act: bool) -> dict:
        'Enqueue step i for the n_in live agents of the current state (no host synchronisation).'
        if _DEBUG_STALL_MS > 0 and (i * 2654435761 >> 7) % 3 == 0:
            time.sleep(_DEBUG_STALL_MS * 1e-3)          # diagnostics: a host that stalls between two steps
        cur, brow_valid = S['cur'], S['brow_valid']
        r = recs[i & 1]
        if rec_free[i & 1] is not None:
            main.wait_event(rec_free[i & 1])
        b_cur, b_nxt = agent._bufs[agent._cur], agent._bufs[1 - agent._cur]
        s_cur, s_nxt = agent._scales[agent._cur], agent._scales[1 - agent._cur]
        args.b_in, args.b_out, args.ld = b_cur.data_ptr(), b_nxt.data_ptr(), b_cur.stride(0)
        args.scale_in, args.scale_out = s_cur.data_ptr(), s_nxt.data_ptr()
        args.pos_in, args.pos_out = pos[cur].data_ptr(), pos[1 - cur].data_ptr()
        args.ts_in, args.ts_out = ts[cur].data_ptr(), ts[1 - cur].data_ptr()
        args.n, args.step = n_in, i
        args.moves, args.thresholds, args.n_thr = moves.data_ptr(), thresholds.data_ptr(), int(thresholds.shape[-1])
        args.act_layer = act_layer.data_ptr() if layered else None
        args.hist_layer = hist_layer.data_ptr() if layered else None
        args.thr_per_layer = 1 if thresholds.dim() == 2 else 0
        args.time_mode = 1 if environment.reference_time_quirk else 0
        args.time_loop = 1 if time_loop else 0
        args.eval_path = int(getattr(agent, 'gemm_path', 0))
        args.have_act = 1 if have_act else 0
        args.act, args.obs_id, args.src_rows, args.act_c, args.obs_c = r['act'].data_ptr(), obs_id.data_ptr(), src_rows.data_ptr(), act_c.data_ptr(), obs_c.data_ptr()
        args.ok, args.at_end, args.term_c = ok.data_ptr(), at_end.data_ptr(), term_c2[i & 1].data_ptr()
        args.hist, args.eval_val, args.eval_idx = hist_t.data_ptr(), eval_val.data_ptr(), eval_idx.data_ptr()
        args.rec_pos, args.rec_odor, args.rec_reached, args.rec_term = r['pos'].data_ptr(), r['odor'].data_ptr(), r['reached'].data_ptr(), r['term'].data_ptr()
        args.counts = counts2[i & 1].data_ptr()
        args.inplace = 1 if inplace else 0
        if inplace:
            if not have_act:
                run_policy(b_cur, s_cur, n_in, r['act'], brow[cur] if brow_valid else None)      # first step: the caller's policy (see olfnav_sim_step)
                args.have_act = 1
            args.act_next = None
            args.brow_in = brow[cur].data_ptr() if brow_valid else None
            args.brow_out = brow[1 - cur].data_ptr()
            args.belief_rows = int(b_cur.shape[0])
        elif fused:
            # The update of this step also delivers the NEXT step's actions (also on the last step of the horizon: the policy rides
            # along for free).  They go to a buffer of their own and move into the record set at the start of the step that plays
            # them: the record set of step i + 1 is still on its way to the host while step i runs.
            if have_act:
                r['act'][:n_in].copy_(act_next_buf[i & 1][:n_in], non_blocking=True)
            args.act_next = act_next_buf[(i + 1) & 1].data_ptr()
            args.brow_in = brow[cur].data_ptr() if brow_valid else None
            args.brow_out = brow[1 - cur].data_ptr()
            args.belief_rows = int(b_cur.shape[0])
        else:
            args.act_next = None
        evs = None
        if timing is not None:
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            for e in evs:
                e.record(main)                                  # creates the underlying cudaEvent_t handles
            args.ev_policy_end, args.ev_update_begin, args.ev_update_end = evs[1].cuda_event, evs[2].cuda_event, evs[3].cuda_event
            evs[0].record(main)
        run_ahead = ahead and i + 1 < horizon
        args.ev_plan_done = plan_events[i & 1].cuda_event if run_ahead else None
        with torch.cuda.device(device):
            check(lib.olfnav_sim_step(model_h, alphas_h, env_h, ctypes.byref(args), stream()))
        counts_host2[i & 1].copy_(counts2[i & 1], non_blocking=True)
        # the host wait of the step: sleep on it when the ranks of a node outnumber half the host cores (8 ranks share the cores of a
        # node), spin otherwise (a blocking wake-up costs tens of microseconds)
        step_done = torch.cuda.Event(blocking=_SLEEP_ON_EVENTS)
        step_done.record(main)
        if timing is not None:
            evs[4].record(main)
        early = None
        if run_ahead:
            with torch.cuda.stream(side):
                side.wait_event(plan_events[i & 1])
                early_host[i & 1][:1].copy_(counts2[i & 1][:1], non_blocking=True)
                early = torch.cuda.Event(blocking=_SLEEP_ON_EVENTS)
                early.record(side)

        # Two-kernel modes: the policy of the NEXT step, enqueued right behind this step: it runs over the n_in rows of the updated
        # buffer (an upper bound of the survivors; rows are independent, the surplus results are never read).
        if i + 1 < horizon:
            nxt = recs[(i + 1) & 1]
            if not fused and rec_free[(i + 1) & 1] is not None:
                main.wait_event(rec_free[(i + 1) & 1])
            if timing is not None:
                p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                p0.record(main)
            if not fused:
                run_policy(b_nxt, s_nxt, n_in, nxt['act'], brow[1 - cur] if inplace else None)
            if timing is not None:
                p1.record(main)
                evs += [p0, p1]

        # record of this step -> pinned host memory on the side stream; the host consumes older records here, with the step enqueued
        if record:
            flush(2)                                             # the host set about to be overwritten has been consumed
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(step_done)
                rec_host[i % 3].copy_(rec_dev[i & 1], non_blocking=True)
                copied = torch.cuda.Event(blocking=True)
                copied.record(copy_stream)
            rec_free[i & 1] = copied
            pending.append((copied, rec_host_views[i % 3], n_in))
        hist._agent_steps = getattr(hist, '_agent_steps', 0) + n_in
        return dict(i=i, n=n_in, step_done=step_done, early=early, evs=evs)

    def advance(m: int) -> None:
        'The host-side state after a step that left m survivors.'
        agent._cur, agent._n_live, agent._action_ids, agent._dropped = 1 - agent._cur, m, None, None
        S['cur'] = 1 - S['cur']
        S['brow_valid'] = fused or inplace
        S['n_live'] = m

    bar = None
    if print_progress:
        try:
            from tqdm.auto import tqdm
            bar = tqdm(total=horizon)
        except Exception:
            pass

    info = launch(0, n, False) if (horizon > 0 and n > 0) else None
    while info is not None:
        i, nxt_info = info['i'], None
        if info['early'] is not None:
            # the survivors of step i are known while its belief update is still running: enqueue step i + 1 behind it
            info['early'].synchronize()
            m = int(early_host[i & 1][0])
            advance(m)
            if m > 0:
                nxt_info = launch(i + 1, m, True)
        info['step_done'].synchronize()
        ch = counts_host2[i & 1]
        m_done, n_fail, n_end = int(ch[0]), int(ch[1]), int(ch[2])
        if info['early'] is None:
            m = m_done
            advance(m)
        if timing is not None:
            timing.append((info['n'], info['evs'], m))
        if n_fail or n_end:
            # rare: some survivors must still go (impossible update / end of the recording): explicit compaction of the state after
            # step i.  A step that was enqueued ahead is discarded first (its outputs are overwritten by the launch that follows).
            if nxt_info is not None:
                if record:
                    pending.pop()
                hist._agent_steps -= nxt_info['n']
                nxt_info = None
            cur = S['cur']
            have_next = i + 1 < horizon
            if fused and S['brow_valid']:
                agent._row_map = brow[cur]
                agent._rows_to_agent_order()
                S['brow_valid'] = False
            keep = ~term_c2[i & 1][:m].bool()
            rows = torch.nonzero(keep).flatten()
            k = int(rows.numel())
            if k:
                pos[1 - cur][:k] = pos[cur][:m][rows]
                ts[1 - cur][:k] = ts[cur][:m][rows]
                if inplace:
                    brow[1 - cur][:k] = brow[cur][:m][rows]       # the belief rows stay where they are: only the row map is compacted
                if have_next:
                    nact = act_next_buf[(i + 1) & 1] if fused else recs[(i + 1) & 1]['act']
                    nact[:k] = nact[:m][rows]
            S['cur'] = 1 - cur
            if inplace:
                agent._n_live = k
            else:
                agent.kill_device(~keep)
            m = k
            S['n_live'] = m
        if bar is not None:
            bar.update(1)
            bar.set_postfix({'done ': f' {n - m}/{n}'})
        if nxt_info is None and i + 1 < horizon and m > 0:
            nxt_info = launch(i + 1, m, True)
        info = nxt_info
    if bar is not None:
        bar.close()
    flush(0)
    n_live, cur = S['n_live'], S['cur']
    if n_live == 0:
        agent._bufs = None
    elif fused and S['brow_valid']:
        agent._row_map = brow[cur][:n_live].clone()      # rows still grouped by action: put back in agent order on first use
    elif inplace:
        agent._bufs = None                               # one buffer, rows scattered: the state is not kept after an in-place run
        agent._n_live = 0
    return hist


class _null_context:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def _run_generic(agent, n, start_points, time_shift, environment, time_loop, horizon, initialization_values,
                 reward_discount, distance_metric, print_progress) -> SimulationHistory:
    '''
    Host-driven loop for agents that only implement the reference's NumPy protocol (custom Agent subclasses):
    the literal step order of the reference's run_test (simulation.py:1452-1497).
    '''
    pos, ts = np.array(start_points), np.array(time_shift)
    agent.initialize_state(n, **initialization_values)
    hist = SimulationHistory(start_points=start_points, environment=environment, agent=agent, time_shift=time_shift,
                             horizon=horizon, reward_discount=reward_discount, distance_metric=distance_metric, used_gpu=False)
    for i in range(horizon):
        action = agent.choose_action()
        pos = environment.move(pos=pos, movement=(action if not environment.has_layers else action[:, 1:]))
        observation = environment.get_observation(pos=pos, time=(ts + i), layer=(0 if not environment.has_layers else action[:, 0]))
        reached = environment.source_reached(pos)
        if agent.space_aware:
            observation = np.hstack((observation[:, None], pos))
        ok = agent.update_state(action=action, observation=observation, source_reached=reached)
        if ok is None:
            ok = np.ones(len(reached), dtype=bool)
        at_end = (ts + i + 1) >= (math.inf if time_loop else environment.timesteps)
        term = reached | at_end | ~np.asarray(ok, dtype=bool)
        # the history records the odor cue only (reference simulation.py:1489), not the position columns of a space-aware observation
        hist.add_step(actions=action, next_positions=pos, observations=(observation[:, 0] if agent.space_aware else observation),
                      reached_source=reached, interupt=term)
        pos, ts = pos[~term], ts[~term]
        agent.kill(simulations_to_kill=term)
        if len(pos) == 0:
            break
    return hist
