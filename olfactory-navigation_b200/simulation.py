'''
run_test / SimulationHistory — device-resident version of the reference's simulation driver
(reference olfactory_navigation/simulation.py:1224-1519 and :34-227).

Same signature, same step order (act on b_t -> move -> observe at the new cell -> update to
b_{t+1} -> kill; simulation.py:1452-1497) and the same recorded quantities, but the whole step
stays on the GPU: positions, time shifts and beliefs never leave HBM; per step only the O(N) record
(action id, position, odor, flags) is copied to pinned host memory, asynchronously, on a side
stream.  Host <-> device boundary: start points / time shifts in (once), records out (per step).

Analysis DataFrames, CSV save/load and plots of the reference's SimulationHistory are host-side
bookkeeping outside the hot path and are not rebuilt here.
'''
from __future__ import annotations

import math
from datetime import datetime

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
        steps = self.steps_taken.astype(float)
        ok = self.reached_source
        n_ok = int(np.sum(ok))
        fails = self.n - n_ok
        s = f'Simulations reached goal: {n_ok}/{self.n} ({fails} failures (reached horizon: {int(np.sum(self.simulations_at_horizon))})) ({100.0 * n_ok / self.n:.2f}% success)\n'
        if self.error_count:
            s += f' - Simulations with errors: {self.error_count}\n'
        dist = np.maximum(self.compute_distance_to_source(), 0.0)
        adr = np.where(ok, self.reward_discount ** steps, 0.0)
        with np.errstate(divide='ignore', invalid='ignore'):
            ratio = np.where(steps > 0, dist / steps, 0.0)

        def ms(x, sel=None):
            x = x if sel is None else x[sel]
            return (float(np.mean(x)), float(np.std(x))) if len(x) else (float('nan'), float('nan'))
        for label, arr in (('Average step count:', steps), ('Extra steps:', steps - dist),
                           ('Average discounted rewards (ADR):', adr), ('Tmin/T:', ratio)):
            a, b = ms(arr), ms(arr, ok)
            s += f' - {label:<38}{a[0]:.3f} +- {a[1]:.2f} (Successful only: {b[0]:.3f} +- {b[1]:.2f})\n'
        return s

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
        out.timestamps = {0: self.timestamps[0], 1: other.timestamps[0]}
        return out


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


def _run_single(agent, n, start_points, time_shift, environment, time_loop, horizon, initialization_values,
                reward_discount, distance_metric, print_progress) -> SimulationHistory:
    from .pomdp import current_device
    device = current_device()
    dev = torch.device('cuda', device)
    env_h = environment.device_handle(device)
    copy_stream = torch.cuda.Stream(device=dev)

    # host -> device, once
    pos = torch.as_tensor(np.ascontiguousarray(start_points), dtype=torch.int32, device=dev).contiguous()
    tshift = torch.as_tensor(np.ascontiguousarray(time_shift), dtype=torch.int64, device=dev).contiguous()
    thresholds = torch.as_tensor(agent.thresholds, dtype=torch.float64, device=dev).contiguous()
    n_thr = int(thresholds.numel())
    moves = torch.as_tensor(np.ascontiguousarray(agent.action_set), dtype=torch.int32, device=dev).contiguous()

    agent.initialize_state(n, **initialization_values)
    hist = SimulationHistory(start_points=start_points, environment=environment, agent=agent, time_shift=time_shift,
                             horizon=horizon, reward_discount=reward_discount, distance_metric=distance_metric, used_gpu=True)

    odor = torch.empty(n, dtype=torch.float64, device=dev)
    obs_id = torch.empty(n, dtype=torch.int32, device=dev)
    reached_u8 = torch.empty(n, dtype=torch.uint8, device=dev)
    pending = []        # (event, pinned buffers) of records still in flight
    T = environment.timesteps

    def flush(upto: int) -> None:
        while len(pending) > upto:
            ev, (h_act, h_pos, h_odor, h_flags), m = pending.pop(0)
            ev.synchronize()
            fl = h_flags.numpy()
            hist.add_step(actions=h_act.numpy()[:m].astype(np.int64), next_positions=h_pos.numpy()[:m].astype(np.int64),
                          observations=h_odor.numpy()[:m], reached_source=fl[0, :m].astype(bool), interupt=fl[1, :m].astype(bool),
                          action_is_id=True)

    n_live = n
    iterator = range(horizon)
    if print_progress:
        try:
            from tqdm.auto import trange
            iterator = trange(horizon)
        except Exception:
            pass
    for i in iterator:
        act_ids = agent.choose_action_device()                                   # (n_live,) int32
        if environment.reference_time_quirk:
            # agent i observes the frame of the i-th smallest (time_shift + i) % T (see Environment.reference_time_quirk)
            t_eff, t_step = torch.sort((tshift[:n_live] + i) % T).values.contiguous(), 0
        else:
            t_eff, t_step = tshift, i
        with torch.cuda.device(device):
            check(lib.olfnav_env_step(env_h, n_live, ptr(pos), ptr(act_ids), ptr(moves), ptr(t_eff), t_step,
                                      ptr(thresholds), n_thr, ptr(odor), ptr(obs_id), ptr(reached_u8), stream()))
        reached = reached_u8[:n_live].bool()
        ok = agent.update_state_device(obs_id[:n_live], reached)
        if time_loop:
            term = reached | ~ok
        else:
            term = reached | ~ok | ((tshift + (i + 1)) >= T)

        # record (device -> pinned host, asynchronously on the side stream)
        flags = torch.stack([reached, term]).to(torch.uint8)
        bufs = (torch.empty(n_live, dtype=torch.int32, pin_memory=True), torch.empty((n_live, 2), dtype=torch.int32, pin_memory=True),
                torch.empty(n_live, dtype=torch.float64, pin_memory=True), torch.empty((2, n_live), dtype=torch.uint8, pin_memory=True))
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(dev))
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ready)
            bufs[0].copy_(act_ids[:n_live], non_blocking=True)
            bufs[1].copy_(pos[:n_live], non_blocking=True)
            bufs[2].copy_(odor[:n_live], non_blocking=True)
            bufs[3].copy_(flags, non_blocking=True)
            for t in (act_ids, flags):
                t.record_stream(copy_stream)
            done = torch.cuda.Event()
            done.record(copy_stream)
        pending.append((done, bufs, n_live))

        keep = ~term
        new_pos = pos[:n_live][keep].contiguous()
        new_ts = tshift[:n_live][keep].contiguous()
        agent.kill_device(term)
        # pos/odor are overwritten by the next step: the record copy must have read them first
        torch.cuda.current_stream(dev).wait_event(done)
        n_live = int(new_pos.shape[0])           # the one host sync of the step
        pos[:n_live] = new_pos
        tshift[:n_live] = new_ts
        flush(2)
        if n_live == 0:
            break
        if print_progress and hasattr(iterator, 'set_postfix'):
            iterator.set_postfix({'done ': f' {n - n_live}/{n}'})
    flush(0)
    return hist
