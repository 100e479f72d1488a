'''
PBVI_Agent — drop-in for the reference's `olfactory_navigation.agents.PBVI_Agent`
(reference olfactory_navigation/agents/pbvi_agent.py:237-874): same constructor, attributes and
`initialize_state / choose_action / update_state / kill` protocol (:698-811), with the belief state
resident in HBM and every O(N*S) operation running in the CUDA library.

Simulation state layout: two (capacity, Sp) float32 belief buffers used as ping-pong targets of the
fused belief step.  Agents that reached the source are dropped by the step itself (the kernel
writes the survivors to rows 0..n-1 of the other buffer), which replaces the reference's
`belief_array[~kill]` gather-copy (:810-811).  `kill()` only copies rows in the rare case where
something else dies (failed update, end of the odor recording).
'''
from __future__ import annotations

import json
import os
import shutil
from datetime import datetime

import numpy as np
import torch

from .._lib import lib, ptr, stream, check
from ..agent import Agent
from ..converter import exact_converter
from ..environment import Environment
from ..pomdp import POMDP, Belief, BeliefSet, ValueFunction, current_device


class BeliefSimulationMixin:
    'Device-resident belief state shared by every belief-based agent (PBVI family, QMDP, Infotaxis).'

    def _init_belief_state(self, n: int, belief: BeliefSet = None) -> None:
        device = current_device()
        dev = torch.device('cuda', device)
        model: POMDP = self.model
        Sp = model.padded_states
        self._device = device
        # release the buffers of a previous run BEFORE allocating the new ones: two belief buffers of a C5 shard are 2 x 61.6 GB
        # of a 180 GB part, a third one does not fit  (a `belief` taken from this agent keeps its own reference alive)
        self._bufs, self._scales = None, None
        # run_test's one-pass update + policy kernel writes the rows grouped by action, every action's last 128-row tile padded:
        # it asks for a few spare rows per buffer (olfnav_fused_out_rows) through _row_capacity
        cap = max(n, 1)
        rows = max(cap, int(getattr(self, '_row_capacity', 0) or 0))
        b0 = torch.empty((rows, Sp), dtype=torch.float32, device=dev)
        s0 = torch.zeros(rows, dtype=torch.float32, device=dev)
        if belief is None:
            if n:
                with torch.cuda.device(device):
                    check(lib.olfnav_belief_init(model.handle(device), ptr(b0), Sp, n, ptr(s0), stream()))
        else:
            assert len(belief) == n, f'The amount of beliefs provided ({len(belief)}) to initialize the state need to match the amount of stimulations to initialize (n={n}).'
            src = belief if isinstance(belief, BeliefSet) else BeliefSet(model, belief.belief_array)
            if src._b.shape[1] != Sp:
                raise ValueError('belief set buffers have an unexpected shape')
            b0[:n].copy_(src._b[:n])                   # never clobber the caller's set
            s0[:n].copy_(src._scale[:n])
        first = BeliefSet(model, _buffers=(b0, s0, n), device=device)
        if getattr(self, '_single_buffer', False):
            # run_test's in-place mode (10^6 agents x 30 793 states = 127 GB: a second buffer does not fit one B200): both "sides" of
            # the ping-pong are the same tensor, rows never move
            self._bufs = [first._b, first._b]
            self._scales = [first._scale, first._scale]
        else:
            self._bufs = [first._b, torch.empty_like(first._b)]
            self._scales = [first._scale, torch.zeros_like(first._scale)]
        self._cur = 0
        self._n_live = n
        self._ok = torch.empty(cap, dtype=torch.uint8, device=dev)
        self._row_map = None             # set by run_test while the rows are grouped by action (see _rows_to_agent_order)
        self._action_ids = None          # device int32 (n_live,), set by choose_action
        self._dropped = None             # bool (n,) rows the last update dropped (source reached)
        self._moves_dev = torch.as_tensor(np.ascontiguousarray(self.action_set), dtype=torch.int32, device=dev).contiguous()
        self.action_played = None
        self.succeeded_update = None

    def _rows_to_agent_order(self) -> None:
        '''
        run_test's one-pass step kernel leaves the live belief rows grouped by action, with a row map (agent i -> row): gather them
        back into agent order (row i = agent i) the first time anything else looks at the state.
        '''
        row_map = getattr(self, '_row_map', None)
        if row_map is None or getattr(self, '_bufs', None) is None:
            self._row_map = None
            return
        self._row_map = None
        m = self._n_live
        if m:
            src, dst = self._cur, 1 - self._cur
            with torch.cuda.device(self._device):
                check(lib.olfnav_copy_rows(ptr(self._bufs[src]), ptr(self._bufs[dst]), self._bufs[src].stride(0), self.model.padded_states,
                                           ptr(row_map), ptr(self._scales[src]), ptr(self._scales[dst]), m, stream()))
            self._cur = dst

    # -- views ----------------------------------------------------------------------------------
    @property
    def belief(self) -> BeliefSet | None:
        self._rows_to_agent_order()
        if getattr(self, '_bufs', None) is None or self._n_live == 0:
            return None
        return BeliefSet(self.model, _buffers=(self._bufs[self._cur], self._scales[self._cur], self._n_live), device=self._device)

    @belief.setter
    def belief(self, value) -> None:
        if value is None:
            self._bufs = None
            self._n_live = 0
        else:
            self._init_belief_state(len(value), value)

    # -- device-side protocol (used by olfactory_navigation_b200.run_test) ------------------------
    def update_state_device(self, obs_ids: torch.Tensor, reached: torch.Tensor) -> torch.Tensor:
        '''
        obs_ids int32 (n,), reached bool (n,) on the device.  Updates the beliefs of the agents that
        did not reach the source, compacting them to rows 0..m-1.  Returns ok (n,) bool: False where
        the update was impossible (normaliser 0).  Agents with reached=True are dropped and MUST be
        killed by the caller (the reference's run_test always does: simulation.py:1480-1497).
        '''
        self._rows_to_agent_order()
        n = self._n_live
        assert self._action_ids is not None and len(self._action_ids) == n, 'choose_action must be called before update_state'
        dev = obs_ids.device
        keep_rows = torch.nonzero(~reached).flatten().to(torch.int32)
        m = int(keep_rows.numel())
        ok_full = torch.ones(n, dtype=torch.bool, device=dev)
        if m:
            act_c = self._action_ids[keep_rows.long()].contiguous()
            obs_c = obs_ids[keep_rows.long()].to(torch.int32).contiguous()
            src, dst = self._cur, 1 - self._cur
            with torch.cuda.device(self._device):
                check(lib.olfnav_belief_step(self.model.handle(self._device), ptr(self._bufs[src]), ptr(self._bufs[dst]),
                                             self._bufs[src].stride(0), m, ptr(act_c), ptr(obs_c),
                                             None if m == n else ptr(keep_rows), None,
                                             ptr(self._scales[src]), ptr(self._scales[dst]), ptr(self._ok), stream()))
            self._cur = dst
            ok_full[keep_rows.long()] = self._ok[:m].bool()
        self._n_live = m
        self._dropped = reached
        self._action_ids = None
        return ok_full

    def kill_device(self, to_kill: torch.Tensor) -> None:
        'to_kill bool (n,) in the indexing used by the last update_state_device call.'
        self._rows_to_agent_order()
        dropped = self._dropped if self._dropped is not None else torch.zeros_like(to_kill)
        self._dropped = None
        if bool((dropped & ~to_kill).any()):
            raise RuntimeError('an agent that reached the source was not killed: its belief was dropped by update_state')
        extra = to_kill[~dropped]                       # among the m compacted rows
        if extra.numel() and bool(extra.any()):
            rows = torch.nonzero(~extra).flatten().to(torch.int32)
            k = int(rows.numel())
            if k:
                src, dst = self._cur, 1 - self._cur
                with torch.cuda.device(self._device):
                    check(lib.olfnav_copy_rows(ptr(self._bufs[src]), ptr(self._bufs[dst]), self._bufs[src].stride(0),
                                               self.model.padded_states, ptr(rows), ptr(self._scales[src]),
                                               ptr(self._scales[dst]), k, stream()))
                self._cur = dst
            self._n_live = k

    # -- reference protocol (NumPy in / out) ---------------------------------------------------------
    def update_state(self, action: np.ndarray, observation: np.ndarray, source_reached: np.ndarray):
        assert getattr(self, '_bufs', None) is not None, 'Agent was not initialized yet, run the initialize_state function first'
        dev = self._bufs[0].device
        ids = self.discretize_observations(observation=np.asarray(observation), action=action, source_reached=np.asarray(source_reached))
        ok = self.update_state_device(torch.as_tensor(ids, dtype=torch.int32, device=dev),
                                      torch.as_tensor(np.asarray(source_reached, dtype=bool), device=dev))
        ok = ok.cpu().numpy()
        self.succeeded_update = ok
        return ok

    def kill(self, simulations_to_kill: np.ndarray) -> None:
        if getattr(self, '_bufs', None) is None:
            return
        mask = torch.as_tensor(np.asarray(simulations_to_kill, dtype=bool), device=self._bufs[0].device)
        self.kill_device(mask)
        if self._n_live == 0:
            self._bufs = None


class PBVI_Agent(BeliefSimulationMixin, Agent):
    def __init__(self,
                 environment: Environment,
                 thresholds: float | list[float] = 3e-6,
                 space_aware: bool = False,
                 spacial_subdivisions: np.ndarray = None,
                 actions: dict | np.ndarray = None,
                 name: str = None,
                 model: POMDP = None,
                 use_reachability: bool = False,
                 environment_converter=None,
                 rng=None,
                 **converter_parameters) -> None:
        # `rng` is accepted and forwarded (the reference drops it); stray kwargs such
        # as `spatial_subdivisions` are swallowed like the reference does (pbvi_agent.py:341).
        super().__init__(environment=environment, thresholds=thresholds, space_aware=space_aware,
                         spacial_subdivisions=spacial_subdivisions, actions=actions, name=name, rng=rng)
        if model is not None:
            self.model = model
        elif callable(environment_converter):
            self.model = environment_converter(agent=self, **converter_parameters)
        else:
            self.model = exact_converter(agent=self)
        self.use_reachability = use_reachability
        self.trained_at = None
        self.value_function: ValueFunction = None
        self._bufs = None
        self._n_live = 0
        self.action_played = None
        self.succeeded_update = None
        self.gemm_path = 0      # 0 auto, 1 FP32 SIMT, 2 tcgen05 (see olfnav_evaluate)

    # -- simulation (pbvi_agent.py:698-811) ---------------------------------------------------------
    def initialize_state(self, n: int = 1, belief: BeliefSet = None) -> None:
        assert self.value_function is not None, 'Agent was not trained, run the training function first...'
        self._init_belief_state(n, belief)

    def choose_action_device(self) -> torch.Tensor:
        'Action ids (int32, device) of the live agents: argmax_g b.alpha_g -> actions[g] (pbvi_agent.py:741).'
        assert getattr(self, '_bufs', None) is not None, 'Agent was not initialized yet, run the initialize_state function first'
        _, _, act = self.value_function.evaluate_device(self.belief, self.gemm_path)
        self._action_ids = act
        return act

    def choose_action(self) -> np.ndarray:
        ids = self.choose_action_device().cpu().numpy().astype(np.int64)
        self.action_played = ids
        return self.action_set[ids, :]

    # -- training -------------------------------------------------------------------------------------
    def train(self, *args, **kwargs):
        raise NotImplementedError('The train function is not implemented, make a PBVI agent subclass to implement the method')

    def _train_with(self, solver, flavour: dict, expansions, update_passes, max_belief_growth, initial_belief,
                    initial_value_function, prune_level, prune_interval, limit_value_function_size, gamma, eps,
                    convergence_stop, history_tracking_level, overwrite_training, print_progress, print_stats):
        'Shared body of every flavour\'s train() (e.g. reference agents/fsvi_agent.py:191-242).'
        if self.value_function is not None:
            if overwrite_training:
                self.trained_at = None
                self.name = '-'.join(self.name.split('-')[:-1])
                self.value_function = None
            else:
                initial_value_function = self.value_function
        result = solver.solve(model=self.model, expansions=expansions, update_passes=update_passes,
                              max_belief_growth=max_belief_growth, initial_belief=initial_belief,
                              initial_value_function=initial_value_function, prune_level=prune_level,
                              prune_interval=prune_interval, limit_value_function_size=limit_value_function_size,
                              gamma=gamma, eps=eps, convergence_stop=convergence_stop, use_gpu=True,
                              use_reachability=self.use_reachability, rng=self.rng,
                              history_tracking_level=history_tracking_level, print_progress=print_progress,
                              print_stats=print_stats, gemm_path=self.gemm_path, **flavour)
        value_function, hist = result[0], result[1]
        if len(result) > 2:
            self.mdp_policy = result[2]
        self.trained_at = datetime.now().strftime('%Y%m%d_%H%M%S')
        self.name += f'-trained_{self.trained_at}'
        self.value_function = value_function
        if print_stats:
            print(hist.summary)
        self.trained = True
        return hist

    # -- persistence (pbvi_agent.py:450-581; load fixed) --------------------------------
    def save(self, folder: str = None, force: bool = False, save_environment: bool = False) -> None:
        assert self.trained_at is not None, 'The agent is not trained, there is nothing to save.'
        folder = f'./Agent-{self.name}' if folder is None else folder + '/Agent-' + self.name
        if not os.path.exists(folder):
            os.mkdir(folder)
        elif len(os.listdir(folder)):
            if not force:
                raise Exception(f'{folder} is not empty. If you want to overwrite the saved model, enable "force".')
            shutil.rmtree(folder)
            os.mkdir(folder)
        meta = {'name': self.name, 'class': self.class_name, 'thresholds': self.thresholds.tolist(),
                'environment_name': self.environment.name, 'environment_saved_at': self.environment.saved_at,
                'space_aware': self.space_aware, 'spacial_subdivisions': self.spacial_subdivisions.tolist(),
                'action_labels': self.action_labels, 'action_set': np.asarray(self.action_set).tolist(),
                'trained_at': self.trained_at}
        with open(folder + '/METADATA.json', 'w') as f:
            json.dump(meta, f, indent=4)
        self.value_function.save(folder=folder, file_name='Value_Function.npy')
        if save_environment:
            self.environment.save(folder=folder, save_arrays=self.environment.data_file_path is None, force=force)
        self.saved_at = os.path.abspath(folder).replace('\\', '/')
        print(f'Agent saved to: {folder}')

    @classmethod
    def load(cls, folder: str, environment: Environment = None) -> 'PBVI_Agent':
        with open(folder + '/METADATA.json', 'r') as f:
            meta = json.load(f)
        if environment is None:
            saved_env = meta.get('environment_saved_at')
            local = [d for d in os.listdir(folder) if d.startswith('Env-')]
            if local:
                saved_env = folder + '/' + local[0]
            assert saved_env is not None, 'The environment was not saved with the agent: pass environment='
            environment = Environment.load(saved_env)
        if meta['class'] != cls.__name__:
            from .. import agents
            cls = getattr(agents, meta['class'])
        inst = cls(environment=environment, thresholds=meta['thresholds'], space_aware=meta['space_aware'],
                   spacial_subdivisions=np.array(meta['spacial_subdivisions']),
                   actions=dict(zip(meta['action_labels'], meta['action_set'])), name=meta['name'])
        inst.value_function = ValueFunction.load(file=folder + '/Value_Function.npy', model=inst.model)
        inst.trained_at = meta['trained_at']
        inst.trained = True
        inst.saved_at = folder
        return inst

    def generate_beliefs_from_trajectory(self, history, trajectory_i: int = 0, initial_belief: Belief = None) -> BeliefSet:
        '''
        The sequence of beliefs the agent goes through along one trajectory of a SimulationHistory (pbvi_agent.py:812-873):
        replays (action, observation) of every recorded step with Belief.update, never as a goal observation; a failed update is
        reported and skipped like in the reference.  Reads the history's arrays (the reference goes through its pandas frames).
        '''
        belief = Belief(self.model) if initial_belief is None else initial_belief
        sequence = [belief]
        done = int(history.done_at_step[trajectory_i])
        steps = len(history) if done < 0 else done
        action_set = np.asarray(self.action_set)
        for t in range(steps):
            vec = np.asarray(history.actions[t][trajectory_i])
            a = int(np.argwhere(np.all(action_set == vec[None, :], axis=1))[0, 0])
            o = [float(history.observations[t][trajectory_i])]
            if self.space_aware:
                o += [float(v) for v in history.positions[t][trajectory_i]]
            obs = np.array([o]) if self.space_aware else np.array(o)
            discrete_o = int(self.discretize_observations(observation=obs, action=vec[None, :], source_reached=np.array([False]))[0])
            try:
                belief = belief.update(a=a, o=discrete_o, use_reachability=self.use_reachability)
                sequence.append(belief)
            except Exception:
                print(f'[Warning] Update of belief failed at step {t + 1}...')
        return BeliefSet(self.model, sequence)

    def modify_environment(self, new_environment: Environment):
        '''
        A new agent on a modified environment; the alpha vectors are resampled to the new grid like the reference does
        (agents/pbvi_agent.py:661-695: cv2.resize of every alpha vector, bilinear with pixel-centre alignment).
        '''
        modified = self.__class__(environment=new_environment, thresholds=self.thresholds, name=self.name)
        if self.value_function is not None:
            H, W = self.model.grid_shape
            H2, W2 = modified.model.grid_shape
            av = self.value_function.alpha_vector_array.reshape(len(self.value_function), H, W)
            resized = np.array([_resize_bilinear(a, (H2, W2)).ravel() for a in av])
            modified.value_function = ValueFunction(modified.model, alpha_vectors=resized, action_list=self.value_function.actions)
            modified.trained, modified.trained_at = self.trained, self.trained_at
        return modified


def _resize_bilinear(a: np.ndarray, new_shape: tuple) -> np.ndarray:
    'cv2.resize(a, (W2, H2)) (INTER_LINEAR, pixel centres aligned); NumPy fallback with the same convention when cv2 is absent.'
    try:
        import cv2
        return cv2.resize(np.ascontiguousarray(a, dtype=np.float64), (int(new_shape[1]), int(new_shape[0])))
    except ImportError:
        pass
    H, W = a.shape
    H2, W2 = new_shape

    def axis(n, m):
        x = (np.arange(m) + 0.5) * (n / m) - 0.5
        x0 = np.floor(x).astype(int)
        f = x - x0
        lo, hi = np.clip(x0, 0, n - 1), np.clip(x0 + 1, 0, n - 1)
        f = np.where(x0 < 0, 0.0, f)
        return lo, hi, f
    ylo, yhi, fy = axis(H, H2)
    xlo, xhi, fx = axis(W, W2)
    top = a[ylo][:, xlo] * (1 - fx)[None, :] + a[ylo][:, xhi] * fx[None, :]
    bot = a[yhi][:, xlo] * (1 - fx)[None, :] + a[yhi][:, xhi] * fx[None, :]
    return top * (1 - fy)[:, None] + bot * fy[:, None]
