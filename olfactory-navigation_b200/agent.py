'''
Agent base class — same constructor, attributes and simulation protocol as the reference's
`olfactory_navigation.Agent` (reference olfactory_navigation/agent.py:144-318, :343-482).
Written from scratch.  Layered environments (actions carry a layer component, thresholds may differ per layer:
agent.py:171-201,281-283,413-417) and space-aware observations (agent.py:206-234,418-434) are supported on the host side;
the model converters reject space-aware agents exactly like the reference's (environment_converter.py:37).
'''
from __future__ import annotations

import inspect
import os
import pickle
import shutil
import sys

import numpy as np

from .environment import Environment
from .util import get_rng


class Agent:
    def __init__(self,
                 environment: Environment,
                 thresholds: float | list[float] = 3e-6,
                 space_aware: bool = False,
                 spacial_subdivisions: np.ndarray = None,
                 actions: dict | np.ndarray = None,
                 name: str = None,
                 rng: int | np.random.Generator = None) -> None:
        self.environment = environment
        self.space_aware = space_aware
        self.trained = False

        # thresholds -> ascending array that starts with -inf and ends with +inf (agent.py:157-204); a dict {layer label: value(s)}
        # or a 2-D array gives one row per layer of a layered environment (the reference's dict branch is broken, SURVEY App. C;
        # this one implements what it documents)
        def _one(t) -> list:
            t = t.tolist() if isinstance(t, np.ndarray) else t
            t = [float(t)] if isinstance(t, (float, int, np.floating, np.integer)) else [float(v) for v in t]
            if -np.inf not in t:
                t = [-np.inf] + t
            if np.inf not in t:
                t = t + [np.inf]
            return sorted(t)
        if isinstance(thresholds, dict):
            assert environment.has_layers, 'Thresholds can only be provided as a dict if the environment is layered.'
            assert len(thresholds) == len(environment.layer_labels), 'Thresholds must be given for each layers.'
            for label in environment.layer_labels:
                assert label in thresholds, f"Environment's layer [{label}] was not matched to any layers of the environment."
            rows = [_one(thresholds[label]) for label in environment.layer_labels]
            assert all(len(r) == len(rows[0]) for r in rows), 'If provided as lists, the threshold lists must all be of the same length.'
            self.thresholds = np.array(rows, dtype=float)
        elif isinstance(thresholds, np.ndarray) and thresholds.ndim == 2:
            assert environment.has_layers and len(thresholds) == len(environment.layers), 'A 2-D thresholds array needs one row per layer.'
            self.thresholds = np.array([_one(r) for r in thresholds], dtype=float)
        else:
            self.thresholds = np.array(_one(thresholds), dtype=float)

        self.spacial_subdivisions = np.array(environment.shape if spacial_subdivisions is None else spacial_subdivisions)
        assert len(self.spacial_subdivisions) == environment.dimensions, \
            'The amount of spacial divisions must match the amount of dimensions of the environment.'
        # grid cell -> subdivision id (agent.py:213-234): equal cells, the remainder split between the first and the last one
        env_shape = np.array(environment.shape)
        std_size = (env_shape / self.spacial_subdivisions).astype(int)
        overflows = env_shape % self.spacial_subdivisions
        edges = []
        for parts, size, over in zip(self.spacial_subdivisions, std_size, overflows):
            sizes = np.repeat(size, parts)
            if parts >= 2:
                sizes = sizes + np.array([np.floor(over / 2), *np.zeros(parts - 2), np.ceil(over / 2)]).astype(int)
            else:
                sizes = sizes + int(over)
            edges.append(np.concatenate(([0], np.cumsum(sizes))).astype(int))
        mapping = np.full(environment.shape, -1)
        cell = 0
        for y0, y1 in zip(edges[0][:-1], edges[0][1:]):
            for x0, x1 in zip(edges[1][:-1], edges[1][1:]):
                mapping[y0:y1, x0:x1] = cell
                cell += 1
        self._environment_to_subdivision_mapping = mapping

        # action set (agent.py:237-297): default N, E, S, W as (dy, dx); a layered environment multiplies it by the layers
        # ([layer, dy, dx], labels l_<layer>_<action>, agent.py:281-283)
        layered = 1 if environment.has_layers else 0
        if actions is None:
            assert environment.dimensions == 2, 'only 2-D environments are supported'
            self.action_set = np.array([[-1, 0], [0, 1], [1, 0], [0, -1]])
            self.action_labels = ['North', 'East', 'South', 'West']
            if layered:
                self.action_set = np.array([[layer, *vec] for layer in environment.layers for vec in self.action_set])
                self.action_labels = [f'l_{layer}_{action}' for layer in environment.layer_labels for action in self.action_labels]
        elif isinstance(actions, np.ndarray):
            self.action_set = actions
            self.action_labels = ['a_' + '_'.join(str(v) for v in vec) for vec in actions]
        else:   # dict label -> vector (the reference's dict branch is broken; this one works)
            self.action_set = np.array(list(actions.values()))
            self.action_labels = list(actions.keys())
        assert self.action_set.ndim == 2 and self.action_set.shape[1] == layered + environment.dimensions, \
            f'The shape of the action_set provided is not right. (Found {self.action_set.shape}; expected (., {layered + environment.dimensions}))'

        if name is None:
            name = self.class_name + '-tresholds-' + '_'.join(str(v) for v in self.thresholds.ravel())
        self.name = name
        self.saved_at = None
        self.is_on_gpu = True          # the arithmetic of this package only exists on the device
        self._alternate_version = None
        self.rng = get_rng(rng)

    @property
    def class_name(self) -> str:
        return self.__class__.__name__

    @property
    def on_gpu(self) -> 'Agent':
        return self

    @property
    def on_cpu(self) -> 'Agent':
        return self

    # -- training / simulation protocol (agent.py:324-482) ---------------------------------------
    def train(self) -> None:
        raise NotImplementedError('The train function is not implemented, make an agent subclass to implement the method')

    def initialize_state(self, n: int = 1) -> None:
        raise NotImplementedError('The initialize_state function is not implemented, make an agent subclass to implement the method')

    def choose_action(self) -> np.ndarray:
        raise NotImplementedError('The choose_action function is not implemented, make an agent subclass to implement the method')

    def update_state(self, action: np.ndarray, observation: np.ndarray, source_reached: np.ndarray):
        raise NotImplementedError('The update_state function is not implemented, make an agent subclass to implement the method')

    def kill(self, simulations_to_kill: np.ndarray) -> None:
        raise NotImplementedError('The kill function is not implemented, make an agent subclass to implement the method')

    def discretize_observations(self, observation: np.ndarray, action: np.ndarray, source_reached: np.ndarray) -> np.ndarray:
        '''
        Odor value -> observation id: bin k with t[k] <= x < t[k+1]; the goal id (= number of bins)
        where the source was reached (agent.py:401-439).  Host/NumPy version of the API; the
        device version is part of olfnav_env_step.
        '''
        observation = np.asarray(observation, dtype=float)
        odor = observation[:, 0] if self.space_aware else observation
        bins = self.thresholds.shape[-1] - 1
        if self.environment.has_layers and self.thresholds.ndim == 2:
            thr = self.thresholds[np.asarray(action)[:, 0].astype(int)]            # the layer is the first action component (agent.py:414-417)
            ids = np.sum(odor[:, None] >= thr[:, 1:], axis=1).astype(int)
        else:
            thr = self.thresholds if self.thresholds.ndim == 1 else self.thresholds[0]
            ids = (np.searchsorted(thr, odor, side='right') - 1).astype(int)
        count = bins
        if self.space_aware:                                                        # agent.py:422-434
            position = np.clip(observation[:, 1:], 0, np.array(self.environment.shape) - 1).astype(int)
            cells = int(np.prod(self.spacial_subdivisions))
            ids = ids * cells + self._environment_to_subdivision_mapping[tuple(position.T)]
            count = bins * cells
        ids[np.asarray(source_reached, dtype=bool)] = count
        return ids

    # -- persistence (agent.py:488-567: pickle by default, subclasses store only what is needed) --
    _transient = ('belief', '_bufs', '_alternate_version', 'action_played', 'succeeded_update', 'last_delta_H', '_action_ids')

    def __getstate__(self):
        'Simulation state (device buffers of the running agents) is never saved with the agent.'
        state = self.__dict__.copy()
        for key in self._transient:
            if key in state:
                state[key] = None
        return state

    def save(self, folder: str = None, force: bool = False, save_environment: bool = False) -> None:
        folder = f'./Agent-{self.name}' if folder is None else folder + f'/Agent-{self.name}'
        if not os.path.exists(folder):
            os.mkdir(folder)
        elif len(os.listdir(folder)) > 0:
            if not force:
                raise Exception(f'{folder} is not empty. If you want to overwrite the saved agent, enable "force".')
            shutil.rmtree(folder)
            os.mkdir(folder)
        with open(folder + '/binary.pkl', 'wb') as f:
            pickle.dump(self, f)
        if save_environment:
            self.environment.save(folder=(folder + f'/Env-{self.environment.name}'))
        self.saved_at = os.path.abspath(folder)

    @classmethod
    def load(cls, folder: str) -> 'Agent':
        'Dispatches to the subclass whose name appears in the folder name (agent.py:557-561), else unpickles.'
        from . import agents
        for name, obj in inspect.getmembers(agents):
            if inspect.isclass(obj) and (name in folder) and issubclass(obj, cls) and (obj is not cls) and ('load' in obj.__dict__):
                return obj.load(folder)
        with open(folder + '/binary.pkl', 'rb') as f:
            return pickle.load(f)

    def modify_environment(self, new_environment: Environment) -> 'Agent':
        'A new agent of the same class, thresholds and name on another environment (agent.py:569-593).'
        return self.__class__(environment=new_environment, thresholds=self.thresholds, name=self.name)

    def size_on_memory(self) -> int:
        'Bytes held by the agent and its components, host arrays and device tensors alike (agent.py:656-667).'
        seen, total, stack = set(), 0, [self]
        while stack:
            obj = stack.pop()
            if id(obj) in seen:
                continue
            seen.add(id(obj))
            if isinstance(obj, np.ndarray):
                total += obj.nbytes
            elif hasattr(obj, 'element_size') and hasattr(obj, 'nelement'):
                total += obj.element_size() * obj.nelement()
            else:
                total += sys.getsizeof(obj)
                if isinstance(obj, dict):
                    stack.extend(obj.values())
                elif isinstance(obj, (list, tuple, set)):
                    stack.extend(obj)
                elif hasattr(obj, '__dict__') and not inspect.ismodule(obj) and not inspect.isclass(obj):
                    stack.extend(v for k, v in vars(obj).items() if k != '_alternate_version')
        return total

    @property
    def movement_set(self) -> np.ndarray:
        'The physical (dy, dx) part of the action vectors (a layered action also carries its layer first).'
        return self.action_set[:, -self.environment.dimensions:]

    @property
    def action_layers(self) -> np.ndarray:
        'Layer of every action id (zeros without layers).'
        return self.action_set[:, 0].astype(int) if self.environment.has_layers else np.zeros(len(self.action_set), dtype=int)
