'''
Agent base class — same constructor, attributes and simulation protocol as the reference's
`olfactory_navigation.Agent` (reference olfactory_navigation/agent.py:144-318, :343-482).
Written from scratch; layered thresholds (dict) and space-aware observations are outside the
hot path and raise NotImplementedError.
'''
from __future__ import annotations

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
        if space_aware:
            raise NotImplementedError('space_aware agents are out of scope of the B200 hot path (SURVEY §8 f4)')

        # thresholds -> ascending array that starts with -inf and ends with +inf (agent.py:157-204)
        if isinstance(thresholds, dict):
            raise NotImplementedError('per-layer thresholds are out of scope of the B200 hot path (SURVEY §8 f4)')
        if isinstance(thresholds, np.ndarray):
            thresholds = thresholds.tolist()
        if isinstance(thresholds, (float, int, np.floating)):
            thresholds = [float(thresholds)]
        thresholds = list(thresholds)
        if -np.inf not in thresholds:
            thresholds = [-np.inf] + thresholds
        if np.inf not in thresholds:
            thresholds = thresholds + [np.inf]
        self.thresholds = np.sort(np.array(thresholds, dtype=float))

        self.spacial_subdivisions = np.array(environment.shape if spacial_subdivisions is None else spacial_subdivisions)
        assert len(self.spacial_subdivisions) == environment.dimensions, \
            'The amount of spacial divisions must match the amount of dimensions of the environment.'

        # action set (agent.py:237-297): default N, E, S, W as (dy, dx)
        if actions is None:
            assert environment.dimensions == 2, 'only 2-D environments are supported'
            self.action_set = np.array([[-1, 0], [0, 1], [1, 0], [0, -1]])
            self.action_labels = ['North', 'East', 'South', 'West']
        elif isinstance(actions, np.ndarray):
            self.action_set = actions
            self.action_labels = ['a_' + '_'.join(str(v) for v in vec) for vec in actions]
        else:   # dict label -> vector (the reference's dict branch is broken; this one works)
            self.action_set = np.array(list(actions.values()))
            self.action_labels = list(actions.keys())
        assert self.action_set.ndim == 2 and self.action_set.shape[1] == environment.dimensions, \
            f'The shape of the action_set provided is not right. (Found {self.action_set.shape}; expected (., {environment.dimensions}))'

        if name is None:
            name = self.class_name + '-tresholds-' + '_'.join(str(v) for v in self.thresholds)
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
        ids = np.searchsorted(self.thresholds, observation, side='right') - 1
        ids = ids.astype(int)
        ids[np.asarray(source_reached, dtype=bool)] = len(self.thresholds) - 1
        return ids
