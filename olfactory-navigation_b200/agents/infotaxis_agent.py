'''
Infotaxis_Agent — no training; picks the action with the largest expected entropy reduction
(reference agents/infotaxis_agent.py:25-363, Vergassola et al.).  choose_action is ONE fused
reduction kernel per step (olfnav_infotaxis_choose): the successor beliefs the reference
materialises (:260-262, up to N*A*O rows of S) never exist.
'''
from __future__ import annotations

import numpy as np
import torch

from .._lib import lib, ptr, stream, check
from ..agent import Agent
from ..converter import exact_converter
from ..environment import Environment
from ..pomdp import POMDP, BeliefSet
from .pbvi_agent import BeliefSimulationMixin


class Infotaxis_Agent(BeliefSimulationMixin, Agent):
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
        super().__init__(environment=environment, thresholds=thresholds, space_aware=space_aware,
                         spacial_subdivisions=spacial_subdivisions, actions=actions, name=name, rng=rng)
        if model is not None:
            self.model = model
        elif callable(environment_converter):
            self.model = environment_converter(agent=self, **converter_parameters)
        else:
            self.model = exact_converter(agent=self)
        self.use_reachability = use_reachability
        self.trained = True            # nothing to train
        self._bufs = None
        self._n_live = 0
        self.action_played = None
        self.succeeded_update = None
        self.last_delta_H = None

    def initialize_state(self, n: int = 1, belief: BeliefSet = None) -> None:
        self._init_belief_state(n, belief)

    def choose_action_device(self, return_delta_H: bool = False) -> torch.Tensor:
        assert getattr(self, '_bufs', None) is not None, 'Agent was not initialized yet, run the initialize_state function first'
        n, dev = self._n_live, self._bufs[0].device
        act = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        dH = torch.empty((max(n, 1), self.model.action_count), dtype=torch.float32, device=dev) if return_delta_H else None
        b = self._bufs[self._cur]
        with torch.cuda.device(self._device):
            check(lib.olfnav_infotaxis_choose(self.model.handle(self._device), ptr(b), b.stride(0), n,
                                              ptr(self._scales[self._cur]), ptr(act), ptr(dH), stream()))
        self._action_ids = act[:n]
        self.last_delta_H = dH[:n] if dH is not None else None
        return self._action_ids

    def choose_action(self) -> np.ndarray:
        ids = self.choose_action_device().cpu().numpy().astype(np.int64)
        self.action_played = ids
        return self.action_set[ids, :]
