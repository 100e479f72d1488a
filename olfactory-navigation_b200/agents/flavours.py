'''
Solver-flavour agents: each only adds `train()` with the reference's signature and forwards to the
device solver (reference agents/{fsvi,pbvi_ra,pbvi_ssra,pbvi_ssga,pbvi_ssea,pbvi_ger,perseus,hsvi}_agent.py).
'''
from __future__ import annotations

from .. import solvers
from ..pomdp import ValueFunction
from .pbvi_agent import PBVI_Agent

_COMMON = dict(expansions=10, update_passes=1, max_belief_growth=10, initial_belief=None, initial_value_function=None,
               prune_level=1, prune_interval=10, limit_value_function_size=-1, gamma=0.99, eps=1e-6,
               convergence_stop=False, history_tracking_level=1, overwrite_training=False, print_progress=True,
               print_stats=True)


def _split(kwargs: dict, flavour_keys: tuple) -> tuple[dict, dict]:
    kwargs.pop('use_gpu', None)          # accepted for signature compatibility; this package only runs on the device
    common = dict(_COMMON)
    flavour = {}
    for k, v in kwargs.items():
        if k in common:
            common[k] = v
        elif k in flavour_keys:
            flavour[k] = v
        else:
            raise TypeError(f'train() got an unexpected keyword argument {k!r}')
    return common, flavour


class FSVI_Agent(PBVI_Agent):
    'reference agents/fsvi_agent.py:113-243: train(..., mdp_policy=None, vi_horizon=1000, ...)'
    mdp_policy: ValueFunction = None

    def train(self, expansions: int = 10, **kwargs):
        common, flavour = _split(dict(kwargs, expansions=expansions), ('mdp_policy', 'vi_horizon'))
        flavour.setdefault('mdp_policy', None)
        flavour.setdefault('vi_horizon', 1000)
        flavour['return_mdp_policy'] = True
        return self._train_with(solvers.FSVI, flavour, **common)


class PBVI_RA_Agent(PBVI_Agent):
    def train(self, expansions: int = 10, **kwargs):
        common, flavour = _split(dict(kwargs, expansions=expansions), ())
        return self._train_with(solvers.PBVI_RA, flavour, **common)


class PBVI_SSRA_Agent(PBVI_Agent):
    def train(self, expansions: int = 10, **kwargs):
        common, flavour = _split(dict(kwargs, expansions=expansions), ())
        return self._train_with(solvers.PBVI_SSRA, flavour, **common)


class PBVI_SSGA_Agent(PBVI_Agent):
    'reference agents/pbvi_ssga_agent.py:176-198: extra kwargs epsilon, epsilon_decay'
    def train(self, expansions: int = 10, **kwargs):
        common, flavour = _split(dict(kwargs, expansions=expansions), ('epsilon', 'epsilon_decay'))
        return self._train_with(solvers.PBVI_SSGA, flavour, **common)


class PBVI_SSEA_Agent(PBVI_Agent):
    def train(self, expansions: int = 10, **kwargs):
        common, flavour = _split(dict(kwargs, expansions=expansions), ())
        return self._train_with(solvers.PBVI_SSEA, flavour, **common)


class PBVI_GER_Agent(PBVI_Agent):
    def train(self, expansions: int = 10, **kwargs):
        common, flavour = _split(dict(kwargs, expansions=expansions), ())
        return self._train_with(solvers.PBVI_GER, flavour, **common)


class Perseus_Agent(PBVI_Agent):
    def train(self, expansions: int = 10, **kwargs):
        common, flavour = _split(dict(kwargs, expansions=expansions), ('use_policy_to_choose_actions',))
        return self._train_with(solvers.Perseus, flavour, **common)


class HSVI_Agent(PBVI_Agent):
    mdp_policy: ValueFunction = None

    def train(self, expansions: int = 10, **kwargs):
        common, flavour = _split(dict(kwargs, expansions=expansions), ('mdp_policy', 'vi_horizon', 'epsilon'))
        flavour['return_mdp_policy'] = True           # reference agents/hsvi_agent.py:184: value_function, hist, mdp_policy = HSVI.solve(...)
        return self._train_with(solvers.HSVI, flavour, **common)
