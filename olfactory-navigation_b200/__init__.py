'''
olfactory-navigation_b200 — B200-native POMDP core behind the Environment / Agent / run_test API of
PimLb/olfactory-navigation.  Import as `olfactory_navigation_b200` (shim module at the repo root).

Public surface (same names as the reference package, reference olfactory_navigation/__init__.py:2-12):
    Environment, Agent, run_test, SimulationHistory, agents.{PBVI,FSVI,QMDP,Infotaxis,...}_Agent
plus the toolkit-surface mirror the agents use: POMDP, Belief, BeliefSet, ValueFunction, solvers.

All arithmetic runs in libolfnav_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/olfnav_b200.h).  There is no CPU fallback: importing without the built library raises.
'''
__version__ = '0.1.0'

from . import _lib                                    # noqa: F401  (raises ImportError when the .so is missing)
from .environment import Environment                  # noqa: F401
from .agent import Agent                              # noqa: F401
from .pomdp import POMDP, Belief, BeliefSet, ValueFunction   # noqa: F401
from . import solvers                                 # noqa: F401
from . import agents                                  # noqa: F401
from .converter import exact_converter, minimal_converter   # noqa: F401
from .simulation import run_test, SimulationHistory   # noqa: F401
from . import synthetic                               # noqa: F401
from . import test_setups                             # noqa: F401
