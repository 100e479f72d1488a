'''
QMDP_Agent — value function = one alpha vector per action, Q(., a) of the fully observable MDP
(reference agents/qmdp_agent.py:100-167); simulation methods inherited from PBVI_Agent.
'''
from __future__ import annotations

from datetime import datetime

from .. import solvers
from .pbvi_agent import PBVI_Agent


class QMDP_Agent(PBVI_Agent):
    def train(self, expansions: int = 1000, initial_value_function=None, gamma: float = 0.99, eps: float = 1e-6,
              use_gpu: bool = True, history_tracking_level: int = 1, overwrite_training: bool = False,
              print_progress: bool = True, print_stats: bool = True):
        if self.value_function is not None:
            if overwrite_training:
                self.trained_at = None
                self.name = '-'.join(self.name.split('-')[:-1])
                self.value_function = None
            else:
                initial_value_function = self.value_function
        value_function, hist = solvers.VI.solve(model=self.model, horizon=expansions,
                                                initial_value_function=initial_value_function, gamma=gamma, eps=eps,
                                                history_tracking_level=history_tracking_level, print_progress=print_progress)
        self.trained_at = datetime.now().strftime('%Y%m%d_%H%M%S')
        self.name += f'-trained_{self.trained_at}'
        self.value_function = value_function
        if print_stats:
            print(hist.summary)
        self.trained = True
        return hist
