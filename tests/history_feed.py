'''
The fixed add_step sequence behind tests/golden/history_T0.npz: fed to the reference's SimulationHistory by
oracle/make_golden_history.py and to the package's by tests/test_history.py.
'''
from datetime import datetime, timedelta

import numpy as np


def feed(make_history, starts, shifts, seed, t0):
    '''The fixed add_step sequence: n runs, horizon 6; every step some runs reach the source, one is interrupted.'''
    rng = np.random.default_rng(seed)
    n = len(starts)
    h = make_history(starts, shifts, 6)
    pos = starts.copy()
    live = np.arange(n)
    moves = np.array([[-1, 0], [0, 1], [1, 0], [0, -1]])
    for step in range(6):
        act = rng.integers(0, 4, len(live))
        pos = pos + moves[act]
        odor = np.round(rng.random(len(live)), 6)
        reached = rng.random(len(live)) < 0.25
        interrupt = (~reached) & (rng.random(len(live)) < 0.1) if step < 5 else np.zeros(len(live), bool)
        h.add_step(actions=moves[act], next_positions=pos, observations=odor, reached_source=reached, interupt=interrupt)
        keep = ~(reached | interrupt)
        live, pos = live[keep], pos[keep]
        if len(live) == 0:
            break
    k = len(h.timestamps[0])
    h.timestamps = {0: [t0 + timedelta(seconds=0.37 * i * (1 + 0.1 * i)) for i in range(k)]}
    return h


STARTS_A = np.array([[3, 5], [4, 6], [5, 7], [8, 2], [6, 6], [2, 9], [7, 7]])
STARTS_B = np.array([[1, 1], [9, 4], [5, 5], [4, 8], [3, 3]])
T0_A = datetime(2024, 1, 2, 23, 59, 58, 500000)            # the second batch of timestamps crosses midnight
T0_B = datetime(2024, 1, 2, 23, 59, 59, 250000)
