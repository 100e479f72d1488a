'''
ORACLE — TEST INFRASTRUCTURE ONLY.  Generates tests/golden/history_T0.npz.

Runs in THIS container only (imports the UNMODIFIED reference from /root/reference through oracle/refloader.py): feeds a fixed
sequence of add_step calls and fixed timestamps (two batches joined with +) to the reference's SimulationHistory
(simulation.py:121-227,550-618) and freezes the three files its save() writes (simulation.py:621-740: the simulations CSV, the
runs analysis and the general analysis).  tests/test_history.py feeds the same sequence to the package's SimulationHistory.

    python oracle/make_golden_history.py
'''
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import refloader  # noqa: E402

on = refloader.load()
from olfactory_navigation import Environment, Agent                 # noqa: E402
from olfactory_navigation.simulation import SimulationHistory       # noqa: E402
import olfactory_navigation_b200.synthetic as syn                   # noqa: E402


sys.path.insert(0, os.path.join(ROOT, 'tests'))
from history_feed import feed, STARTS_A, STARTS_B, T0_A, T0_B      # noqa: E402  (the fixed add_step sequence, shared with tests/test_history.py)

if __name__ == '__main__':
    c = syn.config('T0')
    env = Environment(**c['env_kwargs'])
    agent = Agent(env, thresholds=c['thresholds'])
    mk = lambda s, t, hz: SimulationHistory(start_points=s, environment=env, agent=agent, time_shift=t, horizon=hz, reward_discount=0.99)   # noqa: E731
    ha = feed(mk, STARTS_A, np.arange(len(STARTS_A)) * 3, 1, T0_A)
    hb = feed(mk, STARTS_B, np.arange(len(STARTS_B)) + 40, 2, T0_B)
    both = ha + hb
    both.used_gpu = False          # the reference's __add__ does not carry used_gpu over (simulation.py:550-618) and its save() then raises
    d = tempfile.mkdtemp()
    both.save(file='hist', folder=d, save_analysis=True)
    out = {name: np.array(open(os.path.join(d, f)).read()) for name, f in
           (('csv', 'hist.csv'), ('runs_analysis', 'hist-runs_analysis.csv'), ('general_analysis', 'hist-general_analysis.csv'))}
    # (the reference's table cells are one-element Series — its analysis table has a one-level MultiIndex; the numbers are frozen)
    out['compare'] = np.array(both.compare(ha, hb).map(lambda v: float(np.asarray(v).ravel()[0])).to_csv(index=False))
    out.update(summary=np.array(both.summary), done_at_step=both.done_at_step, reached_source=both.reached_source, n_steps=np.array(len(both.positions)))
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'history_T0.npz'), **out)
    print(out['csv'][()][:1500])
    print(out['general_analysis'][()])
