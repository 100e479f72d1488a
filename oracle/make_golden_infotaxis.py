'''
ORACLE — TEST INFRASTRUCTURE ONLY.  Generates tests/golden/infotaxis_C1.npz (BASELINE configs[0] shape, 50 x 50 states, 'stop'
boundary: every clipped-edge merge rule is exercised) by running the UNMODIFIED reference Infotaxis_Agent.choose_action
(/root/reference/olfactory_navigation/agents/infotaxis_agent.py:244-301) on top of the restated toolkit, like oracle/make_golden.py
does for the tiny configurations.  Build container only (/root/reference cannot travel).

    python oracle/make_golden_infotaxis.py
'''
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import refloader  # noqa: E402

on = refloader.load()
from olfactory_navigation import Environment                      # noqa: E402
from olfactory_navigation.agents import Infotaxis_Agent            # noqa: E402
from pomdp_toolkit import BeliefSet                                # noqa: E402

import olfactory_navigation_b200.synthetic as syn                  # noqa: E402

cfg = 'C1'
c = syn.config(cfg)
env = Environment(**c['env_kwargs'])
it = Infotaxis_Agent(env, thresholds=c['thresholds'], use_reachability=True)
model = it.model
S, A, O = model.state_count, model.action_count, model.observation_count
H, W = env.shape
rng = np.random.default_rng(syn.SEED + 77)

# beliefs: the start belief, beliefs walked from it (some after smelling something), dense random rows, border-heavy rows, point masses
rows = [model.start_probabilities.copy()]
bs = BeliefSet(model, np.tile(model.start_probabilities, (10, 1)))
for step in range(8):
    a = rng.integers(0, A, len(bs))
    o = (rng.random(len(bs)) < 0.3).astype(int) * rng.integers(1, O - 1, len(bs))
    nxt, prov = bs.update(a, o, raise_on_impossible_belief=False, return_provenance=True)
    if len(nxt) == 0:
        break
    bs = nxt
    if step in (1, 4, 7):
        rows += list(bs.belief_array[:4])
d = -np.log(1.0 - rng.random((4, S)))
rows += list(d / d.sum(1, keepdims=True))
border = np.zeros((H, W))
border[0, :] = border[-1, :] = border[:, 0] = border[:, -1] = 1.0
border += 0.01 * rng.random((H, W))
rows.append((border / border.sum()).ravel())
for (y, x) in [(0, 0), (0, W - 1), (H - 1, 0), (H // 2, 0), (0, W // 2), (25, 12)]:
    e = np.zeros((H, W))
    e[y, x] = 1.0
    rows.append(e.ravel())
B = np.asarray(rows, dtype=np.float32).astype(np.float64)
B /= B.sum(1, keepdims=True)
B = B.astype(np.float32).astype(np.float64)

it.belief = BeliefSet(it.model, B)
moves = it.choose_action()
chosen = it.action_played.copy()

# the per-action deltas with the reference's formula (infotaxis_agent.py:257-293), float64
bsB = BeliefSet(it.model, B)
Hb = bsB.entropies
u, p = bsB.unnormalised_successors()
dH = np.zeros((len(B), A))
for b in range(len(B)):
    for a in range(A):
        acc = 0.0
        for o in range(O):
            if p[b, a, o] > 0:
                w = u[b, a, o] / p[b, a, o]
                nz = w > 0
                acc += p[b, a, o] * (-np.sum(w[nz] * np.log(w[nz])))
        dH[b, a] = Hb[b] - acc
assert np.array_equal(np.argmax(dH, axis=1), chosen), 'restated delta-H disagrees with the reference loop'
d_sorted = np.sort(dH, axis=1)
path = os.path.join(ROOT, 'tests', 'golden', f'infotaxis_{cfg}.npz')
np.savez_compressed(path, beliefs=B.astype(np.float32), chosen=chosen.astype(np.int32), moves=moves.astype(np.int32), delta_H=dH, entropies=Hb,
                    margin=(d_sorted[:, -1] - d_sorted[:, -2]))
print(path, os.path.getsize(path) // 1024, 'KiB', len(B), 'beliefs; margins', np.sort(d_sorted[:, -1] - d_sorted[:, -2])[:6])
