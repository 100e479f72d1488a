'''
ORACLE — TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*_M1.npz: DENSE stochastic-transition model.

Runs in THIS container only: the UNMODIFIED reference's minimal_converter (agents/model_based_util/environment_converter.py:137-292)
on the C1 synthetic environment, the reference QMDP / FSVI agents on that model (restated NumPy `pomdp_toolkit` underneath) and the
reference run_test.  Frozen: the model tables, the VI solution, a batch of belief updates / evaluations / one backup / infotaxis
deltas from the oracle toolkit, and run_test trajectories.

    python oracle/make_golden_minimal.py
'''
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import refloader  # noqa: E402

on = refloader.load()
from olfactory_navigation import Environment, run_test            # noqa: E402
from olfactory_navigation.agents import FSVI_Agent, QMDP_Agent     # noqa: E402
from olfactory_navigation.agents.model_based_util.environment_converter import minimal_converter   # noqa: E402
from pomdp_toolkit import BeliefSet, ValueFunction                 # noqa: E402
from pomdp_toolkit import solvers as osolvers                      # noqa: E402

import olfactory_navigation_b200.synthetic as syn                  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
PARTS = [3, 6]


def save(name, **arrays):
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **arrays)
    print(f'{name:28s} {os.path.getsize(path) / 1024:8.1f} KiB')


c = syn.config('C1')
env = Environment(**c['env_kwargs'])
kw = dict(environment_converter=minimal_converter, partitions=PARTS, margin_partitions=True)
ag = QMDP_Agent(env, thresholds=c['thresholds'], **kw)
model = ag.model
S, A, O = model.state_count, model.action_count, model.observation_count
save('converter_M1', transition_table=model.transition_table, observation_table=model.observation_table,
     end_states=model.end_states.astype(np.int32), start_probabilities=model.start_probabilities, partitions=np.array(PARTS),
     expected_rewards=model.expected_rewards_table)

rng = np.random.default_rng(syn.SEED + 99)
n = 120
d = -np.log(1.0 - rng.random((n, S)))
d[rng.random((n, S)) < 0.4] = 0.0
d[:, 0] += 1e-3
B = (d / d.sum(1, keepdims=True)).astype(np.float32).astype(np.float64)
B[:6] = 0
B[np.arange(6), rng.integers(0, S, 6)] = 1.0
act = rng.integers(0, A, n)
obs = rng.integers(0, O, n)
u, p = BeliefSet(model, B).unnormalised_successors()
idx = np.arange(n)
z = p[idx, act, obs]
ok = z > 0
out = np.zeros_like(B)
out[ok] = u[idx, act, obs][ok] / z[ok, None]
save('belief_M1', b_in=B.astype(np.float32), act=act.astype(np.int32), obs=obs.astype(np.int32), b_out=out, norm=z, ok=ok)

vf_q, hist = osolvers.VI.solve(model, horizon=1000, gamma=0.99, eps=1e-6)
save('vi_M1', Q=vf_q.alpha_vector_array, sweeps=np.array(len(hist.value_function_changes)))

G = 11
alphas = np.vstack([vf_q.alpha_vector_array] + [vf_q.alpha_vector_array[rng.integers(0, A)] * (1 + 0.3 * rng.standard_normal(S)) for _ in range(G - A)])
alphas = alphas.astype(np.float32).astype(np.float64)
acts = np.concatenate([np.arange(A), rng.integers(0, A, G - A)])
_, det = osolvers.backup(model, BeliefSet(model, B), ValueFunction(model, alphas, acts), gamma=0.99, append=False,
                         belief_dominance_prune=False, return_details=True)
q_sorted = np.sort(det['q'], axis=1)
save('backup_M1', beliefs=B.astype(np.float32), alphas=alphas.astype(np.float32), alpha_actions=acts.astype(np.int32), discount=np.array(0.99),
     best_a=det['best_a'].astype(np.int32), new_alpha=det['new_alpha'], new_value=det['new_value'], a_margin=(q_sorted[:, -1] - q_sorted[:, -2]))

# infotaxis deltas (formula of agents/infotaxis_agent.py:257-293 on the oracle successors)
bs = BeliefSet(model, B)
Hb = bs.entropies
dH = np.zeros((n, A))
for b in range(n):
    for a in range(A):
        acc = 0.0
        for o in range(O):
            if p[b, a, o] > 0:
                w = u[b, a, o] / p[b, a, o]
                nz = w > 0
                acc += p[b, a, o] * (-np.sum(w[nz] * np.log(w[nz])))
        dH[b, a] = Hb[b] - acc
save('infotaxis_M1', delta_H=dH)

n_sim, horizon = 32, 80
for label, cls, train_kwargs in (('qmdp', QMDP_Agent, dict(expansions=1000)), ('fsvi', FSVI_Agent, dict(expansions=6, max_belief_growth=10))):
    agent = cls(env, thresholds=c['thresholds'], **kw)
    agent.rng = np.random.default_rng(syn.SEED + 1)
    agent.train(print_progress=False, print_stats=False, **train_kwargs)
    sim_rng = np.random.default_rng(syn.SEED + 2)
    starts = env.random_start_points(n_sim, sim_rng)
    tshift = sim_rng.integers(0, env.timesteps, n_sim)
    vfa = agent.value_function
    try:
        hist = run_test(agent, n=n_sim, start_points=starts, time_shift=tshift, horizon=horizon, print_progress=False,
                        print_stats=False, use_gpu=False, batches=1)
    except ValueError as exc:
        # Impossible belief updates DO occur on this coarse model, and the reference's run_test cannot survive one: update_state
        # returns a mask of length N_ok instead of N (agents/pbvi_agent.py:783-790).  Freeze the value function only;
        # tests/test_dense_model.py checks the device run against the intended-semantics protocol loop.
        print(f'  reference run_test failed on {label} (known latent bug): {exc}')
        save(f'runtest_{label}_M1', alphas=vfa.alpha_vector_array.astype(np.float32), alpha_actions=vfa.actions.astype(np.int32),
             start_points=starts.astype(np.int32), time_shift=tshift.astype(np.int64), horizon=np.array(horizon), reference_failed=np.array(True))
        continue
    save(f'runtest_{label}_M1', alphas=vfa.alpha_vector_array.astype(np.float32), alpha_actions=vfa.actions.astype(np.int32),
         start_points=starts.astype(np.int32), time_shift=tshift.astype(np.int64), horizon=np.array(horizon),
         actions=np.array(hist.actions).astype(np.int8), positions=np.array(hist.positions).astype(np.int16),
         observations=np.array(hist.observations).astype(np.float32), reached_source=hist.reached_source,
         done_at_step=hist.done_at_step.astype(np.int32))
print('done')
