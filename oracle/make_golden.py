'''
ORACLE — TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz.

Runs in THIS container only: it imports the UNMODIFIED reference Python from /root/reference
(through oracle/refloader.py: cupy pass-through + h5py/matplotlib stubs) on top of the restated
NumPy `pomdp_toolkit` (oracle/pomdp_toolkit) and freezes inputs/outputs of every hot-path
operation as small fixtures.  /root/reference cannot travel to the GPU box; these files can.

    python oracle/make_golden.py            # rewrites tests/golden/

What is pinned by what:
  * converter_*   : reference Environment + Agent + exact_converter (all in-repo reference code)
  * envstep_*     : reference Environment.move/get_observation/source_reached + Agent.discretize_observations
  * runtest_*     : reference run_test driving reference QMDP_Agent / FSVI_Agent (+ oracle toolkit maths)
  * infotaxis_*   : reference Infotaxis_Agent.choose_action (in-repo einsum + Python loop) (+ oracle successors)
  * belief/evaluate/backup/vi : oracle toolkit only (parity UNPINNED upstream; see oracle/pomdp_toolkit)
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
from olfactory_navigation.agents import FSVI_Agent, QMDP_Agent, Infotaxis_Agent   # noqa: E402
from pomdp_toolkit import BeliefSet, ValueFunction                 # noqa: E402
from pomdp_toolkit import solvers as osolvers                      # noqa: E402

import olfactory_navigation_b200.synthetic as syn                  # noqa: E402  (pure NumPy generator, shared)

OUT = os.path.join(ROOT, 'tests', 'golden')
os.makedirs(OUT, exist_ok=True)


def save(name, **arrays):
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **arrays)
    print(f'{name:28s} {os.path.getsize(path) / 1024:8.1f} KiB')


def f32(x):
    'Round to float32 and back: fixtures hold values a float32 device path receives exactly.'
    return np.asarray(x, dtype=np.float32).astype(np.float64)


def test_beliefs(model, rng, n_random=6, n_walk=10):
    'A mix of beliefs: start, random walks from the start, Dirichlet points, border-heavy and corner masses.'
    S = model.state_count
    H, W = model.state_grid.shape
    rows = [model.start_probabilities.copy()]
    bs = BeliefSet(model, np.tile(model.start_probabilities, (n_walk, 1)))
    for _ in range(6):
        a = rng.integers(0, model.action_count, len(bs))
        o = rng.integers(0, max(model.observation_count - 1, 1), len(bs))
        nxt, prov = bs.update(a, o, raise_on_impossible_belief=False, return_provenance=True)
        if len(nxt) == 0:
            break
        bs = nxt
    rows += list(bs.belief_array)
    d = -np.log(1.0 - rng.random((n_random, S)))
    rows += list(d / d.sum(1, keepdims=True))
    border = np.zeros((H, W))
    border[0, :] = border[-1, :] = border[:, 0] = border[:, -1] = 1.0
    border += 0.01 * rng.random((H, W))
    rows.append((border / border.sum()).ravel())
    for (y, x) in [(0, 0), (0, W - 1), (H - 1, 0), (H - 1, W - 1), (H // 2, 0), (0, W // 2)]:
        e = np.zeros((H, W))
        e[y, x] = 1.0
        rows.append(e.ravel())
    return f32(np.array(rows))


def build(cfg_name, agent_cls):
    c = syn.config(cfg_name)
    env = Environment(**c['env_kwargs'])
    ag = agent_cls(env, thresholds=c['thresholds'])
    return c, env, ag


for cfg in ['T0', 'T1', 'T2', 'T3', 'C1']:
    rng = np.random.default_rng(syn.SEED + sum(map(ord, cfg)))
    c, env, ag = build(cfg, QMDP_Agent)
    model = ag.model
    S, A, O = model.state_count, model.action_count, model.observation_count
    H, W = env.shape

    # ---- converter (reference in-repo code) -----------------------------------------------------
    save(f'converter_{cfg}', shape=np.array(env.shape), thresholds=ag.thresholds, action_set=ag.action_set,
         reachable_states=model.reachable_states[:, :, 0].astype(np.int32), observation_table=model.observation_table,
         end_states=model.end_states.astype(np.int32), start_probabilities=model.start_probabilities,
         source_position=env.source_position, margins=env.margins, expected_rewards=model.expected_rewards_table)

    # ---- env step (reference in-repo code) ------------------------------------------------------------
    n = 200
    pos = np.stack([rng.integers(0, H, n), rng.integers(0, W, n)], axis=1)
    act = rng.integers(0, A, n)
    tsh = rng.integers(0, 3 * env.timesteps, n)
    step = 7
    new_pos = env.move(pos, ag.action_set[act])
    odor = env.get_observation(new_pos, time=(tsh + step))
    reached = env.source_reached(new_pos)
    ids = ag.discretize_observations(observation=odor, action=ag.action_set[act], source_reached=reached)
    save(f'envstep_{cfg}', pos=pos.astype(np.int32), act=act.astype(np.int32), time_shift=tsh.astype(np.int64), step=np.array(step),
         new_pos=new_pos.astype(np.int32), odor=odor, reached=reached, obs_id=ids.astype(np.int32))

    # ---- belief update (oracle toolkit) ------------------------------------------------------------
    B = test_beliefs(model, rng)
    nb = len(B)
    reps = 3
    b_in = np.tile(B, (reps, 1))
    a_in = rng.integers(0, A, len(b_in))
    o_in = rng.integers(0, O, len(b_in))
    u, p = BeliefSet(model, b_in).unnormalised_successors()
    idx = np.arange(len(b_in))
    z = p[idx, a_in, o_in]
    ok = z > 0
    out = np.zeros_like(b_in)
    out[ok] = u[idx, a_in, o_in][ok] / z[ok, None]
    save(f'belief_{cfg}', b_in=b_in.astype(np.float32), act=a_in.astype(np.int32), obs=o_in.astype(np.int32),
         b_out=out, norm=z, ok=ok)

    # ---- VI / evaluate (oracle toolkit) ----------------------------------------------------------------
    vf_q, hist = osolvers.VI.solve(model, horizon=1000, gamma=0.99, eps=1e-6)
    save(f'vi_{cfg}', Q=vf_q.alpha_vector_array, sweeps=np.array(len(hist.value_function_changes)),
         changes=np.array(hist.value_function_changes))
    G = 37
    base = vf_q.alpha_vector_array
    alphas = f32(np.vstack([base] + [base[rng.integers(0, A)] * (1 + 0.2 * rng.standard_normal(S)) for _ in range(G - A)]))
    acts = np.concatenate([np.arange(A), rng.integers(0, A, G - A)])
    vf = ValueFunction(model, alphas, acts)
    scores = B @ alphas.T
    values, actions = vf.evaluate_at(BeliefSet(model, B))
    srt = np.sort(scores, axis=1)
    save(f'evaluate_{cfg}', beliefs=B.astype(np.float32), alphas=alphas.astype(np.float32), alpha_actions=acts.astype(np.int32),
         values=values, argmax=np.argmax(scores, axis=1).astype(np.int32), actions=actions.astype(np.int32),
         margin=(srt[:, -1] - srt[:, -2]))

    # ---- backup (oracle toolkit) ----------------------------------------------------------------------
    Gb = 9
    vf_b = ValueFunction(model, alphas[:Gb], acts[:Gb])
    _, det = osolvers.backup(model, BeliefSet(model, B), vf_b, gamma=0.99, append=False, belief_dominance_prune=False, return_details=True)
    gam = osolvers.gamma_a_o(model, alphas[:Gb], 0.99)
    sc = np.einsum('bs,aogs->baog', B, gam)
    s_sorted = np.sort(sc, axis=3)
    q_sorted = np.sort(det['q'], axis=1)
    save(f'backup_{cfg}', beliefs=B.astype(np.float32), alphas=alphas[:Gb].astype(np.float32), discount=np.array(0.99),
         gamma_ao=gam.astype(np.float32) if cfg != 'C1' else np.zeros(0, np.float32),
         best_g=det['best_g'].astype(np.int32), best_a=det['best_a'].astype(np.int32), new_alpha=det['new_alpha'],
         new_value=det['new_value'], g_margin=(s_sorted[..., -1] - s_sorted[..., -2]), a_margin=(q_sorted[:, -1] - q_sorted[:, -2]))

    # ---- infotaxis (reference in-repo choose_action) ------------------------------------------------------
    if cfg != 'C1':
        it = Infotaxis_Agent(env, thresholds=c['thresholds'], use_reachability=True)
        Bi = B[:12]
        it.belief = BeliefSet(it.model, Bi)
        moves = it.choose_action()
        chosen = it.action_played.copy()
        # the per-action deltas, recomputed with the same formula (infotaxis_agent.py:257-293)
        bs = BeliefSet(it.model, Bi)
        Hb = bs.entropies
        u, p = bs.unnormalised_successors()
        dH = np.zeros((len(Bi), A))
        for b in range(len(Bi)):
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
        save(f'infotaxis_{cfg}', beliefs=Bi.astype(np.float32), chosen=chosen.astype(np.int32), moves=moves.astype(np.int32),
             delta_H=dH, entropies=Hb, margin=(d_sorted[:, -1] - d_sorted[:, -2]))

    # ---- run_test (reference driver + reference agents) ---------------------------------------------------
    n_sim, horizon = (48, 60) if cfg != 'C1' else (32, 120)
    for label, cls, train_kwargs in (('qmdp', QMDP_Agent, dict(expansions=1000)), ('fsvi', FSVI_Agent, dict(expansions=8, max_belief_growth=12))):
        agent = cls(env, thresholds=c['thresholds'])
        agent.rng = np.random.default_rng(syn.SEED + 1)
        agent.train(print_progress=False, print_stats=False, **train_kwargs)
        sim_rng = np.random.default_rng(syn.SEED + 2)
        starts = env.random_start_points(n_sim, sim_rng)
        tshift = sim_rng.integers(0, env.timesteps, n_sim)
        hist = run_test(agent, n=n_sim, start_points=starts, time_shift=tshift, horizon=horizon, print_progress=False,
                        print_stats=False, use_gpu=False, batches=1)
        vfa = agent.value_function
        save(f'runtest_{label}_{cfg}', alphas=vfa.alpha_vector_array.astype(np.float32), alpha_actions=vfa.actions.astype(np.int32),
             start_points=starts.astype(np.int32), time_shift=tshift.astype(np.int64), horizon=np.array(horizon),
             actions=np.array(hist.actions).astype(np.int8), positions=np.array(hist.positions).astype(np.int16),
             observations=np.array(hist.observations).astype(np.float32), reached_source=hist.reached_source,
             done_at_step=hist.done_at_step.astype(np.int32))

print('done')
