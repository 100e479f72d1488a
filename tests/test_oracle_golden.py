'''
The oracle (oracle/pomdp_toolkit, NumPy fp64) against the committed golden vectors, which were
produced in the build container by the UNMODIFIED reference code running on top of it
(oracle/make_golden.py).  CPU only.
'''
import numpy as np
import pytest

from conftest import ALL_CFGS, TINY, golden, make, oracle_model


@pytest.fixture(scope='module')
def tk():
    from oracle import refloader
    refloader.install_oracle_toolkit()
    import pomdp_toolkit
    return pomdp_toolkit


@pytest.mark.parametrize('cfg', ALL_CFGS)
def test_belief_update(cfg, tk):
    g = golden(f'belief_{cfg}')
    _, agent = make(cfg)
    m = oracle_model(agent.model)
    bs = tk.BeliefSet(m, g['b_in'].astype(np.float64))
    new, prov = bs.update(g['act'], g['obs'], raise_on_impossible_belief=False, return_provenance=True)
    assert np.array_equal(prov[:, 0], np.flatnonzero(g['ok']))
    np.testing.assert_allclose(new.belief_array, g['b_out'][g['ok']], rtol=1e-12, atol=0)
    assert (~g['ok']).sum() > 0, 'fixture must contain impossible updates'
    np.testing.assert_allclose(new.belief_array.sum(1), 1.0, rtol=1e-12)


@pytest.mark.parametrize('cfg', ALL_CFGS)
def test_vi_stops_at_918(cfg, tk):
    'The one numeric pin the reference offers: "Converged in 918 iterations" (gamma=.99, eps=1e-6).'
    g = golden(f'vi_{cfg}')
    _, agent = make(cfg)
    vf, hist = tk.solvers.VI.solve(oracle_model(agent.model), horizon=1000, gamma=0.99, eps=1e-6)
    assert len(hist.value_function_changes) == 918 == int(g['sweeps'])
    np.testing.assert_allclose(vf.alpha_vector_array, g['Q'], rtol=1e-12)


@pytest.mark.parametrize('cfg', ALL_CFGS)
def test_evaluate_and_backup(cfg, tk):
    _, agent = make(cfg)
    m = oracle_model(agent.model)
    g = golden(f'evaluate_{cfg}')
    vf = tk.ValueFunction(m, g['alphas'].astype(np.float64), g['alpha_actions'])
    v, a = vf.evaluate_at(tk.BeliefSet(m, g['beliefs'].astype(np.float64)))
    np.testing.assert_allclose(v, g['values'], rtol=1e-12)
    assert np.array_equal(a, g['actions'])
    g = golden(f'backup_{cfg}')
    vf = tk.ValueFunction(m, g['alphas'].astype(np.float64), np.zeros(len(g['alphas']), dtype=int))
    _, det = tk.solvers.backup(m, tk.BeliefSet(m, g['beliefs'].astype(np.float64)), vf, gamma=float(g['discount']),
                               belief_dominance_prune=False, return_details=True)
    assert np.array_equal(det['best_g'], g['best_g']) and np.array_equal(det['best_a'], g['best_a'])
    np.testing.assert_allclose(det['new_alpha'], g['new_alpha'], rtol=1e-12, atol=1e-300)


@pytest.mark.parametrize('cfg', TINY)
def test_infotaxis_closed_form_matches_reference_loop(cfg, tk):
    '''
    dH_a = H(b) - sum_o p log p + sum b_a C + sum b_a log b_a (the form the fused kernel uses) against
    the deltas frozen from the reference's own choose_action loop.
    '''
    g = golden(f'infotaxis_{cfg}')
    _, agent = make(cfg)
    m = oracle_model(agent.model)
    B = g['beliefs'].astype(np.float64)
    bs = tk.BeliefSet(m, B)
    u, p = bs.unnormalised_successors()
    with np.errstate(divide='ignore', invalid='ignore'):
        xlogx = lambda t: np.where(t > 0, t * np.log(t), 0.0)   # noqa: E731
        C = np.sum(xlogx(m.observation_table), axis=2)          # (S, A)
        b_a = np.sum(u, axis=2)                                  # (n, A, S)
        dH = bs.entropies[:, None] - np.sum(xlogx(p), axis=2) + np.einsum('nas,sa->na', b_a, C) + np.sum(xlogx(b_a), axis=2)
    np.testing.assert_allclose(dH, g['delta_H'], rtol=1e-9, atol=1e-12)
    assert np.array_equal(np.argmax(dH, axis=1), g['chosen'])


def test_fsvi_trajectory_draw_order(tk):
    'Oracle and product sample identical (action, observation) sequences from identical generators.'
    from olfactory_navigation_b200 import solvers as dsolvers
    _, agent = make('T1')
    m = oracle_model(agent.model)
    Q = golden('vi_T1')['Q']
    for seed in range(5):
        s1 = tk.solvers.FSVI.trajectory(m, m.start_probabilities, Q, 12, np.random.default_rng(seed))
        s2 = dsolvers.FSVI.trajectory(agent.model, agent.model.start_probabilities, Q, 12, np.random.default_rng(seed))
        assert s1 == s2 and len(s1) > 0


@pytest.mark.parametrize('cfg', ['T0', 'T1', 'T2', 'T3', 'C1'])
@pytest.mark.parametrize('label', ['qmdp', 'fsvi'])
def test_oracle_sim_loop_reproduces_reference_run_test(cfg, label, tk):
    'oracle/sim.py (restated run loop, the CPU baseline) == the reference\'s run_test, step by step, bit exact.'
    import olfactory_navigation_b200.synthetic as syn
    from oracle import sim
    g = golden(f'runtest_{label}_{cfg}')
    c = syn.CONFIGS[cfg]
    data = syn.plume_frames(c['T'], c['h'], c['w'], source_y=c['source'][0], seed=syn.SEED)
    thr = np.sort(np.array([-np.inf] + (c['thresholds'] if isinstance(c['thresholds'], list) else [c['thresholds']]) + [np.inf]))
    setup = sim.build_model(data, c['margins'], c['source'], c['radius'], c['boundary'], thr, c['start_zone'],
                            0.0 if c['start_zone'] == 'odor_present' else None)
    conv = golden(f'converter_{cfg}')
    assert np.array_equal(setup['model'].observation_table, conv['observation_table'])
    assert np.array_equal(setup['model'].reachable_states[:, :, 0], conv['reachable_states'])
    assert np.array_equal(setup['model'].start_probabilities, conv['start_probabilities'])
    vf = tk.ValueFunction(setup['model'], g['alphas'].astype(np.float64), g['alpha_actions'])
    out = sim.run_test(setup, vf, g['start_points'].astype(int), g['time_shift'], int(g['horizon']))
    # the fixture's value function was rounded to float32: trajectories may differ at near-ties only
    same = (out['done_at_step'] == g['done_at_step']) & (out['reached_source'] == g['reached_source'])
    assert same.mean() >= 0.9
    sims, act, pos, odor = out['records'][0]
    assert np.array_equal(pos, g['positions'][0][sims]) and np.array_equal(setup['action_set'][act], g['actions'][0][sims])
    assert np.array_equal(odor.astype(np.float32), g['observations'][0][sims])


def test_oracle_next_solvers_run_and_improve():
    'SSEA / GER / Perseus / HSVI restatements: deterministic under a seed, belief sets grow, value functions stay valid lower bounds.'
    from conftest import make, oracle_model
    import pomdp_toolkit as tk
    _, agent = make('T0')
    om = oracle_model(agent.model)
    q, _ = tk.solvers.VI.solve(om, horizon=1000, gamma=0.99, eps=1e-6, print_progress=False)
    upper = q.alpha_vector_array                      # Q_MDP is an upper bound of the POMDP value
    for name, kw in [('PBVI_SSEA', {}), ('PBVI_GER', {}), ('Perseus', {}), ('HSVI', {'epsilon': 0.05})]:
        runs = [getattr(tk.solvers, name).solve(om, expansions=3, max_belief_growth=4, rng=np.random.default_rng(2), print_progress=False, **kw) for _ in range(2)]
        a0, a1 = runs[0][0].alpha_vector_array, runs[1][0].alpha_vector_array
        assert a0.shape == a1.shape and np.array_equal(a0, a1), name
        hist = runs[0][1]
        assert hist.beliefs_counts == sorted(hist.beliefs_counts) and hist.beliefs_counts[-1] > 1
        B = hist.final_belief_set.belief_array
        assert np.all(np.max(B @ a0.T, axis=1) <= np.max(B @ upper.T, axis=1) + 1e-9), name
