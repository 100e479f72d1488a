'''
SURVEY §8 a6 / f2: every belief-point selection strategy — FSVI (the benchmarked solver), PBVI_RA, PBVI_SSRA, PBVI_SSGA, PBVI_SSEA,
PBVI_GER, Perseus, HSVI — and level-2 pruning on the device against the oracle's restatement, same generator on both sides (every
stochastic choice is an inverse CDF on rng.random() in a documented order, so identical seeds give identical (action, observation)
sequences).  One strict contract for all of them: same belief counts per expansion, same belief points, same number of alpha vectors,
same value function (as a function) within 1e-5 relative.
Reference call sites: agents/fsvi_agent.py:201-224, pbvi_ra_agent.py:168, pbvi_ssra_agent.py:170, pbvi_ssga_agent.py:176-198,
pbvi_ssea_agent.py:170, pbvi_ger_agent.py:170, perseus_agent.py:176-196, hsvi_agent.py:184-208.
'''
import numpy as np
import pytest
import torch

from conftest import make, oracle_model

pytestmark = pytest.mark.gpu

CASES = [('FSVI', {}), ('PBVI_RA', {}), ('PBVI_SSRA', {}), ('PBVI_SSGA', {'epsilon': 0.3}), ('PBVI_SSGA', {'epsilon': 0.9, 'epsilon_decay': 0.5}),
         ('PBVI_SSEA', {}), ('PBVI_GER', {}), ('Perseus', {}), ('Perseus', {'use_policy_to_choose_actions': True}), ('HSVI', {'epsilon': 0.05})]


def _values_on(points, alphas):
    return np.max(points @ alphas.T, axis=1)


@pytest.mark.parametrize('cfg', ['T0', 'T1', 'C1'])
@pytest.mark.parametrize('name,kw', CASES)
def test_solver_matches_oracle(cfg, name, kw):
    import olfactory_navigation_b200 as onb
    from oracle import refloader
    refloader.install_oracle_toolkit()
    import pomdp_toolkit as tk
    _, agent = make(cfg)
    m, om = agent.model, oracle_model(agent.model)
    common = dict(expansions=3, max_belief_growth=4, gamma=0.99, print_progress=False, print_stats=False)
    ref = getattr(tk.solvers, name).solve(om, rng=np.random.default_rng(7), **common, **kw)
    got = getattr(onb.solvers, name).solve(m, rng=np.random.default_rng(7), **common, **kw)
    ref_vf, ref_hist, got_vf, got_hist = ref[0], ref[1], got[0], got[1]
    # same belief points were selected (sizes per expansion and the points themselves)
    assert got_hist.beliefs_counts == ref_hist.beliefs_counts
    gb, rb = got_hist.final_belief_set.belief_array, ref_hist.final_belief_set.belief_array
    np.testing.assert_allclose(gb, rb, rtol=2e-5, atol=1e-9)
    # same value function as a function: values at the selected points and at random simplex points within 1e-5 relative
    rng = np.random.default_rng(3)
    pts = -np.log(1.0 - rng.random((64, m.state_count)))
    pts = np.vstack([rb, pts / pts.sum(1, keepdims=True)])
    vr, vg = _values_on(pts, ref_vf.alpha_vector_array), _values_on(pts, got_vf.alpha_vector_array)
    np.testing.assert_allclose(vg, vr, rtol=1e-5, atol=1e-7)
    assert len(got_vf) == len(ref_vf)


@pytest.mark.parametrize('name', ['FSVI', 'PBVI_SSRA'])
def test_limit_value_function_size_and_convergence_stop(name):
    '''
    `limit_value_function_size` (random drop of max_belief_growth vectors once the set is larger, same rng draw on both sides) and
    `convergence_stop` (stop at the first expansion whose value change on the belief set is below eps): same history as the oracle.
    Reference train() arguments: agents/pbvi_agent.py:584-658.
    '''
    import olfactory_navigation_b200 as onb
    from oracle import refloader
    refloader.install_oracle_toolkit()
    import pomdp_toolkit as tk
    _, agent = make('T1')
    m, om = agent.model, oracle_model(agent.model)
    common = dict(expansions=6, max_belief_growth=3, gamma=0.99, limit_value_function_size=6, print_progress=False, print_stats=False)
    ref = getattr(tk.solvers, name).solve(om, rng=np.random.default_rng(5), **common)
    got = getattr(onb.solvers, name).solve(m, rng=np.random.default_rng(5), **common)
    assert got[1].alpha_vector_counts == ref[1].alpha_vector_counts
    assert got[1].prune_counts == ref[1].prune_counts and max(ref[1].prune_counts) > 0
    rb = ref[1].final_belief_set.belief_array
    np.testing.assert_allclose(got[1].final_belief_set.belief_array, rb, rtol=2e-5, atol=1e-9)
    np.testing.assert_allclose(_values_on(rb, got[0].alpha_vector_array), _values_on(rb, ref[0].alpha_vector_array), rtol=1e-5, atol=1e-7)
    # convergence_stop with a loose eps: both stop after the same number of expansions, before the budget is used up
    common = dict(expansions=12, max_belief_growth=2, gamma=0.5, eps=5e-3, convergence_stop=True, print_progress=False, print_stats=False)
    ref = getattr(tk.solvers, name).solve(om, rng=np.random.default_rng(6), **common)
    got = getattr(onb.solvers, name).solve(m, rng=np.random.default_rng(6), **common)
    assert len(got[1].value_function_changes) == len(ref[1].value_function_changes) < 12
    np.testing.assert_allclose(got[1].value_function_changes, ref[1].value_function_changes, rtol=1e-4, atol=1e-6)


def test_prune_level_2_drops_dominated_vectors():
    import olfactory_navigation_b200 as onb
    from olfactory_navigation_b200.solvers import prune_alphas
    _, agent = make('T0')
    m = agent.model
    dev = torch.device('cuda', 0)
    rng = np.random.default_rng(0)
    Sp = m.padded_states
    base = torch.as_tensor(rng.random((6, Sp)).astype(np.float32), device=dev)
    alphas = torch.cat([base, base[:2] - 0.1, base[3:4].clone(), base[4:5] + 0.0])      # two dominated rows, two duplicates
    alphas[6, 5] = base[0, 5]                                                           # still dominated (>= everywhere, > somewhere)
    actions = torch.arange(alphas.shape[0], dtype=torch.int32, device=dev) % 4
    actions[8], actions[9] = actions[3], actions[4]
    a1, c1 = prune_alphas(alphas, actions, 1)
    assert a1.shape[0] == 8
    a2, c2 = prune_alphas(alphas, actions, 2)
    assert a2.shape[0] == 6 and torch.equal(a2, base)
    # oracle agrees
    from oracle import refloader
    refloader.install_oracle_toolkit()
    import pomdp_toolkit as tk
    om = oracle_model(m)
    H, W = m.grid_shape
    dense = alphas.view(-1, H, Sp // H)[:, :, :W].reshape(alphas.shape[0], -1).double().cpu().numpy()
    ovf = tk.ValueFunction(om, dense, actions.cpu().numpy())
    ovf.prune(2)
    assert len(ovf) == 6


def test_flavour_agents_train_and_run():
    'Every flavour agent trains through the reference train() signature and drives run_test.'
    import olfactory_navigation_b200 as onb
    c = onb.synthetic.config('T1')
    env = onb.Environment(**c['env_kwargs'])
    for cls, kw in ((onb.agents.PBVI_RA_Agent, {}), (onb.agents.PBVI_SSRA_Agent, {}), (onb.agents.PBVI_SSGA_Agent, {}),
                    (onb.agents.PBVI_SSEA_Agent, {}), (onb.agents.PBVI_GER_Agent, {}), (onb.agents.Perseus_Agent, {'use_policy_to_choose_actions': True}),
                    (onb.agents.HSVI_Agent, {'epsilon': 0.05})):
        agent = cls(env, thresholds=c['thresholds'], rng=5)
        hist = agent.train(expansions=3, max_belief_growth=4, prune_level=2, prune_interval=1, print_progress=False, print_stats=False, **kw)
        assert agent.trained and len(agent.value_function) >= 1 and len(hist.alpha_vector_counts) == 3
        sim = onb.run_test(agent, n=16, horizon=30, print_progress=False, print_stats=False)
        assert len(sim.done_at_step) == 16
