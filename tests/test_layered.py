'''
Layered environments (reference environment.py:202-235, agent.py:281-283,413-417,
agents/model_based_util/environment_converter.py:54-85, simulation.py:1456-1463) against golden vectors frozen from the
UNMODIFIED reference (oracle/make_golden_layered.py): converter tables with per-action observation rows, the per-step host
API with a layer per action, and run_test trajectories of QMDP / FSVI agents on the device.
'''
import numpy as np
import pytest

from conftest import golden


def _make(per_layer: bool = False, agent_cls: str = 'QMDP_Agent'):
    import olfactory_navigation_b200 as onb
    c = onb.synthetic.layered_config('L1')
    env = onb.Environment(**c['env_kwargs'])
    thr = c['thresholds_per_layer'] if per_layer else c['thresholds']
    return env, getattr(onb.agents, agent_cls)(env, thresholds=thr), c


@pytest.mark.parametrize('tag', ['L1', 'L1pl'])
def test_layered_converter_tables(tag):
    g = golden(f'converter_{tag}')
    env, agent, _ = _make(per_layer=(tag == 'L1pl'))
    m = agent.model
    assert tuple(g['shape']) == env.shape and list(g['layer_labels']) == env.layer_labels
    assert np.array_equal(g['thresholds'], agent.thresholds)
    assert np.array_equal(g['action_set'], agent.action_set) and list(g['action_labels']) == agent.action_labels
    assert np.array_equal(g['reachable_states'], m.reachable_states[:, :, 0])
    assert np.array_equal(g['observation_table'], m.observation_table)          # bit exact, rows differ per action (layer)
    assert not np.array_equal(m.observation_table[:, 0], m.observation_table[:, 4])
    assert np.array_equal(g['end_states'], m.end_states)
    assert np.array_equal(g['start_probabilities'], m.start_probabilities)


@pytest.mark.parametrize('tag', ['L1', 'L1pl'])
def test_layered_env_step_numpy(tag):
    g = golden(f'envstep_{tag}')
    env, agent, _ = _make(per_layer=(tag == 'L1pl'))
    vec = agent.action_set[g['act']]
    new_pos = env.move(g['pos'].astype(int), vec[:, 1:])
    assert np.array_equal(new_pos, g['new_pos'])
    odor = env.get_observation(new_pos, time=g['time_shift'] + int(g['step']), layer=vec[:, 0])
    assert np.array_equal(odor, g['odor'])
    reached = env.source_reached(new_pos)
    ids = agent.discretize_observations(odor, vec, reached)
    assert np.array_equal(ids, g['obs_id'])


def test_layered_dict_thresholds_and_save_load(tmp_path):
    import olfactory_navigation_b200 as onb
    env, _, c = _make()
    ag = onb.Agent(env, thresholds={'ground': [0.3, 0.6], 'air': [0.5, 0.2]})
    assert ag.thresholds.tolist() == [[-np.inf, 0.3, 0.6, np.inf], [-np.inf, 0.2, 0.5, np.inf]]
    env.save(str(tmp_path), save_arrays=True)
    again = onb.Environment.load(env.saved_at)
    assert again.has_layers and again.layer_labels == env.layer_labels and np.array_equal(again.data, env.data)


@pytest.mark.gpu
@pytest.mark.parametrize('label', ['qmdp', 'fsvi'])
def test_layered_run_test_matches_reference_run(label):
    'Device run_test on a layered environment (per-action likelihood tables in the belief kernel, layer pick in the env step).'
    import olfactory_navigation_b200 as onb
    from test_sim_gpu import _compare_runs
    g = golden(f'runtest_{label}_L1')
    env, _, c = _make()
    agent = onb.agents.PBVI_Agent(env, thresholds=c['thresholds'])
    agent.value_function = onb.ValueFunction(agent.model, g['alphas'].astype(np.float64), g['alpha_actions'])
    hist = onb.run_test(agent, n=len(g['start_points']), start_points=g['start_points'].astype(int), time_shift=g['time_shift'],
                        horizon=int(g['horizon']), print_progress=False, print_stats=False, batches=1)
    same = _compare_runs(hist, g)
    obs = np.array(hist.observations)
    for k in np.flatnonzero(same)[:10]:
        steps = len(g['actions']) if g['done_at_step'][k] < 0 else g['done_at_step'][k]
        assert np.array_equal(obs[:steps, k].astype(np.float32), g['observations'][:steps, k])


@pytest.mark.gpu
def test_layered_per_layer_thresholds_device_equals_host_protocol():
    'Per-layer thresholds (unreachable through the reference constructor): device run == the reference protocol loop on the host API.'
    import olfactory_navigation_b200 as onb
    env, agent, c = _make(per_layer=True, agent_cls='QMDP_Agent')
    agent.train(print_progress=False, print_stats=False)
    rng = np.random.default_rng(3)
    n, horizon = 40, 40
    starts = env.random_start_points(n, rng)
    tshift = rng.integers(0, env.timesteps, n)
    dev = onb.run_test(agent, n=n, start_points=starts, time_shift=tshift, horizon=horizon, print_progress=False, print_stats=False, batches=1)
    pos, ts = starts.copy(), tshift.copy()
    agent.initialize_state(n)
    hist = onb.SimulationHistory(pos, env, agent, ts, horizon)
    for i in range(horizon):
        action = agent.choose_action()
        pos = env.move(pos, action[:, 1:])
        observation = env.get_observation(pos, time=ts + i, layer=action[:, 0])
        reached = env.source_reached(pos)
        ok = agent.update_state(action=action, observation=observation, source_reached=reached)
        term = reached | ~ok
        hist.add_step(actions=action, next_positions=pos, observations=observation, reached_source=reached, interupt=term)
        pos, ts = pos[~term], ts[~term]
        agent.kill(simulations_to_kill=term)
        if len(pos) == 0:
            break
    assert np.array_equal(np.array(dev.done_at_step), np.array(hist.done_at_step))
    assert np.array_equal(np.array(dev.reached_source), np.array(hist.reached_source))


@pytest.mark.gpu
def test_layered_belief_update_and_infotaxis_against_oracle():
    'Per-action likelihood tables: fused belief step and the infotaxis expected-entropy choice against the oracle toolkit.'
    import olfactory_navigation_b200 as onb
    from conftest import oracle_model
    from oracle import refloader
    refloader.install_oracle_toolkit()
    import pomdp_toolkit as tk
    env, agent, c = _make()
    m, om = agent.model, oracle_model(agent.model)
    rng = np.random.default_rng(5)
    n = 96
    d = -np.log(1.0 - rng.random((n, m.state_count)))
    B = (d / d.sum(1, keepdims=True)).astype(np.float32).astype(np.float64)
    act, obs = rng.integers(0, m.action_count, n), rng.integers(0, m.observation_count - 1, n)
    new, prov = onb.BeliefSet(m, B).update(act, obs, raise_on_impossible_belief=False, return_provenance=True)
    ref, rprov = tk.BeliefSet(om, B).update(act, obs, raise_on_impossible_belief=False, return_provenance=True)
    assert np.array_equal(prov[:, 0], rprov[:, 0])
    np.testing.assert_allclose(new.belief_array, ref.belief_array, rtol=1e-5, atol=1e-30)

    it = onb.agents.Infotaxis_Agent(env, thresholds=c['thresholds'])
    it.initialize_state(n, onb.BeliefSet(it.model, B))
    ids = it.choose_action_device(return_delta_H=True).cpu().numpy()
    bs = tk.BeliefSet(om, B)
    u, p = bs.unnormalised_successors()
    Hb = bs.entropies
    dH = np.zeros((n, m.action_count))
    for b in range(n):
        for a in range(m.action_count):
            acc = 0.0
            for o in range(m.observation_count):
                if p[b, a, o] > 0:
                    w = u[b, a, o] / p[b, a, o]
                    nz = w > 0
                    acc += p[b, a, o] * (-np.sum(w[nz] * np.log(w[nz])))
            dH[b, a] = Hb[b] - acc
    np.testing.assert_allclose(it.last_delta_H.cpu().numpy(), dH, rtol=1e-3, atol=2e-4)
    ds = np.sort(dH, axis=1)
    clear = (ds[:, -1] - ds[:, -2]) > 1e-3
    assert np.array_equal(ids[clear], np.argmax(dH, axis=1)[clear])
