'''
Environment stepping, infotaxis choice, agents and run_test on the device against the golden vectors
frozen from the UNMODIFIED reference's run_test / agents / Infotaxis loop (oracle/make_golden.py).
'''
import numpy as np
import pytest
import torch

from conftest import ALL_CFGS, TINY, golden, make

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('cfg', ALL_CFGS)
@pytest.mark.parametrize('quirk', [True, False])
def test_env_step_bit_exact(cfg, quirk):
    from olfactory_navigation_b200._lib import lib, ptr, stream, check
    g = golden(f'envstep_{cfg}')
    env, agent = make(cfg)
    dev = torch.device('cuda', 0)
    n = len(g['act'])
    pos = torch.as_tensor(g['pos'], dtype=torch.int32, device=dev).contiguous()
    act = torch.as_tensor(g['act'], dtype=torch.int32, device=dev)
    moves = torch.as_tensor(agent.action_set, dtype=torch.int32, device=dev).contiguous()
    step = int(g['step'])
    ts = g['time_shift']
    if quirk:       # reference semantics: sorted effective times (Environment.reference_time_quirk)
        t_eff, t_step = np.sort((ts + step) % env.timesteps), 0
        want_odor, want_id = g['odor'], g['obs_id']
    else:
        t_eff, t_step = ts, step
        env.reference_time_quirk = False
        try:
            want_odor = env.get_observation(g['new_pos'].astype(int), time=ts + step)
        finally:
            env.reference_time_quirk = True
        want_id = agent.discretize_observations(want_odor, None, g['reached'])
    tsh = torch.as_tensor(t_eff, dtype=torch.int64, device=dev)
    thr = torch.as_tensor(agent.thresholds, dtype=torch.float64, device=dev)
    odor = torch.empty(n, dtype=torch.float64, device=dev)
    oid = torch.empty(n, dtype=torch.int32, device=dev)
    reached = torch.empty(n, dtype=torch.uint8, device=dev)
    check(lib.olfnav_env_step(env.device_handle(0), n, ptr(pos), ptr(act), ptr(moves), ptr(tsh), t_step, ptr(thr), len(agent.thresholds),
                              ptr(odor), ptr(oid), ptr(reached), stream()))
    assert np.array_equal(pos.cpu().numpy(), g['new_pos'])
    assert np.array_equal(odor.cpu().numpy(), want_odor)
    assert np.array_equal(reached.cpu().numpy().astype(bool), g['reached'])
    assert np.array_equal(oid.cpu().numpy(), want_id)


def _unsure(model):
    'Agents of the last infotaxis choice that went through the float64 re-evaluation.'
    import ctypes
    from olfactory_navigation_b200._lib import lib, stream, check
    k = ctypes.c_int32(0)
    check(lib.olfnav_infotaxis_last_unsure(model.handle(0), ctypes.byref(k), stream()))
    return int(k.value)


@pytest.mark.parametrize('cfg', ALL_CFGS)
def test_infotaxis_vs_reference_loop(cfg):
    '''
    Expected-entropy choice against the UNMODIFIED reference loop (agents/infotaxis_agent.py:257-293; goldens from oracle/make_golden.py
    and oracle/make_golden_infotaxis.py, C1 = BASELINE configs[0] shape with gaps down to 6e-6 between the two best actions).
    ΔH within 1e-5 relative (float64 kernel when the values are asked for); the decision of the production path (FP32 tensor-core sums +
    float64 re-evaluation of the near-tied agents) must be the reference's wherever the gap exceeds 1e-6 of the largest |ΔH| of the row.
    '''
    import olfactory_navigation_b200 as onb
    g = golden(f'infotaxis_{cfg}')
    env, _ = make(cfg)
    c = onb.synthetic.config(cfg)
    it = onb.agents.Infotaxis_Agent(env, thresholds=c['thresholds'])
    B = g['beliefs'].astype(np.float64)
    scale_row = np.abs(g['delta_H']).max(1)
    clear = g['margin'] > 1e-6 * scale_row
    assert clear.sum() >= 0.7 * len(clear)
    it.initialize_state(len(B), onb.BeliefSet(it.model, B))
    ids_exact = it.choose_action_device(return_delta_H=True).cpu().numpy()
    np.testing.assert_allclose(it.last_delta_H.cpu().numpy(), g['delta_H'], rtol=1e-5, atol=1e-7 * scale_row.max())
    assert np.array_equal(ids_exact[clear], g['chosen'][clear])
    ids = it.choose_action_device().cpu().numpy()                       # production path
    assert np.array_equal(ids[clear], g['chosen'][clear])
    assert _unsure(it.model) >= int((g['margin'] < 1e-7).sum())         # exact ties (point masses) are always re-evaluated
    H = onb.BeliefSet(it.model, B).entropies
    np.testing.assert_allclose(H, g['entropies'], rtol=1e-4, atol=1e-5)
    # reference protocol: movement vectors
    it.initialize_state(len(B), onb.BeliefSet(it.model, B))
    mv = it.choose_action()
    assert np.array_equal(mv[clear], g['moves'][clear])


def test_infotaxis_c4_shape_vs_live_oracle():
    '''
    BASELINE configs[3] shape (83 x 371, wrap_vertical): 64 agents at different stages of a run (the shared start belief, beliefs
    after 1 ... 12 steps with and without detections) against the oracle's float64 expected-entropy reduction.
    Production path decisions identical wherever the float64 gap exceeds 1e-6 of the row's largest |ΔH|; ΔH within 1e-5 relative.
    '''
    import olfactory_navigation_b200 as onb
    import replay_check
    sim, setup = replay_check.oracle_setup('C2')
    import pomdp_toolkit as tk
    om = setup['model']
    _, agent = make('C2')
    it = onb.agents.Infotaxis_Agent(agent.environment, thresholds=onb.synthetic.config('C2')['thresholds'])
    rng = np.random.default_rng(21)
    rows = [om.start_probabilities.copy()]
    bs = tk.BeliefSet(om, np.tile(om.start_probabilities, (21, 1)))
    for step in range(12):
        a = rng.integers(0, 4, len(bs))
        o = (rng.random(len(bs)) < (0.5 if step in (3, 7) else 0.0)).astype(int)
        nxt, prov = bs.update(a, o, raise_on_impossible_belief=False, return_provenance=True)
        bs = nxt if len(nxt) >= 8 else bs
        if step % 4 == 3:
            rows += list(bs.belief_array)
    B = np.asarray(rows[:64], dtype=np.float32).astype(np.float64)
    B = (B / B.sum(1, keepdims=True)).astype(np.float32).astype(np.float64)
    dH = replay_check.infotaxis_delta_H(om, B)
    srt = np.sort(dH, axis=1)
    scale_row = np.abs(dH).max(1)
    clear = (srt[:, -1] - srt[:, -2]) > 1e-6 * scale_row
    assert clear.mean() > 0.5
    it.initialize_state(len(B), onb.BeliefSet(it.model, B))
    ids_exact = it.choose_action_device(return_delta_H=True).cpu().numpy()
    np.testing.assert_allclose(it.last_delta_H.cpu().numpy(), dH, rtol=1e-5, atol=1e-7 * scale_row.max())
    assert np.array_equal(ids_exact[clear], np.argmax(dH, axis=1)[clear])
    ids = it.choose_action_device().cpu().numpy()
    assert np.array_equal(ids[clear], np.argmax(dH, axis=1)[clear])
    assert np.array_equal(ids, ids_exact) or (ids != ids_exact).sum() <= (~clear).sum()


def _compare_runs(hist, g, cfg=None, min_same=0.85, quirk=True):
    '''
    Device run against the reference's golden run (north star: "chosen actions bit-exact except at documented ties").
    1. Every decision, environment transition and termination of the device run is re-derived by the oracle from the run's own
       history (tests/replay_check.py): each recorded action must be a float64 argmax of b . alpha or lose to it by < 1e-5 relative.
    2. If the run made no tie decision at all it must equal the golden run everywhere; otherwise the two may only differ from the
       first tie step on (a flipped near-tie sends that agent elsewhere and, through the reference's sorted-time indexing, changes
       the other agents' observations from then on), and identical trajectories must still dominate.
    '''
    import replay_check
    acts, poss = hist.actions, hist.positions
    r_acts, r_poss = g['actions'].astype(int), g['positions'].astype(int)
    div = replay_check.first_divergence(acts, poss, r_acts, r_poss)
    if cfg is None:
        # configurations oracle/sim.py does not model (layered environments): trajectory comparison only
        out = {'ties': -1, 'first_tie_step': -1}
    else:
        out = replay_check.replay(acts, poss, hist.observations, hist.done_at_step, hist.reached_source, cfg, g['alphas'].astype(np.float64),
                                  g['alpha_actions'], g['start_points'].astype(int), g['time_shift'], quirk=quirk)
    if out['ties'] == 0:
        assert div is None and len(acts) == len(r_acts), f'no tie decision in the device run, yet it differs from the reference run at step {div}'
        assert np.array_equal(hist.done_at_step, g['done_at_step']) and np.array_equal(hist.reached_source, g['reached_source'])
    elif div is not None and cfg is not None:
        assert out['first_tie_step'] <= div, f'runs diverge at step {div}, before the first documented tie (step {out["first_tie_step"]})'
    ref_done, ref_reached = g['done_at_step'], g['reached_source']
    n = len(ref_done)
    acts, poss = np.array(acts), np.array(poss)
    same = np.zeros(n, dtype=bool)
    for k in range(n):
        if hist.done_at_step[k] != ref_done[k] or hist.reached_source[k] != ref_reached[k]:
            continue
        steps = len(r_acts) if ref_done[k] < 0 else ref_done[k]
        if len(acts) < steps:
            continue
        same[k] = np.array_equal(acts[:steps, k], r_acts[:steps, k]) and np.array_equal(poss[:steps, k], r_poss[:steps, k])
    assert same.mean() >= min_same, f'only {same.mean():.2%} of trajectories identical to the reference run'
    return same


@pytest.mark.parametrize('cfg', ALL_CFGS)
@pytest.mark.parametrize('label', ['qmdp', 'fsvi'])
def test_run_test_matches_reference_run(cfg, label):
    'Same start points / time shifts / value function as the reference run -> same trajectories.'
    import olfactory_navigation_b200 as onb
    g = golden(f'runtest_{label}_{cfg}')
    env, _ = make(cfg)
    c = onb.synthetic.config(cfg)
    agent = onb.agents.PBVI_Agent(env, thresholds=c['thresholds'])
    agent.value_function = onb.ValueFunction(agent.model, g['alphas'].astype(np.float64), g['alpha_actions'])
    hist = onb.run_test(agent, n=len(g['start_points']), start_points=g['start_points'].astype(int), time_shift=g['time_shift'],
                        horizon=int(g['horizon']), print_progress=False, print_stats=False, batches=1)
    same = _compare_runs(hist, g, cfg)
    # observations of identical trajectories are bit exact (float64 odor data on the device)
    obs = np.array(hist.observations)
    for k in np.flatnonzero(same)[:10]:
        steps = len(g['actions']) if g['done_at_step'][k] < 0 else g['done_at_step'][k]
        assert np.array_equal(obs[:steps, k].astype(np.float32), g['observations'][:steps, k])
    # batching splits and re-concatenates like the reference (simulation.py:1370-1413)
    hist_b = onb.run_test(agent, n=len(g['start_points']), start_points=g['start_points'].astype(int), time_shift=g['time_shift'],
                          horizon=int(g['horizon']), print_progress=False, print_stats=False, batches=3)
    assert hist_b.n == hist.n and len(hist_b.done_at_step) == hist.n


def test_reference_protocol_loop_equals_device_loop():
    'Driving the agent through the NumPy protocol exactly like the reference run_test (simulation.py:1452-1497).'
    import olfactory_navigation_b200 as onb
    g = golden('runtest_fsvi_T1')
    env, _ = make('T1')
    c = onb.synthetic.config('T1')
    agent = onb.agents.PBVI_Agent(env, thresholds=c['thresholds'])
    agent.value_function = onb.ValueFunction(agent.model, g['alphas'].astype(np.float64), g['alpha_actions'])
    n, horizon = len(g['start_points']), int(g['horizon'])
    pos, tshift = g['start_points'].astype(int), g['time_shift'].astype(int)
    agent.initialize_state(n)
    hist = onb.SimulationHistory(pos, env, agent, tshift, horizon)
    for i in range(horizon):
        action = agent.choose_action()
        pos = env.move(pos, action)
        observation = env.get_observation(pos, time=tshift + i)
        reached = env.source_reached(pos)
        ok = agent.update_state(action=action, observation=observation, source_reached=reached)
        term = reached | ~ok
        hist.add_step(actions=action, next_positions=pos, observations=observation, reached_source=reached, interupt=term)
        pos, tshift = pos[~term], tshift[~term]
        agent.kill(simulations_to_kill=term)
        if len(pos) == 0:
            break
    _compare_runs(hist, g, 'T1')


def test_failed_updates_and_extra_kills():
    'Impossible observation -> ok False, agent terminated; kill() with extra victims compacts the rest.'
    import olfactory_navigation_b200 as onb
    env, _ = make('T0')
    c = onb.synthetic.config('T0')
    agent = onb.agents.QMDP_Agent(env, thresholds=c['thresholds'])
    agent.train(print_progress=False, print_stats=False)
    n = 6
    agent.initialize_state(n)
    agent.choose_action()
    reached = np.array([False, True, False, False, False, False])
    ok = agent.update_state(action=None, observation=np.zeros(n), source_reached=reached)
    assert ok.all()
    agent.kill(np.array([False, True, False, True, False, False]))           # extra kill of agent 3
    assert len(agent.belief) == 4
    np.testing.assert_allclose(agent.belief.belief_array.sum(1), 1.0, rtol=1e-5)
    # 'something' can never be observed in the margins: point masses in a margin corner make the update impossible
    m = agent.model
    corner = np.zeros((4, m.state_count))
    corner[:, 0] = 1.0
    corner[3] = m.start_probabilities
    agent.initialize_state(4, onb.BeliefSet(m, corner))
    agent.choose_action()
    okd = agent.update_state_device(torch.ones(4, dtype=torch.int32, device='cuda'), torch.zeros(4, dtype=torch.bool, device='cuda'))
    assert okd.cpu().tolist() == [False, False, False, True]
    agent.kill_device(~okd)
    assert len(agent.belief) == 1
    agent.choose_action()
    okd = agent.update_state_device(torch.zeros(1, dtype=torch.int32, device='cuda'), torch.ones(1, dtype=torch.bool, device='cuda'))
    agent.kill_device(torch.ones(1, dtype=torch.bool, device='cuda'))
    assert agent.belief is None
    with pytest.raises(RuntimeError):
        agent.initialize_state(2)
        agent.choose_action()
        agent.update_state(None, np.zeros(2), np.array([True, False]))
        agent.kill(np.array([False, False]))                                  # reached agent not killed


def test_fsvi_training_on_device_and_save_load(tmp_path):
    import olfactory_navigation_b200 as onb
    env, _ = make('T1')
    c = onb.synthetic.config('T1')
    agent = onb.agents.FSVI_Agent(env, thresholds=c['thresholds'], rng=7)
    hist = agent.train(expansions=6, max_belief_growth=10, print_progress=False, print_stats=False)
    assert agent.trained and len(agent.value_function) >= agent.model.action_count
    assert len(hist.backup_times) == 6 and 'Point Based Value Iteration' in hist.summary
    assert agent.mdp_policy is not None and len(agent.mdp_policy) == 4
    h = onb.run_test(agent, n=64, horizon=80, print_progress=False, print_stats=False)
    assert h.success_count > 0
    agent.save(folder=str(tmp_path))
    again = onb.agents.PBVI_Agent.load(agent.saved_at, environment=env)
    assert isinstance(again, onb.agents.FSVI_Agent)
    np.testing.assert_array_equal(again.value_function.alpha_vector_array, agent.value_function.alpha_vector_array)
    assert np.array_equal(again.value_function.actions, agent.value_function.actions)


def test_fsvi_matches_oracle_solver_alpha_count():
    'Same RNG -> same (a, o) trajectories -> comparable value functions: every oracle alpha has a device twin.'
    import olfactory_navigation_b200 as onb
    from conftest import oracle_model
    from oracle import refloader
    refloader.install_oracle_toolkit()
    import pomdp_toolkit as tk
    env, _ = make('T1')
    c = onb.synthetic.config('T1')
    agent = onb.agents.FSVI_Agent(env, thresholds=c['thresholds'])
    vf_d, _, _ = onb.solvers.FSVI.solve(agent.model, expansions=4, max_belief_growth=8, rng=np.random.default_rng(11), return_mdp_policy=True, print_progress=False)
    vf_o, _, _ = tk.solvers.FSVI.solve(oracle_model(agent.model), expansions=4, max_belief_growth=8, rng=np.random.default_rng(11), return_mdp_policy=True, print_progress=False)
    A_d, A_o = vf_d.alpha_vector_array, vf_o.alpha_vector_array
    matched = sum(np.any(np.all(np.isclose(A_d, a[None, :], rtol=1e-4, atol=1e-6), axis=1)) for a in A_o)
    assert matched >= 0.8 * len(A_o)


def test_robustness_flow_modified_environment():
    '''
    SURVEY §8 f3 / test_setups.py:421-429,646-654: an agent trained on one environment is run in a rescaled one (the device
    env step takes the TEST environment, the belief kernels the agent's model), and an agent moved to the new environment with
    modify_environment runs on its own grid.  The fused device loop must reproduce the host-driven protocol loop exactly.
    '''
    import olfactory_navigation_b200 as onb
    from olfactory_navigation_b200 import simulation
    c = onb.synthetic.config('T1')
    env = onb.Environment(**c['env_kwargs'])
    agent = onb.agents.FSVI_Agent(env, thresholds=c['thresholds'], rng=3)
    agent.train(expansions=5, print_progress=False, print_stats=False)
    n, horizon = 40, 50
    for big, ag in ((env.modify_scale(1.5), agent), (env.modify_scale(1.5), None)):
        if ag is None:
            ag = agent.modify_environment(big)
            assert ag.model.grid_shape == big.shape and len(ag.value_function) == len(agent.value_function)
        starts = big.random_start_points(n, np.random.default_rng(4))
        tshift = np.random.default_rng(5).integers(0, big.timesteps, n)
        if ag is agent:
            # the agent's grid is smaller than the test environment: keep the start points inside the agent's grid
            starts = np.minimum(starts, np.array(env.shape)[None, :] - 1)
        dev = onb.run_test(ag, n=n, start_points=starts, time_shift=tshift, environment=big, horizon=horizon, print_progress=False,
                           print_stats=False, print_warning=False, batches=1)
        host = simulation._run_generic(ag, n, starts, tshift.astype(int), environment=big, time_loop=True, horizon=horizon,
                                       initialization_values={}, reward_discount=0.99, distance_metric='l1', print_progress=False)
        assert np.array_equal(dev.done_at_step, host.done_at_step) and np.array_equal(dev.reached_source, host.reached_source)
        assert np.array_equal(np.array(dev.positions), np.array(host.positions))
        assert np.array_equal(np.array(dev.actions), np.array(host.actions))


def test_one_pass_step_run_matches_reference_run(monkeypatch):
    '''
    run_test with the one-pass update + policy kernel (default wherever olfnav_can_fuse_policy holds; here the T1 golden geometry
    with rows padded to 64 floats) against the reference's golden run, and against the two-kernel loop (OLFNAV_FUSED_STEP=0).
    The kernel keeps the belief rows grouped by action during the run; afterwards the agent's state is back in agent order.
    '''
    import olfactory_navigation_b200 as onb
    from olfactory_navigation_b200._lib import lib
    monkeypatch.setenv('OLFNAV_ROW_PAD', '64')
    g = golden('runtest_fsvi_T1')
    c = onb.synthetic.config('T1')
    env = onb.Environment(**c['env_kwargs'])
    agent = onb.agents.PBVI_Agent(env, thresholds=c['thresholds'])
    agent.value_function = onb.ValueFunction(agent.model, g['alphas'].astype(np.float64), g['alpha_actions'])
    assert lib.olfnav_can_fuse_policy(agent.model.handle(0), agent.value_function.handle(0)) == 1
    kw = dict(n=len(g['start_points']), start_points=g['start_points'].astype(int), time_shift=g['time_shift'], print_progress=False,
              print_stats=False, batches=1)
    got = onb.run_test(agent, horizon=int(g['horizon']), **kw)
    _compare_runs(got, g, 'T1')
    monkeypatch.setenv('OLFNAV_FUSED_STEP', '0')
    ref = onb.run_test(agent, horizon=int(g['horizon']), **kw)
    _compare_runs(ref, g, 'T1')
    # a short run that leaves agents alive: the beliefs the agent is left with are the two-kernel loop's, in agent order
    ref5 = onb.run_test(agent, horizon=5, **kw)
    b_ref, n_ref = agent.belief.belief_array, len(agent.belief)
    monkeypatch.setenv('OLFNAV_FUSED_STEP', '1')
    got5 = onb.run_test(agent, horizon=5, **kw)
    assert np.array_equal(np.array(got5.positions), np.array(ref5.positions)) and len(agent.belief) == n_ref > 0
    np.testing.assert_allclose(agent.belief.belief_array, b_ref, rtol=1e-6, atol=1e-30)


@pytest.mark.parametrize('fused', ['1', '0'])
def test_run_test_c2_shape_replayed_by_the_oracle(fused, monkeypatch):
    '''
    BASELINE shape end to end: 300 agents of a device-trained FSVI agent (the bench's training call) for 25 steps on the 83 x 371
    plume, every decision / move / observation / termination re-derived by the oracle (tests/replay_check.py), with the one-pass
    step kernel and with the two-kernel loop.
    '''
    import olfactory_navigation_b200 as onb
    import replay_check
    monkeypatch.setenv('OLFNAV_FUSED_STEP', fused)
    c = onb.synthetic.config('C2')
    env = onb.Environment(**c['env_kwargs'])
    agent = onb.agents.FSVI_Agent(env, thresholds=c['thresholds'], rng=onb.synthetic.SEED)
    agent.train(expansions=10, print_progress=False, print_stats=False)
    n = 300
    rng = np.random.default_rng(8)
    starts = env.random_start_points(n, rng)
    tshift = rng.integers(0, env.timesteps, n)
    hist = onb.run_test(agent, n=n, start_points=starts, time_shift=tshift, horizon=25, print_progress=False, print_stats=False, batches=1)
    vf = agent.value_function
    out = replay_check.replay(hist.actions, hist.positions, hist.observations, hist.done_at_step, hist.reached_source, 'C2',
                              vf.alpha_vector_array, vf.actions, starts, tshift)
    assert out['decisions'] > 20 * n * 0.8 and out['exact'] >= 0.98 * out['decisions']


@pytest.mark.parametrize('cfg', ['T1', 'C1'])
def test_in_place_run_equals_two_buffer_run(cfg, monkeypatch):
    '''
    OLFNAV_INPLACE=1: ONE belief buffer, the survivors are updated in the rows they occupy (the mode run_test picks by itself when two
    buffers do not fit the device: 10^6 agents x 30 793 states).  Same simulation as the two-buffer loop, step for step (the belief
    kernel writes bit-identical rows in place), and the reference's golden run through the oracle replay.
    '''
    import olfactory_navigation_b200 as onb
    g = golden(f'runtest_fsvi_{cfg}')
    env, _ = make(cfg)
    c = onb.synthetic.config(cfg)
    agent = onb.agents.PBVI_Agent(env, thresholds=c['thresholds'])
    agent.value_function = onb.ValueFunction(agent.model, g['alphas'].astype(np.float64), g['alpha_actions'])
    kw = dict(n=len(g['start_points']), start_points=g['start_points'].astype(int), time_shift=g['time_shift'], horizon=int(g['horizon']),
              print_progress=False, print_stats=False, batches=1)
    monkeypatch.setenv('OLFNAV_INPLACE', '0')
    ref = onb.run_test(agent, **kw)
    monkeypatch.setenv('OLFNAV_INPLACE', '1')
    got = onb.run_test(agent, **kw)
    assert np.array_equal(np.array(got.actions), np.array(ref.actions)) and np.array_equal(np.array(got.positions), np.array(ref.positions))
    assert np.array_equal(got.done_at_step, ref.done_at_step) and np.array_equal(got.reached_source, ref.reached_source)
    _compare_runs(got, g, cfg)
    # the Infotaxis agent through the same loop
    it = onb.agents.Infotaxis_Agent(env, thresholds=c['thresholds'])
    kw['horizon'] = 12
    monkeypatch.setenv('OLFNAV_INPLACE', '0')
    ref = onb.run_test(it, **kw)
    monkeypatch.setenv('OLFNAV_INPLACE', '1')
    got = onb.run_test(it, **kw)
    assert np.array_equal(np.array(got.actions), np.array(ref.actions)) and np.array_equal(got.done_at_step, ref.done_at_step)


@pytest.mark.parametrize('fused', ['1', '0'])
def test_run_ahead_discards_the_step_enqueued_behind_failed_updates(fused, monkeypatch):
    '''
    run_test enqueues step i + 1 as soon as step i's survivor count is known, i.e. before step i's belief update has said whether
    any update was impossible.  Agents that start from a point mass in a margin corner cannot observe anything but "nothing": their
    update fails at the first detection-level observation... here at step 0 for the ones seeded with an impossible prior, so the
    step enqueued ahead must be thrown away, the survivors compacted and the step launched again.  Same history as the loop that
    never runs ahead (OLFNAV_RUN_AHEAD=0) and as the host-driven protocol loop (reference step order, simulation.py:1452-1497), with
    the one-pass kernel and with the two kernels.
    '''
    import olfactory_navigation_b200 as onb
    monkeypatch.setenv('OLFNAV_ROW_PAD', '64')
    monkeypatch.setenv('OLFNAV_FUSED_STEP', fused)
    g = golden('runtest_fsvi_T1')
    c = onb.synthetic.config('T1')
    env = onb.Environment(**c['env_kwargs'])
    agent = onb.agents.PBVI_Agent(env, thresholds=c['thresholds'])
    agent.value_function = onb.ValueFunction(agent.model, g['alphas'].astype(np.float64), g['alpha_actions'])
    m = agent.model
    starts, tshift = g['start_points'].astype(int), g['time_shift'].astype(int)
    n, horizon = len(starts), 20
    # a third of the agents believe they sit in a cell whose successor cells can only ever emit observation 0: any other observation
    # (they do get some: their true position is a real start point) makes the normaliser 0
    prior = np.tile(m.start_probabilities, (n, 1))
    obs_tab = np.asarray(m.observation_table)                       # (S, A, O)
    quiet = (obs_tab[:, :, 0] == 1.0).all(axis=1)
    R = np.asarray(m.reachable_states)[:, :, 0]                      # (S, A): the one successor of every move
    silent = np.flatnonzero(quiet & quiet[R].all(axis=1))           # cells that stay silent for at least two moves
    assert len(silent) > 0
    for k in range(0, n, 3):
        prior[k] = 0.0
        prior[k, silent[k % len(silent)]] = 1.0
    kw = dict(n=n, start_points=starts, time_shift=tshift, horizon=horizon, print_progress=False, print_stats=False, batches=1)
    monkeypatch.setenv('OLFNAV_RUN_AHEAD', '1')
    got = onb.run_test(agent, initialization_values=dict(belief=onb.BeliefSet(m, prior)), **kw)
    monkeypatch.setenv('OLFNAV_RUN_AHEAD', '0')
    ref = onb.run_test(agent, initialization_values=dict(belief=onb.BeliefSet(m, prior)), **kw)
    for a, b in ((got.actions, ref.actions), (got.positions, ref.positions), (got.observations, ref.observations)):
        assert len(a) == len(b) and all(np.array_equal(np.asarray(x), np.asarray(y), equal_nan=True) for x, y in zip(a, b))
    assert np.array_equal(got.done_at_step, ref.done_at_step) and np.array_equal(got.reached_source, ref.reached_source)
    # the protocol loop
    pos, ts = starts.copy(), tshift.copy()
    agent.initialize_state(n, onb.BeliefSet(m, prior))
    hist = onb.SimulationHistory(pos, env, agent, ts, horizon)
    failed = 0
    for i in range(horizon):
        action = agent.choose_action()
        pos = env.move(pos, action)
        observation = env.get_observation(pos, time=ts + i)
        reached = env.source_reached(pos)
        ok = agent.update_state(action=action, observation=observation, source_reached=reached)
        failed += int((~ok & ~reached).sum())
        term = reached | ~ok
        hist.add_step(actions=action, next_positions=pos, observations=observation, reached_source=reached, interupt=term)
        pos, ts = pos[~term], ts[~term]
        agent.kill(simulations_to_kill=term)
        if len(pos) == 0:
            break
    assert failed > 0, 'the scenario must contain impossible updates'
    assert np.array_equal(np.array(got.done_at_step), np.array(hist.done_at_step))
    assert np.array_equal(np.array(got.reached_source), np.array(hist.reached_source))
    for x, y in zip(got.positions, hist.positions):
        assert np.array_equal(np.asarray(x), np.asarray(y))


def test_many_agents_compaction_across_plan_tiles():
    '''
    9 600 agents = 200 copies of the 48 golden start points (several 4 096-agent tiles of the compaction plan).  With every agent
    on its own clock (reference_time_quirk off) the copies are independent, so each must reproduce the 48-agent run exactly, in order.
    '''
    import olfactory_navigation_b200 as onb
    g = golden('runtest_fsvi_T1')
    c = onb.synthetic.config('T1')
    env = onb.Environment(**c['env_kwargs'])
    env.reference_time_quirk = False
    agent = onb.agents.PBVI_Agent(env, thresholds=c['thresholds'])
    agent.value_function = onb.ValueFunction(agent.model, g['alphas'].astype(np.float64), g['alpha_actions'])
    starts, tshift, horizon = g['start_points'].astype(int), g['time_shift'], int(g['horizon'])
    kw = dict(horizon=horizon, print_progress=False, print_stats=False, batches=1)
    one = onb.run_test(agent, n=len(starts), start_points=starts, time_shift=tshift, **kw)
    reps = 200
    many = onb.run_test(agent, n=reps * len(starts), start_points=np.tile(starts, (reps, 1)), time_shift=np.tile(tshift, reps), **kw)
    assert np.array_equal(np.array(many.done_at_step), np.tile(np.array(one.done_at_step), reps))
    assert np.array_equal(np.array(many.reached_source), np.tile(np.array(one.reached_source), reps))
    pos_one, pos_many = np.array(one.positions), np.array(many.positions)
    assert np.array_equal(pos_many, np.tile(pos_one, (1, reps, 1)))


def test_generate_beliefs_from_trajectory():
    'Replay of one recorded trajectory (reference pbvi_agent.py:812-873) == the oracle toolkit updating a single Belief step by step.'
    import olfactory_navigation_b200 as onb
    from conftest import oracle_model
    from oracle import refloader
    refloader.install_oracle_toolkit()
    import pomdp_toolkit as tk
    g = golden('runtest_fsvi_T1')
    env, _ = make('T1')
    c = onb.synthetic.config('T1')
    agent = onb.agents.PBVI_Agent(env, thresholds=c['thresholds'])
    agent.value_function = onb.ValueFunction(agent.model, g['alphas'].astype(np.float64), g['alpha_actions'])
    hist = onb.run_test(agent, n=len(g['start_points']), start_points=g['start_points'].astype(int), time_shift=g['time_shift'],
                        horizon=int(g['horizon']), print_progress=False, print_stats=False, batches=1)
    k = int(np.argmax(np.where(np.array(hist.done_at_step) < 0, len(hist), np.array(hist.done_at_step))))     # the longest trajectory
    seq = agent.generate_beliefs_from_trajectory(hist, k)
    steps = len(hist) if hist.done_at_step[k] < 0 else int(hist.done_at_step[k])
    om = oracle_model(agent.model)
    b = tk.Belief(om)
    rows = [b.values]
    for t in range(steps):
        a = int(np.argwhere(np.all(agent.action_set == np.asarray(hist.actions[t][k])[None, :], axis=1))[0, 0])
        o = int(agent.discretize_observations(np.array([hist.observations[t][k]]), np.asarray(hist.actions[t][k])[None, :], np.array([False]))[0])
        try:
            b = b.update(a=a, o=o)
            rows.append(b.values)
        except Exception:
            pass
    assert len(seq) == len(rows) and len(seq) >= 2
    np.testing.assert_allclose(seq.belief_array, np.array(rows), rtol=2e-5, atol=1e-12)
