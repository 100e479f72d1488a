'''
Host-side mirror (Environment, Agent, exact_converter, discretisation) against the golden vectors
frozen from the reference's in-repo code, and — when /root/reference is present (build container
only) — against the live reference objects.  CPU only.
'''
import numpy as np
import pytest

from conftest import ALL_CFGS, golden, make


@pytest.mark.parametrize('cfg', ALL_CFGS)
def test_converter_tables(cfg):
    g = golden(f'converter_{cfg}')
    env, agent = make(cfg)
    m = agent.model
    assert tuple(g['shape']) == env.shape
    assert np.array_equal(g['thresholds'], agent.thresholds)
    assert np.array_equal(g['action_set'], agent.action_set)
    assert np.array_equal(g['reachable_states'], m.reachable_states[:, :, 0])
    assert np.array_equal(g['observation_table'], m.observation_table)          # bit exact
    assert np.array_equal(g['end_states'], m.end_states)
    assert np.array_equal(g['start_probabilities'], m.start_probabilities)
    assert np.array_equal(g['source_position'], env.source_position)
    assert np.array_equal(g['margins'], env.margins)
    assert np.array_equal(g['expected_rewards'], m.expected_rewards_table)
    assert np.all(m.observation_table.sum(axis=2) == 1.0)


@pytest.mark.parametrize('cfg', ALL_CFGS)
def test_env_step_numpy(cfg):
    g = golden(f'envstep_{cfg}')
    env, agent = make(cfg)
    new_pos = env.move(g['pos'].astype(int), agent.action_set[g['act']])
    assert np.array_equal(new_pos, g['new_pos'])
    odor = env.get_observation(new_pos, time=g['time_shift'] + int(g['step']))
    assert np.array_equal(odor, g['odor'])
    reached = env.source_reached(new_pos)
    assert np.array_equal(reached, g['reached'])
    ids = agent.discretize_observations(odor, agent.action_set[g['act']], reached)
    assert np.array_equal(ids, g['obs_id'])


def test_discretize_documented_example():
    'SURVEY §3.1: thresholds=0.5, odor [0, 1, .5, .2], reached [F,F,F,T] -> [0,1,1,2].'
    import olfactory_navigation_b200 as onb
    env, _ = make('T0')
    ag = onb.Agent(env, thresholds=0.5)
    ids = ag.discretize_observations(np.array([0., 1., 0.5, 0.2]), None, np.array([False, False, False, True]))
    assert ids.tolist() == [0, 1, 1, 2]


def test_agent_constructor_variants():
    import olfactory_navigation_b200 as onb
    env, _ = make('T0')
    assert onb.Agent(env, thresholds=[0.7, 0.2]).thresholds.tolist() == [-np.inf, 0.2, 0.7, np.inf]
    ag = onb.Agent(env, actions={'N': [-1, 0], 'stay': [0, 0]})
    assert ag.action_set.tolist() == [[-1, 0], [0, 0]] and ag.action_labels == ['N', 'stay']
    ag = onb.agents.PBVI_Agent(env, thresholds=0.5, spatial_subdivisions=[2, 2], rng=3)    # stray kwarg swallowed like the reference
    assert ag.model.state_count == env.shape[0] * env.shape[1]
    # space-aware observations (reference agent.py:206-234,418-434): odor bin x position cell, goal id = bins x cells
    sa = onb.Agent(env, thresholds=0.5, space_aware=True, spacial_subdivisions=[2, 3])
    H, W = env.shape
    obs = np.array([[0.9, 0, 0], [0.1, H - 1, W - 1], [0.9, H - 1, 0], [0.3, 5, 5]], dtype=float)
    ids = sa.discretize_observations(obs, None, np.array([False, False, False, True]))
    cells = sa._environment_to_subdivision_mapping
    assert cells.min() == 0 and cells.max() == 5
    assert ids.tolist() == [1 * 6 + cells[0, 0], 0 * 6 + cells[H - 1, W - 1], 1 * 6 + cells[H - 1, 0], 2 * 6]
    with pytest.raises(AssertionError):             # the converters reject space-aware agents like the reference's (environment_converter.py:37)
        onb.agents.PBVI_Agent(env, thresholds=0.5, space_aware=True)


def test_random_start_points_rng_stream():
    env, _ = make('T1')
    a = env.random_start_points(50, np.random.default_rng(5))
    rng = np.random.default_rng(5)
    states = rng.choice(np.arange(env.shape[0] * env.shape[1]), size=50, replace=True, p=env.start_probabilities.ravel())
    assert np.array_equal(a, np.array(np.unravel_index(states, env.shape)).T)
    assert np.all(env.start_probabilities[a[:, 0], a[:, 1]] > 0)


def test_grid_inference_from_move_table():
    from olfactory_navigation_b200.pomdp import POMDP
    for cfg in ['T0', 'T1', 'T2', 'T3']:
        _, agent = make(cfg)
        m = agent.model
        H, W = m.grid_shape
        m2 = POMDP(states=[[0] * W for _ in range(H)], actions=4, observations=m.observation_count,
                   reachable_states_table=m.reachable_states, observation_table=m.observation_table,
                   end_states=m.end_states, start_probabilities=m.start_probabilities)
        assert m2.boundary == m.boundary and np.array_equal(m2.moves, m.moves)


def test_simulation_history_bookkeeping():
    import olfactory_navigation_b200 as onb
    env, agent = make('T0')
    starts = np.array([[3, 5], [4, 6], [5, 7]])
    h = onb.SimulationHistory(starts, env, agent, np.zeros(3, int), horizon=5)
    h.add_step(np.array([0, 1, 2]), starts + 1, np.array([.1, .2, .3]), np.array([False, True, False]), np.array([False, False, False]), action_is_id=True)
    h.add_step(np.array([3, 3]), starts[[0, 2]] + 2, np.array([.4, .5]), np.array([False, False]), np.array([True, False]), action_is_id=True)
    assert h.done_at_step.tolist() == [2, 1, -1] and h.reached_source.tolist() == [False, True, False]
    assert h.actions[1].tolist() == [[0, -1], [-1, -1], [0, -1]]
    assert h.positions[1][1].tolist() == [-1, -1] and h.observations[1][1] == -1
    both = h + h
    assert both.n == 6 and both.done_at_step.tolist() == [2, 1, -1, 2, 1, -1] and len(both.positions[0]) == 6
    assert 'Simulations reached goal: 1/3' in h.summary


def test_against_live_reference_when_available():
    from oracle import refloader
    if not refloader.available():
        pytest.skip('/root/reference only exists in the build container')
    refloader.load()
    from olfactory_navigation import Environment as RefEnv
    from olfactory_navigation.agents import QMDP_Agent as RefQMDP
    import olfactory_navigation_b200 as onb
    for cfg in ['T1', 'C1']:
        c = onb.synthetic.config(cfg)
        renv = RefEnv(**c['env_kwargs'])
        env, agent = make(cfg)
        assert renv.shape == env.shape and renv.name == env.name
        assert np.array_equal(renv.start_probabilities, env.start_probabilities)
        assert np.array_equal(renv.data_bounds, env.data_bounds)
        ragent = RefQMDP(renv, thresholds=c['thresholds'])
        assert ragent.name == agent.name
        assert np.array_equal(ragent.model.observation_table, agent.model.observation_table)
        assert np.array_equal(ragent.model.reachable_states, agent.model.reachable_states)


def test_base_agent_persistence(tmp_path):
    'Base Agent.save/load/modify_environment/size_on_memory (reference agent.py:488-667): pickle by default, subclass dispatch by folder name.'
    import olfactory_navigation_b200 as onb
    env, _ = make('T0')
    ag = onb.agents.Infotaxis_Agent(env, thresholds=[0.2, 0.7])
    ag.save(folder=str(tmp_path))
    folder = str(tmp_path / f'Agent-{ag.name}')
    with pytest.raises(Exception):
        ag.save(folder=str(tmp_path))                       # not empty, no force
    ag.save(folder=str(tmp_path), force=True)
    back = onb.Agent.load(folder)
    assert type(back) is onb.agents.Infotaxis_Agent and back.name == ag.name
    assert np.array_equal(back.thresholds, ag.thresholds)
    assert np.array_equal(back.model.observation_table, ag.model.observation_table)
    env2, _ = make('T1')
    moved = ag.modify_environment(env2)
    assert type(moved) is type(ag) and moved.environment is env2 and np.array_equal(moved.thresholds, ag.thresholds)
    assert ag.size_on_memory() > ag.model.observation_table.nbytes


def test_generic_loop_records_the_odor_cue_of_space_aware_agents():
    '''
    Host-driven protocol loop with a SPACE-AWARE custom agent: the agent receives the (n, 1 + dims) observation, the history records
    the odor cue only (reference simulation.py:1489: `observation[:,0] if agent.space_aware else observation`).
    '''
    import olfactory_navigation_b200 as onb
    from olfactory_navigation_b200 import simulation
    env, _ = make('T1')

    class Wanderer(onb.Agent):
        def initialize_state(self, n=1):
            self.n_alive, self.seen = n, []

        def choose_action(self):
            return self.action_set[self.rng.integers(0, len(self.action_set), self.n_alive)]

        def update_state(self, action, observation, source_reached):
            assert observation.shape == (self.n_alive, 1 + self.environment.dimensions)
            self.seen.append(observation.copy())
            return None

        def kill(self, simulations_to_kill):
            self.n_alive -= int(np.sum(simulations_to_kill))

    agent = Wanderer(env, thresholds=0.5, space_aware=True, rng=2)
    n, horizon = 12, 25
    starts = env.random_start_points(n, np.random.default_rng(0))
    hist = simulation._run_generic(agent, n, starts, np.zeros(n, dtype=int), environment=env, time_loop=True, horizon=horizon,
                                   initialization_values={}, reward_discount=0.99, distance_metric='l1', print_progress=False)
    obs = hist.observations                    # used to raise: (n, 3) rows written into an (n,) record
    assert len(obs) == len(hist) and obs[0].shape == (n,)
    live = np.arange(n)
    for step, seen in enumerate(agent.seen):
        assert np.array_equal(obs[step][live], seen[:, 0])
        assert np.array_equal(hist.positions[step][live], seen[:, 1:].astype(int))
        live = live[hist.done_at_step[live] != step + 1]
    assert len(hist.simulation_dfs) == n


def test_host_busy_probe_is_cheap_and_bounded():
    '''
    run_test only takes the one-pass step kernel on quiet single-rank hosts: the probe behind that decision returns a
    fraction in [0, 1] and costs nothing after its first call in a process.
    '''
    import time
    from olfactory_navigation_b200 import simulation as sim
    v = sim._host_busy_fraction()
    assert 0.0 <= v <= 1.0
    t = time.perf_counter()
    for _ in range(20):
        w = sim._host_busy_fraction()
        assert 0.0 <= w <= 1.0
    assert time.perf_counter() - t < 0.02
