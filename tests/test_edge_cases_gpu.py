'''
Edge cases of the C-ABI hot path: empty and single-row inputs, row counts around the 128-row tensor-core tile, grids whose width is a
multiple of 4 (no padding column: the x-shift must not leak across grid rows), argument errors reported through status codes, and
simulations in which every agent finishes at once.  Parity against the oracle toolkit / the separate kernels.
'''
import numpy as np
import pytest
import torch

from conftest import make, oracle_model

pytestmark = pytest.mark.gpu


def _env(boundary: str, w: int = 10, margins=(3, 3)):
    'A grid whose total width is a multiple of 4: (w + 2 * margin) = 16 columns, 13 rows.'
    import olfactory_navigation_b200 as onb
    data = onb.synthetic.plume_frames(16, 7, w, source_y=3, seed=3)
    return onb.Environment(data_file=data, data_source_position=[3, 0], source_radius=1.0, margins=list(margins),
                           boundary_condition=boundary, start_zone='data_zone')


@pytest.mark.parametrize('boundary', ['stop', 'wrap', 'wrap_vertical', 'wrap_horizontal'])
def test_no_padding_column_grid(boundary):
    'W % 4 == 0: belief step, fused update + evaluate and infotaxis against the oracle.'
    import olfactory_navigation_b200 as onb
    from oracle import refloader
    refloader.install_oracle_toolkit()
    import pomdp_toolkit as tk
    env = _env(boundary)
    agent = onb.agents.QMDP_Agent(env, thresholds=0.5)
    m, om = agent.model, oracle_model(agent.model)
    assert m.grid_shape[1] % 4 == 0 and m.padded_states == m.state_count
    rng = np.random.default_rng(2)
    n = 150
    d = -np.log(1.0 - rng.random((n, m.state_count)))
    B = (d / d.sum(1, keepdims=True)).astype(np.float32).astype(np.float64)
    B[:8] = 0
    H, W = m.grid_shape
    corners = [0, W - 1, (H - 1) * W, H * W - 1, W, 2 * W - 1, (H // 2) * W, (H // 2) * W + W - 1]
    B[np.arange(8), corners] = 1.0                                    # point masses on the edges: the boundary rules decide
    act = np.arange(n) % 4
    obs = np.zeros(n, dtype=int)
    new, prov = onb.BeliefSet(m, B).update(act, obs, raise_on_impossible_belief=False, return_provenance=True)
    ref, rprov = tk.BeliefSet(om, B).update(act, obs, raise_on_impossible_belief=False, return_provenance=True)
    assert np.array_equal(prov[:, 0], rprov[:, 0])
    np.testing.assert_allclose(new.belief_array, ref.belief_array, rtol=1e-5, atol=1e-30)

    alphas = np.abs(rng.standard_normal((9, m.state_count))).astype(np.float32).astype(np.float64)
    vf = onb.ValueFunction(m, alphas, np.arange(9) % 4)
    fnew, fok, fval, farg, fact = onb.BeliefSet(m, B).update_and_evaluate(vf, act, obs)
    okn = fok.cpu().numpy()
    assert np.array_equal(np.flatnonzero(okn), rprov[:, 0])
    np.testing.assert_allclose(fnew.belief_array[okn], ref.belief_array, rtol=1e-5, atol=1e-30)
    np.testing.assert_allclose(fval.cpu().numpy()[okn], (ref.belief_array @ alphas.T).max(1), rtol=1e-5)

    it = onb.agents.Infotaxis_Agent(env, thresholds=0.5)
    it.initialize_state(n, onb.BeliefSet(it.model, B))
    it.choose_action_device(return_delta_H=True)
    bs = tk.BeliefSet(om, B)
    u, p = bs.unnormalised_successors()
    with np.errstate(divide='ignore', invalid='ignore'):
        xlogx = lambda t: np.where(t > 0, t * np.log(t), 0.0)   # noqa: E731
        dH = bs.entropies[:, None] - np.sum(xlogx(p), 2) + np.einsum('nas,sa->na', u.sum(2), np.sum(xlogx(om.observation_table), 2)) + np.sum(xlogx(u.sum(2)), 2)
    np.testing.assert_allclose(it.last_delta_H.cpu().numpy(), dH, rtol=1e-3, atol=2e-4)


def test_empty_and_single_row_calls():
    import olfactory_navigation_b200 as onb
    from olfactory_navigation_b200._lib import lib, ptr, stream, check
    _, agent = make('T1')
    m = agent.model
    dev = torch.device('cuda', 0)
    h = m.handle(0)
    Sp = m.padded_states
    rng = np.random.default_rng(0)
    vf = onb.ValueFunction(m, np.abs(rng.standard_normal((5, m.state_count))), np.arange(5) % 4)
    b = torch.full((4, Sp), 7.0, dtype=torch.float32, device=dev)
    s = torch.ones(4, dtype=torch.float32, device=dev)
    i32 = torch.zeros(4, dtype=torch.int32, device=dev)
    u8 = torch.zeros(4, dtype=torch.uint8, device=dev)
    f32 = torch.zeros(4, dtype=torch.float32, device=dev)
    # n = 0: every entry point returns OK and writes nothing
    check(lib.olfnav_belief_init(h, ptr(b), Sp, 0, ptr(s), stream()))
    check(lib.olfnav_belief_step(h, ptr(b), ptr(b), Sp, 0, ptr(i32), ptr(i32), None, None, ptr(s), ptr(s), ptr(u8), stream()))
    check(lib.olfnav_evaluate(h, vf.handle(0), ptr(b), Sp, 0, ptr(s), ptr(f32), ptr(i32), ptr(i32), 0, stream()))
    check(lib.olfnav_belief_step_evaluate(h, vf.handle(0), ptr(b), 1, ptr(b.clone()), 1, Sp, 0, ptr(i32), ptr(i32), None, ptr(s), ptr(s), ptr(u8), None, None, None, None, stream()))
    check(lib.olfnav_infotaxis_choose(h, ptr(b), Sp, 0, ptr(s), ptr(i32), None, stream()))
    check(lib.olfnav_entropies(h, ptr(b), Sp, 0, ptr(s), ptr(f32), stream()))
    torch.cuda.synchronize()
    assert float(b.min()) == 7.0 and float(b.max()) == 7.0
    assert len(onb.BeliefSet(m, np.zeros((0, m.state_count)))) == 0
    # n = 1 through every path
    B1 = m.start_probabilities[None, :]
    new = onb.BeliefSet(m, B1).update([1], [0])
    fnew, fok, fval, farg, fact = onb.BeliefSet(m, B1).update_and_evaluate(vf, [1], [0])
    assert bool(fok[0]) and torch.equal(fnew._b[:1], new._b[:1])
    v1, a1 = vf.evaluate_at(new, path=1)
    v2, a2 = vf.evaluate_at(new, path=2)
    np.testing.assert_allclose(v1, v2, rtol=1e-5)
    np.testing.assert_allclose(fval.cpu().numpy(), v2, rtol=1e-5)


@pytest.mark.parametrize('n', [127, 128, 129, 257])
def test_rows_around_the_tile_size(n):
    import olfactory_navigation_b200 as onb
    _, agent = make('T0')
    m = agent.model
    rng = np.random.default_rng(n)
    d = -np.log(1.0 - rng.random((n, m.state_count)))
    B = (d / d.sum(1, keepdims=True)).astype(np.float32).astype(np.float64)
    alphas = np.abs(rng.standard_normal((33, m.state_count))).astype(np.float32).astype(np.float64)
    vf = onb.ValueFunction(m, alphas, np.arange(33) % 4)
    act, obs = rng.integers(0, 4, n), np.zeros(n, dtype=int)
    new = onb.BeliefSet(m, B).update(act, obs)
    fnew, fok, fval, farg, fact = onb.BeliefSet(m, B).update_and_evaluate(vf, act, obs)
    assert bool(fok.all()) and torch.equal(fnew._b[:n], new._b[:n])
    sc = new.belief_array @ alphas.T
    for vals in (fval.cpu().numpy(), vf.evaluate_at(new, path=2)[0]):
        np.testing.assert_allclose(vals, sc.max(1), rtol=1e-5)


def test_argument_errors_are_status_codes():
    import olfactory_navigation_b200 as onb
    from olfactory_navigation_b200._lib import lib, ptr, stream, check, OlfnavError
    _, agent = make('T0')
    m = agent.model
    dev = torch.device('cuda', 0)
    h, Sp = m.handle(0), m.padded_states
    b = torch.zeros((2, Sp), dtype=torch.float32, device=dev)
    s = torch.ones(2, dtype=torch.float32, device=dev)
    i32 = torch.zeros(2, dtype=torch.int32, device=dev)
    u8 = torch.zeros(2, dtype=torch.uint8, device=dev)
    with pytest.raises(OlfnavError):          # leading dimension smaller than the padded row
        check(lib.olfnav_belief_step(h, ptr(b), ptr(b), Sp - 4, 2, ptr(i32), ptr(i32), None, None, ptr(s), ptr(s), ptr(u8), stream()))
    with pytest.raises(OlfnavError):          # null action pointer
        check(lib.olfnav_belief_step(h, ptr(b), ptr(b), Sp, 2, None, ptr(i32), None, None, ptr(s), ptr(s), ptr(u8), stream()))
    with pytest.raises(OlfnavError):          # misaligned belief buffer
        check(lib.olfnav_belief_step(h, b.data_ptr() + 4, ptr(b), Sp, 1, ptr(i32), ptr(i32), None, None, ptr(s), ptr(s), ptr(u8), stream()))
    vf = onb.ValueFunction(m, np.ones((3, m.state_count)), np.arange(3))
    with pytest.raises(OlfnavError):          # the fused kernel cannot run in place
        check(lib.olfnav_belief_step_evaluate(h, vf.handle(0), ptr(b), 2, ptr(b), 2, Sp, 2, ptr(i32), ptr(i32), None, ptr(s), ptr(s), ptr(u8), None, None, None, None, stream()))
    torch.cuda.synchronize()                   # the context is still healthy
    assert bool((onb.BeliefSet(m, 2)._b >= 0).all())


def test_everybody_finishes_at_once_and_single_agent_run():
    'All agents start next to the source and reach it on the first step; a run with n = 1.'
    import olfactory_navigation_b200 as onb
    env, _ = make('T0')
    c = onb.synthetic.config('T0')
    agent = onb.agents.QMDP_Agent(env, thresholds=c['thresholds'])
    agent.train(print_progress=False, print_stats=False)
    src = env.source_position
    start = np.tile(np.array([[src[0], src[1] + 2]]), (37, 1))       # two cells east of the source, radius 1: one step west reaches it
    hist = onb.run_test(agent, n=37, start_points=start, time_shift=np.zeros(37, dtype=int), horizon=20, print_progress=False, print_stats=False, batches=1)
    assert bool(np.all(hist.reached_source)) and np.all(np.array(hist.done_at_step) == np.array(hist.done_at_step)[0])
    one = onb.run_test(agent, n=1, start_points=start[:1], time_shift=np.zeros(1, dtype=int), horizon=20, print_progress=False, print_stats=False, batches=1)
    assert bool(one.reached_source[0]) and one.done_at_step[0] == hist.done_at_step[0]


@pytest.mark.parametrize('boundary', ['stop', 'wrap_vertical', 'wrap'])
@pytest.mark.parametrize('rot_inplace', ['0', '1'])
def test_in_place_rotated_traversal_tall_grid(boundary, rot_inplace, monkeypatch):
    '''
    A grid tall enough for several ring chunks per row (200 x 52): in place, CTAs start at different chunks and the one source row that
    is overwritten early comes from the stash — rows must equal the out-of-place update bit for bit, and the oracle within 1e-5.
    '''
    import olfactory_navigation_b200 as onb
    from olfactory_navigation_b200._lib import lib, ptr, stream, check
    from oracle import refloader
    refloader.install_oracle_toolkit()
    import pomdp_toolkit as tk
    monkeypatch.setenv('OLFNAV_BELIEF_ROT_INPLACE', rot_inplace)      # read per call
    data = onb.synthetic.plume_frames(8, 180, 32, source_y=90, seed=5)
    env = onb.Environment(data_file=data, data_source_position=[90, 0], source_radius=2.0, margins=[10, 10], boundary_condition=boundary,
                          start_zone='data_zone')
    agent = onb.agents.QMDP_Agent(env, thresholds=0.5)
    m, om = agent.model, oracle_model(agent.model)
    assert m.grid_shape == (200, 52)
    dev = torch.device('cuda', 0)
    rng = np.random.default_rng(9)
    n = 600                                                          # more CTAs than chunk offsets: every rotation occurs
    d = -np.log(1.0 - rng.random((n, m.state_count)))
    B = (d / d.sum(1, keepdims=True)).astype(np.float32).astype(np.float64)
    for acts in (np.zeros(n, dtype=int), np.full(n, 2), rng.integers(0, 4, n)):      # all north, all south, mixed
        obs = np.zeros(n, dtype=int)
        bs = onb.BeliefSet(m, B)
        a = torch.as_tensor(acts, dtype=torch.int32, device=dev)
        o = torch.as_tensor(obs, dtype=torch.int32, device=dev)
        ok = torch.empty(n, dtype=torch.uint8, device=dev)
        out_b, out_s = torch.empty_like(bs._b), torch.empty_like(bs._scale)
        check(lib.olfnav_belief_step(m.handle(0), ptr(bs._b), ptr(out_b), bs.ld, n, ptr(a), ptr(o), None, None, ptr(bs._scale), ptr(out_s), ptr(ok), stream()))
        ip_b, ip_s = bs._b.clone(), bs._scale.clone()
        check(lib.olfnav_belief_step(m.handle(0), ptr(ip_b), ptr(ip_b), bs.ld, n, ptr(a), ptr(o), None, None, ptr(ip_s), ptr(ip_s), ptr(ok), stream()))
        assert torch.equal(ip_b, out_b)
        assert float(((ip_s - out_s).abs() / out_s).max()) < 1e-6
        ref = tk.BeliefSet(om, B[:40]).update(acts[:40], obs[:40], raise_on_impossible_belief=False)
        got = onb.BeliefSet(m, _buffers=(ip_b, ip_s, n), device=0).belief_array[:40]
        np.testing.assert_allclose(got, ref.belief_array, rtol=1e-5, atol=1e-30)


def test_generate_all_successors_matches_the_oracle():
    '''
    BeliefSet.generate_all_successors (toolkit surface of the reference's own Infotaxis_Agent, agents/infotaxis_agent.py:260-262):
    same rows, same order, same provenance as the oracle; refused (MemoryError) when the successors would not fit the budget.
    '''
    import olfactory_navigation_b200 as onb
    from oracle import refloader
    refloader.install_oracle_toolkit()
    import pomdp_toolkit as tk
    _, agent = make('T1')
    m, om = agent.model, oracle_model(agent.model)
    rng = np.random.default_rng(4)
    d = -np.log(1.0 - rng.random((7, m.state_count)))
    B = (d / d.sum(1, keepdims=True)).astype(np.float32).astype(np.float64)
    B[0] = 0
    B[0, 3] = 1.0                                          # a point mass: most successors are impossible
    got, prov = onb.BeliefSet(m, B).generate_all_successors(return_provenance=True)
    ref, rprov = tk.BeliefSet(om, B).generate_all_successors(raise_on_impossible_belief=False, return_provenance=True)
    assert np.array_equal(prov, rprov) and len(got) == len(ref) < 7 * m.action_count * m.observation_count
    np.testing.assert_allclose(got.belief_array, ref.belief_array, rtol=1e-5, atol=1e-30)
    with pytest.raises(MemoryError):
        onb.BeliefSet(m, B).generate_all_successors(max_bytes=1024)
    with pytest.raises(ValueError):
        onb.BeliefSet(m, B).generate_all_successors(raise_on_impossible_belief=True)
