'''
BASELINE shape (C2/C3/C4: 83 x 371 states, wrap_vertical, Wp = 384 = 96 float4 per grid row) against the LIVE oracle — not against
invariants or against another device path.  The oracle model is oracle/sim.py's own construction of the configuration
(independent of the product's converter); 64 agents x 30 793 states cost it well under a second.

  * belief step  : all 4 actions x both observations, out of place, in place and through a row list (every CTA index gets its own
                   rotated traversal start and direction), dense / start / walked / point-mass rows; 1e-5 relative, ok flags exact
  * evaluate_at  : G = 75 (FSVI size of the bench) and G = 512, FP32 SIMT, tcgen05 3xTF32 and tcgen05 FP16-split kernels;
                   values 1e-5 relative, argmax exact where the float64 margin is clear
  * backup       : B = 256 belief points x G = 48 alpha vectors, both GEMM kernels; best (a, g) exact where clear, values and new
                   vectors 1e-5 relative
  * FSVI training: device solver == oracle solver under the strict contract when both follow the same MDP policy, and the one
                   difference between the two trainings at this shape (75 against 74 alpha vectors) is pinned to exact float64 ties
                   of the MDP policy.
Reference call sites: agents/pbvi_agent.py:741,783-787, agents/fsvi_agent.py:201-224.
'''
import numpy as np
import pytest
import torch

from conftest import make

pytestmark = pytest.mark.gpu

_SETUP = {}


def _oracle():
    'oracle/sim.py model of C2 (installs the oracle toolkit as `pomdp_toolkit` on first use).'
    if 'c2' not in _SETUP:
        import replay_check
        sim, setup = replay_check.oracle_setup('C2')
        _SETUP['c2'] = setup
    return _SETUP['c2']


def _beliefs(model, om, n, rng):
    'n float32-representable rows: the start belief, rows walked from it, dense random rows and point masses next to every border.'
    import pomdp_toolkit as tk          # installed by _oracle()
    S = om.state_count
    H, W = 83, 371
    rows = [om.start_probabilities.copy()]
    walk = tk.BeliefSet(om, np.tile(om.start_probabilities, (n // 2, 1)))
    for _ in range(4):
        walk = walk.update(rng.integers(0, 4, len(walk)), np.zeros(len(walk), dtype=int), raise_on_impossible_belief=False)
    rows += list(walk.belief_array)
    for (y, x) in [(0, 0), (0, W - 1), (H - 1, 0), (H - 1, W - 1), (0, 200), (H - 1, 200), (40, 0), (40, W - 1)]:
        e = np.zeros((H, W))
        e[y, x] = 1.0
        rows.append(e.ravel())
    d = -np.log(1.0 - rng.random((n - len(rows), S)))
    rows += list(d / d.sum(1, keepdims=True))
    return np.array(rows[:n]).astype(np.float32).astype(np.float64)


def _check_tables(m, om):
    'The product model the kernels run on carries the same tables as the independent oracle construction.'
    assert m.state_count == om.state_count and m.grid_shape == (83, 371)
    assert np.array_equal(m.reachable_states.reshape(om.state_count, -1), om.reachable_states.reshape(om.state_count, -1))
    assert np.array_equal(m.observation_table, om.observation_table)
    assert np.array_equal(m.start_probabilities, om.start_probabilities)


@pytest.mark.parametrize('mode', ['out_of_place', 'in_place', 'row_list'])
def test_belief_step_c2_vs_oracle(mode):
    from test_belief_gpu import _step, RTOL, ATOL
    om = _oracle()['model']
    import pomdp_toolkit as tk
    _, agent = make('C2')
    m = agent.model
    _check_tables(m, om)
    rng = np.random.default_rng(42)
    n = 64
    B = _beliefs(m, om, n, rng)
    act = np.arange(n) % 4                                # every action ...
    obs = (np.arange(n) // 4) % 2                         # ... under both observations ('nothing', 'something')
    u, p = tk.BeliefSet(om, B).unnormalised_successors()
    z = p[np.arange(n), act, obs]
    ok_ref = z > 0
    ref = np.zeros_like(B)
    ref[ok_ref] = u[np.arange(n), act, obs][ok_ref] / z[ok_ref, None]
    assert ok_ref.sum() >= 40 and (~ok_ref).sum() >= 1       # point masses in the margins cannot smell anything
    if mode == 'row_list':
        rows = np.sort(rng.choice(n, size=40, replace=False))
        out, ok = _step(m, B, act[rows], obs[rows], src_rows=rows)
        ok_ref, ref = ok_ref[rows], ref[rows]
    else:
        out, ok = _step(m, B, act, obs, inplace=(mode == 'in_place'))
    assert np.array_equal(ok, ok_ref)
    np.testing.assert_allclose(out[ok], ref[ok], rtol=RTOL, atol=ATOL)
    assert np.all(out[~ok] == 0)


@pytest.mark.parametrize('G', [75, 512])
def test_evaluate_c2_vs_oracle(G, monkeypatch):
    import olfactory_navigation_b200 as onb
    om = _oracle()['model']
    _, agent = make('C2')
    m = agent.model
    rng = np.random.default_rng(G)
    n = 64
    B = _beliefs(m, om, n, rng)
    # value-function-like vectors: non-negative, decaying away from a random centre, magnitudes over three decades
    yy, xx = np.meshgrid(np.arange(83), np.arange(371), indexing='ij')
    alphas = np.empty((G, om.state_count))
    for g in range(G):
        cy, cx = rng.integers(0, 83), rng.integers(0, 371)
        alphas[g] = (0.99 ** (np.abs(yy - cy) + np.abs(xx - cx))).ravel() * (1 + 0.1 * rng.random(om.state_count)) * 10.0 ** rng.integers(-2, 1)
    alphas = alphas.astype(np.float32).astype(np.float64)
    acts = rng.integers(0, 4, G)
    sc = B @ alphas.T
    ref_v, ref_g = sc.max(1), sc.argmax(1)
    srt = np.sort(sc, axis=1)
    clear = (srt[:, -1] - srt[:, -2]) > 1e-5 * np.abs(srt[:, -1])
    assert clear.mean() > 0.5
    vf = onb.ValueFunction(m, alphas, acts)
    # a belief set carried with a deferred normaliser (one device update, then compared on ITS normalised rows) as well as fresh rows
    for bs, rows in ((onb.BeliefSet(m, B), B),):
        for path, f16 in ((1, '1'), (2, '1'), (2, '0')):
            monkeypatch.setenv('OLFNAV_GEMM_F16', f16)
            val, arg, act = vf.evaluate_device(bs, path)
            np.testing.assert_allclose(val.cpu().numpy(), ref_v, rtol=1e-5, err_msg=f'path {path} f16 {f16}')
            assert np.array_equal(arg.cpu().numpy()[clear], ref_g[clear]), f'path {path} f16 {f16}'
            assert np.array_equal(act.cpu().numpy()[clear], acts[ref_g][clear])


@pytest.mark.parametrize('f16', ['1', '0'])
def test_backup_c2_vs_oracle(f16, monkeypatch):
    import olfactory_navigation_b200 as onb
    om = _oracle()['model']
    import pomdp_toolkit as tk
    _, agent = make('C2')
    m = agent.model
    monkeypatch.setenv('OLFNAV_GEMM_F16', f16)
    rng = np.random.default_rng(7)
    nb, G = 256, 48 if f16 == '0' else 130          # G = 130: the FP16 split kernel serves the backup from G = 128 up
    if f16 == '1':
        nb = 128                                    # keep the oracle's (A, O, G, S) float64 tensor below 400 MB
    B = _beliefs(m, om, nb, rng)
    yy, xx = np.meshgrid(np.arange(83), np.arange(371), indexing='ij')
    alphas = np.empty((G, om.state_count))
    for g in range(G):
        cy, cx = rng.integers(0, 83), rng.integers(0, 371)
        alphas[g] = (0.99 ** (np.abs(yy - cy) + np.abs(xx - cx))).ravel() * (1 + 0.1 * rng.random(om.state_count))
    alphas = alphas.astype(np.float32).astype(np.float64)
    _, det = tk.solvers.backup(om, tk.BeliefSet(om, B), tk.ValueFunction(om, alphas, np.zeros(G, dtype=int)), gamma=0.99, append=False,
                               belief_dominance_prune=False, return_details=True)
    gam = tk.solvers.gamma_a_o(om, alphas, 0.99)
    sc = np.einsum('bs,aogs->baog', B, gam)
    s_sorted = np.sort(sc, axis=3)
    q_sorted = np.sort(det['q'], axis=1)
    scale = np.abs(det['new_value'])[:, None, None] + 1e-30
    g_clear = (s_sorted[..., -1] - s_sorted[..., -2]) > 1e-5 * scale
    a_clear = (q_sorted[:, -1] - q_sorted[:, -2]) > 1e-5 * np.abs(det['new_value'])
    vf = onb.ValueFunction(m, alphas, np.zeros(G, dtype=int))
    new_alpha, best_a, best_v, best_g = onb.solvers.backup_device(m, onb.BeliefSet(m, B), vf.device_alphas(0), 0.99, 2)
    best_a, best_g = best_a.cpu().numpy(), best_g.cpu().numpy()
    assert g_clear.mean() > 0.3 and a_clear.mean() > 0.3
    assert np.array_equal(best_g[g_clear], det['best_g'][g_clear])
    assert np.array_equal(best_a[a_clear], det['best_a'][a_clear])
    np.testing.assert_allclose(best_v.cpu().numpy(), det['new_value'], rtol=1e-5)
    dense = onb.ValueFunction(m, action_list=np.zeros(nb, int), _device_alphas=new_alpha).alpha_vector_array
    rows = a_clear & np.array([g_clear[b, det['best_a'][b]].all() for b in range(nb)])
    assert rows.sum() > 10
    np.testing.assert_allclose(dense[rows], det['new_alpha'][rows], rtol=1e-5, atol=1e-9)


def test_fsvi_training_c2_strict_and_the_75_vs_74_ties():
    '''
    FSVI at the benchmark shape, device solver against oracle solver (the bench's training call: expansions = 10).
    (0) Value iteration: 918 sweeps on both sides, Q within 1e-6; the two MDP policies differ only at states where the best two
        actions are tied in float64 (wrap_vertical with the source on the middle row makes north and south equivalent).
    (1) With one MDP policy on both sides the two trainings select the same belief points and hold the strict solver contract (same
        number of alpha vectors, same value function) expansion by expansion — until a backup meets an EXACT tie: a belief point
        whose two best actions have the same float64 value (relative gap below 1e-9; measured: 1.9e-16 between "south" and "west" at
        expansion 3).  np.argmax and the device then keep different (equally good) vectors, which is how the default trainings come to
        75 and 74 vectors.  The test finds that backup and proves the tie; a divergence that is not an exact tie fails it.
    '''
    import olfactory_navigation_b200 as onb
    om = _oracle()['model']
    import pomdp_toolkit as tk
    _, agent = make('C2')
    m = agent.model
    vi_o, hist_o = tk.solvers.VI.solve(om, horizon=1000, gamma=0.99, eps=1e-6)
    vi_d, hist_d = onb.solvers.VI.solve(m, horizon=1000, gamma=0.99, eps=1e-6)
    Qo, Qd = vi_o.alpha_vector_array, vi_d.alpha_vector_array
    assert len(hist_o.value_function_changes) == len(hist_d.value_function_changes) == 918
    np.testing.assert_allclose(Qd, Qo, rtol=1e-6)
    po, pd = np.argmax(Qo, axis=0), np.argmax(Qd, axis=0)
    diff = np.flatnonzero(po != pd)
    for Q in (Qo, Qd):
        gap = np.abs(Q[po[diff], diff] - Q[pd[diff], diff])
        assert np.all(gap <= 1e-9 * np.abs(Q[po[diff], diff])), 'MDP policies differ at a state that is not a float64 tie'

    shared = onb.ValueFunction(m, Qo, np.arange(4))
    seed = onb.synthetic.SEED
    prev = None
    tie_found = False
    for E in range(1, 11):
        kw = dict(expansions=E, gamma=0.99, eps=1e-6, print_progress=False, print_stats=False)
        ref_vf, ref_hist = tk.solvers.FSVI.solve(om, rng=np.random.default_rng(seed), mdp_policy=vi_o, **kw)[:2]
        got_vf, got_hist = onb.solvers.FSVI.solve(m, rng=np.random.default_rng(seed), mdp_policy=shared, **kw)[:2]
        # the belief points never depend on the value function in FSVI: identical all the way
        assert got_hist.beliefs_counts == ref_hist.beliefs_counts
        rb = ref_hist.final_belief_set.belief_array
        np.testing.assert_allclose(got_hist.final_belief_set.belief_array, rb, rtol=2e-5, atol=1e-12)
        if len(got_vf) == len(ref_vf) and np.array_equal(np.sort(got_vf.actions), np.sort(ref_vf.actions)):
            # strict contract: same value function as a function, same actions
            rng = np.random.default_rng(3)
            pts = -np.log(1.0 - rng.random((32, om.state_count)))
            pts = np.vstack([rb, pts / pts.sum(1, keepdims=True)])
            np.testing.assert_allclose(np.max(pts @ got_vf.alpha_vector_array.T, axis=1), np.max(pts @ ref_vf.alpha_vector_array.T, axis=1),
                                       rtol=1e-5, atol=1e-9)
            prev = (ref_vf, ref_hist)
            continue
        # first expansion whose backup differs: redo it in float64 from the (common) state before it
        p_vf = prev[0] if prev else tk.ValueFunction(om, om.expected_rewards_table.T.copy(), om.actions)
        nb = rb[(prev[1].beliefs_counts[-1] if prev else 1):]
        _, det = tk.solvers.backup(om, tk.BeliefSet(om, nb), p_vf, gamma=0.99, append=True, belief_dominance_prune=True, return_details=True)
        vf_prev = onb.ValueFunction(m, p_vf.alpha_vector_array, p_vf.actions)
        _, best_a, best_v, best_g = onb.solvers.backup_device(m, onb.BeliefSet(m, nb.astype(np.float32).astype(np.float64)), vf_prev.device_alphas(0), 0.99, 0)
        best_a, best_g = best_a.cpu().numpy(), best_g.cpu().numpy()
        np.testing.assert_allclose(best_v.cpu().numpy(), det['new_value'], rtol=1e-5)
        # the divergence must be explained by exact float64 ties of this backup: a belief point whose two best actions (or whose two
        # best alpha vectors for an (action, observation) of the chosen action) have the same value to 1e-9 relative.  Which side of
        # such a tie a solver lands on depends on the last bit of its beliefs (the device's come from chained float32 updates).
        q = det['q']
        q_sorted = np.sort(q, axis=1)
        tie_a = (q_sorted[:, -1] - q_sorted[:, -2]) <= 1e-9 * np.abs(q_sorted[:, -1])
        sc = np.einsum('bs,aogs->baog', nb, tk.solvers.gamma_a_o(om, p_vf.alpha_vector_array, 0.99))
        srt = np.sort(sc, axis=3)
        scale_b = np.abs(det['new_value'])[:, None, None] + 1e-300
        tie_g_all = (srt[..., -1] - srt[..., -2]) <= 1e-9 * scale_b
        tie_g = np.array([tie_g_all[b, det['best_a'][b]].any() for b in range(len(nb))])
        assert tie_a.any() or tie_g.any(), f'expansion {E}: the solvers part ways without an exact float64 tie in the backup'
        # away from the ties the device makes the oracle's choices
        clear_a = (q_sorted[:, -1] - q_sorted[:, -2]) > 1e-5 * np.abs(q_sorted[:, -1])
        assert np.array_equal(best_a[clear_a], det['best_a'][clear_a])
        g_clear = (srt[..., -1] - srt[..., -2]) > 1e-5 * scale_b
        assert np.array_equal(best_g[g_clear], det['best_g'][g_clear])
        tie_found = True
        break
    # at this shape, seed and training call the tie is met (the 75-vs-74 of the bench); if a future change removes it, the strict
    # contract above has held for all ten expansions instead
    assert tie_found or (len(got_vf) == len(ref_vf) and np.array_equal(np.sort(got_vf.actions), np.sort(ref_vf.actions)))
