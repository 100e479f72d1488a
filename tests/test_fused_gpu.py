'''
One-pass belief update + value-function evaluation (olfnav_belief_step_evaluate, csrc/fused_step2.cu) against the two separate
C-ABI calls it replaces (olfnav_belief_step then olfnav_evaluate) and against the oracle.
Bar: updated rows BIT-EXACT with olfnav_belief_step (same arithmetic, same operation order); `ok` flags exact; normalisers
within 1e-6 relative (different summation order); values within 1e-5 relative; argmax / action indices exact wherever the
float64 margin between the two best candidates exceeds 1e-5 relative (documented ties below that).
The kernel covers models whose grid rows are padded to 64 floats (every grid of >= 128 columns; OLFNAV_ROW_PAD=64 forces it on the
tiny test grids), vertical moves that wrap and horizontal moves that clip — the BASELINE plume configurations; everything else takes
the two calls (BeliefSet.update_and_evaluate does that by itself, the raw entry point answers OLFNAV_ERR_UNSUPPORTED).
'''
import numpy as np
import pytest
import torch

from conftest import ALL_CFGS, make, oracle_model

pytestmark = pytest.mark.gpu

_FUSABLE = {}


def fusable(name):
    '''
    Fresh agents on models the one-pass kernel covers:
      T1p   the T1 golden geometry (15 x 21, wrap_vertical) with rows padded to 64 floats
      wide  9 x 150 wrap_vertical (padded to 192 by the default rule), thresholds [0.3, 0.6] (O = 4)
      diag  12 x 40 wrap_vertical padded to 64, all nine unit moves (diagonals and "stay")
      C2    the BASELINE shape
    '''
    if name in _FUSABLE:
        return _FUSABLE[name]
    import os
    import olfactory_navigation_b200 as onb
    old = os.environ.get('OLFNAV_ROW_PAD')
    try:
        if name in ('T1p', 'diag'):
            os.environ['OLFNAV_ROW_PAD'] = '64'
        if name == 'T1p':
            c = onb.synthetic.config('T1')
            agent = onb.agents.QMDP_Agent(onb.Environment(**c['env_kwargs']), thresholds=c['thresholds'])
        elif name == 'wide':
            rng = np.random.default_rng(5)
            data = rng.random((8, 5, 140)) * (rng.random((8, 5, 140)) < 0.4)
            env = onb.Environment(data, [2, 0], source_radius=1.0, margins=[2, 5], boundary_condition='wrap_vertical')
            agent = onb.agents.PBVI_Agent(env, thresholds=[0.3, 0.6])
        elif name == 'diag':
            rng = np.random.default_rng(6)
            data = rng.random((4, 12, 40)) * (rng.random((4, 12, 40)) < 0.5)
            env = onb.Environment(data, [6, 0], source_radius=0.5, margins=0, boundary_condition='wrap_vertical')
            moves = np.array([[dy, dx] for dy in (-1, 0, 1) for dx in (-1, 0, 1)])
            agent = onb.agents.PBVI_Agent(env, thresholds=[0.3, 0.6], actions=moves)
        else:
            _, agent = make('C2')
    finally:
        if old is None:
            os.environ.pop('OLFNAV_ROW_PAD', None)
        else:
            os.environ['OLFNAV_ROW_PAD'] = old
    assert agent.model.padded_width % 64 == 0
    _FUSABLE[name] = agent
    return agent


def _is_fused(m, vf):
    from olfactory_navigation_b200._lib import lib
    return lib.olfnav_can_fuse_policy(m.handle(0), vf.handle(0)) == 1


def _random_inputs(m, n, G, seed, point_masses=True):
    rng = np.random.default_rng(seed)
    d = -np.log(1.0 - rng.random((n, m.state_count)))
    d[rng.random((n, m.state_count)) < 0.3] = 0.0
    B = (d / d.sum(1, keepdims=True)).astype(np.float32).astype(np.float64)
    if point_masses:
        k = min(12, n)
        B[:k] = 0
        B[np.arange(k), rng.integers(0, m.state_count, k)] = 1.0
    act = rng.integers(0, m.action_count, n)
    obs = rng.integers(0, m.observation_count - 1, n)
    alphas = np.abs(rng.standard_normal((G, m.state_count))).astype(np.float32).astype(np.float64)
    return B, act, obs, alphas, rng.integers(0, m.action_count, G)


def _separate(onb, m, vf, bs, act, obs, src_rows=None):
    from olfactory_navigation_b200._lib import lib, ptr, stream, check
    dev = bs._b.device
    n = len(act)
    a = torch.as_tensor(act, dtype=torch.int32, device=dev)
    o = torch.as_tensor(obs, dtype=torch.int32, device=dev)
    rows = None if src_rows is None else torch.as_tensor(src_rows, dtype=torch.int32, device=dev)
    out_b = torch.zeros((n, bs._b.shape[1]), dtype=torch.float32, device=dev)
    out_s = torch.empty(n, dtype=torch.float32, device=dev)
    ok = torch.empty(n, dtype=torch.uint8, device=dev)
    check(lib.olfnav_belief_step(m.handle(0), ptr(bs._b), ptr(out_b), bs.ld, n, ptr(a), ptr(o), ptr(rows), None, ptr(bs._scale), ptr(out_s), ptr(ok), stream()))
    new = onb.BeliefSet(m, _buffers=(out_b, out_s, n), device=0)
    val, arg, action = vf.evaluate_device(new, 2)
    return new, ok.bool(), val, arg, action


def _compare(new, ok, val, arg, action, ref, vf_actions):
    rnew, rok, rval, rarg, raction = ref
    assert torch.equal(ok, rok)
    assert torch.equal(new._b[:len(ok)], rnew._b[:len(ok)]), 'updated rows must be bit-identical to olfnav_belief_step'
    good = ok.cpu().numpy()
    s, rs = new._scale[:len(ok)].cpu().numpy()[good], rnew._scale[:len(ok)].cpu().numpy()[good]
    np.testing.assert_allclose(s, rs, rtol=1e-6)
    assert np.all(new._scale[:len(ok)].cpu().numpy()[~good] == 0)
    # value / argmax against float64 products of the (identical) updated rows
    rows = rnew.belief_array[good]
    return rows, good


@pytest.mark.parametrize('cfg', ['T1p', 'wide', 'diag', 'C2'] + ALL_CFGS)
@pytest.mark.parametrize('G', [3, 20, 75, 80])
def test_fused_matches_separate_calls(cfg, G):
    import olfactory_navigation_b200 as onb
    one_pass = cfg in ('T1p', 'wide', 'diag', 'C2')
    agent = fusable(cfg) if one_pass else make(cfg)[1]
    m = agent.model
    n = 300
    B, act, obs, alphas, alpha_actions = _random_inputs(m, n, G, seed=G)
    act[5], obs[6] = 99, -3                                           # invalid ids fail the row
    vf = onb.ValueFunction(m, alphas, alpha_actions)
    assert _is_fused(m, vf) == one_pass
    bs = onb.BeliefSet(m, B)
    ref = _separate(onb, m, vf, bs, act, obs)
    new, ok, val, arg, action = bs.update_and_evaluate(vf, act, obs)
    rows, good = _compare(new, ok, val, arg, action, ref, alpha_actions)
    sc = rows @ alphas.T
    srt = np.sort(sc, axis=1)
    clear = (srt[:, -1] - srt[:, -2]) > 1e-5 * np.abs(srt[:, -1])
    np.testing.assert_allclose(val.cpu().numpy()[good], sc.max(1), rtol=1e-5)
    assert np.array_equal(arg.cpu().numpy()[good][clear], sc.argmax(1)[clear])
    assert np.array_equal(action.cpu().numpy()[good][clear], alpha_actions[sc.argmax(1)][clear])
    assert clear.mean() > 0.9


@pytest.mark.parametrize('cfg', ['T1p', 'wide', 'diag', 'C2'])
@pytest.mark.parametrize('mode', ['store_mode', 'store_sync', 'deferred'])
def test_safe_store_variants_write_the_same_rows(cfg, mode, monkeypatch):
    '''
    The variants of the kernel that wait for the FULL completion of their TMA stores before a shared-memory stage is reused
    (OLFNAV_F2_STORE_MODE=1: the transform warps store their own 32 rows; OLFNAV_F2_STORE_SYNC=1: the store thread waits after every
    block; =2: it releases a stage one block later; DESIGN §6a: the cures of the kernel's open fault) against the two separate calls:
    same bar as the default kernel.
    '''
    import olfactory_navigation_b200 as onb
    monkeypatch.setenv('OLFNAV_F2_STORE_MODE' if mode == 'store_mode' else 'OLFNAV_F2_STORE_SYNC', '2' if mode == 'deferred' else '1')
    agent = fusable(cfg)
    m = agent.model
    n, G = 700, 75
    B, act, obs, alphas, alpha_actions = _random_inputs(m, n, G, seed=11)
    vf = onb.ValueFunction(m, alphas, alpha_actions)
    assert _is_fused(m, vf)
    bs = onb.BeliefSet(m, B)
    ref = _separate(onb, m, vf, bs, act, obs)
    new, ok, val, arg, action = bs.update_and_evaluate(vf, act, obs)
    rows, good = _compare(new, ok, val, arg, action, ref, alpha_actions)
    sc = rows @ alphas.T
    np.testing.assert_allclose(val.cpu().numpy()[good], sc.max(1), rtol=1e-5)


@pytest.mark.parametrize('cfg', ['T1p', 'wide', 'diag', 'T0', 'T2'])
def test_fused_against_oracle(cfg):
    import olfactory_navigation_b200 as onb
    from oracle import refloader
    refloader.install_oracle_toolkit()
    import pomdp_toolkit as tk
    agent = fusable(cfg) if cfg in ('T1p', 'wide', 'diag') else make(cfg)[1]
    m, om = agent.model, oracle_model(agent.model)
    n, G = 200, 9
    B, act, obs, alphas, alpha_actions = _random_inputs(m, n, G, seed=11)
    new, ok, val, arg, action = onb.BeliefSet(m, B).update_and_evaluate(onb.ValueFunction(m, alphas, alpha_actions), act, obs)
    ref, prov = tk.BeliefSet(om, B).update(act, obs, raise_on_impossible_belief=False, return_provenance=True)
    okn = ok.cpu().numpy()
    assert np.array_equal(np.nonzero(okn)[0], prov[:, 0])
    np.testing.assert_allclose(new.belief_array[okn], ref.belief_array, rtol=1e-5, atol=1e-30)
    rv, ra = tk.ValueFunction(om, alphas, alpha_actions).evaluate_at(ref)
    np.testing.assert_allclose(val.cpu().numpy()[okn], rv, rtol=1e-5)
    sc = np.sort(ref.belief_array @ alphas.T, axis=1)
    clear = (sc[:, -1] - sc[:, -2]) > 1e-5 * np.abs(sc[:, -1])
    assert np.array_equal(action.cpu().numpy()[okn][clear], np.asarray(ra)[clear])


@pytest.mark.parametrize('cfg', ['T1p', 'wide', 'T0'])
def test_fused_compacting_row_list(cfg):
    'src_rows: row i of the output is the update of source row src_rows[i] (the survivors of a simulation step).'
    import olfactory_navigation_b200 as onb
    agent = fusable(cfg) if cfg in ('T1p', 'wide') else make(cfg)[1]
    m = agent.model
    n, G = 260, 17
    B, act, obs, alphas, alpha_actions = _random_inputs(m, n, G, seed=5)
    vf = onb.ValueFunction(m, alphas, alpha_actions)
    bs = onb.BeliefSet(m, B)
    rows = np.sort(np.random.default_rng(0).choice(n, 141, replace=False))
    ref = _separate(onb, m, vf, bs, act[rows], obs[rows], src_rows=rows)
    new, ok, val, arg, action = bs.update_and_evaluate(vf, act[rows], obs[rows], src_rows=rows)
    _compare(new, ok, val, arg, action, ref, alpha_actions)
    np.testing.assert_allclose(val.cpu().numpy(), ref[2].cpu().numpy(), rtol=2e-5)


def test_fused_full_size():
    'BASELINE shape (83x371, wrap_vertical), 1000 agents (7.8 tiles), G = 75, three chained steps.'
    import olfactory_navigation_b200 as onb
    _, agent = make('C2')
    m = agent.model
    rng = np.random.default_rng(2)
    n, G = 1000, 75
    alphas = np.abs(rng.standard_normal((G, m.state_count))).astype(np.float32).astype(np.float64)
    alpha_actions = rng.integers(0, 4, G)
    vf = onb.ValueFunction(m, alphas, alpha_actions)
    assert _is_fused(m, vf)
    bs = onb.BeliefSet(m, n)
    for step in range(3):
        act = rng.integers(0, 4, n)
        obs = (rng.random(n) < 0.2).astype(int)
        ref = _separate(onb, m, vf, bs, act, obs)
        new, ok, val, arg, action = bs.update_and_evaluate(vf, act, obs)
        _compare(new, ok, val, arg, action, ref, alpha_actions)
        assert bool(ok.all())
        rv, rarg = ref[2].cpu().numpy(), ref[3].cpu().numpy()
        np.testing.assert_allclose(val.cpu().numpy(), rv, rtol=1e-5)
        assert (arg.cpu().numpy() == rarg).mean() > 0.98
        total = new._b.double().sum(1) * new._scale.double()
        assert float((total - 1).abs().max()) < 1e-5
        bs = new


def test_stream_k_split_many_tiles():
    '''
    20 000 agents on the wide grid = 160 tiles on 148 SMs: every CTA owns a fraction of a tile more than one tile, so most tiles are
    cut into two segments whose partial sums and normaliser shares are added by the finishing kernel.  Uniform and skewed action
    mixes (one action with a single agent, one without any).
    '''
    import olfactory_navigation_b200 as onb
    agent = fusable('wide')
    m = agent.model
    rng = np.random.default_rng(9)
    n, G = 20000, 33
    alphas = np.abs(rng.standard_normal((G, m.state_count))).astype(np.float32).astype(np.float64)
    alpha_actions = rng.integers(0, 4, G)
    vf = onb.ValueFunction(m, alphas, alpha_actions)
    bs = onb.BeliefSet(m, n)
    for _ in range(2):
        bs = bs.update(rng.integers(0, 4, n), np.zeros(n, dtype=int))
    for mix in ('uniform', 'skewed'):
        act = rng.integers(0, 4, n) if mix == 'uniform' else np.where(rng.random(n) < 0.7, 1, 3)
        if mix == 'skewed':
            act[1234] = 0
        obs = (rng.random(n) < 0.1).astype(int)
        ref = _separate(onb, m, vf, bs, act, obs)
        new, ok, val, arg, action = bs.update_and_evaluate(vf, act, obs)
        _compare(new, ok, val, arg, action, ref, alpha_actions)
        good = ok.cpu().numpy()
        np.testing.assert_allclose(val.cpu().numpy()[good], ref[2].cpu().numpy()[good], rtol=1e-5)
        assert (arg.cpu().numpy()[good] == ref[3].cpu().numpy()[good]).mean() > 0.995


def test_unsupported_models_take_the_two_calls():
    'The raw entry point answers UNSUPPORTED outside its coverage; BeliefSet.update_and_evaluate issues the two calls instead.'
    import olfactory_navigation_b200 as onb
    from olfactory_navigation_b200._lib import lib, ptr, stream, OK, ERR_UNSUPPORTED
    _, agent = make('T0')                                              # 'stop' boundary, rows padded to 4
    m = agent.model
    B, act, obs, alphas, alpha_actions = _random_inputs(m, 10, 9, seed=1)
    vf = onb.ValueFunction(m, alphas, alpha_actions)
    assert not _is_fused(m, vf)
    bs = onb.BeliefSet(m, B)
    dev = bs._b.device
    cap = int(lib.olfnav_fused_out_rows(m.handle(0), 10))
    out_b, out_s = torch.zeros((cap, bs._b.shape[1]), device=dev), torch.zeros(cap, device=dev)
    i32 = torch.zeros(10, dtype=torch.int32, device=dev)
    u8 = torch.zeros(10, dtype=torch.uint8, device=dev)
    rc = lib.olfnav_belief_step_evaluate(m.handle(0), vf.handle(0), ptr(bs._b), 10, ptr(out_b), cap, bs.ld, 10, ptr(i32), ptr(i32), None,
                                         ptr(bs._scale), ptr(out_s), ptr(u8), None, None, None, None, stream())
    assert rc == ERR_UNSUPPORTED
    new, ok, val, arg, action = bs.update_and_evaluate(vf, act, obs)
    assert len(new) == 10
    # more than 80 alpha vectors on a model the kernel covers otherwise
    agent = fusable('T1p')
    m = agent.model
    B, act, obs, alphas, alpha_actions = _random_inputs(m, 10, 81, seed=1)
    vf = onb.ValueFunction(m, alphas, alpha_actions)
    assert not _is_fused(m, vf)
    new, ok, val, arg, action = onb.BeliefSet(m, B).update_and_evaluate(vf, act, obs)
    sc = new.belief_array[ok.cpu().numpy()] @ alphas.T
    np.testing.assert_allclose(val.cpu().numpy()[ok.cpu().numpy()], sc.max(1), rtol=1e-5)
