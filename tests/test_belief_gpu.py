'''
Fused belief step (olfnav_belief_step) against the golden vectors and the live oracle.
Tolerance: beliefs within 1e-5 relative (north_star), float32 device path vs float64 oracle fed the
same float32-representable inputs; `ok` flags and survivor sets bit-exact.
'''
import numpy as np
import pytest
import torch

from conftest import ALL_CFGS, golden, make, oracle_model

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-30


def _step(model, B, act, obs, inplace=False, src_rows=None):
    'Raw C-ABI call; returns (normalised dense rows, ok).'
    import olfactory_navigation_b200 as onb
    from olfactory_navigation_b200._lib import lib, ptr, stream, check
    dev = torch.device('cuda', 0)
    bs = onb.BeliefSet(model, np.asarray(B, dtype=np.float64))
    n = len(act)
    a = torch.as_tensor(act, dtype=torch.int32, device=dev)
    o = torch.as_tensor(obs, dtype=torch.int32, device=dev)
    ok = torch.zeros(max(n, 1), dtype=torch.uint8, device=dev)
    if inplace:
        out_b, out_s = bs._b, bs._scale
        rows = None if src_rows is None else torch.as_tensor(src_rows, dtype=torch.int32, device=dev)
        check(lib.olfnav_belief_step(model.handle(0), ptr(bs._b), ptr(out_b), bs.ld, n, ptr(a), ptr(o), ptr(rows), ptr(rows),
                                     ptr(bs._scale), ptr(out_s), ptr(ok), stream()))
        sel = slice(None) if src_rows is None else torch.as_tensor(src_rows, dtype=torch.long, device=dev)
        res = onb.BeliefSet(model, _buffers=(out_b[sel].contiguous(), out_s[sel].contiguous(), n), device=0)
    else:
        out_b = torch.full((max(n, 1), bs._b.shape[1]), float('nan'), dtype=torch.float32, device=dev)
        out_s = torch.empty(max(n, 1), dtype=torch.float32, device=dev)
        rows = None if src_rows is None else torch.as_tensor(src_rows, dtype=torch.int32, device=dev)
        check(lib.olfnav_belief_step(model.handle(0), ptr(bs._b), ptr(out_b), bs.ld, n, ptr(a), ptr(o), ptr(rows), None,
                                     ptr(bs._scale), ptr(out_s), ptr(ok), stream()))
        res = onb.BeliefSet(model, _buffers=(out_b, out_s, n), device=0)
        # padding columns must be exactly zero
        H, W = model.grid_shape
        Wp = out_b.shape[1] // H
        if Wp > W and n:
            assert float(out_b[:n].view(n, H, Wp)[:, :, W:].abs().max()) == 0.0
    return res.belief_array, ok[:n].bool().cpu().numpy()


@pytest.mark.parametrize('cfg', ALL_CFGS)
@pytest.mark.parametrize('inplace', [False, True])
def test_golden(cfg, inplace):
    g = golden(f'belief_{cfg}')
    _, agent = make(cfg)
    out, ok = _step(agent.model, g['b_in'], g['act'], g['obs'], inplace=inplace)
    assert np.array_equal(ok, g['ok'])
    np.testing.assert_allclose(out[ok], g['b_out'][ok], rtol=RTOL, atol=ATOL)
    assert np.all(out[~ok] == 0)


@pytest.mark.parametrize('cfg', ALL_CFGS)
@pytest.mark.parametrize('inplace', [False, True])
def test_golden_wide_grid_shape(cfg, inplace, monkeypatch):
    'The 512-thread / 32 KB-stage shape of the ring kernel (chosen for grids wider than a 16 KB stage) on the same golden vectors.'
    monkeypatch.setenv('OLFNAV_BELIEF_THREADS', '512')
    g = golden(f'belief_{cfg}')
    _, agent = make(cfg)
    out, ok = _step(agent.model, g['b_in'], g['act'], g['obs'], inplace=inplace)
    assert np.array_equal(ok, g['ok'])
    np.testing.assert_allclose(out[ok], g['b_out'][ok], rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize('cfg', ['T0', 'T1', 'T2', 'T3'])
def test_compacting_and_inplace_row_lists(cfg):
    g = golden(f'belief_{cfg}')
    _, agent = make(cfg)
    n = len(g['act'])
    rows = np.sort(np.random.default_rng(1).choice(n, size=n // 2, replace=False))
    out, ok = _step(agent.model, g['b_in'], g['act'][rows], g['obs'][rows], src_rows=rows)            # dst = 0..m-1
    assert np.array_equal(ok, g['ok'][rows])
    np.testing.assert_allclose(out[ok], g['b_out'][rows][ok], rtol=RTOL, atol=ATOL)
    out, ok = _step(agent.model, g['b_in'], g['act'][rows], g['obs'][rows], inplace=True, src_rows=rows)
    assert np.array_equal(ok, g['ok'][rows])
    np.testing.assert_allclose(out[ok], g['b_out'][rows][ok], rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize('boundary', ['stop', 'wrap', 'wrap_vertical', 'wrap_horizontal'])
@pytest.mark.parametrize('shape', [(7, 9), (12, 8), (5, 4), (9, 1030), (1, 6), (6, 1)])
def test_all_unit_moves_vs_oracle(boundary, shape):
    'Every (dy, dx) in {-1,0,1}^2 incl. diagonals and "stay", ragged widths, 1-wide grids, > 1 chunk rows.'
    import olfactory_navigation_b200 as onb
    h, w = shape
    rng = np.random.default_rng(h * 1000 + w)
    data = rng.random((4, h, w)) * (rng.random((4, h, w)) < 0.5)
    env = onb.Environment(data, [h // 2, 0], source_radius=0.5, margins=0, boundary_condition=boundary)
    moves = np.array([[dy, dx] for dy in (-1, 0, 1) for dx in (-1, 0, 1)])
    agent = onb.agents.PBVI_Agent(env, thresholds=[0.3, 0.6], actions=moves)
    m, om = agent.model, oracle_model(agent.model)
    from oracle import refloader
    refloader.install_oracle_toolkit()
    import pomdp_toolkit as tk
    n = 45
    d = -np.log(1.0 - rng.random((n, m.state_count)))
    B = (d / d.sum(1, keepdims=True)).astype(np.float32).astype(np.float64)
    B[:9] = 0
    B[np.arange(9), rng.integers(0, m.state_count, 9)] = 1.0          # point masses
    act = np.arange(n) % 9
    obs = rng.integers(0, m.observation_count - 1, n)
    u, p = tk.BeliefSet(om, B).unnormalised_successors()
    z = p[np.arange(n), act, obs]
    ok_ref = z > 0
    ref = np.zeros_like(B)
    ref[ok_ref] = u[np.arange(n), act, obs][ok_ref] / z[ok_ref, None]
    for inplace in (False, True):
        out, ok = _step(m, B, act, obs, inplace=inplace)
        assert np.array_equal(ok, ok_ref)
        np.testing.assert_allclose(out[ok], ref[ok], rtol=RTOL, atol=ATOL)


def test_invalid_ids_fail_the_row():
    _, agent = make('T0')
    m = agent.model
    B = np.tile(m.start_probabilities, (4, 1))
    out, ok = _step(m, B, [0, 99, 1, -1], [0, 0, 77, 0])
    assert ok.tolist() == [True, False, False, False]


def test_chained_steps_track_the_oracle():
    'Deferred normalisation over 40 chained steps: |delta| <= 1e-5 * max(b) (SURVEY §7 trajectory criterion).'
    import olfactory_navigation_b200 as onb
    from oracle import refloader
    refloader.install_oracle_toolkit()
    import pomdp_toolkit as tk
    _, agent = make('C1')
    m, om = agent.model, oracle_model(agent.model)
    rng = np.random.default_rng(3)
    n = 24
    bs, obs_ref = onb.BeliefSet(m, n), tk.BeliefSet(om, np.tile(om.start_probabilities, (n, 1)))
    for _ in range(40):
        act = rng.integers(0, 4, n)
        obs = np.zeros(n, dtype=int)                      # 'nothing' is always possible away from the source
        bs = bs.update(act, obs)
        obs_ref = obs_ref.update(act, obs)
    got, ref = bs.belief_array, obs_ref.belief_array
    assert np.max(np.abs(got - ref) / ref.max(axis=1, keepdims=True)) < 1e-5
    np.testing.assert_allclose(got.sum(1), 1.0, rtol=1e-5)


def test_full_size_properties():
    'BASELINE shape (83x371, wrap_vertical), 2048 agents: mass conservation, zero padding, in-place rows == out-of-place rows.'
    import olfactory_navigation_b200 as onb
    from olfactory_navigation_b200._lib import lib, ptr, stream, check
    _, agent = make('C2')
    m = agent.model
    dev = torch.device('cuda', 0)
    n = 2048
    bs = onb.BeliefSet(m, n)
    rng = np.random.default_rng(0)
    for step in range(3):
        act = torch.as_tensor(rng.integers(0, 4, n), dtype=torch.int32, device=dev)
        obs = torch.as_tensor((rng.random(n) < 0.2).astype(np.int32), device=dev)
        ok = torch.empty(n, dtype=torch.uint8, device=dev)
        out_b, out_s = torch.empty_like(bs._b), torch.empty_like(bs._scale)
        check(lib.olfnav_belief_step(m.handle(0), ptr(bs._b), ptr(out_b), bs.ld, n, ptr(act), ptr(obs), None, None, ptr(bs._scale), ptr(out_s), ptr(ok), stream()))
        ip_b, ip_s = bs._b.clone(), bs._scale.clone()
        check(lib.olfnav_belief_step(m.handle(0), ptr(ip_b), ptr(ip_b), bs.ld, n, ptr(act), ptr(obs), None, None, ptr(ip_s), ptr(ip_s), ptr(ok), stream()))
        assert torch.equal(ip_b, out_b)                                        # rows: bit exact
        # the normaliser is summed in a different chunk order (traversal direction is free out of place): last-ulp differences only
        assert float(((ip_s - out_s).abs() / out_s).max()) < 1e-6
        assert bool(ok.all())
        total = (out_b.double().sum(1) * out_s.double())
        assert float((total - 1).abs().max()) < 1e-5
        H, W = m.grid_shape
        assert float(out_b.view(n, H, -1)[:, :, W:].abs().max()) == 0.0
        bs = onb.BeliefSet(m, _buffers=(out_b, out_s, n), device=0)
