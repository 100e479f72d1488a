'''
Value-function evaluation (olfnav_evaluate) — FP32 SIMT path and tcgen05 3xTF32 path — and the
PBVI backup (gamma build / select / assemble) against golden vectors and the oracle.
Values within 1e-5 relative; argmax / action indices exact wherever the float64 margin between the two
best candidates exceeds 1e-5 relative (documented ties below that).
'''
import numpy as np
import pytest
import torch

from conftest import ALL_CFGS, golden, make, oracle_model

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('cfg', ALL_CFGS)
@pytest.mark.parametrize('path', [1, 2])
def test_evaluate_golden(cfg, path):
    import olfactory_navigation_b200 as onb
    g = golden(f'evaluate_{cfg}')
    _, agent = make(cfg)
    m = agent.model
    vf = onb.ValueFunction(m, g['alphas'].astype(np.float64), g['alpha_actions'])
    val, arg, act = vf.evaluate_device(onb.BeliefSet(m, g['beliefs'].astype(np.float64)), path)
    np.testing.assert_allclose(val.cpu().numpy(), g['values'], rtol=1e-5)
    clear = g['margin'] > 1e-5 * np.abs(g['values'])
    assert clear.sum() >= 0.8 * len(clear)
    assert np.array_equal(arg.cpu().numpy()[clear], g['argmax'][clear])
    assert np.array_equal(act.cpu().numpy()[clear], g['actions'][clear])


@pytest.mark.parametrize('G', [1, 4, 16, 17, 40, 130, 300])
def test_evaluate_paths_agree_many_sizes(G):
    'Row counts that do not fill a 128-row tile, G across every tile width (32/64/128/256 and 2 N-tiles).'
    import olfactory_navigation_b200 as onb
    _, agent = make('C1')
    m = agent.model
    rng = np.random.default_rng(G)
    n = 333
    d = -np.log(1.0 - rng.random((n, m.state_count)))
    B = (d / d.sum(1, keepdims=True)).astype(np.float32).astype(np.float64)
    alphas = rng.standard_normal((G, m.state_count)).astype(np.float32).astype(np.float64)
    acts = rng.integers(0, 4, G)
    vf = onb.ValueFunction(m, alphas, acts)
    bs = onb.BeliefSet(m, B)
    sc = B @ alphas.T
    ref_v, ref_g = sc.max(1), sc.argmax(1)
    srt = np.sort(sc, axis=1)
    clear = (srt[:, -1] - srt[:, -2] > 1e-5 * np.abs(srt).max()) if G > 1 else np.ones(n, bool)
    for path in (1, 2):
        val, arg, act = vf.evaluate_device(bs, path)
        np.testing.assert_allclose(val.cpu().numpy(), ref_v, rtol=0, atol=1e-5 * np.abs(sc).max())   # signed random alphas: error scales with sum|a||b|
        assert np.array_equal(arg.cpu().numpy()[clear], ref_g[clear])
        assert np.array_equal(act.cpu().numpy()[clear], acts[ref_g][clear])


def test_first_maximum_wins_on_exact_ties():
    'Duplicate alpha vectors (backups emit them routinely): the lowest index must win on both paths.'
    import olfactory_navigation_b200 as onb
    _, agent = make('T1')
    m = agent.model
    rng = np.random.default_rng(0)
    base = rng.standard_normal((3, m.state_count)).astype(np.float32).astype(np.float64)
    alphas = np.vstack([base, base, base[::-1]])
    vf = onb.ValueFunction(m, alphas, np.arange(9) % 4)
    B = np.abs(rng.standard_normal((50, m.state_count)))
    B /= B.sum(1, keepdims=True)
    ref = (B.astype(np.float32).astype(np.float64) @ base.T).argmax(1)          # index among the three distinct vectors
    for path in (1, 2):
        _, arg, _ = vf.evaluate_device(onb.BeliefSet(m, B), path)
        assert np.array_equal(arg.cpu().numpy(), ref)


def test_evaluate_full_size_tensor_vs_simt():
    'BASELINE shape, 4096 agents, G=64: tcgen05 path == FP32 SIMT path (values 1e-5, argmax where clear).'
    import olfactory_navigation_b200 as onb
    _, agent = make('C2')
    m = agent.model
    rng = np.random.default_rng(1)
    bs = onb.BeliefSet(m, 4096)
    act = rng.integers(0, 4, 4096)
    for _ in range(3):
        bs = bs.update(act, np.zeros(4096, dtype=int))
    alphas = np.abs(rng.standard_normal((64, m.state_count))).astype(np.float32).astype(np.float64)
    vf = onb.ValueFunction(m, alphas, np.arange(64) % 4)
    v1, g1, _ = vf.evaluate_device(bs, 1)
    v2, g2, _ = vf.evaluate_device(bs, 2)
    np.testing.assert_allclose(v2.cpu().numpy(), v1.cpu().numpy(), rtol=1e-5)
    assert float((g1 == g2).float().mean()) > 0.999


@pytest.mark.parametrize('cfg', ALL_CFGS)
@pytest.mark.parametrize('path', [1, 2])
def test_backup_golden(cfg, path):
    import olfactory_navigation_b200 as onb
    from olfactory_navigation_b200._lib import lib, ptr, stream, check
    g = golden(f'backup_{cfg}')
    _, agent = make(cfg)
    m = agent.model
    bs = onb.BeliefSet(m, g['beliefs'].astype(np.float64))
    vf = onb.ValueFunction(m, g['alphas'].astype(np.float64), np.zeros(len(g['alphas']), dtype=int))
    alphas = vf.device_alphas(0)
    if g['gamma_ao'].size:
        # Gamma_{a,o} itself
        A, O, G, Sp = m.action_count, m.observation_count, alphas.shape[0], m.padded_states
        gam = torch.empty((A * O * G, Sp), dtype=torch.float32, device=alphas.device)
        check(lib.olfnav_backup_gamma(m.handle(0), ptr(alphas), alphas.stride(0), G, float(g['discount']), ptr(gam), Sp, stream()))
        dense = torch.empty((A * O * G, m.state_count), dtype=torch.float32, device=alphas.device)
        check(lib.olfnav_unpack_rows(m.handle(0), ptr(gam), Sp, A * O * G, None, None, ptr(dense), stream()))
        np.testing.assert_allclose(dense.cpu().numpy().reshape(A, O, G, -1), g['gamma_ao'], rtol=2e-6, atol=0)
    new_alpha, best_a, best_v, best_g = onb.solvers.backup_device(m, bs, alphas, float(g['discount']), path)
    best_a, best_g = best_a.cpu().numpy(), best_g.cpu().numpy()
    scale = np.abs(g['new_value'])[:, None, None] + 1e-30
    g_clear = g['g_margin'] > 1e-5 * scale
    assert g_clear.mean() > 0.3
    assert np.array_equal(best_g[g_clear], g['best_g'][g_clear])
    a_clear = g['a_margin'] > 1e-5 * np.abs(g['new_value'])
    assert a_clear.mean() > 0.5
    assert np.array_equal(best_a[a_clear], g['best_a'][a_clear])
    np.testing.assert_allclose(best_v.cpu().numpy(), g['new_value'], rtol=1e-5)
    # assembly: alpha^b = r_{a*} + sum_o Gamma_{a*,o}^{g*}, checked with the DEVICE's own choices so ties do not matter
    from oracle import refloader
    refloader.install_oracle_toolkit()
    import pomdp_toolkit as tk
    om = oracle_model(m)
    gam = tk.solvers.gamma_a_o(om, g['alphas'].astype(np.float64), float(g['discount']))      # (A, O, G, S)
    nb = len(best_a)
    ref = om.expected_rewards_table.T[best_a] + sum(gam[best_a, o, best_g[np.arange(nb), best_a, o], :] for o in range(m.observation_count))
    dense = onb.ValueFunction(m, action_list=np.zeros(nb, int), _device_alphas=new_alpha).alpha_vector_array
    np.testing.assert_allclose(dense, ref, rtol=1e-5, atol=1e-7)
    # and where nothing was tied the vectors are the oracle's
    rows = a_clear & np.array([g_clear[b, g['best_a'][b]].all() for b in range(nb)])
    np.testing.assert_allclose(dense[rows], g['new_alpha'][rows], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize('cfg', ALL_CFGS)
def test_vi_918_sweeps(cfg):
    import olfactory_navigation_b200 as onb
    g = golden(f'vi_{cfg}')
    _, agent = make(cfg)
    vf, hist = onb.solvers.VI.solve(agent.model, horizon=1000, gamma=0.99, eps=1e-6)
    assert len(hist.value_function_changes) == 918
    np.testing.assert_allclose(hist.value_function_changes, g['changes'], rtol=1e-9, atol=1e-15)
    np.testing.assert_allclose(vf.alpha_vector_array, g['Q'], rtol=1e-6)


def test_backup_full_size_properties():
    '''
    BASELINE C3 shape (83x371 states, wrap_vertical), 2 000 belief points against 64 alpha vectors — size-independent properties of the
    backup: (1) the value reported for a point is the value of ITS new vector at that point (selection and assembly are consistent);
    (2) with non-negative alpha vectors the new vector is worth at least the best immediate reward at its point; (3) padding
    columns of the new vectors are zero.
    '''
    import olfactory_navigation_b200 as onb
    _, agent = make('C2')
    m = agent.model
    rng = np.random.default_rng(4)
    B, G = 2000, 64
    bs = onb.BeliefSet(m, B)
    for _ in range(5):
        bs = bs.update(rng.integers(0, 4, B), (rng.random(B) < 0.15).astype(int), raise_on_impossible_belief=False)
    n = len(bs)
    alphas = torch.as_tensor(np.abs(rng.standard_normal((G, m.padded_states))).astype(np.float32), device=bs._b.device)
    H, W = m.grid_shape
    alphas.view(G, H, -1)[:, :, W:] = 0
    new_alpha, best_a, new_v, best_g = onb.solvers.backup_device(m, bs, alphas, 0.99, 2)
    dense = bs.belief_tensor(torch.float64)                                       # normalised rows (n, S), reference state order
    na = new_alpha.view(n, H, -1)[:, :, :W].reshape(n, -1).double()
    at_point = (dense * na).sum(1)
    np.testing.assert_allclose(new_v.double().cpu().numpy(), at_point.cpu().numpy(), rtol=2e-5)
    assert float(new_alpha.view(n, H, -1)[:, :, W:].abs().max()) == 0.0
    r = torch.as_tensor(m.expected_rewards_table.T.copy(), device=dense.device)    # (A, S)
    immediate = (dense @ r.T).max(1).values
    assert bool((at_point >= immediate - 1e-6).all())                              # alphas >= 0: the backup can only add to the immediate reward


@pytest.mark.parametrize('G', [5, 16, 75])
def test_tail_split_of_the_last_wave(G):
    '''
    More than one wave of 128-row tiles (20 000 rows = 157 tiles on 148 SMs): the 9 tiles of the last wave are cut along K into parts
    that run on different SMs and are added by a finishing kernel.  Values and argmax must agree with the FP32 SIMT path and with
    the same rows evaluated in a small batch (no split); the infotaxis sums (store mode) go through the same code.
    '''
    import olfactory_navigation_b200 as onb
    _, agent = make('C1')
    m = agent.model
    rng = np.random.default_rng(G)
    n = 20000
    bs = onb.BeliefSet(m, n)
    for _ in range(3):
        bs = bs.update(rng.integers(0, 4, n), np.zeros(n, dtype=int), raise_on_impossible_belief=False)
    n = len(bs)
    alphas = np.abs(rng.standard_normal((G, m.state_count))).astype(np.float32).astype(np.float64)
    vf = onb.ValueFunction(m, alphas, np.arange(G) % 4)
    v2, a2, _ = vf.evaluate_device(bs, 2)
    v1, a1, _ = vf.evaluate_device(bs, 1)
    np.testing.assert_allclose(v2.cpu().numpy(), v1.cpu().numpy(), rtol=1e-5)
    assert float((a2 == a1).float().mean()) > 0.999
    tail = onb.BeliefSet(m, _buffers=(bs._b[n - 1200:n].contiguous(), bs._scale[n - 1200:n].contiguous(), 1200), device=0)
    vt, at, _ = vf.evaluate_device(tail, 2)
    np.testing.assert_allclose(v2[n - 1200:].cpu().numpy(), vt.cpu().numpy(), rtol=2e-6)
    assert float((a2[n - 1200:] == at).float().mean()) > 0.999
    if G == 5:
        c = onb.synthetic.config('C1')
        it = onb.agents.Infotaxis_Agent(agent.environment, thresholds=c['thresholds'])
        it.initialize_state(n, onb.BeliefSet(it.model, _buffers=(bs._b.clone(), bs._scale.clone(), n), device=0))
        # production path (FP32 tensor-core sums through the tail split + float64 re-evaluation of near ties) against the float64
        # kernel: the same decisions, in the large batch and in the small one
        ids_all = it.choose_action_device().clone()
        ids_all_exact = it.choose_action_device(return_delta_H=True).clone()
        dh = it.last_delta_H.double().cpu().numpy()
        srt = np.sort(dh, axis=1)
        clear = torch.as_tensor((srt[:, -1] - srt[:, -2]) > 1e-6 * np.abs(dh).max(1), device=ids_all.device)
        assert bool((ids_all == ids_all_exact)[clear].all())
        it.initialize_state(1200, onb.BeliefSet(it.model, _buffers=(bs._b[n - 1200:n].clone(), bs._scale[n - 1200:n].clone(), 1200), device=0))
        ids_t = it.choose_action_device()
        assert bool((ids_all[n - 1200:] == ids_t)[clear[n - 1200:]].all())


def test_tail_split_in_the_backup_gemm():
    '''
    Backup selection with several n tiles and a small last wave (896 points x 20 segments of 160 columns = 7 x 25 = 175 units on 148
    SMs: the 27 tail units are cut along K): the tensor path must pick the same alpha per (belief, action, observation) as the FP32
    SIMT path wherever the margin is clear, and the same best action / value.
    '''
    import olfactory_navigation_b200 as onb
    _, agent = make('C1')
    m = agent.model
    rng = np.random.default_rng(12)
    B, G = 896, 150
    bs = onb.BeliefSet(m, B)
    for _ in range(4):
        bs = bs.update(rng.integers(0, 4, B), np.zeros(B, dtype=int), raise_on_impossible_belief=False)
    n = len(bs)
    alphas = torch.as_tensor(np.abs(rng.standard_normal((G, m.padded_states))).astype(np.float32), device=bs._b.device)
    H, W = m.grid_shape
    alphas.view(G, H, -1)[:, :, W:] = 0
    na2, a2, v2, g2 = onb.solvers.backup_device(m, bs, alphas, 0.99, 2)
    na1, a1, v1, g1 = onb.solvers.backup_device(m, bs, alphas, 0.99, 1)
    np.testing.assert_allclose(v2.cpu().numpy(), v1.cpu().numpy(), rtol=2e-5)
    assert float((a2 == a1).float().mean()) > 0.995
    assert float((g2 == g1).float().mean()) > 0.995
    same = (a2 == a1) & (g2 == g1).flatten(1).all(1)
    assert torch.equal(na2[same], na1[same])


@pytest.mark.parametrize('B,G', [(896, 150), (300, 5), (1000, 64), (130, 33)])
def test_backup_fp16_split_gemm_against_simt_and_tf32(B, G, monkeypatch):
    '''
    The backup selection's tensor path runs a 2-way FP16 split with power-of-two block scales (umma_f16_kernel, three kind::f16 MMAs
    per K = 16, cross terms and main term in one accumulator through scale-input-d); OLFNAV_GEMM_F16=0 selects the 3xTF32 kernel.
    Both must agree with the FP32 SIMT path: values to 1e-5 (measured: 5e-7 for FP16, 2e-6 for TF32), same choices.
    Alpha vectors span six decades so that the per-vector scale matters.
    '''
    import olfactory_navigation_b200 as onb
    _, agent = make('C1')
    m = agent.model
    rng = np.random.default_rng(12)
    bs = onb.BeliefSet(m, B)
    for _ in range(4):
        bs = bs.update(rng.integers(0, 4, B), np.zeros(B, dtype=int), raise_on_impossible_belief=False)
    vals = 100 * np.abs(rng.standard_normal((G, m.padded_states))) ** 3 * 10.0 ** rng.integers(-4, 2, (G, 1))
    alphas = torch.as_tensor(vals.astype(np.float32), device=bs._b.device)
    H, W = m.grid_shape
    alphas.view(G, H, -1)[:, :, W:] = 0
    na1, a1, v1, g1 = onb.solvers.backup_device(m, bs, alphas, 0.99, 1)
    monkeypatch.setenv('OLFNAV_GEMM_F16', '1')
    na2, a2, v2, g2 = onb.solvers.backup_device(m, bs, alphas, 0.99, 2)
    monkeypatch.setenv('OLFNAV_GEMM_F16', '0')
    na3, a3, v3, g3 = onb.solvers.backup_device(m, bs, alphas, 0.99, 2)
    for v, a, g in ((v2, a2, g2), (v3, a3, g3)):
        np.testing.assert_allclose(v.cpu().numpy(), v1.cpu().numpy(), rtol=1e-5)
        assert float((a == a1).float().mean()) > 0.995 and float((g == g1).float().mean()) > 0.995
    same = (a2 == a1) & (g2 == g1).flatten(1).all(1)
    assert torch.equal(na2[same], na1[same])


def test_backup_cta_pair_kernel_against_simt_and_single_cta(monkeypatch):
    '''
    umma_f16_pair_kernel (opt-in, OLFNAV_GEMM_PAIR=1): two CTAs of a cluster issue one tcgen05.mma.cta_group::2 over 256 belief rows and
    each loads half of the Gamma tile.  Same split scheme as the single-CTA kernel (two K blocks per accumulator lifetime instead of
    one): values against the FP32 SIMT path to 1e-5 and the same choices; against the single-CTA kernel to 2e-6.
    OLFNAV_GEMM_PAIR=2 makes the call fail if the pair kernel does not apply, so the test knows which kernel ran.
    '''
    import olfactory_navigation_b200 as onb
    _, agent = make('C1')
    m = agent.model
    rng = np.random.default_rng(21)
    B, G = 5333, 150                      # 21 pair tiles x 15 column tiles, ragged last tile; G not a multiple of 16
    bs = onb.BeliefSet(m, B)
    for _ in range(4):
        bs = bs.update(rng.integers(0, 4, B), np.zeros(B, dtype=int), raise_on_impossible_belief=False)
    vals = 100 * np.abs(rng.standard_normal((G, m.padded_states))) ** 3 * 10.0 ** rng.integers(-4, 2, (G, 1))
    alphas = torch.as_tensor(vals.astype(np.float32), device=bs._b.device)
    H, W = m.grid_shape
    alphas.view(G, H, -1)[:, :, W:] = 0
    na1, a1, v1, g1 = onb.solvers.backup_device(m, bs, alphas, 0.99, 1)
    monkeypatch.setenv('OLFNAV_GEMM_PAIR', '0')
    na2, a2, v2, g2 = onb.solvers.backup_device(m, bs, alphas, 0.99, 2)
    monkeypatch.setenv('OLFNAV_GEMM_PAIR', '2')
    na3, a3, v3, g3 = onb.solvers.backup_device(m, bs, alphas, 0.99, 2)
    np.testing.assert_allclose(v3.cpu().numpy(), v1.cpu().numpy(), rtol=1e-5)
    assert float((a3 == a1).float().mean()) > 0.995 and float((g3 == g1).float().mean()) > 0.995
    np.testing.assert_allclose(v3.cpu().numpy(), v2.cpu().numpy(), rtol=2e-6)
    assert float((a3 == a2).float().mean()) > 0.999 and float((g3 == g2).float().mean()) > 0.999
    same = (a3 == a1) & (g3 == g1).flatten(1).all(1)
    assert torch.equal(na3[same], na1[same])
    # a problem the pair kernel does not take is refused under OLFNAV_GEMM_PAIR=2
    small = onb.BeliefSet(m, 300)
    with pytest.raises(RuntimeError, match='pair'):
        onb.solvers.backup_device(m, small, alphas, 0.99, 2)


def test_backup_fp16_split_handles_unnormalised_and_dead_rows():
    'Rows carried with a deferred normaliser (sum << 1) and an all-zero row: the row scale comes from the normaliser, dead rows give zeros.'
    import olfactory_navigation_b200 as onb
    from olfactory_navigation_b200._lib import lib, ptr, stream, check
    _, agent = make('C1')
    m = agent.model
    rng = np.random.default_rng(5)
    B, G = 200, 40
    bs = onb.BeliefSet(m, B)
    for _ in range(3):
        bs = bs.update(rng.integers(0, 4, B), np.zeros(B, dtype=int), raise_on_impossible_belief=False)
    n = len(bs)
    alphas = torch.as_tensor(np.abs(rng.standard_normal((G, m.padded_states))).astype(np.float32), device=bs._b.device)
    H, W = m.grid_shape
    alphas.view(G, H, -1)[:, :, W:] = 0
    factor = torch.as_tensor(10.0 ** rng.uniform(-25, 0, n), dtype=torch.float32, device=bs._b.device)
    raw = (bs._b[:n] * (bs._scale[:n] * factor)[:, None]).contiguous()          # rows with sums from 1e-25 to 1
    inv = (1.0 / factor).contiguous()
    raw[7] = 0
    inv[7] = float('inf')
    scaled = onb.BeliefSet(m, _buffers=(raw, inv, n), device=0)
    _, a1, v1, g1 = onb.solvers.backup_device(m, scaled, alphas, 0.99, 1)
    _, a2, v2, g2 = onb.solvers.backup_device(m, scaled, alphas, 0.99, 2)
    keep = torch.ones(n, dtype=torch.bool, device=raw.device)
    keep[7] = False
    np.testing.assert_allclose(v2[keep].cpu().numpy(), v1[keep].cpu().numpy(), rtol=1e-5)
    assert float((g2[keep] == g1[keep]).float().mean()) > 0.995 and float((a2[keep] == a1[keep]).float().mean()) > 0.995


@pytest.mark.parametrize('G', [128, 515])
def test_evaluate_fp16_split_gemm(G, monkeypatch):
    '''
    From G = 128 up evaluate_at's tensor path is the 2-way FP16 split kernel (OLFNAV_GEMM_F16=0: 3xTF32).  Both against float64 on
    rows carried with a deferred normaliser, signed alpha vectors whose magnitudes span eight decades.
    '''
    import olfactory_navigation_b200 as onb
    _, agent = make('C1')
    m = agent.model
    rng = np.random.default_rng(G)
    n = 700
    bs = onb.BeliefSet(m, n)
    for _ in range(5):
        bs = bs.update(rng.integers(0, 4, n), np.zeros(n, dtype=int), raise_on_impossible_belief=False)
    B = bs.belief_array
    alphas = (rng.standard_normal((G, m.state_count)) * 10.0 ** rng.integers(-6, 3, (G, 1))).astype(np.float32).astype(np.float64)
    acts = rng.integers(0, 4, G)
    vf = onb.ValueFunction(m, alphas, acts)
    sc = B @ alphas.T
    norm = np.abs(B) @ np.abs(alphas).T
    ref_g = sc.argmax(1)
    srt = np.sort(sc, axis=1)
    clear = srt[:, -1] - srt[:, -2] > 1e-5 * norm.max(1)
    assert clear.mean() > 0.5
    for flag in ('1', '0'):
        monkeypatch.setenv('OLFNAV_GEMM_F16', flag)
        val, arg, act = vf.evaluate_device(bs, 2)
        err = np.abs(val.cpu().numpy() - sc.max(1)) / norm[np.arange(len(B)), ref_g]
        assert err.max() < 1e-5, (flag, err.max())
        assert np.array_equal(arg.cpu().numpy()[clear], ref_g[clear])
