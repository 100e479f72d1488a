'''
Per-row measurements of the hot path on one B200: belief step (out-of-place, in-place, 10^6 agents),
evaluate vs G (SIMT / tcgen05), PBVI backup at C3 scale, infotaxis choice at C4 scale.  One JSON line per row.
CUDA-event timing, >= 3 warm-ups, inputs far larger than L2.  Usage: python scripts/bench_rows.py [rows...]
'''
import json
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import olfactory_navigation_b200 as onb
from olfactory_navigation_b200._lib import lib, ptr, stream, check

torch.cuda.set_device(0); dev = torch.device('cuda', 0)
rows = sys.argv[1:] or ['belief', 'evaluate', 'backup', 'infotaxis', 'million']
peaks = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')) else {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0}
HBM = peaks['hbm_gbs']

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

def emit(**kw): print(json.dumps(kw), flush=True)

c = onb.synthetic.config('C2'); env = onb.Environment(**c['env_kwargs'])
agent = onb.agents.QMDP_Agent(env, thresholds=c['thresholds']); m = agent.model
S, Sp = m.state_count, m.padded_states
rng = np.random.default_rng(0)

def evolved_beliefs(n, steps=3):
    bs = onb.BeliefSet(m, n)
    for _ in range(steps):
        bs = bs.update(rng.integers(0, 4, n), (rng.random(n) < 0.15).astype(int))
    return bs

if 'belief' in rows:
    n = 100_000
    bs = evolved_beliefs(n)
    out_b, out_s = torch.empty_like(bs._b), torch.empty_like(bs._scale)
    ok = torch.empty(n, dtype=torch.uint8, device=dev)
    act = torch.as_tensor(rng.integers(0, 4, n), dtype=torch.int32, device=dev); obs = torch.zeros(n, dtype=torch.int32, device=dev)
    h = m.handle(0)
    t = timeit(lambda: check(lib.olfnav_belief_step(h, ptr(bs._b), ptr(out_b), Sp, n, ptr(act), ptr(obs), None, None, ptr(bs._scale), ptr(out_s), ptr(ok), stream())))
    emit(row='a1 belief_step out-of-place', n=n, S=S, ms=t, GBps=8 * S * n / t / 1e6, frac_hbm=8 * S * n / t / 1e6 / HBM, agent_steps_per_s=n / t * 1e3)
    t = timeit(lambda: check(lib.olfnav_belief_step(h, ptr(bs._b), ptr(bs._b), Sp, n, ptr(act), ptr(obs), None, None, ptr(bs._scale), ptr(bs._scale), ptr(ok), stream())))
    emit(row='a1 belief_step in-place', n=n, S=S, ms=t, GBps=8 * S * n / t / 1e6, frac_hbm=8 * S * n / t / 1e6 / HBM, agent_steps_per_s=n / t * 1e3)
    for a_fixed, name in ((0, 'N'), (1, 'E')):
        act1 = torch.full((n,), a_fixed, dtype=torch.int32, device=dev)
        t = timeit(lambda: check(lib.olfnav_belief_step(h, ptr(bs._b), ptr(out_b), Sp, n, ptr(act1), ptr(obs), None, None, ptr(bs._scale), ptr(out_s), ptr(ok), stream())))
        emit(row=f'a1 belief_step all-{name}', n=n, ms=t, GBps=8 * S * n / t / 1e6, frac_hbm=8 * S * n / t / 1e6 / HBM)
    del bs, out_b, out_s

if 'evaluate' in rows:
    n = 100_000
    bs = evolved_beliefs(n)
    for G in (4, 16, 36, 64, 128, 512, 2048):
        alphas = np.abs(rng.standard_normal((G, S))).astype(np.float32).astype(np.float64)
        vf = onb.ValueFunction(m, alphas, np.arange(G) % 4)
        for path in ((1, 2) if G <= 16 else (2,)):
            t = timeit(lambda: vf.evaluate_device(bs, path), iters=5 if G > 256 else 10)
            flops = 2.0 * n * S * G
            f16 = path == 2 and G >= 64 and os.environ.get('OLFNAV_GEMM_F16', '1') != '0'     # evaluate.cu: 2-way FP16 split planes from G = 64 up
            emit(row='a4 evaluate', G=G, path='simt' if path == 1 else ('tcgen05-fp16x2' if f16 else 'tcgen05-3xtf32'), n=n, ms=t,
                 read_GBps=4 * S * n / t / 1e6, frac_hbm=4 * S * n / t / 1e6 / HBM,
                 useful_TFLOPs=flops / t / 1e9, issued_TFLOPs=3 * 2.0 * n * Sp * (((G + 15) // 16 * 16) if G <= 128 else ((G + 127) // 128 * 128)) / t / 1e9)
        del vf
    del bs

if 'backup' in rows:
    B = 10_000
    bs = evolved_beliefs(B, steps=6)
    for G in (64, 256, 512, 1000, 2048):
        alphas = torch.as_tensor(np.abs(rng.standard_normal((G, Sp))).astype(np.float32), device=dev)
        t = timeit(lambda: onb.solvers.backup_device(m, bs, alphas, 0.99, 2), iters=3, warm=1)
        flops = 2.0 * B * S * G * m.action_count * m.observation_count
        f16 = os.environ.get('OLFNAV_GEMM_F16', '1') != '0' and G >= 128       # solver.cu: 2-way FP16 split from G = 128 up, 3xTF32 below
        emit(row='a6 backup (gamma + select + assemble)', B=B, G=G, ms=t, belief_backups_per_s=B / t * 1e3, useful_TFLOPs=flops / t / 1e9,
             scheme='fp16x2 (3 kind::f16 MMAs per K=16)' if f16 else 'tf32x3 (3 kind::tf32 MMAs per K=8)',
             issued_TFLOPs=3 * flops * (((G + 15) // 16 * 16) / G) / t / 1e9)
        del alphas
    del bs

if 'infotaxis' in rows:
    c4 = onb.agents.Infotaxis_Agent(env, thresholds=c['thresholds'])
    for n in (10_000, 100_000):
        bs = evolved_beliefs(n)
        c4.initialize_state(n, bs)
        t = timeit(lambda: c4.choose_action_device(), iters=3, warm=1)
        emit(row='a9 infotaxis choose', n=n, ms=t, decisions_per_s=n / t * 1e3, read_GBps=4 * S * n / t / 1e6, frac_hbm=4 * S * n / t / 1e6 / HBM)
        del bs
    c4._bufs = None

if 'million' in rows:
    torch.cuda.empty_cache()
    n = 1_000_000
    free, total = torch.cuda.mem_get_info()
    need = n * Sp * 4
    emit(row='north-star memory', free_GB=free / 1e9, total_GB=total / 1e9, beliefs_GB=need / 1e9)
    if free > need * 1.03:
        b = torch.empty((n, Sp), dtype=torch.float32, device=dev); scale = torch.empty(n, dtype=torch.float32, device=dev)
        h = m.handle(0)
        check(lib.olfnav_belief_init(h, ptr(b), Sp, n, ptr(scale), stream()))
        ok = torch.empty(n, dtype=torch.uint8, device=dev)
        act = torch.as_tensor(rng.integers(0, 4, n), dtype=torch.int32, device=dev); obs = torch.zeros(n, dtype=torch.int32, device=dev)
        t = timeit(lambda: check(lib.olfnav_belief_step(h, ptr(b), ptr(b), Sp, n, ptr(act), ptr(obs), None, None, ptr(scale), ptr(scale), ptr(ok), stream())), iters=5, warm=3)
        mass = float((b[:1000].double().sum(1) * scale[:1000].double() - 1).abs().max())
        emit(row='a1 belief_step IN PLACE, 10^6 agents x 83x371 (north star)', n=n, ms=t, GBps=8 * S * n / t / 1e6, frac_hbm=8 * S * n / t / 1e6 / HBM,
             agent_steps_per_s=n / t * 1e3, all_ok=bool(ok.all()), max_mass_error=mass)
