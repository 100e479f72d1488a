#!/usr/bin/env python
'''
bench.py — agent belief-steps/s of the device-resident run_test step (BASELINE.json configs[1]):
run_test of a trained FSVI agent on the 55x247 plume with 14/62 margins (83x371 = 30 793 states),
100 000 parallel agents per GPU, synthetic seeded plume (olfactory_navigation_b200.synthetic 'C2').

A "step" = one pass of the hot path over all live agents: choose_action (value-function evaluation)
-> move / observe / source check / discretise -> fused belief update (+ compaction of finished agents).

  value     device-resident loop (inputs already in HBM), agent-steps / s, all ranks
  e2e       the same through the public API (`olfactory_navigation_b200.run_test`) with HOST inputs and
            the per-step history records copied back to pinned host memory inside the timed region
  roofline  fused belief-step kernel: 8 * S * live-agents algorithmic bytes per launch / CUDA-event time
  cpu_baseline / --impl reference   the oracle's NumPy restatement of the reference loop (oracle/sim.py)
            on the host cores (the reference is pure Python over the absent pomdp_toolkit and cannot travel)

Launch: python bench.py --gpus N --steps K --warmup W   (N > 1 under torchrun, one rank per GPU).
'''
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIG = 'C2'
POLICY = 'fsvi'             # 'infotaxis' for --config C4
MAIN = 'steps'              # 'backup' for --config C3
N_AGENTS = 100_000
WORKLOADS = {'C2': '55x247 plume + 14/62 margins = 83x371 = 30793 states, 4 actions, 3 observations, wrap_vertical',
             'C5': '110x494 plume + 28/124 margins = 166x742 = 123172 states, 4 actions, 3 observations, wrap_vertical'}
# DRAM bytes per updated agent from ncu --set full captures (dram__bytes_read.sum + dram__bytes_write.sum per launch / rows updated)
TRAFFIC = {'fused2': ((12.809274e9 + 12.781066e9) / 100_000, 'profiles/r2p_ncu_full_raw_fused2.csv (C2, 100k agents, G=75: 12.81 GB read + 12.78 GB written per launch)'),
           'belief_step': ((12.737818e9 + 12.680899e9) / 99_953, 'profiles/r2q_ncu_full_raw_belief_step.csv (C4 run, 99 953 live agents, rows padded to 384: 12.74 GB read + 12.68 GB written per launch)')}
TRAIN = dict(expansions=10)          # the reference notebook's FSVI training call (experiments/pomdp_agent_training.ipynb cell 4)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--agents', type=int, default=None, help='agents per GPU (default: 100 000 for C2, 125 000 for C5)')
    ap.add_argument('--config', default='C2', choices=['C2', 'C3', 'C4', 'C5'],
                    help='C2 = BASELINE configs[1] (the metric\'s configuration, default); C3 = configs[2], PBVI backups of 10 000 belief points '
                         '(the line\'s value is belief-backups/s at --backup-alphas, FSVI training wall time reported next to it); '
                         'C4 = configs[3], Infotaxis agent on the C2 plume; C5 = configs[4], the 2x-resolution plume of the scaling sweep (10^6 agents over 8 GPUs)')
    ap.add_argument('--cpu-agents', type=int, default=None, help='agents of the bounded CPU sample (default: 256 for C2, 64 for C5: same CPU work per sample)')
    ap.add_argument('--cpu-steps', type=int, default=12)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-backup', action='store_true')
    ap.add_argument('--backup-points', type=int, default=10_000)
    ap.add_argument('--backup-alphas', type=int, default=512)
    args = ap.parse_args()
    global CONFIG, POLICY, MAIN
    MAIN = 'backup' if args.config == 'C3' else 'steps'
    POLICY = 'infotaxis' if args.config == 'C4' else 'fsvi'
    CONFIG = 'C5' if args.config == 'C5' else 'C2'          # C3 and C4 run on the C2 plume
    if args.agents is None:
        args.agents = N_AGENTS if CONFIG == 'C2' else 125_000
    if args.cpu_agents is None:
        args.cpu_agents = 256 if CONFIG == 'C2' else 64
    return args


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    'nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe).'
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self) -> dict:
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[k] for r in self.rows if len(r) >= 7 for k in range(4) if r[3 + k].lower().startswith('active')})
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(sm)}


# =================================================================================================
_CPU_SETUP = None
_SYN = None


def synthetic_module():
    '''
    The synthetic plume generator, loaded by file path: it is pure NumPy, and importing it through the package would dlopen
    libolfnav_b200.so (the package __init__ loads the library) — the CPU reference arm must not map any of the repo's native code.
    '''
    global _SYN
    if _SYN is None:
        import importlib.util
        spec = importlib.util.spec_from_file_location('_olfnav_synthetic', os.path.join(ROOT, 'olfactory-navigation_b200', 'synthetic.py'))
        _SYN = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_SYN)
    return _SYN


def cpu_threads() -> int:
    '''
    Let the CPU arm use every host core: torchrun exports OMP_NUM_THREADS=1 to its workers, which would pin NumPy's BLAS to one
    thread.  Returns the BLAS thread count actually in force (the oracle's element-wise NumPy passes are single-threaded by nature).
    '''
    try:
        from threadpoolctl import threadpool_limits, threadpool_info
        threadpool_limits(limits=os.cpu_count())
        return max([int(i.get('num_threads', 1)) for i in threadpool_info()] or [1])
    except Exception:
        return 1


def cpu_setup():
    'Oracle-side model of the same workload (NumPy fp64), built once.'
    global _CPU_SETUP
    if _CPU_SETUP is not None:
        return _CPU_SETUP
    cpu_threads()
    from oracle import refloader
    refloader.install_oracle_toolkit()
    from oracle import sim
    syn = synthetic_module()
    c = syn.CONFIGS[CONFIG]
    data = syn.plume_frames(c['T'], c['h'], c['w'], source_y=c['source'][0], seed=syn.SEED)
    thr = np.array([-np.inf, c['thresholds'], np.inf])
    _CPU_SETUP = (sim, sim.build_model(data, c['margins'], c['source'], c['radius'], c['boundary'], thr, c['start_zone'], 0.0))
    return _CPU_SETUP


def cpu_sample(alphas: np.ndarray, actions: np.ndarray, n: int, steps: int, seed: int = 0) -> dict:
    sim, setup = cpu_setup()
    import pomdp_toolkit as tk
    vf = tk.ValueFunction(setup['model'], alphas, actions)
    r = sim.timed_agent_steps(setup, vf, n, steps, seed)
    return r


def oracle_value_function(G_target: int = None):
    'The value function of the reference arm, trained without a GPU: the oracle\'s own FSVI with the bench\'s training call (untimed setup).'
    sim, setup = cpu_setup()
    import pomdp_toolkit as tk
    syn = synthetic_module()
    vf = tk.solvers.FSVI.solve(setup['model'], gamma=0.99, eps=1e-6, rng=np.random.default_rng(syn.SEED), print_progress=False, **TRAIN)[0]
    return vf.alpha_vector_array, vf.actions


def run_reference(args) -> None:
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = cpu_threads()
    alphas, actions = oracle_value_function()
    for _ in range(max(args.warmup, 0)):
        cpu_sample(alphas, actions, min(args.cpu_agents, 32), 2)
    total_steps, total_time = 0, 0.0
    for k in range(args.steps):
        r = cpu_sample(alphas, actions, args.cpu_agents, args.cpu_steps, seed=k)
        total_steps += r['agent_steps']
        total_time += r['seconds']
    value = total_steps / total_time
    sample = (f'oracle/sim.py NumPy fp64 restatement of the reference loop (reference = pure Python over the absent pomdp_toolkit); '
              f'{args.cpu_agents} agents x {args.cpu_steps} sim steps per bench step on the {CONFIG} model, FSVI value function trained by the oracle (expansions=10, G={len(alphas)}), '
              f'single process, NumPy/BLAS threads')
    line = {'impl': 'reference', 'metric': 'agent_belief_steps_per_s', 'value': value, 'unit': 'agent-steps/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total_time / max(args.steps, 1), 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': f'{CONFIG} run_test of a trained FSVI agent: ' + WORKLOADS[CONFIG] + ', bounded CPU sample', 'agents_per_step': args.cpu_agents,
                       'alpha_vectors': int(len(alphas))},
            'cpu_baseline': {'value': value, 'unit': 'agent-steps/s', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': value, 'unit': 'agent-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


# =================================================================================================
def measure_backup(onb, model, world, rank, dev, args):
    '''
    One PBVI backup (Gamma_{a,o} build -> belief x Gamma tcgen05 GEMM with per-(a,o) argmax -> alpha assembly) of B belief points
    against G alpha vectors; the belief points are split across the ranks and the new alpha vectors all-gathered (strong scaling:
    B is the whole job's).  Device time over 3 repetitions after 2 warm-ups, max over ranks.
    '''
    import torch
    import torch.distributed as dist
    from olfactory_navigation_b200 import parallel
    B, G = args.backup_points, args.backup_alphas
    rng = np.random.default_rng(99)                                   # same belief points and alphas on every rank
    bs = onb.BeliefSet(model, B)
    for _ in range(6):
        bs = bs.update(rng.integers(0, model.action_count, B), (rng.random(B) < 0.15).astype(int))
    alphas = torch.as_tensor(np.abs(rng.standard_normal((G, model.padded_states))).astype(np.float32), device=dev)
    for _ in range(2):
        parallel.backup_sharded(model, bs, alphas, 0.99)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        new_alpha, best_a = parallel.backup_sharded(model, bs, alphas, 0.99)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    assert new_alpha.shape[0] == B
    flops = 2.0 * B * model.padded_states * ((G + 15) // 16 * 16) * model.action_count * model.observation_count
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
    f16 = os.environ.get('OLFNAV_GEMM_F16', '1') != '0'              # the backup GEMM runs the 2-way FP16 split unless switched off
    tensor_peak = peaks.get('bf16_tflops', 1590.0) / (1.0 if f16 else 2.0)
    issued = 3.0 * flops / (ms * 1e-3) / 1e12 / world
    return {'value': B / (ms * 1e-3), 'unit': 'belief-backups/s', 'belief_points': B, 'alpha_vectors': G, 'ms_per_backup': ms, 'scaling': 'strong',
            'collective': 'all_gather of the per-belief selection ((1 + A*O) int32 per belief point: best action + alpha indices) per backup; every rank assembles all new alpha vectors from its own Gamma' if world > 1 else 'none (1 rank)',
            'roofline': {'kernel': 'umma_f16_kernel' if f16 else 'umma_ts_kernel', 'bound': 'tensor', 'achieved': issued, 'peak': tensor_peak,
                         'unit': 'TFLOP/s', 'frac': issued / tensor_peak,
                         'note': ('issued FP16 flop (3 kind::f16 passes of the 2-way split scheme) per GPU over the WHOLE backup time (Gamma build, plane '
                                  'split, GEMM, assembly, all-gather); peak = measured cuBLAS bf16 burst') if f16 else
                                 ('issued TF32 flop (3 passes of the split-precision scheme) per GPU over the WHOLE backup time (Gamma build, GEMM, '
                                  'assembly, all-gather); peak = measured cuBLAS bf16 burst / 2 (TF32 runs at half the bf16 rate)')}}


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist
    import olfactory_navigation_b200 as onb
    from olfactory_navigation_b200._lib import lib, ptr, stream, check

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    # ---- workload: environment, trained agent (untimed) --------------------------------------------
    c = onb.synthetic.config(CONFIG)
    env = onb.Environment(**c['env_kwargs'])
    if POLICY == 'infotaxis':
        agent = onb.agents.Infotaxis_Agent(env, thresholds=c['thresholds'], rng=onb.synthetic.SEED)
        G, train_s = 0, None
    else:
        # FSVI training through the Agent API (under torchrun every backup of the solve is sharded over the ranks, the selection
        # all-gathered with NCCL); timed once after a warm-up training, wall clock between device synchronisations
        agent = onb.agents.FSVI_Agent(env, thresholds=c['thresholds'], rng=onb.synthetic.SEED)
        agent.train(print_progress=False, print_stats=False, **TRAIN)
        agent = onb.agents.FSVI_Agent(env, thresholds=c['thresholds'], rng=onb.synthetic.SEED)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t_train = time.perf_counter()
        agent.train(print_progress=False, print_stats=False, **TRAIN)
        torch.cuda.synchronize(dev)
        train_s = time.perf_counter() - t_train
        G = len(agent.value_function)
    model = agent.model
    fused = (POLICY == 'fsvi' and os.environ.get('OLFNAV_FUSED_STEP', '1') != '0'
             and lib.olfnav_can_fuse_policy(model.handle(local), agent.value_function.handle(local)) == 1)
    S = model.state_count
    n = args.agents
    rng = np.random.default_rng(onb.synthetic.SEED + 17 * rank)         # every rank simulates its own shard of agents
    starts = env.random_start_points(n, rng)
    tshift = rng.integers(0, env.timesteps, n)

    from olfactory_navigation_b200 import simulation as onb_sim

    def device_loop(steps: int, timed: bool):
        '''
        The device-resident step loop: the same `olfnav_sim_step` call per step as run_test (policy -> environment step ->
        compaction plan -> fused belief update), start points already uploaded, no history records copied out.
        Kernel times come from CUDA events recorded by the library on the launching stream around the policy and the
        fused belief update.
        '''
        timing = [] if timed else None
        common = dict(environment=env, time_loop=True, horizon=steps, initialization_values={}, reward_discount=0.99,
                      distance_metric='l1', print_progress=False)
        hist = onb_sim._run_single(agent, n, starts, tshift, record=False, timing=timing, **common)
        torch.cuda.synchronize(dev)
        loop_start = timing[0][1][0] if timing else None      # CUDA event recorded when the first step starts (beliefs, start points and
                                                              # time shifts are resident in HBM by then)
        bs_ms, ev_ms, bs_bytes = [], [], []
        for (live, evs, counts) in (timing or []):
            # policy: inside the call on the first step (evs[0..1]), enqueued ahead of time for every later step (evs[5..6])
            ev_ms.append(evs[5].elapsed_time(evs[6]) if len(evs) > 6 else evs[0].elapsed_time(evs[1]))
            bs_ms.append(evs[2].elapsed_time(evs[3]))
            bs_bytes.append(8.0 * S * counts)            # survivors actually updated by the kernel
        return hist._agent_steps, bs_ms, bs_bytes, ev_ms, loop_start, getattr(hist, '_step_mode', 'two_kernel')

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- value: device-resident loop -------------------------------------------------------------------
    device_loop(args.warmup, False)
    clocks = ClockSampler(local)
    barrier()
    if rank == 0:
        clocks.start()
    launches0 = lib.olfnav_kernel_launches()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    torch.cuda.nvtx.range_push('timed')
    agent_steps, bs_ms, bs_bytes, ev_ms, loop_start, step_mode = device_loop(args.steps, True)
    fused = step_mode == 'one_pass'
    torch.cuda.nvtx.range_pop()
    t1.record()
    barrier()
    launches = int(lib.olfnav_kernel_launches() - launches0)
    # `value`: the K steps with inputs already resident in HBM, i.e. from the start of the first step (the one-off upload of the
    # start points and the broadcast of the initial belief happen before it); `e2e` below times everything from host buffers
    setup_ms = t0.elapsed_time(loop_start)
    elapsed_ms = loop_start.elapsed_time(t1)

    # ---- e2e: public API, host buffers in, history records out ------------------------------------------
    onb.run_test(agent, n=n, start_points=starts, time_shift=tshift, horizon=args.warmup, print_progress=False, print_stats=False, batches=1)
    barrier()
    t2 = torch.cuda.Event(enable_timing=True)
    t3 = torch.cuda.Event(enable_timing=True)
    t2.record()
    hist = onb.run_test(agent, n=n, start_points=starts, time_shift=tshift, horizon=args.steps, print_progress=False, print_stats=False, batches=1)
    t3.record()
    barrier()
    clk = clocks.stop() if rank == 0 else None
    e2e_ms = t2.elapsed_time(t3)
    e2e_steps = int(np.sum(np.where(hist.done_at_step >= 0, hist.done_at_step, len(hist))))
    d2h = 22.0 * n + 16.0                                      # one record buffer per step (float64 odor + 2 x int32 position + int32 action + 2 flag
                                                               # bytes per agent slot, copied whole) + the 16-byte survivor counts
    h2d = 16.0 * n / max(len(hist), 1)                         # start points (2 x int32) + time shifts (int64), once per run

    # ---- second half of the BASELINE metric: PBVI belief-backups/s (C3: B = 10 000 belief points, sharded over the ranks,
    #      new alpha vectors all-gathered with NCCL); reported inside the same line under "backup" ----------------------------
    backup = measure_backup(onb, model, world, rank, dev, args) if not args.no_backup else None

    # ---- reduce over ranks (device time = max over ranks; work = sum) ------------------------------------
    stats = torch.tensor([elapsed_ms, e2e_ms, float(agent_steps), float(e2e_steps)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        elapsed_ms, e2e_ms, agent_steps, e2e_steps = float(mx[0]), float(mx[1]), float(sm[2]), float(sm[3])
        d2h, h2d = d2h * world, h2d * world

    if rank == 0:
        peak, peak_src = measured_peaks()
        ach = sum(bs_bytes) / (sum(bs_ms) * 1e-3) / 1e9 if sum(bs_ms) > 0 else 0.0
        Sp = model.padded_states
        if fused:
            kernel = 'fused2_kernel<%d>' % ((G + 15) // 16 * 16)
            evaluate_path = 'inside the update kernel (tcgen05 fp16x2-split, A operand = the updated row)'
            what = ('one-pass belief update + policy: 8*S algorithmic bytes per agent-step for BOTH (read b, write b\'); the CUDA events '
                    'bracket the plan kernels, fused2_kernel and the finishing kernel')
        else:
            kernel = 'belief_step_tma_kernel<256>'
            evaluate_path = ('infotaxis: tcgen05 3xtf32 sums + float64 re-evaluation of near ties' if POLICY == 'infotaxis' else
                             'simt-fp32' if G <= 8 else ('tcgen05-fp16x2-split' if (G >= 64 and os.environ.get('OLFNAV_GEMM_F16', '1') != '0') else 'tcgen05-3xtf32'))
            what = 'belief update alone: 8*S algorithmic bytes per agent-step; the policy is a second pass (4*S more bytes), see evaluate_ms_per_step'
        traffic_per_agent, traffic_src = TRAFFIC.get('fused2' if fused else 'belief_step', (None, None))
        label = {'fsvi': 'run_test of a trained FSVI agent', 'infotaxis': 'run_test of an Infotaxis agent (BASELINE configs[3])'}[POLICY]
        line = {
            'metric': 'agent_belief_steps_per_s', 'value': agent_steps / (elapsed_ms * 1e-3), 'unit': 'agent-steps/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': elapsed_ms / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'{args.config} {label}: ' + WORKLOADS[CONFIG],
                       'agents_per_gpu': n, 'alpha_vectors': G, 'evaluate_path': evaluate_path, 'one_pass_step': bool(fused), 'step_mode': step_mode,
                       'step_mode_note': ('multi-rank jobs take the two-kernel step by default: the one-pass kernel (N = 1 default, 4.8 ms/step) has an open '
                                          'fault on fully loaded 8-GPU boxes (DESIGN.md 6a); OLFNAV_FUSED_STEP=1 forces it') if (world > 1 and not fused) else None,
                       'sharding': 'agents split across ranks, no data-path collective',
                       'layout': 'grid rows padded to %d floats (%d padded states for %d states)' % (model.padded_width, Sp, S),
                       'l2': 'belief buffers (2 x %.1f GB) far exceed the 126 MB L2' % (n * Sp * 4 / 1e9)},
            'e2e': {'value': e2e_steps / (e2e_ms * 1e-3), 'unit': 'agent-steps/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
            'gpu_launches': launches, 'setup_ms_outside_value': setup_ms,
            'clocks': clk,
            'roofline': {'kernel': kernel, 'bound': 'hbm', 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak, 'what': what,
                         # DRAM bytes per launch: a constant recorded from one ncu --set full capture of this kernel (file named in
                         # traffic_source), scaled by the rows this run's launches updated — not measured by this run
                         'traffic': traffic_per_agent * (sum(bs_bytes) / (8.0 * S)) / max(len(bs_bytes), 1) if (bs_bytes and traffic_per_agent) else None,
                         'traffic_unit': 'bytes per launch (recorded ncu constant x updated rows)', 'traffic_source': traffic_src,
                         'peak_source': peak_src, 'algorithmic_bytes_per_agent_step': 8 * S,
                         'step_frac_of_8S_bound': (8.0 * S * agent_steps / world) / (elapsed_ms * 1e-3) / 1e9 / peak,
                         'share_of_step': sum(bs_ms) / elapsed_ms, 'evaluate_ms_per_step': statistics.mean(ev_ms) if ev_ms else None,
                         'belief_step_ms_per_step': statistics.mean(bs_ms) if bs_ms else None},
        }
        if train_s is not None:
            line['training'] = {'wall_s': train_s, 'call': 'FSVI_Agent.train(expansions=10)', 'alpha_vectors': G, 'n_gpus': world,
                                'sharded_backups': world > 1}
        if backup is not None:
            line['backup'] = backup
        if not args.no_cpu_baseline and world == 1 and POLICY == 'fsvi':
            vf = agent.value_function
            r = cpu_sample(vf.alpha_vector_array, vf.actions, args.cpu_agents, args.cpu_steps)
            line['cpu_baseline'] = {'value': r['per_s'], 'unit': 'agent-steps/s', 'cores': cpu_threads(), 'kind': 'port',
                                    'sample': f'oracle/sim.py (NumPy fp64 restatement of the reference loop) on {args.cpu_agents} agents x {args.cpu_steps} steps '
                                              f'of the same model and value function (G={G}); {r["agent_steps"]} agent-steps in {r["seconds"]:.1f} s; single process, BLAS threads'}
        if MAIN == 'backup' and backup is not None:
            # --config C3: the backup is the line; the stepping numbers ride along
            steps_line = {k: line[k] for k in ('metric', 'value', 'unit', 'ms_per_step', 'e2e', 'roofline')}
            line.update({'metric': 'pbvi_belief_backups_per_s', 'value': backup['value'], 'unit': backup['unit'], 'ms_per_step': backup['ms_per_backup'],
                         'scaling': 'strong', 'roofline': backup['roofline'], 'stepping': steps_line})
            line['config']['workload'] = (f'C3 PBVI backup of {backup["belief_points"]} belief points against {backup["alpha_vectors"]} alpha vectors on the C2 plume '
                                          f'({WORKLOADS[CONFIG]}); belief points split across the ranks, selection all-gathered')
            line['e2e'] = {'value': backup['value'], 'unit': backup['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0,
                           'note': 'belief points and alpha vectors are device resident in training: the backup has no host leg'}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
