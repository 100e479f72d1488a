'''
Multi-GPU partitioning of the hot path: one process per GPU, `torch.distributed`.

* Simulation: agents are independent, so the start points are split contiguously across ranks exactly like the
  reference splits them into batches (reference simulation.py:1370-1413) and the per-rank histories are
  concatenated like SimulationHistory.__add__ (:550-618).  NO collective on the data path; one object gather
  of the (small) histories at the end.
* Training: the belief points of one backup are split across ranks, every rank holds the full alpha set and builds Gamma,
  and only the per-belief SELECTION (best action + alpha indices) is all-gathered (NCCL over NVLink on GPUs, gloo in the CPU
  tests); every rank then assembles all new alpha vectors locally.
Works with any initialised process group; with none it degrades to the single-process behaviour.
'''
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n: int, rank: int, world_size: int) -> tuple[int, int]:
    'Contiguous split with the remainder on the first ranks (same rule as the reference\'s batch sizes, simulation.py:1372-1374).'
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def run_test_sharded(agent, n: int, start_points: np.ndarray, time_shift: np.ndarray, **kwargs):
    '''
    run_test over every rank's slice of the agents.  Returns the concatenated SimulationHistory on rank 0
    (None elsewhere).  `start_points` / `time_shift` are the full arrays on every rank.
    '''
    from .simulation import run_test
    rank, ws = world()
    lo, hi = shard_bounds(n, rank, ws)
    env = kwargs.get('environment', None)
    if env is None:
        env = getattr(agent, 'environment', None)
    if ws > 1 and rank == 0 and getattr(env, 'reference_time_quirk', False) and len(np.unique(np.asarray(time_shift))) > 1:
        # Under the reference's sorted-time indexing (environment.py:598-601, reproduced by default) agent i observes the frame of the
        # i-th smallest time AMONG THE AGENTS OF ITS BATCH: a shard is a batch of its own, so a sharded run is a valid run of the
        # reference's loop per shard but not the same trajectories as one batch over all agents.
        import warnings
        warnings.warn('run_test_sharded: Environment.reference_time_quirk is on and the time shifts differ between agents: every shard '
                      'observes by its own sorted-time order, results are not comparable with a single-batch run '
                      '(set environment.reference_time_quirk = False for shard-invariant trajectories)', stacklevel=2)
    kwargs.setdefault('print_progress', False)
    kwargs.setdefault('print_stats', False)
    hist = run_test(agent, n=hi - lo, start_points=np.asarray(start_points)[lo:hi], time_shift=np.asarray(time_shift)[lo:hi], **kwargs) if hi > lo else None
    if ws == 1:
        return hist
    if hist is not None:
        hist.environment, hist.agent = None, None              # keep the gathered objects small and picklable
    gathered = [None] * ws if rank == 0 else None
    dist.gather_object(hist, gathered, dst=0)
    if rank != 0:
        return None
    out = None
    for h in gathered:
        if h is None:
            continue
        h.environment, h.agent = agent.environment, agent
        out = h if out is None else out + h
    return out


def assert_same_rng(rng: np.random.Generator, device) -> None:
    '''
    Sharded training keeps the host-side expand step replicated: every rank must draw the same numbers.  Compares a digest of the
    generator state across the ranks (does not advance the generator) and raises on every rank if they differ.
    '''
    rank, ws = world()
    if ws == 1:
        return
    import hashlib
    import pickle
    digest = int.from_bytes(hashlib.sha256(pickle.dumps(rng.bit_generator.state)).digest()[:7], 'little')
    t = torch.tensor([digest, -digest], dtype=torch.int64, device=device if dist.get_backend() == 'nccl' else 'cpu')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if int(t[0]) != -int(t[1]):
        raise RuntimeError('sharded training: the ranks hold different random generators (seed the agent / solver identically on every rank)')


def all_gather_rows(rows: torch.Tensor, extra: torch.Tensor = None, counts: list = None):
    '''
    Concatenate a (k_r, D) tensor (k_r may differ per rank) over all ranks, plus an optional (k_r,) companion.
    One padded all_gather of the payload; the row counts are gathered first unless the caller knows them (`counts`, e.g. from
    shard_bounds: saves a collective and its host synchronisation per backup).
    '''
    rank, ws = world()
    if ws == 1:
        return rows, extra
    if counts is None:
        k = torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device)
        counts = [torch.zeros_like(k) for _ in range(ws)]
        dist.all_gather(counts, k)
        counts = [int(c) for c in counts]
    assert len(counts) == ws and counts[rank] == rows.shape[0]
    kmax = max(max(counts), 1)
    pad = torch.zeros((kmax, rows.shape[1]), dtype=rows.dtype, device=rows.device)
    pad[:rows.shape[0]] = rows
    parts = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(parts, pad)
    out = torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)
    out_extra = None
    if extra is not None:
        epad = torch.zeros(kmax, dtype=extra.dtype, device=extra.device)
        epad[:extra.shape[0]] = extra
        eparts = [torch.empty_like(epad) for _ in range(ws)]
        dist.all_gather(eparts, epad)
        out_extra = torch.cat([p[:c] for p, c in zip(eparts, counts)], dim=0)
    return out, out_extra


def all_gather_alphas(new_alpha: torch.Tensor, actions: torch.Tensor):
    'New alpha vectors (B_r, Sp) + actions (B_r,) of every rank, in rank order.'
    return all_gather_rows(new_alpha, actions)


def backup_sharded(model, belief_set, alphas: torch.Tensor, gamma: float, path: int = 0, exchange: str = 'indices', return_values: bool = False):
    '''
    One PBVI backup with the belief points split across ranks: each rank runs the selection GEMM (argmax over the alpha set for
    every (belief, action, observation)) on rows [lo, hi) of the (replicated) belief set.  Returns (new_alpha, best_a) for ALL
    belief points on every rank, in the original order.

    exchange='indices' (default): the ranks all-gather only the selection — best action + A*O alpha indices per belief point,
    (1 + A*O) ints — and every rank assembles all B new vectors from its own copy of Gamma (a streaming kernel, ~1 ms for
    10 000 points): 0.5 MB on the wire instead of B x Sp floats (1.2 GB at C3).  exchange='alphas': all-gather the assembled
    vectors (the first implementation; kept for comparison).  Both give bit-identical results (same Gamma, same indices).
    return_values=True also returns the value of every new vector at its belief point (gathered with the selection; FSVI / HSVI
    keep a new vector only where it improves on the current value function).
    '''
    from .pomdp import BeliefSet
    from .solvers import backup_device, backup_select_device, backup_assemble_device
    rank, ws = world()
    lo, hi = shard_bounds(len(belief_set), rank, ws)
    part = BeliefSet(model, _buffers=(belief_set._b[lo:hi], belief_set._scale[lo:hi], hi - lo), device=belief_set.device)
    counts = [b - a for a, b in (shard_bounds(len(belief_set), r, ws) for r in range(ws))]
    if exchange == 'alphas' or ws == 1:
        if hi > lo:
            new_alpha, best_a, best_v, _ = backup_device(model, part, alphas, gamma, path)
        else:
            new_alpha = torch.empty((0, alphas.shape[1]), dtype=alphas.dtype, device=alphas.device)
            best_a = torch.empty(0, dtype=torch.int32, device=alphas.device)
            best_v = torch.empty(0, dtype=torch.float32, device=alphas.device)
        out_alpha, out_a = all_gather_alphas(new_alpha, best_a)
        if not return_values:
            return out_alpha, out_a
        out_v = all_gather_rows(best_v[:, None], counts=counts)[0][:, 0].contiguous() if ws > 1 else best_v
        return out_alpha, out_a, out_v
    G = alphas.shape[0]
    ctx, best_g, best_a, best_v = backup_select_device(model, part, alphas, gamma, path)
    # (B_r, 2 + A*O) int32: best action, the value's bit pattern, the alpha indices
    packed = torch.cat([best_a[:, None], best_v.view(torch.int32)[:, None], best_g.reshape(hi - lo, -1)], dim=1).contiguous()
    packed_all, _ = all_gather_rows(packed, counts=counts)
    best_a_all = packed_all[:, 0].contiguous()
    best_v_all = packed_all[:, 1].contiguous().view(torch.float32)
    best_g_all = packed_all[:, 2:].contiguous()
    new_alpha = backup_assemble_device(model, ctx, G, best_g_all, best_a_all, belief_set.device)
    return (new_alpha, best_a_all, best_v_all) if return_values else (new_alpha, best_a_all)
