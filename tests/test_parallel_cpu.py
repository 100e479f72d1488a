'Host-side multi-GPU logic on CPU: world_size-2 gloo process group (127.0.0.1), no CUDA.'
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def test_shard_bounds_cover_everything():
    from olfactory_navigation_b200.parallel import shard_bounds
    for n in [0, 1, 7, 8, 100, 100_003]:
        for ws in [1, 2, 3, 8]:
            spans = [shard_bounds(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[r][1] == spans[r + 1][0] for r in range(ws - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank: int, ws: int, port: int, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dist.init_process_group('gloo', rank=rank, world_size=ws)
    try:
        from olfactory_navigation_b200 import parallel
        import olfactory_navigation_b200 as onb
        assert parallel.world() == (rank, ws)
        # ragged all-gather of alpha vectors: rank r contributes r + 2 rows
        k = rank + 2
        rows = torch.full((k, 5), float(rank)) + torch.arange(k)[:, None] * 0.1
        acts = torch.arange(k, dtype=torch.int32) + 10 * rank
        got, got_a = parallel.all_gather_alphas(rows, acts)
        want = torch.cat([torch.full((r + 2, 5), float(r)) + torch.arange(r + 2)[:, None] * 0.1 for r in range(ws)])
        want_a = torch.cat([torch.arange(r + 2, dtype=torch.int32) + 10 * r for r in range(ws)])
        assert torch.equal(got, want) and torch.equal(got_a, want_a)
        # caller-known counts (shard_bounds): same result without the count collective
        got2, _ = parallel.all_gather_rows(rows, counts=[r + 2 for r in range(ws)])
        assert torch.equal(got2, want)
        # empty contribution from one rank
        e, ea = parallel.all_gather_alphas(rows[:0] if rank == 0 else rows, acts[:0] if rank == 0 else acts)
        assert e.shape[0] == sum((r + 2) for r in range(1, ws))
        # sharded training refuses generators that differ between the ranks (the expand step is replicated on the host)
        parallel.assert_same_rng(np.random.default_rng(5), 'cpu')
        try:
            parallel.assert_same_rng(np.random.default_rng(5 + rank), 'cpu')
            raise AssertionError('different generators were accepted')
        except RuntimeError:
            pass
        # history gather + concatenation in rank order (the pickled histories travel through gather_object)
        c = onb.synthetic.config('T0')
        env = onb.Environment(**c['env_kwargs'])
        agent = onb.Agent(env, thresholds=c['thresholds'])
        n = 7
        starts = np.stack([np.arange(n) + 3, np.arange(n) + 4], axis=1)
        lo, hi = parallel.shard_bounds(n, rank, ws)
        h = onb.SimulationHistory(starts[lo:hi], env, agent, np.zeros(hi - lo, int), horizon=3)
        m = hi - lo
        h.add_step(np.full(m, rank), starts[lo:hi] + 1, np.full(m, 0.5 + rank), np.arange(m) % 2 == 0, np.zeros(m, bool), action_is_id=True)
        h.environment, h.agent = None, None
        gathered = [None] * ws if rank == 0 else None
        dist.gather_object(h, gathered, dst=0)
        if rank == 0:
            out = None
            for g in gathered:
                g.environment, g.agent = env, agent
                out = g if out is None else out + g
            assert out.n == n and np.array_equal(out.start_points, starts)
            assert out.positions[0][:, 0].tolist() == (starts[:, 0] + 1).tolist()
            exp_reached = np.concatenate([np.arange(b - a) % 2 == 0 for a, b in (parallel.shard_bounds(n, r, ws) for r in range(ws))])
            assert np.array_equal(out.reached_source, exp_reached)
        q.put((rank, 'ok'))
    except Exception as e:                       # surface the failure in the parent
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_gloo_world_size_2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == 'ok', f'rank {rank}: {msg}'
