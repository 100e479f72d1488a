'''
Device solvers — the `pomdp_toolkit.solvers` surface the reference's agents call:
  VI.solve                 reference agents/qmdp_agent.py:145-153 (and implicitly FSVI, agents/fsvi_agent.py:154-158)
  FSVI.solve               reference agents/fsvi_agent.py:201-224
  PBVI_RA/SSRA/SSGA/SSEA   reference agents/pbvi_ra_agent.py:168, pbvi_ssra_agent.py:170,
                           pbvi_ssga_agent.py:176-198, pbvi_ssea_agent.py:170
  PBVI_GER/Perseus/HSVI    reference agents/pbvi_ger_agent.py:170, perseus_agent.py:176-196, hsvi_agent.py:184-208

The backup runs on the device in three steps through the C ABI:
  olfnav_backup_gamma    Gamma_{a,o}^g = gamma * O_o(R_a(.)) * alpha_g(R_a(.))          streaming kernel
  olfnav_backup_select   g*(b,a,o) = argmax_g b.Gamma_{a,o}^g ; a*(b)                    tcgen05 GEMM + segmented argmax
  olfnav_backup_assemble alpha^b = r_{a*} + sum_o Gamma_{a*,o}^{g*}                      streaming kernel
Host code keeps only the control flow, the (belief-free) trajectory sampling and the
duplicate-vector pruning.  Random draws follow the same documented order as the oracle so that
identical generators give identical (action, observation) sequences.
'''
from __future__ import annotations

import os
import time

import numpy as np
import torch

from ._lib import lib, ptr, stream, check
from .pomdp import POMDP, Belief, BeliefSet, ValueFunction, current_device


class SolverHistory:
    'Statistics of a solve(); fields follow the reference\'s TrainingHistory (agents/pbvi_agent.py:139-234).'
    def __init__(self, model: POMDP, tracking_level: int, gamma: float, eps: float, solver: str) -> None:
        self.model, self.tracking_level, self.gamma, self.eps, self.solver = model, tracking_level, gamma, eps, solver
        self.expansion_times, self.backup_times, self.pruning_times = [], [], []
        self.alpha_vector_counts, self.beliefs_counts, self.prune_counts = [], [], []
        self.value_function_changes, self.iteration_times = [], []
        self.value_functions, self.belief_sets = [], []
        self.total_time = 0.0

    @property
    def summary(self) -> str:
        m = self.model
        if self.solver == 'VI':
            n, tot = len(self.value_function_changes), self.total_time
            return (f'Summary of Value Iteration run\n  - Model: {m.state_count}-state, {m.action_count}-action\n'
                    f'  - Converged in {n} iterations and {tot:.4f} seconds\n'
                    f'  - Took on average {(tot / max(n, 1)):.4f}s per iteration\n')
        ne, nb = len(self.expansion_times), len(self.backup_times)
        s = (f'Summary of Point Based Value Iteration run\n'
             f'  - Model: {m.state_count} state, {m.action_count} action, {m.observation_count} observations\n'
             f'  - Converged or stopped after {ne} expansion steps and {nb} backup steps.\n'
             f'  - Resulting value function has {self.alpha_vector_counts[-1] if self.alpha_vector_counts else 0} alpha vectors.\n'
             f'  - Converged in {self.total_time:.4f}s\n\n')
        if ne:
            s += (f'  - Expand function took on average {np.mean(self.expansion_times):.4f}s and yielded on average '
                  f'{np.mean(self.beliefs_counts):.2f} beliefs per iteration.\n')
        if nb:
            s += (f'  - Backup function took on average {np.mean(self.backup_times):.4f}s and yielded on average '
                  f'{np.mean(np.diff([0] + self.alpha_vector_counts)):.2f} alpha vectors per iteration.\n')
        if self.pruning_times:
            s += (f'  - Pruning function took on average {np.mean(self.pruning_times):.4f}s and yielded on average prunings of '
                  f'{np.mean(self.prune_counts):.2f} alpha vectors per iteration.\n')
        return s


# =================================================================================================
class VI:
    @staticmethod
    def solve(model: POMDP, horizon: int = 1000, initial_value_function: ValueFunction = None, gamma: float = 0.99,
              eps: float = 1e-6, use_gpu: bool = True, use_reachability: bool = True, history_tracking_level: int = 1,
              print_progress: bool = True, check_every: int = 16):
        '''
        Q_a(s) = r_a(s) + gamma * max_a' Q_a'(R_a(s)); stops after the first sweep whose
        max |V_k - V_{k-1}| < eps / (1 - gamma) (918 sweeps at gamma=.99, eps=1e-6, the count the
        reference's notebooks print).  float64 sweeps on the device; the per-sweep change is read
        back every `check_every` sweeps and the stop sweep is then replayed exactly.
        '''
        device = current_device()
        dev = torch.device('cuda', device)
        hist = SolverHistory(model, history_tracking_level, gamma, eps, 'VI')
        t0 = time.perf_counter()
        Sp, A = model.padded_states, model.action_count
        h = model.handle(device)
        Q = torch.zeros((2, A, Sp), dtype=torch.float64, device=dev)
        if initial_value_function is not None:
            Q[0] = initial_value_function.device_alphas(device).double()
        threshold = eps / (1.0 - gamma)
        changes = torch.zeros(max(horizon, 1), dtype=torch.float64, device=dev)
        done, cur = 0, 0
        while done < horizon:
            block = min(check_every, horizon - done)
            snapshot = Q[cur].clone()
            for k in range(block):
                check(lib.olfnav_vi_sweep(h, ptr(Q[cur]), ptr(Q[1 - cur]), Sp, gamma, ptr(changes[done + k:]), stream()))
                cur = 1 - cur
            ch = changes[done:done + block].cpu().numpy()
            below = np.flatnonzero(ch < threshold)
            if len(below):
                # replay from the snapshot up to (and including) the first sweep below the threshold
                stop = int(below[0]) + 1
                Q[0] = snapshot
                cur = 0
                changes[done:done + block] = 0
                for k in range(stop):
                    check(lib.olfnav_vi_sweep(h, ptr(Q[cur]), ptr(Q[1 - cur]), Sp, gamma, ptr(changes[done + k:]), stream()))
                    cur = 1 - cur
                done += stop
                break
            done += block
        hist.value_function_changes = changes[:done].cpu().numpy().tolist()
        alphas = Q[cur].float().contiguous()
        vf = ValueFunction(model, action_list=model.actions, _device_alphas=alphas)
        torch.cuda.synchronize(dev)
        hist.total_time = time.perf_counter() - t0
        hist.iteration_times = [hist.total_time / max(done, 1)] * done
        hist.alpha_vector_counts.append(len(vf))
        return vf, hist


# =================================================================================================
def _gamma_free(model: POMDP, G: int, path: int) -> bool:
    '''
    Gamma-free backup (olfnav_backup_select_alphas / _assemble_alphas): the FP16 operand planes are built straight from the alpha
    vectors and the selected Gamma rows recomputed at assembly — bit-identical vectors, no A*O*G*Sp floats in memory (783 MB at
    G = 512, 3.1 GB at G = 2048).  Measured at B = 10 000, G = 512: selection 11.7 ms against 12.1 ms, assembly slower (the rows are
    gathered through the move instead of streamed), so it is taken when Gamma would use more than a quarter of the free device memory,
    or with OLFNAV_BACKUP_GAMMA=0; OLFNAV_BACKUP_GAMMA=1 always keeps Gamma in memory.  Tensor path of grid models from G = 128 up only.
    '''
    if model.dense or path == 1 or G < 128 or os.environ.get('OLFNAV_GEMM_F16', '1') == '0':
        return False
    flag = os.environ.get('OLFNAV_BACKUP_GAMMA', '')
    if flag in ('0', '1'):
        return flag == '0'
    need = model.action_count * model.observation_count * G * model.padded_states * 4
    return need > 0.25 * torch.cuda.mem_get_info(torch.device('cuda', current_device()))[0]


def backup_device(model: POMDP, belief_set: BeliefSet, alphas: torch.Tensor, gamma: float, path: int = 0):
    '''
    One point-based backup of the alpha set `alphas` (G, Sp) at every belief of the set.
    Returns device tensors: new_alpha (B, Sp) f32, best_a (B,) i32, new_value (B,) f32 (value of the
    new vector at its belief), best_g (B, A, O) i32.
    '''
    ctx, best_g, best_a, best_v = backup_select_device(model, belief_set, alphas, gamma, path)
    new_alpha = backup_assemble_device(model, ctx, alphas.shape[0], best_g, best_a, belief_set.device)
    return new_alpha, best_a, best_v, best_g


def backup_select_device(model: POMDP, belief_set: BeliefSet, alphas: torch.Tensor, gamma: float, path: int = 0):
    '''
    The first two thirds of a backup: Gamma (or its operand planes) + per-belief selection.  Returns (ctx, best_g (B, A, O), best_a (B,),
    best_v (B,)); backup_assemble_device turns (ctx, best_g, best_a) — of ANY belief points, e.g. gathered from other ranks — into
    alpha vectors.  ctx is the FP32 Gamma tensor, or ('alphas', alphas, gamma) on the path that never materialises it.
    '''
    device = belief_set.device
    dev = torch.device('cuda', device)
    h = model.handle(device)
    A, O, Sp = model.action_count, model.observation_count, model.padded_states
    G, B = alphas.shape[0], len(belief_set)
    best_g = torch.empty((max(B, 1), A, O), dtype=torch.int32, device=dev)
    best_a = torch.empty(max(B, 1), dtype=torch.int32, device=dev)
    best_v = torch.empty(max(B, 1), dtype=torch.float32, device=dev)
    if _gamma_free(model, G, path):
        alphas = alphas.contiguous()
        if B:
            with torch.cuda.device(device):
                check(lib.olfnav_backup_select_alphas(h, ptr(belief_set._b), belief_set.ld, B, ptr(belief_set._scale), ptr(alphas), alphas.stride(0), G,
                                                      gamma, ptr(best_g), ptr(best_a), ptr(best_v), stream()))
        return ('alphas', alphas, gamma), best_g[:B], best_a[:B], best_v[:B]
    gam = torch.empty((A * O * G, Sp), dtype=torch.float32, device=dev)
    with torch.cuda.device(device):
        check(lib.olfnav_backup_gamma(h, ptr(alphas), alphas.stride(0), G, gamma, ptr(gam), Sp, stream()))
        if B:
            check(lib.olfnav_backup_select(h, ptr(belief_set._b), belief_set.ld, B, ptr(belief_set._scale), ptr(gam), Sp, G,
                                           ptr(best_g), ptr(best_a), ptr(best_v), path, stream()))
    return gam, best_g[:B], best_a[:B], best_v[:B]


def backup_assemble_device(model: POMDP, ctx, G: int, best_g: torch.Tensor, best_a: torch.Tensor, device: int = None):
    'alpha_b = r_{a*} + sum_o Gamma[a*][o][g*(b, a*, o)] for every row of (best_g, best_a); ctx from backup_select_device.'
    device = current_device() if device is None else device
    Sp = model.padded_states
    B = int(best_a.shape[0])
    dev = ctx[1].device if isinstance(ctx, tuple) else ctx.device
    new_alpha = torch.empty((B, Sp), dtype=torch.float32, device=dev)
    if B:
        with torch.cuda.device(device):
            if isinstance(ctx, tuple):
                _, alphas, gamma = ctx
                check(lib.olfnav_backup_assemble_alphas(model.handle(device), ptr(alphas), alphas.stride(0), G, gamma, ptr(best_g.contiguous()),
                                                        ptr(best_a.contiguous()), B, ptr(new_alpha), Sp, stream()))
            else:
                check(lib.olfnav_backup_assemble(model.handle(device), ptr(ctx), Sp, G, ptr(best_g.contiguous()), ptr(best_a.contiguous()), B,
                                                 ptr(new_alpha), Sp, stream()))
    return new_alpha


def _unique_rows(alpha: torch.Tensor, actions: torch.Tensor):
    'Drop exact duplicate (vector, action) rows, keep the first occurrence and the original order.'
    if alpha.shape[0] <= 1:
        return alpha, actions
    keyed = torch.cat([actions.float()[:, None], alpha], dim=1)
    _, inverse = torch.unique(keyed, dim=0, return_inverse=True)
    first = torch.full((int(inverse.max()) + 1,), alpha.shape[0], dtype=torch.long, device=alpha.device)
    first.scatter_reduce_(0, inverse, torch.arange(alpha.shape[0], device=alpha.device), reduce='amin')
    keep = torch.sort(first).values
    return alpha[keep], actions[keep]


class _PBVI:
    name = 'PBVI'
    full_backup = True
    dominance_prune = False

    @classmethod
    def prepare(cls, model, gamma, eps, print_progress, **kw) -> dict:
        return {}

    @classmethod
    def expand(cls, model, belief_set, alphas, actions, max_generation, rng, **kw) -> BeliefSet:
        raise NotImplementedError

    @classmethod
    def solve(cls, model: POMDP, expansions: int = 10, update_passes: int = 1, max_belief_growth: int = 10,
              initial_belief=None, initial_value_function: ValueFunction = None, prune_level: int = 1,
              prune_interval: int = 10, limit_value_function_size: int = -1, gamma: float = 0.99, eps: float = 1e-6,
              convergence_stop: bool = False, use_gpu: bool = True, use_reachability: bool = True,
              rng: np.random.Generator = None, history_tracking_level: int = 1, print_progress: bool = True,
              print_stats: bool = True, gemm_path: int = 0, **flavour):
        rng = np.random.default_rng() if rng is None else rng
        device = current_device()
        dev = torch.device('cuda', device)
        # Under torchrun (one process per GPU, process group initialised) the backups of this solve are sharded over the ranks.
        # The expand step draws on the host: every rank must hold the same generator state (same seed), so that all ranks select the
        # same belief points — checked once per solve through the first draw.
        from . import parallel
        sharded = parallel.world()[1] > 1 and os.environ.get('OLFNAV_SHARDED_TRAIN', '1') != '0'
        if sharded:
            parallel.assert_same_rng(rng, dev)
        hist = SolverHistory(model, history_tracking_level, gamma, eps, cls.name)
        t_start = time.perf_counter()
        extra = cls.prepare(model, gamma, eps, print_progress, use_reachability=use_reachability, **flavour)

        if initial_belief is None:
            belief_set = BeliefSet(model, 1)
        elif isinstance(initial_belief, Belief):
            belief_set = BeliefSet(model, initial_belief.values[None, :])
        else:
            belief_set = initial_belief

        if initial_value_function is None:
            vf0 = ValueFunction(model, model.expected_rewards_table.T.copy(), model.actions)
        else:
            vf0 = initial_value_function
        alphas = vf0.device_alphas(device).clone()
        actions = torch.as_tensor(vf0.actions, dtype=torch.int32, device=dev)

        def values_of(bs, al, ac):
            v, _, _ = ValueFunction(model, action_list=ac.cpu().numpy(), _device_alphas=al).evaluate_device(bs, gemm_path)
            return v

        old_values = values_of(belief_set, alphas, actions)
        backup_count = 0
        for _ in range(expansions):
            t0 = time.perf_counter()
            new_beliefs = cls.expand(model, belief_set, alphas, actions, max_belief_growth, rng, **extra)
            belief_set = belief_set.union(new_beliefs)
            torch.cuda.synchronize(dev)
            hist.expansion_times.append(time.perf_counter() - t0)
            hist.beliefs_counts.append(len(belief_set))

            t0 = time.perf_counter()
            target = belief_set if cls.full_backup else new_beliefs
            for _ in range(update_passes):
                if len(target) == 0:
                    break
                if sharded:
                    # belief points split across the ranks of the process group, selection all-gathered (parallel.backup_sharded):
                    # every rank ends up with the same new vectors, bit-identical to the single-GPU backup
                    new_alpha, best_a, new_v = parallel.backup_sharded(model, target, alphas, gamma, gemm_path, return_values=True)
                else:
                    new_alpha, best_a, new_v, _ = backup_device(model, target, alphas, gamma, gemm_path)
                if cls.dominance_prune:
                    keep = new_v > values_of(target, alphas, actions)
                    new_alpha, best_a = new_alpha[keep], best_a[keep]
                new_alpha, best_a = _unique_rows(new_alpha, best_a)
                if cls.full_backup:
                    alphas, actions = new_alpha, best_a
                else:
                    alphas, actions = _unique_rows(torch.cat([alphas, new_alpha]), torch.cat([actions, best_a]))
                backup_count += 1
            torch.cuda.synchronize(dev)
            hist.backup_times.append(time.perf_counter() - t0)

            t0 = time.perf_counter()
            pruned = 0
            if prune_interval > 0 and (backup_count % prune_interval) == 0 and prune_level >= 1:
                before = alphas.shape[0]
                alphas, actions = prune_alphas(alphas, actions, prune_level)
                pruned = before - alphas.shape[0]
            if 0 < limit_value_function_size < alphas.shape[0]:
                G = alphas.shape[0]
                drop = rng.choice(G, size=min(max_belief_growth, G - 1), replace=False)
                keep = np.ones(G, dtype=bool)
                keep[drop] = False
                keep_t = torch.as_tensor(keep, device=dev)
                alphas, actions = alphas[keep_t], actions[keep_t]
                pruned += int((~keep).sum())
            hist.pruning_times.append(time.perf_counter() - t0)
            hist.prune_counts.append(pruned)
            hist.alpha_vector_counts.append(int(alphas.shape[0]))

            new_values = values_of(belief_set, alphas, actions)
            change = float((new_values[:len(old_values)] - old_values).abs().max()) if len(old_values) else float('inf')
            old_values = new_values
            hist.value_function_changes.append(change)
            if convergence_stop and change < eps:
                break

        value_function = ValueFunction(model, action_list=actions.cpu().numpy(), _device_alphas=alphas.contiguous())
        hist.total_time = time.perf_counter() - t_start
        hist.final_belief_set = belief_set
        if 'mdp_policy' in extra and flavour.get('return_mdp_policy', False):
            return value_function, hist, extra['mdp_policy']
        return value_function, hist


def _chain_updates(model: POMDP, b0: BeliefSet, row: int, seq) -> BeliefSet:
    'Beliefs visited by applying the (action, observation) sequence to row `row` of b0, one kernel launch per step.'
    device = b0.device
    dev = b0._b.device
    n = len(seq)
    Sp = model.padded_states
    out_b = torch.empty((max(n, 1), Sp), dtype=torch.float32, device=dev)
    out_s = torch.empty(max(n, 1), dtype=torch.float32, device=dev)
    ok = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
    if n == 0:
        return BeliefSet(model, _buffers=(out_b, out_s, 0), device=device)
    acts = torch.as_tensor([a for a, _ in seq], dtype=torch.int32, device=dev)
    obss = torch.as_tensor([o for _, o in seq], dtype=torch.int32, device=dev)
    h = model.handle(device)
    with torch.cuda.device(device):
        for k in range(n):
            src_b = b0._b[row:row + 1] if k == 0 else out_b[k - 1:k]
            src_s = b0._scale[row:row + 1] if k == 0 else out_s[k - 1:k]
            check(lib.olfnav_belief_step(h, ptr(src_b), ptr(out_b[k:k + 1]), Sp, 1, ptr(acts[k:k + 1]), ptr(obss[k:k + 1]),
                                         None, None, ptr(src_s), ptr(out_s[k:k + 1]), ptr(ok[k:k + 1]), stream()))
    if not bool(ok[:n].all()):
        raise ValueError('Impossible belief update during expand')
    return BeliefSet(model, _buffers=(out_b, out_s, n), device=device)


class FSVI(_PBVI):
    '''
    Forward Search Value Iteration (Shani et al., cited at agents/fsvi_agent.py:12): one trajectory
    per expansion from the initial belief under the MDP policy; the (action, observation) sequence
    is belief-free so it is sampled on the host and the beliefs are produced by chained belief-step
    launches.  Backup on the new beliefs only, appended (agents/fsvi_agent.py:139).
    '''
    name = 'FSVI'
    full_backup = False
    dominance_prune = True

    @classmethod
    def prepare(cls, model, gamma, eps, print_progress, mdp_policy=None, vi_horizon=1000, use_reachability=True, **kw):
        if mdp_policy is None:
            mdp_policy, _ = VI.solve(model, horizon=vi_horizon, gamma=gamma, eps=eps, print_progress=print_progress)
        return {'mdp_policy': mdp_policy, 'Q': mdp_policy.alpha_vector_array}

    @staticmethod
    def trajectory(model: POMDP, b0: np.ndarray, Q: np.ndarray, max_generation: int, rng) -> list:
        cdf = np.cumsum(b0)
        s = min(int(np.searchsorted(cdf, rng.random() * cdf[-1], side='right')), model.state_count - 1)
        seq = []
        for _ in range(max_generation):
            a = int(np.argmax(Q[:, s]))
            s_p = model.transition(s, a, rng.random() if model.dense else 0.0)
            o = model.observe(s_p, a, rng.random())
            seq.append((a, o))
            s = s_p
            if model.end_mask[s]:
                break
        return seq

    @classmethod
    def expand(cls, model, belief_set, alphas, actions, max_generation, rng, mdp_policy=None, Q=None, **kw) -> BeliefSet:
        b0 = belief_set.belief_tensor(torch.float64)[0].cpu().numpy() if len(belief_set) else model.start_probabilities
        seq = cls.trajectory(model, b0, Q, max_generation, rng)
        return _chain_updates(model, belief_set, 0, seq)


class PBVI_SSRA(_PBVI):
    'Stochastic Simulation with Random Action (same draw order as the oracle).'
    name = 'PBVI_SSRA'

    @classmethod
    def choose_action(cls, model, k, belief_set, alphas, actions, rng, **kw) -> int:
        return min(int(rng.random() * model.action_count), model.action_count - 1)

    @classmethod
    def expand(cls, model, belief_set, alphas, actions, max_generation, rng, **kw) -> BeliefSet:
        m = min(max_generation, len(belief_set))
        dense = belief_set.belief_tensor(torch.float64)[:m].cpu().numpy()
        pieces = None
        for k in range(m):
            cdf = np.cumsum(dense[k])
            s = min(int(np.searchsorted(cdf, rng.random() * cdf[-1], side='right')), model.state_count - 1)
            a = cls.choose_action(model, k, belief_set, alphas, actions, rng, **kw)
            s_p = model.transition(s, a, rng.random() if model.dense else 0.0)
            o = model.observe(s_p, a, rng.random())
            nb = _chain_updates(model, belief_set, k, [(a, o)])
            pieces = nb if pieces is None else pieces.union(nb)
        return pieces if pieces is not None else BeliefSet(model, np.zeros((0, model.state_count)))


class PBVI_SSGA(PBVI_SSRA):
    'Stochastic Simulation with Greedy Action, epsilon-greedy (agents/pbvi_ssga_agent.py:176-198).'
    name = 'PBVI_SSGA'

    @classmethod
    def prepare(cls, model, gamma, eps, print_progress, epsilon=0.1, epsilon_decay=1.0, **kw):
        return {'epsilon_state': {'epsilon': float(epsilon), 'decay': float(epsilon_decay)}}

    @classmethod
    def choose_action(cls, model, k, belief_set, alphas, actions, rng, epsilon_state=None, **kw) -> int:
        if rng.random() < epsilon_state['epsilon']:
            return min(int(rng.random() * model.action_count), model.action_count - 1)
        one = BeliefSet(model, _buffers=(belief_set._b[k:k + 1], belief_set._scale[k:k + 1], 1), device=belief_set.device)
        _, _, act = ValueFunction(model, action_list=actions.cpu().numpy(), _device_alphas=alphas).evaluate_device(one)
        return int(act[0])

    @classmethod
    def expand(cls, model, belief_set, alphas, actions, max_generation, rng, epsilon_state=None, **kw) -> BeliefSet:
        out = super().expand(model, belief_set, alphas, actions, max_generation, rng, epsilon_state=epsilon_state)
        epsilon_state['epsilon'] *= epsilon_state['decay']
        return out


class PBVI_RA(_PBVI):
    'Random belief selection: Dirichlet(1) points from rng.random((n, S)) (same draws as the oracle).'
    name = 'PBVI_RA'

    @classmethod
    def expand(cls, model, belief_set, alphas, actions, max_generation, rng, **kw) -> BeliefSet:
        pts = -np.log(1.0 - rng.random((max_generation, model.state_count)))
        return BeliefSet(model, pts / np.sum(pts, axis=1, keepdims=True))


def _pair_reduce(mode: int, X: torch.Tensor, sx, Y: torch.Tensor, sy, W: torch.Tensor = None, vmax: float = 0.0, vmin: float = 0.0,
                 row_len: int = None) -> torch.Tensor:
    'out[i, j] = reduce_s f(X[i]*sx[i], Y[j]*sy[j])  (olfnav_pair_reduce: 0 = L1 distance, 1 = GER error term, 2 = dominated flag).'
    nx, ny = X.shape[0], Y.shape[0]
    out = torch.empty((nx, ny), dtype=torch.float32, device=X.device)
    row_len = X.shape[1] if row_len is None else row_len
    for i0 in range(0, nx, 65535):
        i1 = min(nx, i0 + 65535)
        check(lib.olfnav_pair_reduce(mode, ptr(X[i0:i1]), X.stride(0), None if sx is None else ptr(sx[i0:i1]), i1 - i0,
                                     ptr(Y), Y.stride(0), None if sy is None else ptr(sy), ny,
                                     None if W is None else ptr(W), 0 if W is None else W.stride(0), vmax, vmin, row_len,
                                     ptr(out[i0:i1]), stream()))
    return out


def _batched_updates(model: POMDP, bs: BeliefSet, rows, acts, obss):
    '''
    One belief-step launch producing b_{a,o} for every (source row, action, observation) triple.
    Returns (BeliefSet of len(rows) rows, ok (bool tensor), p = P(o | b, a) (float32 tensor, 0 where impossible)).
    '''
    dev = bs._b.device
    n = len(rows)
    Sp = model.padded_states
    r = torch.as_tensor(np.asarray(rows), dtype=torch.int32, device=dev).contiguous()
    a = torch.as_tensor(np.asarray(acts), dtype=torch.int32, device=dev).contiguous()
    o = torch.as_tensor(np.asarray(obss), dtype=torch.int32, device=dev).contiguous()
    out_b = torch.empty((max(n, 1), Sp), dtype=torch.float32, device=dev)
    out_s = torch.empty(max(n, 1), dtype=torch.float32, device=dev)
    ok = torch.zeros(max(n, 1), dtype=torch.uint8, device=dev)
    if n:
        with torch.cuda.device(bs.device):
            check(lib.olfnav_belief_step(model.handle(bs.device), ptr(bs._b), ptr(out_b), Sp, n, ptr(a), ptr(o), ptr(r), None,
                                         ptr(bs._scale), ptr(out_s), ptr(ok), stream()))
    okb = ok[:n].bool()
    p = torch.where(okb, 1.0 / out_s[:n].clamp_min(1e-38), torch.zeros_like(out_s[:n]))
    return BeliefSet(model, _buffers=(out_b, out_s, n), device=bs.device), okb, p


def _take_rows(model: POMDP, bs: BeliefSet, rows) -> BeliefSet:
    idx = torch.as_tensor(np.asarray(rows, dtype=np.int64), device=bs._b.device)
    return BeliefSet(model, _buffers=(bs._b[idx].contiguous() if len(idx) else bs._b[:1].clone(),
                                      bs._scale[idx].contiguous() if len(idx) else bs._scale[:1].clone(), len(idx)), device=bs.device)


def prune_alphas(alphas: torch.Tensor, actions: torch.Tensor, level: int):
    '''
    ValueFunction pruning (prune_level of every train(), reference agents/pbvi_agent.py:584-658): level >= 1 drops exact
    duplicates, level >= 2 also vectors point-wise dominated by another single vector (olfnav_pair_reduce, DOMINATED).
    '''
    if level >= 1:
        alphas, actions = _unique_rows(alphas, actions)
    if level >= 2 and alphas.shape[0] > 1:
        a = alphas.contiguous()
        dom = _pair_reduce(2, a, None, a, None)                    # dom[i, j] = 1 if alpha_j dominates alpha_i
        keep = dom.sum(dim=1) == 0
        alphas, actions = a[keep], actions[keep]
    return alphas, actions


class PBVI_SSEA(_PBVI):
    '''
    Stochastic Simulation with Exploratory Action (agents/pbvi_ssea_agent.py:170): for each of the first `max_generation`
    beliefs simulate one step per action and keep the successor farthest (L1) from the current set (which grows as successors
    are picked).  All candidates come from ONE batched belief-step launch and the distances from two olfnav_pair_reduce
    launches (candidates x set, candidates x candidates); the sequential pick is host bookkeeping on those matrices.
    Draws per belief, per action: rng.random() for the state, rng.random() for the observation (same order as the oracle).
    '''
    name = 'PBVI_SSEA'

    @classmethod
    def expand(cls, model, belief_set, alphas, actions, max_generation, rng, **kw) -> BeliefSet:
        m, A = min(max_generation, len(belief_set)), model.action_count
        dense = belief_set.belief_tensor(torch.float64)[:m].cpu().numpy()
        rows, acts, obss = [], [], []
        for k in range(m):
            cdf = np.cumsum(dense[k])
            for a in range(A):
                s = min(int(np.searchsorted(cdf, rng.random() * cdf[-1], side='right')), model.state_count - 1)
                s_p = model.transition(s, a, rng.random() if model.dense else 0.0)
                obss.append(model.observe(s_p, a, rng.random()))
                rows.append(k)
                acts.append(a)
        cand, ok, _ = _batched_updates(model, belief_set, rows, acts, obss)
        if not bool(ok.all()):
            raise ValueError('Impossible belief update during expand')
        Sp = model.padded_states
        n = len(belief_set)
        d_set = _pair_reduce(0, cand._b, cand._scale, belief_set._b[:n], belief_set._scale[:n], row_len=Sp).min(dim=1).values.cpu().numpy()
        d_cc = _pair_reduce(0, cand._b, cand._scale, cand._b[:len(cand)], cand._scale[:len(cand)], row_len=Sp).cpu().numpy()
        chosen = []
        for k in range(m):
            best, best_d = -1, -1.0
            for a in range(A):
                c = k * A + a
                d = float(min([d_set[c]] + [d_cc[c, j] for j in chosen]))
                if d > best_d:
                    best, best_d = c, d
            chosen.append(best)
        return _take_rows(model, cand, chosen)


class PBVI_GER(_PBVI):
    '''
    Greedy Error Reduction (agents/pbvi_ger_agent.py:10,170; Pineau et al. 2006): see the oracle's PBVI_GER for the definition.
    All A*O successors of every belief come from one batched belief-step launch; the error bound of every successor against
    every belief of the set is one olfnav_pair_reduce launch (GER mode) followed by a row minimum.  Deterministic.
    '''
    name = 'PBVI_GER'

    @classmethod
    def prepare(cls, model, gamma, eps, print_progress, **kw) -> dict:
        return {'gamma': gamma}

    @classmethod
    def expand(cls, model, belief_set, alphas, actions, max_generation, rng, gamma=0.99, **kw) -> BeliefSet:
        n, A, O, Sp = len(belief_set), model.action_count, model.observation_count, model.padded_states
        vf = ValueFunction(model, action_list=actions.cpu().numpy(), _device_alphas=alphas)
        _, g_star, _ = vf.evaluate_device(belief_set)
        alpha_b = alphas[g_star.long()].contiguous()
        r = model.expected_rewards_table
        vmax, vmin = float(np.max(r)) / (1.0 - gamma), float(np.min(r)) / (1.0 - gamma)
        rows = np.repeat(np.arange(n), A * O)
        acts = np.tile(np.repeat(np.arange(A), O), n)
        obss = np.tile(np.arange(O), n * A)
        succ, ok, p = _batched_updates(model, belief_set, rows, acts, obss)
        e = _pair_reduce(1, succ._b, succ._scale, belief_set._b[:n], belief_set._scale[:n], W=alpha_b, vmax=vmax, vmin=vmin, row_len=Sp)
        eps_bao = torch.where(ok, e.min(dim=1).values, torch.zeros_like(p)).double().view(n, A, O)
        contrib = (p.double().view(n, A, O) * eps_bao).cpu().numpy()
        p_h = p.view(n, A, O).cpu().numpy()
        err = contrib.sum(axis=2)
        a_star = np.argmax(err, axis=1)
        score = err[np.arange(n), a_star]
        o_star = np.argmax(contrib[np.arange(n), a_star, :], axis=1)
        order = np.argsort(-score, kind='stable')[:max_generation]
        pick = [int(i * A * O + a_star[i] * O + o_star[i]) for i in order if p_h[i, a_star[i], o_star[i]] > 0]
        return _take_rows(model, succ, pick)


class Perseus(_PBVI):
    '''
    Perseus (agents/perseus_agent.py:99,176-196; Spaan & Vlassis 2005): see the oracle's Perseus for the definition.
    The stage backs up ALL beliefs against the stage's starting value function in one batched device backup (every per-belief
    backup of the sequential algorithm uses that same value function, so they are independent), takes the cross values
    candidate_i . b_j from one FP32 matrix product, and replays the randomised improve-only selection on the host.
    '''
    name = 'Perseus'

    @classmethod
    def walk(cls, model, belief_set, alphas, actions, max_generation, rng, use_policy) -> BeliefSet:
        n = len(belief_set)
        b0 = belief_set.belief_tensor(torch.float64)[n - 1].cpu().numpy()
        cdf = np.cumsum(b0)
        s = min(int(np.searchsorted(cdf, rng.random() * cdf[-1], side='right')), model.state_count - 1)
        cur = _take_rows(model, belief_set, [n - 1])
        vf = ValueFunction(model, action_list=actions.cpu().numpy(), _device_alphas=alphas) if use_policy else None
        out = None
        for _ in range(max_generation):
            if use_policy:
                a = int(vf.evaluate_device(cur)[2][0])
            else:
                a = min(int(rng.random() * model.action_count), model.action_count - 1)
            s_p = model.transition(s, a, rng.random() if model.dense else 0.0)
            o = model.observe(s_p, a, rng.random())
            cur = _chain_updates(model, cur, 0, [(a, o)])
            out = cur if out is None else out.union(cur)
            s = s_p
            if model.end_mask[s]:
                break
        return out if out is not None else BeliefSet(model, np.zeros((0, model.state_count)))

    @classmethod
    def stage(cls, model, belief_set, alphas, actions, gamma, rng, gemm_path=0):
        n = len(belief_set)
        vf = ValueFunction(model, action_list=actions.cpu().numpy(), _device_alphas=alphas)
        v_old, g_old, _ = vf.evaluate_device(belief_set, gemm_path)
        new_alpha, best_a, new_v, _ = backup_device(model, belief_set, alphas, gamma, gemm_path)
        improved = new_v >= v_old
        cand = torch.where(improved[:, None], new_alpha, alphas[g_old.long()])
        cand_a = torch.where(improved, best_a, actions[g_old.long()])
        cross = (cand @ belief_set._b[:n].T) * belief_set._scale[:n][None, :]          # cross[i, j] = cand_i . b_j   (FP32)
        cross_h, v_old_h = cross.double().cpu().numpy(), v_old.double().cpu().numpy()
        todo = np.arange(n)
        kept = []
        v_new = np.full(n, -np.inf)
        while len(todo):
            i = int(todo[min(int(rng.random() * len(todo)), len(todo) - 1)])
            kept.append(i)
            v_new = np.maximum(v_new, cross_h[i])
            v_new[i] = max(v_new[i], v_old_h[i])            # the kept vector is at least as good as the old one at its own belief
            todo = todo[v_new[todo] < v_old_h[todo]]
        idx = torch.as_tensor(kept, dtype=torch.long, device=alphas.device)
        return cand[idx].contiguous(), cand_a[idx].contiguous()

    @classmethod
    def solve(cls, model, expansions=10, update_passes=1, max_belief_growth=10, initial_belief=None, initial_value_function=None,
              prune_level=1, prune_interval=10, limit_value_function_size=-1, gamma=0.99, eps=1e-6, convergence_stop=False,
              use_gpu=True, use_reachability=True, rng=None, history_tracking_level=1, print_progress=True, print_stats=True,
              use_policy_to_choose_actions=False, gemm_path=0, **kw):
        rng = np.random.default_rng() if rng is None else rng
        device = current_device()
        dev = torch.device('cuda', device)
        # Under torchrun (one process per GPU, process group initialised) the backups of this solve are sharded over the ranks.
        # The expand step draws on the host: every rank must hold the same generator state (same seed), so that all ranks select the
        # same belief points — checked once per solve through the first draw.
        from . import parallel
        sharded = parallel.world()[1] > 1 and os.environ.get('OLFNAV_SHARDED_TRAIN', '1') != '0'
        if sharded:
            parallel.assert_same_rng(rng, dev)
        hist = SolverHistory(model, history_tracking_level, gamma, eps, cls.name)
        t_start = time.perf_counter()
        if initial_belief is None:
            belief_set = BeliefSet(model, 1)
        elif isinstance(initial_belief, Belief):
            belief_set = BeliefSet(model, initial_belief.values[None, :])
        else:
            belief_set = initial_belief
        vf0 = ValueFunction(model, model.expected_rewards_table.T.copy(), model.actions) if initial_value_function is None else initial_value_function
        alphas = vf0.device_alphas(device).clone()
        actions = torch.as_tensor(vf0.actions, dtype=torch.int32, device=dev)

        def values_of(bs, al, ac):
            return ValueFunction(model, action_list=ac.cpu().numpy(), _device_alphas=al).evaluate_device(bs, gemm_path)[0]
        old_values = values_of(belief_set, alphas, actions)
        backup_count = 0
        for _ in range(expansions):
            t0 = time.perf_counter()
            belief_set = belief_set.union(cls.walk(model, belief_set, alphas, actions, max_belief_growth, rng, bool(use_policy_to_choose_actions)))
            torch.cuda.synchronize(dev)
            hist.expansion_times.append(time.perf_counter() - t0)
            hist.beliefs_counts.append(len(belief_set))
            t0 = time.perf_counter()
            for _ in range(update_passes):
                alphas, actions = cls.stage(model, belief_set, alphas, actions, gamma, rng, gemm_path)
                backup_count += 1
            torch.cuda.synchronize(dev)
            hist.backup_times.append(time.perf_counter() - t0)
            t0 = time.perf_counter()
            before = alphas.shape[0]
            if prune_interval > 0 and backup_count % prune_interval == 0:
                alphas, actions = prune_alphas(alphas, actions, prune_level)
            hist.pruning_times.append(time.perf_counter() - t0)
            hist.prune_counts.append(before - alphas.shape[0])
            hist.alpha_vector_counts.append(int(alphas.shape[0]))
            new_values = values_of(belief_set, alphas, actions)
            change = float((new_values[:len(old_values)] - old_values).abs().max()) if len(old_values) else float('inf')
            old_values = new_values
            hist.value_function_changes.append(change)
            if convergence_stop and change < eps:
                break
        value_function = ValueFunction(model, action_list=actions.cpu().numpy(), _device_alphas=alphas.contiguous())
        hist.total_time = time.perf_counter() - t_start
        hist.final_belief_set = belief_set
        return value_function, hist


class HSVI(_PBVI):
    '''
    Heuristic Search Value Iteration (agents/hsvi_agent.py:100-102,184-208; Smith & Simmons 2004): see the oracle's HSVI for
    the definition (upper bound = the MDP policy's Q rows, lower bound = the alpha vectors, one deterministic descent per
    expansion).  Per depth: the O successors of the current belief under a* are one batched belief-step launch, both bounds
    are evaluated on them with the device evaluate.
    '''
    name = 'HSVI'
    full_backup = False
    dominance_prune = True

    @classmethod
    def prepare(cls, model, gamma, eps, print_progress, mdp_policy=None, vi_horizon=1000, epsilon=0.99, use_reachability=True, **kw):
        if mdp_policy is None:
            mdp_policy, _ = VI.solve(model, horizon=vi_horizon, gamma=gamma, eps=eps, print_progress=print_progress)
        return {'mdp_policy': mdp_policy, 'epsilon': float(epsilon), 'gamma': gamma}

    @classmethod
    def expand(cls, model, belief_set, alphas, actions, max_generation, rng, mdp_policy=None, epsilon=0.99, gamma=0.99, **kw) -> BeliefSet:
        O = model.observation_count
        lower = ValueFunction(model, action_list=actions.cpu().numpy(), _device_alphas=alphas)
        cur = _take_rows(model, belief_set, [0])
        out = None
        for depth in range(max_generation):
            a = int(mdp_policy.evaluate_device(cur)[2][0])                         # first argmax_a b.Q_a
            succ, ok, p = _batched_updates(model, cur, [0] * O, [a] * O, list(range(O)))
            up = mdp_policy.evaluate_device(succ)[0].double()
            lo = lower.evaluate_device(succ)[0].double()
            excess = torch.where(ok, p.double() * ((up - lo) - epsilon * gamma ** (-(depth + 1))), torch.full_like(up, -np.inf)).cpu().numpy()
            o = int(np.argmax(excess))
            if not (excess[o] > 0):
                break
            cur = _take_rows(model, succ, [o])
            out = cur if out is None else out.union(cur)
        return out if out is not None else BeliefSet(model, np.zeros((0, model.state_count)))
