'''
Device-backed mirror of the `pomdp_toolkit` surface the reference's agents use:
POMDP, Belief, BeliefSet, ValueFunction.  Host-visible attributes keep the reference's layouts
(float64 / int64 NumPy, state-major); the arithmetic runs in the CUDA library through the C ABI
(include/olfnav_b200.h).  Nothing here imports oracle/.

Two model families: grid models with one deterministic successor per (state, action) — what `exact_converter` builds
(reference agents/model_based_util/environment_converter.py:7-134; stencil kernels) — and dense stochastic-transition
models over a flat state list — what `minimal_converter` builds (:137-292; dense kernels, csrc/dense.cu).
'''
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

from . import _lib
from ._lib import lib, ptr, stream, check


def current_device() -> int:
    _lib.require_cuda()
    return torch.cuda.current_device()


class POMDP:
    '''
    POMDP(states, actions, observations, reachable_states_table, observation_table, end_states,
          start_probabilities) as called at environment_converter.py:125-133, plus the grid
    description the kernels need: `grid_shape`, `moves` (A x 2, (dy, dx)) and `boundary`.
    When those are not given they are recovered from the reachable-states table.
    '''
    def __init__(self, states, actions, observations,
                 reachable_states_table: np.ndarray = None,
                 transition_table: np.ndarray = None,
                 observation_table: np.ndarray = None,
                 end_states=None,
                 start_probabilities: np.ndarray = None,
                 grid_shape: tuple = None, moves: np.ndarray = None, boundary: str = None) -> None:
        arr = np.array(states, dtype=object) if not isinstance(states, (int, np.integer)) else None
        self.dense = transition_table is not None          # dense stochastic transitions (minimal_converter): flat state list
        if arr is not None and arr.ndim == 2:
            gshape = arr.shape
        elif grid_shape is not None:
            gshape = tuple(grid_shape)
        elif self.dense:
            gshape = (1, int(states) if arr is None else len(arr))
        else:
            raise NotImplementedError('a flat state list needs a transition_table (dense model); grid models need a 2-D state grid')
        if grid_shape is not None:
            gshape = tuple(int(v) for v in grid_shape)
        self.state_count = int(gshape[0] * gshape[1])
        self.state_grid = np.arange(self.state_count).reshape(gshape)
        self.states = np.arange(self.state_count)
        self.state_labels = [str(s) for s in arr.ravel()] if arr is not None else [f's_{i}' for i in range(self.state_count)]

        self.action_labels = [f'a_{i}' for i in range(actions)] if isinstance(actions, (int, np.integer)) else [str(a) for a in actions]
        self.action_count = len(self.action_labels)
        self.actions = np.arange(self.action_count)
        self.observation_labels = [f'o_{i}' for i in range(observations)] if isinstance(observations, (int, np.integer)) else [str(o) for o in observations]
        self.observation_count = len(self.observation_labels)
        self.observations = np.arange(self.observation_count)
        S, A, O = self.state_count, self.action_count, self.observation_count

        if self.dense:
            # POMDP(..., transition_table = T[s, a, s']) as called at environment_converter.py:282-290: every state is "reachable",
            # with probability T (so sampling walks the cumulative sum of the row, like the oracle's)
            self._transition_table = np.ascontiguousarray(transition_table, dtype=np.float64)
            assert self._transition_table.shape == (S, A, S), 'transition_table must be (S, A, S)'
            self.reachable_states = np.broadcast_to(np.arange(S)[None, None, :], (S, A, S))
            self.reachable_state_count = S
            self.reachable_probabilities = self._transition_table
            self.observation_table = np.ascontiguousarray(observation_table, dtype=np.float64)
            assert self.observation_table.shape == (S, A, O)
            self.end_states = np.asarray([] if end_states is None else end_states, dtype=np.int64)
            self.end_mask = np.zeros(S, dtype=bool)
            self.end_mask[self.end_states] = True
            self.expected_rewards_table = np.sum(self._transition_table * self.end_mask[None, None, :], axis=2)
            self.start_probabilities = (np.full(S, 1.0 / S) if start_probabilities is None
                                        else np.ascontiguousarray(start_probabilities, dtype=np.float64).ravel())
            self.moves, self.boundary = None, None
            self.is_on_gpu = True
            self._handles = {}
            return
        if reachable_states_table is None:
            raise NotImplementedError('a model needs a reachable_states_table (grid) or a transition_table (dense)')
        R = np.asarray(reachable_states_table, dtype=np.int64)
        if R.ndim == 2:
            R = R[:, :, None]
        if R.shape != (S, A, 1):
            raise NotImplementedError('only one reachable state per (state, action) is supported')
        self.reachable_states = R
        self.reachable_state_count = 1
        self.reachable_probabilities = np.ones((S, A, 1))

        self.observation_table = np.ascontiguousarray(observation_table, dtype=np.float64)
        assert self.observation_table.shape == (S, A, O)
        self.end_states = np.asarray([] if end_states is None else end_states, dtype=np.int64)
        self.end_mask = np.zeros(S, dtype=bool)
        self.end_mask[self.end_states] = True
        self.expected_rewards_table = self.end_mask[R[:, :, 0]].astype(float)
        self.start_probabilities = (np.full(S, 1.0 / S) if start_probabilities is None
                                    else np.ascontiguousarray(start_probabilities, dtype=np.float64).ravel())

        # grid description
        H, W = gshape
        if moves is None or boundary is None:
            moves, boundary = _infer_grid(R[:, :, 0], H, W)
        self.moves = np.ascontiguousarray(moves, dtype=np.int32).reshape(A, 2)
        self.boundary = boundary
        # row padding rule of the device layout (olfnav_padded_width_for): 0 = by width (64 floats from 128 columns up, else 4);
        # OLFNAV_ROW_PAD=4|64 forces one for the models built while it is set
        self.row_pad = int(os.environ.get('OLFNAV_ROW_PAD', '0') or 0)
        expected = _move_table(H, W, self.moves, boundary)
        if not np.array_equal(expected, R[:, :, 0]):
            raise NotImplementedError('reachable_states_table is not a unit-step grid move table')

        self.is_on_gpu = True
        self._handles = {}

    # toolkit attributes used at agents/infotaxis_agent.py:272 (small models only; host side)
    @property
    def reachable_transition_observation_table(self) -> np.ndarray:
        a_idx = np.arange(self.action_count)[None, :, None]
        return self.observation_table[self.reachable_states, a_idx, :]

    @property
    def on_gpu(self) -> 'POMDP':
        return self

    @property
    def on_cpu(self) -> 'POMDP':
        return self

    @property
    def grid_shape(self) -> tuple:
        return self.state_grid.shape

    @property
    def padded_width(self) -> int:
        H, W = self.grid_shape
        return int(lib.olfnav_padded_width_for(W, 0 if self.dense else getattr(self, 'row_pad', 0))) if not self.dense else (W + 3) // 4 * 4

    @property
    def padded_states(self) -> int:
        return self.grid_shape[0] * self.padded_width

    def handle(self, device: int = None):
        device = current_device() if device is None else device
        if device not in self._handles:
            h = ctypes.c_void_p()
            H, W = self.grid_shape
            end = np.ascontiguousarray(self.end_mask.astype(np.uint8))
            if self.dense:
                check(lib.olfnav_model_create_dense(ctypes.byref(h), device, self.state_count, self.action_count, self.observation_count,
                                                    ptr(self._transition_table), ptr(self.observation_table), ptr(end),
                                                    ptr(self.start_probabilities)))
                self._handles[device] = _Handle(h, lib.olfnav_model_destroy)
                return self._handles[device].h
            check(lib.olfnav_model_create_padded(ctypes.byref(h), device, H, W, self.action_count, self.observation_count,
                                                 ptr(self.moves), _lib.BOUNDARY[self.boundary], ptr(self.observation_table),
                                                 ptr(end), ptr(self.start_probabilities), getattr(self, 'row_pad', 0)))
            self._handles[device] = _Handle(h, lib.olfnav_model_destroy)
        return self._handles[device].h

    def __getstate__(self):
        state = self.__dict__.copy()
        state['_handles'] = {}
        return state

    # sampling helpers (host side, inverse CDF on a uniform draw; same convention as the oracle's)
    @property
    def transition_table(self) -> np.ndarray:
        if self.dense:
            return self._transition_table
        tt = np.zeros((self.state_count, self.action_count, self.state_count))
        tt[np.arange(self.state_count)[:, None], np.arange(self.action_count)[None, :], self.reachable_states[:, :, 0]] = 1.0
        return tt

    def transition(self, s: int, a: int, u: float = 0.0) -> int:
        "s' ~ T(.|s, a) by inverse CDF from the uniform draw u (deterministic on grid models)."
        if not self.dense:
            return int(self.reachable_states[s, a, 0])
        cdf = np.cumsum(self._transition_table[s, a])
        return min(int(np.searchsorted(cdf, u * cdf[-1], side='right')), self.state_count - 1)

    def observe(self, s_p: int, a: int, u: float) -> int:
        cdf = np.cumsum(self.observation_table[s_p, a])
        return min(int(np.searchsorted(cdf, u * cdf[-1], side='right')), self.observation_count - 1)


class _Handle:
    def __init__(self, h, destroy):
        self.h, self._destroy = h, destroy

    def __del__(self):
        try:
            self._destroy(self.h)
        except Exception:
            pass


def _move_table(H: int, W: int, moves: np.ndarray, boundary: str) -> np.ndarray:
    'R[s, a] for unit-step moves under the boundary condition (reference environment.py:711-769).'
    wrap_y = boundary in ('wrap', 'wrap_vertical')
    wrap_x = boundary in ('wrap', 'wrap_horizontal')
    y, x = np.divmod(np.arange(H * W), W)
    out = np.empty((H * W, len(moves)), dtype=np.int64)
    for a, (dy, dx) in enumerate(moves):
        y2 = (y + dy) % H if wrap_y else np.clip(y + dy, 0, H - 1)
        x2 = (x + dx) % W if wrap_x else np.clip(x + dx, 0, W - 1)
        out[:, a] = y2 * W + x2
    return out


def _infer_grid(R: np.ndarray, H: int, W: int):
    'Recover (moves, boundary) from a reachable-states table built by exact_converter.'
    S, A = R.shape
    probe = (H // 2) * W + (W // 2) if H > 2 and W > 2 else None
    if probe is None:
        raise NotImplementedError('grid too small to infer the move set; pass moves= and boundary=')
    dest = R[probe]
    moves = np.stack([dest // W - probe // W, dest % W - probe % W], axis=1)
    for boundary in ('stop', 'wrap', 'wrap_vertical', 'wrap_horizontal'):
        if np.array_equal(_move_table(H, W, moves, boundary), R):
            return moves, boundary
    raise NotImplementedError('reachable_states_table does not describe unit-step grid moves')


# =================================================================================================
class BeliefSet:
    '''
    (N, S) set of beliefs resident in HBM in the padded, unnormalised layout (see
    include/olfnav_b200.h).  Mirrors the toolkit's BeliefSet as used at
    agents/pbvi_agent.py:717-726,783-790,811 and agents/infotaxis_agent.py:232,257-270,335-363.
    `beliefs` may be an (N, S) NumPy array / list of Belief (uploaded) or an int N (N copies of
    the start probabilities, built on the device without an O(N) Python loop).
    '''
    def __init__(self, model: POMDP, beliefs=None, *, _buffers=None, device: int = None) -> None:
        self.model = model
        self.is_on_gpu = True
        self.device = current_device() if device is None else device
        dev = torch.device('cuda', self.device)
        Sp = model.padded_states
        if _buffers is not None:
            self._b, self._scale, self._n = _buffers
            return
        if isinstance(beliefs, (int, np.integer)):
            n = int(beliefs)
            self._b = torch.empty((max(n, 1), Sp), dtype=torch.float32, device=dev)
            self._scale = torch.empty(max(n, 1), dtype=torch.float32, device=dev)
            self._n = n
            with torch.cuda.device(self.device):
                check(lib.olfnav_belief_init(model.handle(self.device), ptr(self._b), Sp, n, ptr(self._scale), stream()))
            return
        if isinstance(beliefs, np.ndarray):
            dense = np.ascontiguousarray(beliefs, dtype=np.float64).reshape(-1, model.state_count)
        else:
            beliefs = list(beliefs)
            dense = np.array([np.asarray(b.values) for b in beliefs], dtype=np.float64).reshape(-1, model.state_count)
        n = len(dense)
        self._b = torch.empty((max(n, 1), Sp), dtype=torch.float32, device=dev)
        self._scale = torch.ones(max(n, 1), dtype=torch.float32, device=dev)
        self._n = n
        if n:
            d = torch.from_numpy(dense).to(dev)
            with torch.cuda.device(self.device):
                check(lib.olfnav_pack_rows(model.handle(self.device), ptr(d), None, ptr(self._b), Sp, n, stream()))
            torch.cuda.current_stream(dev).synchronize()

    def __len__(self) -> int:
        return self._n

    @property
    def ld(self) -> int:
        return self._b.stride(0)

    @property
    def on_gpu(self) -> 'BeliefSet':
        return self

    @property
    def on_cpu(self) -> 'BeliefSet':
        return self

    @property
    def belief_array(self) -> np.ndarray:
        'Normalised (N, S) float64 host copy in the reference layout.'
        return self.belief_tensor(torch.float64).cpu().numpy()

    def belief_tensor(self, dtype=torch.float32) -> torch.Tensor:
        'Normalised dense (N, S) device tensor.'
        out = torch.empty((self._n, self.model.state_count), dtype=dtype, device=self._b.device)
        if self._n:
            with torch.cuda.device(self.device):
                check(lib.olfnav_unpack_rows(self.model.handle(self.device), ptr(self._b), self.ld, self._n, ptr(self._scale),
                                             ptr(out) if dtype == torch.float64 else None,
                                             ptr(out) if dtype == torch.float32 else None, stream()))
        return out

    @property
    def belief_list(self) -> list:
        return [Belief(self.model, row) for row in self.belief_array]

    @property
    def entropies(self) -> np.ndarray:
        return self.entropies_tensor().double().cpu().numpy()

    def entropies_tensor(self) -> torch.Tensor:
        out = torch.empty(self._n, dtype=torch.float32, device=self._b.device)
        with torch.cuda.device(self.device):
            check(lib.olfnav_entropies(self.model.handle(self.device), ptr(self._b), self.ld, self._n, ptr(self._scale), ptr(out), stream()))
        return out

    def update(self, actions, observations, raise_on_impossible_belief: bool = True, use_reachability: bool = True,
               return_provenance: bool = False):
        '''
        Toolkit-compatible out-of-place update: rows whose normaliser is 0 are dropped (or raise);
        provenance (n_ok, 3) = (source row, action, observation).
        '''
        dev = self._b.device
        act = torch.as_tensor(np.asarray(actions), dtype=torch.int32, device=dev).contiguous()
        obs = torch.as_tensor(np.asarray(observations), dtype=torch.int32, device=dev).contiguous()
        assert act.shape == (self._n,) and obs.shape == (self._n,)
        out_b = torch.empty_like(self._b)
        out_s = torch.empty_like(self._scale)
        ok = torch.empty(max(self._n, 1), dtype=torch.uint8, device=dev)
        with torch.cuda.device(self.device):
            check(lib.olfnav_belief_step(self.model.handle(self.device), ptr(self._b), ptr(out_b), self.ld, self._n,
                                         ptr(act), ptr(obs), None, None, ptr(self._scale), ptr(out_s), ptr(ok), stream()))
        ok_b = ok[:self._n].bool()
        if bool(ok_b.all()):
            new = BeliefSet(self.model, _buffers=(out_b, out_s, self._n), device=self.device)
            rows = torch.arange(self._n, device=dev)
        else:
            if raise_on_impossible_belief:
                raise ValueError(f'Impossible belief update for rows {torch.nonzero(~ok_b).flatten().tolist()}')
            rows = torch.nonzero(ok_b).flatten()
            n_ok = len(rows)
            cb = torch.empty((max(n_ok, 1), self._b.shape[1]), dtype=torch.float32, device=dev)
            cs = torch.empty(max(n_ok, 1), dtype=torch.float32, device=dev)
            if n_ok:
                r32 = rows.to(torch.int32).contiguous()
                with torch.cuda.device(self.device):
                    check(lib.olfnav_copy_rows(ptr(out_b), ptr(cb), self.ld, self.model.padded_states, ptr(r32), ptr(out_s), ptr(cs), n_ok, stream()))
            new = BeliefSet(self.model, _buffers=(cb, cs, n_ok), device=self.device)
        if return_provenance:
            prov = torch.stack([rows.long(), act.long()[rows], obs.long()[rows]], dim=1).cpu().numpy()
            return new, prov
        return new

    def update_and_evaluate(self, value_function: 'ValueFunction', actions, observations, src_rows=None):
        '''
        BeliefSet.update followed by ValueFunction.evaluate_at of the updated set (agents/pbvi_agent.py:783-787 then :741) as
        ONE device pass (olfnav_belief_step_evaluate).  Every row is kept (failed rows: ok False, scale 0).
        The kernel writes the updated rows grouped by action; this convenience wrapper gathers them back into input order
        (run_test keeps them where they are and carries the row map instead).
        Returns (new BeliefSet, ok bool tensor, values f32, alpha index i32, action i32) — device tensors.
        '''
        dev = self._b.device
        act = torch.as_tensor(np.asarray(actions), dtype=torch.int32, device=dev).contiguous()
        obs = torch.as_tensor(np.asarray(observations), dtype=torch.int32, device=dev).contiguous()
        n = int(act.shape[0])
        rows = None if src_rows is None else torch.as_tensor(np.asarray(src_rows), dtype=torch.int32, device=dev).contiguous()
        assert obs.shape == (n,) and (rows is not None or n == self._n)
        h = self.model.handle(self.device)
        if lib.olfnav_can_fuse_policy(h, value_function.handle(self.device)) != 1:
            # outside what the one-pass kernel covers (more than 80 alpha vectors, clipped vertical / wrapped horizontal moves,
            # narrow unpadded grids, dense models): the two calls it stands for
            out_b = torch.empty((max(n, 1), self._b.shape[1]), dtype=torch.float32, device=dev)
            out_s = torch.empty(max(n, 1), dtype=torch.float32, device=dev)
            ok = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
            with torch.cuda.device(self.device):
                check(lib.olfnav_belief_step(h, ptr(self._b), ptr(out_b), self.ld, n, ptr(act), ptr(obs), ptr(rows), None, ptr(self._scale),
                                             ptr(out_s), ptr(ok), stream()))
            new = BeliefSet(self.model, _buffers=(out_b, out_s, n), device=self.device)
            val, arg, action = value_function.evaluate_device(new)
            return new, ok[:n].bool(), val, arg, action
        cap = max(int(lib.olfnav_fused_out_rows(h, n)), 1)
        out_b = torch.empty((cap, self._b.shape[1]), dtype=torch.float32, device=dev)
        out_s = torch.zeros(cap, dtype=torch.float32, device=dev)
        ok = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
        val = torch.empty(max(n, 1), dtype=torch.float32, device=dev)
        arg = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        action = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        row_out = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
        with torch.cuda.device(self.device):
            check(lib.olfnav_belief_step_evaluate(h, value_function.handle(self.device), ptr(self._b), self._b.shape[0], ptr(out_b), cap,
                                                  self.ld, n, ptr(act), ptr(obs), ptr(rows), ptr(self._scale), ptr(out_s), ptr(ok),
                                                  ptr(val), ptr(arg), ptr(action), ptr(row_out), stream()))
        idx = row_out[:n].long()
        new = BeliefSet(self.model, _buffers=(out_b[idx].contiguous() if n else out_b[:1], out_s[idx].contiguous() if n else out_s[:1], n), device=self.device)
        return new, ok[:n].bool(), val[:n], arg[:n], action[:n]

    def union(self, other: 'BeliefSet') -> 'BeliefSet':
        b = torch.cat([self._b[:self._n], other._b[:other._n]], dim=0)
        s = torch.cat([self._scale[:self._n], other._scale[:other._n]], dim=0)
        return BeliefSet(self.model, _buffers=(b, s, self._n + other._n), device=self.device)

    def generate_all_successors(self, use_reachability: bool = True, raise_on_impossible_belief: bool = False,
                                return_provenance: bool = False, max_bytes: int = 8 << 30):
        '''
        Every successor b_{a,o} of every belief of the set (toolkit surface used at agents/infotaxis_agent.py:260-262): ONE
        olfnav_belief_step launch over the (belief, action, observation) triples, impossible successors (normaliser 0) dropped.
        Returns the BeliefSet (rows ordered by belief, then action, then observation) and, with return_provenance, the (k, 3) int
        array of (source belief, action, observation) of the rows kept.
        This MATERIALISES n * A * O rows of S floats and exists for API fidelity on small sets (the reference's own
        Infotaxis_Agent can then run on this toolkit surface); Infotaxis_Agent.choose_action of this package never does
        (olfnav_infotaxis_choose).  Sets whose successors exceed `max_bytes` are refused.
        '''
        m = self.model
        n, A, O = self._n, m.action_count, m.observation_count
        total = n * A * O
        need = total * m.padded_states * 4
        if need > max_bytes:
            raise MemoryError(f'generate_all_successors would materialise {total} rows ({need / 2**30:.1f} GiB): use '
                              'Infotaxis_Agent.choose_action / olfnav_infotaxis_choose, which never builds them')
        dev = self._b.device
        idx = torch.arange(total, device=dev)
        rows = (idx // (A * O)).to(torch.int32).contiguous()
        act = ((idx // O) % A).to(torch.int32).contiguous()
        obs = (idx % O).to(torch.int32).contiguous()
        out_b = torch.empty((max(total, 1), self._b.shape[1]), dtype=torch.float32, device=dev)
        out_s = torch.empty(max(total, 1), dtype=torch.float32, device=dev)
        ok = torch.zeros(max(total, 1), dtype=torch.uint8, device=dev)
        if total:
            with torch.cuda.device(self.device):
                check(lib.olfnav_belief_step(m.handle(self.device), ptr(self._b), ptr(out_b), self.ld, total, ptr(act), ptr(obs), ptr(rows), None,
                                             ptr(self._scale), ptr(out_s), ptr(ok), stream()))
        keep = torch.nonzero(ok[:total].bool()).flatten()
        if raise_on_impossible_belief and int(keep.numel()) < total:
            raise ValueError('Impossible belief update (the observation has probability 0 under the belief and action)')
        new = BeliefSet(m, _buffers=(out_b[keep].contiguous() if int(keep.numel()) else out_b[:1], out_s[keep].contiguous() if int(keep.numel()) else out_s[:1],
                                     int(keep.numel())), device=self.device)
        if return_provenance:
            prov = torch.stack([rows.long()[keep], act.long()[keep], obs.long()[keep]], dim=1).cpu().numpy()
            return new, prov
        return new


class Belief:
    'One belief point (host values, float64); Belief(model) = start probabilities (agents/pbvi_agent.py:717,838).'
    def __init__(self, model: POMDP, values: np.ndarray = None) -> None:
        self.model = model
        self.values = model.start_probabilities.copy() if values is None else np.asarray(values, dtype=float)
        assert self.values.shape == (model.state_count,)
        self.is_on_gpu = True

    def update(self, a: int, o: int, use_reachability: bool = True, throw_error: bool = True) -> 'Belief':
        'a1 for N = 1 (agents/pbvi_agent.py:869); raises on an impossible update.'
        new = BeliefSet(self.model, self.values[None, :]).update([a], [o], raise_on_impossible_belief=throw_error)
        if len(new) == 0:
            return Belief(self.model, np.zeros_like(self.values))
        return Belief(self.model, new.belief_array[0])

    @property
    def on_gpu(self) -> 'Belief':
        return self

    @property
    def on_cpu(self) -> 'Belief':
        return self


# =================================================================================================
class ValueFunction:
    '''
    Alpha-vector set (G, S) + actions (G,), mirrored on the device in the padded layout with
    split-precision operand planes.  Call sites: agents/pbvi_agent.py:523,574-577,690-692,741.
    On-disk layout (the toolkit's is unknown): one (G, S+1) float64 .npy, column 0 = action.
    '''
    def __init__(self, model: POMDP, alpha_vectors=None, action_list=None, *, _device_alphas: torch.Tensor = None) -> None:
        self.model = model
        self.is_on_gpu = True
        self._host = None
        self._dev = None          # (G, Sp) float32 padded, device
        self._handles = {}
        if _device_alphas is not None:
            self._dev = _device_alphas
            self._actions = np.asarray(action_list, dtype=np.int64).ravel()
            assert len(self._actions) == self._dev.shape[0]
        else:
            if alpha_vectors is None:
                alpha_vectors, action_list = np.zeros((0, model.state_count)), np.zeros(0, dtype=np.int64)
            self._host = np.ascontiguousarray(alpha_vectors, dtype=np.float64).reshape(-1, model.state_count)
            self._actions = np.asarray(action_list, dtype=np.int64).ravel()
            assert len(self._actions) == len(self._host)

    def __len__(self) -> int:
        return len(self._actions)

    @property
    def actions(self) -> np.ndarray:
        return self._actions

    @property
    def alpha_vector_array(self) -> np.ndarray:
        if self._host is None:
            G = len(self)
            out = torch.empty((G, self.model.state_count), dtype=torch.float64, device=self._dev.device)
            if G:
                dev_i = self._dev.device.index
                with torch.cuda.device(dev_i):
                    check(lib.olfnav_unpack_rows(self.model.handle(dev_i), ptr(self._dev), self._dev.stride(0), G, None, ptr(out), None, stream()))
            self._host = out.cpu().numpy()
        return self._host

    def device_alphas(self, device: int = None) -> torch.Tensor:
        '(G, Sp) float32 padded alpha vectors on the device.'
        device = current_device() if device is None else device
        if self._dev is None or self._dev.device.index != device:
            dev = torch.device('cuda', device)
            G = len(self)
            out = torch.zeros((max(G, 1), self.model.padded_states), dtype=torch.float32, device=dev)
            if G:
                d = torch.from_numpy(self.alpha_vector_array).to(dev)
                with torch.cuda.device(device):
                    check(lib.olfnav_pack_rows(self.model.handle(device), ptr(d), None, ptr(out), out.stride(0), G, stream()))
                torch.cuda.current_stream(dev).synchronize()
            self._dev = out[:G] if G else out[:0]
        return self._dev

    def handle(self, device: int = None):
        device = current_device() if device is None else device
        if device not in self._handles:
            assert len(self) > 0, 'empty value function'
            alphas = self.device_alphas(device)
            acts = torch.as_tensor(self._actions, dtype=torch.int32, device=alphas.device).contiguous()
            h = ctypes.c_void_p()
            with torch.cuda.device(device):
                check(lib.olfnav_alphas_create(ctypes.byref(h), self.model.handle(device), ptr(alphas), alphas.stride(0),
                                               ptr(acts), len(self), stream()))
            torch.cuda.current_stream(alphas.device).synchronize()
            self._handles[device] = _Handle(h, lib.olfnav_alphas_destroy)
        return self._handles[device].h

    def evaluate_device(self, belief_set: BeliefSet, path: int = 0):
        'Device tensors (values f32, alpha index i32, action i32) for every belief of the set.'
        n, dev = len(belief_set), belief_set._b.device
        val = torch.empty(max(n, 1), dtype=torch.float32, device=dev)
        arg = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        act = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        with torch.cuda.device(belief_set.device):
            check(lib.olfnav_evaluate(self.model.handle(belief_set.device), self.handle(belief_set.device), ptr(belief_set._b),
                                      belief_set.ld, n, ptr(belief_set._scale), ptr(val), ptr(arg), ptr(act), path, stream()))
        return val[:n], arg[:n], act[:n]

    def evaluate_at(self, belief, path: int = 0):
        'V(b) = max_g b.alpha_g, action of the first argmax (agents/pbvi_agent.py:741).'
        if isinstance(belief, Belief):
            v, _, a = self.evaluate_device(BeliefSet(self.model, belief.values[None, :]), path)
            return float(v[0]), int(a[0])
        v, _, a = self.evaluate_device(belief, path)
        return v.double().cpu().numpy(), a.long().cpu().numpy()

    @property
    def on_gpu(self) -> 'ValueFunction':
        return self

    @property
    def on_cpu(self) -> 'ValueFunction':
        return self

    def save(self, folder: str = './', file_name: str = 'Value_Function.npy') -> None:
        if not file_name.endswith('.npy'):
            file_name += '.npy'
        np.save(os.path.join(folder, file_name), np.hstack([self._actions[:, None].astype(float), self.alpha_vector_array]))

    @classmethod
    def load(cls, file: str, model: POMDP) -> 'ValueFunction':
        arr = np.load(file)
        return cls(model, alpha_vectors=arr[:, 1:], action_list=arr[:, 0].astype(np.int64))

    def __getstate__(self):
        state = self.__dict__.copy()
        state['_host'] = self.alpha_vector_array
        state['_dev'] = None
        state['_handles'] = {}
        return state
