'''
ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.

CPU (NumPy, float64) restatement of the subset of the third-party PyPI package
`pomdp-toolkit` (constraint `>=0.0.4`, reference `requirements.txt:7`, no lock file) that the
reference's hot path calls.  The real package is NOT present in /root/reference nor installable
here (no network), and the reference ships no test, golden vector or fixture for it, so this
restatement is anchored on:

  (i)   the reference's call sites (cited per symbol below, paths relative to /root/reference),
  (ii)  the model the reference builds in
        olfactory_navigation/agents/model_based_util/environment_converter.py:7-134,
  (iii) the published algorithms the reference's docstrings cite (Pineau et al. PBVI
        `agents/pbvi_agent.py:239`, Shani et al. FSVI `agents/fsvi_agent.py:12`,
        Vergassola et al. infotaxis `agents/infotaxis_agent.py:24`),
  (iv)  the solver logs preserved in the reference notebooks (VI "Converged in 918 iterations"
        for gamma=.99, eps=1e-6: experiments/pomdp_agent_training.ipynb:115-116; model
        construction log experiments/q_mdp_agent_training.ipynb:60-80).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this package; the product (olfactory-navigation_b200/) never does.

Semantics DEFINED here because upstream is unseen (see DESIGN.md "Oracle decisions"):
  * impossible update (normaliser == 0): row dropped, provenance tells which rows survived;
  * ties: first maximum wins (np.argmax) everywhere;
  * entropy: natural log, 0*log(0) := 0;
  * rewards: 1 for every transition that lands in an end state (non-absorbing, which is what
    makes VI take 918 sweeps on every model size in the notebooks);
  * provenance: int array (n, 3) with columns (source belief row, action, observation).
'''
from __future__ import annotations

import os
import numpy as np

__version__ = '0.0.4+oracle'

__all__ = ['POMDP', 'Belief', 'BeliefSet', 'ValueFunction', 'solvers']


class POMDP:
    '''
    Model container.  Constructed by the reference at
    agents/model_based_util/environment_converter.py:125-133 (exact) and :282-290 (minimal) as
    POMDP(states, actions, observations, reachable_states_table|transition_table,
          observation_table, end_states, start_probabilities).

    Tables (float64 / int64 like the reference):
      reachable_states              (S, A, R) int   successor ids
      reachable_probabilities       (S, A, R)       T(s, a, r); 1/R when no transition table
                                                    (log: experiments/q_mdp_agent_training.ipynb:61-70)
      observation_table             (S, A, O)       P(o | arrive in s', a)
      expected_rewards_table        (S, A)          sum_r T(s,a,r) * 1[R[s,a,r] in end_states]
      reachable_transition_observation_table (S, A, R, O) = T(s,a,r) * O[R[s,a,r], a, o]
                                                    (indexed `[:, a_vec, :, o_vec]` at
                                                    agents/infotaxis_agent.py:272)
    '''
    def __init__(self,
                 states,
                 actions,
                 observations,
                 reachable_states_table: np.ndarray = None,
                 transition_table: np.ndarray = None,
                 observation_table: np.ndarray = None,
                 end_states=None,
                 start_probabilities: np.ndarray = None,
                 ) -> None:
        # States: an int count, a flat list of labels or a 2-D grid of labels
        # (environment_converter.py:43 passes a list of rows of labels)
        if isinstance(states, (int, np.integer)):
            self.state_count = int(states)
            self.state_labels = [f's_{i}' for i in range(self.state_count)]
            self.state_grid = np.arange(self.state_count)[None, :]
        else:
            arr = np.array(states, dtype=object)
            self.state_count = int(arr.size)
            self.state_labels = [str(s) for s in arr.ravel()]
            self.state_grid = np.arange(self.state_count).reshape(arr.shape if arr.ndim > 1 else (1, -1))
        self.states = np.arange(self.state_count)

        self.action_labels = [f'a_{i}' for i in range(actions)] if isinstance(actions, (int, np.integer)) else [str(a) for a in actions]
        self.action_count = len(self.action_labels)
        self.actions = np.arange(self.action_count)

        self.observation_labels = [f'o_{i}' for i in range(observations)] if isinstance(observations, (int, np.integer)) else [str(o) for o in observations]
        self.observation_count = len(self.observation_labels)
        self.observations = np.arange(self.observation_count)

        S, A, O = self.state_count, self.action_count, self.observation_count

        # Transitions
        if reachable_states_table is not None:
            self.reachable_states = np.asarray(reachable_states_table, dtype=np.int64)
            if self.reachable_states.ndim == 2:
                self.reachable_states = self.reachable_states[:, :, None]
            assert self.reachable_states.shape[:2] == (S, A)
            R = self.reachable_states.shape[2]
            if transition_table is not None:
                tt = np.asarray(transition_table, dtype=float)
                self.reachable_probabilities = np.take_along_axis(tt, self.reachable_states, axis=2)
            else:
                self.reachable_probabilities = np.full((S, A, R), 1.0 / R)
        else:
            assert transition_table is not None, 'Either reachable_states_table or transition_table is required'
            tt = np.asarray(transition_table, dtype=float)
            assert tt.shape == (S, A, S)
            R = int(np.max(np.sum(tt > 0, axis=2)))
            order = np.argsort(-(tt > 0).astype(np.int8), axis=2, kind='stable')[:, :, :R]
            self.reachable_states = order.astype(np.int64)
            self.reachable_probabilities = np.take_along_axis(tt, self.reachable_states, axis=2)
        self.reachable_state_count = self.reachable_states.shape[2]
        self._transition_table = None if transition_table is None else np.asarray(transition_table, dtype=float)

        # Observations
        self.observation_table = np.asarray(observation_table, dtype=float)
        assert self.observation_table.shape == (S, A, O)

        # End states and rewards ("reaching an end state ... will give a reward of 1",
        # experiments/q_mdp_agent_training.ipynb:76)
        self.end_states = np.asarray([] if end_states is None else end_states, dtype=np.int64)
        end_mask = np.zeros(S, dtype=bool)
        end_mask[self.end_states] = True
        self.end_mask = end_mask
        self.expected_rewards_table = np.sum(self.reachable_probabilities * end_mask[self.reachable_states], axis=2)

        # Start probabilities
        if start_probabilities is None:
            self.start_probabilities = np.full(S, 1.0 / S)
        else:
            self.start_probabilities = np.asarray(start_probabilities, dtype=float).ravel()
            assert self.start_probabilities.shape == (S,)

        # T * O on reachable states
        a_idx = np.arange(A)[None, :, None]
        self.reachable_transition_observation_table = (
            self.reachable_probabilities[:, :, :, None] * self.observation_table[self.reachable_states, a_idx, :])

        self.is_on_gpu = False

    # -- dense views (small models only) -------------------------------------------------------
    @property
    def transition_table(self) -> np.ndarray:
        if self._transition_table is None:
            S, A = self.state_count, self.action_count
            tt = np.zeros((S, A, S))
            s_idx = np.arange(S)[:, None, None]
            a_idx = np.arange(A)[None, :, None]
            np.add.at(tt, (s_idx, a_idx, self.reachable_states), self.reachable_probabilities)
            self._transition_table = tt
        return self._transition_table

    @property
    def transition_observation_table(self) -> np.ndarray:
        # (S, A, S', O), used at agents/infotaxis_agent.py:274 (dense: small models only; cached)
        if getattr(self, '_dense_to_table', None) is None:
            self._dense_to_table = self.transition_table[:, :, :, None] * np.transpose(self.observation_table, (1, 0, 2))[None, :, :, :]
        return self._dense_to_table

    # -- twin objects (agents/pbvi_agent.py:396-397,433-434); the oracle is CPU-only --------------
    @property
    def on_gpu(self) -> 'POMDP':
        return self

    @property
    def on_cpu(self) -> 'POMDP':
        return self

    # -- sampling helpers used by the solvers' expand functions --------------------------------
    def transition(self, s: int, a: int, u: float) -> int:
        'Sample s\' ~ T(.|s,a) by inverse CDF from the uniform draw u.'
        if self.reachable_state_count == 1:
            return int(self.reachable_states[s, a, 0])
        cdf = np.cumsum(self.reachable_probabilities[s, a])
        r = min(int(np.searchsorted(cdf, u * cdf[-1], side='right')), self.reachable_state_count - 1)
        return int(self.reachable_states[s, a, r])

    def observe(self, s_p: int, a: int, u: float) -> int:
        'Sample o ~ O(.|s\',a) by inverse CDF from the uniform draw u.'
        cdf = np.cumsum(self.observation_table[s_p, a])
        return min(int(np.searchsorted(cdf, u * cdf[-1], side='right')), self.observation_count - 1)


def _push(model: POMDP, belief_rows: np.ndarray, actions: np.ndarray) -> np.ndarray:
    '''
    b_a(s') = sum_{s,r : R[s,a,r] = s'} T(s,a,r) * b(s)   for every row (SURVEY App. B "push").
    belief_rows (n, S), actions (n,) -> (n, S).
    '''
    n, S = belief_rows.shape
    out = np.zeros((n, S))
    rows = np.arange(n)[:, None]
    for r in range(model.reachable_state_count):
        dest = model.reachable_states[:, actions, r].T            # (n, S)
        w = model.reachable_probabilities[:, actions, r].T * belief_rows
        np.add.at(out, (rows, dest), w)
    return out


class Belief:
    '''
    One belief point.  Call sites: agents/pbvi_agent.py:717 (`Belief(model)` -> start
    probabilities), :838, :869 (`belief.update(a=, o=, use_reachability=)`, wrapped in
    try/except :867-872 => raises on an impossible update).
    '''
    def __init__(self, model: POMDP, values: np.ndarray = None) -> None:
        self.model = model
        self.values = model.start_probabilities.copy() if values is None else np.asarray(values, dtype=float)
        assert self.values.shape == (model.state_count,)
        self.is_on_gpu = False

    def update(self, a: int, o: int, use_reachability: bool = True, throw_error: bool = True) -> 'Belief':
        u = _push(self.model, self.values[None, :], np.array([a]))[0] * self.model.observation_table[:, a, o]
        z = np.sum(u)
        if z == 0.0:
            if throw_error:
                raise ValueError('Impossible belief: the observation has probability 0 under this belief and action.')
            return Belief(self.model, u)
        return Belief(self.model, u / z)

    @property
    def entropy(self) -> float:
        b = self.values
        nz = b > 0
        return float(-np.sum(b[nz] * np.log(b[nz])))

    def random_state(self, u: float) -> int:
        cdf = np.cumsum(self.values)
        return min(int(np.searchsorted(cdf, u * cdf[-1], side='right')), self.model.state_count - 1)

    @property
    def on_gpu(self) -> 'Belief':
        return self

    @property
    def on_cpu(self) -> 'Belief':
        return self


class BeliefSet:
    '''
    (N, S) set of beliefs.  Call sites: agents/pbvi_agent.py:717,719,721-726,783-790,811;
    agents/infotaxis_agent.py:232,257-270,335-342,363.
    '''
    def __init__(self, model: POMDP, beliefs) -> None:
        self.model = model
        if isinstance(beliefs, np.ndarray):
            self._belief_array = np.asarray(beliefs, dtype=float).reshape(-1, model.state_count)
        else:
            beliefs = list(beliefs)
            self._belief_array = (np.array([b.values for b in beliefs], dtype=float)
                                  if len(beliefs) else np.zeros((0, model.state_count)))
        self.is_on_gpu = False

    @property
    def belief_array(self) -> np.ndarray:
        return self._belief_array

    @property
    def belief_list(self) -> list:
        return [Belief(self.model, row) for row in self._belief_array]

    def __len__(self) -> int:
        return len(self._belief_array)

    @property
    def on_gpu(self) -> 'BeliefSet':
        return self

    @property
    def on_cpu(self) -> 'BeliefSet':
        return self

    @property
    def entropies(self) -> np.ndarray:
        b = self._belief_array
        with np.errstate(divide='ignore', invalid='ignore'):
            t = np.where(b > 0, b * np.log(b), 0.0)
        return -np.sum(t, axis=1)

    def update(self,
               actions: np.ndarray,
               observations: np.ndarray,
               raise_on_impossible_belief: bool = True,
               use_reachability: bool = True,
               return_provenance: bool = False):
        '''
        b'_n(s') = O[s', a_n, o_n] * sum_s T(s'|s, a_n) b_n(s) / Z_n   (SURVEY §8 a1).
        Rows with Z_n == 0 are dropped (or raise).  provenance[:, 0] = surviving source rows.
        '''
        actions = np.asarray(actions, dtype=np.int64).ravel()
        observations = np.asarray(observations, dtype=np.int64).ravel()
        n = len(self)
        assert actions.shape == (n,) and observations.shape == (n,)

        u = _push(self.model, self._belief_array, actions) * self.model.observation_table[:, actions, observations].T
        z = np.sum(u, axis=1)
        ok = z > 0
        if raise_on_impossible_belief and not np.all(ok):
            raise ValueError(f'Impossible belief update for rows {np.flatnonzero(~ok).tolist()}')

        new_set = BeliefSet(self.model, u[ok] / z[ok, None])
        if return_provenance:
            provenance = np.stack([np.flatnonzero(ok), actions[ok], observations[ok]], axis=1)
            return new_set, provenance
        return new_set

    def unnormalised_successors(self) -> tuple[np.ndarray, np.ndarray]:
        'u[n,a,o,:] = O[:,a,o] * push_a(b_n) and p[n,a,o] = sum_s u  (helper for successors/backup).'
        n, S = self._belief_array.shape
        A, O = self.model.action_count, self.model.observation_count
        u = np.empty((n, A, O, S))
        for a in range(A):
            b_a = _push(self.model, self._belief_array, np.full(n, a))
            u[:, a, :, :] = b_a[:, None, :] * self.model.observation_table[:, a, :].T[None, :, :]
        return u, np.sum(u, axis=3)

    def generate_all_successors(self,
                                use_reachability: bool = True,
                                raise_on_impossible_belief: bool = True,
                                return_provenance: bool = False):
        '''
        All b_{a,o} with P(o|b,a) > 0, ordered lexicographically by (belief row, action,
        observation); provenance (n, 3) = (b, a, o)  (agents/infotaxis_agent.py:260-267).
        '''
        u, p = self.unnormalised_successors()
        n, A, O, S = u.shape
        ok = p > 0
        if raise_on_impossible_belief and not np.all(ok):
            raise ValueError('Impossible successor belief encountered')
        b_idx, a_idx, o_idx = np.nonzero(ok)
        succ = u[b_idx, a_idx, o_idx, :] / p[b_idx, a_idx, o_idx][:, None]
        new_set = BeliefSet(self.model, succ)
        if return_provenance:
            return new_set, np.stack([b_idx, a_idx, o_idx], axis=1)
        return new_set

    def union(self, other: 'BeliefSet') -> 'BeliefSet':
        return BeliefSet(self.model, np.vstack([self._belief_array, other._belief_array]))


class ValueFunction:
    '''
    Set of alpha vectors.  Call sites: agents/pbvi_agent.py:523 (save), :574-577 (load),
    :690-692 (alpha_vector_array, actions, ctor with alpha_vectors=, action_list=), :741
    (evaluate_at).  The on-disk layout of the real package is unknown (no sample in the
    reference tree); the layout defined here is one (G, S+1) float64 array, column 0 = action.
    '''
    def __init__(self, model: POMDP, alpha_vectors=None, action_list=None) -> None:
        self.model = model
        if alpha_vectors is None:
            alpha_vectors = np.zeros((0, model.state_count))
            action_list = np.zeros(0, dtype=np.int64)
        self._alpha_vector_array = np.asarray(alpha_vectors, dtype=float).reshape(-1, model.state_count)
        self._actions = np.asarray(action_list, dtype=np.int64).ravel()
        assert len(self._actions) == len(self._alpha_vector_array)
        self.is_on_gpu = False

    @property
    def alpha_vector_array(self) -> np.ndarray:
        return self._alpha_vector_array

    @property
    def actions(self) -> np.ndarray:
        return self._actions

    def __len__(self) -> int:
        return len(self._alpha_vector_array)

    @property
    def on_gpu(self) -> 'ValueFunction':
        return self

    @property
    def on_cpu(self) -> 'ValueFunction':
        return self

    def evaluate_at(self, belief) -> tuple:
        '''
        V(b) = max_g b . alpha_g ; action = actions[first argmax]  (agents/pbvi_agent.py:741).
        Accepts a BeliefSet (-> arrays) or a Belief (-> scalars).
        '''
        if isinstance(belief, Belief):
            v = self._alpha_vector_array @ belief.values
            g = int(np.argmax(v))
            return float(v[g]), int(self._actions[g])
        v = belief.belief_array @ self._alpha_vector_array.T
        g = np.argmax(v, axis=1)
        return v[np.arange(len(g)), g], self._actions[g]

    def best_alpha_indices(self, belief_set: BeliefSet) -> np.ndarray:
        return np.argmax(belief_set.belief_array @ self._alpha_vector_array.T, axis=1)

    def extend(self, other: 'ValueFunction') -> None:
        self._alpha_vector_array = np.vstack([self._alpha_vector_array, other._alpha_vector_array])
        self._actions = np.concatenate([self._actions, other._actions])

    def prune(self, level: int = 1) -> int:
        '''
        level 0: nothing; level >= 1: drop exact duplicate (vector, action) rows keeping the first
        occurrence and the original order; level >= 2: also drop vectors point-wise dominated
        by another single vector.  Returns how many were removed.
        '''
        before = len(self)
        if level >= 1 and before > 0:
            keyed = np.hstack([self._actions[:, None].astype(float), self._alpha_vector_array])
            _, first = np.unique(keyed, axis=0, return_index=True)
            keep = np.sort(first)
            self._alpha_vector_array = self._alpha_vector_array[keep]
            self._actions = self._actions[keep]
        if level >= 2 and len(self) > 1:
            av = self._alpha_vector_array
            keep = np.ones(len(av), dtype=bool)
            for i in range(len(av)):
                others = keep.copy()
                others[i] = False
                if np.any(np.all(av[others] >= av[i][None, :], axis=1) & np.any(av[others] > av[i][None, :], axis=1)):
                    keep[i] = False
            self._alpha_vector_array = av[keep]
            self._actions = self._actions[keep]
        return before - len(self)

    def save(self, folder: str = './', file_name: str = 'Value_Function.npy') -> None:
        if not file_name.endswith('.npy'):
            file_name += '.npy'
        np.save(os.path.join(folder, file_name), np.hstack([self._actions[:, None].astype(float), self._alpha_vector_array]))

    @classmethod
    def load(cls, file: str, model: POMDP) -> 'ValueFunction':
        arr = np.load(file)
        return cls(model, alpha_vectors=arr[:, 1:], action_list=arr[:, 0].astype(np.int64))

    def plot(self, *args, **kwargs) -> None:
        raise NotImplementedError('Plotting is out of scope for the oracle')


from . import solvers  # noqa: E402  (needs the classes above)
