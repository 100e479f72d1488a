'''
ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see package docstring).

Restated solvers of `pomdp_toolkit.solvers` as called by the reference:
  VI.solve        agents/qmdp_agent.py:145-153 (+ implicitly by FSVI/HSVI, agents/fsvi_agent.py:154-158)
  FSVI.solve      agents/fsvi_agent.py:201-224
  PBVI_RA/SSRA/SSGA/SSEA.solve   agents/pbvi_ra_agent.py:168, pbvi_ssra_agent.py:170,
                  pbvi_ssga_agent.py:176-198, pbvi_ssea_agent.py:170
  PBVI_GER / Perseus / HSVI      agents/pbvi_ger_agent.py:170, perseus_agent.py:176-196,
                  hsvi_agent.py:184-208 -- share the same backup; their expand heuristics are
                  SURVEY §8(f2) "next" and raise NotImplementedError here.

Maths: SURVEY.md Appendix B.
  backup   Gamma_{a,o}^g(s) = gamma * sum_r T(s,a,r) O[R[s,a,r],a,o] alpha_g(R[s,a,r])
           g*(b,a,o) = argmax_g b.Gamma_{a,o}^g ;  alpha_a^b = r_a + sum_o Gamma_{a,o}^{g*}
           a*(b) = argmax_a b.alpha_a^b ; emit (alpha_{a*}^b, a*)
  VI       Q_a(s) = r_a(s) + gamma * sum_r T(s,a,r) max_a' Q_a'(R[s,a,r])
           stop after the first sweep whose max |V_k - V_{k-1}| < eps / (1 - gamma); with
           gamma=.99, eps=1e-6 and V_0 = 0 this is sweep 918 on every model that has an end
           state, which is the only numeric pin the reference's notebooks offer
           (experiments/pomdp_agent_training.ipynb:115-116 and four other notebooks).

Random draws: every stochastic choice consumes `rng.random()` uniforms through inverse-CDF
sampling, in a documented order, so that a device implementation fed the same generator state
reproduces the same (action, observation) sequences.
'''
from __future__ import annotations

import time
import numpy as np

from . import POMDP, Belief, BeliefSet, ValueFunction


class SolverHistory:
    '''
    Statistics holder returned by every solve().  Field names follow the reference's (never
    instantiated) TrainingHistory, agents/pbvi_agent.py:139-234; the reference only reads
    `.summary` (agents/fsvi_agent.py:236-237).
    '''
    def __init__(self, model: POMDP, tracking_level: int, gamma: float, eps: float, solver: str) -> None:
        self.model = model
        self.tracking_level = tracking_level
        self.gamma = gamma
        self.eps = eps
        self.solver = solver
        self.run_ts = time.time()
        self.expansion_times: list[float] = []
        self.backup_times: list[float] = []
        self.pruning_times: list[float] = []
        self.alpha_vector_counts: list[int] = []
        self.beliefs_counts: list[int] = []
        self.prune_counts: list[int] = []
        self.value_function_changes: list[float] = []
        self.iteration_times: list[float] = []
        self.value_functions: list[ValueFunction] = []
        self.belief_sets: list[BeliefSet] = []
        self.total_time = 0.0

    @property
    def summary(self) -> str:
        m = self.model
        if self.solver == 'VI':
            n = len(self.iteration_times)
            tot = float(np.sum(self.iteration_times))
            return (f'Summary of Value Iteration run\n'
                    f'  - Model: {m.state_count}-state, {m.action_count}-action\n'
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
# Value iteration on the fully observable MDP
# =================================================================================================
class VI:
    @staticmethod
    def sweep(model: POMDP, Q: np.ndarray, gamma: float) -> np.ndarray:
        'One Bellman sweep on Q (A, S) -> (A, S).'
        V = np.max(Q, axis=0)
        return model.expected_rewards_table.T + gamma * np.einsum('sar,sar->as', model.reachable_probabilities, V[model.reachable_states])

    @staticmethod
    def solve(model: POMDP,
              horizon: int = 1000,
              initial_value_function: ValueFunction = None,
              gamma: float = 0.99,
              eps: float = 1e-6,
              use_gpu: bool = False,
              use_reachability: bool = True,
              history_tracking_level: int = 1,
              print_progress: bool = True):
        hist = SolverHistory(model, history_tracking_level, gamma, eps, 'VI')
        if initial_value_function is None:
            Q = np.zeros((model.action_count, model.state_count))
        else:
            Q = initial_value_function.alpha_vector_array.copy()
        V = np.max(Q, axis=0)
        threshold = eps / (1.0 - gamma)
        for _ in range(horizon):
            t0 = time.perf_counter()
            Q = VI.sweep(model, Q, gamma)
            V_new = np.max(Q, axis=0)
            change = float(np.max(np.abs(V_new - V)))
            V = V_new
            hist.iteration_times.append(time.perf_counter() - t0)
            hist.value_function_changes.append(change)
            if change < threshold:
                break
        vf = ValueFunction(model, Q, model.actions)
        hist.alpha_vector_counts.append(len(vf))
        hist.total_time = float(np.sum(hist.iteration_times))
        return vf, hist


# =================================================================================================
# Point-based value iteration: shared backup + pluggable expand
# =================================================================================================
def gamma_a_o(model: POMDP, alpha_vectors: np.ndarray, gamma: float) -> np.ndarray:
    '''
    Gamma[a, o, g, s] = gamma * sum_r T(s,a,r) * O[R[s,a,r], a, o] * alpha_g(R[s,a,r])
    (SURVEY App. B "Backup (pull)").
    '''
    A, O = model.action_count, model.observation_count
    G, S = alpha_vectors.shape
    out = np.zeros((A, O, G, S))
    for a in range(A):
        for r in range(model.reachable_state_count):
            dest = model.reachable_states[:, a, r]                                      # (S,)
            w = model.reachable_transition_observation_table[:, a, r, :]                # (S, O)
            out[a] += w.T[:, None, :] * alpha_vectors[:, dest][None, :, :]
    return gamma * out


def backup(model: POMDP,
           belief_set: BeliefSet,
           value_function: ValueFunction,
           gamma: float = 0.99,
           append: bool = False,
           belief_dominance_prune: bool = True,
           return_details: bool = False):
    '''
    One point-based backup of `value_function` at the points of `belief_set`.
    Returns the new ValueFunction (unioned with the old one when append=True).
    '''
    B = belief_set.belief_array                                   # (B, S)
    Gm = gamma_a_o(model, value_function.alpha_vector_array, gamma)   # (A, O, G, S)
    A, O, G, S = Gm.shape

    # Step 2: best alpha per (b, a, o)
    scores = np.einsum('bs,aogs->baog', B, Gm)                    # (B, A, O, G)
    best_g = np.argmax(scores, axis=3)                            # (B, A, O) first maximum
    a_idx = np.arange(A)[None, :, None]
    o_idx = np.arange(O)[None, None, :]
    best_vec = Gm[a_idx, o_idx, best_g, :]                        # (B, A, O, S)
    alpha_a = model.expected_rewards_table.T[None, :, :] + np.sum(best_vec, axis=2)   # (B, A, S)

    # Step 3: best action per belief
    q = np.einsum('bas,bs->ba', alpha_a, B)
    best_a = np.argmax(q, axis=1)
    new_alpha = alpha_a[np.arange(len(B)), best_a, :]
    new_value = q[np.arange(len(B)), best_a]

    keep = np.ones(len(B), dtype=bool)
    if belief_dominance_prune:
        old_value = np.max(B @ value_function.alpha_vector_array.T, axis=1)
        keep = new_value > old_value

    new_vf = ValueFunction(model, new_alpha[keep], best_a[keep])
    new_vf.prune(1)
    if append:
        merged = ValueFunction(model, value_function.alpha_vector_array.copy(), value_function.actions.copy())
        merged.extend(new_vf)
        merged.prune(1)
        new_vf = merged
    if return_details:
        return new_vf, dict(best_g=best_g, best_a=best_a, new_alpha=new_alpha, new_value=new_value, keep=keep, q=q)
    return new_vf


class _PBVI:
    name = 'PBVI'
    full_backup = True        # back up the whole belief set and replace the value function
    dominance_prune = False

    # -- to be provided by subclasses -------------------------------------------------------
    @classmethod
    def expand(cls, model, belief_set, value_function, max_generation, rng, **kw) -> BeliefSet:
        raise NotImplementedError

    @classmethod
    def prepare(cls, model, gamma, eps, print_progress, **kw) -> dict:
        return {}

    @classmethod
    def solve(cls,
              model: POMDP,
              expansions: int = 10,
              update_passes: int = 1,
              max_belief_growth: int = 10,
              initial_belief: BeliefSet | Belief = None,
              initial_value_function: ValueFunction = None,
              prune_level: int = 1,
              prune_interval: int = 10,
              limit_value_function_size: int = -1,
              gamma: float = 0.99,
              eps: float = 1e-6,
              convergence_stop: bool = False,
              use_gpu: bool = False,
              use_reachability: bool = True,
              rng: np.random.Generator = None,
              history_tracking_level: int = 1,
              print_progress: bool = True,
              print_stats: bool = True,
              **flavour):
        rng = np.random.default_rng() if rng is None else rng
        hist = SolverHistory(model, history_tracking_level, gamma, eps, cls.name)
        t_start = time.perf_counter()

        extra = cls.prepare(model, gamma, eps, print_progress, use_reachability=use_reachability, **flavour)

        # Initial belief set
        if initial_belief is None:
            belief_set = BeliefSet(model, [Belief(model)])
        elif isinstance(initial_belief, Belief):
            belief_set = BeliefSet(model, [initial_belief])
        else:
            belief_set = initial_belief

        # Initial value function: immediate expected rewards, one vector per action
        if initial_value_function is None:
            value_function = ValueFunction(model, model.expected_rewards_table.T.copy(), model.actions)
        else:
            value_function = ValueFunction(model, initial_value_function.alpha_vector_array.copy(), initial_value_function.actions.copy())

        old_values = np.max(belief_set.belief_array @ value_function.alpha_vector_array.T, axis=1)
        backup_count = 0
        for _ in range(expansions):
            # 1. Expand
            t0 = time.perf_counter()
            new_beliefs = cls.expand(model, belief_set, value_function, max_belief_growth, rng, **extra)
            belief_set = belief_set.union(new_beliefs)
            hist.expansion_times.append(time.perf_counter() - t0)
            hist.beliefs_counts.append(len(belief_set))

            # 2. Backup
            t0 = time.perf_counter()
            target = belief_set if cls.full_backup else new_beliefs
            for _ in range(update_passes):
                if len(target) == 0:
                    break
                value_function = backup(model, target, value_function, gamma,
                                        append=(not cls.full_backup),
                                        belief_dominance_prune=cls.dominance_prune)
                backup_count += 1
            hist.backup_times.append(time.perf_counter() - t0)

            # 3. Prune / limit
            t0 = time.perf_counter()
            pruned = 0
            if prune_interval > 0 and (backup_count % prune_interval) == 0:
                pruned = value_function.prune(prune_level)
            if 0 < limit_value_function_size < len(value_function):
                drop = rng.choice(len(value_function), size=min(max_belief_growth, len(value_function) - 1), replace=False)
                keep = np.ones(len(value_function), dtype=bool)
                keep[drop] = False
                value_function = ValueFunction(model, value_function.alpha_vector_array[keep], value_function.actions[keep])
                pruned += int(np.sum(~keep))
            hist.pruning_times.append(time.perf_counter() - t0)
            hist.prune_counts.append(pruned)
            hist.alpha_vector_counts.append(len(value_function))

            # 4. Convergence
            new_values = np.max(belief_set.belief_array @ value_function.alpha_vector_array.T, axis=1)
            change = float(np.max(np.abs(new_values[:len(old_values)] - old_values))) if len(old_values) else np.inf
            old_values = new_values
            hist.value_function_changes.append(change)
            if history_tracking_level >= 2:
                hist.value_functions.append(value_function)
                hist.belief_sets.append(belief_set)
            if convergence_stop and change < eps:
                break

        hist.total_time = time.perf_counter() - t_start
        hist.final_belief_set = belief_set
        if 'mdp_policy' in extra and flavour.get('return_mdp_policy', False):
            return value_function, hist, extra['mdp_policy']
        return value_function, hist


# -------------------------------------------------------------------------------------------------
class FSVI(_PBVI):
    '''
    Forward Search Value Iteration (Shani et al. 2007, cited at agents/fsvi_agent.py:12).
    expand: one trajectory from the first (initial) belief under the MDP policy:
        u0 -> s ~ b0 ; repeat max_generation times:
            a = argmax_a Q_mdp[a, s] ; (u -> s' ~ T(.|s,a) only when R > 1) ; u -> o ~ O(.|s',a)
            b <- update(b, a, o) ; collect b ; stop once s' is an end state.
    backup: on the newly collected beliefs only, appended to the value function
    ("performs the backup only on set of beliefs generated by the expand function",
    agents/fsvi_agent.py:139).
    '''
    name = 'FSVI'
    full_backup = False
    dominance_prune = True

    @classmethod
    def prepare(cls, model, gamma, eps, print_progress, mdp_policy=None, vi_horizon=1000, use_reachability=True, **kw):
        if mdp_policy is None:
            mdp_policy, _ = VI.solve(model, horizon=vi_horizon, gamma=gamma, eps=eps, print_progress=print_progress)
        return {'mdp_policy': mdp_policy}

    @staticmethod
    def trajectory(model: POMDP, b0: np.ndarray, Q: np.ndarray, max_generation: int, rng) -> list[tuple[int, int]]:
        'The (action, observation) sequence of one forward-search trajectory (host-side, belief-free).'
        cdf = np.cumsum(b0)
        s = min(int(np.searchsorted(cdf, rng.random() * cdf[-1], side='right')), model.state_count - 1)
        seq = []
        for _ in range(max_generation):
            a = int(np.argmax(Q[:, s]))
            s_p = model.transition(s, a, rng.random() if model.reachable_state_count > 1 else 0.0)
            o = model.observe(s_p, a, rng.random())
            seq.append((a, o))
            s = s_p
            if model.end_mask[s]:
                break
        return seq

    @classmethod
    def expand(cls, model, belief_set, value_function, max_generation, rng, mdp_policy=None, **kw) -> BeliefSet:
        b = Belief(model, belief_set.belief_array[0])
        out = []
        for a, o in cls.trajectory(model, b.values, mdp_policy.alpha_vector_array, max_generation, rng):
            b = b.update(a, o)
            out.append(b)
        return BeliefSet(model, out)


class PBVI_SSRA(_PBVI):
    '''Stochastic Simulation with Random Action (Pineau et al. 2003): for up to max_generation
    beliefs b of the set (in order): u -> s ~ b ; u -> a uniform ; (u -> s') ; u -> o ; add update(b,a,o).'''
    name = 'PBVI_SSRA'

    @classmethod
    def choose_action(cls, model, b, value_function, rng, **kw) -> int:
        return min(int(rng.random() * model.action_count), model.action_count - 1)

    @classmethod
    def expand(cls, model, belief_set, value_function, max_generation, rng, **kw) -> BeliefSet:
        out = []
        for row in belief_set.belief_array[:max_generation]:
            b = Belief(model, row)
            s = b.random_state(rng.random())
            a = cls.choose_action(model, b, value_function, rng, **kw)
            s_p = model.transition(s, a, rng.random() if model.reachable_state_count > 1 else 0.0)
            o = model.observe(s_p, a, rng.random())
            out.append(b.update(a, o))
        return BeliefSet(model, out)


class PBVI_SSGA(PBVI_SSRA):
    '''Stochastic Simulation with Greedy Action: epsilon-greedy on the current value function
    (`epsilon`, `epsilon_decay` kwargs, agents/pbvi_ssga_agent.py:176-198).'''
    name = 'PBVI_SSGA'

    @classmethod
    def prepare(cls, model, gamma, eps, print_progress, epsilon=0.1, epsilon_decay=1.0, **kw):
        return {'epsilon_state': {'epsilon': float(epsilon), 'decay': float(epsilon_decay)}}

    @classmethod
    def choose_action(cls, model, b, value_function, rng, epsilon_state=None, **kw) -> int:
        if rng.random() < epsilon_state['epsilon']:
            return min(int(rng.random() * model.action_count), model.action_count - 1)
        return value_function.evaluate_at(b)[1]

    @classmethod
    def expand(cls, model, belief_set, value_function, max_generation, rng, epsilon_state=None, **kw) -> BeliefSet:
        out = super().expand(model, belief_set, value_function, max_generation, rng, epsilon_state=epsilon_state)
        epsilon_state['epsilon'] *= epsilon_state['decay']
        return out


class PBVI_SSEA(_PBVI):
    '''Stochastic Simulation with Exploratory Action: for each belief simulate one step per
    action and keep the successor farthest (L1) from the current set.'''
    name = 'PBVI_SSEA'

    @classmethod
    def expand(cls, model, belief_set, value_function, max_generation, rng, **kw) -> BeliefSet:
        current = belief_set.belief_array
        out = []
        for row in belief_set.belief_array[:max_generation]:
            b = Belief(model, row)
            best, best_d = None, -1.0
            for a in range(model.action_count):
                s = b.random_state(rng.random())
                s_p = model.transition(s, a, rng.random() if model.reachable_state_count > 1 else 0.0)
                o = model.observe(s_p, a, rng.random())
                b_a = b.update(a, o)
                d = float(np.min(np.sum(np.abs(current - b_a.values[None, :]), axis=1)))
                if d > best_d:
                    best, best_d = b_a, d
            out.append(best)
            current = np.vstack([current, best.values[None, :]])
        return BeliefSet(model, out)


class PBVI_RA(_PBVI):
    '''Random belief selection: max_generation random points of the simplex (normalised
    exponentials of uniforms, i.e. Dirichlet(1)), drawn as rng.random(S) per point.'''
    name = 'PBVI_RA'

    @classmethod
    def expand(cls, model, belief_set, value_function, max_generation, rng, **kw) -> BeliefSet:
        pts = -np.log(1.0 - rng.random((max_generation, model.state_count)))
        return BeliefSet(model, pts / np.sum(pts, axis=1, keepdims=True))


class PBVI_GER(_PBVI):
    '''
    Greedy Error Reduction (Pineau, Gordon & Thrun 2006, the PBVI variant the reference cites at agents/pbvi_agent.py:239
    and describes at agents/pbvi_ger_agent.py:10: "choosing belief points that will most decrease the error in the value
    function").  For a candidate successor b' the error bound is
        eps(b') = min_{b in B} sum_s (b'(s) - b(s)) * ( (Vmax - alpha_b(s)) if b'(s) >= b(s) else (Vmin - alpha_b(s)) ),
    alpha_b = the best vector at b, Vmax = max r / (1 - gamma), Vmin = min r / (1 - gamma).
    Per belief b:  err(b, a) = sum_o P(o|b,a) eps(b_{a,o});  a* = first argmax_a;  o* = first argmax_o P(o|b,a*) eps(b_{a*,o});
    the `max_generation` beliefs with the largest err(b, a*) (stable order) contribute their b_{a*,o*}.  Deterministic.
    '''
    name = 'PBVI_GER'

    @classmethod
    def prepare(cls, model, gamma, eps, print_progress, **kw) -> dict:
        return {'gamma': gamma}

    @classmethod
    def expand(cls, model, belief_set, value_function, max_generation, rng, gamma=0.99, **kw) -> BeliefSet:
        B = belief_set.belief_array
        n = len(B)
        av = value_function.alpha_vector_array
        alpha_b = av[np.argmax(B @ av.T, axis=1)]                      # (n, S)
        r = model.expected_rewards_table
        vmax, vmin = float(np.max(r)) / (1.0 - gamma), float(np.min(r)) / (1.0 - gamma)
        u, p = belief_set.unnormalised_successors()                    # (n, A, O, S), (n, A, O)
        A, O = model.action_count, model.observation_count
        eps_bao = np.zeros((n, A, O))
        for i in range(n):
            for a in range(A):
                for o in range(O):
                    if p[i, a, o] <= 0:
                        continue
                    bp = u[i, a, o] / p[i, a, o]
                    d = bp[None, :] - B                                # (n, S)
                    w = np.where(d >= 0, vmax - alpha_b, vmin - alpha_b)
                    eps_bao[i, a, o] = float(np.min(np.sum(d * w, axis=1)))
        contrib = p * eps_bao                                          # (n, A, O)
        err = np.sum(contrib, axis=2)                                  # (n, A)
        a_star = np.argmax(err, axis=1)
        score = err[np.arange(n), a_star]
        o_star = np.argmax(contrib[np.arange(n), a_star, :], axis=1)
        order = np.argsort(-score, kind='stable')[:max_generation]
        out = [u[i, a_star[i], o_star[i]] / p[i, a_star[i], o_star[i]] for i in order if p[i, a_star[i], o_star[i]] > 0]
        return BeliefSet(model, np.array(out).reshape(-1, model.state_count))


class Perseus(_PBVI):
    '''
    Perseus (Spaan & Vlassis 2005; reference agents/perseus_agent.py:99,176-196).
    Expand: a random walk of `max_generation` steps continuing from the LAST belief of the set (the initial belief the first
    time): s ~ b once, then per step a = uniform random action (or the value function's greedy action at the current belief when
    `use_policy_to_choose_actions`), s' = T(s, a), o ~ O(.|s', a), b <- update(b, a, o); stops early in an end state.
    Draws: rng.random() for s, then per step [rng.random() for a when random] and rng.random() for o.
    Backup: one randomised improve-only stage over the WHOLE belief set with the value function V of the stage's start:
        todo = all beliefs; repeat: i = todo[min(int(rng.random() * len(todo)), len(todo) - 1)];
        alpha = backup(b_i, V); keep alpha if alpha.b_i >= V(b_i) else the best old vector at b_i;
        drop from todo every belief whose value under the vectors kept so far is >= its value under V.
    The kept vectors replace the value function.
    '''
    name = 'Perseus'

    @classmethod
    def prepare(cls, model, gamma, eps, print_progress, use_policy_to_choose_actions=False, **kw) -> dict:
        return {'use_policy': bool(use_policy_to_choose_actions)}

    @classmethod
    def expand(cls, model, belief_set, value_function, max_generation, rng, use_policy=False, **kw) -> BeliefSet:
        b = Belief(model, belief_set.belief_array[-1])
        s = b.random_state(rng.random())
        out = []
        for _ in range(max_generation):
            if use_policy:
                a = int(value_function.actions[np.argmax(value_function.alpha_vector_array @ b.values)])
            else:
                a = min(int(rng.random() * model.action_count), model.action_count - 1)
            s_p = model.transition(s, a, 0.0)
            o = model.observe(s_p, a, rng.random())
            b = b.update(a, o)
            out.append(b)
            s = s_p
            if model.end_mask[s]:
                break
        return BeliefSet(model, out)

    @classmethod
    def stage(cls, model, belief_set, value_function, gamma, rng) -> ValueFunction:
        B = belief_set.belief_array
        av, ac = value_function.alpha_vector_array, value_function.actions
        scores = B @ av.T
        v_old, g_old = np.max(scores, axis=1), np.argmax(scores, axis=1)
        todo = np.arange(len(B))
        kept_v, kept_a = [], []
        v_new = np.full(len(B), -np.inf)
        while len(todo):
            i = int(todo[min(int(rng.random() * len(todo)), len(todo) - 1)])
            _, det = backup(model, BeliefSet(model, B[i:i + 1]), value_function, gamma, belief_dominance_prune=False, return_details=True)
            if det['new_value'][0] >= v_old[i]:
                vec, act = det['new_alpha'][0], int(det['best_a'][0])
            else:
                vec, act = av[g_old[i]], int(ac[g_old[i]])
            kept_v.append(vec)
            kept_a.append(act)
            v_new = np.maximum(v_new, B @ vec)
            todo = todo[v_new[todo] < v_old[todo]]
        return ValueFunction(model, np.array(kept_v), np.array(kept_a))

    @classmethod
    def solve(cls, model, expansions=10, update_passes=1, max_belief_growth=10, initial_belief=None, initial_value_function=None,
              prune_level=1, prune_interval=10, limit_value_function_size=-1, gamma=0.99, eps=1e-6, convergence_stop=False,
              use_gpu=False, use_reachability=True, rng=None, history_tracking_level=1, print_progress=True, print_stats=True,
              use_policy_to_choose_actions=False, **kw):
        rng = np.random.default_rng() if rng is None else rng
        hist = SolverHistory(model, history_tracking_level, gamma, eps, cls.name)
        t_start = time.perf_counter()
        if initial_belief is None:
            belief_set = BeliefSet(model, [Belief(model)])
        elif isinstance(initial_belief, Belief):
            belief_set = BeliefSet(model, [initial_belief])
        else:
            belief_set = initial_belief
        if initial_value_function is None:
            value_function = ValueFunction(model, model.expected_rewards_table.T.copy(), model.actions)
        else:
            value_function = ValueFunction(model, initial_value_function.alpha_vector_array.copy(), initial_value_function.actions.copy())
        old_values = np.max(belief_set.belief_array @ value_function.alpha_vector_array.T, axis=1)
        backup_count = 0
        for _ in range(expansions):
            t0 = time.perf_counter()
            belief_set = belief_set.union(cls.expand(model, belief_set, value_function, max_belief_growth, rng, use_policy=use_policy_to_choose_actions))
            hist.expansion_times.append(time.perf_counter() - t0)
            hist.beliefs_counts.append(len(belief_set))
            t0 = time.perf_counter()
            for _ in range(update_passes):
                value_function = cls.stage(model, belief_set, value_function, gamma, rng)
                backup_count += 1
            hist.backup_times.append(time.perf_counter() - t0)
            t0 = time.perf_counter()
            pruned = value_function.prune(prune_level) if (prune_interval > 0 and backup_count % prune_interval == 0) else 0
            hist.pruning_times.append(time.perf_counter() - t0)
            hist.prune_counts.append(pruned)
            hist.alpha_vector_counts.append(len(value_function))
            new_values = np.max(belief_set.belief_array @ value_function.alpha_vector_array.T, axis=1)
            change = float(np.max(np.abs(new_values[:len(old_values)] - old_values))) if len(old_values) else np.inf
            old_values = new_values
            hist.value_function_changes.append(change)
            if convergence_stop and change < eps:
                break
        hist.total_time = time.perf_counter() - t_start
        hist.final_belief_set = belief_set
        return value_function, hist


class HSVI(_PBVI):
    '''
    Heuristic Search Value Iteration (Smith & Simmons 2004; reference agents/hsvi_agent.py:100-102,184-208: parameters
    `mdp_policy`, `vi_horizon`, `epsilon`, and "backup only on the beliefs generated by the expand function").
    Lower bound = the alpha vectors, upper bound = the MDP policy (Q_MDP): Vup(b) = max_a b.Q_a.
    Expand: one deterministic descent of at most `max_generation` steps from the initial belief (row 0):
        a* = first argmax_a b.Q_a;   for every o with P(o|b,a*) > 0:  gap_o = Vup(b_o) - Vlow(b_o);
        excess_o = P(o|b,a*) * (gap_o - epsilon * gamma^-(depth+1));   o* = first argmax_o excess_o;
        stop when excess_{o*} <= 0, else b <- b_{a*,o*} and collect it.
    Backup on the new beliefs only, appended, kept only where they improve (like FSVI).
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
    def expand(cls, model, belief_set, value_function, max_generation, rng, mdp_policy=None, epsilon=0.99, gamma=0.99, **kw) -> BeliefSet:
        Q = mdp_policy.alpha_vector_array                                 # (A, S)
        av = value_function.alpha_vector_array
        b = belief_set.belief_array[0]
        out = []
        for depth in range(max_generation):
            a = int(np.argmax(Q @ b))
            u, p = BeliefSet(model, b[None, :]).unnormalised_successors()
            u, p = u[0, a], p[0, a]                                       # (O, S), (O,)
            excess = np.full(model.observation_count, -np.inf)
            for o in range(model.observation_count):
                if p[o] <= 0:
                    continue
                bo = u[o] / p[o]
                gap = float(np.max(Q @ bo)) - float(np.max(av @ bo))
                excess[o] = p[o] * (gap - epsilon * gamma ** (-(depth + 1)))
            o = int(np.argmax(excess))
            if not (excess[o] > 0):
                break
            b = u[o] / p[o]
            out.append(b)
        return BeliefSet(model, np.array(out).reshape(-1, model.state_count))
