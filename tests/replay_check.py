'''
Replay of a recorded simulation through the ORACLE (test infrastructure): every decision and every environment transition of a
device run is re-derived in float64 NumPy from the run's own history.

A comparison of final trajectories with the reference's golden run cannot tell a documented tie from a bug: a float32 near-tie
flips one argmax, the agent walks elsewhere, and — with the reference's sorted-time indexing (environment.py:598-601) — every
other agent's observations change from that step on.  The replay avoids the cascade: at every step it rebuilds the beliefs of
the LIVE agents from the run's recorded (action, observation) pairs with the oracle's update, and asserts that
  * the recorded action is an argmax of b . alpha in float64, or loses to it by less than `rel` relative (a documented tie),
  * positions, odor values, reached flags and terminations are exactly what the oracle's environment functions give,
  * agents whose update is impossible (normaliser 0) are terminated without success.
Only then are differences with the golden trajectories accepted, and only from the first tie on.
'''
import numpy as np


def oracle_setup(cfg: str, thresholds=None):
    'oracle/sim.py model of a synthetic configuration (independent of the product\'s converter).'
    from oracle import refloader
    refloader.install_oracle_toolkit()
    from oracle import sim
    import olfactory_navigation_b200.synthetic as syn
    c = syn.CONFIGS[cfg]
    data = syn.plume_frames(c['T'], c['h'], c['w'], source_y=c['source'][0], seed=syn.SEED)
    t = c['thresholds'] if thresholds is None else thresholds
    thr = np.concatenate([[-np.inf], np.atleast_1d(np.asarray(t, dtype=float)), [np.inf]])
    return sim, sim.build_model(data, c['margins'], c['source'], c['radius'], c['boundary'], thr, c['start_zone'],
                                0.0 if c['start_zone'] == 'odor_present' else None)


def replay(hist_actions, hist_positions, hist_observations, done_at_step, reached_source, cfg, alphas, alpha_actions, start_points,
           time_shift, *, quirk=True, rel=1e-5, infotaxis=False, thresholds=None):
    '''
    hist_*: dense per-step arrays over all n simulations as SimulationHistory exposes them (actions = movement vectors).
    Returns dict(decisions, exact, ties, first_tie_step).  Raises AssertionError on the first inconsistency.
    '''
    sim, setup = oracle_setup(cfg, thresholds)
    import pomdp_toolkit as tk
    model, geom, data, thr, action_set = setup['model'], setup['geom'], setup['data'], setup['thresholds'], setup['action_set']
    A = np.asarray(alphas, dtype=np.float64) if alphas is not None else None
    aact = np.asarray(alpha_actions) if alpha_actions is not None else None
    n = len(start_points)
    pos, tshift = np.array(start_points, dtype=int), np.array(time_shift, dtype=int)
    live = np.arange(n)
    B = np.tile(model.start_probabilities, (n, 1))
    done_at_step, reached_source = np.asarray(done_at_step), np.asarray(reached_source)
    decisions = exact = ties = 0
    first_tie = -1
    steps = len(hist_actions)
    for i in range(steps):
        if len(live) == 0:
            break
        mv = np.asarray(hist_actions[i])[live]
        ids = np.array([int(np.flatnonzero(np.all(action_set == m[None, :], axis=1))[0]) for m in mv])
        # ---- decision ----
        if infotaxis:
            dH = infotaxis_delta_H(model, B)
            top = dH.max(1)
            chosen = dH[np.arange(len(live)), ids]
            best_ids = np.argmax(dH, axis=1)
            slack = rel * np.maximum(np.abs(top), 1e-300)
        else:
            V = B @ A.T
            top = V.max(1)
            per_action = np.full((len(live), len(action_set)), -np.inf)
            for a in range(len(action_set)):
                if np.any(aact == a):
                    per_action[:, a] = V[:, aact == a].max(1)
            chosen = per_action[np.arange(len(live)), ids]
            best_ids = aact[np.argmax(V, axis=1)]
            slack = rel * np.abs(top)
        bad = chosen < top - slack
        assert not bad.any(), (f'step {i}: {bad.sum()} recorded actions are not argmax within {rel:g} relative '
                               f'(first: sim {live[bad][0]}, chosen {chosen[bad][0]!r} vs best {top[bad][0]!r})')
        same = ids == best_ids
        decisions += len(live)
        exact += int(same.sum())
        ties += int((~same).sum())
        if (~same).any() and first_tie < 0:
            first_tie = i
        # ---- environment (reference environment.py:711-769, 554-652, 655-680; agent.py:401-439) ----
        pos = sim.move(geom, pos, action_set[ids])
        assert np.array_equal(pos, np.asarray(hist_positions[i])[live]), f'step {i}: positions differ from the oracle move'
        t_eff = tshift + i
        if quirk:
            odor = sim.get_observation(geom, data, pos, t_eff)
        else:
            dp = pos - np.array([geom['my'], geom['mx']])[None, :]
            ok_zone = np.all((dp >= 0) & (dp < np.array([geom['h'], geom['w']])[None, :]), axis=1)
            odor = np.zeros(len(pos))
            odor[ok_zone] = data[(t_eff % geom['T'])[ok_zone], dp[ok_zone, 0], dp[ok_zone, 1]]
        rec = np.asarray(hist_observations[i], dtype=float)[live]
        assert np.array_equal(odor.astype(rec.dtype), rec) or np.array_equal(odor.astype(np.float32), rec.astype(np.float32)), \
            f'step {i}: recorded odor differs from the oracle get_observation'
        reached = sim.source_reached(geom, pos)
        obs_ids = sim.discretize(thr, odor, reached)
        # ---- belief update (oracle BeliefSet.update) ----
        nb, prov = tk.BeliefSet(model, B).update(ids, obs_ids, raise_on_impossible_belief=False, return_provenance=True)
        ok = np.isin(np.arange(len(live)), prov[:, 0])
        term = reached | ~ok
        # ---- bookkeeping against the run's own flags ----
        d = done_at_step[live]
        assert np.array_equal(d == i + 1, term), f'step {i}: terminations differ (oracle {np.flatnonzero(term)}, run {np.flatnonzero(d == i + 1)})'
        assert np.array_equal(reached_source[live][term], reached[term]), f'step {i}: reached_source flags differ'
        keep = ~term
        B = nb.belief_array[keep[ok]]
        pos, tshift, live = pos[keep], tshift[keep], live[keep]
    assert np.all(done_at_step[live] == -1), 'simulations alive at the end must have done_at_step == -1'
    return dict(decisions=decisions, exact=exact, ties=ties, first_tie_step=first_tie)


def infotaxis_delta_H(model, B):
    'reference agents/infotaxis_agent.py:257-293 in float64 (oracle/sim.py infotaxis_choose without the argmax).'
    import pomdp_toolkit as tk
    bs = tk.BeliefSet(model, B)
    Hb = bs.entropies
    u, p = bs.unnormalised_successors()
    with np.errstate(divide='ignore', invalid='ignore'):
        w = np.where(p[..., None] > 0, u / p[..., None], 0.0)
        Hao = -np.sum(np.where(w > 0, w * np.log(w), 0.0), axis=3)
    return Hb[:, None] - np.sum(p * Hao, axis=2)


def first_divergence(acts, poss, g_acts, g_poss):
    'First step at which two dense histories differ anywhere (None if identical over the common length).'
    for i in range(min(len(acts), len(g_acts))):
        if not (np.array_equal(acts[i], g_acts[i]) and np.array_equal(poss[i], g_poss[i])):
            return i
    return None
