'''
ORACLE — TEST INFRASTRUCTURE ONLY (parity unpinned upstream; see oracle/pomdp_toolkit/__init__.py).

NumPy restatement of the reference's per-step simulation path, used (a) as the checker of the device
run loop and (b) as the CPU baseline timed by bench.py (`cpu_baseline`, `--impl reference`): the
reference itself is pure Python on top of `pomdp_toolkit` and cannot travel to the GPU box.

Each function cites the reference code it follows (paths relative to /root/reference/olfactory_navigation/).
'''
from __future__ import annotations

import time

import numpy as np

from pomdp_toolkit import POMDP, BeliefSet, ValueFunction


def build_model(data: np.ndarray, margins, source_data_pos, source_radius: float, boundary: str, thresholds: np.ndarray,
                start_zone: str = 'data_zone', odor_present_threshold: float = None,
                action_set: np.ndarray = None) -> dict:
    '''
    Environment geometry + exact_converter in one go (environment.py:277-423 for margins / shape / source /
    start probabilities, agents/model_based_util/environment_converter.py:7-134 for the POMDP tables).
    Returns a dict with the oracle POMDP and the arrays the step functions need.
    '''
    T, h, w = data.shape
    margins = np.asarray(margins)
    if margins.ndim == 1:
        margins = np.stack([margins, margins], axis=1)
    H, W = h + margins[0].sum(), w + margins[1].sum()
    my, mx = int(margins[0, 0]), int(margins[1, 0])
    source = np.asarray(source_data_pos) + np.array([my, mx])
    action_set = np.array([[-1, 0], [0, 1], [1, 0], [0, -1]]) if action_set is None else np.asarray(action_set)
    thr = np.asarray(thresholds, dtype=float)

    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing='ij')
    disc = ((yy - source[0]) ** 2 + (xx - source[1]) ** 2) <= source_radius ** 2           # environment_converter.py:44-45
    start = np.zeros((H, W))
    if start_zone == 'data_zone':                                                          # environment.py:400-402
        start[my:my + h, mx:mx + w] = 1.0
    else:                                                                                  # environment.py:404-407
        t0 = 0 if odor_present_threshold is None else odor_present_threshold
        start[my:my + h, mx:mx + w] = (np.mean((data > t0).astype(int), axis=0) > 0).astype(float)
    start[disc] = 0.0                                                                      # environment.py:421-423
    start /= start.sum()

    O = len(thr)
    fields = np.zeros((H, W, O - 1))                                                       # environment_converter.py:66-74
    d = data[:, :, :, None]
    fields[my:my + h, mx:mx + w, :] = np.average((d >= thr[:-1]) & (d < thr[1:]), axis=0)
    S, A = H * W, len(action_set)
    obs = np.empty((S, A, O))
    obs[:, :, :O - 1] = fields.reshape(S, 1, O - 1)
    margin_mask = np.ones((H, W), dtype=bool)
    margin_mask[my:my + h, mx:mx + w] = False
    obs[margin_mask.ravel(), :, 0] = 1.0                                                   # :86-89
    obs[:, :, O - 1] = 0.0                                                                 # :91-94
    obs[disc.ravel(), :, :] = 0.0
    obs[disc.ravel(), :, O - 1] = 1.0
    assert np.all(obs.sum(axis=2) == 1.0)                                                  # :97

    geom = dict(H=H, W=W, my=my, mx=mx, h=h, w=w, T=T, source=source, r2=source_radius ** 2, boundary=boundary)
    cells = np.stack([yy.ravel(), xx.ravel()], axis=1)
    reach = np.stack([np.ravel_multi_index(tuple(move(geom, cells, mv[None, :]).T), (H, W)) for mv in action_set], axis=1)   # :111-122
    model = POMDP(states=[[0] * W for _ in range(H)], actions=len(action_set), observations=O,
                  reachable_states_table=reach[:, :, None], observation_table=obs,
                  end_states=np.flatnonzero(disc.ravel()).tolist(), start_probabilities=start.ravel())
    return dict(model=model, geom=geom, data=data, thresholds=thr, action_set=action_set, start=start)


def move(geom: dict, pos: np.ndarray, movement: np.ndarray) -> np.ndarray:
    'environment.py:711-769'
    new = pos + movement
    for axis, size, wrap in ((0, geom['H'], geom['boundary'] in ('wrap', 'wrap_vertical')),
                             (1, geom['W'], geom['boundary'] in ('wrap', 'wrap_horizontal'))):
        if wrap:
            new[:, axis] = np.where(new[:, axis] < 0, new[:, axis] + size, new[:, axis])
            new[:, axis] = np.where(new[:, axis] >= size, new[:, axis] - size, new[:, axis])
        else:
            new[:, axis] = np.clip(new[:, axis], 0, size - 1)
    return new


def get_observation(geom: dict, data: np.ndarray, pos: np.ndarray, time_idx: np.ndarray) -> np.ndarray:
    '''
    environment.py:554-652, including its array-time indexing: `np.where(time == unique_times[:, None])[0]`
    is in sorted-time order, i.e. agent i reads the frame of the i-th smallest time (:598-601,649-651).
    '''
    t = np.sort(np.asarray(time_idx) % geom['T'])
    dp = pos - np.array([geom['my'], geom['mx']])[None, :]
    ok = np.all((dp >= 0) & (dp < np.array([geom['h'], geom['w']])[None, :]), axis=1)
    out = np.zeros(len(pos))
    out[ok] = data[t[ok], dp[ok, 0], dp[ok, 1]]
    return out


def source_reached(geom: dict, pos: np.ndarray) -> np.ndarray:
    'environment.py:655-680'
    return np.sum((pos - geom['source'][None, :]) ** 2, axis=-1) <= geom['r2']


def discretize(thresholds: np.ndarray, odor: np.ndarray, reached: np.ndarray) -> np.ndarray:
    'agent.py:401-439 (non layered, not space aware)'
    ids = np.argwhere((odor[:, None] >= thresholds[:-1][None, :]) & (odor[:, None] < thresholds[1:][None, :]))[:, 1]
    ids[reached] = len(thresholds) - 1
    return ids


def run_test(setup: dict, value_function: ValueFunction, start_points: np.ndarray, time_shift: np.ndarray, horizon: int,
             infotaxis: bool = False) -> dict:
    '''
    simulation.py:1452-1497 with the PBVI_Agent glue (agents/pbvi_agent.py:729-811) or, when
    `infotaxis`, Infotaxis_Agent.choose_action (agents/infotaxis_agent.py:244-301).
    Returns per-step records and the count of agent-steps performed.
    '''
    model, geom, data, thr, action_set = setup['model'], setup['geom'], setup['data'], setup['thresholds'], setup['action_set']
    n = len(start_points)
    pos, tshift = np.array(start_points), np.array(time_shift)
    belief = BeliefSet(model, np.tile(model.start_probabilities, (n, 1)))
    running = np.arange(n)
    done_at, reached_src = np.full(n, -1), np.zeros(n, dtype=bool)
    agent_steps, records = 0, []
    for i in range(horizon):
        if infotaxis:
            act = infotaxis_choose(model, belief)
        else:
            _, act = value_function.evaluate_at(belief)                                    # pbvi_agent.py:741
        pos = move(geom, pos, action_set[act])                                             # simulation.py:1457
        odor = get_observation(geom, data, pos, tshift + i)                                # :1461
        reached = source_reached(geom, pos)                                                # :1466
        ids = discretize(thr, odor, reached)                                               # pbvi_agent.py:780
        belief, prov = belief.update(act, ids, raise_on_impossible_belief=False, return_provenance=True)   # :783-787
        ok = np.isin(np.arange(len(act)), prov[:, 0])                                      # intended length-N mask
        term = reached | ~ok                                                               # simulation.py:1480-1483 (time_loop=True)
        agent_steps += len(act)
        records.append((running.copy(), act.copy(), pos.copy(), odor.copy()))
        done_at[running[term]] = i + 1
        reached_src[running[reached]] = True
        keep_ok = ~term[ok]                                                                # pbvi_agent.py:810-811
        belief = BeliefSet(model, belief.belief_array[keep_ok])
        pos, tshift, running = pos[~term], tshift[~term], running[~term]
        if len(pos) == 0:
            break
    return dict(done_at_step=done_at, reached_source=reached_src, records=records, agent_steps=agent_steps)


def infotaxis_choose(model: POMDP, belief: BeliefSet) -> np.ndarray:
    'agents/infotaxis_agent.py:257-293 without materialising the successors (same numbers).'
    Hb = belief.entropies
    u, p = belief.unnormalised_successors()
    with np.errstate(divide='ignore', invalid='ignore'):
        w = np.where(p[..., None] > 0, u / p[..., None], 0.0)
        Hao = -np.sum(np.where(w > 0, w * np.log(w), 0.0), axis=3)
    dH = Hb[:, None] - np.sum(p * Hao, axis=2)
    return np.argmax(dH, axis=1)


def timed_agent_steps(setup: dict, value_function: ValueFunction, n: int, steps: int, seed: int = 0) -> dict:
    'CPU baseline sample: n agents for `steps` steps from random start points; returns agent-steps/s.'
    rng = np.random.default_rng(seed)
    S = setup['model'].state_count
    states = rng.choice(S, size=n, p=setup['model'].start_probabilities)
    starts = np.array(np.unravel_index(states, (setup['geom']['H'], setup['geom']['W']))).T
    tshift = rng.integers(0, setup['geom']['T'], n)
    t0 = time.perf_counter()
    out = run_test(setup, value_function, starts, tshift, steps)
    dt = time.perf_counter() - t0
    return dict(agent_steps=out['agent_steps'], seconds=dt, per_s=out['agent_steps'] / dt)
