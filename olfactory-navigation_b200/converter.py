'''
exact_converter — environment -> grid POMDP, one state per cell.  Same tables as the reference's
exact_converter (reference agents/model_based_util/environment_converter.py:7-134):
  * end states: cells within the source radius (:44-45);
  * observation_table[s, a, o]: time-average of the odor-bin indicator inside the data zone,
    'nothing' = 1 in the margins (:66-89), goal observation = 1 on end states and every other
    observation 0 there (:91-94); rows must sum to exactly 1.0 (:97);
  * reachable_states_table from environment.move on every cell (:111-122);
  * start probabilities = environment.start_probabilities.ravel() (:132).
'''
from __future__ import annotations

import numpy as np

from .pomdp import POMDP


def exact_converter(agent) -> POMDP:
    env = agent.environment
    thr = agent.thresholds
    action_set = np.asarray(agent.action_set)
    moves = action_set[:, 1:] if env.has_layers else action_set          # the physical part of a layered action (:113)
    assert env.dimensions == 2, 'This converter only works for 2D environments...'
    assert not agent.space_aware, 'This model is not compatible with space_aware agents as the observations would not match the expected observations of this model.'
    H, W = env.shape
    S, A, O = H * W, len(moves), thr.shape[-1]          # bins (= len(thr) - 1) + goal

    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing='ij')
    end_mask = (((yy - env.source_position[0]) ** 2 + (xx - env.source_position[1]) ** 2) <= env.source_radius ** 2).ravel()
    end_states = np.flatnonzero(end_mask).tolist()

    # time-averaged probability of each odor bin inside the data zone (one field per layer, thresholds possibly per layer, :54-65)
    (y0, y1), (x0, x1) = env.data_bounds

    def field(frames, t):
        data = frames[:, :, :, None]
        in_bin = (data >= t[:-1][None, None, None, :]) & (data < t[1:][None, None, None, :])
        f = np.zeros((H, W, O - 1))
        f[y0:y1, x0:x1, :] = np.average(in_bin, axis=0)
        return f

    obs = np.empty((S, A, O))
    if env.has_layers:
        fields = [field(env.data[layer], thr[layer] if thr.ndim == 2 else thr) for layer in env.layers]
        for a in range(A):
            obs[:, a, :O - 1] = fields[int(action_set[a, 0])].reshape(S, O - 1)
    else:
        obs[:, :, :O - 1] = field(env.data, thr).reshape(S, 1, O - 1)
    margin = np.ones((H, W), dtype=bool)
    margin[y0:y1, x0:x1] = False
    obs[margin.ravel(), :, 0] = 1.0
    obs[:, :, O - 1] = 0.0
    obs[end_states, :, :] = 0.0
    obs[end_states, :, O - 1] = 1.0
    assert np.all(np.sum(obs, axis=2) == 1.0), 'Observation table malformed, something is wrong...'

    labels = ['nothing'] + ([f'something_l{i}' for i in range(O - 2)] if O > 3 else ['something']) + ['goal']

    cells = np.stack([yy.ravel(), xx.ravel()], axis=1)
    reach = np.stack([np.ravel_multi_index(tuple(env.move(cells, mv[None, :]).T), (H, W)) for mv in moves], axis=1)

    state_grid = [[f's_{x}_{y}' for x in range(W)] for y in range(H)]
    return POMDP(states=state_grid, actions=agent.action_labels, observations=labels,
                 reachable_states_table=reach[:, :, None], observation_table=obs, end_states=end_states,
                 start_probabilities=env.start_probabilities.ravel(),
                 grid_shape=(H, W), moves=moves,
                 boundary=('stop' if env.boundary_condition == 'clip' else env.boundary_condition))


def minimal_converter(agent, partitions=[3, 6], margin_partitions: bool = False) -> POMDP:
    '''
    A handful of partition cells + one goal state with stochastic transitions between them — same tables as the reference's
    minimal_converter (reference agents/model_based_util/environment_converter.py:137-292):
      * cells: `partitions` (y, x) cells over the data zone plus one ring of margin cells per axis (margin_partitions=True), or over
        the whole grid (False; that branch of the reference cannot run: its observation block reshapes to partitions + 2, :264);
      * T[cell, a, cell'] = fraction of the cell's grid points that environment.move sends into cell' without reaching the source,
        T[cell, a, goal] = fraction that reaches it; the goal state is absorbing (:207-224);
      * O: 'nothing' = 1 everywhere, data cells get the time-and-space average of each odor bin (strict inequalities, :250-255),
        the goal state observes 'goal' (:262-275);  uniform start probabilities (:277-279).
    '''
    env = agent.environment
    thr = agent.thresholds
    action_set = np.asarray(agent.action_set)
    assert not agent.space_aware, 'This model is not compatible with space_aware agents as the observations would not match the expected observations of this model.'
    assert thr.ndim == 1, 'This model doesnt allow for layer specific thresholds.'
    shape = np.array(env.shape)
    partitions = np.array(partitions)
    if margin_partitions:
        cell_shape = (np.array(env.data_shape) / partitions).astype(int)
        bounds = [np.array([0, *((np.arange(n + 1) * cell_shape[ax]) + env.margins[ax, 0]), shape[ax]]) for ax, n in enumerate(partitions)]
        cell_counts = int(np.prod(partitions + 2))
    else:
        cell_shape = (shape / partitions).astype(int)
        bounds = [np.arange(n + 1) * cell_shape[ax] for ax, n in enumerate(partitions)]
        cell_counts = int(np.prod(partitions))
    # cell ids follow the reference's meshgrid(indexing='xy') order: the FIRST axis varies fastest
    grid_cells = np.full(env.shape, -1)
    cell = 0
    for x0, x1 in zip(bounds[1][:-1], bounds[1][1:]):
        for y0, y1 in zip(bounds[0][:-1], bounds[0][1:]):
            grid_cells[y0:y1, x0:x1] = cell
            cell += 1

    S, A = cell_counts + 1, len(action_set)
    yy, xx = np.meshgrid(np.arange(env.shape[0]), np.arange(env.shape[1]), indexing='ij')
    points = np.stack([yy.ravel(), xx.ravel()], axis=1)
    point_cell = grid_cells[points[:, 0], points[:, 1]]
    moves = action_set[:, 1:] if env.has_layers else action_set
    T = np.full((S, A, S), -1.0)
    for a, mv in enumerate(moves):
        new_points = env.move(points, mv[None, :])
        new_cell = grid_cells[new_points[:, 0], new_points[:, 1]]
        at_source = env.source_reached(new_points)
        for c in range(cell_counts):
            inside = point_cell == c
            count = np.sum(inside)
            with np.errstate(divide='ignore', invalid='ignore'):
                T[c, a, :cell_counts] = np.bincount(new_cell[inside & ~at_source & (new_cell >= 0)], minlength=cell_counts) / count
                T[c, a, -1] = np.sum(inside & at_source) / count
    T[-1, :, :] = 0.0
    T[-1, :, -1] = 1.0

    O = len(thr)                                   # bins + goal
    labels = ['nothing'] + ([f'something_l{i}' for i in range(O - 2)] if O > 3 else ['something']) + ['goal']
    obs = np.zeros((S, A, O))
    obs[:-1, :, 0] = 1.0
    data_bounds = [np.arange(n + 1) * cell_shape[ax] for ax, n in enumerate(partitions)]
    cell_obs = []
    for x0, x1 in zip(data_bounds[1][:-1], data_bounds[1][1:]):
        for y0, y1 in zip(data_bounds[0][:-1], data_bounds[0][1:]):
            levels = []
            for lo, hi in zip(thr[:-1], thr[1:]):
                if env.has_layers:
                    block = env.data[:, :, y0:y1, x0:x1]
                    levels.append(np.average((block > lo) & (block < hi), axis=(1, 2, 3)))
                else:
                    block = env.data[:, y0:y1, x0:x1]
                    levels.append(np.average((block > lo) & (block < hi)))
            cell_obs.append(levels)
    data_cell_ids = np.arange(cell_counts).reshape(partitions + 2)[1:-1, 1:-1].ravel()
    for i, cid in enumerate(data_cell_ids):
        if env.has_layers:
            layers = action_set[:, 0].astype(int)
            for o in range(O - 1):
                obs[cid, :, o] = np.asarray(cell_obs[i][o])[layers]
        else:
            obs[cid, :, :O - 1] = cell_obs[i]
    obs[-1, :, -1] = 1.0

    start = np.ones(S) / S
    return POMDP(states=[f'cell_{c}' for c in range(cell_counts)] + ['goal'], actions=agent.action_labels, observations=labels,
                 transition_table=T, observation_table=obs, end_states=[cell_counts], start_probabilities=start)
