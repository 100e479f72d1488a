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
