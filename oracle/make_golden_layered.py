'''
ORACLE — TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*_L1*.npz: LAYERED environment.

Runs in THIS container only (imports the UNMODIFIED reference from /root/reference through oracle/refloader.py on top of the
restated `pomdp_toolkit`).  Pinned by the reference's own code: Environment(layers=...), Agent action set / labels with layers,
exact_converter with per-action observation tables (shared thresholds, and per-layer thresholds set on the agent object:
the reference's dict-threshold constructor branch is broken), move / get_observation(layer=action[:, 0]) /
discretize_observations, and run_test of QMDP / FSVI agents.

    python oracle/make_golden_layered.py
'''
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import refloader  # noqa: E402

on = refloader.load()
from olfactory_navigation import Environment, run_test            # noqa: E402
from olfactory_navigation.agents import FSVI_Agent, QMDP_Agent     # noqa: E402
from olfactory_navigation.agents.model_based_util.environment_converter import exact_converter   # noqa: E402

import olfactory_navigation_b200.synthetic as syn                  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def save(name, **arrays):
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **arrays)
    print(f'{name:28s} {os.path.getsize(path) / 1024:8.1f} KiB')


c = syn.layered_config('L1')
env = Environment(**c['env_kwargs'])
rng = np.random.default_rng(syn.SEED + 77)
H, W = env.shape

for tag, per_layer in (('L1', False), ('L1pl', True)):
    ag = QMDP_Agent(env, thresholds=c['thresholds'])
    if per_layer:
        thr = np.array([[-np.inf, *row, np.inf] for row in c['thresholds_per_layer']])
        ag.thresholds = thr
        ag.model = exact_converter(ag)
    model = ag.model
    A = model.action_count
    save(f'converter_{tag}', shape=np.array(env.shape), thresholds=ag.thresholds, action_set=ag.action_set,
         action_labels=np.array(ag.action_labels), layer_labels=np.array(env.layer_labels),
         reachable_states=model.reachable_states[:, :, 0].astype(np.int32), observation_table=model.observation_table,
         end_states=model.end_states.astype(np.int32), start_probabilities=model.start_probabilities,
         source_position=env.source_position, margins=env.margins)

    n = 200
    pos = np.stack([rng.integers(0, H, n), rng.integers(0, W, n)], axis=1)
    act = rng.integers(0, A, n)
    tsh = rng.integers(0, 3 * env.timesteps, n)
    step = 5
    vec = ag.action_set[act]
    new_pos = env.move(pos, vec[:, 1:])
    odor = env.get_observation(new_pos, time=(tsh + step), layer=vec[:, 0])
    reached = env.source_reached(new_pos)
    ids = ag.discretize_observations(observation=odor, action=vec, source_reached=reached)
    save(f'envstep_{tag}', pos=pos.astype(np.int32), act=act.astype(np.int32), time_shift=tsh.astype(np.int64), step=np.array(step),
         new_pos=new_pos.astype(np.int32), odor=odor, reached=reached, obs_id=ids.astype(np.int32))

# run_test with shared thresholds (the reference's working path)
n_sim, horizon = 48, 60
for label, cls, train_kwargs in (('qmdp', QMDP_Agent, dict(expansions=1000)), ('fsvi', FSVI_Agent, dict(expansions=8, max_belief_growth=12))):
    agent = cls(env, thresholds=c['thresholds'])
    agent.rng = np.random.default_rng(syn.SEED + 1)
    agent.train(print_progress=False, print_stats=False, **train_kwargs)
    sim_rng = np.random.default_rng(syn.SEED + 2)
    starts = env.random_start_points(n_sim, sim_rng)
    tshift = sim_rng.integers(0, env.timesteps, n_sim)
    hist = run_test(agent, n=n_sim, start_points=starts, time_shift=tshift, horizon=horizon, print_progress=False,
                    print_stats=False, use_gpu=False, batches=1)
    vfa = agent.value_function
    save(f'runtest_{label}_L1', alphas=vfa.alpha_vector_array.astype(np.float32), alpha_actions=vfa.actions.astype(np.int32),
         start_points=starts.astype(np.int32), time_shift=tshift.astype(np.int64), horizon=np.array(horizon),
         actions=np.array(hist.actions).astype(np.int8), positions=np.array(hist.positions).astype(np.int16),
         observations=np.array(hist.observations).astype(np.float32), reached_source=hist.reached_source,
         done_at_step=hist.done_at_step.astype(np.int32))
print('done')
