'''
ORACLE — TEST INFRASTRUCTURE ONLY.  Generates tests/golden/envmod_*.npz.

Runs in THIS container only (imports the UNMODIFIED reference from /root/reference through oracle/refloader.py) and freezes
what the reference's Environment(shape=, multiplier=) / modify / modify_scale (environment.py:293-377,1061-1161) and
PBVI_Agent.modify_environment (agents/pbvi_agent.py:661-695) produce on the synthetic plumes.

    python oracle/make_golden_env.py
'''
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import refloader  # noqa: E402

on = refloader.load()
from olfactory_navigation import Environment            # noqa: E402
from olfactory_navigation.agents import QMDP_Agent      # noqa: E402
import olfactory_navigation_b200.synthetic as syn       # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')

VARIANTS = {
    'shape':      lambda e: e.modify(shape=[int(e.shape[0] * 1.4), int(e.shape[1] * 1.25)]),
    'multiplier': lambda e: e.modify(multiplier=[1.2, 0.8]),
    'scale':      lambda e: e.modify_scale(1.5),
    'shrink':     lambda e: e.modify_scale(0.75),
    'radius':     lambda e: e.modify(source_radius=e.source_radius + 1, boundary_condition='wrap'),
}

for cfg in ['T1', 'C1']:
    c = syn.config(cfg)
    for name, make in VARIANTS.items():
        env = Environment(**c['env_kwargs'])
        mod = make(env)
        rng = np.random.default_rng(17)
        n = 64
        pos = np.stack([rng.integers(0, mod.shape[0], n), rng.integers(0, mod.shape[1], n)], axis=1)
        t = int(rng.integers(0, 3 * mod.timesteps))
        odor = mod.get_observation(pos, time=t)
        arrays = dict(shape=np.array(mod.shape), data_shape=np.array(mod.data_shape), margins=mod.margins,
                      source_position=mod.source_position, data_source_position=mod.data_source_position,
                      source_radius=np.array(mod.source_radius), start_probabilities=mod.start_probabilities,
                      pos=pos.astype(np.int32), time=np.array(t), odor=odor, data=np.asarray(mod.data).astype(np.float64))
        if name in ('scale', 'shape') and cfg == 'T1':
            ag = QMDP_Agent(Environment(**c['env_kwargs']), thresholds=c['thresholds'])
            ag.train(expansions=1000, print_progress=False, print_stats=False)
            # The reference's modify_environment itself raises (it hands self.thresholds back as an ndarray, which Agent.__init__
            # does not accept: agent.py:157-204 -> AttributeError).  Its resampling expression (agents/pbvi_agent.py:688-691) is
            # evaluated here verbatim on a freshly built agent of the modified environment.
            import cv2
            new_agent = QMDP_Agent(make(Environment(**c['env_kwargs'])), thresholds=c['thresholds'])
            vf = ag.value_function
            reshaped = np.array([cv2.resize(av, np.array(new_agent.model.state_grid.shape)[::-1]).ravel()
                                 for av in vf.alpha_vector_array.reshape(len(vf), *ag.model.state_grid.shape)])
            arrays.update(alphas=vf.alpha_vector_array, alphas_modified=reshaped, alpha_actions=np.asarray(vf.actions).astype(np.int32))
        path = os.path.join(OUT, f'envmod_{name}_{cfg}.npz')
        np.savez_compressed(path, **arrays)
        print(f'envmod_{name}_{cfg}: shape {mod.shape} data {mod.data_shape} margins {mod.margins.tolist()} {os.path.getsize(path) / 1024:.1f} KiB')
