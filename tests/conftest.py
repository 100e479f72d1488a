import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
TINY = ['T0', 'T1', 'T2', 'T3']
ALL_CFGS = TINY + ['C1']


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def golden(name: str):
    return np.load(os.path.join(GOLDEN, name + '.npz'))


@pytest.fixture(scope='session')
def onb():
    import olfactory_navigation_b200
    return olfactory_navigation_b200


_CACHE = {}


def make(cfg: str, agent_cls_name: str = 'QMDP_Agent'):
    'Environment + agent of our package for a synthetic config (cached per process).'
    key = (cfg, agent_cls_name)
    if key not in _CACHE:
        import olfactory_navigation_b200 as onb
        c = onb.synthetic.config(cfg)
        env = onb.Environment(**c['env_kwargs'])
        agent = getattr(onb.agents, agent_cls_name)(env, thresholds=c['thresholds'])
        _CACHE[key] = (env, agent)
    return _CACHE[key]


def oracle_model(model):
    'The oracle POMDP with the same tables as one of our models.'
    from oracle import refloader
    refloader.install_oracle_toolkit()
    import pomdp_toolkit
    H, W = model.grid_shape
    return pomdp_toolkit.POMDP(states=[[f's_{x}_{y}' for x in range(W)] for y in range(H)], actions=model.action_labels,
                               observations=model.observation_labels, reachable_states_table=model.reachable_states,
                               observation_table=model.observation_table, end_states=model.end_states.tolist(),
                               start_probabilities=model.start_probabilities)
