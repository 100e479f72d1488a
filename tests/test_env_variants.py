'''
SURVEY §8 f3 on the host: Environment(shape=, multiplier=) / modify / modify_scale, save / load and the alpha resampling of
PBVI_Agent.modify_environment against golden vectors frozen from the UNMODIFIED reference (oracle/make_golden_env.py).
'''
import os

import numpy as np
import pytest

from conftest import golden

VARIANTS = {
    'shape':      lambda e: e.modify(shape=[int(e.shape[0] * 1.4), int(e.shape[1] * 1.25)]),
    'multiplier': lambda e: e.modify(multiplier=[1.2, 0.8]),
    'scale':      lambda e: e.modify_scale(1.5),
    'shrink':     lambda e: e.modify_scale(0.75),
    'radius':     lambda e: e.modify(source_radius=e.source_radius + 1, boundary_condition='wrap'),
}


@pytest.mark.parametrize('cfg', ['T1', 'C1'])
@pytest.mark.parametrize('name', list(VARIANTS))
def test_modified_environment_matches_reference(cfg, name, onb):
    g = golden(f'envmod_{name}_{cfg}')
    c = onb.synthetic.config(cfg)
    mod = VARIANTS[name](onb.Environment(**c['env_kwargs']))
    assert tuple(mod.shape) == tuple(g['shape']) and tuple(mod.data_shape) == tuple(g['data_shape'])
    assert np.array_equal(mod.margins, g['margins'])
    assert np.array_equal(mod.source_position, g['source_position']) and np.array_equal(mod.data_source_position, g['data_source_position'])
    assert float(mod.source_radius) == float(g['source_radius'])
    np.testing.assert_allclose(mod.data, g['data'], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(mod.start_probabilities, g['start_probabilities'], rtol=1e-12, atol=0)
    np.testing.assert_allclose(mod.get_observation(g['pos'].astype(int), time=int(g['time'])), g['odor'], rtol=1e-12, atol=1e-15)


def test_save_load_round_trip(tmp_path, onb):
    c = onb.synthetic.config('T1')
    env = VARIANTS['scale'](onb.Environment(**c['env_kwargs']))
    env.save(folder=str(tmp_path), save_arrays=True)
    back = onb.Environment.load(env.saved_at)
    assert back.shape == env.shape and back.name == env.name and back.boundary_condition == env.boundary_condition
    assert np.array_equal(back.data, env.data) and np.array_equal(back.start_probabilities, env.start_probabilities)
    assert np.array_equal(back.margins, env.margins) and np.array_equal(back.source_position, env.source_position)
    pos = np.array([[3, 4], [10, 12]])
    assert np.array_equal(back.get_observation(pos, time=5), env.get_observation(pos, time=5))
    # from a data file: METADATA only, the environment is rebuilt (and resampled) from the file
    path = os.path.join(str(tmp_path), 'plume.npy')
    np.save(path, c['env_kwargs']['data_file'])
    kw = dict(c['env_kwargs'], data_file=path)
    env2 = onb.Environment(**kw).modify_scale(1.5)
    env2.save(folder=str(tmp_path), force=True)
    assert not os.path.exists(env2.saved_at + '/data.npy')
    back2 = onb.Environment.load(env2.saved_at)
    assert back2.shape == env2.shape and np.array_equal(back2.data, env2.data)
    with pytest.raises(Exception):
        env2.save(folder=str(tmp_path))            # not empty, no force


def test_alpha_resampling_matches_cv2_expression(onb):
    from olfactory_navigation_b200.agents.pbvi_agent import _resize_bilinear
    for name in ('scale', 'shape'):
        g = golden(f'envmod_{name}_T1')
        c = onb.synthetic.config('T1')
        env = onb.Environment(**c['env_kwargs'])
        H, W = env.shape
        H2, W2 = (int(v) for v in g['shape'])
        got = np.array([_resize_bilinear(a.reshape(H, W), (H2, W2)).ravel() for a in g['alphas']])
        np.testing.assert_allclose(got, g['alphas_modified'], rtol=1e-6, atol=1e-9)
