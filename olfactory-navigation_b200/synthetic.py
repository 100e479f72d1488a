'''
Seeded synthetic plume environments of the shapes named in BASELINE.json.

Generator precedent in the reference: Gaussian-probability Bernoulli frames in
experiments/sea-robins/minimal_model.ipynb cells 2-7.  Detection probability
    p(y, x) = p0 * exp(-(y - y_c)^2 / (2 sigma(x)^2)) * exp(-x / L),  sigma(x) = sigma0 + k sqrt(x)
and frames c[t, y, x] = Bernoulli(p) * U(0, 1).  T is a power of two so that time-averaged bin
probabilities sum to exactly 1.0 (the reference asserts that with `==`,
agents/model_based_util/environment_converter.py:97).
'''
from __future__ import annotations

import numpy as np

SEED = 12131415  # borrowed from the reference's only environment metadata (experiments/Env-55_247-.../METADATA.json:53)

# name -> (T, h, w, margins [y, x], source data position [y, x], radius, boundary, start_zone, thresholds)
CONFIGS = {
    'C1': dict(T=64, h=30, w=30, margins=[10, 10], source=[15, 0], radius=2.0, boundary='stop',
               start_zone='data_zone', thresholds=[0.25, 0.5, 0.75]),
    'C2': dict(T=256, h=55, w=247, margins=[14, 62], source=[27, 0], radius=2.0, boundary='wrap_vertical',
               start_zone='odor_present', thresholds=0.5),
    'C5': dict(T=256, h=110, w=494, margins=[28, 124], source=[55, 0], radius=4.0, boundary='wrap_vertical',
               start_zone='odor_present', thresholds=0.5),
    # tiny shapes for fast parity tests
    'T0': dict(T=16, h=9, w=13, margins=[3, 4], source=[4, 0], radius=1.0, boundary='stop',
               start_zone='data_zone', thresholds=0.5),
    'T1': dict(T=16, h=9, w=13, margins=[3, 4], source=[4, 0], radius=1.0, boundary='wrap_vertical',
               start_zone='odor_present', thresholds=[0.3, 0.6]),
    'T2': dict(T=16, h=8, w=11, margins=[2, 3], source=[4, 1], radius=1.0, boundary='wrap',
               start_zone='data_zone', thresholds=0.5),
    'T3': dict(T=16, h=8, w=11, margins=[2, 3], source=[4, 1], radius=1.0, boundary='wrap_horizontal',
               start_zone='data_zone', thresholds=0.5),
}


def plume_frames(T: int, h: int, w: int, source_y: float, seed: int = SEED,
                 p0: float = 0.9, sigma0: float = 1.0, k: float = 0.45, L: float = None) -> np.ndarray:
    'Odor frames (T, h, w) float64 in [0, 1); the source sits on the left edge at row source_y.'
    rng = np.random.default_rng(seed)
    L = 0.8 * w if L is None else L
    y = np.arange(h, dtype=float)[:, None]
    x = np.arange(w, dtype=float)[None, :]
    sigma = sigma0 + k * np.sqrt(x)
    p = p0 * np.exp(-((y - source_y) ** 2) / (2.0 * sigma ** 2)) * np.exp(-x / L)
    hits = rng.random((T, h, w)) < p[None, :, :]
    return hits * rng.random((T, h, w))


def config(name: str) -> dict:
    'Constructor kwargs for an Environment (reference olfactory_navigation.Environment or ours).'
    c = CONFIGS[name]
    data = plume_frames(c['T'], c['h'], c['w'], source_y=c['source'][0], seed=SEED)
    kwargs = dict(data_file=data,
                  data_source_position=list(c['source']),
                  source_radius=c['radius'],
                  margins=list(c['margins']),
                  boundary_condition=c['boundary'],
                  start_zone=c['start_zone'],
                  odor_present_threshold=(0.0 if c['start_zone'] == 'odor_present' else None))
    return dict(env_kwargs=kwargs, thresholds=c['thresholds'])


def layered_config(name: str = 'L1') -> dict:
    '''
    A tiny LAYERED environment (reference environment.py:202-235): two layers over the T1 geometry — a narrow plume
    ("ground") and a wider, weaker one ("air").  thresholds: shared by the layers; thresholds_per_layer: one row per layer.
    '''
    assert name == 'L1'
    c = CONFIGS['T1']
    ground = plume_frames(c['T'], c['h'], c['w'], source_y=c['source'][0], seed=SEED)
    air = plume_frames(c['T'], c['h'], c['w'], source_y=c['source'][0], seed=SEED + 7, p0=0.7, sigma0=2.0, k=0.8)
    kwargs = dict(data_file=np.stack([ground, air]), data_source_position=list(c['source']), source_radius=c['radius'],
                  layers=['ground', 'air'], margins=list(c['margins']), boundary_condition=c['boundary'],
                  start_zone='data_zone')        # 'odor_present' breaks the reference's constructor on layered arrays (environment.py:404)
    return dict(env_kwargs=kwargs, thresholds=[0.3, 0.6], thresholds_per_layer=np.array([[0.3, 0.6], [0.2, 0.5]]))
