'''
Environment — host-side mirror of the reference's `olfactory_navigation.Environment`
(reference olfactory_navigation/environment.py) for the part of it that sits on the hot path:
geometry (margins / shape / data bounds / source, :277-377), start probabilities (:383-423),
`move` (:711-769), `get_observation` (:554-652), `source_reached` (:655-680) and
`random_start_points` (:683-708).  The per-step functions also exist on the device
(`olfnav_env_step`, include/olfnav_b200.h); the NumPy versions here are the host API the
reference exposes and are what `exact_converter` uses to build the move table.

Out of scope (raise NotImplementedError): h5 files, layers, `shape=` /
`multiplier=` resampling, save/load, plotting, modify()/modify_scale().
'''
from __future__ import annotations

import ctypes

import numpy as np

from .util import get_rng

_BOUNDARIES = ('stop', 'clip', 'wrap', 'wrap_vertical', 'wrap_horizontal')


class Environment:
    def __init__(self,
                 data_file: str | np.ndarray,
                 data_source_position: list | np.ndarray,
                 source_radius: float = 1.0,
                 layers: bool | list[str] = False,
                 shape: list | np.ndarray = None,
                 margins: int | list | np.ndarray = 0,
                 multiplier: list | np.ndarray = [1.0, 1.0],
                 interpolation_method: str = 'Linear',
                 preprocess_data: bool = False,
                 boundary_condition: str = 'stop',
                 start_zone: str | np.ndarray = 'data_zone',
                 odor_present_threshold: float = None,
                 name: str = None) -> None:
        if layers not in (False, None):
            raise NotImplementedError('layered environments are out of scope of the B200 hot path (SURVEY §8 f4)')
        if shape is not None or not np.allclose(np.asarray(multiplier, dtype=float), 1.0):
            raise NotImplementedError('shape= / multiplier= resampling is out of scope of the B200 hot path (SURVEY §8 f3)')

        if isinstance(data_file, str):
            if not data_file.endswith('.npy'):
                raise NotImplementedError('only .npy files or in-memory arrays are supported')
            self.data_file_path = data_file
            data = np.load(data_file)
        elif isinstance(data_file, np.ndarray):
            self.data_file_path = None
            data = data_file
        else:
            raise NotImplementedError('data_file must be a .npy path or a numpy array')
        assert data.ndim == 3, 'data must have shape (time, y, x)'

        self._data = np.ascontiguousarray(data, dtype=np.float64)
        self.saved_at = None
        self.layers, self.layer_labels, self.has_layers = None, None, False
        self.timesteps = int(self._data.shape[0])
        self.data_shape = tuple(int(v) for v in self._data.shape[1:])
        self.dimensions = 2
        self.data_source_position = np.array(data_source_position)
        self.original_data_source_position = self.data_source_position
        self.interpolation_method = interpolation_method
        self.data_processed = True

        # margins -> (dims, 2) array of (low, high)
        if isinstance(margins, (int, np.integer)):
            self.margins = np.full((2, 2), int(margins), dtype=int)
        else:
            m = np.array(margins)
            if m.shape == (2,):
                self.margins = np.stack([m, m], axis=1)
            elif m.shape == (2, 2):
                self.margins = m.copy()
            else:
                raise ValueError('The array or lists of Margins provided have a shape not supported. (Supported formats (2,) or (2,2))')
        assert self.margins.dtype.kind == 'i', 'margins should be integers'

        self.shape = tuple(int(d + lo + hi) for d, (lo, hi) in zip(self.data_shape, self.margins))
        self.data_bounds = np.stack([self.margins[:, 0], self.margins[:, 0] + np.array(self.data_shape)], axis=1)
        self.source_position = self.data_source_position + self.margins[:, 0]
        self.source_radius = source_radius

        assert boundary_condition in _BOUNDARIES + ('no',), f'unknown boundary condition {boundary_condition}'
        self.boundary_condition = boundary_condition

        # start probabilities (environment.py:383-423)
        inner = tuple(slice(int(lo), int(hi)) for lo, hi in self.data_bounds)
        probs = np.zeros(self.shape)
        self.start_type = start_zone if isinstance(start_zone, str) else 'custom'
        if isinstance(start_zone, np.ndarray):
            if start_zone.shape == (2, 2):
                probs[tuple(slice(int(lo), int(hi)) for lo, hi in start_zone)] = 1.0
                self.start_type += '_' + '_'.join(str(v) for v in start_zone.ravel())
            elif start_zone.shape == self.shape:
                probs = start_zone.astype(float).copy()
            else:
                raise ValueError('If an np.ndarray is provided for the start_zone it has to be |dim| x 2...')
        elif start_zone == 'data_zone':
            probs[inner] = 1.0
        elif start_zone == 'odor_present':
            thr = 0 if odor_present_threshold is None else odor_present_threshold
            probs[inner] = (np.mean((self._data > thr).astype(int), axis=0) > 0).astype(float)
        else:
            raise ValueError('start_zone value is wrong')
        self.odor_present_threshold = odor_present_threshold
        probs[self._source_mask()] = 0.0
        self.start_probabilities = probs / np.sum(probs)

        if name is None:
            name = '_'.join(str(v) for v in self.shape)
            name += '-marg_' + '_'.join('_'.join(str(v) for v in row) for row in self.margins)
            name += f'-edge_{self.boundary_condition}-start_{self.start_type}'
            name += '-source_' + '_'.join(str(v) for v in self.source_position) + f'_radius{self.source_radius}'
        self.name = name

        self.is_on_gpu = False
        self._device_handles = {}
        # Reference quirk (environment.py:598-601,649-651): with an ARRAY of times the reference indexes
        # `np.where(time == unique_times[:, None])[0]`, which is in sorted-time order, so agent i observes
        # the frame of the i-th smallest time.  Invisible for a scalar time_shift.  True reproduces it
        # (drop-in parity, the default); False gives every agent its own time.
        self.reference_time_quirk = True

    # ---------------------------------------------------------------------------------------------
    def _source_mask(self) -> np.ndarray:
        yy, xx = np.meshgrid(np.arange(self.shape[0]), np.arange(self.shape[1]), indexing='ij')
        return ((yy - self.source_position[0]) ** 2 + (xx - self.source_position[1]) ** 2) <= self.source_radius ** 2

    @property
    def data(self) -> np.ndarray:
        return self._data

    @property
    def _data_is_numpy(self) -> bool:
        return True

    @property
    def on_gpu(self) -> 'Environment':
        return self

    @property
    def on_cpu(self) -> 'Environment':
        return self

    # -- per-step host API (NumPy) ---------------------------------------------------------------
    def get_observation(self, pos: np.ndarray, time: int | np.ndarray = 0, layer: int | np.ndarray = 0):
        'data[time % T, pos - low margins], 0.0 outside the data zone (environment.py:554-652).'
        pos = np.asarray(pos)
        single = pos.ndim == 1
        p = pos[None, :] if single else pos
        t = np.broadcast_to(np.asarray(time) % self.timesteps, (len(p),))
        if self.reference_time_quirk and np.ndim(time) > 0:
            t = np.sort(t)
        dp = p - self.margins[:, 0][None, :]
        inside = np.all((dp >= 0) & (dp < np.array(self.data_shape)[None, :]), axis=1)
        out = np.zeros(len(p), dtype=float)
        out[inside] = self._data[t[inside], dp[inside, 0], dp[inside, 1]]
        return float(out[0]) if single else out

    def source_reached(self, pos: np.ndarray):
        'Squared distance to the source centre <= radius^2 (environment.py:655-680).'
        pos = np.asarray(pos)
        single = pos.ndim == 1
        p = pos[None, :] if single else pos
        hit = np.sum((p - self.source_position[None, :]) ** 2, axis=-1) <= self.source_radius ** 2
        return bool(hit[0]) if single else hit

    def random_start_points(self, n: int = 1, rng=None) -> np.ndarray:
        'n start cells drawn from start_probabilities (environment.py:683-708); same RNG consumption.'
        assert n > 0, 'n has to be a strictly positive number (>0)'
        rng = get_rng(rng)
        states = rng.choice(np.arange(int(np.prod(self.shape))), size=n, replace=True, p=self.start_probabilities.ravel())
        return np.array(np.unravel_index(states, self.shape)).T

    def move(self, pos: np.ndarray, movement: np.ndarray) -> np.ndarray:
        'pos + movement under the boundary condition (environment.py:711-769).'
        pos = np.asarray(pos)
        single = pos.ndim == 1
        new = (pos + movement).reshape(-1, 2).copy()
        h, w = self.shape
        wrap_y = self.boundary_condition in ('wrap', 'wrap_vertical')
        wrap_x = self.boundary_condition in ('wrap', 'wrap_horizontal')
        if self.boundary_condition in _BOUNDARIES:
            for axis, size, wrap in ((0, h, wrap_y), (1, w, wrap_x)):
                if wrap:
                    new[new[:, axis] < 0, axis] += size
                    new[new[:, axis] >= size, axis] -= size
                else:
                    new[:, axis] = np.clip(new[:, axis], 0, size - 1)
        return new[0] if single else new

    def distance_to_source(self, point: np.ndarray, metric: str = 'manhattan'):
        point = np.asarray(point)
        single = point.ndim == 1
        p = point[None, :] if single else point
        if metric != 'manhattan':
            raise NotImplementedError('This distance metric has not yet been implemented')
        d = np.sum(np.abs(self.source_position[None, :] - p), axis=-1) - self.source_radius
        return float(d[0]) if single else d

    # -- device side -----------------------------------------------------------------------------
    def device_handle(self, device: int):
        'olfnav_env handle on `device` (created on first use, cached).'
        if device not in self._device_handles:
            from . import _lib
            _lib.require_cuda()
            h = ctypes.c_void_p()
            T, dh, dw = self._data.shape
            _lib.check(_lib.lib.olfnav_env_create(ctypes.byref(h), device, self.shape[0], self.shape[1],
                                                  _lib.BOUNDARY.get(self.boundary_condition, 0), _lib.ptr(self._data), T, dh, dw,
                                                  int(self.margins[0, 0]), int(self.margins[1, 0]),
                                                  float(self.source_position[0]), float(self.source_position[1]),
                                                  float(self.source_radius)))
            self._device_handles[device] = _EnvHandle(h)
        return self._device_handles[device].h

    def __getstate__(self):
        state = self.__dict__.copy()
        state['_device_handles'] = {}
        return state


class _EnvHandle:
    def __init__(self, h):
        self.h = h

    def __del__(self):
        try:
            from . import _lib
            _lib.lib.olfnav_env_destroy(self.h)
        except Exception:
            pass
