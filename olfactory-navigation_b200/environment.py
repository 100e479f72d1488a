'''
Environment — host-side mirror of the reference's `olfactory_navigation.Environment`
(reference olfactory_navigation/environment.py) for the part of it that sits on the hot path:
geometry (margins / shape / data bounds / source, :277-377), start probabilities (:383-423),
`move` (:711-769), `get_observation` (:554-652), `source_reached` (:655-680) and
`random_start_points` (:683-708).  The per-step functions also exist on the device
(`olfnav_env_step`, include/olfnav_b200.h); the NumPy versions here are the host API the
reference exposes and are what `exact_converter` uses to build the move table.

shape= / multiplier= resampling, save / load and modify / modify_scale are host-side one-offs kept in
NumPy / SciPy.  Layered data (layer, time, y, x) is supported.  Out of scope (raise NotImplementedError): h5 files, `shape=` /
`multiplier=` resampling, save/load, plotting, modify()/modify_scale().
'''
from __future__ import annotations

import ctypes

import numpy as np

from .util import get_rng

_BOUNDARIES = ('stop', 'clip', 'wrap', 'wrap_vertical', 'wrap_horizontal')


def _resize_array(array: np.ndarray, new_shape: tuple, interpolation: str) -> np.ndarray:
    '''
    Resample a 2-D frame onto new_shape with endpoints aligned (index 0 -> 0, last -> last), the convention of the
    reference's _resize_array (environment.py:25-55, scipy RegularGridInterpolator over linspace indices).
    '''
    from scipy.interpolate import RegularGridInterpolator
    axes = [np.linspace(0, n - 1, n) for n in array.shape]
    new_axes = [np.linspace(0, n - 1, m) for n, m in zip(array.shape, new_shape)]
    grid = np.meshgrid(*new_axes, indexing='ij')
    return RegularGridInterpolator(tuple(axes), array, method=interpolation)(tuple(grid))


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
        # layers (environment.py:202-235): data of shape (layer, time, y, x); a list names the layers
        has_layers = isinstance(layers, (list, tuple)) or layers is True

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
        assert data.ndim == (4 if has_layers else 3), 'data must have shape (time, y, x), or (layer, time, y, x) with layers'

        self._data = np.ascontiguousarray(data, dtype=np.float64)
        self._original_data = self._data
        self._preprocess_data = preprocess_data
        self.saved_at = None
        self.layers, self.layer_labels, self.has_layers = None, None, has_layers
        if has_layers:
            self.layers = np.arange(len(self._data))
            if isinstance(layers, (list, tuple)):
                assert len(layers) == len(self._data), 'The amount of layers provided dont match the amount in the dataset.'
                self.layer_labels = [str(layer) for layer in layers]
            else:
                self.layer_labels = [str(layer) for layer in range(len(self._data))]
        self.timesteps = int(self._data.shape[-3])
        self.data_shape = tuple(int(v) for v in self._data.shape[-2:])
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

        # shape= / multiplier= (environment.py:293-337): the data zone is resampled to shape - margins, then scaled around the
        # source by `multiplier` (the margins absorb the difference), every frame by the same separable interpolation as the
        # reference's _resize_array (environment.py:25-55).  Done once here: the device path needs dense frames in HBM.
        original_data_shape = self.data_shape
        new_data_shape = None
        if shape is not None:
            shape = np.array(shape)
            assert np.all(shape > np.sum(self.margins, axis=1)), 'The shape of the environment must be strictly larger than the sum of margins.'
            new_data_shape = (shape - np.sum(self.margins, axis=1)).astype(int)
            self.data_source_position = (self.data_source_position * (new_data_shape / np.array(self.data_shape))).astype(int)
            self.data_shape = tuple(int(v) for v in new_data_shape)
        else:
            shape = np.array(self.data_shape) + np.sum(self.margins, axis=1)
        multiplier = np.array(multiplier, dtype=float)
        with np.errstate(divide='ignore', invalid='ignore'):
            low_max = (self.margins[:, 0] / self.data_source_position) + 1
            high_max = 1 + (self.margins[:, 1] / (np.array(self.data_shape) - self.data_source_position))
            max_mult = np.nan_to_num(np.min(np.vstack([low_max, high_max]), axis=0), nan=np.inf)   # 0 / 0: no margin AND source on the edge
        assert np.all(multiplier == 1.0) or np.all(multiplier <= max_mult), f'The multiplier given is larger than allowed (the values should be lower than {max_mult})'
        scaled_shape = (np.array(self.data_shape) * multiplier).astype(int)
        new_source = (self.data_source_position * multiplier).astype(int)
        self.margins = self.margins.copy()
        self.margins[:, 0] -= (new_source - self.data_source_position)
        self.margins[:, 1] = shape - (self.margins[:, 0] + scaled_shape)
        self.data_source_position = new_source
        self.data_shape = tuple(int(v) for v in scaled_shape)
        if self.data_shape != original_data_shape:
            frames = self._data.reshape((-1,) + self._data.shape[-2:])
            frames = np.stack([_resize_array(frame, self.data_shape, interpolation_method.lower()) for frame in frames])
            self._data = frames.reshape(self._data.shape[:-2] + self.data_shape)
        self.data_processed = True

        self.shape = tuple(int(d + lo + hi) for d, (lo, hi) in zip(self.data_shape, self.margins))
        self.data_bounds = np.stack([self.margins[:, 0], self.margins[:, 0] + np.array(self.data_shape)], axis=1)
        self.source_position = self.data_source_position + self.margins[:, 0]
        self.source_radius = source_radius

        assert boundary_condition in _BOUNDARIES + ('no',), f'unknown boundary condition {boundary_condition}'
        self.boundary_condition = boundary_condition

        # start probabilities (environment.py:383-423)
        inner = tuple(slice(int(lo), int(hi)) for lo, hi in self.data_bounds)
        probs = np.zeros(self.shape)
        if isinstance(start_zone, str) and start_zone.startswith('custom_'):       # a start_type string handed back by modify()
            start_zone = np.array(start_zone.split('_')[1:]).reshape((2, 2)).astype(int)
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
            frames0 = self._data[0] if self.has_layers else self._data       # layered: the first layer decides (environment.py:408)
            if self.data_shape != original_data_shape and not preprocess_data:
                # reference quirk (environment.py:402-413): on lazily resampled data the map is the FRACTION of frames with odor,
                # not the 0/1 indicator used for data that is already at its final shape
                probs[inner] = np.mean((frames0 > thr).astype(float), axis=0)
            else:
                probs[inner] = (np.mean((frames0 > thr).astype(int), axis=0) > 0).astype(float)
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
        'data[(layer,) time % T, pos - low margins], 0.0 outside the data zone (environment.py:554-652).'
        pos = np.asarray(pos)
        single = pos.ndim == 1
        p = pos[None, :] if single else pos
        t = np.broadcast_to(np.asarray(time) % self.timesteps, (len(p),))
        if self.reference_time_quirk and np.ndim(time) > 0:
            t = np.sort(t)
        dp = p - self.margins[:, 0][None, :]
        inside = np.all((dp >= 0) & (dp < np.array(self.data_shape)[None, :]), axis=1)
        out = np.zeros(len(p), dtype=float)
        if self.has_layers:
            lay = np.broadcast_to(np.asarray(layer), (len(p),))
            if self.reference_time_quirk and np.ndim(layer) > 0:
                lay = np.sort(lay)             # same indexing quirk as for the times (environment.py:594-595,645)
            out[inside] = self._data[lay[inside], t[inside], dp[inside, 0], dp[inside, 1]]
        else:
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

    # -- persistence and variants (environment.py:809-996, :1061-1161) ---------------------------------
    def save(self, folder: str = None, save_arrays: bool = False, force: bool = False) -> None:
        'METADATA.json (+ data.npy, start_probabilities.npy with save_arrays) in <folder>/Env-<name>; same keys as the reference.'
        import json, os, shutil
        assert save_arrays or ((self.data_file_path is not None) and (self.start_type is not None)), \
            "The environment was not created from a data file so 'save_arrays' has to be set to True."
        folder = f'./Env-{self.name}' if folder is None else folder + '/Env-' + self.name
        if not os.path.exists(folder):
            os.mkdir(folder)
        elif len(os.listdir(folder)) > 0:
            if not force:
                raise Exception(f'{folder} is not empty. If you want to overwrite the saved environment, enable "force".')
            shutil.rmtree(folder)
            os.mkdir(folder)
        if self.start_type.startswith('custom') and len(self.start_type.split('_')) == 1 and not save_arrays:
            raise Exception('Start probabilities have been set from a custom array, please enable save_arrays to be able to reconstruct the environment later.')
        arguments = {'name': self.name}
        if self.data_file_path is not None:
            arguments['data_file_path'] = self.data_file_path
        arguments.update(timesteps=int(self.timesteps), data_shape=list(self.data_shape), dimensions=self.dimensions,
                         margins=self.margins.tolist(), shape=list(self.shape), data_bounds=self.data_bounds.tolist(),
                         original_data_source_position=np.asarray(self.original_data_source_position).tolist(),
                         data_source_position=self.data_source_position.tolist(), layers=(self.layer_labels if self.has_layers else False),
                         source_position=self.source_position.tolist(), source_radius=self.source_radius,
                         interpolation_method=self.interpolation_method, preprocess_data=self._preprocess_data,
                         data_processed=self.data_processed, boundary_condition=self.boundary_condition, start_type=self.start_type)
        if self.odor_present_threshold is not None:
            arguments['odor_present_threshold'] = self.odor_present_threshold
        with open(folder + '/METADATA.json', 'w') as f:
            json.dump(arguments, f, indent=4)
        if save_arrays:
            np.save(folder + '/data.npy', self._data)
            np.save(folder + '/start_probabilities.npy', self.start_probabilities)
        self.saved_at = os.path.abspath(folder).replace('\\', '/')
        print(f'Environment saved to: {folder}')

    @classmethod
    def load(cls, folder: str) -> 'Environment':
        import json, os
        assert os.path.exists(folder), "Folder doesn't exist..."
        assert folder.rstrip('/').split('/')[-1].startswith('Env-'), 'The folder provided is not the data of en Environment object.'
        with open(folder + '/METADATA.json', 'r') as f:
            a = json.load(f)
        if os.path.exists(folder + '/data.npy') and os.path.exists(folder + '/start_probabilities.npy'):
            env = cls.__new__(cls)
            env._data = np.ascontiguousarray(np.load(folder + '/data.npy'), dtype=np.float64)
            env.start_probabilities = np.load(folder + '/start_probabilities.npy')
            env.name, env.timesteps, env.dimensions = a['name'], a['timesteps'], a['dimensions']
            env.data_shape, env.shape = tuple(a['data_shape']), tuple(a['shape'])
            env.margins, env.data_bounds = np.array(a['margins']), np.array(a['data_bounds'])
            env.original_data_source_position = np.array(a['original_data_source_position'])
            env.data_source_position, env.source_position = np.array(a['data_source_position']), np.array(a['source_position'])
            env.source_radius = a['source_radius']
            env.has_layers = bool(a.get('layers'))
            env.layers = np.arange(len(env._data)) if env.has_layers else None
            env.layer_labels = list(a['layers']) if env.has_layers else None
            env.interpolation_method, env._preprocess_data, env.data_processed = a['interpolation_method'], a['preprocess_data'], True
            env.boundary_condition = a['boundary_condition']
            env.data_file_path, env.odor_present_threshold, env.start_type = a.get('data_file_path'), a.get('odor_present_threshold'), a.get('start_type')
            env.is_on_gpu, env._device_handles, env.reference_time_quirk = False, {}, True
        else:
            start_zone = a['start_type']
            bounds = None
            if start_zone.startswith('custom'):
                bounds = np.array(start_zone.split('_')[1:]).reshape((a['dimensions'], 2)).astype(int)
            env = cls(data_file=a['data_file_path'], data_source_position=a['original_data_source_position'], source_radius=a['source_radius'],
                      layers=a['layers'], shape=a['shape'], margins=a['margins'], interpolation_method=a['interpolation_method'],
                      preprocess_data=a['preprocess_data'], boundary_condition=a['boundary_condition'],
                      start_zone=(bounds if bounds is not None else start_zone), odor_present_threshold=a.get('odor_present_threshold'),
                      name=a['name'])
        env.saved_at = os.path.abspath(folder)
        return env

    def _source_data(self):
        'What a modified copy is rebuilt from: the data file, else the ORIGINAL (unresampled) frames.'
        return self.data_file_path if self.data_file_path is not None else self._original_data

    def modify(self, data_source_position=None, source_radius: float = None, shape=None, margins=None, multiplier=None,
               interpolation_method: str = None, boundary_condition: str = None) -> 'Environment':
        'A copy with one or more parameters changed (environment.py:1061-1123).'
        return Environment(data_file=self._source_data(),
                           data_source_position=(data_source_position if data_source_position is not None else self.original_data_source_position),
                           source_radius=(source_radius if source_radius is not None else self.source_radius),
                           layers=(self.layer_labels if self.has_layers else False),
                           shape=(shape if shape is not None else self.shape),
                           margins=(margins if margins is not None else self.margins),
                           multiplier=(multiplier if multiplier is not None else [1.0, 1.0]),
                           interpolation_method=(interpolation_method if interpolation_method is not None else self.interpolation_method),
                           preprocess_data=self._preprocess_data,
                           boundary_condition=(boundary_condition if boundary_condition is not None else self.boundary_condition),
                           start_zone=self.start_type, odor_present_threshold=self.odor_present_threshold, name=self.name)

    def modify_scale(self, scale_factor: float) -> 'Environment':
        'Everything scaled by one factor: shape, margins, source radius and data shape (environment.py:1126-1161).'
        return Environment(data_file=self._source_data(), data_source_position=self.original_data_source_position,
                           source_radius=self.source_radius * scale_factor, layers=(self.layer_labels if self.has_layers else False),
                           shape=(np.array(self.shape) * scale_factor).astype(int),
                           margins=(self.margins * scale_factor).astype(int), multiplier=[1.0, 1.0],
                           interpolation_method=self.interpolation_method, preprocess_data=self._preprocess_data,
                           boundary_condition=self.boundary_condition, start_zone=self.start_type,
                           odor_present_threshold=self.odor_present_threshold, name=self.name)

    # -- device side -----------------------------------------------------------------------------
    def device_handle(self, device: int):
        'olfnav_env handle on `device` (created on first use, cached).'
        if device not in self._device_handles:
            from . import _lib
            _lib.require_cuda()
            h = ctypes.c_void_p()
            T, dh, dw = self._data.shape[-3:]
            L = len(self.layers) if self.has_layers else 1
            _lib.check(_lib.lib.olfnav_env_create_layered(ctypes.byref(h), device, self.shape[0], self.shape[1],
                                                          _lib.BOUNDARY.get(self.boundary_condition, 0), _lib.ptr(self._data), L, T, dh, dw,
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
