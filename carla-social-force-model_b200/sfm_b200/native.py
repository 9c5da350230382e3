"""ctypes binding of ``libsfm_b200.so`` (C ABI in ``include/sfm_b200.h``).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  Loading it needs the CUDA runtime but no
GPU; *using* it needs a B200: ``Context()`` raises ``SfmError`` otherwise -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), 'csrc')
LIB_PATH = os.environ.get('SFM_LIB') or os.path.join(HERE, 'libsfm_b200.so')     # SFM_LIB: tuning builds only

FORCE_CLASSES = ('acceleration_force', 'pedestrian_force', 'border_force', 'static_obstacle_force',
                 'dynamic_obstacle_force')                      # pedestrian_simulation.py:37-48 dict order
ACCELERATION, PEDESTRIAN, BORDER, STATIC_OBSTACLE, DYNAMIC_OBSTACLE = range(5)
ABI_VERSION = 5

# every symbol include/sfm_b200.h declares (tests/test_abi.py checks the header against this list and the .so)
SYMBOLS = ('sfm_abi_version', 'sfm_last_error', 'sfm_device_count', 'sfm_create', 'sfm_destroy', 'sfm_set_stream',
           'sfm_synchronize', 'sfm_set_params', 'sfm_set_origin', 'sfm_set_partition', 'sfm_upload_state',
           'sfm_update_kinematics', 'sfm_update_targets', 'sfm_download_state', 'sfm_set_borders', 'sfm_set_obstacles',
           'sfm_force', 'sfm_enumerate_pairs', 'sfm_count_point_evaluations', 'sfm_step', 'sfm_tick_host', 'sfm_tick_records', 'sfm_apply_force',
           'sfm_host_column_gather', 'sfm_host_column_equal', 'sfm_host_register', 'sfm_host_unregister', 'sfm_download_force',
           'sfm_download_class_force', 'sfm_gather_buffer', 'sfm_stage', 'sfm_step_begin', 'sfm_step_end',
           'sfm_set_reorder_interval', 'sfm_reorder_slots', 'sfm_get_slot_order', 'sfm_set_slot_order',
           'sfm_force_accumulator', 'sfm_set_profiling', 'sfm_reset_stats', 'sfm_get_stats',
           # lifecycle
           'sfm_set_mode_machines', 'sfm_set_traffic', 'sfm_tick_modes', 'sfm_download_modes', 'sfm_set_routes',
           'sfm_advance_waypoints', 'sfm_download_routes', 'sfm_append_pedestrians', 'sfm_despawn_finished', 'sfm_lifecycle_counters', 'sfm_set_vehicles',
           'sfm_advance_vehicles', 'sfm_download_vehicles', 'sfm_record_begin', 'sfm_record_frame',
           'sfm_download_frames',
           # peer-memory exchange (K7)
           'sfm_peer_export', 'sfm_peer_import', 'sfm_step_peer', 'sfm_peer_barrier', 'sfm_peer_status')


class SfmError(RuntimeError):
    pass


class MoussaidParams(C.Structure):
    _fields_ = [('lambda_weight', C.c_double), ('A', C.c_double), ('gamma', C.c_double), ('n', C.c_double),
                ('n_prime', C.c_double), ('epsilon', C.c_double), ('perception_threshold', C.c_double)]


class Params(C.Structure):
    _fields_ = [('step_length', C.c_double), ('tau', C.c_double), ('max_speed_factor', C.c_double),
                ('border_a', C.c_double), ('border_b', C.c_double),
                ('ped', MoussaidParams), ('static_obs', MoussaidParams), ('dynamic_obs', MoussaidParams),
                ('use_ped_radius', C.c_int32), ('enable', C.c_int32 * 5)]


class Stats(C.Structure):
    _fields_ = [('launches', C.c_int64), ('steps', C.c_int64), ('ms_pairs', C.c_double), ('ms_cells', C.c_double),
                ('ms_segments', C.c_double), ('ms_integrate', C.c_double), ('pair_launches', C.c_int64),
                ('fixup_rows', C.c_int64), ('pair_evaluations', C.c_int64), ('ms_lifecycle', C.c_double), ('graph_replays', C.c_int64),
                ('local_tile_pairs', C.c_int64)]


NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC',
              '-shared']


def build(force=False, verbose=False):
    """Compile ``csrc/sfm_api.cu`` into ``libsfm_b200.so`` next to this file (nvcc cross-compiles without a GPU)."""
    sources = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(('.cu', '.cuh'))]
    sources.append(os.path.join(os.path.dirname(os.path.dirname(HERE)), 'include', 'sfm_b200.h'))
    if not force and os.path.exists(LIB_PATH):
        if os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(s) for s in sources):
            return LIB_PATH
    nvcc = os.environ.get('NVCC', 'nvcc')
    cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', LIB_PATH,
                                                                            os.path.join(CSRC, 'sfm_api.cu')]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise SfmError('nvcc failed:\n' + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


_lib = None


def lib():
    """Load the shared library (once) and declare the prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SfmError(f'{LIB_PATH} is missing: run `python -c "import __graft_entry__ as g; g.build()"` first; '
                       'there is no CPU fallback')
    L = C.CDLL(LIB_PATH)
    p_ctx, p_d, p_u8, p_i64 = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint8), C.POINTER(C.c_int64)
    i64 = C.c_int64
    sig = {
        'sfm_abi_version': (C.c_int, []),
        'sfm_last_error': (C.c_char_p, []),
        'sfm_device_count': (C.c_int, [C.POINTER(C.c_int)]),
        'sfm_create': (C.c_int, [C.c_int, C.POINTER(p_ctx)]),
        'sfm_destroy': (C.c_int, [p_ctx]),
        'sfm_set_stream': (C.c_int, [p_ctx, C.c_void_p]),
        'sfm_synchronize': (C.c_int, [p_ctx]),
        'sfm_set_params': (C.c_int, [p_ctx, C.POINTER(Params)]),
        'sfm_set_origin': (C.c_int, [p_ctx, C.c_double, C.c_double, C.c_double]),
        'sfm_set_partition': (C.c_int, [p_ctx, C.c_int, C.c_int, i64]),
        'sfm_upload_state': (C.c_int, [p_ctx, i64, p_d, p_d, p_d, p_d, p_d, p_u8]),
        'sfm_update_kinematics': (C.c_int, [p_ctx, i64, p_d, p_d]),
        'sfm_update_targets': (C.c_int, [p_ctx, i64, p_d, p_d, p_u8]),
        'sfm_download_state': (C.c_int, [p_ctx, i64, p_d, p_d]),
        'sfm_set_borders': (C.c_int, [p_ctx, i64, p_d, p_d, p_i64, p_d]),
        'sfm_set_obstacles': (C.c_int, [p_ctx, C.c_int, i64, p_d, p_d, p_i64, p_d]),
        'sfm_force': (C.c_int, [p_ctx, C.c_int, i64, p_d]),
        'sfm_enumerate_pairs': (C.c_int, [p_ctx, C.c_int, i64, p_i64, p_i64]),
        'sfm_count_point_evaluations': (C.c_int, [p_ctx, C.c_int, p_i64, p_i64]),
        'sfm_step': (C.c_int, [p_ctx, C.c_int, C.c_int]),
        'sfm_tick_host': (C.c_int, [p_ctx, i64, p_d, p_d, p_d, p_d]),
        'sfm_apply_force': (C.c_int, [p_ctx, i64, p_d, p_d]),
        'sfm_tick_records': (C.c_int, [p_ctx, i64, C.c_void_p, i64, p_i64, C.c_double, C.c_int, p_i64, i64, C.c_int,
                                      C.POINTER(C.c_int)]),
        'sfm_host_register': (C.c_int, [C.c_void_p, C.c_size_t]),
        'sfm_host_unregister': (C.c_int, [C.c_void_p]),
        'sfm_host_column_gather': (C.c_int, [C.c_void_p, i64, i64, i64, i64, C.c_void_p]),
        'sfm_host_column_equal': (C.c_int, [C.c_void_p, i64, i64, i64, i64, C.c_void_p, C.POINTER(C.c_int)]),
        'sfm_download_force': (C.c_int, [p_ctx, i64, p_d]),
        'sfm_download_class_force': (C.c_int, [p_ctx, C.c_int, i64, p_d]),
        'sfm_gather_buffer': (C.c_int, [p_ctx, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
        'sfm_stage': (C.c_int, [p_ctx]),
        'sfm_set_reorder_interval': (C.c_int, [p_ctx, C.c_int]),
        'sfm_reorder_slots': (C.c_int, [p_ctx]),
        'sfm_get_slot_order': (C.c_int, [p_ctx, C.c_int64, C.c_void_p]),
        'sfm_set_slot_order': (C.c_int, [p_ctx, C.c_int64, C.c_void_p]),
        'sfm_step_begin': (C.c_int, [p_ctx]),
        'sfm_step_end': (C.c_int, [p_ctx, C.c_int]),
        'sfm_force_accumulator': (C.c_int, [p_ctx, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
        'sfm_set_profiling': (C.c_int, [p_ctx, C.c_int]),
        'sfm_reset_stats': (C.c_int, [p_ctx]),
        'sfm_get_stats': (C.c_int, [p_ctx, C.POINTER(Stats)]),
        'sfm_set_mode_machines': (C.c_int, [p_ctx, i64, p_d, p_d, p_d, p_d, p_d, C.c_double]),
        'sfm_set_traffic': (C.c_int, [p_ctx, i64, p_d, p_d, p_d]),
        'sfm_tick_modes': (C.c_int, [p_ctx, C.c_double]),
        'sfm_download_modes': (C.c_int, [p_ctx, i64, p_u8, p_d, p_d, p_d]),
        'sfm_set_routes': (C.c_int, [p_ctx, i64, p_i64, p_d, p_u8, C.c_double, C.c_int]),
        'sfm_advance_waypoints': (C.c_int, [p_ctx]),
        'sfm_download_routes': (C.c_int, [p_ctx, i64, p_i64, p_u8, p_d]),
        'sfm_append_pedestrians': (C.c_int, [p_ctx, i64, p_d, p_d, p_d, p_d, p_d, p_u8, p_d, p_d, p_d, p_d, p_d, p_i64, p_d, p_u8]),
        'sfm_despawn_finished': (C.c_int, [p_ctx, p_i64, p_i64]),
        'sfm_lifecycle_counters': (C.c_int, [p_ctx, p_i64]),
        'sfm_set_vehicles': (C.c_int, [p_ctx, i64, p_d, p_d, p_d, p_d, C.c_double, C.c_double]),
        'sfm_advance_vehicles': (C.c_int, [p_ctx, C.c_double]),
        'sfm_download_vehicles': (C.c_int, [p_ctx, i64, p_d, p_i64, i64, p_d]),
        'sfm_record_begin': (C.c_int, [p_ctx, i64]),
        'sfm_record_frame': (C.c_int, [p_ctx, C.c_double]),
        'sfm_download_frames': (C.c_int, [p_ctx, i64, i64, p_d, p_u8, p_d, p_i64]),
        'sfm_peer_export': (C.c_int, [p_ctx, C.c_void_p]),
        'sfm_peer_import': (C.c_int, [p_ctx, C.c_void_p]),
        'sfm_step_peer': (C.c_int, [p_ctx, C.c_int, C.c_int]),
        'sfm_peer_barrier': (C.c_int, [p_ctx]),
        'sfm_peer_status': (C.c_int, [p_ctx, p_i64, C.POINTER(C.c_int)]),
    }
    assert set(sig) == set(SYMBOLS)
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    if L.sfm_abi_version() != ABI_VERSION:
        raise SfmError('libsfm_b200.so ABI version mismatch; rebuild')
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise SfmError(lib().sfm_last_error().decode(errors='replace'))


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != shape:
        raise ValueError(f'expected shape {shape}, got {a.shape}')
    return a


def _ptr(a, ctype=C.c_double):
    return a.ctypes.data_as(C.POINTER(ctype)) if a is not None else None


class PinnedArray:
    """Keeps a numpy array page-locked (``sfm_host_register``) and alive until ``release`` -- DMA straight out of it."""

    def __init__(self, array):
        self.array = array                       # the reference keeps the memory from being freed while it is registered
        self.ptr = array.ctypes.data
        span = (len(array) - 1) * array.strides[0] + array.dtype.itemsize if len(array) else 0
        self.ok = span > 0 and lib().sfm_host_register(C.c_void_p(self.ptr), span) == 0

    def release(self):
        if self.ok:
            lib().sfm_host_unregister(C.c_void_p(self.ptr))
            self.ok = False
        self.array = None


def column_gather(state, field, width):
    """Raw bytes of one column of a structured array as a packed uint8 [n, width] array (works for object columns)."""
    out = np.empty((len(state), width), dtype=np.uint8)
    _check(lib().sfm_host_column_gather(C.c_void_p(state.ctypes.data), state.strides[0], state.dtype.fields[field][1],
                                        width, len(state), C.c_void_p(out.ctypes.data)))
    return out


def column_equal(state, field, packed):
    same = C.c_int()
    _check(lib().sfm_host_column_equal(C.c_void_p(state.ctypes.data), state.strides[0], state.dtype.fields[field][1],
                                       packed.shape[1], len(state), C.c_void_p(packed.ctypes.data), C.byref(same)))
    return bool(same.value)


def pack_point_set(rings):
    """list of (P_k, 2) arrays -> (offsets int64 [K+1], points float64 [sum P_k, 2])  -- the CSR wire format."""
    sizes = np.fromiter((len(r) for r in rings), dtype=np.int64, count=len(rings))
    offsets = np.zeros(len(rings) + 1, dtype=np.int64)
    np.cumsum(sizes, out=offsets[1:])
    if len(rings):
        points = np.ascontiguousarray(np.concatenate([np.asarray(r, dtype=np.float64).reshape(-1, 2) for r in rings]))
    else:
        points = np.zeros((0, 2))
    return offsets, points


def moussaid_from_config(section, defaults):
    """Read one Moussaid section exactly like forces.py:66-72 / :200-206 (missing keys fall back to the code defaults)."""
    m = MoussaidParams()
    m.lambda_weight = section.get('lambda', defaults[0])
    m.A = section.get('A', defaults[1])
    m.gamma = section.get('gamma', defaults[2])
    m.n = section.get('n', defaults[3])
    m.n_prime = section.get('n_prime', defaults[4])
    m.epsilon = section.get('epsilon', defaults[5])
    m.perception_threshold = section.get('perception_threshold', defaults[6])
    return m


_MOUSSAID_DEFAULTS = (2.0, 4.5, 0.35, 2.0, 3.0, 0.005, 20)


def params_from_config(sfm_config, step_length, enable=None, strict=False):
    """Translate the parsed sfm_config.toml into ``Params``, reading the keys the reference's code reads.

    Bug-compatible by default: ``max_speed_factor`` (not the shipped ``max_speed_multiplier``)
    and ``[goal_force].tau`` (not the shipped ``[acceleration_force].tau``) are honoured, both defaulting to the shipped
    values.  With ``strict=True`` missing mandatory sections raise ``KeyError`` like forces.py:66,134,197,199 do.
    """
    p = Params()
    p.step_length = step_length
    p.tau = sfm_config.get('goal_force', {}).get('tau', 0.5)
    p.max_speed_factor = sfm_config.get('max_speed_factor', 1.3)
    border = sfm_config['border_force'] if strict else sfm_config.get('border_force', {})
    p.border_a, p.border_b = border.get('a', 3.0), border.get('b', 0.1)
    get = (lambda k: sfm_config[k]) if strict else (lambda k: sfm_config.get(k, {}))
    p.ped = moussaid_from_config(get('pedestrian_force'), _MOUSSAID_DEFAULTS)
    p.static_obs = moussaid_from_config(get('static_obstacle_force'), _MOUSSAID_DEFAULTS)
    p.dynamic_obs = moussaid_from_config(get('dynamic_obstacle_force'), _MOUSSAID_DEFAULTS)
    p.use_ped_radius = int(bool(sfm_config.get('use_ped_radius', False)))
    switches = sfm_config.get('forces', {}) if enable is None else enable
    for k, name in enumerate(FORCE_CLASSES):
        p.enable[k] = int(bool(switches.get(name, False)))
    return p


class Context:
    """One device context = one rank's pedestrians + the replicated point sets."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        self._lib = lib()
        _check(self._lib.sfm_create(int(device), C.byref(self._h)))
        self.device = int(device)
        self.n = 0

    def close(self):
        if self._h:
            self._lib.sfm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- configuration
    def set_params(self, params):
        self.params = params
        _check(self._lib.sfm_set_params(self._h, C.byref(params)))

    def set_stream(self, cuda_stream):
        _check(self._lib.sfm_set_stream(self._h, C.c_void_p(cuda_stream) if cuda_stream else None))

    def set_origin(self, ox, oy, oz=0.0):
        _check(self._lib.sfm_set_origin(self._h, ox, oy, oz))

    def set_partition(self, world, rank, rows_pad):
        _check(self._lib.sfm_set_partition(self._h, world, rank, rows_pad))

    def synchronize(self):
        _check(self._lib.sfm_synchronize(self._h))

    # -- state
    def upload_state(self, loc, vel, next_waypoint, radius, target_speed, mode):
        n = len(loc)
        loc, vel, wp = _f64(loc, (n, 3)), _f64(vel, (n, 3)), _f64(next_waypoint, (n, 3))
        radius, speed = _f64(radius, (n,)), _f64(target_speed, (n,))
        mode = np.ascontiguousarray(mode, dtype=np.uint8)
        _check(self._lib.sfm_upload_state(self._h, n, _ptr(loc), _ptr(vel), _ptr(wp), _ptr(radius), _ptr(speed),
                                          _ptr(mode, C.c_uint8)))
        self.n = n

    def update_kinematics(self, loc, vel):
        loc, vel = _f64(loc, (self.n, 3)), _f64(vel, (self.n, 3))
        _check(self._lib.sfm_update_kinematics(self._h, self.n, _ptr(loc), _ptr(vel)))

    def update_targets(self, next_waypoint=None, target_speed=None, mode=None):
        wp = _f64(next_waypoint, (self.n, 3)) if next_waypoint is not None else None
        sp = _f64(target_speed, (self.n,)) if target_speed is not None else None
        md = np.ascontiguousarray(mode, dtype=np.uint8) if mode is not None else None
        _check(self._lib.sfm_update_targets(self._h, self.n, _ptr(wp), _ptr(sp), _ptr(md, C.c_uint8)))

    def download_state(self, loc=None, vel=None):
        loc = np.empty((self.n, 3)) if loc is None else loc
        vel = np.empty((self.n, 3)) if vel is None else vel
        _check(self._lib.sfm_download_state(self._h, self.n, _ptr(loc), _ptr(vel)))
        return loc, vel

    # -- point sets
    def set_borders(self, borders, section_center, section_length):
        n = len(borders)
        if n == 0:
            _check(self._lib.sfm_set_borders(self._h, 0, None, None, None, None))
            return
        offsets, points = pack_point_set(borders)
        center = _f64(np.asarray(section_center, dtype=np.float64).reshape(-1, 2), (n, 2))
        length = _f64(np.asarray(section_length, dtype=np.float64).reshape(-1), (n,))
        _check(self._lib.sfm_set_borders(self._h, n, _ptr(center), _ptr(length), _ptr(offsets, C.c_int64), _ptr(points)))

    def set_obstacles(self, which, centers, rings, velocities=None):
        n = len(rings)
        if n == 0:
            _check(self._lib.sfm_set_obstacles(self._h, which, 0, None, None, None, None))
            return
        offsets, points = pack_point_set(rings)
        centers = _f64(np.asarray(centers, dtype=np.float64).reshape(-1, 2), (n, 2))
        vel = _f64(np.asarray(velocities, dtype=np.float64).reshape(-1, 2), (n, 2)) if velocities is not None else None
        _check(self._lib.sfm_set_obstacles(self._h, which, n, _ptr(centers), _ptr(vel), _ptr(offsets, C.c_int64),
                                           _ptr(points)))

    def set_obstacles_csr(self, which, centers, velocities, offsets, points):
        n = len(centers)
        centers, points = _f64(centers, (n, 2)), _f64(points)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        vel = _f64(velocities, (n, 2)) if velocities is not None else None
        _check(self._lib.sfm_set_obstacles(self._h, which, n, _ptr(centers), _ptr(vel), _ptr(offsets, C.c_int64),
                                           _ptr(points)))

    # -- forces and stepping
    def force(self, force_class, out=None):
        out = np.empty((self.n, 3)) if out is None else out
        _check(self._lib.sfm_force(self._h, force_class, self.n, _ptr(out)))
        return out

    def enumerate_pairs(self, force_class, capacity=1 << 22):
        buf = np.empty((capacity, 3), dtype=np.int64)
        count = C.c_int64()
        _check(self._lib.sfm_enumerate_pairs(self._h, force_class, capacity, _ptr(buf, C.c_int64), C.byref(count)))
        if count.value > capacity:
            return self.enumerate_pairs(force_class, int(count.value))
        t = buf[:count.value]
        return t[np.lexsort((t[:, 1], t[:, 0]))]

    def count_point_evaluations(self, force_class):
        """-> (pairs inside the cutoff, distances the reference's argmin ranges over) for a cell-list class."""
        pairs, evals = C.c_int64(), C.c_int64()
        _check(self._lib.sfm_count_point_evaluations(self._h, force_class, C.byref(pairs), C.byref(evals)))
        return pairs.value, evals.value

    def step(self, n_steps=1, integrate_positions=True):
        _check(self._lib.sfm_step(self._h, int(n_steps), int(bool(integrate_positions))))

    def tick_host(self, loc, vel, new_vel, new_loc=None):
        """Host-buffer tick; arrays must be C-contiguous float64 [n, 3] (pinned memory makes the copies asynchronous)."""
        _check(self._lib.sfm_tick_host(self._h, self.n, _ptr(loc), _ptr(vel), _ptr(new_vel), _ptr(new_loc)))

    def tick_records(self, state, sim_time, tick_modes, identity=None):
        """One drop-in tick on the structured pedestrian table ``state`` (in place: new velocities land in ``state['vel']``,
        and with ``tick_modes`` the applied target speeds in ``state['target_speed']``).  Returns the lifecycle counters.

        ``identity``: ``None``, or ``('adopt' | 'check', field)`` -- the 8-byte column ``field`` (the ``mode`` object
        pointers) is adopted as the table's identity, or checked against the adopted one on the device; with ``'check'``
        the return value is ``(counters, changed)`` and a changed table is handed back untouched."""
        fields = state.dtype.fields
        offsets = np.array([fields[k][1] for k in ('loc', 'vel', 'next_waypoint', 'radius', 'target_speed')], dtype=np.int64)
        counters = np.zeros(4, dtype=np.int64)
        changed = C.c_int(0)
        id_mode, id_off = 0, 0
        if identity is not None:
            id_mode, id_off = {'adopt': 1, 'check': 2}[identity[0]], fields[identity[1]][1]
        _check(self._lib.sfm_tick_records(self._h, len(state), C.c_void_p(state.ctypes.data), state.strides[0],
                                          _ptr(offsets, C.c_int64), float(sim_time), int(bool(tick_modes)),
                                          _ptr(counters, C.c_int64), int(id_off), id_mode, C.byref(changed)))
        if identity is not None and identity[0] == 'check':
            return counters, bool(changed.value)
        return counters

    def apply_force(self, force, out=None):
        """calculate_new_velocities on the device for a caller-composed force [n, 3]; returns the new velocities."""
        force = _f64(force, (self.n, 3))
        out = np.empty((self.n, 3)) if out is None else out
        _check(self._lib.sfm_apply_force(self._h, self.n, _ptr(force), _ptr(out)))
        return out

    def download_force(self, out=None):
        out = np.empty((self.n, 3)) if out is None else out
        _check(self._lib.sfm_download_force(self._h, self.n, _ptr(out)))
        return out

    # -- multi-GPU plumbing and accounting
    def gather_buffer(self):
        ptr, nbytes = C.c_void_p(), C.c_size_t()
        _check(self._lib.sfm_gather_buffer(self._h, C.byref(ptr), C.byref(nbytes)))
        return ptr.value, nbytes.value

    def step_begin(self):
        _check(self._lib.sfm_step_begin(self._h))

    def step_end(self, integrate_positions=True):
        _check(self._lib.sfm_step_end(self._h, int(bool(integrate_positions))))

    def force_accumulator(self):
        ptr, nbytes = C.c_void_p(), C.c_size_t()
        _check(self._lib.sfm_force_accumulator(self._h, C.byref(ptr), C.byref(nbytes)))
        return ptr.value, nbytes.value

    def stage(self):
        _check(self._lib.sfm_stage(self._h))

    # -- staged slot order (csrc/k8_order.cuh): rows staged along a Hilbert curve, so the pair kernel's local path applies
    def set_reorder_interval(self, ticks):
        _check(self._lib.sfm_set_reorder_interval(self._h, int(ticks)))

    def reorder_slots(self):
        _check(self._lib.sfm_reorder_slots(self._h))

    def slot_order(self):
        """int32 [n]: the staged slot of every row (a permutation of range(n); the identity unless reordering is on)."""
        out = np.empty(self.n, dtype=np.int32)
        _check(self._lib.sfm_get_slot_order(self._h, self.n, out.ctypes.data_as(C.c_void_p)))
        return out

    def set_slot_order(self, slot_of_row):
        a = np.ascontiguousarray(slot_of_row, dtype=np.int32)
        if a.shape != (self.n,):
            raise ValueError(f'expected shape {(self.n,)}, got {a.shape}')
        _check(self._lib.sfm_set_slot_order(self._h, self.n, a.ctypes.data_as(C.c_void_p)))

    # -- peer-memory exchange (K7)
    PEER_HANDLE_BYTES = 3 * 64

    def peer_export(self):
        buf = C.create_string_buffer(self.PEER_HANDLE_BYTES)
        _check(self._lib.sfm_peer_export(self._h, buf))
        return buf.raw

    def peer_import(self, handles_by_rank):
        blob = b''.join(handles_by_rank)
        _check(self._lib.sfm_peer_import(self._h, C.create_string_buffer(blob, len(blob))))

    def step_peer(self, n_steps=1, integrate_positions=True):
        _check(self._lib.sfm_step_peer(self._h, int(n_steps), int(bool(integrate_positions))))

    def peer_barrier(self):
        _check(self._lib.sfm_peer_barrier(self._h))

    def peer_status(self):
        n, bad = C.c_int64(), C.c_int()
        _check(self._lib.sfm_peer_status(self._h, C.byref(n), C.byref(bad)))
        stalled = [r for r in range(32) if bad.value >> r & 1]
        return dict(barriers=n.value, timed_out=bool(bad.value), stalled_ranks=stalled)

    # -- lifecycle on the device
    def set_mode_machines(self, initial_target_speed, crossing_speed, crossing_safety_margin, mode_target_speed=None,
                          next_mode_time=None, waiting_time=5.0):
        n = self.n
        ini, cro, mar = _f64(initial_target_speed, (n,)), _f64(crossing_speed, (n,)), _f64(crossing_safety_margin, (n,))
        spd = _f64(ini if mode_target_speed is None else mode_target_speed, (n,))
        nxt = _f64(np.full(n, -1.0) if next_mode_time is None else next_mode_time, (n,))     # ped_mode_manager.py:27
        _check(self._lib.sfm_set_mode_machines(self._h, n, _ptr(ini), _ptr(cro), _ptr(mar), _ptr(spd), _ptr(nxt),
                                               float(waiting_time)))

    def set_traffic(self, centers, velocities, extents):
        v = len(centers)
        if v == 0:
            _check(self._lib.sfm_set_traffic(self._h, 0, None, None, None))
            return
        c, u, e = (_f64(np.asarray(a, dtype=np.float64).reshape(-1, 2), (v, 2)) for a in (centers, velocities, extents))
        _check(self._lib.sfm_set_traffic(self._h, v, _ptr(c), _ptr(u), _ptr(e)))

    def tick_modes(self, sim_time):
        _check(self._lib.sfm_tick_modes(self._h, float(sim_time)))

    def download_modes(self):
        """-> dict(mode uint8, mode_target_speed, next_mode_time, target_speed)"""
        n = self.n
        mode, spd, nxt, tgt = np.empty(n, dtype=np.uint8), np.empty(n), np.empty(n), np.empty(n)
        _check(self._lib.sfm_download_modes(self._h, n, _ptr(mode, C.c_uint8), _ptr(spd), _ptr(nxt), _ptr(tgt)))
        return dict(mode=mode, mode_target_speed=spd, next_mode_time=nxt, target_speed=tgt)

    def download_mode_codes(self):
        mode = np.empty(self.n, dtype=np.uint8)
        _check(self._lib.sfm_download_modes(self._h, self.n, _ptr(mode, C.c_uint8), None, None, None))
        return mode

    def set_routes(self, routes, distance_threshold=2.0, fused=True):
        """``routes``: per pedestrian a list of (waypoint(3), crossing_road) tuples -- SimulationRunner.waypoint_dict."""
        sizes = np.fromiter((len(r) for r in routes), dtype=np.int64, count=len(routes))
        offsets = np.zeros(len(routes) + 1, dtype=np.int64)
        np.cumsum(sizes, out=offsets[1:])
        total = int(offsets[-1])
        wps = np.zeros((max(total, 1), 3))
        cross = np.zeros(max(total, 1), dtype=np.uint8)
        k = 0
        for r in routes:
            for wp, crossing in r:
                wps[k], cross[k] = wp, bool(crossing)
                k += 1
        self.set_routes_csr(offsets, wps, cross, distance_threshold, fused)

    def set_routes_csr(self, offsets, waypoints, crossing, distance_threshold=2.0, fused=True):
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        wps = _f64(waypoints)
        cross = np.ascontiguousarray(crossing, dtype=np.uint8)
        _check(self._lib.sfm_set_routes(self._h, self.n, _ptr(offsets, C.c_int64), _ptr(wps), _ptr(cross, C.c_uint8),
                                        float(distance_threshold), int(bool(fused))))

    def advance_waypoints(self):
        _check(self._lib.sfm_advance_waypoints(self._h))

    def download_routes(self):
        n = self.n
        cursor, finished, wp = np.empty(n, dtype=np.int64), np.empty(n, dtype=np.uint8), np.empty((n, 3))
        _check(self._lib.sfm_download_routes(self._h, n, _ptr(cursor, C.c_int64), _ptr(finished, C.c_uint8), _ptr(wp)))
        return cursor, finished.astype(bool), wp

    def append_pedestrians(self, loc, vel, next_waypoint, radius, target_speed, mode, machines=None, routes=None):
        """Spawn ``m`` pedestrians at the end of the table.  ``machines`` = (initial_target_speed, crossing_speed,
        crossing_safety_margin, mode_target_speed, next_mode_time) columns; ``routes`` = per pedestrian a list of
        (waypoint(3), crossing_road) tuples -- both required when the context carries them."""
        m = len(loc)
        loc, vel, wp = _f64(loc, (m, 3)), _f64(vel, (m, 3)), _f64(next_waypoint, (m, 3))
        radius, speed = _f64(radius, (m,)), _f64(target_speed, (m,))
        mode = np.ascontiguousarray(mode, dtype=np.uint8)
        mach = [None] * 5 if machines is None else [_f64(a, (m,)) for a in machines]
        offsets = wps = cross = None
        if routes is not None:
            sizes = np.fromiter((len(r) for r in routes), dtype=np.int64, count=m)
            offsets = np.zeros(m + 1, dtype=np.int64)
            np.cumsum(sizes, out=offsets[1:])
            wps = np.zeros((max(int(offsets[-1]), 1), 3))
            cross = np.zeros(max(int(offsets[-1]), 1), dtype=np.uint8)
            k = 0
            for r in routes:
                for w, crossing in r:
                    wps[k], cross[k] = w, bool(crossing)
                    k += 1
        _check(self._lib.sfm_append_pedestrians(self._h, m, _ptr(loc), _ptr(vel), _ptr(wp), _ptr(radius), _ptr(speed),
                                                _ptr(mode, C.c_uint8), *[_ptr(a) for a in mach],
                                                _ptr(offsets, C.c_int64), _ptr(wps), _ptr(cross, C.c_uint8)))
        self.n += m

    def despawn_finished(self):
        """Remove the pedestrians that arrived with no waypoint left; returns how many were removed."""
        after, removed = C.c_int64(), C.c_int64()
        _check(self._lib.sfm_despawn_finished(self._h, C.byref(after), C.byref(removed)))
        self.n = after.value
        return removed.value

    def lifecycle_counters(self):
        out = np.zeros(4, dtype=np.int64)
        _check(self._lib.sfm_lifecycle_counters(self._h, _ptr(out, C.c_int64)))
        return dict(zip(('crossings_started', 'idle_wakeups', 'handovers', 'finished'), (int(v) for v in out)))

    def set_vehicles(self, centers, yaw_deg, velocities, extents, resolution=0.1, size_factor=float(np.sqrt(2.0))):
        v = len(centers)
        self.n_vehicles = v
        if v == 0:
            _check(self._lib.sfm_set_vehicles(self._h, 0, None, None, None, None, resolution, size_factor))
            return
        c, u, e = (_f64(np.asarray(a, dtype=np.float64).reshape(-1, 2), (v, 2)) for a in (centers, velocities, extents))
        yaw = _f64(yaw_deg, (v,))
        _check(self._lib.sfm_set_vehicles(self._h, v, _ptr(c), _ptr(yaw), _ptr(u), _ptr(e), float(resolution),
                                          float(size_factor)))

    def advance_vehicles(self, dt):
        _check(self._lib.sfm_advance_vehicles(self._h, float(dt)))

    def download_vehicles(self):
        """-> (centers [V, 2], list of rings [(P_v, 2)])"""
        v = self.n_vehicles
        centers, offsets = np.empty((v, 2)), np.empty(v + 1, dtype=np.int64)
        _check(self._lib.sfm_download_vehicles(self._h, v, _ptr(centers), _ptr(offsets, C.c_int64), 0, None))
        points = np.empty((int(offsets[-1]), 2))
        _check(self._lib.sfm_download_vehicles(self._h, v, None, None, len(points), _ptr(points)))
        return centers, [points[offsets[k]:offsets[k + 1]] for k in range(v)]

    def record_begin(self, capacity_frames):
        _check(self._lib.sfm_record_begin(self._h, int(capacity_frames)))

    def record_frame(self, sim_time):
        _check(self._lib.sfm_record_frame(self._h, float(sim_time)))

    def download_frames(self, first=0, count=None):
        """-> (times [F], xyv [F, n, 4] = (x, y, v_x, v_y), mode uint8 [F, n])"""
        have = C.c_int64()
        _check(self._lib.sfm_download_frames(self._h, 0, 0, None, None, None, C.byref(have)))
        count = have.value - first if count is None else count
        xyv, mode, times = np.empty((count, self.n, 4)), np.empty((count, self.n), dtype=np.uint8), np.empty(count)
        if count:
            _check(self._lib.sfm_download_frames(self._h, first, count, _ptr(xyv), _ptr(mode, C.c_uint8), _ptr(times), None))
        return times, xyv, mode

    def set_profiling(self, enabled):
        _check(self._lib.sfm_set_profiling(self._h, int(bool(enabled))))

    def reset_stats(self):
        _check(self._lib.sfm_reset_stats(self._h))

    def stats(self):
        s = Stats()
        _check(self._lib.sfm_get_stats(self._h, C.byref(s)))
        return {name: getattr(s, name) for name, _ in Stats._fields_}
