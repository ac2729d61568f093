"""ctypes binding of libce2e.so (C ABI declared in include/ce2e.h).

The shared library is built in-tree by `__graft_entry__.build()` (or `python -m
env_build_b200.build`) into env_build_b200/csrc/libce2e.so.  There is NO fallback: if
the library is missing, or a call fails, an exception is raised.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, 'csrc')
INCLUDE = os.path.join(os.path.dirname(_HERE), 'include')
LIB_PATH = os.environ.get('CE2E_LIB') or os.path.join(CSRC, 'libce2e.so')   # CE2E_LIB: A/B-test another build

MAX_PATHS = 4
MAX_VEH = 256
TASK_ID = dict(left=0, straight=1, right=2)

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-fmad=false',
              '-std=c++17', '-shared', '-Xcompiler', '-fPIC,-ffp-contract=off']


class Ce2eError(RuntimeError):
    pass


class TurnClasses(ctypes.Structure):
    _fields_ = [('tc', ctypes.c_int8 * MAX_VEH)]


class Config(ctypes.Structure):
    """ce2e_config of include/ce2e.h."""
    _fields_ = [('L', ctypes.c_double), ('W', ctypes.c_double), ('lane_width', ctypes.c_double),
                ('lane_number', ctypes.c_int), ('crossroad_size', ctypes.c_double), ('expected_v', ctypes.c_double),
                ('w_devi_v', ctypes.c_double), ('w_devi_y', ctypes.c_double), ('w_devi_phi', ctypes.c_double),
                ('w_punish_yaw_rate', ctypes.c_double), ('w_punish_steer', ctypes.c_double),
                ('w_punish_a_x', ctypes.c_double)]


_c = ctypes
_vp, _i, _i64, _d = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_double

# name -> (restype, argtypes); must list every symbol include/ce2e.h declares
SIGNATURES = {
    'ce2e_version': (_i, []),
    'ce2e_last_error': (_c.c_char_p, []),
    'ce2e_launch_count': (_i64, []),
    'ce2e_config_set': (_i, [_c.POINTER(Config)]),
    'ce2e_config_get': (_i, [_c.POINTER(Config)]),
    'ce2e_set_fast_trig': (_i, [_i]),
    'ce2e_set_tma': (_i, [_i]),
    'ce2e_last_step_kernel': (_i, []),
    'ce2e_paths_create': (_i, [_i, _i, _c.POINTER(_c.c_int32), _c.POINTER(_vp), _c.POINTER(_vp),
                               _c.POINTER(_vp), _c.POINTER(_vp)]),
    'ce2e_paths_destroy': (_i, [_vp]),
    'ce2e_action_transform': (_i, [_vp, _vp, _i64, _vp]),
    'ce2e_dynamics_step': (_i, [_vp, _i64, _vp, _d, _vp, _i64, _vp, _i, _i64, _vp]),
    'ce2e_find_closest_point': (_i, [_vp, _i, _vp, _vp, _i, _i, _vp, _vp, _i64, _vp]),
    'ce2e_grid_build_host': (_i, [_vp, _vp, _c.c_int32, _vp, _vp, _i64]),
    'ce2e_index_points': (_i, [_vp, _i, _vp, _i, _vp, _i64, _vp]),
    'ce2e_tracking_error': (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i64, _i64, _vp]),
    'ce2e_compute_rewards': (_i, [_i, _vp, _i64, _vp, _i, _i, _vp, _vp, _i64, _vp]),
    'ce2e_compute_next_obses': (_i, [_vp, _i, _vp, _vp, _i64, _vp, _c.POINTER(TurnClasses), _i, _i, _i,
                                     _vp, _i64, _i64, _vp]),
    'ce2e_env_step': (_i, [_vp, _vp, _vp, _i64, _vp, _c.POINTER(TurnClasses), _i, _i, _i, _vp, _i64, _vp, _vp, _vp,
                           _vp, _i64, _vp]),
    'ce2e_env_reset': (_i, [_vp, _c.c_uint64, _vp, _vp, _i, _vp, _i64, _vp, _vp, _i, _i, _i64, _vp]),
    'ce2e_env_step_reset': (_i, [_vp, _vp, _vp, _i64, _vp, _c.POINTER(TurnClasses), _i, _i, _i, _vp, _i64, _vp, _vp, _vp,
                                 _vp, _vp, _c.c_uint64, _vp, _i, _vp, _i64, _vp]),
    'ce2e_philox4x32': (None, [_c.POINTER(_c.c_uint32), _c.POINTER(_c.c_uint32), _c.POINTER(_c.c_uint32)]),
    'ce2e_judge_done': (_i, [_i, _vp, _i64, _vp, _i, _i, _i, _vp, _i64, _vp]),
    'ce2e_veh_predict': (_i, [_vp, _i64, _c.POINTER(TurnClasses), _i, _vp, _i64, _i64, _vp]),
    'ce2e_rollout_step': (_i, [_vp, _i, _vp, _vp, _i64, _vp, _c.POINTER(TurnClasses), _i, _i, _i,
                               _vp, _i64, _vp, _vp, _i64, _vp]),
    'ce2e_rollout_step_backward': (_i, [_vp, _i, _vp, _vp, _i64, _vp, _i, _i, _vp, _i64, _vp, _vp, _i64, _vp, _i64,
                                        _vp]),
    'ce2e_select_vehicles': (_i, [_i, _vp, _vp, _i, _vp, _i, _vp, _vp, _i64, _i64, _vp]),
    'ce2e_rollout_horizon': (_i, [_vp, _i, _vp, _vp, _i64, _vp, _c.POINTER(TurnClasses), _i, _i, _i, _vp, _i64, _vp,
                                  _i64, _vp]),
    'ce2e_ss': (_i, [_vp, _i64, _vp, _i64, _i, _i, _d, _vp, _i64, _vp]),
}

_lib = None


def build(verbose=False):
    """Compile csrc/ce2e.cu for sm_100a into csrc/libce2e.so (nvcc cross-compiles without a GPU)."""
    src = os.path.join(CSRC, 'ce2e.cu')
    cmd = ['nvcc'] + NVCC_FLAGS + ['-I', INCLUDE, '-I', CSRC, '-o', LIB_PATH, src]
    if verbose:
        cmd.insert(1, '-Xptxas=-v')
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise Ce2eError('nvcc failed:\n%s\n%s' % (' '.join(cmd), res.stderr))
    return res.stderr if verbose else LIB_PATH


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh', '.h'))]
    deps.append(os.path.join(INCLUDE, 'ce2e.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def load():
    """dlopen libce2e.so and type its entry points.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Ce2eError('%s not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                        '(there is no CPU fallback)' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().ce2e_last_error().decode('utf-8', 'replace')
        if rc in (-1, -2, -3, -4):
            raise ValueError('libce2e: %s (code %d)' % (msg, rc))
        raise Ce2eError('libce2e: %s (code %d)' % (msg, rc))


def launch_count():
    return int(load().ce2e_launch_count())


def make_turn_classes(classes):
    t = TurnClasses()
    if len(classes) > MAX_VEH:
        raise ValueError('at most %d vehicles per row' % MAX_VEH)
    for i, c in enumerate(classes):
        t.tc[i] = int(c)
    return t


def set_tma(mode):
    """See ce2e_set_tma in include/ce2e.h: False / 0 cp.async kernel, True / 1 TMA kernel (default),
    2 / 3 TMA kernel with the overlapped / balanced work split forced.  Returns the previous mode."""
    return int(load().ce2e_set_tma(int(mode)))


def set_fast_trig(enable):
    """See ce2e_set_fast_trig in include/ce2e.h (off by default).  Returns the previous setting."""
    return bool(load().ce2e_set_fast_trig(int(bool(enable))))
