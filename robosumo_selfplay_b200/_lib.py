"""ctypes binding of librs_b200.so (the C ABI declared in include/rs_b200.h).

There is no CPU fallback: if the CUDA library is missing or cannot be loaded the import of
any op raises.  The library is built in-tree by `python -m robosumo_selfplay_b200.build`.
"""
import ctypes
import os

from .morphology import rs_agent_model

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('RS_B200_LIB', os.path.join(_HERE, 'librs_b200.so'))   # env override: kernel-variant experiments
_lib = None

c_void_p, c_int, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_float


class rs_config(ctypes.Structure):
    _fields_ = [('num_envs', c_int), ('frame_skip', c_int), ('timestep_limit', c_int), ('newton_iters', c_int),
                ('timestep', c_float), ('ring_limit', c_float), ('init_pos_noise', c_float), ('init_vel_noise', c_float),
                ('seed', ctypes.c_uint64), ('device', c_int), ('reserved', c_int)]


class rs_mlp_job(ctypes.Structure):
    _fields_ = [('params', c_void_p), ('obs', c_void_p), ('obs_row_stride', ctypes.c_longlong), ('mean', c_void_p), ('value', c_void_p),
                ('activation', c_int), ('reserved', c_int)]


# every symbol include/rs_b200.h declares: name -> (restype, argtypes)
class rs_rollout_io(ctypes.Structure):
    _fields_ = [(k, c_void_p) for k in ('params0', 'params1', 'obs', 'rew', 'done', 'info', 'episode', 'mb_obs', 'mb_actions', 'mb_values', 'mb_nlp',
                                        'mb_opp_nlp', 'mb_dones', 'mb_shaping', 'mb_main', 'ep_done', 'ep_info', 'scratch')]


SYMBOLS = {
    'rs_agent_model_size': (c_int, []),
    'rs_last_error': (ctypes.c_char_p, []),
    'rs_launch_count': (ctypes.c_longlong, []),
    'rs_create': (c_int, [ctypes.POINTER(rs_config), ctypes.POINTER(rs_agent_model), ctypes.POINTER(c_void_p)]),
    'rs_destroy': (None, [c_void_p]),
    'rs_dims': (c_int, [c_void_p] + [ctypes.POINTER(c_int)] * 7),
    'rs_reset': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    'rs_set_state': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'rs_get_state': (c_int, [c_void_p] * 6),
    'rs_step': (c_int, [c_void_p] * 7 + [c_int, c_void_p]),
    'rs_step_host': (c_int, [c_void_p] * 7 + [c_int]),
    'rs_forward_debug': (c_int, [c_void_p] * 6),
    'rs_get_diag': (c_int, [c_void_p] * 3),
    'rs_tc_selftest': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'rs_param_count': (c_int, [c_int, c_int]),
    'rs_mlp_forward': (c_int, [c_void_p, c_int, c_int, c_void_p, ctypes.c_longlong, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    'rs_mlp_forward_multi': (c_int, [ctypes.POINTER(rs_mlp_job), c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'rs_rollout': (c_int, [c_void_p, c_int, ctypes.POINTER(rs_rollout_io), c_int, ctypes.c_ulonglong, ctypes.c_uint, c_int, c_void_p]),
    'rs_rollout_sample': (c_int, [c_int, c_int] + [c_void_p] * 6 + [ctypes.c_ulonglong, ctypes.c_uint, c_int] + [c_void_p] * 6),
    'rs_neglogp': (c_int, [c_int, c_int] + [c_void_p] * 5),
    'rs_vtrace': (c_int, [c_int, c_int, c_float, c_float, c_float, c_float] + [c_void_p] * 10),
    'rs_adv_moments': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    'rs_ppo_workspace_floats': (ctypes.c_longlong, [c_int, c_int, c_int]),
    'rs_ppo_grad': (c_int, [c_void_p, c_int, c_int] + [c_void_p] * 7 + [c_int, ctypes.c_longlong, c_void_p, c_float, c_float, c_float,
                            c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    'rs_adam_step': (c_int, [c_void_p] * 4 + [c_int, c_int, c_float, c_float, c_float, ctypes.c_longlong, c_float, c_float, c_float,
                             c_void_p, ctypes.c_longlong, c_void_p, c_void_p]),
    'rs_ppo_minibatch_step': (c_int, [c_void_p] * 3 + [c_int, c_int] + [c_void_p] * 7 + [c_int, c_float, c_float, c_float, c_float, c_float,
                                      ctypes.c_longlong] + [c_void_p] * 3 + [c_int] + [c_void_p] * 3 + [c_int, c_void_p]),
    'rs_epoch_split': (c_int, [c_void_p, ctypes.c_longlong, c_int, ctypes.c_longlong, ctypes.c_longlong, c_void_p, c_void_p, c_void_p]),
    'rs_epoch_split_host': (c_int, [c_void_p, c_int, ctypes.c_longlong, c_int, ctypes.c_longlong, ctypes.c_longlong, c_void_p, c_void_p]),
    'rs_adv_moments_multi': (c_int, [c_void_p, c_void_p, c_int, c_int, ctypes.c_longlong, c_void_p, c_void_p, c_void_p, c_void_p]),
    'rs_status_latch': (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    'rs_seed': (c_int, [c_void_p, ctypes.c_ulonglong]),
    'rs_legacy_shuffle': (c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_longlong]),
    'rs_legacy_shuffle32': (c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_longlong]),
    'rs_ppo_stats': (c_int, [c_void_p, c_void_p, c_int, c_int, ctypes.c_longlong, c_void_p, c_void_p]),
    'rs_ppo_minibatch_step_peer': (c_int, [c_void_p] * 4 + [c_int, c_int] + [c_void_p] * 7 + [c_int, ctypes.c_longlong, c_float, c_float, c_float, c_float, c_float,
                                           ctypes.c_longlong] + [c_void_p] * 6 + [c_int, c_void_p]),
    'rs_peer_create': (c_int, [c_int, c_int, ctypes.c_longlong, c_int, ctypes.POINTER(c_void_p)]),
    'rs_peer_handle_bytes': (c_int, []),
    'rs_peer_export': (c_int, [c_void_p, c_void_p]),
    'rs_peer_connect': (c_int, [c_void_p, c_void_p]),
    'rs_peer_send_buffer': (c_void_p, [c_void_p]),
    'rs_peer_allreduce': (c_int, [c_void_p, c_void_p, ctypes.c_longlong, c_void_p]),
    'rs_peer_error': (c_int, [c_void_p]),
    'rs_peer_destroy': (None, [c_void_p]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("robosumo_selfplay_b200: %s not found. Build it with `python -m robosumo_selfplay_b200.build` "
                               "(requires nvcc); there is no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.rs_agent_model_size() != ctypes.sizeof(rs_agent_model):
            raise RuntimeError("rs_agent_model layout mismatch between morphology.py and rs_b200.h")
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError("rs_b200 error %d: %s" % (rc, lib().rs_last_error().decode()))
