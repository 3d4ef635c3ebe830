"""ctypes binding of the C ABI (include/dq_decoding.h -> libdq_decoding.so).

There is no fallback: if the shared library is missing or a call fails, this raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DQ_DECODING_LIB selects an alternative build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("DQ_DECODING_LIB") or os.path.join(_HERE, "libdq_decoding.so")

OK, EINVAL, ECUDA, ESTATE = 0, -1, -2, -3
MODEL = {"X": 0, "DP": 1}
REFEREE_JOINT, REFEREE_SPLIT = 0, 1
(INFO_NUM_ACTIONS, INFO_OBS_CHANNELS, INFO_OBS_SIDE, INFO_MASK_WORDS, INFO_STATE_WORDS,
 INFO_STATE_STRIDE, INFO_NUM_STABS, INFO_N_TYPE3, INFO_N_TYPE1, INFO_RNG_BLOCKS) = range(10)
QINFO_NUM_PARAMS, QINFO_PACKED_ROWS, QINFO_NUM_TENSORS, QINFO_FLOPS_PER_SAMPLE = range(4)


class DQError(RuntimeError):
    pass


_lib = None
_vp, _i, _i64, _u64, _u32, _dbl, _f = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_uint32, C.c_double, C.c_float

_SIGNATURES = {
    "dq_last_error": (C.c_char_p, []),
    "dq_version": (_i, []),
    "dq_launch_count": (_i64, []),
    "dq_env_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _dbl, _dbl, _i64, _u64, _i64, _i]),
    "dq_env_destroy": (_i, [_vp]),
    "dq_env_info": (_i, [_vp, _i, C.POINTER(_i64)]),
    "dq_env_set_noise": (_i, [_vp, _dbl, _dbl]),
    "dq_env_set_max_attempts": (_i, [_vp, _i]),
    "dq_env_set_referee_lut": (_i, [_vp, _i, _vp, _i64, _vp, _i64]),
    "dq_env_reset": (_i, [_vp, _vp, _vp, _vp]),
    "dq_env_step": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "dq_env_step_random": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "dq_env_rollout_random": (_i, [_vp, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "dq_env_reset_host": (_i, [_vp, _vp, _vp]),
    "dq_env_step_host": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "dq_env_step_host_begin": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "dq_env_step_host_end": (_i, [_vp]),
    "dq_policy_random_legal_host": (_i, [_vp, _vp, _u32, _vp]),
    "dq_env_reset_host_packed": (_i, [_vp, _vp, _vp]),
    "dq_env_step_host_packed": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "dq_unpack_observations_host": (_i, [_vp, _i64, _i64, _i, _i, _vp]),
    "dq_env_get_state": (_i, [_vp, _vp, _vp]),
    "dq_env_set_state": (_i, [_vp, _vp, _vp]),
    "dq_policy_random_legal": (_i, [_vp, _vp, _u32, _vp, _vp]),
    "dq_policy_seek": (_i, [_vp, _u32, _vp]),
    "dq_policy_random_legal_next": (_i, [_vp, _vp, _vp, _vp]),
    "dq_env_packed_obs": (_i, [_vp, C.POINTER(_vp), C.POINTER(_i64), C.POINTER(_i64)]),
    "dq_qnet_create": (_i, [C.POINTER(_vp), _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i64, _i]),
    "dq_qnet_destroy": (_i, [_vp]),
    "dq_qnet_info": (_i, [_vp, _i, C.POINTER(_i64)]),
    "dq_qnet_param_layout": (_i, [_vp, _vp, _vp]),
    "dq_qnet_pack_obs": (_i, [_vp, _vp, _vp, _i64, _i64, _vp]),
    "dq_qnet_forward": (_i, [_vp, _vp, _vp, _i64, _i64, _vp, _i, _u64, _vp]),
    "dq_qnet_forward_tc": (_i, [_vp, _vp, _vp, _i64, _i64, _vp, _vp]),
    "dq_qnet_prepare_tc": (_i, [_vp, _vp, _vp]),
    "dq_qnet_forward_tc_train": (_i, [_vp, _vp, _vp, _i64, _i64, _vp, _u64, _vp]),
    "dq_qnet_backward_tc": (_i, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    "dq_qnet_fold_head": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "dq_qnet_tc_activation": (_i, [_vp, _i, C.POINTER(_vp), C.POINTER(_i64)]),
    "dq_qnet_backward": (_i, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    "dq_qnet_activation": (_i, [_vp, _i, C.POINTER(_vp), C.POINTER(_i64)]),
    "dq_adam_step": (_i, [_vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _i64, _f, _vp]),
    "dq_dqn_targets": (_i, [_vp, _vp, _vp, _vp, _f, _i64, _i, _vp, _vp]),
    "dq_dqn_loss_grad": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _vp, _vp]),
    "dq_policy_eps_greedy": (_i, [_vp, _vp, _i64, _i, _i, _u32, _u64, _u32, _vp, _dbl, _i, _vp, _vp]),
    "dq_replay_sample": (_i, [_vp, _vp, _vp, _vp, _i, _i64, _i64, _i, _i, _i, _i64, _u64, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dq_comm_create": (_i, [C.POINTER(_vp), _i, _i, _i64, _i]),
    "dq_comm_destroy": (_i, [_vp]),
    "dq_comm_handle": (_i, [_vp, _vp]),
    "dq_comm_connect": (_i, [_vp, _vp]),
    "dq_comm_next_grads": (_i, [_vp, C.POINTER(_vp)]),
    "dq_comm_allreduce_adam": (_i, [_vp, _vp, _vp, _vp, _f, _f, _f, _f, _i64, _vp]),
    "dq_comm_status": (_i, [_vp, C.POINTER(_i)]),
}


def exported_symbols():
    return sorted(_SIGNATURES)


def _share_host_cores():
    """Several ranks on one box (torchrun sets LOCAL_WORLD_SIZE): each gets its share of the CPUs for the library's host threads
    (DQ_HOST_THREADS, read once when the pool starts) instead of every rank starting one thread per core."""
    if "DQ_HOST_THREADS" in os.environ:
        return
    try:
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", "1"))
    except ValueError:
        local_world = 1
    if local_world > 1:
        cpus = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        os.environ["DQ_HOST_THREADS"] = str(max(1, cpus // local_world))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DQError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C deepq_decoding_b200/csrc` (there is no CPU fallback)" % LIB_PATH)
        _share_host_cores()
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc):
    if rc != OK:
        raise DQError("dq error %d: %s" % (rc, lib().dq_last_error().decode(errors="replace")))


def launch_count():
    return int(lib().dq_launch_count())
