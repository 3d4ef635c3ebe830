"""TEST INFRASTRUCTURE: the fp32 SIMT half of the Q-network / DQN C ABI (dq_qnet.cu up to the tensor-core path)
executed on the CPU by the kernel emulation (tests/host/cuda_emu.h, tests/host/emu_build.py), on numpy buffers.
The bf16 tcgen05 path cannot run here and is not part of this library.  Nothing under deepq_decoding_b200/ imports this.
"""
import ctypes as C
import os

import numpy as np

import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "host"))
import emu_build as B  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "host", "libdq_qnet_emu.so")
CUT = ("// ================================================================================================\n"
       "// bf16 tensor-core inference path (acting)")
TAIL = "static void tc_free(dq_qnet*) {}\n"
QINFO_NUM_PARAMS, QINFO_PACKED_ROWS, QINFO_NUM_TENSORS, QINFO_FLOPS_PER_SAMPLE = range(4)

OUT_TC = os.path.join(HERE, "host", "libdq_qnet_tc_emu.so")
_HOST = os.path.join(HERE, "host")
TC_SWAPS = (("// [tcgen05 kernels: begin]", "// [tcgen05 kernels: end]", open(os.path.join(_HOST, "tc_ref_gemm.inc")).read()),
            ("// [tcgen05 kernels: begin]", "// [tcgen05 kernels: end]", open(os.path.join(_HOST, "tc_ref_dw.inc")).read()))

_L = None
_LTC = None


def lib_tc():
    """The WHOLE of dq_qnet.cu on the CPU, with the tcgen05 kernels swapped for plain loops (tests/host/tc_emu.h): the bf16
    forward / backward drivers and every kernel around the GEMMs run as written."""
    global _LTC
    if _LTC is None:
        deps = [os.path.join(_HOST, f) for f in ("tc_emu.h", "tc_ref_gemm.inc", "tc_ref_dw.inc")]
        path = B.build(OUT_TC, [os.path.join(B.CSRC, "dq_env.cu"), (os.path.join(B.CSRC, "dq_qnet.cu"), None, "", TC_SWAPS)],
                       extra_flags=("-include", os.path.join(_HOST, "tc_emu.h")), deps=deps)
        L = C.CDLL(path)
        _signatures(L)
        vp, i, i64, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64
        L.dq_qnet_prepare_tc.argtypes = [vp, vp, vp]
        L.dq_qnet_forward_tc.argtypes = [vp, vp, vp, i64, i64, vp, vp]
        L.dq_qnet_forward_tc_train.argtypes = [vp, vp, vp, i64, i64, vp, u64, vp]
        L.dq_qnet_backward_tc.argtypes = [vp, vp, vp, i64, i64, vp, vp, vp]
        _LTC = L
    return _LTC


def _signatures(L):
    vp, i, i64, u64, u32, f, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_uint32, C.c_float, C.c_double
    L.dq_last_error.restype = C.c_char_p
    L.dq_qnet_create.argtypes = [C.POINTER(vp), i, i, i, vp, vp, vp, i, vp, vp, i, i, i64, i]
    L.dq_qnet_destroy.argtypes = [vp]
    L.dq_qnet_info.argtypes = [vp, i, C.POINTER(i64)]
    L.dq_qnet_param_layout.argtypes = [vp, vp, vp]
    L.dq_qnet_pack_obs.argtypes = [vp, vp, vp, i64, i64, vp]
    L.dq_qnet_forward.argtypes = [vp, vp, vp, i64, i64, vp, i, u64, vp]
    L.dq_qnet_backward.argtypes = [vp, vp, vp, i64, i64, vp, vp, vp]
    L.dq_qnet_activation.argtypes = [vp, i, C.POINTER(vp), C.POINTER(i64)]


def lib():
    global _L
    if _L is None:
        path = B.build(OUT, [os.path.join(B.CSRC, "dq_env.cu"), (os.path.join(B.CSRC, "dq_qnet.cu"), CUT, TAIL)])
        L = C.CDLL(path)
        vp, i, i64, u64, u32, f, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_uint32, C.c_float, C.c_double
        L.dq_last_error.restype = C.c_char_p
        L.dq_qnet_create.argtypes = [C.POINTER(vp), i, i, i, vp, vp, vp, i, vp, vp, i, i, i64, i]
        L.dq_qnet_destroy.argtypes = [vp]
        L.dq_qnet_info.argtypes = [vp, i, C.POINTER(i64)]
        L.dq_qnet_param_layout.argtypes = [vp, vp, vp]
        L.dq_qnet_pack_obs.argtypes = [vp, vp, vp, i64, i64, vp]
        L.dq_qnet_forward.argtypes = [vp, vp, vp, i64, i64, vp, i, u64, vp]
        L.dq_qnet_backward.argtypes = [vp, vp, vp, i64, i64, vp, vp, vp]
        L.dq_qnet_fold_head.argtypes = [vp, vp, vp, vp, vp]
        L.dq_qnet_activation.argtypes = [vp, i, C.POINTER(vp), C.POINTER(i64)]
        L.dq_adam_step.argtypes = [vp, vp, vp, vp, i64, f, f, f, f, i64, f, vp]
        L.dq_dqn_targets.argtypes = [vp, vp, vp, vp, f, i64, i, vp, vp]
        L.dq_dqn_loss_grad.argtypes = [vp, vp, vp, i64, i, vp, vp, vp]
        L.dq_policy_eps_greedy.argtypes = [vp, vp, i64, i, i, u32, u64, u32, vp, dbl, i, vp, vp]
        L.dq_replay_sample.argtypes = [vp, vp, vp, vp, i, i64, i64, i, i, i, i64, u64, u32, vp, vp, vp, vp, vp, vp, vp]
        _L = L
    return _L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def check(rc, L=None):
    if rc != 0:
        raise RuntimeError("emulated ABI call failed (%d): %s" % (rc, (L or lib()).dq_last_error().decode()))


class EmuQNet:
    """Same construction arguments as deepq_decoding_b200.qnet.QNetwork; weights in Keras layouts in and out."""

    def __init__(self, cc_layers, ff_layers, input_shape, num_actions, dueling=True, max_batch=256, tc=False):
        self.L = lib_tc() if tc else lib()
        self.cc, self.ff = [list(map(int, l)) for l in cc_layers], [[int(l[0]), float(l[1])] for l in ff_layers]
        self.C_in, self.H = int(input_shape[0]), int(input_shape[1])
        self.A, self.max_batch = int(num_actions), int(max_batch)
        arr = lambda vals, t: np.asarray(vals, dtype=t)
        self._keep = (arr([l[0] for l in self.cc], np.int32), arr([l[1] for l in self.cc], np.int32), arr([l[2] for l in self.cc], np.int32),
                      arr([l[0] for l in self.ff], np.int32), arr([l[1] for l in self.ff], np.float32))
        f, k, s, u, dr = self._keep
        h = C.c_void_p()
        check(self.L.dq_qnet_create(C.byref(h), self.C_in, self.H, len(self.cc), _p(f), _p(k), _p(s), len(self.ff), _p(u), _p(dr),
                                    self.A, int(dueling), self.max_batch, 0))
        self.h = h
        self.num_params = self._info(QINFO_NUM_PARAMS)
        self.rows = self._info(QINFO_PACKED_ROWS)
        nt = self._info(QINFO_NUM_TENSORS) // 2
        off, shp = np.zeros(2 * nt, np.int64), np.zeros(2 * nt, np.int64)
        check(self.L.dq_qnet_param_layout(self.h, _p(off), _p(shp)))
        self.layout = [(int(off[2 * t]), int(off[2 * t + 1]), int(shp[2 * t]), int(shp[2 * t + 1])) for t in range(nt)]
        side, c = self.H, self.C_in
        for filt, ksz, st in self.cc:
            side, c = (side - ksz) // st + 1, filt
        self.flat_c, self.flat_p = c, side * side
        self.params = np.zeros(self.num_params, np.float32)

    def _info(self, what):
        v = C.c_int64()
        check(self.L.dq_qnet_info(self.h, what, C.byref(v)))
        return int(v.value)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.dq_qnet_destroy(self.h)
            self.h = None

    def flatten_perm(self):      # ours[row] = keras[perm[row]]: (position, channel) vs Keras' Flatten (channel, position)
        p, c = np.meshgrid(np.arange(self.flat_p), np.arange(self.flat_c), indexing="ij")
        return (c * self.flat_p + p).reshape(-1)

    def set_keras_weights(self, conv, dense):
        pairs = [(np.asarray(k, np.float32).reshape(-1, k.shape[-1]), b) for k, b in conv]
        for i, (k, b) in enumerate(dense):
            k = np.asarray(k, np.float32)
            pairs.append((k[self.flatten_perm()] if i == 0 else k, b))
        assert len(pairs) == len(self.layout)
        for (wo, bo, K, N), (k, b) in zip(self.layout, pairs):
            assert k.shape == (K, N)
            self.params[wo:wo + K * N] = k.reshape(-1)
            self.params[bo:bo + N] = np.asarray(b, np.float32)

    def keras_grads(self, flat):
        """flat gradient buffer -> tensors in the order / layouts of TorchQNet.parameters() (conv OIHW, dense (in,out))."""
        out, nconv = [], len(self.cc)
        cin = self.C_in
        for t, (wo, bo, K, N) in enumerate(self.layout):
            w = flat[wo:wo + K * N].reshape(K, N)
            if t < nconv:
                ksz = self.cc[t][1]
                w = w.reshape(ksz, ksz, cin, N).transpose(3, 2, 0, 1)          # HWIO -> OIHW
                cin = N
            elif t == nconv:
                inv = np.empty(K, np.int64); inv[self.flatten_perm()] = np.arange(K)
                w = w[inv]
            out += [np.ascontiguousarray(w), flat[bo:bo + N].copy()]
        return out

    def pack(self, obs_u8):
        obs_u8 = np.ascontiguousarray(obs_u8, dtype=np.uint8)
        b = obs_u8.shape[0]
        packed = np.zeros((self.rows, b), np.uint64)
        check(self.L.dq_qnet_pack_obs(self.h, _p(obs_u8), _p(packed), b, b, None))
        return packed

    def forward(self, obs_u8, train=False, dropout_seed=0):
        packed = self.pack(obs_u8)
        b = packed.shape[1]
        q = np.zeros((b, self.A), np.float32)
        check(self.L.dq_qnet_forward(self.h, _p(self.params), _p(packed), b, b, _p(q), int(train), dropout_seed, None))
        return q, packed

    def forward_tc(self, packed, train=False, dropout_seed=0):
        """bf16 path (tc=True handles only): prepare + forward; train keeps the head unfolded and applies dropout."""
        b = packed.shape[1]
        q = np.zeros((b, self.A), np.float32)
        check(self.L.dq_qnet_prepare_tc(self.h, _p(self.params), None), self.L)
        if train:
            check(self.L.dq_qnet_forward_tc_train(self.h, _p(self.params), _p(packed), b, b, _p(q), dropout_seed, None), self.L)
        else:
            check(self.L.dq_qnet_forward_tc(self.h, _p(self.params), _p(packed), b, b, _p(q), None), self.L)
        return q

    def backward_tc(self, packed, dq):
        b = packed.shape[1]
        dq = np.ascontiguousarray(dq, dtype=np.float32)
        grads = np.zeros(self.num_params, np.float32)
        check(self.L.dq_qnet_backward_tc(self.h, _p(self.params), _p(packed), b, b, _p(dq), _p(grads), None), self.L)
        return grads

    def fold_head(self):
        """(w [K][A], b [A]) with Q = h @ w + b for the output h of the last hidden dense layer (dq_qnet_fold_head)."""
        K = self.layout[-2][2]
        w, b = np.zeros((K, self.A), np.float32), np.zeros(self.A, np.float32)
        check(self.L.dq_qnet_fold_head(self.h, _p(self.params), _p(w), _p(b), None))
        return w, b

    def backward(self, packed, dq):
        b = packed.shape[1]
        dq = np.ascontiguousarray(dq, dtype=np.float32)
        grads = np.zeros(self.num_params, np.float32)
        check(self.L.dq_qnet_backward(self.h, _p(self.params), _p(packed), b, b, _p(dq), _p(grads), None))
        return grads
