"""Q-network: host-side mirror of the reference's model builder over the CUDA kernels.

`build_convolutional_nn(cc_layers, ff_layers, input_shape, num_actions)` has the signature of
example_notebooks/Function_Library.py:338-377 and returns a `QNetwork` instead of a Keras model; the
dueling head keras-rl adds (`enable_dueling_network=True`, dueling_type 'avg') is part of the network
here.  Weights live in ONE flat fp32 device buffer (so the optimizer, the hard target copy and the
gradient all-reduce are single kernels / single collectives); `load_weights` / `save_weights` read and
write the Keras HDF5 layout of `final_dqn_weights.h5f`, so the reference's agents run here and
agents trained here load back into Keras.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib, h5lite


class QNetwork:
    def __init__(self, cc_layers, ff_layers, input_shape, num_actions, dueling=True, max_batch=4096, device="cuda:0", seed=0):
        """cc_layers: [[filters, kernel, stride], ...]; ff_layers: [[units, dropout_rate], ...] (reference convention)."""
        self.cc_layers = [[int(v) for v in l] for l in cc_layers]
        self.ff_layers = [[int(l[0]), float(l[1])] for l in ff_layers]
        self.input_shape = tuple(int(v) for v in input_shape)
        C_in, H, W_ = self.input_shape
        if H != W_:
            raise ValueError("square observations only")
        self.num_actions, self.dueling, self.max_batch = int(num_actions), bool(dueling), int(max_batch)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.DQError("QNetwork needs a CUDA device (no CPU fallback)")
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        self.L = _lib.lib()
        arr = lambda vals, t: (t * len(vals))(*vals)
        f = arr([l[0] for l in self.cc_layers], C.c_int)
        k = arr([l[1] for l in self.cc_layers], C.c_int)
        s = arr([l[2] for l in self.cc_layers], C.c_int)
        u = arr([l[0] for l in self.ff_layers], C.c_int) if self.ff_layers else None
        dr = arr([l[1] for l in self.ff_layers], C.c_float) if self.ff_layers else None
        h = C.c_void_p()
        torch.cuda.init()
        _lib.check(self.L.dq_qnet_create(C.byref(h), C_in, H, len(self.cc_layers), f, k, s, len(self.ff_layers), u, dr,
                                         self.num_actions, int(self.dueling), self.max_batch, idx))
        self._h = h
        self.num_params = self._info(_lib.QINFO_NUM_PARAMS)
        self.packed_rows = self._info(_lib.QINFO_PACKED_ROWS)
        self.flops_per_sample = self._info(_lib.QINFO_FLOPS_PER_SAMPLE)
        nt = self._info(_lib.QINFO_NUM_TENSORS) // 2
        off = (C.c_int64 * (2 * nt))()
        shp = (C.c_int64 * (2 * nt))()
        _lib.check(self.L.dq_qnet_param_layout(self._h, off, shp))
        self.layout = [(int(off[2 * t]), int(off[2 * t + 1]), int(shp[2 * t]), int(shp[2 * t + 1])) for t in range(nt)]
        self.params = torch.zeros(self.num_params, dtype=torch.float32, device=self.device)
        self._q = torch.zeros((self.max_batch, self.num_actions), dtype=torch.float32, device=self.device)
        self._packed = None
        self._tc_dirty = True            # bf16 weight copies of the tensor-core path are stale
        # geometry of the last conv output, for the Flatten permutation
        side, c = H, C_in
        for filt, ksz, st in self.cc_layers:
            side, c = (side - ksz) // st + 1, filt
        self._flat_c, self._flat_p = c, side * side
        self.init_glorot(seed)

    # ---- plumbing -------------------------------------------------------------------------------
    def _info(self, what):
        v = C.c_int64()
        _lib.check(self.L.dq_qnet_info(self._h, what, C.byref(v)))
        return int(v.value)

    def close(self):
        if getattr(self, "_h", None):
            self.L.dq_qnet_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights --------------------------------------------------------------------------------
    def _flatten_perm(self):
        """Row permutation of the first dense kernel: ours is (position, channel), Keras' Flatten is (channel, position)."""
        p, c = np.meshgrid(np.arange(self._flat_p), np.arange(self._flat_c), indexing="ij")
        return (c * self._flat_p + p).reshape(-1)        # ours[row] = keras[perm[row]]

    def set_keras_weights(self, conv, dense):
        """conv: [(HWIO kernel, bias)], dense: [(in,out kernel, bias)] in Keras layouts (dueling head last)."""
        flat = np.zeros(self.num_params, np.float32)
        pairs = [(np.asarray(k, np.float32).reshape(-1, k.shape[-1]), b) for k, b in conv]
        for i, (k, b) in enumerate(dense):
            k = np.asarray(k, np.float32)
            if i == 0:
                k = k[self._flatten_perm()]
            pairs.append((k, b))
        if len(pairs) != len(self.layout):
            raise ValueError("expected %d weight tensors pairs, got %d" % (len(self.layout), len(pairs)))
        for (wo, bo, K, N), (k, b) in zip(self.layout, pairs):
            if k.shape != (K, N) or np.asarray(b).shape != (N,):
                raise ValueError("weight shape %s does not match layer (%d,%d)" % (k.shape, K, N))
            flat[wo:wo + K * N] = k.reshape(-1)
            flat[bo:bo + N] = np.asarray(b, np.float32)
        self.params.copy_(torch.from_numpy(flat))
        self.params_changed()

    def get_keras_weights(self):
        flat = self.params.detach().cpu().numpy()
        conv, dense = [], []
        cin = self.input_shape[0]
        for t, (wo, bo, K, N) in enumerate(self.layout):
            k, b = flat[wo:wo + K * N].reshape(K, N).copy(), flat[bo:bo + N].copy()
            if t < len(self.cc_layers):
                ksz = self.cc_layers[t][1]
                conv.append((k.reshape(ksz, ksz, cin, N), b))
                cin = N
            else:
                if t == len(self.cc_layers):
                    inv = np.empty_like(self._flatten_perm())
                    inv[self._flatten_perm()] = np.arange(len(inv))
                    k = k[inv]
                dense.append((k, b))
        return conv, dense

    def init_glorot(self, seed=0):
        """Keras defaults: glorot_uniform kernels, zero biases."""
        rng = np.random.default_rng(seed)
        conv, dense, cin = [], [], self.input_shape[0]
        for t, (wo, bo, K, N) in enumerate(self.layout):
            if t < len(self.cc_layers):
                ksz = self.cc_layers[t][1]
                lim = np.sqrt(6.0 / (K + N * ksz * ksz))
                conv.append((rng.uniform(-lim, lim, size=(ksz, ksz, cin, N)).astype(np.float32), np.zeros(N, np.float32)))
                cin = N
            else:
                lim = np.sqrt(6.0 / (K + N))
                dense.append((rng.uniform(-lim, lim, size=(K, N)).astype(np.float32), np.zeros(N, np.float32)))
        self.set_keras_weights(conv, dense)

    def load_weights(self, path):
        """Keras / keras-rl HDF5 weight file (e.g. trained_models/d5_dp/0.007/final_dqn_weights.h5f)."""
        f = h5lite.H5File(path)
        names = f.keys("/")
        conv_n = sorted((k for k in names if k.startswith("conv2d_") and f.keys("/" + k)), key=lambda s: int(s.split("_")[1]))
        dense_n = sorted((k for k in names if k.startswith("dense_") and f.keys("/" + k)), key=lambda s: int(s.split("_")[1]))
        get = lambda n: (lambda g: (f[g + "/kernel:0"], f[g + "/bias:0"]))("/%s/%s" % (n, f.keys("/" + n)[0]))
        self.set_keras_weights([get(n) for n in conv_n], [get(n) for n in dense_n])

    def save_weights(self, path, overwrite=True):
        conv, dense = self.get_keras_weights()
        tree, names = {}, []
        for i, (k, b) in enumerate(conv):
            n = "conv2d_%d" % (i + 1)
            names.append(n)
            tree[n] = {"@weight_names": [n + "/kernel:0", n + "/bias:0"], n: {"kernel:0": k, "bias:0": b}}
        for i, (k, b) in enumerate(dense):
            n, sub = "dense_%d" % (i + 1), "dense_%d_1" % (i + 1)
            names.append(n)
            tree[n] = {"@weight_names": [sub + "/kernel:0", sub + "/bias:0"], sub: {"kernel:0": k, "bias:0": b}}
        tree["@layer_names"] = names
        tree["@backend"] = "tensorflow"
        tree["@keras_version"] = "2.2.2"
        h5lite.write_h5(path, tree)

    # ---- compute --------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def pack(self, obs_u8):
        """uint8 [B,C,H,W] (device) -> packed [rows, B] int64 bit patterns."""
        B = obs_u8.shape[0]
        out = torch.empty((self.packed_rows, B), dtype=torch.int64, device=self.device)
        _lib.check(self.L.dq_qnet_pack_obs(self._h, C.c_void_p(obs_u8.data_ptr()), C.c_void_p(out.data_ptr()), B, B, self._stream()))
        return out

    def forward_packed(self, packed_ptr, stride, batch, out=None, train=False, dropout_seed=0, params=None, precision="fp32"):
        """precision: "fp32" (SIMT kernels, Keras-fp32 arithmetic) or "bf16" (tcgen05 tensor cores, fp32 accumulation)."""
        q = self._q[:batch] if out is None else out
        p = self.params if params is None else params
        if precision == "bf16":
            if params is not None:
                raise ValueError("the bf16 tensor-core path uses the network's own parameters (its staged bf16 copies)")
            if self._tc_dirty:
                _lib.check(self.L.dq_qnet_prepare_tc(self._h, C.c_void_p(p.data_ptr()), self._stream()))
                self._tc_dirty = False
            if train:           # mixed-precision training forward: dropout applied, activations kept for backward_packed(precision="bf16")
                _lib.check(self.L.dq_qnet_forward_tc_train(self._h, C.c_void_p(p.data_ptr()), C.c_void_p(packed_ptr), stride, batch,
                                                           C.c_void_p(q.data_ptr()), int(dropout_seed), self._stream()))
                return q
            _lib.check(self.L.dq_qnet_forward_tc(self._h, C.c_void_p(p.data_ptr()), C.c_void_p(packed_ptr), stride, batch,
                                                 C.c_void_p(q.data_ptr()), self._stream()))
            return q
        _lib.check(self.L.dq_qnet_forward(self._h, C.c_void_p(p.data_ptr()), C.c_void_p(packed_ptr), stride, batch,
                                          C.c_void_p(q.data_ptr()), int(train), int(dropout_seed), self._stream()))
        return q

    def forward(self, obs, train=False, dropout_seed=0, precision="fp32"):
        """Q values for uint8/bool/int observations [B,C,H,W] (numpy or torch) -- model.predict_on_batch."""
        t = torch.as_tensor(obs)
        if t.dim() == 3:
            t = t[None]
        t = t.to(self.device).to(torch.uint8).contiguous()
        packed = self.pack(t)
        self._packed = packed
        return self.forward_packed(packed.data_ptr(), t.shape[0], t.shape[0], train=train, dropout_seed=dropout_seed, precision=precision)

    def backward(self, dq, grads=None, precision="fp32"):
        """Gradient (flat fp32) of sum(dq * Q) for the batch of the last forward(train=True) call."""
        g = torch.zeros(self.num_params, dtype=torch.float32, device=self.device) if grads is None else grads
        B = self._packed.shape[1]
        return self.backward_packed(self._packed.data_ptr(), B, B, dq.contiguous(), g, precision=precision)

    def params_changed(self):
        """Tell the network its flat parameter buffer was modified in place (optimizer step, copy).  Handles that share the buffer
        (`share_params_with`) are told too."""
        self._tc_dirty = True
        for other in getattr(self, "_sharers", ()):
            other._tc_dirty = True

    def share_params_with(self, owner):
        """Make this handle a second view of `owner`'s parameters (own activation buffers and staged bf16 weights, same flat fp32
        buffer): what lets two forwards of one network run side by side on two streams."""
        self.params = owner.params
        self._tc_dirty = True
        if not hasattr(owner, "_sharers"):
            owner._sharers = []
        owner._sharers.append(self)

    def backward_packed(self, packed_ptr, stride, batch, dq, grads, params=None, precision="fp32"):
        p = self.params if params is None else params
        if precision == "bf16":         # after forward_packed(train=True, precision="bf16") on the same batch
            _lib.check(self.L.dq_qnet_backward_tc(self._h, C.c_void_p(p.data_ptr()), C.c_void_p(packed_ptr), stride, batch,
                                                  C.c_void_p(dq.data_ptr()), C.c_void_p(grads.data_ptr()), self._stream()))
            return grads
        _lib.check(self.L.dq_qnet_backward(self._h, C.c_void_p(p.data_ptr()), C.c_void_p(packed_ptr), stride, batch,
                                           C.c_void_p(dq.data_ptr()), C.c_void_p(grads.data_ptr()), self._stream()))
        return grads

    def activation(self, index, batch):
        """Post-activation output of layer `index` of the last forward, [batch, per_sample] (tests)."""
        ptr, per = C.c_void_p(), C.c_int64()
        _lib.check(self.L.dq_qnet_activation(self._h, index, C.byref(ptr), C.byref(per)))
        return device_view(ptr.value, (batch, per.value), "<f4", self.device).clone()


    def tc_activation(self, index, batch):
        """bf16 activation `index` of the last tensor-core forward as float32 [batch, per_sample] (tests)."""
        ptr, per = C.c_void_p(), C.c_int64()
        _lib.check(QNetwork._L().dq_qnet_tc_activation(self._h, index, C.byref(ptr), C.byref(per)))
        raw = device_view(ptr.value, (batch, per.value), "<i2", self.device)
        return raw.view(torch.bfloat16).float()

    @staticmethod
    def _L():
        return _lib.lib()


class _RawDeviceArray:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def device_view(ptr, shape, typestr, device):
    """torch view of library-owned device memory (no copy)."""
    return torch.as_tensor(_RawDeviceArray(ptr, shape, typestr), device=device)


def build_convolutional_nn(cc_layers, ff_layers, input_shape, num_actions, **kw):
    """Reference signature (example_notebooks/Function_Library.py:338); returns a QNetwork."""
    return QNetwork(cc_layers, ff_layers, input_shape, num_actions, **kw)
