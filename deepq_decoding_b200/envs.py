"""Host-side mirror of the reference environment API over the CUDA path.

  VecSurfaceCodeEnv                                  N lattices per call, tensors stay on the device
  Surface_Code_Environment_Multi_Decoding_Cycles    the reference's class name, ctor signature and
                                                     attribute set (example_notebooks/Environments.py:10-385)
                                                     as an N=1 adapter, so reference drivers run unchanged

Both call the C ABI (include/dq_decoding.h) through ctypes; torch is used for device memory and
streams only.  There is no CPU path: constructing an env without the CUDA library raises.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from . import referee as _referee


class Box:
    """Stand-in for gym.spaces.Box (gym is only used for these two spaces, Environments.py:78-84)."""

    def __init__(self, low, high, shape, dtype):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype


class Discrete:
    def __init__(self, n):
        self.n = n


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class VecSurfaceCodeEnv:
    """`n_envs` independent Surface_Code_Environment_Multi_Decoding_Cycles instances on one GPU.

    Constructor arguments up to `static_decoder` are the reference's (Environments.py:45).
    `static_decoder` may be a `referee.RefereeLUT`, None (the package's shipped table for this
    (d, error_model)), or any object with the reference's `.predict` (tabulated once).

    reset()        -> obs uint8 [N,C,H,H]
    step(actions)  -> (obs, reward float32 [N], done bool [N], info) ; info = {"lifetime": int32 [N],
                      "legal_mask": uint64-as-int64 [N,W]}.  With auto_reset (default) finished
                      lattices restart inside the call, as vectorised RL loops expect.
    The returned tensors are the env's own buffers, overwritten by the next call (the reference
    also returns its own `board_state` array, mutated in place).
    """

    def __init__(self, d=5, p_phys=0.01, p_meas=0.01, error_model="DP", use_Y=True, volume_depth=3,
                 static_decoder=None, n_envs=1, seed=0, env_id_base=0, device="cuda:0", auto_reset=True):
        if error_model not in _lib.MODEL:
            raise ValueError("specified error model not currently supported!")      # reference only prints (:67)
        if d % 2 != 1:
            raise Exception("for the surface code d must be odd!")                  # Function_Library.py:38-39
        self.L = _lib.lib()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.DQError("VecSurfaceCodeEnv needs a CUDA device (no CPU fallback)")
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", dev_index)
        self.d, self.error_model, self.use_Y, self.volume_depth = d, error_model, bool(use_Y), volume_depth
        self._p_phys, self._p_meas = float(p_phys), float(p_meas)
        self.n_envs, self.seed, self.env_id_base, self.auto_reset = int(n_envs), int(seed), int(env_id_base), bool(auto_reset)
        torch.cuda.init()
        h = C.c_void_p()
        _lib.check(self.L.dq_env_create(C.byref(h), d, _lib.MODEL[error_model], int(self.use_Y), volume_depth,
                                        self._p_phys, self._p_meas, self.n_envs, self.seed, self.env_id_base, dev_index))
        self._h = h
        info = lambda k: self._info(k)
        self.num_actions = info(_lib.INFO_NUM_ACTIONS)
        self.n_action_layers = (self.num_actions - 1) // (d * d)
        self.identity_index = self.num_actions - 1
        self.obs_channels, self.obs_side = info(_lib.INFO_OBS_CHANNELS), info(_lib.INFO_OBS_SIDE)
        self.mask_words = info(_lib.INFO_MASK_WORDS)
        self.state_words, self.state_stride = info(_lib.INFO_STATE_WORDS), info(_lib.INFO_STATE_STRIDE)
        self.observation_space = Box(low=0, high=1, shape=(self.obs_channels, self.obs_side, self.obs_side), dtype=np.uint8)
        self.action_space = Discrete(self.num_actions)
        self.multi_cycle = True

        N, dev = self.n_envs, self.device
        self.obs = torch.zeros((N, self.obs_channels, self.obs_side, self.obs_side), dtype=torch.uint8, device=dev)
        self.reward = torch.zeros(N, dtype=torch.float32, device=dev)
        self.done = torch.zeros(N, dtype=torch.uint8, device=dev)
        self.lifetime = torch.zeros(N, dtype=torch.int32, device=dev)
        self.legal_mask = torch.zeros((N, self.mask_words), dtype=torch.int64, device=dev)   # uint64 bit patterns
        self._actions = torch.zeros(N, dtype=torch.int32, device=dev)
        self._host = None
        self.static_decoder = None
        self.set_referee(static_decoder)

    # ---- plumbing ----
    def _info(self, what):
        v = C.c_int64()
        _lib.check(self.L.dq_env_info(self._h, what, C.byref(v)))
        return int(v.value)

    def close(self):
        if getattr(self, "_h", None):
            self.L.dq_env_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_referee(self, static_decoder):
        if static_decoder is None:
            static_decoder = _referee.shipped(self.d, self.error_model)
        elif not isinstance(static_decoder, _referee.RefereeLUT):
            static_decoder = _referee.from_predict(static_decoder, self.d, self.error_model)
        if static_decoder.d != self.d or static_decoder.error_model != self.error_model:
            raise ValueError("referee was built for d=%d %s" % (static_decoder.d, static_decoder.error_model))
        a, b = static_decoder.device_tables(self.device)
        self._lut_keepalive = (a, b)
        _lib.check(self.L.dq_env_set_referee_lut(self._h, static_decoder.mode, _ptr(a), a.numel(),
                                                 _ptr(b), 0 if b is None else b.numel()))
        self.static_decoder = static_decoder

    @property
    def p_phys(self):
        return self._p_phys

    @p_phys.setter
    def p_phys(self, v):
        self._p_phys = float(v)
        _lib.check(self.L.dq_env_set_noise(self._h, self._p_phys, self._p_meas))

    @property
    def p_meas(self):
        return self._p_meas

    @p_meas.setter
    def p_meas(self, v):
        self._p_meas = float(v)
        _lib.check(self.L.dq_env_set_noise(self._h, self._p_phys, self._p_meas))

    # ---- device API ----
    def reset(self, out=None):
        obs = self.obs if out is None else out
        _lib.check(self.L.dq_env_reset(self._h, _ptr(obs), _ptr(self.legal_mask), _stream(self.device)))
        self.lifetime.zero_()
        self.done.zero_()
        return obs

    def step(self, actions, out=None):
        if not (torch.is_tensor(actions) and actions.dtype == torch.int32 and actions.device == self.device
                and actions.is_contiguous()):
            actions = self._actions.copy_(torch.as_tensor(actions).to(torch.int32), non_blocking=True)
        obs = self.obs if out is None else out
        _lib.check(self.L.dq_env_step(self._h, _ptr(actions), _ptr(obs), _ptr(self.reward), _ptr(self.done),
                                      _ptr(self.lifetime), _ptr(self.legal_mask), int(self.auto_reset),
                                      _stream(self.device)))
        return obs, self.reward, self.done.bool(), {"lifetime": self.lifetime, "legal_mask": self.legal_mask}

    def random_legal_actions(self, step_index, out=None):
        act = self._actions if out is None else out
        _lib.check(self.L.dq_policy_random_legal(self._h, _ptr(self.legal_mask), int(step_index) & 0xFFFFFFFF,
                                                 _ptr(act), _stream(self.device)))
        return act

    def rollout_random(self, n_steps, obs_ring=None, first_slot=0, keep=("reward", "done", "lifetime", "actions")):
        """`n_steps` steps under the built-in uniform-random-legal policy in ONE launch (dq_env_rollout_random): the loop
        `for _ in range(n_steps): env.step(random legal action)` of the reference, bit-identical to making the steps one by
        one.  `obs_ring` uint8 [slots, N, C, H, H] (optional): step s fills slot (first_slot + s) % slots.  Returns a dict of
        per-step device tensors [n_steps, N] (legal: [n_steps, N, W]) for the names in `keep`."""
        n_steps, N, dev = int(n_steps), self.n_envs, self.device
        mk = lambda name, shape, dt: torch.empty(shape, dtype=dt, device=dev) if name in keep else None
        out = dict(reward=mk("reward", (n_steps, N), torch.float32), done=mk("done", (n_steps, N), torch.uint8),
                   lifetime=mk("lifetime", (n_steps, N), torch.int32), legal=mk("legal", (n_steps, N, self.mask_words), torch.int64),
                   actions=mk("actions", (n_steps, N), torch.int32))
        slots = 1
        if obs_ring is not None:
            if obs_ring.dtype != torch.uint8 or not obs_ring.is_contiguous() or tuple(obs_ring.shape[1:]) != tuple(self.obs.shape):
                raise ValueError("obs_ring must be a contiguous uint8 [slots, N, C, H, H] tensor")
            slots = obs_ring.shape[0]
        q = lambda x: C.c_void_p(0 if x is None else x.data_ptr())
        _lib.check(self.L.dq_env_rollout_random(self._h, n_steps, q(obs_ring), slots, int(first_slot), q(out["reward"]), q(out["done"]),
                                                q(out["lifetime"]), q(out["legal"]), q(out["actions"]), int(self.auto_reset),
                                                _stream(self.device)))
        return {k: v for k, v in out.items() if v is not None}

    # ---- host-buffer API (what the e2e benchmark times) ----
    def _host_buffers(self):
        """Pinned host buffers of the host-buffer calls, with their ctypes pointers and numpy views made ONCE (a step of 16 384 lattices
        lasts ~100 us: rebuilding six pointers and five views per call was a tenth of it)."""
        if self._host is None:
            N = self.n_envs
            pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
            hb = dict(actions=pin((N,), torch.int32), obs=pin(tuple(self.obs.shape), torch.uint8),
                      reward=pin((N,), torch.float32), done=pin((N,), torch.uint8),
                      lifetime=pin((N,), torch.int32), legal=pin((N, self.mask_words), torch.int64))
            hb["ptr"] = {k: _ptr(v) for k, v in hb.items()}
            hb["np"] = dict(actions=hb["actions"].numpy(), obs=hb["obs"].numpy(), reward=hb["reward"].numpy(),
                            done=hb["done"].numpy().view(np.bool_),          # the library writes 0 / 1
                            lifetime=hb["lifetime"].numpy(), legal=hb["legal"].numpy().view(np.uint64))
            hb["info"] = {"lifetime": hb["np"]["lifetime"], "legal_mask": hb["np"]["legal"]}
            self._host = hb
        return self._host

    def _packed_host_buffer(self):
        hb = self._host_buffers()
        if "packed" not in hb:
            hb["packed"] = torch.empty((self.state_words - ROW_BM, self.state_stride), dtype=torch.int64).pin_memory()
            hb["ptr"]["packed"] = _ptr(hb["packed"])
            hb["np"]["packed"] = hb["packed"].numpy().view(np.uint64)
        return hb["packed"]

    def reset_host(self, packed=False):
        """`packed=True`: the observations come back as the bit-packed rows the Q-network consumes (uint64
        [C*PW][state_stride], see `unpack_observations`) instead of uint8 [N, C, H, H]: 7.5x fewer bytes over PCIe.
        The returned arrays are views of the handle's pinned buffers: the next host-buffer call overwrites them."""
        hb = self._host_buffers()
        P, V = hb["ptr"], hb["np"]
        if packed:
            self._packed_host_buffer()
            _lib.check(self.L.dq_env_reset_host_packed(self._h, P["packed"], P["legal"]))
            return V["packed"], V["legal"]
        _lib.check(self.L.dq_env_reset_host(self._h, P["obs"], P["legal"]))
        return V["obs"], V["legal"]

    def step_host(self, actions, packed=False):
        hb = self._host_buffers()
        P, V = hb["ptr"], hb["np"]
        V["actions"][:] = np.asarray(actions, dtype=np.int32)
        if packed:
            self._packed_host_buffer()
            _lib.check(self.L.dq_env_step_host_packed(self._h, P["actions"], P["packed"], P["reward"], P["done"], P["lifetime"], P["legal"],
                                                      int(self.auto_reset)))
            return V["packed"], V["reward"], V["done"], hb["info"]
        _lib.check(self.L.dq_env_step_host(self._h, P["actions"], P["obs"], P["reward"], P["done"], P["lifetime"], P["legal"], int(self.auto_reset)))
        return V["obs"], V["reward"], V["done"], hb["info"]

    def step_host_begin(self, actions=None):
        """First half of `step_host` (dq_env_step_host_begin): queues the copy-in, the launch and the copy-outs and returns at once.
        `actions=None` uses what is already in the pinned action buffer (e.g. written by `random_legal_actions_host`).  Two
        environments stepped begin(A) begin(B) end(A) begin(A) end(B) ... overlap one's GPU work with the other's host-side work."""
        hb = self._host_buffers()
        P = hb["ptr"]
        if actions is not None:
            hb["np"]["actions"][:] = np.asarray(actions, dtype=np.int32)
        _lib.check(self.L.dq_env_step_host_begin(self._h, P["actions"], P["obs"], P["reward"], P["done"], P["lifetime"], P["legal"], int(self.auto_reset)))

    def step_host_end(self):
        hb = self._host_buffers()
        _lib.check(self.L.dq_env_step_host_end(self._h))
        V = hb["np"]
        return V["obs"], V["reward"], V["done"], hb["info"]

    def random_legal_actions_host(self, step_index):
        """The random-legal policy on the host (dq_policy_random_legal_host): picks from the legal masks of the latest host-buffer
        call into the pinned action buffer, which is returned (and is what `step_host_begin()` sends)."""
        hb = self._host_buffers()
        _lib.check(self.L.dq_policy_random_legal_host(self._h, hb["ptr"]["legal"], int(step_index) & 0xFFFFFFFF, hb["ptr"]["actions"]))
        return hb["np"]["actions"]

    # ---- packed state (parity tests, checkpoints) ----
    def get_state_words(self):
        w = torch.empty((self.state_words, self.state_stride), dtype=torch.int64, device=self.device)
        _lib.check(self.L.dq_env_get_state(self._h, _ptr(w), _stream(self.device)))
        return w

    def set_state_words(self, w):
        w = w.to(self.device, torch.int64).contiguous()
        assert tuple(w.shape) == (self.state_words, self.state_stride)
        _lib.check(self.L.dq_env_set_state(self._h, _ptr(w), _stream(self.device)))
        torch.cuda.current_stream(self.device).synchronize()

    def decode_state(self, i=None):
        """Packed words -> the reference's attributes, as numpy (DESIGN.md section 2 layout)."""
        w = self.get_state_words().cpu().numpy().view(np.uint64)
        idx = range(self.n_envs) if i is None else [i]
        out = [decode_state_column(w[:, k], self.d, self.volume_depth, self.n_action_layers, self.error_model, self.use_Y)
               for k in idx]
        return out if i is None else out[0]

    def legal_actions(self, i=0):
        m = self.legal_mask[i].cpu().numpy().view(np.uint64)
        return {a for a in range(self.num_actions) if (int(m[a >> 6]) >> (a & 63)) & 1}

    def render(self, i=0, mode="ansi"):
        """ASCII picture of lattice i: hidden Pauli frame beside the latest faulty syndrome slice.
        (The reference env has no render(); this is the gym method the north star asks for.)"""
        st = self.decode_state(i)
        d = self.d
        sym = ".XYZ"
        lines = ["lattice %d  lifetime=%d done=%s" % (i, st["lifetime"], st["done"])]
        last = st["faulty_syndromes"][-1]
        for r in range(d + 1):
            srow = " ".join("#" if last[r, c] else ("+" if _referee.plaquette_present(d, r, c) else " ") for c in range(d + 1))
            lines.append("  " + srow)
            if r < d:
                lines.append("   " + " ".join(sym[int(v)] for v in st["hidden_state"][r]))
        text = "\n".join(lines)
        if mode == "human":
            print(text)
        return text


ROW_XB, ROW_ZB, ROW_META, ROW_ACT, ROW_SUM, ROW_BM = 0, 1, 2, 3, 6, 7


def unpack_observations(rows, n_envs, d, channels):
    """Packed observation rows (uint64 [channels*PW][stride]; bit x*H+y of layer l of lattice i lives in rows[l*PW + word][i])
    -> the reference's board_state layout, uint8 [n_envs, channels, H, H] (EN/Environments.py:273-314)."""
    H = 2 * d + 1
    P = H * H
    pw = (P + 63) // 64
    rows = np.ascontiguousarray(np.asarray(rows).view(np.uint64)[:, :n_envs])
    assert rows.shape[0] == channels * pw
    by_lattice = np.ascontiguousarray(rows.reshape(channels, pw, n_envs).transpose(2, 0, 1))       # [n, C, PW]
    bits = np.unpackbits(by_lattice.view(np.uint8).reshape(n_envs, channels, pw * 8), axis=2, bitorder="little")
    return np.ascontiguousarray(bits[:, :, :P]).reshape(n_envs, channels, H, H)


def _grid(word, g, rows, cols):
    word = int(word)
    return np.array([[(word >> (r * g + c)) & 1 for c in range(cols)] for r in range(rows)], dtype=np.int64)


def decode_state_column(col, d, vd, layers, error_model, use_Y):
    """One column of the packed state matrix -> the reference's attributes."""
    g, H = d + 1, 2 * d + 1
    pw = (H * H + 63) // 64
    xb, zb = _grid(col[ROW_XB], g, d, d), _grid(col[ROW_ZB], g, d, d)
    hidden = np.where(xb & zb, 2, np.where(xb, 1, np.where(zb, 3, 0)))
    meta = int(col[ROW_META])
    completed = np.zeros(layers * d * d + 1, np.int64)
    for l in range(layers):
        completed[l * d * d:(l + 1) * d * d] = _grid(col[ROW_ACT + l], g, d, d).reshape(-1)
    board = np.zeros((vd + layers, H, H), np.int64)            # the cached, rendered observation layers
    for l in range(vd + layers):
        big = sum(int(col[ROW_BM + l * pw + i]) << (64 * i) for i in range(pw))
        board[l] = np.array([(big >> k) & 1 for k in range(H * H)]).reshape(H, H)
    faulty = board[:vd, ::2, ::2].copy()
    return dict(hidden_state=hidden, completed_actions=completed, faulty_syndromes=faulty, board_state=board,
                summed_syndrome_volume=_grid(col[ROW_SUM], g, g, g),
                lifetime=meta & 0xFFFFFFFF, attempts=(meta >> 32) & 0x7FFFFFFF, done=bool(meta >> 63))


class Surface_Code_Environment_Multi_Decoding_Cycles:
    """The reference's environment class, backed by the CUDA kernel (one lattice).

    Same constructor, `reset()`, `step(action) -> (board_state, reward, done, {})` and attribute
    set as example_notebooks/Environments.py:10-385: num_actions, identity_index, n_action_layers,
    observation_space, action_space, board_state, hidden_state, current_true_syndrome,
    completed_actions, legal_actions (set), lifetime, done, p_phys / p_meas (assignable),
    padding_syndrome, padding_actions, indicate_identity, multi_cycle.  Extra keyword arguments
    (seed, env_id, device) choose the random stream and GPU.  `static_decoder` may be a
    `RefereeLUT`, an object with the reference's `.predict`, or None for the shipped table.
    """

    def __init__(self, d=5, p_phys=0.01, p_meas=0.01, error_model="DP", use_Y=True, volume_depth=3,
                 static_decoder=None, seed=0, env_id=0, device="cuda:0"):
        self._vec = VecSurfaceCodeEnv(d, p_phys, p_meas, error_model, use_Y, volume_depth, static_decoder,
                                      n_envs=1, seed=seed, env_id_base=env_id, device=device, auto_reset=False)
        v = self._vec
        self.d, self.error_model, self.use_Y, self.volume_depth = d, error_model, use_Y, volume_depth
        self.static_decoder = v.static_decoder
        self.num_actions, self.n_action_layers, self.identity_index = v.num_actions, v.n_action_layers, v.identity_index
        self.observation_space, self.action_space = v.observation_space, v.action_space
        self.identity_indicator = self.generate_identity_indicator(d)
        self.board_state = np.zeros(v.observation_space.shape, int)
        self.hidden_state = np.zeros((d, d), int)
        self.current_true_syndrome = np.zeros((d + 1, d + 1), int)
        self.completed_actions = np.zeros(self.num_actions, int)
        self.legal_actions = set()
        self.done = False
        self.lifetime = 0
        self.multi_cycle = True

    p_phys = property(lambda self: self._vec.p_phys, lambda self, v: setattr(self._vec, "p_phys", v))
    p_meas = property(lambda self: self._vec.p_meas, lambda self, v: setattr(self._vec, "p_meas", v))

    def _sync(self, obs, legal):
        self.board_state[...] = obs[0]
        m = legal[0]
        self.legal_actions = {a for a in range(self.num_actions) if (int(m[a >> 6]) >> (a & 63)) & 1}
        st = self._vec.decode_state(0)
        self.hidden_state = st["hidden_state"]
        self.completed_actions = st["completed_actions"]
        self.lifetime = int(st["lifetime"])
        self.done = bool(st["done"])
        self.current_true_syndrome = true_syndrome_of(self.hidden_state)

    def reset(self):
        obs, legal = self._vec.reset_host()
        self._sync(obs, legal)
        return self.board_state

    def step(self, action):
        obs, reward, done, info = self._vec.step_host(np.array([int(action)], np.int32))
        self._sync(obs, info["legal_mask"])
        return self.board_state, float(reward[0]), self.done, {}

    def render(self, mode="ansi"):
        return self._vec.render(0, mode)

    def close(self):
        self._vec.close()

    # -- constant-layout helpers kept for API parity (Environments.py:273-324, 374-385) --
    def padding_syndrome(self, syndrome_in):
        d = self.d
        out = np.zeros((2 * d + 1, 2 * d + 1), int)
        x = np.arange(2 * d + 1)
        odd = x % 2 == 1
        out[0, odd] = out[2 * d, odd] = 1
        out[odd, 0] = out[odd, 2 * d] = 1
        xx, yy = np.meshgrid(x, x, indexing="ij")
        out[(xx % 2 == 1) & (yy % 2 == 1) & ((xx + yy) % 4 == 0)] = 1
        out[::2, ::2] = np.asarray(syndrome_in)
        return out

    def padding_actions(self, actions_in):
        d = self.d
        out = np.zeros((2 * d + 1, 2 * d + 1), int)
        for idx, taken in enumerate(actions_in):
            if taken:
                out[2 * (idx // d) + 1, 2 * (idx % d) + 1] = 1
        return out

    def indicate_identity(self, board_state):
        for k in range(self.n_action_layers):
            board_state[self.volume_depth + k] = board_state[self.volume_depth + k] + self.identity_indicator
        return board_state

    def generate_identity_indicator(self, d):
        ind = np.ones((2 * d + 1, 2 * d + 1), int)
        ind[1::2, 1::2] = 0
        return ind


def true_syndrome_of(hidden):
    """Host-side syndrome of a Pauli frame (Function_Library.py:162-184), for attribute parity only."""
    hidden = np.asarray(hidden)
    d = hidden.shape[0]
    xb = ((hidden == 1) | (hidden == 2)).astype(np.int64)
    zb = ((hidden == 2) | (hidden == 3)).astype(np.int64)
    syn = np.zeros((d + 1, d + 1), int)
    for a in range(d + 1):
        for b in range(d + 1):
            if not _referee.plaquette_present(d, a, b):
                continue
            plane = xb if (a + b) % 2 == 1 else zb
            syn[a, b] = plane[max(a - 1, 0):min(a + 1, d), max(b - 1, 0):min(b + 1, d)].sum() % 2
    return syn
