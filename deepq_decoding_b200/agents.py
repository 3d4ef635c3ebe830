"""keras-rl surface used by the reference (`rl.agents.dqn.DQNAgent`, `rl.policy.*`, `rl.memory.SequentialMemory`,
`rl.callbacks.FileLogger`, `keras.optimizers.Adam`) over the CUDA path.

The reference drives a FORK of keras-rl whose source is not in the reference repo; the semantics below are
upstream keras-rl 0.4.x plus the fork's extras as reconstructed in SURVEY.md (sections 3.1, 3.3, 8a rows 16-20,
Appendix B) from the call sites (cluster_scripts/d5_dp/0.001/Single_Point_Training_Script.py:109-207), the
stdout of README.md:408-481 and the shipped training_history.json files.

Vectorisation.  `env` may be a `VecSurfaceCodeEnv` with N lattices (or the reference-named N=1 adapter).  One
agent iteration acts on all N lattices at once; `step` counters, warm-up, epsilon annealing, target-network period
and `nb_steps` are all measured in ENV TRANSITIONS (N per iteration), so the reference's hyper-parameters keep
their meaning and N=1 reproduces the reference loop order exactly: act -> step -> store -> sample -> update.
The whole iteration runs on the device (packed observations, no byte boards); the host only launches kernels and,
every `flush_interval` iterations, drains the per-step (reward, done, lifetime) rows to do episode bookkeeping.
"""
import ctypes as C
import json
import math
import time

import numpy as np
import torch

from . import _lib
from .envs import VecSurfaceCodeEnv, Surface_Code_Environment_Multi_Decoding_Cycles
from .episodes import EpisodeBook
from .qnet import QNetwork


# ---- configuration objects with the reference's names ----------------------------------------------
class Adam:
    """keras.optimizers.Adam(lr) (Keras 2 defaults)."""

    def __init__(self, lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.lr, self.beta_1, self.beta_2, self.epsilon = float(lr), float(beta_1), float(beta_2), float(epsilon)


class SequentialMemory:
    """rl.memory.SequentialMemory(limit, window_length=1): `limit` transitions.  Storage is a device ring of packed
    observations allocated by the agent once the environment (N, rows) is known: ceil(limit / N) + 1 slots."""

    def __init__(self, limit, window_length=1):
        if window_length != 1:
            raise ValueError("window_length must be 1 (the reference never uses another value)")
        self.limit, self.window_length = int(limit), 1
        self.ring = None

    @property
    def nb_entries(self):
        return 0 if self.ring is None else self.ring.filled * self.ring.n


class GreedyQPolicy:
    def __init__(self, masked_greedy=False):
        self.masked_greedy = bool(masked_greedy)
        self.eps = 0.0


class EpsGreedyQPolicy:
    def __init__(self, eps=0.1, masked_greedy=False):
        self.eps, self.masked_greedy = float(eps), bool(masked_greedy)


class BoltzmannQPolicy:        # imported by the reference scripts, never used
    def __init__(self, *a, **k):
        raise NotImplementedError("BoltzmannQPolicy is imported but never used by the reference")


class LinearAnnealedPolicy:
    """value(step) = max(value_min, value_max - (value_max - value_min) * step / nb_steps) while training,
    value_test while testing (exact against `mean_eps` of the shipped training histories)."""

    def __init__(self, inner_policy, attr, value_max, value_min, value_test, nb_steps):
        self.inner_policy, self.attr = inner_policy, attr
        self.value_max, self.value_min, self.value_test, self.nb_steps = float(value_max), float(value_min), float(value_test), float(nb_steps)

    def value(self, step, training=True):
        if not training:
            return self.value_test
        a = -(self.value_max - self.value_min) / self.nb_steps
        return max(self.value_min, a * float(step) + self.value_max)

    @property
    def masked_greedy(self):
        return self.inner_policy.masked_greedy


class FileLogger:
    """rl.callbacks.FileLogger(filepath, interval): JSON dict of per-episode lists, rewritten every `interval` episodes."""

    def __init__(self, filepath, interval=None):
        self.filepath, self.interval = filepath, interval
        self._last = 0

    def on_flush(self, history, force=False):
        n = len(history.get("episode", []))
        if force or self.interval is None or n - self._last >= self.interval:
            with open(self.filepath, "w") as f:
                json.dump(history, f, default=lambda o: o.item() if hasattr(o, "item") else str(o))      # numpy scalars
            self._last = n


class History:
    def __init__(self):
        self.history = {}

    def add(self, **kw):
        for k, v in kw.items():
            self.history.setdefault(k, []).append(v)

    def extend(self, **kw):
        """Many episodes at once: every value is a sequence of the same length (numpy arrays become Python scalars, as `add` stores)."""
        for k, v in kw.items():
            self.history.setdefault(k, []).extend(v.tolist() if hasattr(v, "tolist") else v)


class QNetSpec:
    """What `build_convolutional_nn` returns: the architecture, built into a `QNetwork` by the agent."""

    def __init__(self, cc_layers, ff_layers, input_shape, num_actions):
        self.cc_layers, self.ff_layers, self.input_shape, self.num_actions = cc_layers, ff_layers, tuple(input_shape), num_actions
        self._weights_path = None

    def load_weights(self, path):       # dqn.model.load_weights(path) before the agent is compiled
        self._weights_path = path


def build_convolutional_nn(cc_layers, ff_layers, input_shape, num_actions):
    """example_notebooks/Function_Library.py:338-377 (cc_layers [[filters, kernel, stride]], ff_layers [[units, dropout]])."""
    return QNetSpec(cc_layers, ff_layers, input_shape, num_actions)


# ---- replay ring -------------------------------------------------------------------------------------
class ReplayRing:
    """Device ring of packed observations [slot][row][npad] + per-slot action / reward / terminal rows."""

    def __init__(self, capacity, rows, npad, n, device):
        self.capacity, self.rows, self.npad, self.n = int(capacity), rows, npad, n
        self.obs = torch.zeros((self.capacity, rows, npad), dtype=torch.int64, device=device)
        self.act = torch.zeros((self.capacity, n), dtype=torch.int32, device=device)
        self.rew = torch.zeros((self.capacity, n), dtype=torch.float32, device=device)
        self.term = torch.zeros((self.capacity, n), dtype=torch.uint8, device=device)
        self.head, self.filled, self.pushed = 0, 0, 0

    def push_obs(self, rows_view):
        self.head = self.pushed % self.capacity
        self.obs[self.head].copy_(rows_view, non_blocking=True)
        self.filled = min(self.pushed, self.capacity - 1)
        self.pushed += 1

    def push_outcome(self, actions, reward, done):
        self.act[self.head].copy_(actions, non_blocking=True)
        self.rew[self.head].copy_(reward, non_blocking=True)
        self.term[self.head].copy_(done, non_blocking=True)

    def state_dict(self):
        return dict(obs=self.obs.cpu(), act=self.act.cpu(), rew=self.rew.cpu(), term=self.term.cpu(),
                    head=self.head, filled=self.filled, pushed=self.pushed)

    def load_state_dict(self, d):
        self.obs.copy_(d["obs"]); self.act.copy_(d["act"]); self.rew.copy_(d["rew"]); self.term.copy_(d["term"])
        self.head, self.filled, self.pushed = d["head"], d["filled"], d["pushed"]


def _vec(env):
    return env._vec if isinstance(env, Surface_Code_Environment_Multi_Decoding_Cycles) else env


class DQNAgent:
    def __init__(self, model, nb_actions, memory, nb_steps_warmup=1000, target_model_update=10000, policy=None,
                 test_policy=None, gamma=0.99, enable_dueling_network=False, enable_double_dqn=True, batch_size=32,
                 train_interval=1, memory_interval=1, delta_clip=np.inf, dueling_type="avg", updates_per_step=1,
                 seed=0, device="cuda:0", flush_interval=64, process_group=None, act_precision="fp32", target_precision="fp32",
                 collective="fused", train_precision="fp32"):
        if not enable_double_dqn:
            raise NotImplementedError("the reference always runs double DQN (keras-rl default)")
        if dueling_type != "avg" or delta_clip != np.inf or memory_interval != 1:
            raise NotImplementedError("only the reference's settings (dueling avg, delta_clip inf, memory_interval 1)")
        self.spec = model
        self.nb_actions, self.memory = int(nb_actions), memory
        self.nb_steps_warmup, self.target_model_update = int(nb_steps_warmup), int(target_model_update)
        self.policy = policy if policy is not None else EpsGreedyQPolicy()
        self.test_policy = test_policy if test_policy is not None else GreedyQPolicy()
        self.gamma, self.dueling = float(gamma), bool(enable_dueling_network)
        self.batch_size, self.train_interval, self.updates_per_step = int(batch_size), int(train_interval), int(updates_per_step)
        self.seed, self.device, self.flush_interval = int(seed), torch.device(device), int(flush_interval)
        self.process_group = process_group          # torch.distributed group for the gradient all-reduce (None = single GPU)
        if collective not in ("fused", "nccl"):
            raise ValueError("collective must be 'fused' (one peer-memory all-reduce+Adam kernel) or 'nccl' (all_reduce, then Adam)")
        self.collective = collective
        self.comm = None
        self.act_precision = act_precision          # "fp32" (SIMT) or "bf16" (tcgen05 tensor cores) for action selection
        self.target_precision = target_precision    # precision of the two no-grad forwards on s' inside an update (Q_online, Q_target)
        if train_precision not in ("fp32", "bf16"):
            raise ValueError("train_precision must be 'fp32' (Keras arithmetic) or 'bf16' (tcgen05 forward and backward, fp32 master weights)")
        self.train_precision = train_precision      # precision of the forward / backward pass of an update; Adam and the weights stay fp32
        self.optimizer = None
        self.model = None                           # QNetwork, built in compile()
        self.step, self.updates = 0, 0
        self.policy_step = 0        # iterations acted so far over ALL fit / test calls: the step index of the policy's Philox stream
        self.L = _lib.lib()

    # ---- reference API ----
    def compile(self, optimizer, metrics=None, max_envs=16384):
        self.optimizer = optimizer
        s = self.spec
        self.model = QNetwork(s.cc_layers, s.ff_layers, s.input_shape, s.num_actions, dueling=self.dueling,
                              max_batch=max(max_envs, self.batch_size), device=self.device, seed=self.seed)
        if getattr(s, "_weights_path", None):
            self.model.load_weights(s._weights_path)
        n = self.model.num_params
        self.target_params = self.model.params.clone()
        self.target_model = None
        if self.target_precision == "bf16":         # the target network needs its own handle for its staged bf16 weights
            self.target_model = QNetwork(s.cc_layers, s.ff_layers, s.input_shape, s.num_actions, dueling=self.dueling,
                                         max_batch=self.batch_size, device=self.device, seed=self.seed)
            self.target_model.params = self.target_params
        # a second handle on the ONLINE parameters for the no-grad forward of an update, so that it can run beside the training forward
        # (which uses the first handle's tensor-core buffers when train_precision is bf16)
        self.online_twin = None                     # built by update() when it is first needed
        self.grads = torch.zeros(n, dtype=torch.float32, device=self.model.device)
        self.adam_m, self.adam_v = torch.zeros_like(self.grads), torch.zeros_like(self.grads)
        if self.process_group is not None and self.collective == "fused":
            import torch.distributed as dist
            if dist.get_world_size(self.process_group) > 1:
                from .parallel import FusedAllreduceAdam
                self.comm = FusedAllreduceAdam(n, self.model.device, self.process_group)
        B, A, rows, dev = self.batch_size, self.nb_actions, self.model.packed_rows, self.model.device
        self._s0 = torch.zeros((rows, B), dtype=torch.int64, device=dev)
        self._s1 = torch.zeros((rows, B), dtype=torch.int64, device=dev)
        self._ba = torch.zeros(B, dtype=torch.int32, device=dev)
        self._br = torch.zeros(B, dtype=torch.float32, device=dev)
        self._bt = torch.zeros(B, dtype=torch.uint8, device=dev)
        self._qo = torch.zeros((B, A), dtype=torch.float32, device=dev)
        self._qt = torch.zeros((B, A), dtype=torch.float32, device=dev)
        self._q = torch.zeros((B, A), dtype=torch.float32, device=dev)
        self._dq = torch.zeros((B, A), dtype=torch.float32, device=dev)
        self._y = torch.zeros(B, dtype=torch.float32, device=dev)
        self._stats = torch.zeros(2, dtype=torch.float32, device=dev)
        self._actions = None
        self._side = None
        return self

    def load_weights(self, path):
        self.model.load_weights(path)
        self.target_params.copy_(self.model.params)
        if getattr(self, "target_model", None) is not None:
            self.target_model.params_changed()

    def save_weights(self, path, overwrite=True):
        self.model.save_weights(path, overwrite)

    def save_memory(self, path=None):
        """The reference pickles `dqn.memory` to memory.p (SPTS:156-157); here the replay ring's tensors (CPU copies).
        Returns the snapshot dict; with `path` also torch.save()s it."""
        snap = None if self.memory.ring is None else self.memory.ring.state_dict()
        if path is not None and snap is not None:
            torch.save(snap, path)
        return snap

    def load_memory(self, snap_or_path, env):
        """Restore a replay snapshot taken with the same number of lattices and ring capacity (the continue scripts load
        memory.p before fit, Single_Point_Continue_Training_Script.py:109-136)."""
        snap = torch.load(snap_or_path, weights_only=False) if isinstance(snap_or_path, str) else snap_or_path
        if snap is None:
            return
        v = _vec(env)
        cap, rows, npad = snap["obs"].shape
        self.memory.ring = ReplayRing(cap, rows, npad, v.n_envs, self.model.device)
        self.memory.ring.load_state_dict(snap)

    def _st(self):
        return C.c_void_p(torch.cuda.current_stream(self.model.device).cuda_stream)

    def _policy_params(self, policy, training):
        if isinstance(policy, LinearAnnealedPolicy):
            return policy.value(self.step, training), policy.masked_greedy
        return policy.eps, policy.masked_greedy          # GreedyQPolicy: 0; an EpsGreedyQPolicy handed in as test policy keeps its own eps

    def _act(self, v, rows_ptr, step_index, eps, masked):
        """Q(s) for all lattices from the packed rows inside the env state, then the eps-greedy pick."""
        N = v.n_envs
        q = self.model.forward_packed(rows_ptr, v.state_stride, N, precision=self.act_precision)
        if self._actions is None or self._actions.numel() != N:
            self._actions = torch.zeros(N, dtype=torch.int32, device=self.model.device)
        _lib.check(self.L.dq_policy_eps_greedy(C.c_void_p(q.data_ptr()), C.c_void_p(v.legal_mask.data_ptr()), N, v.mask_words,
                                               self.nb_actions, v.env_id_base, v.seed, step_index & 0xFFFFFFFF, None, float(eps),
                                               int(masked), C.c_void_p(self._actions.data_ptr()), self._st()))
        return self._actions

    def forward(self, observation):
        """Greedy (test-policy) action for ONE byte observation [C,H,W] -- notebook 3's production-decoding call.
        No legal-action set is available here, so the argmax runs over all actions, as in the notebook."""
        q = self.model.forward(np.asarray(observation)[None])
        return int(torch.argmax(q[0]).item())

    def train_on_ring(self, ring, draw_index):
        """One double-DQN update from the replay ring (sample -> targets -> forward/backward -> Adam)."""
        B, A, st, m = self.batch_size, self.nb_actions, self._st(), self.model
        p = lambda t: C.c_void_p(t.data_ptr())
        _lib.check(self.L.dq_replay_sample(p(ring.obs), p(ring.act), p(ring.rew), p(ring.term), ring.rows, ring.npad, ring.n,
                                           ring.capacity, ring.head, ring.filled, B, self.seed, draw_index & 0xFFFFFFFF,
                                           p(self._s0), p(self._s1), p(self._ba), p(self._br), p(self._bt), None, st))
        self.update(self._s0, self._s1, self._ba, self._br, self._bt)

    def update(self, s0, s1, actions, reward, terminal):
        """backward() of keras-rl for one batch of packed transitions (all device tensors)."""
        B, A, st, m = s0.shape[1], self.nb_actions, self._st(), self.model
        p = lambda t: C.c_void_p(t.data_ptr())
        tp = self.train_precision
        if self.target_model is not None:
            # Q_target(s'), Q_online(s') and the training forward on s are three independent chains of five to seven small kernels: they run
            # on three streams (the target network and -- when the training forward also uses the tensor-core buffers -- a twin of the
            # online network have their own handles, i.e. their own activation buffers and staged weights)
            main = torch.cuda.current_stream(m.device)
            if self._side is None:
                self._side = (torch.cuda.Stream(device=m.device), torch.cuda.Stream(device=m.device))
                self._ev = tuple(torch.cuda.Event() for _ in range(3))
            fork, join_t, join_o = self._ev
            fork.record(main)
            with torch.cuda.stream(self._side[0]):
                self._side[0].wait_event(fork)
                self.target_model.forward_packed(s1.data_ptr(), B, B, out=self._qt[:B], precision="bf16")
                join_t.record(self._side[0])
            if tp == "bf16" and self.online_twin is None:
                sp = self.spec
                self.online_twin = QNetwork(sp.cc_layers, sp.ff_layers, sp.input_shape, sp.num_actions, dueling=self.dueling,
                                            max_batch=self.batch_size, device=self.device, seed=self.seed)
                self.online_twin.share_params_with(m)
            online = self.online_twin if tp == "bf16" else m
            with torch.cuda.stream(self._side[1]):
                self._side[1].wait_event(fork)
                online.forward_packed(s1.data_ptr(), B, B, out=self._qo[:B], precision="bf16")
                join_o.record(self._side[1])
            self.updates += 1
            m.forward_packed(s0.data_ptr(), B, B, out=self._q[:B], train=True, dropout_seed=(self.seed << 20) ^ self.updates, precision=tp)
            main.wait_event(join_t); main.wait_event(join_o)
            _lib.check(self.L.dq_dqn_targets(p(self._qo), p(self._qt), p(reward), p(terminal), self.gamma, B, A, p(self._y), st))
        else:
            m.forward_packed(s1.data_ptr(), B, B, out=self._qo[:B])
            m.forward_packed(s1.data_ptr(), B, B, out=self._qt[:B], params=self.target_params)
            _lib.check(self.L.dq_dqn_targets(p(self._qo), p(self._qt), p(reward), p(terminal), self.gamma, B, A, p(self._y), st))
            self.updates += 1
            m.forward_packed(s0.data_ptr(), B, B, out=self._q[:B], train=True, dropout_seed=(self.seed << 20) ^ self.updates, precision=tp)
        _lib.check(self.L.dq_dqn_loss_grad(p(self._q), p(actions), p(self._y), B, A, p(self._dq), p(self._stats), st))
        o = self.optimizer
        if self.comm is not None:          # gradient mean over the ranks fused with the Adam step (csrc/dq_comm.cu)
            m.backward_packed(s0.data_ptr(), B, B, self._dq, self.comm.grads(), precision=tp)
            self.comm.step(m.params, self.adam_m, self.adam_v, o, self.updates, st)
            m.params_changed()
            return
        m.backward_packed(s0.data_ptr(), B, B, self._dq, self.grads, precision=tp)
        scale = 1.0
        if self.process_group is not None:
            import torch.distributed as dist
            dist.all_reduce(self.grads, group=self.process_group)
            scale = 1.0 / dist.get_world_size(self.process_group)
        _lib.check(self.L.dq_adam_step(p(m.params), p(self.adam_m), p(self.adam_v), p(self.grads), m.num_params, o.lr, o.beta_1,
                                       o.beta_2, o.epsilon, self.updates, scale, st))
        m.params_changed()

    # ---- fit --------------------------------------------------------------------------------------
    def fit(self, env, nb_steps, action_repetition=1, callbacks=None, verbose=1, visualize=False, nb_max_start_steps=0,
            start_step_policy=None, log_interval=10000, nb_max_episode_steps=None, episode_averaging_length=1000,
            success_threshold=1e5, stopping_patience=1e9, min_nb_steps=0, single_cycle=False):
        if action_repetition != 1 or nb_max_start_steps != 0 or single_cycle:
            raise NotImplementedError("only the settings the reference uses (action_repetition=1, no start steps, multi-cycle)")
        v = _vec(env)
        N, dev = v.n_envs, self.model.device
        if N > self.model.max_batch:
            raise ValueError("compile(max_envs=...) must cover the environment's %d lattices" % N)
        # Data-parallel fit: warm-up, the target period, train_interval and termination must fall on the same iteration on every
        # rank (every update is a collective), so all ranks count the LARGEST shard's transitions per iteration.
        step_inc = N
        if self.process_group is not None:
            import torch.distributed as dist
            if dist.get_world_size(self.process_group) > 1:
                t = torch.tensor([N], dtype=torch.int64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.process_group)
                step_inc = int(t.item())
        rows_ptr, nrows, stride = C.c_void_p(), C.c_int64(), C.c_int64()
        _lib.check(self.L.dq_env_packed_obs(v._h, C.byref(rows_ptr), C.byref(nrows), C.byref(stride)))
        from .qnet import device_view
        rows_view = device_view(rows_ptr.value, (nrows.value, stride.value), "<i8", dev)
        if self.memory.ring is None:
            self.memory.ring = ReplayRing(max(2, -(-self.memory.limit // N) + 1), nrows.value, stride.value, N, dev)
        ring = self.memory.ring
        if ring.pushed > 0:
            # resumed ring (an earlier fit, or load_memory): its newest slot holds (s, a, r) of a transition whose s' was never
            # stored.  The first observation of this fit takes that slot over, so the sampler can never pair it with this episode.
            ring.pushed -= 1
        self.step = 0               # keras-rl's Agent.fit restarts the step counter (warm-up and the eps schedule with it)
        K = self.flush_interval
        h_rew = torch.zeros((K, N), dtype=torch.float32, device=dev)
        h_done = torch.zeros((K, N), dtype=torch.uint8, device=dev)
        h_life = torch.zeros((K, N), dtype=torch.int32, device=dev)
        hist = History()
        ep_reward, ep_steps = np.zeros(N), np.zeros(N, np.int64)
        book = EpisodeBook(episode_averaging_length, success_threshold, stopping_patience, min_nb_steps)
        eps_sum, eps_n = 0.0, 0
        t_start = t_last = time.time()
        it, stop = 0, False
        p = lambda t: C.c_void_p(t.data_ptr())
        v.reset()
        self._stats.zero_()
        upd_window = 0
        while self.step < nb_steps and not stop:
            eps, masked = self._policy_params(self.policy, True)
            ring.push_obs(rows_view)
            actions = self._act(v, rows_ptr.value, self.policy_step, eps, masked)
            self.policy_step += 1
            _lib.check(self.L.dq_env_step(v._h, p(actions), None, p(v.reward), p(v.done), p(v.lifetime), p(v.legal_mask), 1, self._st()))
            ring.push_outcome(actions, v.reward, v.done)
            k = it % K
            h_rew[k].copy_(v.reward, non_blocking=True); h_done[k].copy_(v.done, non_blocking=True); h_life[k].copy_(v.lifetime, non_blocking=True)
            self.step += step_inc
            it += 1
            if self.step > self.nb_steps_warmup:
                eps_sum += eps; eps_n += 1
                if it % self.train_interval == 0 and ring.filled >= 1:
                    for _ in range(self.updates_per_step):
                        self.train_on_ring(ring, self.updates)
                        upd_window += 1
            if (self.step // self.target_model_update) != ((self.step - step_inc) // self.target_model_update):
                self.target_params.copy_(self.model.params)
                if self.target_model is not None:
                    self.target_model.params_changed()
            if k == K - 1 or self.step >= nb_steps:
                # ---- drain: episode bookkeeping on the host, in (iteration, lattice) order
                rew, done, life = h_rew[:k + 1].cpu().numpy(), h_done[:k + 1].cpu().numpy(), h_life[:k + 1].cpu().numpy()
                stats = self._stats.cpu().numpy().copy(); self._stats.zero_()
                loss = stats[0] / (upd_window * self.batch_size) if upd_window else float("nan")
                mean_q = stats[1] / (upd_window * self.batch_size) if upd_window else float("nan")
                mean_eps = eps_sum / eps_n if eps_n else float("nan")
                upd_window, eps_sum, eps_n = 0, 0.0, 0
                now = time.time()
                per_episode_s = (now - t_last) / max(1, int(done.sum()))
                ep_life, ep_nb, ep_rew, ep_len = [], [], [], []
                for j in range(k + 1):
                    ep_reward += rew[j]; ep_steps += 1
                    idx = np.nonzero(done[j])[0]
                    if idx.size:
                        ep_life.append(life[j, idx]); ep_nb.append(np.full(idx.size, int(self.step - (k - j) * step_inc), np.int64))
                        ep_rew.append(ep_reward[idx]); ep_len.append(ep_steps[idx])
                        ep_reward[idx] = 0.0; ep_steps[idx] = 0
                if ep_life:                                            # rolling / best / stop rules for the whole drain at once: episodes.py
                    nbs = np.concatenate(ep_nb)
                    entry = book.finish_many(np.concatenate(ep_life), nbs)
                    m = nbs.size
                    hist.extend(loss=[loss] * m, mean_q=[mean_q] * m, mean_eps=[mean_eps] * m, episode_reward=np.concatenate(ep_rew),
                                nb_episode_steps=np.concatenate(ep_len), nb_steps=nbs, duration=[per_episode_s] * m, **entry)
                stop = stop or book.stop
                if self.comm is not None:
                    self.comm.check()                   # a peer that never arrived: abort here, not after the run
                if self.process_group is not None:      # ranks must leave the loop together (every update is a collective)
                    import torch.distributed as dist
                    flag = torch.tensor([1 if stop else 0], dtype=torch.int32, device=dev)
                    dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.process_group)
                    stop = bool(flag.item())
                t_last = now
                for cb in (callbacks or []):
                    if hasattr(cb, "on_flush"):
                        cb.on_flush(hist.history)
                if verbose and (it // K) % max(1, int(log_interval // max(1, K * N))) == 0 and book.win_n:
                    print("step %d  episodes %d  rolling lifetime %.1f  best %.1f  eps %.3f  loss %.4g  mean_q %.3f  %.0f env-steps/s" % (
                        self.step, book.episode, book.rolling, book.best_avg, eps, loss, mean_q,
                        self.step / max(1e-9, now - t_start)), flush=True)
        if self.comm is not None:
            self.comm.check()
        for cb in (callbacks or []):
            if hasattr(cb, "on_flush"):
                cb.on_flush(hist.history, force=True)
        return hist

    # ---- test -------------------------------------------------------------------------------------
    def test(self, env, nb_episodes=1, action_repetition=1, callbacks=None, visualize=False, nb_max_episode_steps=None,
             nb_max_start_steps=0, start_step_policy=None, verbose=1, interval=100, single_cycle=False, max_iterations=None):
        """Greedy evaluation with the test policy.  Every lattice plays ceil(nb_episodes / N) episodes (a fixed quota
        per lattice keeps the lifetime estimate unbiased; taking the first episodes to finish would favour short
        ones); the first nb_episodes records in lattice-major order are returned.  History keys as in the fork:
        'episode_lifetime', 'episode_lifetimes_rolling_avg' (cumulative mean), 'episode_reward', 'nb_steps'."""
        v = _vec(env)
        N, dev = v.n_envs, self.model.device
        quota = -(-int(nb_episodes) // N)
        rows_ptr, nrows, stride = C.c_void_p(), C.c_int64(), C.c_int64()
        _lib.check(self.L.dq_env_packed_obs(v._h, C.byref(rows_ptr), C.byref(nrows), C.byref(stride)))
        eps, masked = self._policy_params(self.test_policy, False)
        p = lambda t: C.c_void_p(t.data_ptr())
        v.reset()
        count = torch.zeros(N, dtype=torch.int64, device=dev)
        ep_rew = torch.zeros(N, dtype=torch.float32, device=dev)
        ep_len = torch.zeros(N, dtype=torch.int64, device=dev)
        rec_life = torch.zeros((quota, N), dtype=torch.int32, device=dev)
        rec_rew = torch.zeros((quota, N), dtype=torch.float32, device=dev)
        rec_len = torch.zeros((quota, N), dtype=torch.int64, device=dev)
        lane = torch.arange(N, device=dev)
        it = 0
        while True:
            actions = self._act(v, rows_ptr.value, self.policy_step, eps, masked)
            self.policy_step += 1
            _lib.check(self.L.dq_env_step(v._h, p(actions), None, p(v.reward), p(v.done), p(v.lifetime), p(v.legal_mask), 1, self._st()))
            it += 1
            done = v.done.bool()
            ep_rew += v.reward; ep_len += 1
            take = done & (count < quota)
            slot = torch.clamp(count, max=quota - 1)
            rec_life[slot, lane] = torch.where(take, v.lifetime, rec_life[slot, lane])
            rec_rew[slot, lane] = torch.where(take, ep_rew, rec_rew[slot, lane])
            rec_len[slot, lane] = torch.where(take, ep_len, rec_len[slot, lane])
            count += take.long()
            ep_rew = torch.where(done, torch.zeros_like(ep_rew), ep_rew)
            ep_len = torch.where(done, torch.zeros_like(ep_len), ep_len)
            if it % 32 == 0 and bool((count >= quota).all()):
                break
            if max_iterations is not None and it >= max_iterations:
                break
        finished = (torch.arange(quota, device=dev)[:, None] < count[None, :]).T.reshape(-1).cpu().numpy()
        life = rec_life.T.reshape(-1).cpu().numpy()[finished][:nb_episodes]        # lattice-major
        rew = rec_rew.T.reshape(-1).cpu().numpy()[finished][:nb_episodes]
        length = rec_len.T.reshape(-1).cpu().numpy()[finished][:nb_episodes]
        hist = History()
        cum = np.cumsum(life) / np.arange(1, len(life) + 1)
        hist.history = {"episode_lifetime": [int(x) for x in life], "episode_lifetimes_rolling_avg": [float(x) for x in cum],
                        "episode_reward": [float(x) for x in rew], "nb_steps": [int(x) for x in length], "iterations": it}
        if verbose and len(life):
            print("tested %d episodes on %d lattices: mean lifetime %.2f (+- %.2f)" % (
                len(life), N, float(life.mean()), float(life.std() / math.sqrt(len(life)))), flush=True)
        return hist
