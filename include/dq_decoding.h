/*
 * dq_decoding.h -- C ABI of the B200-native DeepQ-Decoding hot path.
 *
 * Drop-in boundary for the reference's per-step environment / DQN inner loop.  The
 * reference has no FFI (it is one Python process); each entry point below names the
 * reference interface it replaces (paths relative to the reference repo root,
 * EN = example_notebooks/, SPTS = cluster_scripts/d5_dp/0.001/Single_Point_Training_Script.py).
 * INTEGRATION.md shows the ctypes stub a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative DQ_E* code on failure;
 *     dq_last_error() returns a thread-local message for the last failure;
 *   - the caller owns every device/host buffer it passes (the library never frees
 *     them); the library owns only the opaque handles and their internal state;
 *   - device entry points are asynchronous on the passed stream (a cudaStream_t
 *     cast to void*; NULL = the legacy default stream) and never synchronise;
 *     *_host entry points take HOST pointers, copy in/out on an internal stream and
 *     return after the results have landed;
 *   - a handle is not thread-safe; distinct handles are independent;
 *   - there is no CPU fallback: without a CUDA device every call fails with DQ_ECUDA.
 *
 * Shapes (N = n_envs, A = num_actions, C = volume_depth + action layers, H = 2d+1,
 * W = ceil(A/64)):
 *   obs        uint8  [N][C][H][H]   the reference's board_state (EN/Environments.py:91), 0/1 cells (any alignment;
 *                                    16-byte aligned buffers get 128-bit stores)
 *   legal_mask uint64 [N][W]         bit a of word a/64 set  <=>  a in env.legal_actions
 *   actions    int32  [N]            action index as passed to env.step(); values outside [0,A) act as identity
 *   reward     float  [N]            1.0f / 0.0f   (EN/Environments.py:146-149)
 *   done       uint8  [N]            env.done after the step
 *   lifetime   int32  [N]            env.lifetime after the step (the finished episode's value when done)
 */
#ifndef DQ_DECODING_H
#define DQ_DECODING_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DQ_OK        0
#define DQ_EINVAL   -1   /* bad argument */
#define DQ_ECUDA    -2   /* CUDA runtime failure / no device */
#define DQ_ESTATE   -3   /* call not valid in this state (e.g. step before referee set) */

#define DQ_MODEL_X   0   /* error_model == "X"  */
#define DQ_MODEL_DP  1   /* error_model == "DP" */

#define DQ_REFEREE_JOINT 0  /* one table over all d*d-1 stabilizers, 2-bit class X+2Z            */
#define DQ_REFEREE_SPLIT 1  /* table A: X class over type-3 stabilizers; table B: Z class, type-1 */

/* dq_env_info selectors */
#define DQ_INFO_NUM_ACTIONS   0
#define DQ_INFO_OBS_CHANNELS  1
#define DQ_INFO_OBS_SIDE      2
#define DQ_INFO_MASK_WORDS    3
#define DQ_INFO_STATE_WORDS   4   /* rows of the packed state matrix  */
#define DQ_INFO_STATE_STRIDE  5   /* columns (n_envs rounded up to 32) */
#define DQ_INFO_NUM_STABS     6
#define DQ_INFO_N_TYPE3       7
#define DQ_INFO_N_TYPE1       8
#define DQ_INFO_RNG_BLOCKS    9   /* Philox blocks per volume attempt (B of the RNG contract) */
#define DQ_INFO_HOST_EXPAND  10   /* 1 if the *_host calls move observations bit-packed and expand them on the host (default; DQ_HOST_EXPAND=0 turns it off) */

typedef struct dq_env dq_env;
typedef void* dq_stream;

const char* dq_last_error(void);
int dq_version(void);

/* Replaces Surface_Code_Environment_Multi_Decoding_Cycles.__init__ (EN/Environments.py:45-97)
 * for n_envs independent lattices living on `device`.  d in {3,5,7}; volume_depth in [1,8].
 * seed / env_id_base select the Philox streams: env i uses stream id
 * env_id_base + i, so a sharded run reproduces the unsharded one. */
int dq_env_create(dq_env** out, int d, int error_model, int use_Y, int volume_depth,
                  double p_phys, double p_meas, int64_t n_envs, uint64_t seed,
                  int64_t env_id_base, int device);
int dq_env_destroy(dq_env* env);
int dq_env_info(const dq_env* env, int what, int64_t* out);

/* env.p_phys / env.p_meas are assignable in the reference (SPTS:200-201). */
int dq_env_set_noise(dq_env* env, double p_phys, double p_meas);
/* Documented deviation: the reference redraws an all-trivial syndrome volume forever (EN/Environments.py:158-170 never
 * returns at p_phys = p_meas = 0 on a clean frame); a kernel must end, so after max_attempts all-trivial attempts on one
 * volume the trivial volume is accepted.  Default 2^20; range [1, 2^20].  Tests lower it to exercise the branch. */
int dq_env_set_max_attempts(dq_env* env, int max_attempts);

/* Replaces the duck-typed static_decoder.predict + argmax (EN/Environments.py:144,150) by
 * table lookups.  Tables are DEVICE pointers, 2 bits per entry, 4 entries per byte, entry i
 * in bits 2*(i&3) of byte i>>2.  The library keeps the pointers (caller keeps them alive).
 * JOINT: lut_a has 2^(d*d-1) entries; index bit k = true-syndrome bit of the k-th stabilizer in
 * "joint order": plaquette-grid rows 1..d-1 left to right (d stabilizers each), then the top/bottom
 * boundary stabilizers by column 1..d-1 (odd columns lie on row 0, even ones on row d).
 * SPLIT: lut_a has 2^(#type-3) entries (logical-X class bit), lut_b 2^(#type-1) entries
 * (logical-Z class bit; ignored for DQ_MODEL_X), each indexed in the stabilizer draw order of
 * EN/Function_Library.py:186-233 (bulk row-major, top, bottom, left, right) restricted to its type. */
int dq_env_set_referee_lut(dq_env* env, int mode, const void* dev_lut_a, int64_t bytes_a,
                           const void* dev_lut_b, int64_t bytes_b);

/* env.reset() for every lattice (EN/Environments.py:99-115).  obs / legal_mask may be NULL. */
int dq_env_reset(dq_env* env, uint8_t* obs, uint64_t* legal_mask, dq_stream stream);

/* env.step(action) for every lattice (EN/Environments.py:118-204).  With auto_reset != 0 a
 * lattice that finishes (done=1) is reset inside the same call: done/lifetime/reward describe
 * the finished step, obs/legal_mask the first state of the new episode.  With auto_reset == 0
 * the reference's behaviour is kept (done stays set until dq_env_reset).  Any output may be NULL. */
int dq_env_step(dq_env* env, const int32_t* actions, uint8_t* obs, float* reward, uint8_t* done,
                int32_t* lifetime, uint64_t* legal_mask, int auto_reset, dq_stream stream);

/* dq_env_step with the uniform random-legal policy built in: every lattice draws its own action exactly as
 * dq_policy_random_legal_next would (same Philox word, step index = the handle's device-side counter, advanced by
 * the launch), then steps.  One kernel per (policy, step) pair; actions_out (optional) receives the picks. */
int dq_env_step_random(dq_env* env, uint8_t* obs, float* reward, uint8_t* done, int32_t* lifetime,
                       uint64_t* legal_mask, int32_t* actions_out, int auto_reset, dq_stream stream);
/* n_steps successive dq_env_step_random calls in ONE launch (bit-identical to making them one by one): lattices are
 * independent, so each tile of lattices runs its n_steps without waiting for the slowest tile of every step.  This is
 * the loop `for _ in range(n_steps): env.step(random legal action)` over Environments.py:118-204.
 *   obs_ring   [ring_slots][N][C][H][H] or NULL: step s writes slot (first_slot + s) % ring_slots
 *   reward, done, lifetime, actions_out   [n_steps][N] or NULL;  legal [n_steps][N][W] or NULL */
int dq_env_rollout_random(dq_env* env, int n_steps, uint8_t* obs_ring, int ring_slots, int first_slot, float* reward,
                          uint8_t* done, int32_t* lifetime, uint64_t* legal, int32_t* actions_out, int auto_reset,
                          dq_stream stream);

/* Same two calls with HOST buffers (the path bench.py's e2e number times).  The observation crosses PCIe bit-packed and is expanded into
 * h_obs by the library's host threads.  The small arrays do not travel by DMA: the kernel reads h_actions from, and stores reward / done /
 * lifetime / legal masks into, pinned host memory -- the caller's own buffers when they are pinned (cudaHostAlloc / cudaHostRegister, looked
 * up once per buffer address and handle: a buffer passed here must keep its pinned-or-pageable nature while the handle lives), otherwise a
 * pinned block of the handle that is memcpy'd to / from the pageable buffers.  h_actions must not be modified between a _begin and its _end.
 * DQ_HOST_ZEROCOPY=0 / DQ_HOST_DIRECT=0 / DQ_HOST_EXPAND=0 in the environment select the DMA-copy / staged / byte-copy forms (same results). */
int dq_env_reset_host(dq_env* env, uint8_t* h_obs, uint64_t* h_legal_mask);
int dq_env_step_host(dq_env* env, const int32_t* h_actions, uint8_t* h_obs, float* h_reward,
                     uint8_t* h_done, int32_t* h_lifetime, uint64_t* h_legal_mask, int auto_reset);
/* dq_env_step_host in two halves.  _begin queues the copy-in of the actions, the launch and every copy-out on the handle's own
 * stream and returns at once (the buffers must stay valid, and pinned for the copies to be asynchronous); _end waits for the results
 * and finishes the observations.  One call may be in flight per handle; two handles driven begin(A) begin(B) end(A) begin(A) end(B) ...
 * overlap one handle's kernel and PCIe traffic with the other's host-side work.
 * What crosses PCIe for h_obs are the bit-packed bitmap rows (7.5x fewer bytes at d = 5); host threads of the library expand them
 * into the caller's byte buffer in _end, range by range while the later ranges are still in flight (DQ_HOST_THREADS = threads per
 * process, default: the CPUs the process may run on, less one; the Python binding divides them by LOCAL_WORLD_SIZE; DQ_HOST_EXPAND=0 makes the kernel write bytes and copies all of them back instead). */
int dq_env_step_host_begin(dq_env* env, const int32_t* h_actions, uint8_t* h_obs, float* h_reward, uint8_t* h_done,
                           int32_t* h_lifetime, uint64_t* h_legal_mask, int auto_reset);
int dq_env_step_host_end(dq_env* env);
/* The uniform pick over the sorted legal actions (dq_policy_random_legal, same Philox word) computed on the host from host-resident
 * legal masks: the policy of the env-only benchmark for callers that live on the host side of the *_host calls. */
int dq_policy_random_legal_host(const dq_env* env, const uint64_t* h_legal_mask, uint32_t step_index, int32_t* h_actions);
/* The same two calls returning the observations PACKED -- uint64 h_packed[C*PW][STATE_STRIDE], bit x*H+y of the PW-word
 * bitmap of layer l of lattice i in h_packed[l*PW + word][i], i.e. the rows dq_env_packed_obs exposes and the Q-network
 * consumes -- instead of one byte per cell: 7.5x fewer bytes over PCIe at d = 5.  (The reference returns board_state as an
 * int64 array, EN/Environments.py:204; deepq_decoding_b200.envs.unpack_observations restores that form on the host.) */
int dq_env_reset_host_packed(dq_env* env, uint64_t* h_packed, uint64_t* h_legal_mask);
int dq_env_step_host_packed(dq_env* env, const int32_t* h_actions, uint64_t* h_packed, float* h_reward, uint8_t* h_done,
                            int32_t* h_lifetime, uint64_t* h_legal_mask, int auto_reset);
/* The expansion on its own, with the library's host threads: packed rows as returned above (stride = STATE_STRIDE) -> uint8
 * h_obs[n][channels][2d+1][2d+1] of 0/1, the board_state of EN/Environments.py:204 for every lattice (padding_syndrome /
 * padding_actions layout, :273-314). */
int dq_unpack_observations_host(const uint64_t* h_packed, int64_t stride, int64_t n, int d, int channels, uint8_t* h_obs);

/* Packed per-lattice state, uint64 [STATE_WORDS][STATE_STRIDE] on the device (layout in
 * DESIGN.md section 2): what env.hidden_state / completed_actions / lifetime / done hold in the
 * reference.  Parity tests inspect and inject it; also the checkpoint format. */
int dq_env_get_state(dq_env* env, uint64_t* dev_words, dq_stream stream);
int dq_env_set_state(dq_env* env, const uint64_t* dev_words, dq_stream stream);

/* Uniform pick over the sorted legal actions of every lattice (the random-legal policy the
 * env-only benchmark uses; also the epsilon branch of EpsGreedyQPolicy, SPTS:110-114).
 * Draw = Philox(stream id, step_index, block 0, domain 1) word 0; index = floor(u*n/2^32). */
int dq_policy_random_legal(const dq_env* env, const uint64_t* legal_mask, uint32_t step_index,
                           int32_t* actions, dq_stream stream);

/* Same pick, with the step index kept in device memory by the handle: dq_policy_seek sets it, every
 * dq_policy_random_legal_next launch uses it and then advances it by one ON THE DEVICE, so a captured
 * CUDA graph of (policy, step) pairs draws fresh words on every replay. */
int dq_policy_seek(dq_env* env, uint32_t step_index, dq_stream stream);
int dq_policy_random_legal_next(dq_env* env, const uint64_t* legal_mask, int32_t* actions, dq_stream stream);

/* ---------------------------------------------------------------------------------------------------
 * Q-network + DQN inner loop.  Replaces build_convolutional_nn + the keras-rl dueling head
 * (EN/Function_Library.py:338-377), DQNAgent.forward/backward, the policies and SequentialMemory
 * (keras-rl fork, call sites SPTS:109-152, 166-167).  fp32 throughout.
 *
 * A network input is a PACKED observation: one (2d+1)^2-cell bitmap per observation layer, PW =
 * ceil((2d+1)^2/64) uint64 words each, stored as packed[layer*PW + word][sample] with row stride `stride`
 * (>= batch).  Rows ROW_BM.. of the environment state matrix (dq_env_packed_obs) have exactly this
 * form, so acting needs no byte observation at all; dq_qnet_pack_obs converts the reference's uint8
 * board_state for callers that have only that.
 * Parameters, gradients and Adam moments are caller-owned flat fp32 device buffers of
 * DQ_QINFO_NUM_PARAMS elements laid out as dq_qnet_param_layout describes. */
typedef struct dq_qnet dq_qnet;
#define DQ_QINFO_NUM_PARAMS       0
#define DQ_QINFO_PACKED_ROWS      1
#define DQ_QINFO_NUM_TENSORS      2
#define DQ_QINFO_FLOPS_PER_SAMPLE 3   /* forward, 2*MAC */

/* cc_layers / ff_layers of build_convolutional_nn: conv l has filters[l] kernels[l]xkernels[l] stride strides[l]
 * ('valid', ReLU); hidden dense i has units[i] (+ReLU, Dropout(dropout[i]) in training mode); then
 * Dense(num_actions) linear and, if dueling, keras-rl's Dense(num_actions+1) + 'avg' combination. */
int dq_qnet_create(dq_qnet** out, int in_channels, int in_side, int n_conv, const int* filters, const int* kernels,
                   const int* strides, int n_dense, const int* units, const float* dropout, int num_actions,
                   int dueling, int64_t max_batch, int device);
int dq_qnet_destroy(dq_qnet* net);
int dq_qnet_info(const dq_qnet* net, int what, int64_t* out);
/* For tensor pair t (conv layers first, then dense): offsets[2t] / offsets[2t+1] = kernel / bias offset in the
 * flat buffer, shapes[2t] x shapes[2t+1] = kernel rows (inputs) x columns (outputs).  Conv kernels are Keras
 * HWIO flattened; rows of the first dense kernel are in (position, channel) order. */
int dq_qnet_param_layout(const dq_qnet* net, int64_t* offsets, int64_t* shapes);
int dq_qnet_pack_obs(dq_qnet* net, const uint8_t* obs, uint64_t* packed, int64_t stride, int64_t batch, dq_stream stream);
/* model.predict_on_batch: Q[batch][num_actions].  train != 0 applies dropout (mask from dropout_seed) and keeps
 * the activations for dq_qnet_backward. */
int dq_qnet_forward(dq_qnet* net, const float* params, const uint64_t* packed, int64_t stride, int64_t batch,
                    float* q_out, int train, uint64_t dropout_seed, dq_stream stream);
/* Same Q values through the bf16 tensor-core path (tcgen05.mma, fp32 accumulation in TMEM) for every layer after the
 * first; inference / acting only.  Returns DQ_EINVAL for shapes the path does not cover (channel / unit counts must be
 * multiples of 8, >= 1 hidden dense layer); there is no silent fallback.  The first layer expands the packed bits into
 * a bf16 {0,1} tile in shared memory, every later layer gathers its A rows from the previous bf16 activation. */
int dq_qnet_forward_tc(dq_qnet* net, const float* params, const uint64_t* packed, int64_t stride, int64_t batch,
                       float* q_out, dq_stream stream);
/* Stages the bf16 weight copies dq_qnet_forward_tc multiplies with; call after every change of `params`. */
int dq_qnet_prepare_tc(dq_qnet* net, const float* params, dq_stream stream);
int dq_qnet_tc_activation(dq_qnet* net, int index, void** dev_ptr, int64_t* per_sample);   /* tests: bf16 activations */
/* bf16 tensor-core TRAINING step (model.train_on_batch of keras-rl's DQNAgent.backward, SPTS:119-152, in mixed precision: fp32
 * master weights, bf16 operands, fp32 accumulation in TMEM, fp32 gradients).  dq_qnet_forward_tc_train is dq_qnet_forward_tc with
 * dropout applied (mask from dropout_seed, the same Philox words as dq_qnet_forward) and Dense(num_actions) / the dueling layer kept
 * apart; dq_qnet_backward_tc then overwrites `grads` with the gradient of sum_b sum_a dq[b][a]*Q[b][a] for that batch: per layer
 * dW = A^T x dY (contraction over batch x positions, split over CTAs, fp32 atomics) and dCol = dY x W^T, both tcgen05 GEMMs.
 * Same shape coverage and error behaviour as dq_qnet_forward_tc; call dq_qnet_prepare_tc after every change of `params`. */
int dq_qnet_forward_tc_train(dq_qnet* net, const float* params, const uint64_t* packed, int64_t stride, int64_t batch,
                             float* q_out, uint64_t dropout_seed, dq_stream stream);
int dq_qnet_backward_tc(dq_qnet* net, const float* params, const uint64_t* packed, int64_t stride, int64_t batch,
                        const float* dq, float* grads, dq_stream stream);
/* Inference only: Dense(num_actions), keras-rl's dueling Dense(num_actions + 1) and its 'avg' combine (SPTS:119-130,
 * enable_dueling_network=True) are all linear, so Q = h * w_out + b_out with w_out [K][num_actions] (K = units of the last
 * hidden dense layer) and b_out [num_actions], both device fp32.  dq_qnet_prepare_tc stages this map and dq_qnet_forward_tc ends in it
 * (one launch instead of three; DQ_QNET_FOLD_HEAD=0 in the environment keeps the layers apart). */
int dq_qnet_fold_head(const dq_qnet* net, const float* params, float* w_out, float* b_out, dq_stream stream);
/* Gradient of sum_b sum_a dq[b][a]*Q[b][a] for the batch of the last dq_qnet_forward call; grads is overwritten. */
int dq_qnet_backward(dq_qnet* net, const float* params, const uint64_t* packed, int64_t stride, int64_t batch,
                     const float* dq, float* grads, dq_stream stream);
int dq_qnet_activation(dq_qnet* net, int index, float** dev_ptr, int64_t* per_sample);   /* tests */
/* keras.optimizers.Adam (Keras 2): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps); g *= grad_scale first. */
int dq_adam_step(float* params, float* m, float* v, const float* grads, int64_t n, float lr, float beta1, float beta2,
                 float eps, int64_t t, float grad_scale, dq_stream stream);
/* double DQN: y = r + gamma*(1-terminal)*Q_target(s', argmax_a Q_online(s', a)) */
int dq_dqn_targets(const float* q_online_next, const float* q_target_next, const float* reward, const uint8_t* terminal,
                   float gamma, int64_t batch, int num_actions, float* y, dq_stream stream);
/* loss = mean_b 0.5*(y_b - Q[b][a_b])^2 (delta_clip = inf); dq = dloss/dQ; stats (optional, float[2]) += {sum of the
 * per-sample losses, sum of max_a Q}. */
int dq_dqn_loss_grad(const float* q, const int32_t* actions, const float* y, int64_t batch, int num_actions, float* dq,
                     float* stats, dq_stream stream);
/* EpsGreedyQPolicy / GreedyQPolicy with env.legal_actions (SPTS:110-115, 166-167): with probability eps uniform over
 * the legal actions, else argmax over all actions (masked_greedy = 0) or over the legal ones (1).  Draws as in
 * dq_policy_random_legal plus word 1 for the eps test.  dev_step_counter (optional, uint32[2] on the device) replaces
 * step_index and is advanced by the launch, for CUDA-graph capture. */
int dq_policy_eps_greedy(const float* q, const uint64_t* legal_mask, int64_t n, int mask_words, int num_actions,
                         uint32_t env_id_base, uint64_t seed, uint32_t step_index, uint32_t* dev_step_counter, double eps,
                         int masked_greedy, int32_t* actions, dq_stream stream);
/* SequentialMemory.sample on a ring of packed observations ring_obs[slot][row][npad] with per-slot action / reward /
 * terminal arrays [slot][n]: draws `batch` (slot, lattice) pairs uniformly over the `filled` most recent complete
 * transitions (slot `head` holds the newest observation, whose transition is not complete yet) and gathers
 * s, s' (= next slot), a, r, terminal.  Draw = Philox(sample index, draw_index, domain 2). */
int dq_replay_sample(const uint64_t* ring_obs, const int32_t* ring_act, const float* ring_rew, const uint8_t* ring_term,
                     int rows, int64_t npad, int64_t n, int capacity, int head, int filled, int64_t batch, uint64_t seed,
                     uint32_t draw_index, uint64_t* s0, uint64_t* s1, int32_t* act, float* rew, uint8_t* term,
                     int32_t* picked, dq_stream stream);
/* Device pointer to the packed observation rows inside the env state (uint64 [C*PW][STATE_STRIDE]). */
int dq_env_packed_obs(dq_env* env, uint64_t** dev_rows, int64_t* n_rows, int64_t* stride);

/* ---- data-parallel training exchange (one process per GPU, NVLink peer memory) -------------------------------------
 * The reference has no exchange step (one agent per process; cluster_scripts/Controller.py runs one Slurm job per
 * hyper-parameter point).  Sharding the lattices of ONE run over the GPUs of a box adds exactly one: the mean of the
 * flat gradient before the Adam step of DQNAgent.backward (Single_Point_Training_Script.py:119-130).  dq_comm does
 * that mean and the Adam update in one kernel per rank, reading every rank's gradient through CUDA-IPC peer mappings.
 *   create -> handle (64 bytes, exchange them out of band, e.g. torch.distributed.all_gather) -> connect;
 *   per update: next_grads (where dq_qnet_backward must write this update's gradient) -> allreduce_adam.
 * Every rank must call allreduce_adam the same number of times.  A rank whose peers do not arrive within ~20 s gives
 * up waiting and sets the sticky flag dq_comm_status reports. */
#define DQ_COMM_HANDLE_BYTES 64
typedef struct dq_comm dq_comm;
int dq_comm_create(dq_comm** out, int rank, int world, int64_t n_floats, int device);
int dq_comm_destroy(dq_comm* comm);
int dq_comm_handle(dq_comm* comm, void* handle64);
int dq_comm_connect(dq_comm* comm, const void* handles /* world x 64 bytes, rank order */);
int dq_comm_next_grads(dq_comm* comm, float** grads);
/* params/m/v: this rank's flat buffers (n_floats, 16-byte aligned).  Same arithmetic as dq_adam_step on the rank-ordered
 * mean of the gradients; t = 1-based update index. */
int dq_comm_allreduce_adam(dq_comm* comm, float* params, float* m, float* v, float lr, float beta1, float beta2, float eps,
                           int64_t t, dq_stream stream);
int dq_comm_status(dq_comm* comm, int* timed_out);

/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
int64_t dq_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DQ_DECODING_H */
