/*
 * dq_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C, array-based restatement of the reference's surface-code
 * environment hot path, used ONLY by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py as the checker for the CUDA
 * path.  Nothing in deepq_decoding_b200/ may call into this file.
 *
 * Reference followed (paths relative to /root/reference, CS = cluster_scripts/d5_dp,
 * which is byte-identical to example_notebooks/ for Environments.py and equal to
 * example_notebooks/Function_Library.py minus its keras import block):
 *   lattice table ............ CS/Function_Library.py:13-51   (generateSurfaceCodeLattice)
 *   Pauli product ............ CS/Function_Library.py:54-62   (multiplyPaulis)
 *   error sampling ........... CS/Function_Library.py:67-122  (generate_error / _X_ / _DP_)
 *   syndrome ................. CS/Function_Library.py:152-174 (generate_surface_code_syndrome_NoFT_efficient)
 *   faulty syndrome .......... CS/Function_Library.py:176-223 (generate_faulty_syndrome)
 *   frame update ............. CS/Function_Library.py:226-241 (obtain_new_error_configuration)
 *   action -> move ........... CS/Function_Library.py:243-294 (index_to_move)
 *   homology label ........... CS/Function_Library.py:296-326 (generate_one_hot_labels_surface_code)
 *   env ctor/reset/step ...... CS/Environments.py:45-97, 99-115, 118-204, 206-235
 *   legal moves .............. CS/Environments.py:238-271, 187-196
 *   observation padding ...... CS/Environments.py:273-314
 *
 * The one thing that is NOT the reference's is the random stream: the reference
 * draws from numpy's global MT19937 one scalar at a time; here (and in the CUDA
 * path, and in oracle/ref_harness.py which drives the UNMODIFIED reference
 * classes) every draw is a word of a counter-based Philox4x32-10 stream, see
 * "RNG contract" below and DESIGN.md section 3.  Parity pin: this file is checked
 * against the imported reference by tests/test_oracle_vs_reference.py (runs
 * where /root/reference exists) and against tests/golden/ fixtures generated
 * from the imported reference by tests/golden/make_golden.py.
 *
 * The data layout is deliberately naive (one int8 per site, loops over sites)
 * so that it shares no bit-trick with the CUDA implementation.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXD 7
#define MAXG (MAXD + 1)          /* plaquette grid is (d+1) x (d+1) */
#define MAXVD 8
#define MAXA (3 * MAXD * MAXD + 1)
#define MAXW 3                   /* 64-bit words in a legal-action mask */

/* ------------------------------------------------------------------ */
/* RNG contract: Philox4x32-10, key = (seed_lo, seed_hi),
 * counter = (env_id, n, block, domain).
 *   domain 0 (noise):  n = volume-attempt index of that env (monotone from 0,
 *                      never reset), block b in [0,B), B = 32*ceil(items/128),
 *                      items = vd*(2 d^2 - 1).  Draw i (0 <= i < items) is word
 *                      (i / B) of block (i % B).  Draws 0 .. vd*d^2-1 are the
 *                      data-qubit draws, slice-major then qubit row-major; the
 *                      rest are measurement draws, slice-major then stabilizer
 *                      in the reference's own draw order (FL:176-223: bulk
 *                      row-major, top, bottom, left, right).
 *   domain 1 (policy): n = step index, block 0, word 0 -> uniform pick.
 * A draw u "fires with probability p" iff u < T(p), T(p) = floor(p * 2^32)
 * (saturating).  Depolarising: X iff u < T/3, Y iff T/3 <= u < 2T/3, Z iff
 * 2T/3 <= u < T (integer divisions) -- one draw per qubit.                    */
static inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                 uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static uint32_t threshold_u32(double p) {
    if (!(p > 0.0)) return 0u;
    double t = floor(p * 4294967296.0);
    if (t >= 4294967295.0) return 0xFFFFFFFFu;
    return (uint32_t)t;
}

/* ------------------------------------------------------------------ */
typedef struct {
    int8_t hidden[MAXD][MAXD];            /* Pauli frame, 0=I 1=X 2=Y 3=Z          */
    int8_t true_syn[MAXG][MAXG];          /* current_true_syndrome                 */
    int8_t faulty[MAXVD][MAXG][MAXG];     /* the volume shown in board_state[0:vd] */
    int8_t summed[MAXG][MAXG];            /* summed_syndrome_volume != 0           */
    int8_t completed[MAXA];               /* completed_actions                     */
    int8_t acted[MAXD * MAXD];            /* acted_on_qubits                       */
    int8_t legal[MAXA];                   /* legal_actions                         */
    int32_t lifetime;
    int32_t done;
    uint32_t attempts;                    /* volume-attempt counter (RNG index)    */
    uint32_t env_id;
} env_t;

typedef struct {
    int d, vd, model, use_Y, layers, A, H, C, ns, nq, B;
    uint32_t T, T1, T2, Tm;
    uint32_t k0, k1;
    int64_t n;
    int8_t ptype[MAXG][MAXG];             /* 0 absent, 1 / 3 plaquette type        */
    int stab_a[MAXG * MAXG], stab_b[MAXG * MAXG];   /* compact (draw) order        */
    int n3, n1;                           /* #type-3 / #type-1 stabilizers         */
    int joint_a[MAXG * MAXG], joint_b[MAXG * MAXG];   /* joint-table bit order       */
    int ref_mode;                         /* -1 none, 0 joint, 1 split             */
    int max_attempts;                     /* 0 = the reference's behaviour (redraw for ever); n > 0 = the product's documented
                                             deviation: after n all-trivial attempts the trivial volume is accepted       */
    const uint8_t* lut_a;                 /* joint table, or X-class table         */
    const uint8_t* lut_b;                 /* Z-class table (split, DP only)        */
    env_t* envs;
} oracle_t;

/* FL:13-51.  Plaquette (a,b), 0<=a,b<=d, type ((a+b)%2)*2+1; removed when
 * (a==0 && b even) || (a==d && b odd) || (b==0 && a odd) || (b==d && a even).  */
static void build_lattice(oracle_t* o) {
    int d = o->d;
    for (int a = 0; a <= d; ++a)
        for (int b = 0; b <= d; ++b) {
            int t = ((a + b) % 2) * 2 + 1;
            if (a == 0 && b % 2 == 0) t = 0;
            if (a == d && b % 2 == 1) t = 0;
            if (b == 0 && a % 2 == 1) t = 0;
            if (b == d && a % 2 == 0) t = 0;
            o->ptype[a][b] = (int8_t)t;
        }
    /* FL:176-223 draw order of generate_faulty_syndrome */
    int k = 0, g = d + 1;
    for (int a = 1; a < g - 1; ++a)
        for (int b = 1; b < g - 1; ++b) { o->stab_a[k] = a; o->stab_b[k] = b; ++k; }
    int nb = g / 2 - 1;
    for (int x = 0; x < nb; ++x) { o->stab_a[k] = 0;     o->stab_b[k] = 2 * x + 1; ++k; }
    for (int x = 0; x < nb; ++x) { o->stab_a[k] = g - 1; o->stab_b[k] = 2 * x + 2; ++k; }
    for (int x = 0; x < nb; ++x) { o->stab_a[k] = 2 * x + 2; o->stab_b[k] = 0;     ++k; }
    for (int x = 0; x < nb; ++x) { o->stab_a[k] = 2 * x + 1; o->stab_b[k] = g - 1; ++k; }
    o->ns = k;
    /* joint referee table bit order (include/dq_decoding.h): grid rows 1..d-1 left to right,
     * then the top/bottom boundary stabilizers by column */
    int m = 0;
    for (int a = 1; a < d; ++a)
        for (int b = 0; b <= d; ++b) if (o->ptype[a][b]) { o->joint_a[m] = a; o->joint_b[m] = b; ++m; }
    for (int b = 1; b < d; ++b) { o->joint_a[m] = (b % 2) ? 0 : d; o->joint_b[m] = b; ++m; }
    o->n3 = o->n1 = 0;
    for (int i = 0; i < k; ++i) {
        if (o->ptype[o->stab_a[i]][o->stab_b[i]] == 3) o->n3++; else o->n1++;
    }
}

/* FL:54-62 */
static inline int mult_paulis(int a, int b) {
    static const int8_t tab[4][4] = {{0,1,2,3},{1,0,3,2},{2,3,0,1},{3,2,1,0}};
    return tab[a][b];
}

/* FL:152-174: every present plaquette adjacent to qubit (i,j) whose type differs
 * from the error on that qubit toggles.  Qubit (i,j) touches (i,j),(i,j+1),(i+1,j),(i+1,j+1). */
static void true_syndrome(const oracle_t* o, env_t* e) {
    int d = o->d;
    memset(e->true_syn, 0, sizeof(e->true_syn));
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) {
            int err = e->hidden[i][j];
            if (!err) continue;
            for (int k = 0; k < 4; ++k) {
                int a = i + (k >> 1), b = j + (k & 1);
                int t = o->ptype[a][b];
                if (t != 0 && t != err) e->true_syn[a][b] ^= 1;
            }
        }
}

/* FL:296-326 */
static int homology_label(const oracle_t* o, const env_t* e) {
    int X = 0, Z = 0;
    for (int x = 0; x < o->d; ++x) if (e->hidden[x][0] == 1 || e->hidden[x][0] == 2) X ^= 1;
    for (int y = 0; y < o->d; ++y) if (e->hidden[0][y] == 3 || e->hidden[0][y] == 2) Z ^= 1;
    return X + 2 * Z;
}

static inline int lut2(const uint8_t* lut, uint64_t idx) {
    return (lut[idx >> 2] >> ((idx & 3) * 2)) & 3;
}

/* Stand-in for static_decoder.predict + argmax (Environments.py:144,150): the
 * referee is a table over the true syndrome (same table the CUDA path gets).   */
static int referee_class(const oracle_t* o, const env_t* e) {
    uint64_t all = 0, i3 = 0, i1 = 0; int c3 = 0, c1 = 0;
    for (int k = 0; k < o->ns; ++k) {
        int a = o->stab_a[k], b = o->stab_b[k];
        uint64_t bit = (uint64_t)e->true_syn[a][b];
        if (o->ptype[a][b] == 3) i3 |= bit << c3++; else i1 |= bit << c1++;
        all |= (uint64_t)e->true_syn[o->joint_a[k]][o->joint_b[k]] << k;
    }
    if (o->ref_mode == 0) return lut2(o->lut_a, all);
    if (o->ref_mode == 1) {
        int c = lut2(o->lut_a, i3) & 1;
        if (o->model == 1 && o->lut_b) c |= (lut2(o->lut_b, i1) & 1) << 1;
        return c;
    }
    return 0;
}

/* One noise word of the contract. */
static inline uint32_t noise_word(const oracle_t* o, const env_t* e, uint32_t attempt, int i,
                                  uint32_t* cache, int* cache_blk) {
    int b = i % o->B, w = i / o->B;
    if (*cache_blk != b) { philox4x32_10(e->env_id, attempt, (uint32_t)b, 0u, o->k0, o->k1, cache); *cache_blk = b; }
    return cache[w];
}

/* Environments.py:158-176 (step) == :216-235 (initialize_state): draw volumes of
 * vd noisy slices until the summed faulty syndrome is non-trivial.  Rejected
 * volumes keep their errors in the hidden frame and still advance lifetime.    */
static void new_volume(const oracle_t* o, env_t* e) {
    int d = o->d, g = d + 1, tries = 0;
    for (;;) {
        uint32_t attempt = e->attempts++;
        uint32_t cache[4]; int cb = -1;
        memset(e->summed, 0, sizeof(e->summed));
        int any = 0;
        for (int j = 0; j < o->vd; ++j) {
            for (int q = 0; q < o->nq; ++q) {           /* FL:67-122 */
                uint32_t u = noise_word(o, e, attempt, j * o->nq + q, cache, &cb);
                int err = 0;
                if (o->model == 0) err = (u < o->T) ? 1 : 0;
                else err = (u < o->T1) ? 1 : (u < o->T2) ? 2 : (u < o->T) ? 3 : 0;
                if (err) { int r = q / d, c = q % d; e->hidden[r][c] = (int8_t)mult_paulis(err, e->hidden[r][c]); }
            }
            true_syndrome(o, e);
            memset(e->faulty[j], 0, sizeof(e->faulty[j]));
            for (int k = 0; k < o->ns; ++k) {           /* FL:176-223 */
                uint32_t u = noise_word(o, e, attempt, o->vd * o->nq + j * o->ns + k, cache, &cb);
                int a = o->stab_a[k], b = o->stab_b[k];
                int v = e->true_syn[a][b] ^ (u < o->Tm ? 1 : 0);
                e->faulty[j][a][b] = (int8_t)v;
                if (v) { e->summed[a][b] = 1; any = 1; }
            }
            e->lifetime += 1;
        }
        (void)g;
        if (any || (o->max_attempts > 0 && ++tries >= o->max_attempts)) break;
    }
}

/* Environments.py:238-271 */
static void reset_legal_moves(const oracle_t* o, env_t* e) {
    int d = o->d;
    memset(e->completed, 0, sizeof(e->completed));
    memset(e->acted, 0, sizeof(e->acted));
    memset(e->legal, 0, sizeof(e->legal));
    e->legal[o->A - 1] = 1;
    for (int q = 0; q < o->nq; ++q) {
        int r = q / d, c = q % d, adj = 0;
        for (int k = 0; k < 4; ++k) {
            int a = r + (k >> 1), b = c + (k & 1);
            if (o->ptype[a][b] != 0 && e->summed[a][b]) adj = 1;
        }
        if (adj) for (int l = 0; l < o->layers; ++l) e->legal[q + l * o->nq] = 1;
    }
}

static void env_reset(const oracle_t* o, env_t* e) {   /* Environments.py:99-115, 206-235 */
    e->done = 0; e->lifetime = 0;
    memset(e->hidden, 0, sizeof(e->hidden));
    memset(e->true_syn, 0, sizeof(e->true_syn));
    new_volume(o, e);
    reset_legal_moves(o, e);
}

/* Environments.py:118-204 */
static float env_step(const oracle_t* o, env_t* e, int action) {
    int d = o->d, nq = o->nq;
    int done_identity = (action == o->A - 1) || e->completed[action] == 1;
    /* FL:243-294 index_to_move + FL:226-241 */
    if (action < o->A - 1) {
        int layer = action / nq, q = action % nq, type;
        if (o->model == 0) type = 1;
        else if (o->use_Y) type = layer + 1;
        else type = (layer == 0) ? 1 : 3;
        e->hidden[q / d][q % d] = (int8_t)mult_paulis(type, e->hidden[q / d][q % d]);
    }
    true_syndrome(o, e);
    int anyons = 0;
    for (int a = 0; a <= d; ++a) for (int b = 0; b <= d; ++b) anyons += e->true_syn[a][b];
    int label = homology_label(o, e);
    float reward = 0.0f;
    if (label == 0 && anyons == 0) reward = 1.0f;
    else if (referee_class(o, e) != label) e->done = 1;

    if (done_identity) {
        new_volume(o, e);
        reset_legal_moves(o, e);
    } else {
        e->completed[action] = 1;
        int q = action % nq;
        if (!e->acted[q]) {
            e->acted[q] = 1;
            int r = q / d, c = q % d;
            for (int dr = -1; dr <= 1; ++dr) for (int dc = -1; dc <= 1; ++dc) {
                if (!dr && !dc) continue;
                int rr = r + dr, cc = c + dc;
                if (rr < 0 || rr >= d || cc < 0 || cc >= d) continue;
                for (int l = 0; l < o->layers; ++l) e->legal[rr * d + cc + l * nq] = 1;
            }
        }
    }
    return reward;
}

/* Environments.py:273-314: board_state as uint8 [C][H][H]. */
static void write_obs(const oracle_t* o, const env_t* e, uint8_t* obs) {
    int d = o->d, H = o->H;
    memset(obs, 0, (size_t)o->C * H * H);
    for (int j = 0; j < o->vd; ++j) {
        uint8_t* L = obs + (size_t)j * H * H;
        for (int x = 0; x < H; ++x) for (int y = 0; y < H; ++y) {
            int v = 0;
            if ((x == 0 || x == 2 * d) && y % 2 == 1) v = 1;
            if ((y == 0 || y == 2 * d) && x % 2 == 1) v = 1;
            if (x % 2 == 0 && y % 2 == 0) v = e->faulty[j][x / 2][y / 2];
            else if (x % 2 == 1 && y % 2 == 1 && (x + y) % 4 == 0) v = 1;
            L[x * H + y] = (uint8_t)v;
        }
    }
    for (int l = 0; l < o->layers; ++l) {
        uint8_t* L = obs + (size_t)(o->vd + l) * H * H;
        for (int q = 0; q < o->nq; ++q)
            if (e->completed[q + l * o->nq]) L[(2 * (q / d) + 1) * H + 2 * (q % d) + 1] = 1;
    }
}

static void write_legal(const oracle_t* o, const env_t* e, uint64_t* mask, int W) {
    for (int w = 0; w < W; ++w) mask[w] = 0;
    for (int a = 0; a < o->A; ++a) if (e->legal[a]) mask[a >> 6] |= 1ull << (a & 63);
}

/* ------------------------------------------------------------------ */
/* exported API (ctypes)                                               */

void* dqo_create(int d, int model, int use_Y, int vd, double p_phys, double p_meas,
                 int64_t n_envs, uint64_t seed, int64_t env_id_base) {
    if (d < 3 || d > MAXD || d % 2 == 0 || vd < 1 || vd > MAXVD || n_envs < 1) return NULL;
    oracle_t* o = (oracle_t*)calloc(1, sizeof(oracle_t));
    o->d = d; o->vd = vd; o->model = model; o->use_Y = use_Y; o->n = n_envs;
    o->nq = d * d;
    o->layers = (model == 0) ? 1 : (use_Y ? 3 : 2);
    o->A = o->layers * o->nq + 1;
    o->H = 2 * d + 1; o->C = vd + o->layers;
    build_lattice(o);
    int items = vd * (o->nq + o->ns);
    o->B = 32 * ((items + 127) / 128);
    o->k0 = (uint32_t)seed; o->k1 = (uint32_t)(seed >> 32);
    o->ref_mode = -1;
    o->envs = (env_t*)calloc((size_t)n_envs, sizeof(env_t));
    for (int64_t i = 0; i < n_envs; ++i) o->envs[i].env_id = (uint32_t)(env_id_base + i);
    o->T = threshold_u32(p_phys); o->Tm = threshold_u32(p_meas);
    o->T1 = o->T / 3; o->T2 = (uint32_t)((2ull * o->T) / 3);
    return o;
}

void dqo_destroy(void* h) { oracle_t* o = (oracle_t*)h; if (o) { free(o->envs); free(o); } }

void dqo_set_max_attempts(void* h, int n) { ((oracle_t*)h)->max_attempts = n; }

void dqo_set_noise(void* h, double p_phys, double p_meas) {
    oracle_t* o = (oracle_t*)h;
    o->T = threshold_u32(p_phys); o->Tm = threshold_u32(p_meas);
    o->T1 = o->T / 3; o->T2 = (uint32_t)((2ull * o->T) / 3);
}

/* mode 0: lut_a = joint 2-bit table over all ns stabilizers (joint order: rows 1..d-1, then top/bottom by column);
 * mode 1: lut_a = X-class table over type-3 stabilizers, lut_b = Z-class table over
 *         type-1 stabilizers (each in compact order restricted to its type).    */
void dqo_set_referee(void* h, int mode, const uint8_t* lut_a, const uint8_t* lut_b) {
    oracle_t* o = (oracle_t*)h; o->ref_mode = mode; o->lut_a = lut_a; o->lut_b = lut_b;
}

int dqo_info(void* h, int what) {
    oracle_t* o = (oracle_t*)h;
    switch (what) { case 0: return o->A; case 1: return o->C; case 2: return o->H;
                    case 3: return o->ns; case 4: return o->B; case 5: return o->n3; case 6: return o->n1;
                    case 7: return (o->A + 63) / 64; }
    return -1;
}

void dqo_reset(void* h, uint8_t* obs, uint64_t* legal) {
    oracle_t* o = (oracle_t*)h;
    int W = (o->A + 63) / 64; size_t L = (size_t)o->C * o->H * o->H;
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < o->n; ++i) {
        env_reset(o, &o->envs[i]);
        if (obs) write_obs(o, &o->envs[i], obs + i * L);
        if (legal) write_legal(o, &o->envs[i], legal + i * W, W);
    }
}

/* lifetime[i] = env.lifetime after the step (the finished episode's value when
 * done); with auto_reset the returned obs / legal mask belong to the new episode. */
void dqo_step(void* h, const int32_t* actions, uint8_t* obs, float* reward, uint8_t* done,
              int32_t* lifetime, uint64_t* legal, int auto_reset) {
    oracle_t* o = (oracle_t*)h;
    int W = (o->A + 63) / 64; size_t L = (size_t)o->C * o->H * o->H;
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < o->n; ++i) {
        env_t* e = &o->envs[i];
        float r = env_step(o, e, actions[i]);
        if (reward) reward[i] = r;
        if (done) done[i] = (uint8_t)e->done;
        if (lifetime) lifetime[i] = e->lifetime;
        if (auto_reset && e->done) env_reset(o, e);
        if (obs) write_obs(o, e, obs + i * L);
        if (legal) write_legal(o, e, legal + i * W, W);
    }
}

/* Uniform pick over the sorted legal actions: index = floor(u * n / 2^32). */
void dqo_random_legal_actions(void* h, const uint64_t* legal, uint32_t step, int32_t* actions) {
    oracle_t* o = (oracle_t*)h;
    int W = (o->A + 63) / 64;
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < o->n; ++i) {
        uint32_t w[4];
        philox4x32_10(o->envs[i].env_id, step, 0u, 1u, o->k0, o->k1, w);
        int n = 0;
        for (int a = 0; a < o->A; ++a) n += (int)((legal[i * W + (a >> 6)] >> (a & 63)) & 1);
        int pick = (int)(((uint64_t)w[0] * (uint64_t)n) >> 32), act = o->A - 1;
        for (int a = 0; a < o->A; ++a)
            if ((legal[i * W + (a >> 6)] >> (a & 63)) & 1) { if (pick-- == 0) { act = a; break; } }
        actions[i] = act;
    }
}

/* Debug / parity views: hidden frame [d*d], true syndrome [(d+1)^2], counters. */
void dqo_get_env(void* h, int64_t i, int8_t* hidden, int8_t* true_syn, int32_t* lifetime,
                 int32_t* done, uint32_t* attempts) {
    oracle_t* o = (oracle_t*)h; env_t* e = &o->envs[i]; int d = o->d;
    if (hidden)   for (int r = 0; r < d; ++r) for (int c = 0; c < d; ++c) hidden[r * d + c] = e->hidden[r][c];
    if (true_syn) for (int a = 0; a <= d; ++a) for (int b = 0; b <= d; ++b) true_syn[a * (d + 1) + b] = e->true_syn[a][b];
    if (lifetime) *lifetime = e->lifetime;
    if (done) *done = e->done;
    if (attempts) *attempts = e->attempts;
}

/* Parity tests inject a hidden frame (row-major Paulis) into env i. */
void dqo_set_hidden(void* h, int64_t i, const int8_t* hidden) {
    oracle_t* o = (oracle_t*)h; env_t* e = &o->envs[i]; int d = o->d;
    for (int r = 0; r < d; ++r) for (int c = 0; c < d; ++c) e->hidden[r][c] = hidden[r * d + c];
    true_syndrome(o, e);
}

/* Stateless helpers for unit parity of single functions. */
void dqo_syndrome_of(void* h, const int8_t* hidden, int8_t* syn, int* label) {
    oracle_t* o = (oracle_t*)h; env_t e; memset(&e, 0, sizeof(e)); int d = o->d;
    for (int r = 0; r < d; ++r) for (int c = 0; c < d; ++c) e.hidden[r][c] = hidden[r * d + c];
    true_syndrome(o, &e);
    for (int a = 0; a <= d; ++a) for (int b = 0; b <= d; ++b) syn[a * (d + 1) + b] = e.true_syn[a][b];
    if (label) *label = homology_label(o, &e);
}

void dqo_stab_order(void* h, int32_t* a, int32_t* b, int32_t* type) {
    oracle_t* o = (oracle_t*)h;
    for (int k = 0; k < o->ns; ++k) { a[k] = o->stab_a[k]; b[k] = o->stab_b[k]; type[k] = o->ptype[o->stab_a[k]][o->stab_b[k]]; }
}

void dqo_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
    philox4x32_10(c0, c1, c2, c3, k0, k1, out);
}

/* torchrun exports OMP_NUM_THREADS=1; the CPU baseline must use the cores it reports */
void dqo_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int dqo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
