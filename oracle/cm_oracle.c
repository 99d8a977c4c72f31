/*
 * cm_oracle.c — CPU restatement of the Com-MARL rollout hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the *checker*, never the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may build, load or call it.  The shipped
 * path (com_marl_b200/csrc) never links or falls back to it.
 *
 * It restates, in plain scalar C and in the reference's own data model (a grid of cell codes,
 * per-entity position lists, sequential loops), the algorithms of
 *   envs/ma_gym/envs/predator_prey/predator_prey.py   (PredatorPrey dynamics + observations)
 *   envs/ma_gym/envs/coverage/coverage.py             (Coverage dynamics + observations)
 *   custom_implement/env_communication.py             (adjacency graph, FC/FL/IID/GE channels)
 *   custom_implement/gilbert_elliot_loss_model.py     (GE link Markov chain)
 *   garage/sampler/vec_env_executor.py                (step-all + time limit + auto-reset)
 * of cnuns/Com-MARL; every function cites the file:line it follows.  Nothing is copied: the
 * reference is Python over string cells, this is C over int16 cell codes.
 *
 * PARITY PINNING: the reference ships no tests or golden vectors, so this oracle is
 * pinned against the fixtures in tests/golden/ (npz) — outputs of the unmodified reference itself, recorded by
 * tests/golden/make_golden.py with injected prey-move / packet-loss streams
 * (tests/test_oracle_golden.py).
 *
 * Random streams.  "Injected" mode consumes caller-supplied prey candidates, channel uniforms and
 * spawn positions (the reference-parity mode).  "Generated" mode draws them from Philox4x32-10
 * keyed by (seed; env id, tick, stream|episode<<8, index) — the production stream specification,
 * which the CUDA engine implements independently and must match bit for bit.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (see oracle/Makefile).  -ffp-contract=off keeps
 * the fp64 reward expressions in the reference's left-to-right evaluation order.
 */
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>

enum { ORC_PP = 0, ORC_CO = 1 };
enum { CH_FC = 0, CH_FL = 1, CH_IID = 2, CH_GE = 3 };
enum { STREAM_SPAWN = 1, STREAM_PREY = 2, STREAM_CHAN = 3, STREAM_ACT = 4 };
enum { ORC_OK = 0, ORC_EBADACTION = -1, ORC_ECONFIG = -2, ORC_ESPAWN = -3 };

#define CELL_WALL 32767

typedef struct {
    int32_t scenario, n, p, G, R, T, load, L, max_path_length;
    int32_t n_empty_cells, rcom2, chan_type, loss_apply, ge_init;
    float p_loss, pgb, pbg, ge_bad_rate;
    double capture_reward, step_cost, moving_cost, penalty, lazy_penalty, revisit_penalty, final_reward;
    uint64_t seed;
    int64_t env_id0;            /* global id of env 0 of this shard (RNG keys use global ids) */
    const uint8_t *wall;        /* [G*G] 1 = wall (Coverage) */
    const float *lut_row;       /* [G]   obs row feature   */
    const float *lut_col;       /* [G]   obs col feature   */
    const float *lut_t;         /* [T+1] obs time feature (PredatorPrey) */
} orc_cfg;

typedef struct {
    int16_t *grid;              /* [B][G*G] 0 empty | +k agent k | -k prey k | CELL_WALL */
    int8_t *apos;               /* [B][n][2] */
    int8_t *ppos;               /* [B][p][2] */
    uint8_t *alive;             /* [B][p] */
    uint8_t *visited;           /* [B][G*G] */
    int32_t *t;                 /* [B] env step count */
    int32_t *total_capture;     /* [B] */
    uint8_t *success;           /* [B] latched like env.success */
    uint32_t *episode;          /* [B] resets so far */
    uint32_t *tick;             /* [B] steps since creation (never reset) */
    uint8_t *ge_state;          /* [B][n][n] last Gilbert-Elliot link state */
} orc_state;

typedef struct {
    const int8_t *actions;      /* [B][n] */
    const int8_t *prey_cand;    /* [B][p][5] or NULL -> Philox */
    const int8_t *spawn_agent;  /* [B][E][n][2] or NULL -> Philox */
    const int8_t *spawn_prey;   /* [B][E][p][2] */
    int32_t spawn_episodes;     /* E */
    const float *chan_u;        /* [B][planes][n][n] or NULL -> Philox */
    int32_t chan_planes;
    int32_t auto_reset;
    float *obs;                 /* [B][n][D] */
    double *reward;             /* [B] */
    uint8_t *done;              /* [B] */
    int32_t *counts;            /* [B][6] */
    uint8_t *prey_alive_out;    /* [B][p] flags returned by the step (pre-reset) */
    uint8_t *adj;               /* [B][n][n] */
    uint8_t *chan;              /* [B][L][n][n] */
    float *ave_deg;             /* [B] */
} orc_io;

/* action -> displacement: 0 down(+row) 1 left(-col) 2 up(-row) 3 right(+col) 4 noop
 * (predator_prey.py:240-253, 640-646; coverage.py:336-345) */
static const int DR[5] = {1, 0, -1, 0, 0};
static const int DC[5] = {0, -1, 0, 1, 0};

/* ---------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al., SC'11) — the published algorithm, restated.
 * ------------------------------------------------------------------------------------------- */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

static void rng_block(const orc_cfg *cfg, int64_t env, uint32_t tick, uint32_t stream, uint32_t episode,
                      uint32_t index, uint32_t out[4])
{
    out[0] = (uint32_t)(cfg->env_id0 + env);
    out[1] = tick;
    out[2] = stream | (episode << 8);
    out[3] = index;
    philox4x32_10(out, (uint32_t)cfg->seed, (uint32_t)(cfg->seed >> 32));
}

/* exported for tests: raw generator access */
void orc_philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4])
{
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    philox4x32_10(out, (uint32_t)seed, (uint32_t)(seed >> 32));
}

static inline float u24(uint32_t w) { return (float)(w >> 8) * 5.9604644775390625e-08f; /* 2^-24 */ }

/* u32 -> prey move through the cumulative distribution of (.175,.175,.175,.175,.3)
 * (predator_prey.py:54,401), thresholds = floor(cdf * 2^32). */
static inline int prey_move_from_bits(uint32_t w)
{
    return (w >= 751619276u) + (w >= 1503238553u) + (w >= 2254857830u) + (w >= 3006477107u);
}

/* ---------------------------------------------------------------------------------------------
 * helpers over one env's grid
 * ------------------------------------------------------------------------------------------- */
static inline int inb(const orc_cfg *cfg, int r, int c) { return r >= 0 && r < cfg->G && c >= 0 && c < cfg->G; }

/* predator_prey.py:309-329 _neighbour_agents: agents in the 4-neighbourhood of pos (pos may be
 * out of bounds; each neighbour is bounds-checked).  ids[] receives their 0-based indices. */
static int neighbour_agents(const orc_cfg *cfg, const int16_t *grid, int r, int c, int ids[4])
{
    static const int NR[4] = {1, -1, 0, 0}, NC[4] = {0, 0, 1, -1};
    int k = 0;
    for (int q = 0; q < 4; ++q) {
        int rr = r + NR[q], cc = c + NC[q];
        if (inb(cfg, rr, cc)) {
            int16_t v = grid[rr * cfg->G + cc];
            if (v > 0 && v != CELL_WALL) { if (ids) ids[k] = v - 1; ++k; }
        }
    }
    return k;
}

/* predator_prey.py:353-382 _neighbour_objects: counts of agents and preys around pos */
static void neighbour_objects(const orc_cfg *cfg, const int16_t *grid, int r, int c, int *na, int *np_, int ids[4])
{
    static const int NR[4] = {1, -1, 0, 0}, NC[4] = {0, 0, 1, -1};
    int a = 0, pr = 0;
    for (int q = 0; q < 4; ++q) {
        int rr = r + NR[q], cc = c + NC[q];
        if (inb(cfg, rr, cc)) {
            int16_t v = grid[rr * cfg->G + cc];
            if (v > 0 && v != CELL_WALL) { ids[a++] = v - 1; }
            else if (v < 0) ++pr;
        }
    }
    *na = a; *np_ = pr;
}

/* predator_prey.py:123-144,384-385: corner 2, border 3, interior -> self.load (sic) */
static int n_adjacent_grid(const orc_cfg *cfg, int r, int c)
{
    int e = cfg->G - 1;
    int re = (r == 0 || r == e), ce = (c == 0 || c == e);
    if (re && ce) return 2;
    if (re || ce) return 3;
    return cfg->load;
}

/* ---------------------------------------------------------------------------------------------
 * observations
 * ------------------------------------------------------------------------------------------- */
/* predator_prey.py:173-204 get_neighbors/get_agent_obs: [(2R+1)^2 agent window][(2R+1)^2 prey window]
 * [row/m, col/(m-1), t/T]; out-of-range cells stay 0. */
static void pp_obs(const orc_cfg *cfg, const orc_state *st, int64_t b, float *obs)
{
    const int G = cfg->G, R = cfg->R, w = 2 * R + 1, D = 2 * w * w + 3;
    const int16_t *grid = st->grid + b * G * G;
    for (int i = 0; i < cfg->n; ++i) {
        float *o = obs + (b * cfg->n + i) * D;
        int r0 = st->apos[(b * cfg->n + i) * 2], c0 = st->apos[(b * cfg->n + i) * 2 + 1];
        memset(o, 0, sizeof(float) * D);
        for (int row = (r0 - R > 0 ? r0 - R : 0); row < (r0 + R + 1 < G ? r0 + R + 1 : G); ++row)
            for (int col = (c0 - R > 0 ? c0 - R : 0); col < (c0 + R + 1 < G ? c0 + R + 1 : G); ++col) {
                int16_t v = grid[row * G + col];
                int idx = (row - (r0 - R)) * w + (col - (c0 - R));
                if (v > 0) o[idx] = 1.0f;
                if (v < 0) o[w * w + idx] = 1.0f;
            }
        o[2 * w * w + 0] = cfg->lut_row[r0];
        o[2 * w * w + 1] = cfg->lut_col[c0];
        o[2 * w * w + 2] = cfg->lut_t[st->t[b]];
    }
}

/* coverage.py:198-212,448-480: 3 x (2R+1)^2 window [wall | agent | visited] + rounded (row, col);
 * out-of-grid cells count as wall. */
static void co_obs(const orc_cfg *cfg, const orc_state *st, int64_t b, float *obs)
{
    const int G = cfg->G, R = cfg->R, w = 2 * R + 1, D = 3 * w * w + 2;
    const int16_t *grid = st->grid + b * G * G;
    const uint8_t *vis = st->visited + b * G * G;
    for (int i = 0; i < cfg->n; ++i) {
        float *o = obs + (b * cfg->n + i) * D;
        int r0 = st->apos[(b * cfg->n + i) * 2], c0 = st->apos[(b * cfg->n + i) * 2 + 1];
        memset(o, 0, sizeof(float) * D);
        for (int row = r0 - R; row <= r0 + R; ++row)
            for (int col = c0 - R; col <= c0 + R; ++col) {
                int idx = (row - (r0 - R)) * w + (col - (c0 - R));
                if (!inb(cfg, row, col)) { o[idx] = 1.0f; continue; }
                int16_t v = grid[row * G + col];
                if (v == CELL_WALL) o[idx] = 1.0f;
                else if (v > 0) o[w * w + idx] = 1.0f;
                if (vis[row * G + col]) o[2 * w * w + idx] = 1.0f;
            }
        o[3 * w * w + 0] = cfg->lut_row[r0];
        o[3 * w * w + 1] = cfg->lut_col[c0];
    }
}

/* ---------------------------------------------------------------------------------------------
 * communication state (env_communication.py:91-157, 200-243; gilbert_elliot_loss_model.py:84-150)
 * ------------------------------------------------------------------------------------------- */
typedef struct { const orc_cfg *cfg; int64_t b; uint32_t tick, episode; const float *u; int planes; int next_plane; } chan_src;

static float chan_uniform(chan_src *s, int plane, int i, int j)
{
    const int n = s->cfg->n;
    if (s->u) return s->u[(((int64_t)s->b * s->planes + plane) * n + i) * n + j];
    uint32_t q = (uint32_t)((plane * n + i) * n + j), blk[4];
    rng_block(s->cfg, s->b, s->tick, STREAM_CHAN, s->episode, q >> 2, blk);
    return u24(blk[q & 3]);
}

/* one Gilbert-Elliot transition of the whole link matrix (gilbert_elliot_loss_model.py:137-148):
 * g2b draw first, then b2g draw; "+ eye" freezes the diagonal. */
static void ge_transition(chan_src *s, uint8_t *state)
{
    const orc_cfg *cfg = s->cfg;
    const int n = cfg->n, pg = s->next_plane, pb = s->next_plane + 1;
    s->next_plane += 2;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            float eye = (i == j) ? 1.0f : 0.0f;
            int g2b = (chan_uniform(s, pg, i, j) + eye) < cfg->pgb;
            int b2g = (chan_uniform(s, pb, i, j) + eye) < cfg->pbg;
            int sij = state[i * n + j];
            int g_next = sij && !(sij && g2b);
            int b_next = !sij && b2g;
            state[i * n + j] = (uint8_t)(g_next || b_next);
        }
}

static int comm_update(const orc_cfg *cfg, orc_state *st, int64_t b, const orc_io *io, int at_reset)
{
    const int n = cfg->n, L = cfg->L;
    uint8_t *adj = io->adj + b * n * n;
    uint8_t *ch = io->chan + b * L * n * n;
    /* env_communication.py:218-243 get_graph: Rcom==0 -> all ones, ave_deg = n; else
     * cdist(pos) <= sqrt(2 Rcom^2)  <=>  dr^2+dc^2 <= 2 Rcom^2 on integer coordinates. */
    if (cfg->rcom2 < 0) {
        memset(adj, 1, (size_t)n * n);
        io->ave_deg[b] = (float)n;
    } else {
        int deg = 0;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                int dr = st->apos[(b * n + i) * 2] - st->apos[(b * n + j) * 2];
                int dc = st->apos[(b * n + i) * 2 + 1] - st->apos[(b * n + j) * 2 + 1];
                int a = (dr * dr + dc * dc) <= cfg->rcom2;
                adj[i * n + j] = (uint8_t)a;
                deg += a;
            }
        io->ave_deg[b] = (float)deg / (float)n; /* float32 sum(axis=1).mean() (:232) */
    }
    chan_src s = { cfg, b, st->tick[b], st->episode[b], io->chan_u, io->chan_planes, 0 };
    switch (cfg->chan_type) {
    case CH_FC: memset(ch, 1, (size_t)L * n * n); break;                          /* :93-95 */
    case CH_FL:                                                                     /* :97-100 */
        memset(ch, 0, (size_t)L * n * n);
        for (int l = 0; l < L; ++l) for (int i = 0; i < n; ++i) ch[(l * n + i) * n + i] = 1;
        break;
    case CH_IID:                                                                    /* :200-214 */
        for (int l = 0; l < L; ++l)
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j) {
                    float eye = (i == j) ? 1.0f : 0.0f;
                    ch[(l * n + i) * n + j] = (uint8_t)((chan_uniform(&s, l, i, j) + eye) >= cfg->p_loss);
                }
        break;
    case CH_GE: {                                                                   /* :106-157 */
        uint8_t *state = st->ge_state + b * n * n;
        if (cfg->ge_init == -1 && cfg->loss_apply == 0) return ORC_ECONFIG; /* broken in the reference (:121) */
        if (at_reset) {
            if (cfg->ge_init == 1) memset(state, 1, (size_t)n * n);
            else if (cfg->ge_init == 0) memset(state, 0, (size_t)n * n);
            else {                                        /* gilbert_elliot_loss_model.py:84-87 */
                for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j)
                    state[i * n + j] = (uint8_t)(chan_uniform(&s, 0, i, j) >= cfg->ge_bad_rate);
                s.next_plane = 1;
            }
            if (cfg->loss_apply == 0) {
                for (int l = 0; l < L; ++l) memcpy(ch + l * n * n, state, (size_t)n * n);
            } else {                                      /* include_prev=True, L-1 transitions */
                memcpy(ch, state, (size_t)n * n);
                for (int l = 1; l < L; ++l) { ge_transition(&s, state); memcpy(ch + l * n * n, state, (size_t)n * n); }
            }
        } else if (cfg->loss_apply == 0) {
            ge_transition(&s, state);
            for (int l = 0; l < L; ++l) memcpy(ch + l * n * n, state, (size_t)n * n);
        } else {
            for (int l = 0; l < L; ++l) { ge_transition(&s, state); memcpy(ch + l * n * n, state, (size_t)n * n); }
        }
        break; }
    default: return ORC_ECONFIG;
    }
    return ORC_OK;
}

/* ---------------------------------------------------------------------------------------------
 * reset / spawn  (predator_prey.py:150-171,206-232; coverage.py:172-196,221-246)
 * ------------------------------------------------------------------------------------------- */
typedef struct { const orc_cfg *cfg; int64_t b; uint32_t tick, episode, ctr; uint32_t blk[4]; } spawn_src;

static void spawn_draw(spawn_src *s, int lo, int span, int *r, int *c)
{
    if ((s->ctr & 1u) == 0) rng_block(s->cfg, s->b, s->tick, STREAM_SPAWN, s->episode, s->ctr >> 1, s->blk);
    uint32_t wr = s->blk[2 * (s->ctr & 1u)], wc = s->blk[2 * (s->ctr & 1u) + 1];
    s->ctr++;
    *r = lo + (int)(((uint64_t)wr * (uint32_t)span) >> 32);
    *c = lo + (int)(((uint64_t)wc * (uint32_t)span) >> 32);
}

static int reset_env(const orc_cfg *cfg, orc_state *st, int64_t b, const orc_io *io)
{
    const int G = cfg->G, n = cfg->n, p = cfg->p;
    int16_t *grid = st->grid + b * G * G;
    uint32_t ep = st->episode[b];
    for (int k = 0; k < G * G; ++k) grid[k] = (cfg->scenario == ORC_CO && cfg->wall[k]) ? CELL_WALL : 0;
    if (cfg->scenario == ORC_CO) memset(st->visited + b * G * G, 0, (size_t)G * G);
    spawn_src s = { cfg, b, st->tick[b], ep, 0, {0, 0, 0, 0} };
    const int lo = cfg->scenario == ORC_CO ? 1 : 0, span = cfg->scenario == ORC_CO ? G - 2 : G;
    for (int i = 0; i < n; ++i) {
        int r, c;
        if (io->spawn_agent) {
            if ((int32_t)ep >= io->spawn_episodes) return ORC_ESPAWN;
            const int8_t *q = io->spawn_agent + (((int64_t)b * io->spawn_episodes + ep) * n + i) * 2;
            r = q[0]; c = q[1];
            if (!inb(cfg, r, c) || grid[r * G + c] != 0) return ORC_ESPAWN;
        } else {
            for (int tries = 0;; ++tries) {   /* rejection sampling, predator_prey.py:155-159 / coverage.py:184-194 */
                if (tries > 1 << 20) return ORC_ESPAWN;
                spawn_draw(&s, lo, span, &r, &c);
                if (grid[r * G + c] == 0) break;
            }
        }
        st->apos[(b * n + i) * 2] = (int8_t)r; st->apos[(b * n + i) * 2 + 1] = (int8_t)c;
        grid[r * G + c] = (int16_t)(i + 1);
        if (cfg->scenario == ORC_CO) st->visited[b * G * G + r * G + c] = 1;   /* coverage.py:188 */
    }
    for (int j = 0; j < p; ++j) {
        int r, c;
        if (io->spawn_prey) {
            const int8_t *q = io->spawn_prey + (((int64_t)b * io->spawn_episodes + ep) * p + j) * 2;
            r = q[0]; c = q[1];
            if (!inb(cfg, r, c) || grid[r * G + c] != 0) return ORC_ESPAWN;
        } else {
            for (int tries = 0;; ++tries) {   /* vacant and no agent in the 4-neighbourhood, predator_prey.py:164-168 */
                if (tries > 1 << 20) return ORC_ESPAWN;
                spawn_draw(&s, lo, span, &r, &c);
                if (grid[r * G + c] == 0 && neighbour_agents(cfg, grid, r, c, NULL) == 0) break;
            }
        }
        st->ppos[(b * p + j) * 2] = (int8_t)r; st->ppos[(b * p + j) * 2 + 1] = (int8_t)c;
        grid[r * G + c] = (int16_t)(-(j + 1));
        st->alive[b * p + j] = 1;
    }
    st->t[b] = 0;
    st->total_capture[b] = 0;
    st->episode[b] = ep + 1;
    return ORC_OK;
}

static int emit_obs_comm(const orc_cfg *cfg, orc_state *st, int64_t b, const orc_io *io, int at_reset)
{
    if (cfg->scenario == ORC_PP) pp_obs(cfg, st, b, io->obs); else co_obs(cfg, st, b, io->obs);
    return comm_update(cfg, st, b, io, at_reset);
}

/* reset the masked envs (mask NULL = all) and write their observation / comm outputs */
int orc_reset(const orc_cfg *cfg, orc_state *st, const orc_io *io, int64_t B, const uint8_t *mask)
{
    for (int64_t b = 0; b < B; ++b) {
        if (mask && !mask[b]) continue;
        int rc = reset_env(cfg, st, b, io);
        if (rc) return rc;
        /* comm draws at reset are keyed with the post-increment episode */
        rc = emit_obs_comm(cfg, st, b, io, 1);
        if (rc) return rc;
    }
    return ORC_OK;
}

/* ---------------------------------------------------------------------------------------------
 * PredatorPrey.step  (predator_prey.py:494-519)
 * ------------------------------------------------------------------------------------------- */
static int pp_step_env(const orc_cfg *cfg, orc_state *st, int64_t b, const orc_io *io, int *env_done)
{
    const int G = cfg->G, n = cfg->n, p = cfg->p;
    int16_t *grid = st->grid + b * G * G;
    const int8_t *act = io->actions + b * n;
    st->t[b] += 1;
    int moved_sum = 0;
    for (int i = 0; i < n; ++i) {                         /* :497-500 -> :240-261, lower index first */
        int a = act[i];
        if (a < 0 || a > 4) return ORC_EBADACTION;
        if (a != 4) {
            ++moved_sum;
            int r = st->apos[(b * n + i) * 2], c = st->apos[(b * n + i) * 2 + 1];
            int nr = r + DR[a], nc = c + DC[a];
            if (inb(cfg, nr, nc) && grid[nr * G + nc] == 0) {
                grid[r * G + c] = 0;
                grid[nr * G + nc] = (int16_t)(i + 1);
                st->apos[(b * n + i) * 2] = (int8_t)nr; st->apos[(b * n + i) * 2 + 1] = (int8_t)nc;
            }
        }
    }
    int capture = 0, penalty = 0, watching_sum = 0;
    uint8_t watching[512];
    memset(watching, 0, sizeof watching);
    if (cfg->load != 2 && cfg->load != 3 && cfg->load != 4) return ORC_ECONFIG;   /* :77-79 */
    for (int j = 0; j < p; ++j) {                         /* :417-432 / :460-478 */
        if (!st->alive[b * p + j]) continue;
        int r = st->ppos[(b * p + j) * 2], c = st->ppos[(b * p + j) * 2 + 1];
        int ids[4], k, preyn = 0;
        if (cfg->load == 2) k = neighbour_agents(cfg, grid, r, c, ids);
        else neighbour_objects(cfg, grid, r, c, &k, &preyn, ids);
        for (int q = 0; q < k; ++q) watching[ids[q]] = 1;
        if (k >= 1) {
            int need = cfg->load;
            if (cfg->load != 2) {                         /* :469-470 */
                int avail = n_adjacent_grid(cfg, r, c) - preyn;
                need = cfg->load < avail ? cfg->load : avail;
            }
            if (need <= k) { ++capture; st->alive[b * p + j] = 0; }
            else ++penalty;
        }
        /* prey_random_move :396-407 */
        if (st->alive[b * p + j]) {
            int mv = 4;
            uint32_t blk[8];
            if (!io->prey_cand) {
                rng_block(cfg, b, st->tick[b], STREAM_PREY, st->episode[b], (uint32_t)(2 * j), blk);
                rng_block(cfg, b, st->tick[b], STREAM_PREY, st->episode[b], (uint32_t)(2 * j + 1), blk + 4);
            }
            for (int tr = 0; tr < 5; ++tr) {
                int cnd = io->prey_cand ? io->prey_cand[((int64_t)b * p + j) * 5 + tr] : prey_move_from_bits(blk[tr]);
                if (neighbour_agents(cfg, grid, r + DR[cnd], c + DC[cnd], NULL) == 0) { mv = cnd; break; }
            }
            if (mv != 4) {                                /* __update_prey_pos :276-299 */
                int nr = r + DR[mv], nc = c + DC[mv];
                if (inb(cfg, nr, nc) && grid[nr * G + nc] == 0) {
                    grid[r * G + c] = 0;
                    grid[nr * G + nc] = (int16_t)(-(j + 1));
                    st->ppos[(b * p + j) * 2] = (int8_t)nr; st->ppos[(b * p + j) * 2 + 1] = (int8_t)nc;
                }
            }
        } else {
            grid[r * G + c] = 0;                          /* :300-301 captured prey leaves the grid */
        }
    }
    for (int i = 0; i < n; ++i) watching_sum += watching[i];
    /* :434 / :480 — fixed left-to-right fp64 evaluation */
    double rew = (cfg->step_cost + cfg->capture_reward * capture + cfg->moving_cost * moved_sum / n);
    if (cfg->load == 2) rew = rew + cfg->penalty * penalty;
    io->reward[b] = rew;
    int32_t *cn = io->counts + b * 6;
    cn[0] = capture; cn[1] = moved_sum; cn[2] = penalty; cn[3] = watching_sum; cn[4] = 0; cn[5] = 0;
    int any_alive = 0;
    for (int j = 0; j < p; ++j) { any_alive |= st->alive[b * p + j]; io->prey_alive_out[b * p + j] = st->alive[b * p + j]; }
    *env_done = 0;
    if (st->t[b] >= cfg->T || !any_alive) {               /* :511-517 */
        st->success[b] = any_alive ? 0 : 1;
        *env_done = 1;
    }
    return ORC_OK;
}

/* ---------------------------------------------------------------------------------------------
 * Coverage.step  (coverage.py:319-401, get_reward :299-317)
 * ------------------------------------------------------------------------------------------- */
static int co_step_env(const orc_cfg *cfg, orc_state *st, int64_t b, const orc_io *io, int *env_done)
{
    const int G = cfg->G, n = cfg->n;
    int16_t *grid = st->grid + b * G * G;
    uint8_t *vis = st->visited + b * G * G;
    const int8_t *act = io->actions + b * n;
    st->t[b] += 1;
    int cap = 0, moved = 0, pen = 0, lazy = 0, rev = 0;
    for (int i = 0; i < n; ++i) {
        int a = act[i];
        if (a < 0 || a > 4) return ORC_EBADACTION;
        if (a == 4) { ++lazy; continue; }                 /* :349-351 */
        ++moved;
        int r = st->apos[(b * n + i) * 2], c = st->apos[(b * n + i) * 2 + 1];
        int nr = r + DR[a], nc = c + DC[a];
        if (inb(cfg, nr, nc) && grid[nr * G + nc] == 0) { /* :354 */
            if (!vis[nr * G + nc]) { vis[nr * G + nc] = 1; ++cap; } else ++rev;
            grid[r * G + c] = 0;
            grid[nr * G + nc] = (int16_t)(i + 1);
            st->apos[(b * n + i) * 2] = (int8_t)nr; st->apos[(b * n + i) * 2 + 1] = (int8_t)nc;
        } else ++pen;                                     /* :373-375 */
    }
    double final_reward = 0;
    int dones = 0;
    st->total_capture[b] += cap;
    if (st->total_capture[b] == cfg->n_empty_cells) { final_reward = cfg->final_reward; dones = 1; }   /* :378-382 */
    if (st->t[b] >= cfg->T) { st->success[b] = dones ? 1 : 0; dones = 1; }                           /* :385-390 */
    /* :300-306, mean(x) = sum(x)/len(x) (:601-602) */
    double dn = (double)n;
    double rew = cfg->step_cost
        + cfg->capture_reward * ((double)cap / dn)
        + cfg->moving_cost * ((double)moved / dn)
        + cfg->penalty * ((double)pen / dn)
        + cfg->lazy_penalty * ((double)lazy / dn)
        + cfg->revisit_penalty * ((double)rev / dn)
        + final_reward;
    io->reward[b] = rew;
    int32_t *cn = io->counts + b * 6;
    cn[0] = cap; cn[1] = moved; cn[2] = pen; cn[3] = rev; cn[4] = lazy; cn[5] = 0;
    *env_done = dones;
    return ORC_OK;
}

/* step every env the way VecEnvExecutor.step does (vec_env_executor.py:19-45): env.step, time limit,
 * reset-on-done; the observation and comm state written are the post-reset ones when auto_reset. */
int orc_step(const orc_cfg *cfg, orc_state *st, const orc_io *io, int64_t B)
{
    if (cfg->n > 512) return ORC_ECONFIG;
    for (int64_t b = 0; b < B; ++b) {
        int env_done = 0;
        st->tick[b] += 1;                                 /* all draws of this step are keyed with the new tick */
        int rc = cfg->scenario == ORC_PP ? pp_step_env(cfg, st, b, io, &env_done) : co_step_env(cfg, st, b, io, &env_done);
        if (rc) return rc;
        int done = env_done;
        if (cfg->max_path_length > 0 && st->t[b] >= cfg->max_path_length) done = 1;   /* :33-35 */
        io->done[b] = (uint8_t)done;
        if (done && io->auto_reset) {
            rc = reset_env(cfg, st, b, io);
            if (rc) return rc;
            rc = emit_obs_comm(cfg, st, b, io, 1);
        } else {
            rc = emit_obs_comm(cfg, st, b, io, 0);
        }
        if (rc) return rc;
    }
    return ORC_OK;
}

/* inverse-CDF action sampling over the masked probabilities (stream specification for
 * Categorical.sample, comm_categorical_mlp_policy.py:109-110): sequential fp32 cumulative sum. */
int orc_sample_actions(const orc_cfg *cfg, const orc_state *st, int64_t B, const float *probs, const float *u_in, int8_t *actions)
{
    const int n = cfg->n;
    for (int64_t b = 0; b < B; ++b)
        for (int i = 0; i < n; ++i) {
            const float *p = probs + (b * n + i) * 5;
            float u;
            if (u_in) u = u_in[b * n + i];
            else {
                uint32_t blk[4];
                rng_block(cfg, b, st->tick[b], STREAM_ACT, st->episode[b], (uint32_t)(i >> 2), blk);
                u = u24(blk[i & 3]);
            }
            float c = 0.0f; int a = -1, last = 4;
            for (int k = 0; k < 5; ++k) { if (p[k] > 0.0f) last = k; }
            for (int k = 0; k < 5; ++k) { c += p[k]; if (u < c) { a = k; break; } }
            actions[b * n + i] = (int8_t)(a < 0 ? last : a);
        }
    return ORC_OK;
}
