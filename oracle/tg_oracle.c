/*
 * TEST INFRASTRUCTURE ONLY -- plain-C CPU restatement of the reference Treasure
 * Game dynamics (batched), used as the checker for the CUDA path and as the
 * "C port" leg of bench.py's cpu_baseline.  The product library never links,
 * loads or calls this file.
 *
 * Parity status: PINNED -- tests/test_oracle_golden.py replays the
 * reference-generated trajectories in tests/golden/ (tools/gen_golden.py, made
 * by running the unmodified reference) through this file with the recorded
 * uniform draws injected and compares every state field after every gym step;
 * tests/test_oracle_vs_reference.py does the same live where /root/reference
 * exists.
 *
 * The arithmetic follows the reference at its own granularity (pixel probes,
 * per-object loops, recursive trigger propagation); nothing here is shared with
 * the CUDA kernels, which use a different formulation (cell bit-tests, packed
 * state), so agreement between the two is meaningful.
 *
 * Citations: impl = gym_treasure_game/envs/_treasure_game_impl/_treasure_game_impl.py
 *            opts = .../_move_options.py   opt = .../_option.py   objs = .../_objects.py
 *            tg   = gym_treasure_game/envs/treasure_game.py
 *
 * Build: see oracle/Makefile  (gcc -O2 -fPIC -shared -ffp-contract=off; single-threaded, callers thread over batches)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define S 48            /* _scale.py:8-9 */
#define INCR 4          /* impl:46-47 */
#define HALFW 12        /* player_width // 2, impl:49 */
#define MAXOBJ 32
#define MAXTRIG 64
#define TICK_CAP 4096   /* per option; the reference would hang instead */

enum { K_DOOR = 0, K_HANDLE = 1, K_KEY = 2, K_BOLT = 3, K_GOLD = 4 };
enum { A_NOP = 0, A_UP, A_DOWN, A_LEFT, A_RIGHT, A_JUMP, A_INTERACT };   /* _actions.py:7-13 */

typedef struct { int32_t kind, cx, cy, flag; } tgo_obj;
typedef struct { int32_t k1, i1, v1, k2, i2, v2; } tgo_trig;

typedef struct {
    int cw, ch, nobj, ntrig;
    char tiles[64][64];
    tgo_obj obj[MAXOBJ];
    tgo_trig trig[MAXTRIG];
    int obs_dim, start_cx, start_cy;
    int n_of_kind[5];
    int idx_in_kind[MAXOBJ];
    int obj_of_kind[5][MAXOBJ];
} level_t;

typedef struct {
    int kind, cx, cy, x, y, val, pt;
    double angle, radius;
} obj_t;

typedef struct {
    const level_t *lv;
    char cells[64][64];          /* impl:204-216 is cell-granular (rows alias) */
    obj_t o[MAXOBJ];
    int px, py, facing, ticker;
    int bag[MAXOBJ], nbag;
    int total_actions;
    /* episode accounting (not in the reference; mirrors the CUDA product) */
    int64_t ep_return; int32_t ep_steps;
    int error;
    /* rng */
    uint32_t draws, d0;          /* d0: draw index when the current batch call began (Philox block numbering) */
    int64_t env_id;
} env_t;

typedef struct {
    const level_t *lv;
    int64_t n, first_id;
    uint64_t seed;
    int max_steps, auto_reset;
    env_t *e;
    double *tape; int64_t *tape_off; int has_tape;
    int64_t stats[8];
} batch_t;

/* ---------------------------------------------------------------- RNG ---- */
/* Philox4x32-10 (Salmon et al., SC'11), restated from the published round
 * function.  Blocks are numbered per batch call: draw j of a call that began at draw index d0 is word (j & 3)
 * of ctr = (d0, j >> 2, env_id_lo, env_id_hi), key = (seed_lo, seed_hi); u = w / 2^32 (the reference's
 * transforms are applied to u unchanged).  Every tgo_batch_* entry point that can draw sets d0 first. */
static void philox4x32_10(const uint32_t c[4], const uint32_t k[2], uint32_t out[4])
{
    uint32_t c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3], k0 = k[0], k1 = k[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void tgo_philox(const uint32_t c[4], const uint32_t k[2], uint32_t out[4]) { philox4x32_10(c, k, out); }

static double draw(const batch_t *b, env_t *e)
{
    uint32_t d = e->draws++;
    if (b->has_tape) {
        int64_t i = e - b->e;
        return b->tape[b->tape_off[i] + d];
    }
    uint32_t j = d - e->d0;                          /* draw j of this batch call */
    uint32_t c[4] = { e->d0, j >> 2, (uint32_t)e->env_id, (uint32_t)((uint64_t)e->env_id >> 32) };
    uint32_t k[2] = { (uint32_t)b->seed, (uint32_t)(b->seed >> 32) }, w[4];
    philox4x32_10(c, k, w);
    return w[j & 3] * (1.0 / 4294967296.0);          /* one 32-bit word per draw: u = w / 2^32 (exact in a double) */
}

static double uniform(const batch_t *b, env_t *e, double lo, double hi)
{   /* CPython random.uniform: a + (b-a) * random() -- two rounded operations */
    volatile double span = hi - lo;
    volatile double prod = span * draw(b, e);
    return lo + prod;
}

/* ------------------------------------------------------------- level ---- */
void *tgo_level_new(const char *tiles, int cw, int ch, const tgo_obj *objs, int nobj,
                    const tgo_trig *trigs, int ntrig)
{
    if (cw > 60 || ch > 60 || nobj > MAXOBJ || ntrig > MAXTRIG) return NULL;
    level_t *lv = calloc(1, sizeof *lv);
    lv->cw = cw; lv->ch = ch; lv->nobj = nobj; lv->ntrig = ntrig;
    for (int y = 0; y < ch; y++) memcpy(lv->tiles[y], tiles + (size_t)y * cw, cw);
    memcpy(lv->obj, objs, sizeof(tgo_obj) * nobj);
    memcpy(lv->trig, trigs, sizeof(tgo_trig) * ntrig);
    lv->obs_dim = 2;
    for (int i = 0; i < nobj; i++) {
        int k = objs[i].kind;
        lv->idx_in_kind[i] = lv->n_of_kind[k];
        lv->obj_of_kind[k][lv->n_of_kind[k]++] = i;
        /* impl:368-378 + objs get_state: handle 1, bolt 1, key 2, gold 2, door 0 */
        lv->obs_dim += (k == K_HANDLE || k == K_BOLT) ? 1 : (k == K_KEY || k == K_GOLD) ? 2 : 0;
    }
    lv->start_cx = lv->start_cy = -1;          /* impl:173-176 */
    for (int y = 0; y < ch && lv->start_cx < 0; y++)
        for (int x = 0; x < cw; x++)
            if (lv->tiles[y][x] != '/') { lv->start_cx = x; lv->start_cy = y; break; }
    return lv;
}
void tgo_level_free(void *lv) { free(lv); }
int tgo_obs_dim(const void *lv) { return ((const level_t *)lv)->obs_dim; }
int tgo_count(const void *lv, int kind) { return ((const level_t *)lv)->n_of_kind[kind]; }

/* -------------------------------------------------- tile predicates ---- */
static char tile(const env_t *e, int x, int y)            /* impl:218-225 */
{
    if (x >= e->lv->cw * S || x < 0) return '/';
    if (y >= e->lv->ch * S || y < 0) return '/';
    return e->cells[y / S][x / S];
}
static char tile_cell(const env_t *e, int xc, int yc)      /* impl:227-230 */
{   return tile(e, xc * S + S / 2, yc * S + S / 2); }

static int up_clear(const env_t *e)                        /* impl:232-238 */
{
    for (int dx = -INCR; dx <= INCR; dx += INCR)
        for (int dy = -INCR; dy < 0; dy++)
            if (tile(e, e->px + dx, e->py + dy) != ' ') return 0;
    return 1;
}
static int can_go_up(const env_t *e)                       /* impl:240-250 */
{
    static const int dys[3] = { -INCR, 0, S - INCR };
    if (e->py <= 1) return 0;
    for (int i = 0; i < 3; i++)
        for (int dx = -HALFW; dx <= HALFW; dx += 2 * HALFW)
            if (tile(e, e->px + dx, e->py + dys[i]) == 'L') return 1;
    return 0;
}
static int can_go_down(const env_t *e)                     /* impl:252-257 */
{
    for (int dy = 0; dy < S + INCR; dy++)
        for (int dx = -HALFW; dx <= HALFW; dx += 2 * HALFW)
            if (tile(e, e->px + dx, e->py + dy) == 'L') return 1;
    return 0;
}
static int side_free(const env_t *e, int x)                /* impl:259-281 */
{
    static const int dys[2] = { INCR, S - INCR };
    for (int i = 0; i < 2; i++) {
        char t = tile(e, x, e->py + dys[i]);
        if (t == '/' || t == 'D') return 0;
    }
    return 1;
}
static int can_go_left(const env_t *e)  { return side_free(e, e->px - HALFW - INCR); }
static int can_go_right(const env_t *e) { return side_free(e, e->px + HALFW + INCR); }
static int can_fall(const env_t *e)                        /* impl:283-288 */
{
    static const int dxs[2] = { -HALFW + 2, -2 + HALFW }, dys[2] = { 0, S + 2 };
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++)
            if (tile(e, e->px + dxs[i], e->py + dys[j]) != ' ') return 0;
    return 1;
}
static int near_obj(const env_t *e, const obj_t *o)        /* objs:46-53 at (px, py+24.0) */
{
    double cx = o->x + S / 2.0, cy = o->y + S / 2.0;
    double d = pow(e->px - cx, 2) + pow(e->py + S / 2.0 - cy, 2);
    return sqrt(d) < o->radius;
}
static void player_cell(const env_t *e, int *xc, int *yc)  /* impl:441-445 (floor division) */
{
    int x = e->px, y = e->py + S / 2;
    *xc = (x >= 0) ? x / S : -((-x + S - 1) / S);
    *yc = (y >= 0) ? y / S : -((-y + S - 1) / S);
}
static int bag_has(const env_t *e, int kind)               /* impl:418-428 */
{
    for (int i = 0; i < e->nbag; i++) if (e->o[e->bag[i]].kind == kind) return 1;
    return 0;
}

/* ------------------------------------------------------ trigger graph ---- */
static void set_val(const batch_t *b, env_t *e, int oi, int v);

static void wiggle(const batch_t *b, env_t *e, obj_t *h)   /* objs:127-131 */
{   h->angle = h->val ? uniform(b, e, 0.85, 1.0) : uniform(b, e, 0, 0.15); }

static void fire(const batch_t *b, env_t *e, int oi, int v)  /* objs:76-94 */
{
    const level_t *lv = e->lv;
    int kind = e->o[oi].kind, idx = lv->idx_in_kind[oi];
    e->o[oi].pt = 1;
    for (int t = 0; t < lv->ntrig; t++) {      /* registration order == file order, objs:65-71 */
        const tgo_trig *tr = &lv->trig[t];
        if (tr->k1 == kind && tr->i1 == idx && (tr->v1 != 0) == (v != 0)) {
            int tgt = lv->obj_of_kind[tr->k2][tr->i2];
            if (!e->o[tgt].pt) set_val(b, e, tgt, tr->v2 != 0);
        }
    }
    e->o[oi].pt = 0;
}
static void set_val(const batch_t *b, env_t *e, int oi, int v)  /* objs:145-149,175-178,231-235 */
{
    obj_t *o = &e->o[oi];
    if (o->val == v) return;
    o->val = v;
    if (o->kind == K_HANDLE) wiggle(b, e, o);
    else if (o->kind == K_DOOR) e->cells[o->cy][o->cx] = v ? 'D' : ' ';   /* objs:246-253 */
    fire(b, e, oi, v);
}
static void flip(const batch_t *b, env_t *e, int oi)       /* objs:117-122 */
{
    if (uniform(b, e, 0, 1) <= 0.8) set_val(b, e, oi, !e->o[oi].val);
    else wiggle(b, e, &e->o[oi]);
}
static void drop_key(env_t *e)                             /* impl:434-439 */
{
    for (int i = 0; i < e->nbag; i++) {
        obj_t *o = &e->o[e->bag[i]];
        if (o->kind == K_KEY) {
            memmove(&e->bag[i], &e->bag[i + 1], sizeof(int) * (e->nbag - i - 1));
            e->nbag--;
            o->cx = o->cy = -1; o->x = o->y = -S;          /* objs:34-38 */
            return;
        }
    }
}

/* --------------------------------------------------------- reset ---- */
static void env_reset(const batch_t *b, env_t *e)          /* impl:55-73 */
{
    const level_t *lv = e->lv;
    for (int y = 0; y < lv->ch; y++) memcpy(e->cells[y], lv->tiles[y], lv->cw);
    for (int i = 0; i < lv->nobj; i++) {                   /* impl:119-166, file order */
        obj_t *o = &e->o[i];
        o->kind = lv->obj[i].kind; o->cx = lv->obj[i].cx; o->cy = lv->obj[i].cy;
        o->x = o->cx * S; o->y = o->cy * S; o->val = lv->obj[i].flag != 0; o->pt = 0;
        o->radius = (o->kind == K_HANDLE) ? S * 0.75 : S / 2.0;            /* objs:23,115 */
        o->angle = 0.0;
        if (o->kind == K_HANDLE)                                           /* objs:111-114 */
            o->angle = o->val ? uniform(b, e, 0.85, 1.0) : uniform(b, e, 0, 0.15);
        else if (o->kind == K_DOOR)                                        /* objs:226 */
            e->cells[o->cy][o->cx] = o->val ? 'D' : ' ';
    }
    /* impl:168-178 with CPython random.gauss (fresh pair; second call uses gauss_next) */
    double x2pi = draw(b, e) * (2.0 * M_PI);
    double g2rad = sqrt(-2.0 * log(1.0 - draw(b, e)));
    double z0 = cos(x2pi) * g2rad, z1 = sin(x2pi) * g2rad;
    int nx = (int)(0 + z0 * (S / 24.0));
    int ny = (int)fabs(0 + z1 * (S / 36.0));
    e->px = e->py = 0;
    if (lv->start_cx >= 0) { e->px = lv->start_cx * S + S / 2 + nx; e->py = lv->start_cy * S + ny; }
    e->nbag = 0; e->ticker = 0; e->facing = 1; e->total_actions = 0;
    e->ep_return = 0; e->ep_steps = 0;
}

/* ------------------------------------------------- primitive tick ---- */
static int noisy(const batch_t *b, env_t *e, int val)      /* impl:361-366 */
{
    double mid = val / 2.0, r;
    if (val < mid) r = uniform(b, e, val, mid); else r = uniform(b, e, mid, val);
    return (int)nearbyint(r);                              /* round-half-even == Python round */
}

static int tick(const batch_t *b, env_t *e, int act)       /* impl:290-359 */
{
    const level_t *lv = e->lv;
    int xd = 0, yd = 0;
    e->total_actions++;
    switch (act) {
    case A_UP:    if (can_go_up(e)) yd = noisy(b, e, -INCR); break;
    case A_DOWN:  if (can_go_down(e)) yd = noisy(b, e, INCR); break;
    case A_LEFT:  if (can_go_left(e)) { xd = noisy(b, e, -INCR); e->facing = 0; } break;
    case A_RIGHT: if (can_go_right(e)) { xd = noisy(b, e, INCR); e->facing = 1; } break;
    case A_JUMP:
        if (!can_go_down(e) && up_clear(e)) { e->ticker = 22; if (draw(b, e) > 0.25) e->ticker = 23; }
        break;
    case A_INTERACT:
        for (int i = 0; i < lv->nobj; i++) {               /* impl:322-329 */
            if (!near_obj(e, &e->o[i])) continue;
            if (e->o[i].kind == K_HANDLE) flip(b, e, i);
            else if (e->o[i].kind == K_BOLT && bag_has(e, K_KEY)) { set_val(b, e, i, 0); drop_key(e); }
        }
        break;
    default: break;
    }
    if (e->ticker > 0) { if (up_clear(e)) yd = -INCR; e->ticker--; }       /* impl:331-337 */
    else if (can_fall(e)) { e->ticker = 0; yd = INCR; }
    e->px += xd;                                                           /* impl:339 */
    if (can_fall(e) && yd > 0) {                                           /* impl:341-348 */
        while (yd > 0) { e->py++; yd--; if (!can_fall(e)) yd = 0; }
    } else e->py += yd;
    for (int i = 0; i < lv->nobj; i++) {                                   /* impl:350-354 */
        obj_t *o = &e->o[i];
        if ((o->kind == K_KEY || o->kind == K_GOLD) && near_obj(e, o)) {
            o->cx = lv->cw - 1 - e->nbag; o->cy = lv->ch - 1;
            o->x = o->cx * S; o->y = o->cy * S;
            if (e->nbag < MAXOBJ) e->bag[e->nbag++] = i;
        }
    }
    return act == A_JUMP ? -5 : -1;                                        /* impl:15-16,356-359 */
}

/* ------------------------------------------------------ option layer ---- */
static int is_object_at(const env_t *e, int xc, int yc)    /* impl:402-409 */
{
    for (int i = 0; i < e->lv->nobj; i++) {
        const obj_t *o = &e->o[i];
        if (o->cx == xc && o->cy == yc && (o->kind != K_DOOR || o->val)) return 1;
    }
    return 0;
}
static int is_closed_door_at(const env_t *e, int xc, int yc)   /* impl:411-416 */
{
    for (int i = 0; i < e->lv->nobj; i++) {
        const obj_t *o = &e->o[i];
        if (o->kind == K_DOOR && o->val && o->cx == xc && o->cy == yc) return 1;
    }
    return 0;
}
/* opts:43-67 (s=-1) / opts:115-139 (s=+1); returns 0 for None */
static int walk_target(const env_t *e, int pcx, int pcy, int s, int *tx)
{
    int xc = pcx + s, yc = pcy;
    for (;;) {
        if (tile_cell(e, xc, yc - 1) == 'L' || tile_cell(e, xc, yc + 1) == 'L'
            || tile_cell(e, xc + s, yc) == '/' || is_object_at(e, xc, yc)
            || is_closed_door_at(e, xc + s, yc) || tile_cell(e, xc + s, yc + 1) == ' ') break;
        xc += s;
        if (xc < 0) return 0;
    }
    *tx = xc;
    return 1;
}
static int walk_can_run(const env_t *e, int s)             /* opts:23-41 / opts:95-113 */
{
    int pcx, pcy, tx;
    player_cell(e, &pcx, &pcy);
    if (!walk_target(e, pcx, pcy, s, &tx)) return 0;
    for (int xc = pcx; s < 0 ? xc >= tx : xc <= tx; xc += s) {
        if (tile_cell(e, xc, pcy) != ' ') return 0;
        if (tile_cell(e, xc, pcy + 1) == ' ') return 0;
    }
    return 1;
}
static int drop_target(const env_t *e, int pcx, int pcy, int s, int *tx)   /* opts:211-221 / 406-416 */
{
    int xc = pcx + s, yc = pcy + 1;
    while (tile_cell(e, xc, yc) == ' ') { yc++; if (yc >= e->lv->ch) return 0; }
    *tx = xc;
    return 1;
}
static int landing(const env_t *e, int xc, int yc)         /* opts:281-287 / 351-357 */
{   return tile_cell(e, xc, yc) == ' ' && tile_cell(e, xc, yc + 1) == '/'; }
static int jump_target(const env_t *e, int pcx, int pcy, int s, int *tx)   /* opts:269-279 / 339-349 */
{
    if (landing(e, pcx + s, pcy - 1)) { *tx = pcx + s; return 1; }
    if (landing(e, pcx + 2 * s, pcy - 1)) { *tx = pcx + 2 * s; return 1; }
    return 0;
}
static int aligned(const env_t *e, int tx)                 /* close_enough_*: |48 tx + 24 - px| < 4 */
{   return fabs((tx * S + S / 2.0) - e->px) < INCR; }

static int can_run(const env_t *e, int k)
{
    int xc, yc, s;
    switch (k) {
    case 0: return walk_can_run(e, -1);
    case 1: return walk_can_run(e, +1);
    case 2: return can_go_up(e);                           /* opts:165-166 */
    case 3: return can_go_down(e);                         /* opts:181-182 */
    case 4:                                                /* opts:446-455 */
        for (int i = 0; i < e->lv->nobj; i++)
            if (near_obj(e, &e->o[i])) {
                if (e->o[i].kind == K_HANDLE) return 1;
                if (e->o[i].kind == K_BOLT && bag_has(e, K_KEY)) return 1;
            }
        return 0;
    case 5: case 6:                                        /* opts:199-209 / 394-404 */
        player_cell(e, &xc, &yc); s = (k == 5) ? -1 : 1;
        return tile_cell(e, xc + s, yc) == ' ' && tile_cell(e, xc + s, yc + 1) == ' ';
    case 7: case 8:                                        /* opts:254-267 / 324-337 */
        player_cell(e, &xc, &yc); s = (k == 7) ? -1 : 1;
        if (tile_cell(e, xc, yc - 1) != ' ' || tile_cell(e, xc + s, yc - 1) != ' ') return 0;
        return landing(e, xc + s, yc - 1) || landing(e, xc + 2 * s, yc - 1);
    }
    return 0;
}

/* opt:20-36 with the nine policies of opts.  *ran=0 reproduces `return None`. */
static int run_option(const batch_t *b, env_t *e, int k, int *ran, int *nticks)
{
    *ran = 0; *nticks = 0;
    if (k < 0 || k > 8 || !can_run(e, k)) return 0;
    int tot = 0, done = 0, first = 1, tx = 0, have = 1, pcx, pcy, n = 0;
    int s = (k == 0 || k == 5 || k == 7) ? -1 : 1;
    player_cell(e, &pcx, &pcy);
    if (k <= 1) have = walk_target(e, pcx, pcy, s, &tx);
    else if (k == 5 || k == 6) have = drop_target(e, pcx, pcy, s, &tx);
    else if (k >= 7) have = jump_target(e, pcx, pcy, s, &tx);
    if (!have) { e->error = 1; return 0; }                 /* reference: TypeError in close_enough_x(None) */
    *ran = 1;
    while (!done) {
        int act;
        if (k <= 1) {                                      /* opts:74-85 / 146-157 */
            if (aligned(e, tx)) done = 1;
            act = (k == 0) ? A_LEFT : A_RIGHT;
        } else if (k == 2) {                               /* opts:168-173 */
            if (!can_go_up(e)) { done = 1; act = A_NOP; } else act = A_UP;
        } else if (k == 3) {                               /* opts:184-189 */
            if (!can_go_down(e)) { done = 1; act = A_NOP; } else act = A_DOWN;
        } else if (k == 4) {                               /* opts:457-460 */
            done = 1; act = A_INTERACT;
        } else if (k <= 6) {                               /* opts:231-244 / 426-439 */
            if (aligned(e, tx)) { if (!can_fall(e)) done = 1; act = A_NOP; }
            else act = (k == 5) ? A_LEFT : A_RIGHT;
        } else {                                           /* opts:297-314 / 367-384 */
            if (first) act = A_JUMP;
            else if (aligned(e, tx)) { if (!can_fall(e)) done = 1; act = A_NOP; }
            else {
                int blocked = !(s < 0 ? can_go_left(e) : can_go_right(e));
                if (!can_fall(e) && blocked) act = (s < 0) ? A_RIGHT : A_LEFT;
                else act = (s < 0) ? A_LEFT : A_RIGHT;
            }
        }
        first = 0;
        tot += tick(b, e, act);
        if (++n >= TICK_CAP && !done) { e->error = 1; break; }
    }
    *nticks = n;
    return tot;
}

/* ------------------------------------------------------- observation ---- */
static void write_obs(const env_t *e, double *out)         /* impl:368-378 + objs get_state */
{
    const level_t *lv = e->lv;
    double W = lv->cw * S, H = lv->ch * S;
    int k = 0;
    out[k++] = (double)e->px / W; out[k++] = (double)e->py / H;
    for (int i = 0; i < lv->nobj; i++) {
        const obj_t *o = &e->o[i];
        if (o->kind == K_HANDLE) out[k++] = o->angle;
        else if (o->kind == K_BOLT) out[k++] = o->val ? 1.0 : 0.0;
        else if (o->kind == K_KEY || o->kind == K_GOLD) { out[k++] = (double)o->x / W; out[k++] = (double)o->y / H; }
    }
}
static int is_done(const env_t *e)                         /* tg:95 */
{
    int xc, yc; player_cell(e, &xc, &yc);
    return bag_has(e, K_GOLD) && yc == 0;
}

/* ----------------------------------------------------- save / restore ---- */
/* impl:447-481 init_with_state, quirks included (see py_oracle.OracleEnv.init_with_state):
 * every key / gold / bolt reads the FIRST slot carrying its name (desc.index), -99 keeps the current
 * value through a float round trip, handles keep previously_triggered = True afterwards. */
static void env_init_with_state(const batch_t *b, env_t *e, const double *in)
{
    const level_t *lv = e->lv;
    double st[2 + 2 * MAXOBJ], cur[2 + 2 * MAXOBJ];
    int od = lv->obs_dim;
    int first_key = -1, first_gold = -1, first_bolt = -1, k = 2;
    double W = lv->cw * S, H = lv->ch * S;
    write_obs(e, cur);
    for (int i = 0; i < lv->nobj; i++) {           /* slot of the first object of each name (impl:380-400) */
        int kind = lv->obj[i].kind;
        if (kind == K_HANDLE) k += 1;
        else if (kind == K_BOLT) { if (first_bolt < 0) first_bolt = k; k += 1; }
        else if (kind == K_KEY) { if (first_key < 0) first_key = k; k += 2; }
        else if (kind == K_GOLD) { if (first_gold < 0) first_gold = k; k += 2; }
    }
    for (int v = 0; v < od; v++) st[v] = (in[v] == -99) ? cur[v] : in[v];
    e->facing = 1;
    e->px = (int)(st[0] * W); e->py = (int)(st[1] * H);
    k = 2;
    for (int i = 0; i < lv->nobj; i++) {
        obj_t *o = &e->o[i];
        if (o->kind == K_KEY || o->kind == K_GOLD) {
            int slot = (o->kind == K_KEY) ? first_key : first_gold;
            o->x = (int)(st[slot] * W); o->y = (int)(st[slot + 1] * H);        /* objs:40-44 move_to_xy */
            o->cx = (int)(o->x / (double)S); o->cy = (int)(o->y / (double)S);
            k += 2;
        } else if (o->kind == K_HANDLE) {
            int old = o->val;                                                  /* objs:133-143 set_angle */
            o->angle = st[k];
            o->val = !(o->angle <= 0.15);
            if (o->val != old) fire(b, e, i, o->val);
            o->pt = 1;                                                         /* impl:473 */
            k += 1;
        } else if (o->kind == K_BOLT) {
            set_val(b, e, i, st[first_bolt] > 0.5);
            k += 1;
        }
    }
}

/* --------------------------------------------------------------- batch ---- */
void *tgo_batch_new(const void *level, int64_t n, int64_t first_env_id, uint64_t seed,
                    int max_episode_steps, int auto_reset)
{
    batch_t *b = calloc(1, sizeof *b);
    b->lv = level; b->n = n; b->first_id = first_env_id; b->seed = seed;
    b->max_steps = max_episode_steps; b->auto_reset = auto_reset;
    b->e = calloc((size_t)n, sizeof(env_t));
    for (int64_t i = 0; i < n; i++) { b->e[i].lv = b->lv; b->e[i].env_id = first_env_id + i; }
    return b;
}
void tgo_batch_free(void *bp)
{   batch_t *b = bp; free(b->e); free(b->tape); free(b->tape_off); free(b); }

void tgo_batch_set_tape(void *bp, const double *tape, const int64_t *offsets)
{
    batch_t *b = bp;
    free(b->tape); free(b->tape_off); b->tape = NULL; b->tape_off = NULL; b->has_tape = 0;
    if (!tape) return;
    int64_t tot = offsets[b->n];
    b->tape = malloc(sizeof(double) * (size_t)(tot ? tot : 1));
    b->tape_off = malloc(sizeof(int64_t) * (size_t)(b->n + 1));
    memcpy(b->tape, tape, sizeof(double) * (size_t)tot);
    memcpy(b->tape_off, offsets, sizeof(int64_t) * (size_t)(b->n + 1));
    b->has_tape = 1;
}

void tgo_batch_reset(void *bp, const uint8_t *mask, double *obs)
{
    batch_t *b = bp; int od = b->lv->obs_dim;
    for (int64_t i = 0; i < b->n; i++) {
        b->e[i].d0 = b->e[i].draws;
        if (!mask || mask[i]) { b->e[i].error = 0; env_reset(b, &b->e[i]); }
        if (obs) write_obs(&b->e[i], obs + i * od);
    }
}

/* done: bit0 terminated (tg:95), bit1 truncated (episode step limit) */
void tgo_batch_step(void *bp, const int32_t *actions, double *obs, float *reward, uint8_t *done,
                    uint8_t *ran_out, int32_t *ticks_out)
{
    batch_t *b = bp; int od = b->lv->obs_dim;
    int64_t s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0, s7 = 0;
    for (int64_t i = 0; i < b->n; i++) {
        env_t *e = &b->e[i];
        int ran, nt, err0 = e->error;
        e->d0 = e->draws;
        int r = run_option(b, e, actions[i], &ran, &nt);
        e->ep_return += r; e->ep_steps += 1;
        int term = is_done(e);
        int trunc = (b->max_steps > 0 && e->ep_steps >= b->max_steps);
        int d = term | (trunc << 1);
        s4 += nt; s5 += ran; s7 += (e->error && !err0);
        if (d) { s0 += 1; s1 += term; s2 += e->ep_return; s3 += e->ep_steps; }
        if (d && b->auto_reset) env_reset(b, e);
        if (obs) write_obs(e, obs + i * od);
        if (reward) reward[i] = (float)r;
        if (done) done[i] = (uint8_t)d;
        if (ran_out) ran_out[i] = (uint8_t)ran;
        if (ticks_out) ticks_out[i] = nt;
    }
    b->stats[0] += s0; b->stats[1] += s1; b->stats[2] += s2; b->stats[3] += s3;
    b->stats[4] += s4; b->stats[5] += s5; b->stats[6] += b->n; b->stats[7] += s7;
}

/* one primitive action per env (impl:290-359), same accounting as tgo_batch_step */
void tgo_batch_prim_step(void *bp, const int32_t *actions, double *obs, float *reward, uint8_t *done)
{
    batch_t *b = bp; int od = b->lv->obs_dim;
    for (int64_t i = 0; i < b->n; i++) {
        env_t *e = &b->e[i];
        int err0 = e->error;
        e->d0 = e->draws;
        int r = tick(b, e, actions[i]);
        e->ep_return += r; e->ep_steps += 1;
        int term = is_done(e);
        int trunc = (b->max_steps > 0 && e->ep_steps >= b->max_steps);
        int d = term | (trunc << 1);
        b->stats[4] += 1; b->stats[5] += 1; b->stats[6] += 1; b->stats[7] += (e->error && !err0);
        if (d) { b->stats[0] += 1; b->stats[1] += term; b->stats[2] += e->ep_return; b->stats[3] += e->ep_steps; }
        if (d && b->auto_reset) env_reset(b, e);
        if (obs) write_obs(e, obs + i * od);
        if (reward) reward[i] = (float)r;
        if (done) done[i] = (uint8_t)d;
    }
}

void tgo_batch_init_with_state(void *bp, const double *states, const uint8_t *mask)
{
    batch_t *b = bp; int od = b->lv->obs_dim;
    for (int64_t i = 0; i < b->n; i++)
        if (!mask || mask[i]) { b->e[i].d0 = b->e[i].draws; env_init_with_state(b, &b->e[i], states + i * od); }
}

/* previously_triggered flag of every handle, [N][n_handles] */
void tgo_batch_get_pt(void *bp, uint8_t *pt)
{
    batch_t *b = bp; const level_t *lv = b->lv; int nh = lv->n_of_kind[K_HANDLE];
    for (int64_t i = 0; i < b->n; i++) {
        int h = 0;
        for (int j = 0; j < lv->nobj; j++) if (b->e[i].o[j].kind == K_HANDLE) pt[i * nh + h++] = (uint8_t)b->e[i].o[j].pt;
    }
}

void tgo_batch_mask(void *bp, uint8_t *mask)               /* tg:83-89 */
{
    batch_t *b = bp;
    for (int64_t i = 0; i < b->n; i++)
        for (int k = 0; k < 9; k++) mask[i * 9 + k] = (uint8_t)can_run(&b->e[i], k);
}

void tgo_batch_stats(void *bp, int64_t *out8) { memcpy(out8, ((batch_t *)bp)->stats, 64); }

/* Test hook: the per-env draw index (the 32-bit counter the Philox blocks are numbered from), e.g. to start a
 * batch just below 2^32 and compare the wrap with the CUDA library ("RNG"). */
void tgo_batch_set_draws(void *bp, const uint32_t *draws)
{
    batch_t *b = bp;
    for (int64_t i = 0; i < b->n; i++) { b->e[i].draws = draws[i]; b->e[i].d0 = draws[i]; }
}

/* Flat state dump for differential tests.  Any pointer may be NULL.
 * pos[N*2]=px,py; misc[N*4]=facing,ticker,total_actions,draws; doors/handles/bolts[N*count] 0/1;
 * angles[N*nh]; items[N*ni*4]=x,y,cx,cy; bag[N*4] item index per bag slot or -1 (an item can appear twice);
 * acct[N*3]=ep_return,ep_steps,error */
void tgo_batch_get(void *bp, int32_t *pos, int32_t *misc, uint8_t *doors, uint8_t *handles,
                   uint8_t *bolts, double *angles, int32_t *items, int32_t *bag, int64_t *acct)
{
    batch_t *b = bp; const level_t *lv = b->lv;
    int nd = lv->n_of_kind[K_DOOR], nh = lv->n_of_kind[K_HANDLE], nb = lv->n_of_kind[K_BOLT];
    int ni = lv->n_of_kind[K_KEY] + lv->n_of_kind[K_GOLD];
    for (int64_t i = 0; i < b->n; i++) {
        const env_t *e = &b->e[i];
        if (pos) { pos[i * 2] = e->px; pos[i * 2 + 1] = e->py; }
        if (misc) { misc[i * 4] = e->facing; misc[i * 4 + 1] = e->ticker; misc[i * 4 + 2] = e->total_actions; misc[i * 4 + 3] = (int32_t)e->draws; }
        int item_no[MAXOBJ], c = 0, d = 0, h = 0, bo = 0;
        for (int j = 0; j < lv->nobj; j++) {
            const obj_t *o = &e->o[j];
            item_no[j] = -1;
            if (o->kind == K_DOOR) { if (doors) doors[i * nd + d] = (uint8_t)o->val; d++; }
            else if (o->kind == K_HANDLE) { if (handles) handles[i * nh + h] = (uint8_t)o->val; if (angles) angles[i * nh + h] = o->angle; h++; }
            else if (o->kind == K_BOLT) { if (bolts) bolts[i * nb + bo] = (uint8_t)o->val; bo++; }
            else { item_no[j] = c; if (items) { int32_t *p = items + (i * ni + c) * 4; p[0] = o->x; p[1] = o->y; p[2] = o->cx; p[3] = o->cy; } c++; }
        }
        if (bag) for (int j = 0; j < 4; j++) bag[i * 4 + j] = (j < e->nbag) ? item_no[e->bag[j]] : -1;
        if (acct) { acct[i * 3] = e->ep_return; acct[i * 3 + 1] = e->ep_steps; acct[i * 3 + 2] = e->error; }
    }
}
