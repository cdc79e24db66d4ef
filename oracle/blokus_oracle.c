/*
 * blokus_oracle.c -- CPU restatement of the Blokus env hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product path (blokus_rl_b200) never does.
 *
 * PARITY PINNED ONLY ON THE REFERENCE'S RENDERED GAMES; everything those renders do not show is
 * UNPINNED.  The reference's env arithmetic lives in two un-vendored, un-pinned git
 * dependencies (colosseumrl: /root/reference/setup.py:11; blokus-gym: setup.py:33) that are
 * absent from /root/reference and from this image, and the reference has no tests
 * (.github/workflows/analysis.yml:10-57 is lint only).  The only outputs of the real engine in
 * /root/reference are the renders under docs/images: three complete games (80 transitions), two
 * positions and one observation; tests/golden/make_ref_render_golden.py decodes them and
 * tests/test_ref_render_golden.py replays them here (legality of every reference move, side to
 * move incl. auto-skips, game end at the last frame and not before, winners vs the file names,
 * start corners, observation planes).  Not shown by the renders and therefore still unpinned:
 * action-id order / action strings, score bonuses, observation orientation for movers 1..3.
 * This file restates the *contract* visible at the reference's call sites plus standard Blokus
 * rules (rules R1-R13):
 *   new_state      blokus_rl/colossumrl/blokus_wrapper.py:80-87     -> orc_reset
 *   next_state     blokus_wrapper.py:89-106                          -> orc_step
 *   valid_actions  blokus_wrapper.py:108-132, 233-246                -> orc_legal_mask
 *   get_winners    blokus_wrapper.py:164-186                         -> orc_winners / orc_terminal_values
 *   canonical_board  blokus_wrapper.py:134-146, models/blokus_nnet.py:99 -> orc_observe
 *   board_contents blokus_wrapper.py:208-218, 259-266                -> orc_board_contents
 *   action table   blokus_wrapper.py:281-324                         -> orc_create / orc_action_cells
 * The executable pins that do exist are checked in tests/test_oracle_kat.py:
 *   30,433 actions on 20x20 (blokus_nnet.py:17,97), observation (8,20,20) (blokus_nnet.py:99),
 *   terminal vector 3/1/-1 (blokus_wrapper.py:177-185), colours 0..4 (blokus_wrapper.py:259).
 *
 * Two independent legality implementations live here and are cross-checked by the tests:
 *   - naive, cell by cell on the uint8 board (orc_legal_mask)                   -- obviously right
 *   - bit-parallel on incrementally kept row words (orc_fast_legal_mask)        -- the timed CPU baseline
 * A third, pure-Python set-based one is oracle/naive.py.
 *
 * Every configuration lives in an orc_ctx (no globals): contexts are immutable after orc_create, so any
 * number of threads may play different states through one context at once.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXN 20
#define MAXP 4
#define NPIECES 21
#define MAXORI 91
#define MAXCELLS 5
#define MAXFIELDS (MAXORI * MAXN)
#define ROWPAD 8                      /* zero rows after the board: anchor rows may read up to 4 past the end */

typedef struct {
    int32_t N, P;
    uint8_t board[MAXN * MAXN];   /* 0 empty, 1..P = colour of player 0..P-1 (R1); what the naive rules read */
    uint32_t inv[MAXP];           /* bit i set = piece i still in hand (R4) */
    int16_t score[MAXP];          /* squares placed (R10 default) */
    uint8_t lastmono[MAXP];       /* most recent placement was the monomino */
    uint8_t mover;                /* player to move */
    uint8_t done;
    uint16_t ply;                 /* placements made in this game */
    uint32_t game;                /* games finished before this one (auto-reset counter) */
    uint32_t rows[MAXP][MAXN];    /* the same board as row words (bit x = column x); what the bit-parallel rules read */
} orc_state;

typedef struct {
    int piece, local, h, w, n;
    int8_t dy[MAXCELLS], dx[MAXCELLS];
} orient_t;

typedef struct orc_ctx {
    int N, P, rule;
    int nori, nact, nfields;
    orient_t ori[MAXORI];
    int psize[NPIECES];
    int obase[MAXORI + 1];        /* first action id of each orientation */
    int fbase[MAXORI + 1];        /* first field (orientation, anchor row) of each orientation */
    int foff[MAXFIELDS + 1];      /* first action id of each field */
    int16_t *act_o, *act_y, *act_x;
} orc_ctx;

/* Shapes, rows top->bottom, SURVEY.md Appendix B order. */
static const char *SHAPES[NPIECES] = {
    "#", "#/#", "#/#/#", "##/#.", "#/#/#/#", "##/#./#.", "#./##/#.", "##/##", "#./##/.#",
    "#/#/#/#/#", "##/#./#./#.", "#./##/#./#.", "##/##/#.", "##/#./##", "###/#../#..",
    "#../###/#..", "#./#./##/.#", "#../###/.#.", "#../##./.##", "#../###/..#", ".#./###/.#."};

/* ---- orientation enumeration (independent of blokus_rl_b200/tables.py) ---- */
typedef struct { int n; int y[MAXCELLS], x[MAXCELLS]; } shape_t;

static void shape_norm(shape_t *s) {
    int my = 99, mx = 99, i, j;
    for (i = 0; i < s->n; i++) { if (s->y[i] < my) my = s->y[i]; if (s->x[i] < mx) mx = s->x[i]; }
    for (i = 0; i < s->n; i++) { s->y[i] -= my; s->x[i] -= mx; }
    for (i = 0; i < s->n; i++)           /* sort cells by (y, x) */
        for (j = i + 1; j < s->n; j++)
            if (s->y[j] < s->y[i] || (s->y[j] == s->y[i] && s->x[j] < s->x[i])) {
                int t = s->y[i]; s->y[i] = s->y[j]; s->y[j] = t;
                t = s->x[i]; s->x[i] = s->x[j]; s->x[j] = t;
            }
}
static int shape_cmp(const shape_t *a, const shape_t *b) {
    for (int i = 0; i < a->n; i++) {
        if (a->y[i] != b->y[i]) return a->y[i] < b->y[i] ? -1 : 1;
        if (a->x[i] != b->x[i]) return a->x[i] < b->x[i] ? -1 : 1;
    }
    return 0;
}

static void build_orientations(orc_ctx *c) {
    c->nori = 0;
    for (int p = 0; p < NPIECES; p++) {
        shape_t base; base.n = 0;
        int y = 0, x = 0;
        for (const char *ch = SHAPES[p]; *ch; ch++) {
            if (*ch == '/') { y++; x = 0; continue; }
            if (*ch == '#') { base.y[base.n] = y; base.x[base.n] = x; base.n++; }
            x++;
        }
        c->psize[p] = base.n;
        shape_t list[8]; int nl = 0;
        shape_t cur = base;
        for (int r = 0; r < 4; r++) {
            shape_t rot = cur;                       /* rotate: (y, x) -> (x, -y) */
            for (int i = 0; i < cur.n; i++) { rot.y[i] = cur.x[i]; rot.x[i] = -cur.y[i]; }
            cur = rot;
            for (int m = 0; m < 2; m++) {
                shape_t s = cur;
                if (m) for (int i = 0; i < s.n; i++) s.x[i] = -s.x[i];
                shape_norm(&s);
                int dup = 0;
                for (int k = 0; k < nl; k++) if (shape_cmp(&list[k], &s) == 0) dup = 1;
                if (!dup) list[nl++] = s;
            }
        }
        for (int i = 0; i < nl; i++)                 /* sort orientations lexicographically */
            for (int j = i + 1; j < nl; j++)
                if (shape_cmp(&list[j], &list[i]) < 0) { shape_t t = list[i]; list[i] = list[j]; list[j] = t; }
        for (int i = 0; i < nl; i++) {
            orient_t *o = &c->ori[c->nori++];
            o->piece = p; o->local = i; o->n = list[i].n; o->h = 0; o->w = 0;
            for (int k = 0; k < o->n; k++) {
                o->dy[k] = (int8_t)list[i].y[k]; o->dx[k] = (int8_t)list[i].x[k];
                if (list[i].y[k] + 1 > o->h) o->h = list[i].y[k] + 1;
                if (list[i].x[k] + 1 > o->w) o->w = list[i].x[k] + 1;
            }
        }
    }
}

orc_ctx *orc_create(int N, int P, int score_rule) {
    if (N < 5 || N > MAXN || (P != 2 && P != 4)) return 0;
    orc_ctx *c = calloc(1, sizeof(orc_ctx));
    if (!c) return 0;
    c->N = N; c->P = P; c->rule = score_rule;
    build_orientations(c);
    int cap = 0;
    for (int o = 0; o < c->nori; o++) cap += (N - c->ori[o].h + 1) * (N - c->ori[o].w + 1);
    c->act_o = malloc(sizeof(int16_t) * cap); c->act_y = malloc(sizeof(int16_t) * cap);
    c->act_x = malloc(sizeof(int16_t) * cap);
    c->nact = 0; c->nfields = 0;
    for (int o = 0; o < c->nori; o++) {               /* id order: (piece, orientation, ay, ax) */
        c->obase[o] = c->nact;
        c->fbase[o] = c->nfields;
        for (int y = 0; y + c->ori[o].h <= N; y++) {
            c->foff[c->nfields++] = c->nact;
            for (int x = 0; x + c->ori[o].w <= N; x++) {
                c->act_o[c->nact] = (int16_t)o; c->act_y[c->nact] = (int16_t)y; c->act_x[c->nact] = (int16_t)x;
                c->nact++;
            }
        }
    }
    c->obase[c->nori] = c->nact;
    c->fbase[c->nori] = c->nfields;
    c->foff[c->nfields] = c->nact;
    return c;
}
void orc_destroy(orc_ctx *c) {
    if (!c) return;
    free(c->act_o); free(c->act_y); free(c->act_x); free(c);
}

int orc_num_actions(const orc_ctx *c) { return c->nact; }
int orc_num_orientations(const orc_ctx *c) { return c->nori; }
int orc_num_fields(const orc_ctx *c) { return c->nfields; }
int orc_state_size(void) { return (int)sizeof(orc_state); }
int orc_piece_size(const orc_ctx *c, int p) { return c->psize[p]; }

/* cells_yx: up to 10 bytes (y0,x0,y1,x1,...). returns ncells; meta = {piece, orient, ay, ax} */
int orc_action_cells(const orc_ctx *c, int a, uint8_t *cells_yx, int32_t *meta) {
    if (a < 0 || a >= c->nact) return -1;
    const orient_t *o = &c->ori[c->act_o[a]];
    for (int k = 0; k < o->n; k++) {
        cells_yx[2 * k] = (uint8_t)(c->act_y[a] + o->dy[k]);
        cells_yx[2 * k + 1] = (uint8_t)(c->act_x[a] + o->dx[k]);
    }
    if (meta) { meta[0] = o->piece; meta[1] = c->act_o[a]; meta[2] = c->act_y[a]; meta[3] = c->act_x[a]; }
    return o->n;
}

/* R3: start corners. */
static void start_corner(const orc_ctx *c, int p, int *y, int *x) {
    int n = c->N - 1;
    if (c->P == 2) { *y = p ? n : 0; *x = p ? n : 0; return; }
    *y = (p & 2) ? n : 0; *x = (p & 1) ? n : 0;
}

void orc_reset(const orc_ctx *c, orc_state *s, uint32_t game) {
    memset(s, 0, sizeof(*s));
    s->N = c->N; s->P = c->P; s->game = game;
    for (int p = 0; p < c->P; p++) s->inv[p] = (1u << NPIECES) - 1;
}

/* ---- naive legality: R5, R6 ---- */
static inline int cell(const orc_ctx *c, const orc_state *s, int y, int x) {
    if (y < 0 || x < 0 || y >= c->N || x >= c->N) return 0;
    return s->board[y * c->N + x];
}

static int placement_legal(const orc_ctx *c, const orc_state *s, int p, int a) {
    const orient_t *o = &c->ori[c->act_o[a]];
    if (!((s->inv[p] >> o->piece) & 1)) return 0;
    int col = p + 1, first = (s->inv[p] == (1u << NPIECES) - 1);
    int cy, cx, corner = 0;
    start_corner(c, p, &cy, &cx);
    for (int k = 0; k < o->n; k++) {
        int y = c->act_y[a] + o->dy[k], x = c->act_x[a] + o->dx[k];
        if (cell(c, s, y, x) != 0) return 0;                                  /* overlap */
        if (cell(c, s, y - 1, x) == col || cell(c, s, y + 1, x) == col ||
            cell(c, s, y, x - 1) == col || cell(c, s, y, x + 1) == col) return 0; /* own edge contact */
        if (first) { if (y == cy && x == cx) corner = 1; }
        else if (cell(c, s, y - 1, x - 1) == col || cell(c, s, y - 1, x + 1) == col ||
                 cell(c, s, y + 1, x - 1) == col || cell(c, s, y + 1, x + 1) == col) corner = 1;
    }
    return corner;
}

/* mask: nact bytes 0/1. returns number of legal actions for player p (any player, not only the mover) */
int orc_legal_mask(const orc_ctx *c, const orc_state *s, int p, uint8_t *mask) {
    int cnt = 0;
    for (int a = 0; a < c->nact; a++) {
        int l = s->done ? 0 : placement_legal(c, s, p, a);
        if (mask) mask[a] = (uint8_t)l;
        cnt += l;
    }
    return cnt;
}

static int has_move_naive(const orc_ctx *c, const orc_state *s, int p) {
    for (int a = 0; a < c->nact; a++) if (placement_legal(c, s, p, a)) return 1;
    return 0;
}

/* ---- bit-parallel legality (second, independent formulation; also the timed CPU engine) ----
 * fr[y] = cells of row y that are empty and not edge-adjacent to p's colour; dg[y] = cells diagonal to p's colour
 * (or the start corner on the first move).  frs[dx][y] = fr[y] >> dx, so a footprint cell (dy, dx) at anchor row ay is
 * frs[dx][ay + dy] and one (orientation, anchor row) FIELD -- bit x = anchor column x is legal -- is the AND of its
 * cells' free words and the OR of their diagonal words.  The loops over anchor rows are written so that the compiler
 * vectorises them (AVX2 / AVX-512 clones are picked at load time). */
typedef struct { uint32_t frs[MAXCELLS][MAXN + ROWPAD], dgs[MAXCELLS][MAXN + ROWPAD]; } rows_t;

static void build_rows(const orc_ctx *c, const orc_state *s, int p, rows_t *r) {
    const int N = c->N;
    const uint32_t full = (1u << N) - 1;
    uint32_t occ[MAXN + 2], own[MAXN + 2];
    occ[0] = own[0] = occ[N + 1] = own[N + 1] = 0;
    for (int y = 0; y < N; y++) {
        uint32_t o = 0;
        for (int q = 0; q < c->P; q++) o |= s->rows[q][y];
        occ[y + 1] = o; own[y + 1] = s->rows[p][y];
    }
    int first = (s->inv[p] == (1u << NPIECES) - 1), cy, cx;
    start_corner(c, p, &cy, &cx);
    memset(r, 0, sizeof(*r));
    for (int y = 0; y < N; y++) {
        uint32_t o = own[y + 1], ud = own[y] | own[y + 2];
        uint32_t adj = ud | (o << 1) | (o >> 1);
        uint32_t fr = ~(occ[y + 1] | adj) & full;
        uint32_t dg = first ? ((y == cy) ? (1u << cx) : 0) : (((ud << 1) | (ud >> 1)) & full);
        for (int dx = 0; dx < MAXCELLS; dx++) { r->frs[dx][y] = fr >> dx; r->dgs[dx][y] = dg >> dx; }
    }
}

/* all fields of player p in id order (fields of pieces not in hand are 0); returns the number of legal actions */
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
__attribute__((target_clones("avx512f", "avx2", "default")))
#endif
static int eval_fields(const orc_ctx *c, const orc_state *s, int p, uint32_t *fields) {
    rows_t r;
    build_rows(c, s, p, &r);
    const int N = c->N;
    int cnt = 0;
    for (int o = 0; o < c->nori; o++) {
        const orient_t *ot = &c->ori[o];
        const int R = N - ot->h + 1;
        uint32_t *f = fields + c->fbase[o];
        if (!((s->inv[p] >> ot->piece) & 1)) { for (int ay = 0; ay < R; ay++) f[ay] = 0; continue; }
        uint32_t fit[MAXN], touch[MAXN];
        const uint32_t *f0 = r.frs[ot->dx[0]] + ot->dy[0], *d0 = r.dgs[ot->dx[0]] + ot->dy[0];
        for (int ay = 0; ay < R; ay++) { fit[ay] = f0[ay]; touch[ay] = d0[ay]; }
        for (int k = 1; k < ot->n; k++) {
            const uint32_t *fk = r.frs[ot->dx[k]] + ot->dy[k], *dk = r.dgs[ot->dx[k]] + ot->dy[k];
            for (int ay = 0; ay < R; ay++) { fit[ay] &= fk[ay]; touch[ay] |= dk[ay]; }
        }
        for (int ay = 0; ay < R; ay++) {      /* bits beyond N-w are 0 because fr is masked to N bits */
            const uint32_t v = fit[ay] & touch[ay];
            f[ay] = v;
            cnt += __builtin_popcount(v);
        }
    }
    return cnt;
}

static void fields_to_mask(const orc_ctx *c, const uint32_t *fields, uint8_t *mask) {
    memset(mask, 0, (size_t)c->nact);
    for (int i = 0; i < c->nfields; i++) {
        uint32_t f = fields[i];
        while (f) { int x = __builtin_ctz(f); f &= f - 1; mask[c->foff[i] + x] = 1; }
    }
}

int orc_fast_legal_mask(const orc_ctx *c, const orc_state *s, int p, uint8_t *mask) {
    if (s->done) { if (mask) memset(mask, 0, (size_t)c->nact); return 0; }
    uint32_t fields[MAXFIELDS];
    int cnt = eval_fields(c, s, p, fields);
    if (mask) fields_to_mask(c, fields, mask);
    return cnt;
}

static int has_move_fast(const orc_ctx *c, const orc_state *s, int p) {
    uint32_t fields[MAXFIELDS];
    return eval_fields(c, s, p, fields) > 0;
}

/* ---- scoring, winners: R10, R11 ---- */
int orc_final_score(const orc_ctx *c, const orc_state *s, int p) {
    int sc = s->score[p];
    if (c->rule == 1 && s->inv[p] == 0) sc += 15 + (s->lastmono[p] ? 5 : 0);
    return sc;
}
/* bitmask of winners (players with the best final score); 0 while the game is running (R9) */
int orc_winners(const orc_ctx *c, const orc_state *s) {
    if (!s->done) return 0;
    int best = -32768, m = 0;
    for (int p = 0; p < c->P; p++) { int v = orc_final_score(c, s, p); if (v > best) best = v; }
    for (int p = 0; p < c->P; p++) if (orc_final_score(c, s, p) == best) m |= 1 << p;
    return m;
}
/* blokus_wrapper.py:177-185: -1 everywhere, 3 for a sole winner, 1 for each tied winner; zeros while running */
void orc_terminal_values(const orc_ctx *c, const orc_state *s, float *v) {
    int w = orc_winners(c, s), nw = __builtin_popcount(w);
    for (int p = 0; p < c->P; p++) v[p] = !s->done ? 0.f : (((w >> p) & 1) ? (nw == 1 ? 3.f : 1.f) : -1.f);
}

/* ---- step: R7-R9.  returns 0 ok, 1 illegal action (state unchanged), 2 game already over ---- */
static void place(const orc_ctx *c, orc_state *s, int p, int a) {
    const orient_t *o = &c->ori[c->act_o[a]];
    for (int k = 0; k < o->n; k++) {
        const int y = c->act_y[a] + o->dy[k], x = c->act_x[a] + o->dx[k];
        s->board[y * c->N + x] = (uint8_t)(p + 1);
        s->rows[p][y] |= 1u << x;
    }
    s->inv[p] &= ~(1u << o->piece);
    s->score[p] = (int16_t)(s->score[p] + o->n);
    s->lastmono[p] = (uint8_t)(o->piece == 0);
    s->ply++;
}
static int step_impl(const orc_ctx *c, orc_state *s, int a, int fast) {
    if (s->done) return 2;
    int p = s->mover;
    if (a < 0 || a >= c->nact || !placement_legal(c, s, p, a)) return 1;
    place(c, s, p, a);
    for (int k = 1; k <= c->P; k++) {                /* R8: auto-skip players without a move */
        int q = (p + k) % c->P;
        if (fast ? has_move_fast(c, s, q) : has_move_naive(c, s, q)) { s->mover = (uint8_t)q; return 0; }
    }
    s->done = 1;                                    /* R9: nobody can move; mover stays = last mover */
    return 0;
}
int orc_step(const orc_ctx *c, orc_state *s, int a) { return step_impl(c, s, a, 0); }
int orc_fast_step(const orc_ctx *c, orc_state *s, int a) { return step_impl(c, s, a, 1); }

/* ---- observation R13: planes 0..P-1 occupancy, planes P..2P-1 one-hot mover broadcast ---- */
void orc_observe(const orc_ctx *c, const orc_state *s, float *obs) {
    int nn = c->N * c->N;
    memset(obs, 0, sizeof(float) * 2 * c->P * nn);
    for (int i = 0; i < nn; i++) if (s->board[i]) obs[(s->board[i] - 1) * nn + i] = 1.f;
    for (int i = 0; i < nn; i++) obs[(c->P + s->mover) * nn + i] = 1.f;
}
void orc_board_contents(const orc_ctx *c, const orc_state *s, uint8_t *out) { memcpy(out, s->board, (size_t)(c->N * c->N)); }

/* ---- engine state format (DESIGN.md "state layout"): u32 words
 *   [q*N + y]        row y of player q, bit x = column x
 *   [P*N + q]        inventory of player q
 *   [P*N + P]        meta: mover(0..3) | done<<4 | lastmono<<8 | ply<<16
 *   [P*N + P + 1]    game counter
 *   [P*N + P + 2..3] scores, int16 x 4 little endian
 * Packed from the uint8 board (not from the row words the fast path keeps), so a drift between the two shows up.
 */
int orc_state_words(const orc_ctx *c) { return c->P * c->N + c->P + 4; }
void orc_pack(const orc_ctx *c, const orc_state *s, uint32_t *w) {
    int nw = orc_state_words(c);
    memset(w, 0, sizeof(uint32_t) * nw);
    for (int y = 0; y < c->N; y++)
        for (int x = 0; x < c->N; x++) { int v = s->board[y * c->N + x]; if (v) w[(v - 1) * c->N + y] |= 1u << x; }
    int b = c->P * c->N;
    uint32_t lm = 0;
    for (int p = 0; p < c->P; p++) { w[b + p] = s->inv[p]; lm |= (uint32_t)(s->lastmono[p] & 1) << p; }
    w[b + c->P] = (uint32_t)s->mover | ((uint32_t)s->done << 4) | (lm << 8) | ((uint32_t)s->ply << 16);
    w[b + c->P + 1] = s->game;
    for (int p = 0; p < c->P; p++) w[b + c->P + 2 + (p >> 1)] |= (uint32_t)(uint16_t)s->score[p] << (16 * (p & 1));
}
/* 1 when the row words agree with the uint8 board (the two representations the two rule sets read) */
int orc_rows_consistent(const orc_ctx *c, const orc_state *s) {
    for (int q = 0; q < c->P; q++)
        for (int y = 0; y < c->N; y++) {
            uint32_t w = 0;
            for (int x = 0; x < c->N; x++) if (s->board[y * c->N + x] == q + 1) w |= 1u << x;
            if (w != s->rows[q][y]) return 0;
        }
    return 1;
}
void orc_unpack(const orc_ctx *c, const uint32_t *w, orc_state *s) {
    memset(s, 0, sizeof(*s));
    s->N = c->N; s->P = c->P;
    for (int q = 0; q < c->P; q++)
        for (int y = 0; y < c->N; y++) {
            s->rows[q][y] = w[q * c->N + y];
            for (int x = 0; x < c->N; x++) if ((w[q * c->N + y] >> x) & 1) s->board[y * c->N + x] = (uint8_t)(q + 1);
        }
    int b = c->P * c->N;
    uint32_t m = w[b + c->P];
    for (int p = 0; p < c->P; p++) {
        s->inv[p] = w[b + p]; s->lastmono[p] = (uint8_t)((m >> (8 + p)) & 1);
        s->score[p] = (int16_t)(uint16_t)(w[b + c->P + 2 + (p >> 1)] >> (16 * (p & 1)));
    }
    s->mover = (uint8_t)(m & 15); s->done = (uint8_t)((m >> 4) & 1); s->ply = (uint16_t)(m >> 16);
    s->game = w[b + c->P + 1];
}

/* ---- counter-based RNG: Philox-4x32-10 (Salmon et al., SC'11), key=(seed_lo, seed_hi^env), ctr=(ply>>2, game, stream, 0) ---- */
void orc_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t *out) {
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
/* the sampler's draw: k = mulhi32(u, n_legal) with u = word (ply & 3) of the block at counter (ply >> 2, game, stream, 0) */
static uint32_t draw_k(const orc_state *s, uint64_t seed, uint32_t env_id, uint32_t stream, int n) {
    uint32_t r[4];
    orc_philox(s->ply >> 2, s->game, stream, 0, (uint32_t)seed, (uint32_t)(seed >> 32) ^ env_id, r);
    return (uint32_t)(((uint64_t)r[s->ply & 3] * (uint64_t)(uint32_t)n) >> 32);
}
/* uniform legal action for the mover of s: k-th legal id ascending. -1 if none */
int orc_sample_action(const orc_ctx *c, const orc_state *s, uint64_t seed, uint32_t env_id, uint32_t stream,
                      uint8_t *scratch_mask, int fast) {
    int n = fast ? orc_fast_legal_mask(c, s, s->mover, scratch_mask) : orc_legal_mask(c, s, s->mover, scratch_mask);
    if (n == 0) return -1;
    uint32_t k = draw_k(s, seed, env_id, stream, n);
    for (int a = 0; a < c->nact; a++) if (scratch_mask[a]) { if (k == 0) return a; k--; }
    return -1;
}
/* the same pick from the fields (ascending field order = ascending id order): no 30 K-byte scan */
static int pick_from_fields(const orc_ctx *c, const uint32_t *fields, uint32_t k) {
    for (int i = 0; i < c->nfields; i++) {
        const uint32_t n = (uint32_t)__builtin_popcount(fields[i]);
        if (k < n) {
            uint32_t f = fields[i];
            while (k--) f &= f - 1;
            return c->foff[i] + __builtin_ctz(f);
        }
        k -= n;
    }
    return -1;
}

/* Random play of one env for `plies` plies with auto-reset; logs actions; returns plies executed.
 * counters: [0] steps, [1] games finished.  Written the slow, obvious way on purpose (mask, then a linear scan): the
 * lockstep tests use it as the independent statement of "uniform legal play"; orc_play_many below is the fast form. */
int64_t orc_random_play(const orc_ctx *c, orc_state *s, uint64_t seed, uint32_t env_id, int plies, int auto_reset, int fast,
                        int32_t *action_log, int64_t *counters) {
    uint8_t *mask = malloc((size_t)c->nact);
    int64_t done_plies = 0;
    for (int i = 0; i < plies; i++) {
        if (s->done) break;
        int a = orc_sample_action(c, s, seed, env_id, 0, mask, fast);
        if (a < 0) break;
        if (action_log) action_log[i] = a;
        if (fast) orc_fast_step(c, s, a); else orc_step(c, s, a);
        done_plies++;
        if (counters) { counters[0]++; if (s->done) counters[1]++; }
        if (s->done && auto_reset) { uint32_t g = s->game + 1; orc_reset(c, s, g); }   /* eager, like the engine */
    }
    free(mask);
    return done_plies;
}

/* ---- many envs, the fast way: what bench.py times as the CPU arm and what the full-size parity tests compare with ----
 * Each env i (global id env_id0 + i * env_stride) plays `plies` plies of uniform-random legal play with auto-reset,
 * exactly the engine's random-play workload: per ply ONE evaluation of the new mover's fields (it doubles as the
 * "can this player move" test of the auto-skip rule), the FULL byte mask written (mask_out != NULL: env i's row is
 * mask_out + i * mask_stride, as the metric demands -- stride 0 keeps one cache-resident row per thread; NULL in the
 * parity runs that only need checksums), the next action drawn from the fields.
 *   traj[i]            running hash over the env's sampled actions: h = h * 0x9E3779B97F4A7C15 + (action + 2), from h = 0,
 *                      one update per ply including the draw on the state after the last ply
 *   cnt_sum[t]         sum over envs of legal_count(env, t) * (2 * global_id + 1), t = 0 (fresh masks) .. plies
 *   sel_plies[nsel]    plies t at which per-env mask checksums are taken (ascending), into
 *   ids_sum[i*nsel+j]  sum over legal ids of (id + 1)
 *   words_sum[i*nsel+j] sum over legal ids of 2^(id & 31) * (2 * (id >> 5) + 1)   (= sum of bit-mask words * (2g + 1))
 *   counters[0..2]     steps, games finished, sum of legal counts seen by the sampler
 * Every accumulator is a wrapping uint64: partial results of several threads add up. */
void orc_play_many(const orc_ctx *c, orc_state *states, int64_t n, uint64_t seed, uint32_t env_id0, uint32_t env_stride,
                   int plies, int auto_reset, uint8_t *mask_out, int64_t mask_stride, uint64_t *traj, uint64_t *cnt_sum,
                   const int32_t *sel_plies, int nsel, uint64_t *ids_sum, uint64_t *words_sum, int64_t *counters) {
    uint32_t fields[MAXFIELDS];
    int64_t n_steps = 0, n_games = 0, n_legal = 0;     /* local: several threads' counter rows may share a cache line */
    for (int64_t i = 0; i < n; i++) {
        orc_state *s = &states[i];
        const uint32_t gid = env_id0 + (uint32_t)i * env_stride;
        const uint64_t wgt = 2ull * gid + 1ull;
        uint64_t h = traj ? traj[i] : 0;
        int sel = 0;
        int cnt = 0;
        if (s->done) memset(fields, 0, sizeof(uint32_t) * (size_t)c->nfields);
        else cnt = eval_fields(c, s, s->mover, fields);
        for (int t = 0;; t++) {
            /* outputs of "step t": the mover's legal set and the action drawn from it */
            if (mask_out) fields_to_mask(c, fields, mask_out + i * mask_stride);
            int a = cnt > 0 ? pick_from_fields(c, fields, draw_k(s, seed, gid, 0, cnt)) : -1;
            h = h * 0x9E3779B97F4A7C15ull + (uint64_t)(int64_t)(a + 2);
            if (cnt_sum) cnt_sum[t] += (uint64_t)cnt * wgt;
            if (sel < nsel && sel_plies[sel] == t) {
                uint64_t si = 0, sw = 0;
                for (int f = 0; f < c->nfields && cnt > 0; f++) {
                    uint32_t v = fields[f];
                    while (v) {
                        const uint32_t id = (uint32_t)c->foff[f] + (uint32_t)__builtin_ctz(v);
                        v &= v - 1;
                        si += id + 1;
                        sw += (1ull << (id & 31)) * (2ull * (id >> 5) + 1ull);
                    }
                }
                if (ids_sum) ids_sum[i * nsel + sel] = si;
                if (words_sum) words_sum[i * nsel + sel] = sw;
                sel++;
            }
            if (t == plies) break;
            if (a < 0) {                            /* finished game without auto-reset: the env stays as it is */
                cnt = 0;
                continue;
            }
            n_steps++; n_legal += cnt;
            const int p = s->mover;
            place(c, s, p, a);
            cnt = 0;
            int found = 0;
            for (int k = 1; k <= c->P && !found; k++) {      /* R8 */
                const int q = (p + k) % c->P;
                cnt = eval_fields(c, s, q, fields);
                if (cnt > 0) { s->mover = (uint8_t)q; found = 1; }
            }
            if (!found) {
                s->done = 1;                                 /* R9 */
                n_games++;
                if (auto_reset) {
                    const uint32_t g = s->game + 1;
                    orc_reset(c, s, g);
                    cnt = eval_fields(c, s, 0, fields);
                } else {
                    memset(fields, 0, sizeof(uint32_t) * (size_t)c->nfields);
                }
            }
        }
        if (traj) traj[i] = h;
    }
    if (counters) { counters[0] += n_steps; counters[1] += n_games; counters[2] += n_legal; }
}

/* Uniform-random playout from `root` to the end of the game with the engine's playout stream (stream 1, key
 * seed_hi ^ game_index): final scores, winners bitmask, plies played; optional action log (0xFFFF terminated, like the
 * engine's).  stop_player >= 0 stops as soon as it is that player's turn.  Returns plies played. */
int orc_playout_stream(const orc_ctx *c, const orc_state *root, uint64_t seed, uint32_t game_index, uint32_t stream, int stop_player,
                       orc_state *out, int16_t *final_scores, int32_t *winners, uint16_t *action_log, int log_stride) {
    uint32_t fields[MAXFIELDS];
    orc_state s = *root;
    int nply = 0;
    if (!s.done) {
        int cnt = eval_fields(c, &s, s.mover, fields);
        /* the root's mover is evaluated first; a root whose mover is stuck is over (the engine does the same) */
        if (cnt == 0) s.done = 1;
        while (!s.done) {
            if ((int)s.mover == stop_player) break;
            const int a = pick_from_fields(c, fields, draw_k(&s, seed, game_index, stream, cnt));
            if (action_log && nply < log_stride - 1) action_log[nply] = (uint16_t)a;
            const int p = s.mover;
            place(c, &s, p, a);
            nply++;
            int found = 0;
            for (int k = 1; k <= c->P && !found; k++) {
                const int q = (p + k) % c->P;
                cnt = eval_fields(c, &s, q, fields);
                if (cnt > 0) { s.mover = (uint8_t)q; found = 1; }
            }
            if (!found) s.done = 1;
        }
    }
    if (action_log) action_log[nply < log_stride - 1 ? nply : log_stride - 1] = 0xFFFFu;
    if (final_scores) for (int p = 0; p < c->P; p++) final_scores[p] = (int16_t)orc_final_score(c, &s, p);
    if (winners) *winners = orc_winners(c, &s);
    if (out) *out = s;
    return nply;
}
/* the playout kernel's stream (1); the fused search kernel's in-kernel playouts use stream 2 */
int orc_playout(const orc_ctx *c, const orc_state *root, uint64_t seed, uint32_t game_index, int stop_player,
                orc_state *out, int16_t *final_scores, int32_t *winners, uint16_t *action_log, int log_stride) {
    return orc_playout_stream(c, root, seed, game_index, 1, stop_player, out, final_scores, winners, action_log, log_stride);
}
