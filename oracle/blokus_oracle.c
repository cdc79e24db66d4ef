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
 *   action table   blokus_wrapper.py:281-324                         -> orc_init / orc_action_cells
 * The executable pins that do exist are checked in tests/test_oracle_kat.py:
 *   30,433 actions on 20x20 (blokus_nnet.py:17,97), observation (8,20,20) (blokus_nnet.py:99),
 *   terminal vector 3/1/-1 (blokus_wrapper.py:177-185), colours 0..4 (blokus_wrapper.py:259).
 *
 * Two independent legality implementations live here and are cross-checked by the tests:
 *   - naive, cell by cell on a uint8 board (orc_legal_mask)        -- obviously right
 *   - bit-parallel on row words (orc_fast_legal_mask)              -- the timed CPU baseline
 * A third, pure-Python set-based one is oracle/naive.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXN 20
#define MAXP 4
#define NPIECES 21
#define MAXORI 91
#define MAXCELLS 5

typedef struct {
    int32_t N, P;
    uint8_t board[MAXN * MAXN];   /* 0 empty, 1..P = colour of player 0..P-1 (R1) */
    uint32_t inv[MAXP];           /* bit i set = piece i still in hand (R4) */
    int16_t score[MAXP];          /* squares placed (R10 default) */
    uint8_t lastmono[MAXP];       /* most recent placement was the monomino */
    uint8_t mover;                /* player to move */
    uint8_t done;
    uint16_t ply;                 /* placements made in this game */
    uint32_t game;                /* games finished before this one (auto-reset counter) */
} orc_state;

typedef struct {
    int piece, local, h, w, n;
    int8_t dy[MAXCELLS], dx[MAXCELLS];
} orient_t;

/* Shapes, rows top->bottom, SURVEY.md Appendix B order. */
static const char *SHAPES[NPIECES] = {
    "#", "#/#", "#/#/#", "##/#.", "#/#/#/#", "##/#./#.", "#./##/#.", "##/##", "#./##/.#",
    "#/#/#/#/#", "##/#./#./#.", "#./##/#./#.", "##/##/#.", "##/#./##", "###/#../#..",
    "#../###/#..", "#./#./##/.#", "#../###/.#.", "#../##./.##", "#../###/..#", ".#./###/.#."};

static int g_N = 0, g_P = 0, g_rule = 0;
static int g_nori = 0, g_nact = 0;
static orient_t g_ori[MAXORI];
static int g_psize[NPIECES];
static int g_obase[MAXORI + 1];
static int16_t *g_act_o = 0, *g_act_y = 0, *g_act_x = 0;

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

static void build_orientations(void) {
    g_nori = 0;
    for (int p = 0; p < NPIECES; p++) {
        shape_t base; base.n = 0;
        int y = 0, x = 0;
        for (const char *c = SHAPES[p]; *c; c++) {
            if (*c == '/') { y++; x = 0; continue; }
            if (*c == '#') { base.y[base.n] = y; base.x[base.n] = x; base.n++; }
            x++;
        }
        g_psize[p] = base.n;
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
            orient_t *o = &g_ori[g_nori++];
            o->piece = p; o->local = i; o->n = list[i].n; o->h = 0; o->w = 0;
            for (int c = 0; c < o->n; c++) {
                o->dy[c] = (int8_t)list[i].y[c]; o->dx[c] = (int8_t)list[i].x[c];
                if (list[i].y[c] + 1 > o->h) o->h = list[i].y[c] + 1;
                if (list[i].x[c] + 1 > o->w) o->w = list[i].x[c] + 1;
            }
        }
    }
}

int orc_init(int N, int P, int score_rule) {
    if (N < 5 || N > MAXN || (P != 2 && P != 4)) return -1;
    g_N = N; g_P = P; g_rule = score_rule;
    build_orientations();
    free(g_act_o); free(g_act_y); free(g_act_x);
    int cap = 0;
    for (int o = 0; o < g_nori; o++) cap += (N - g_ori[o].h + 1) * (N - g_ori[o].w + 1);
    g_act_o = malloc(sizeof(int16_t) * cap); g_act_y = malloc(sizeof(int16_t) * cap);
    g_act_x = malloc(sizeof(int16_t) * cap);
    g_nact = 0;
    for (int o = 0; o < g_nori; o++) {               /* id order: (piece, orientation, ay, ax) */
        g_obase[o] = g_nact;
        for (int y = 0; y + g_ori[o].h <= N; y++)
            for (int x = 0; x + g_ori[o].w <= N; x++) {
                g_act_o[g_nact] = (int16_t)o; g_act_y[g_nact] = (int16_t)y; g_act_x[g_nact] = (int16_t)x;
                g_nact++;
            }
    }
    g_obase[g_nori] = g_nact;
    return 0;
}

int orc_num_actions(void) { return g_nact; }
int orc_num_orientations(void) { return g_nori; }
int orc_state_size(void) { return (int)sizeof(orc_state); }
int orc_piece_size(int p) { return g_psize[p]; }

/* cells_yx: up to 10 bytes (y0,x0,y1,x1,...). returns ncells; meta = {piece, orient, ay, ax} */
int orc_action_cells(int a, uint8_t *cells_yx, int32_t *meta) {
    if (a < 0 || a >= g_nact) return -1;
    const orient_t *o = &g_ori[g_act_o[a]];
    for (int c = 0; c < o->n; c++) {
        cells_yx[2 * c] = (uint8_t)(g_act_y[a] + o->dy[c]);
        cells_yx[2 * c + 1] = (uint8_t)(g_act_x[a] + o->dx[c]);
    }
    if (meta) { meta[0] = o->piece; meta[1] = g_act_o[a]; meta[2] = g_act_y[a]; meta[3] = g_act_x[a]; }
    return o->n;
}

/* R3: start corners. */
static void start_corner(int p, int *y, int *x) {
    int n = g_N - 1;
    if (g_P == 2) { *y = p ? n : 0; *x = p ? n : 0; return; }
    *y = (p & 2) ? n : 0; *x = (p & 1) ? n : 0;
}

void orc_reset(orc_state *s, uint32_t game) {
    memset(s, 0, sizeof(*s));
    s->N = g_N; s->P = g_P; s->game = game;
    for (int p = 0; p < g_P; p++) s->inv[p] = (1u << NPIECES) - 1;
}

/* ---- naive legality: R5, R6 ---- */
static inline int cell(const orc_state *s, int y, int x) {
    if (y < 0 || x < 0 || y >= g_N || x >= g_N) return 0;
    return s->board[y * g_N + x];
}

static int placement_legal(const orc_state *s, int p, int a) {
    const orient_t *o = &g_ori[g_act_o[a]];
    if (!((s->inv[p] >> o->piece) & 1)) return 0;
    int col = p + 1, first = (s->inv[p] == (1u << NPIECES) - 1);
    int cy, cx, corner = 0;
    start_corner(p, &cy, &cx);
    for (int c = 0; c < o->n; c++) {
        int y = g_act_y[a] + o->dy[c], x = g_act_x[a] + o->dx[c];
        if (cell(s, y, x) != 0) return 0;                                  /* overlap */
        if (cell(s, y - 1, x) == col || cell(s, y + 1, x) == col ||
            cell(s, y, x - 1) == col || cell(s, y, x + 1) == col) return 0; /* own edge contact */
        if (first) { if (y == cy && x == cx) corner = 1; }
        else if (cell(s, y - 1, x - 1) == col || cell(s, y - 1, x + 1) == col ||
                 cell(s, y + 1, x - 1) == col || cell(s, y + 1, x + 1) == col) corner = 1;
    }
    return corner;
}

/* mask: g_nact bytes 0/1. returns number of legal actions for player p (any player, not only the mover) */
int orc_legal_mask(const orc_state *s, int p, uint8_t *mask) {
    int cnt = 0;
    for (int a = 0; a < g_nact; a++) {
        int l = s->done ? 0 : placement_legal(s, p, a);
        if (mask) mask[a] = (uint8_t)l;
        cnt += l;
    }
    return cnt;
}

static int has_move_naive(const orc_state *s, int p) {
    for (int a = 0; a < g_nact; a++) if (placement_legal(s, p, a)) return 1;
    return 0;
}

/* ---- bit-parallel legality (second, independent formulation; also the timed CPU engine) ---- */
typedef struct { uint32_t fr[MAXN + 5], dg[MAXN + 5]; } rows_t;

static void build_rows(const orc_state *s, int p, rows_t *r) {
    uint32_t own[MAXN + 2], occ[MAXN + 2], full = (g_N >= 32) ? 0xffffffffu : ((1u << g_N) - 1);
    memset(own, 0, sizeof(own)); memset(occ, 0, sizeof(occ));
    for (int y = 0; y < g_N; y++)
        for (int x = 0; x < g_N; x++) {
            int c = s->board[y * g_N + x];
            if (c) occ[y + 1] |= 1u << x;
            if (c == p + 1) own[y + 1] |= 1u << x;
        }
    int first = (s->inv[p] == (1u << NPIECES) - 1), cy, cx;
    start_corner(p, &cy, &cx);
    memset(r, 0, sizeof(*r));
    for (int y = 0; y < g_N; y++) {
        uint32_t o = own[y + 1], ud = own[y] | own[y + 2];
        uint32_t adj = ud | (o << 1) | (o >> 1);
        r->fr[y] = ~(occ[y + 1] | adj) & full;
        r->dg[y] = first ? ((y == cy) ? (1u << cx) : 0) : (((ud << 1) | (ud >> 1)) & full);
    }
}

static inline uint32_t field_bits(const rows_t *r, const orient_t *o, int ay) {
    uint32_t fit = 0xffffffffu, touch = 0;
    for (int c = 0; c < o->n; c++) {
        fit &= r->fr[ay + o->dy[c]] >> o->dx[c];
        touch |= r->dg[ay + o->dy[c]] >> o->dx[c];
    }
    return fit & touch;   /* bits beyond N-w are 0 because fr is masked to N bits */
}

int orc_fast_legal_mask(const orc_state *s, int p, uint8_t *mask) {
    rows_t r; int cnt = 0;
    if (mask) memset(mask, 0, (size_t)g_nact);
    if (s->done) return 0;
    build_rows(s, p, &r);
    for (int o = 0; o < g_nori; o++) {
        if (!((s->inv[p] >> g_ori[o].piece) & 1)) continue;
        int W = g_N - g_ori[o].w + 1;
        for (int ay = 0; ay + g_ori[o].h <= g_N; ay++) {
            uint32_t f = field_bits(&r, &g_ori[o], ay);
            cnt += __builtin_popcount(f);
            if (mask) while (f) { int x = __builtin_ctz(f); f &= f - 1; mask[g_obase[o] + ay * W + x] = 1; }
        }
    }
    return cnt;
}

static int has_move_fast(const orc_state *s, int p) {
    rows_t r; build_rows(s, p, &r);
    for (int o = 0; o < g_nori; o++) {
        if (!((s->inv[p] >> g_ori[o].piece) & 1)) continue;
        for (int ay = 0; ay + g_ori[o].h <= g_N; ay++) if (field_bits(&r, &g_ori[o], ay)) return 1;
    }
    return 0;
}

/* ---- scoring, winners: R10, R11 ---- */
int orc_final_score(const orc_state *s, int p) {
    int sc = s->score[p];
    if (g_rule == 1 && s->inv[p] == 0) sc += 15 + (s->lastmono[p] ? 5 : 0);
    return sc;
}
/* bitmask of winners (players with the best final score); 0 while the game is running (R9) */
int orc_winners(const orc_state *s) {
    if (!s->done) return 0;
    int best = -32768, m = 0;
    for (int p = 0; p < g_P; p++) { int v = orc_final_score(s, p); if (v > best) best = v; }
    for (int p = 0; p < g_P; p++) if (orc_final_score(s, p) == best) m |= 1 << p;
    return m;
}
/* blokus_wrapper.py:177-185: -1 everywhere, 3 for a sole winner, 1 for each tied winner; zeros while running */
void orc_terminal_values(const orc_state *s, float *v) {
    int w = orc_winners(s), nw = __builtin_popcount(w);
    for (int p = 0; p < g_P; p++) v[p] = !s->done ? 0.f : (((w >> p) & 1) ? (nw == 1 ? 3.f : 1.f) : -1.f);
}

/* ---- step: R7-R9.  returns 0 ok, 1 illegal action (state unchanged), 2 game already over ---- */
static int step_impl(orc_state *s, int a, int fast) {
    if (s->done) return 2;
    int p = s->mover;
    if (a < 0 || a >= g_nact || !placement_legal(s, p, a)) return 1;
    const orient_t *o = &g_ori[g_act_o[a]];
    for (int c = 0; c < o->n; c++)
        s->board[(g_act_y[a] + o->dy[c]) * g_N + g_act_x[a] + o->dx[c]] = (uint8_t)(p + 1);
    s->inv[p] &= ~(1u << o->piece);
    s->score[p] = (int16_t)(s->score[p] + o->n);
    s->lastmono[p] = (uint8_t)(o->piece == 0);
    s->ply++;
    for (int k = 1; k <= g_P; k++) {                /* R8: auto-skip players without a move */
        int q = (p + k) % g_P;
        if (fast ? has_move_fast(s, q) : has_move_naive(s, q)) { s->mover = (uint8_t)q; return 0; }
    }
    s->done = 1;                                    /* R9: nobody can move; mover stays = last mover */
    return 0;
}
int orc_step(orc_state *s, int a) { return step_impl(s, a, 0); }
int orc_fast_step(orc_state *s, int a) { return step_impl(s, a, 1); }

/* ---- observation R13: planes 0..P-1 occupancy, planes P..2P-1 one-hot mover broadcast ---- */
void orc_observe(const orc_state *s, float *obs) {
    int nn = g_N * g_N;
    memset(obs, 0, sizeof(float) * 2 * g_P * nn);
    for (int i = 0; i < nn; i++) if (s->board[i]) obs[(s->board[i] - 1) * nn + i] = 1.f;
    for (int i = 0; i < nn; i++) obs[(g_P + s->mover) * nn + i] = 1.f;
}
void orc_board_contents(const orc_state *s, uint8_t *out) { memcpy(out, s->board, (size_t)(g_N * g_N)); }

/* ---- engine state format (DESIGN.md "state layout"): u32 words
 *   [q*N + y]        row y of player q, bit x = column x
 *   [P*N + q]        inventory of player q
 *   [P*N + P]        meta: mover(0..3) | done<<4 | lastmono<<8 | ply<<16
 *   [P*N + P + 1]    game counter
 *   [P*N + P + 2..3] scores, int16 x 4 little endian
 */
int orc_state_words(void) { return g_P * g_N + g_P + 4; }
void orc_pack(const orc_state *s, uint32_t *w) {
    int nw = orc_state_words();
    memset(w, 0, sizeof(uint32_t) * nw);
    for (int y = 0; y < g_N; y++)
        for (int x = 0; x < g_N; x++) { int c = s->board[y * g_N + x]; if (c) w[(c - 1) * g_N + y] |= 1u << x; }
    int b = g_P * g_N;
    uint32_t lm = 0;
    for (int p = 0; p < g_P; p++) { w[b + p] = s->inv[p]; lm |= (uint32_t)(s->lastmono[p] & 1) << p; }
    w[b + g_P] = (uint32_t)s->mover | ((uint32_t)s->done << 4) | (lm << 8) | ((uint32_t)s->ply << 16);
    w[b + g_P + 1] = s->game;
    for (int p = 0; p < g_P; p++) w[b + g_P + 2 + (p >> 1)] |= (uint32_t)(uint16_t)s->score[p] << (16 * (p & 1));
}
void orc_unpack(const uint32_t *w, orc_state *s) {
    memset(s, 0, sizeof(*s));
    s->N = g_N; s->P = g_P;
    for (int q = 0; q < g_P; q++)
        for (int y = 0; y < g_N; y++)
            for (int x = 0; x < g_N; x++) if ((w[q * g_N + y] >> x) & 1) s->board[y * g_N + x] = (uint8_t)(q + 1);
    int b = g_P * g_N;
    uint32_t m = w[b + g_P];
    for (int p = 0; p < g_P; p++) {
        s->inv[p] = w[b + p]; s->lastmono[p] = (uint8_t)((m >> (8 + p)) & 1);
        s->score[p] = (int16_t)(uint16_t)(w[b + g_P + 2 + (p >> 1)] >> (16 * (p & 1)));
    }
    s->mover = (uint8_t)(m & 15); s->done = (uint8_t)((m >> 4) & 1); s->ply = (uint16_t)(m >> 16);
    s->game = w[b + g_P + 1];
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
/* uniform legal action for the mover of s: k = mulhi32(u, n_legal), k-th legal id ascending. -1 if none */
int orc_sample_action(const orc_state *s, uint64_t seed, uint32_t env_id, uint32_t stream, uint8_t *scratch_mask, int fast) {
    int n = fast ? orc_fast_legal_mask(s, s->mover, scratch_mask) : orc_legal_mask(s, s->mover, scratch_mask);
    if (n == 0) return -1;
    uint32_t r[4];
    /* one Philox block serves four consecutive plies: counter (ply >> 2, game, stream, 0), word ply & 3 */
    orc_philox(s->ply >> 2, s->game, stream, 0, (uint32_t)seed, (uint32_t)(seed >> 32) ^ env_id, r);
    uint32_t k = (uint32_t)(((uint64_t)r[s->ply & 3] * (uint64_t)(uint32_t)n) >> 32);
    for (int a = 0; a < g_nact; a++) if (scratch_mask[a]) { if (k == 0) return a; k--; }
    return -1;
}

/* Random play of one env for `plies` plies with auto-reset; logs actions; returns plies executed.
 * counters: [0] steps, [1] games finished, [2] sum of legal counts seen by the sampler */
int64_t orc_random_play(orc_state *s, uint64_t seed, uint32_t env_id, int plies, int auto_reset, int fast,
                        int32_t *action_log, int64_t *counters) {
    uint8_t *mask = malloc((size_t)g_nact);
    int64_t done_plies = 0;
    for (int i = 0; i < plies; i++) {
        if (s->done) break;
        int a = orc_sample_action(s, seed, env_id, 0, mask, fast);
        if (a < 0) break;
        if (action_log) action_log[i] = a;
        if (fast) orc_fast_step(s, a); else orc_step(s, a);
        done_plies++;
        if (counters) { counters[0]++; if (s->done) counters[1]++; }
        if (s->done && auto_reset) { uint32_t g = s->game + 1; orc_reset(s, g); }   /* eager, like the engine */
    }
    free(mask);
    return done_plies;
}
