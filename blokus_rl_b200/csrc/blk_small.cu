// blk_small.cu -- step / legal-mask kernel for boards with N <= 7 (the reference's "simplified Blokus":
// config/ppo_blokus_7x7.yml and the 7x7 AlphaZero experiments, docs/README.md:51,73).  Compiled once per board
// size: nvcc -DBLK_SMALL_N=7 (see blokus_rl_b200/build.py).
//
// Mapping: ONE THREAD PER ENV.  A 7x7 bitboard fits one 64-bit word (row stride 8: bit 8y + x; column 7 and
// row 7 are zero padding, so shifted reads never wrap into a neighbouring row), which makes an orientation's
// legality at ALL anchors a chain of 64-bit ANDs / ORs over shifted copies of the "free" and "diagonal contact"
// boards -- 32 envs per warp instruction instead of one env per warp with 7 of 32 lanes busy (the warp-per-env
// kernel in blk_kernels.cuh, which stays the path for N >= 8).
// The action-id-ordered mask words are assembled in registers by straight-line code whose bit positions are all
// compile-time constants (blk_small_fields.inc, generated) and parked in shared memory; the block then streams
// states and masks to HBM cooperatively (coalesced 16 B stores through the same byte LUT as the big kernel).
//
// Players that cannot move (auto-skip, R8), the end of the game (R9) and auto-reset are resolved in ROUNDS over a
// compacted work list: round 1 evaluates every env's next candidate, later rounds only the envs whose candidate
// had no move, so the rare second and third evaluations do not drag whole warps along.  A fresh board's mask
// is a constant and is copied, not evaluated.
#include "blk_kernels.cuh"

#ifndef BLK_SMALL_N
#error "compile with -DBLK_SMALL_N=<5|6|7>"
#endif
#define BLK_SMALL_DECL
#include "blk_small_fields.inc"
#undef BLK_SMALL_DECL

#ifndef BLK_SMALL_T
#define BLK_SMALL_T 256          // envs per block and pass at P = 2 (P = 4 states are larger: half of it)
#endif

namespace blk {

namespace {

constexpr int kST2 = BLK_SMALL_T, kST4 = BLK_SMALL_T / 2;
constexpr int kSN = BLK_SMALL_N;
constexpr int kSA = BLK_SMALL_A;
constexpr int kSMW = (kSA + 31) / 32;      // mask words that carry action bits
constexpr int kSRS = kSMW | 1;             // odd row stride in shared memory: thread t, word k -> bank (t*RS + k) % 32

constexpr uint64_t small_full() {
    uint64_t f = 0;
    for (int y = 0; y < kSN; ++y) f |= static_cast<uint64_t>((1u << kSN) - 1u) << (8 * y);
    return f;
}
constexpr uint64_t kSFull = small_full();

// control word of a slot between rounds
constexpr uint32_t kCtlFresh = 1u << 9, kCtlMoved = 1u << 8, kCtlEnded = 1u << 11, kCtlIllegal = 1u << 12,
                   kCtlWasDone = 1u << 13;

__device__ __forceinline__ uint64_t rows_to_board(const uint32_t *rows) {
    uint32_t lo = 0u, hi = 0u;
#pragma unroll
    for (int y = 0; y < kSN; ++y) {
        if (y < 4) lo |= rows[y] << (8 * y);
        else hi |= rows[y] << (8 * (y - 4));
    }
    return (static_cast<uint64_t>(hi) << 32) | lo;
}

__device__ __forceinline__ void board_to_rows(uint64_t b, uint32_t *rows) {
#pragma unroll
    for (int y = 0; y < kSN; ++y) rows[y] = static_cast<uint32_t>(b >> (8 * y)) & ((1u << kSN) - 1u);
}

// "free and not edge-adjacent to own colour" and "diagonal contact / start corner" boards of player q (R5, R6)
template <int kP>
__device__ __forceinline__ void small_prep(uint64_t own, uint64_t occ, bool first, int q, uint64_t &fr, uint64_t &dg) {
    const uint64_t ud = (own << 8) | (own >> 8);
    const uint64_t adj = ud | (own << 1) | (own >> 1);
    fr = ~(occ | adj) & kSFull;
    const int n1 = kSN - 1;
    const int cy = (kP == 2) ? (q ? n1 : 0) : ((q & 2) ? n1 : 0);
    const int cx = (kP == 2) ? (q ? n1 : 0) : ((q & 1) ? n1 : 0);
    dg = first ? (1ull << (8 * cy + cx)) : (((ud << 1) | (ud >> 1)) & kSFull);
}

// All 91 orientations at all anchors: writes the kSMW mask words of one env to `dst` and returns the legal count.
// ck[j] receives the number of legal actions in words [0, 16 (j + 1)): the sampler's first search level.
constexpr int kSCk = (kSMW + 15) / 16 - 1;     // checkpoints (4 at N = 7)
__device__ __forceinline__ int small_eval(uint64_t fr, uint64_t dg, uint32_t inv, uint32_t *dst, uint16_t *ck) {
    uint64_t fs[5][5], ds[5][5];
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
        for (int x = 0; x < 5; ++x) {
            fs[r][x] = (r + x <= 4) ? (fr >> (8 * r + x)) : 0ull;
            ds[r][x] = (r + x <= 4) ? (dg >> (8 * r + x)) : 0ull;
        }
    int cnt = 0;
    uint32_t acc = 0u, L0 = 0u, L1 = 0u;
    uint64_t pm_ = 0ull;
#define SM_EMIT(k, v)                                                              \
    {                                                                              \
        const uint32_t v_ = (v);                                                   \
        dst[k] = v_;                                                               \
        cnt += __popc(v_);                                                         \
        if (((k) & 15) == 15 && (k) / 16 < kSCk) ck[(k) / 16] = static_cast<uint16_t>(cnt); \
    }
#define SM_PIECE(p) pm_ = 0ull - static_cast<uint64_t>((inv >> (p)) & 1u);
#define SM_ORIENT(o, p, n, y0, x0, y1, x1, y2, x2, y3, x3, y4, x4)                                              \
    {                                                                                                          \
        const uint64_t l_ = (fs[y0][x0] & fs[y1][x1] & fs[y2][x2] & fs[y3][x3] & fs[y4][x4]) &                 \
                            (ds[y0][x0] | ds[y1][x1] | ds[y2][x2] | ds[y3][x3] | ds[y4][x4]) & pm_;           \
        L0 = static_cast<uint32_t>(l_); L1 = static_cast<uint32_t>(l_ >> 32);                                  \
    }
#define SM_FIELD(h, s, W, fill) acc |= (((h) ? L1 : L0) >> (s) & ((1u << (W)) - 1u)) << (fill);
#define SM_FIELD_X(h, s, W, fill, k)                                         \
    {                                                                        \
        const uint32_t f_ = ((h) ? L1 : L0) >> (s) & ((1u << (W)) - 1u);     \
        SM_EMIT(k, acc | (f_ << (fill)))                                     \
        acc = f_ >> (32 - (fill));                                           \
    }
#define SM_FLUSH(k) { SM_EMIT(k, acc) acc = 0u; }
#define SM_TAIL(k) SM_EMIT(k, acc)
#include "blk_small_fields.inc"
#undef SM_EMIT
#undef SM_PIECE
#undef SM_ORIENT
#undef SM_FIELD
#undef SM_FIELD_X
#undef SM_FLUSH
#undef SM_TAIL
    return cnt;
}

__device__ __forceinline__ int small_score(const uint32_t *st, int q, int P, int score_rule) {
    const uint32_t packed = st[P * kSN + P + 2 + (q >> 1)];
    int s = static_cast<int>(static_cast<int16_t>((packed >> (16 * (q & 1))) & 0xffffu));
    if (score_rule == 1 && st[P * kSN + q] == 0u) s += 15 + (((st[P * kSN + P] >> (8 + q)) & 1u) ? 5 : 0);
    return s;
}

}  // namespace

// kN only makes the three per-size translation units instantiate distinct symbols
template <int kN, int kP, int kFmt, bool kSample>
__global__ void __launch_bounds__(kP == 2 ? kST2 : kST4, kP == 2 ? 512 / kST2 : 384 / kST4) small_step_kernel(const SmallParams sp) {
    static_assert(kN == kSN, "one translation unit per board size");
    constexpr int T = kP == 2 ? kST2 : kST4;        // envs (slots) per block and pass
    constexpr int N = kSN, P = kP;
    constexpr int SW = P * N + P + 4, SWP = SW | 1;
    constexpr int kMeta = P * N + P;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *s_mask = reinterpret_cast<uint32_t *>(smem_raw);            // [T][kSRS]
    uint32_t *s_state = s_mask + T * kSRS;                                // [T][SWP] env states in the HBM word format
    uint32_t *s_ctl = s_state + T * SWP;                                  // [T]
    int *s_cnt = reinterpret_cast<int *>(s_ctl + T);                      // [T] legal count of the final mover
    int16_t *s_fs = reinterpret_cast<int16_t *>(s_cnt + T);               // [T][4] scores after the step, before auto-reset
    uint16_t *s_ck = reinterpret_cast<uint16_t *>(s_fs + 4 * T);          // [T][4] cumulative legal counts per 16 mask words
    uint16_t *s_list = s_ck + 4 * T;                                      // [2][T] work lists of the rounds
    int *s_len = reinterpret_cast<int *>(s_list + 2 * T);                 // [2] (+2 pad)
    uint2 *s_lut = reinterpret_cast<uint2 *>(s_len + 4);                  // [256] byte -> 8 bytes of 0/1
    int32_t *s_obase = reinterpret_cast<int32_t *>(s_lut + 256);          // [92]
    uint32_t *s_oinfo = reinterpret_cast<uint32_t *>(s_obase + 92);       // [92]
    uint64_t *s_ocells = reinterpret_cast<uint64_t *>(s_oinfo + 92);      // [92] footprints, row stride 8
    uint32_t *s_first = reinterpret_cast<uint32_t *>(s_ocells + 92);      // [kSMW] mask of the fresh board
    uint16_t *s_fck = reinterpret_cast<uint16_t *>(s_first + kSMW);       // [4] its cumulative counts per 16 words

    const blk_step_args &a = sp.a;
    const int tid = threadIdx.x;
    for (int i = tid; i < 256; i += T) s_lut[i] = reinterpret_cast<const uint2 *>(sp.tables + kOffLut)[i];
    for (int i = tid; i < 92; i += T) {
        s_obase[i] = reinterpret_cast<const int32_t *>(sp.tables + sp.t.off_obase)[i];
        s_oinfo[i] = i < kOrients ? reinterpret_cast<const uint32_t *>(sp.tables + sp.t.off_oinfo)[i] : 0u;
        s_ocells[i] = i < kOrients ? sp.ocells64[i] : 0ull;
    }
    for (int i = tid; i < kSMW; i += T) s_first[i] = sp.first_mask[i];
    if (tid == 0) {
        int run = 0;
        for (int k = 0; k < kSMW; ++k) {
            run += __popc(sp.first_mask[k]);
            if ((k & 15) == 15 && k / 16 < kSCk) s_fck[k / 16] = static_cast<uint16_t>(run);
        }
    }

    const int64_t n = a.n;
    for (int64_t base = static_cast<int64_t>(blockIdx.x) * T; base < n; base += static_cast<int64_t>(gridDim.x) * T) {
        const int m = static_cast<int>(min(static_cast<int64_t>(T), n - base));
        // ---- phase 0: coalesced load of the block's env states ----
        if (a.state_index == nullptr) {
            for (int i = tid; i < m * SW; i += T) {
                const int slot = i / SW, w = i - slot * SW;
                s_state[slot * SWP + w] = __ldg(a.state_in + base * SW + i);
            }
        } else {                                             // states gathered out of a pool (search trees)
            for (int i = tid; i < m * SW; i += T) {
                const int slot = i / SW, w = i - slot * SW;
                s_state[slot * SWP + w] = __ldg(a.state_in + static_cast<int64_t>(__ldg(a.state_index + base + slot)) * SW + w);
            }
        }
        if (tid < 2) s_len[tid] = 0;
        __syncthreads();

        // ---- phase A: apply the action (R5, R6 checked against the mover's boards), scores, round-1 candidate ----
        uint32_t *st = s_state + tid * SWP;
        uint32_t *row = s_mask + tid * kSRS;
        bool active = false;                       // evaluates a candidate in round 1
        if (tid < m) {
            uint32_t meta = st[kMeta];
            const bool was_done = (meta >> 4) & 1u;
            const int mover = meta & 15u;
            const int act = a.action != nullptr ? __ldg(a.action + base + tid) : BLK_ACTION_NONE;
            uint32_t ctl = 0u;
            bool moved = false;
            if (act != BLK_ACTION_NONE) {
                bool legal = !was_done && act >= 0 && act < kSA;
                if (legal) {
                    int lo = 0, hi = kOrients - 1;
                    while (lo < hi) {
                        const int mid = (lo + hi + 1) >> 1;
                        if (s_obase[mid] <= act) lo = mid; else hi = mid - 1;
                    }
                    const uint32_t oi = s_oinfo[lo];
                    const int piece = oi & 31, w = (oi >> 12) & 15, ncells = (oi >> 16) & 15;
                    const int W = N + 1 - w;
                    const int rem = act - s_obase[lo];
                    const int ay = rem / W, ax = rem - ay * W;
                    const uint64_t pm = s_ocells[lo] << (8 * ay + ax);
                    uint64_t occ = 0ull, own = 0ull;
#pragma unroll
                    for (int q = 0; q < P; ++q) {
                        const uint64_t b = rows_to_board(st + q * N);
                        occ |= b;
                        if (q == mover) own = b;
                    }
                    const uint32_t inv = st[P * N + mover];
                    uint64_t fr, dg;
                    small_prep<P>(own, occ, inv == kFullInv, mover, fr, dg);
                    legal = ((inv >> piece) & 1u) && (pm & ~fr) == 0ull && (pm & dg) != 0ull;
                    if (legal) {
                        board_to_rows(own | pm, st + mover * N);
                        st[P * N + mover] = inv & ~(1u << piece);
                        uint32_t &sc = st[kMeta + 2 + (mover >> 1)];
                        const int sh = 16 * (mover & 1);
                        sc = (sc & ~(0xffffu << sh)) | (((((sc >> sh) & 0xffffu) + ncells) & 0xffffu) << sh);
                        meta = (meta & ~(1u << (8 + mover))) | ((piece == 0 ? 1u : 0u) << (8 + mover));
                        meta += 1u << 16;
                        st[kMeta] = meta;
                        moved = true;
                    }
                }
                if (!legal) ctl |= kCtlIllegal;
            }
#pragma unroll
            for (int q = 0; q < P; ++q) s_fs[4 * tid + q] = static_cast<int16_t>(small_score(st, q, P, sp.g.score_rule));
            s_cnt[tid] = 0;
            if (was_done) {
                ctl |= kCtlWasDone | kCtlEnded;
                for (int k = 0; k < kSMW; ++k) row[k] = 0u;          // finished game: empty mask
            } else {
                const int cand = moved ? mover : (mover == 0 ? P - 1 : mover - 1);
                ctl |= static_cast<uint32_t>(cand) | (static_cast<uint32_t>(moved ? P : 1) << 4) | (moved ? kCtlMoved : 0u);
                active = true;
            }
            s_ctl[tid] = ctl;
        }

        // ---- rounds: evaluate the next candidate of every slot that still needs a mover ----
        int cur = 0, len = 0;
        for (int round = 0;; ++round) {
            int slot = -1;
            if (round == 0) slot = active ? tid : -1;
            else if (tid < len) slot = s_list[cur * T + tid];
            const int nxt = cur ^ 1;
            if (slot >= 0) {
                uint32_t *s2 = s_state + slot * SWP;
                uint32_t *r2 = s_mask + slot * kSRS;
                uint32_t ctl = s_ctl[slot];
                int cand = ctl & 15u, tries = (ctl >> 4) & 15u;
                uint64_t brd[P], occ = 0ull;
#pragma unroll
                for (int q = 0; q < P; ++q) { brd[q] = rows_to_board(s2 + q * N); occ |= brd[q]; }
                uint32_t inv;
                uint64_t fr, dg;
                for (;;) {
                    cand = (cand + 1 == P) ? 0 : cand + 1;
                    uint64_t own = brd[0];
#pragma unroll
                    for (int q = 1; q < P; ++q) if (q == cand) own = brd[q];
                    inv = s2[P * N + cand];
                    small_prep<P>(own, occ, inv == kFullInv, cand, fr, dg);
                    // every placement covers a cell that is free AND a diagonal contact (or the start corner): without
                    // such a cell the player is stuck and the full evaluation is skipped (it would find nothing);
                    // the last candidate is always evaluated, which also leaves the empty mask behind
                    if ((fr & dg) != 0ull || tries == 1) break;
                    --tries;
                }
                const int cnt = small_eval(fr, dg, inv, r2, s_ck + 4 * slot);
                --tries;
                if (cnt > 0) {
                    s2[kMeta] = (s2[kMeta] & ~15u) | static_cast<uint32_t>(cand);
                    s_cnt[slot] = cnt;
                } else if (tries > 0) {
                    s_list[nxt * T + atomicAdd(&s_len[nxt], 1)] = static_cast<uint16_t>(slot);
                } else if (ctl & kCtlMoved) {
                    ctl |= kCtlEnded;                              // nobody can move: the game is over (R9)
                    if (a.options & BLK_OPT_AUTO_RESET) {
                        const uint32_t game = s2[kMeta + 1] + 1u;
                        for (int w = 0; w < P * N; ++w) s2[w] = 0u;
#pragma unroll
                        for (int q = 0; q < P; ++q) s2[P * N + q] = kFullInv;
                        s2[kMeta] = 0u; s2[kMeta + 1] = game; s2[kMeta + 2] = 0u; s2[kMeta + 3] = 0u;
                        ctl |= kCtlFresh;                          // a fresh board's mask is a constant: the sampler and the
                        s_cnt[slot] = sp.first_count;               // write-back read s_first instead of this slot's row
                    } else {
                        s2[kMeta] |= 1u << 4;                      // done; the mover stays the last mover
                    }
                }                                                  // else: mask-only call on a stuck mover: empty mask, not ended
                s_ctl[slot] = (ctl & ~0xffu) | static_cast<uint32_t>(cand) | (static_cast<uint32_t>(tries) << 4);
            }
            __syncthreads();
            len = s_len[nxt];
            if (tid == 0) s_len[cur] = 0;                          // the list just consumed becomes the next round's target
            __syncthreads();
            if (len == 0) break;
            cur = nxt;
        }

        // ---- per-env outputs: sampler, counts, flags, terminal vector, scores ----
        if (tid < m) {
            const int64_t env = base + tid;
            const uint32_t ctl = s_ctl[tid];
            const int cnt = s_cnt[tid];
            const bool ended = (ctl & kCtlEnded) != 0u;
            const uint32_t *mrow = (ctl & kCtlFresh) ? s_first : row;
            const uint16_t *mck = (ctl & kCtlFresh) ? s_fck : s_ck + 4 * tid;
            if (a.legal_count != nullptr) a.legal_count[env] = cnt;
            if (kSample) {
                int pick = -1;
                if (cnt > 0) {
                    const uint32_t meta = st[kMeta];
                    const uint32_t ply = meta >> 16;
                    const uint32_t u = philox_word(philox4(ply >> 2, st[kMeta + 1], 0u, 0u, static_cast<uint32_t>(a.seed),
                                                           static_cast<uint32_t>(a.seed >> 32) ^ (a.env_id_base + static_cast<uint32_t>(env))), ply);
                    int k = static_cast<int>(__umulhi(u, static_cast<uint32_t>(cnt)));
                    int w = 0;
#pragma unroll
                    for (int j = kSCk - 1; j >= 0; --j) {                      // first level: which run of 16 words
                        const int cj = mck[j];
                        if (w == 0 && k >= cj) { w = 16 * (j + 1); k -= cj; }
                    }
                    uint32_t word = mrow[w];
                    for (;;) {
                        const int c = __popc(word);
                        if (k < c) break;
                        k -= c;
                        word = mrow[++w];
                    }
                    pick = (w << 5) + kth_set_bit(word, k);
                }
                a.next_action[env] = pick;
            }
            if (a.flags != nullptr)
                a.flags[env] = static_cast<uint8_t>((ended ? BLK_FLAG_DONE : 0) | ((ctl & kCtlIllegal) ? BLK_FLAG_ILLEGAL : 0) |
                                                    ((kFmt == 4 && cnt > a.mask_stride) ? BLK_FLAG_TRUNCATED : 0));
            int fs[4], best = -32768, nbest = 0;
#pragma unroll
            for (int q = 0; q < P; ++q) {
                fs[q] = s_fs[4 * tid + q];
                if (fs[q] > best) { best = fs[q]; nbest = 1; } else if (fs[q] == best) ++nbest;
            }
#pragma unroll
            for (int q = 0; q < P; ++q) {
                if (a.scores != nullptr) a.scores[env * P + q] = static_cast<int16_t>(fs[q]);
                if (a.terminal != nullptr)
                    a.terminal[env * P + q] = !ended ? 0.f : (fs[q] == best ? (nbest == 1 ? 3.f : 1.f) : -1.f);
            }
        }

        // ---- cooperative, coalesced write-back of states and masks ----
        if (a.state_out != nullptr)
            for (int i = tid; i < m * SW; i += T) {
                const int slot = i / SW, w = i - slot * SW;
                a.state_out[base * SW + i] = s_state[slot * SWP + w];
            }
        // one warp per env row, lanes across the row: no index divisions, 128 B (bits) / 512 B (bytes) per warp store
        const int warp = tid >> 5, lane = tid & 31;
        if (a.obs != nullptr) {
            // canonical_board of the resulting state (R13): planes 0..P-1 occupancy, planes P..2P-1 one-hot mover
            constexpr int NN = N * N, kObs = 2 * P * NN;
            for (int slot = warp; slot < m; slot += T / 32) {
                const uint32_t *s2 = s_state + slot * SWP;
                float *ob = a.obs + (base + slot) * kObs;
                const int mover = s2[kMeta] & 15u;
#pragma unroll 4
                for (int j = lane; j < kObs; j += 32) {
                    const int plane = j / NN, c = j - plane * NN, y = c / N, x = c - y * N;
                    const bool on = plane < P ? ((s2[plane * N + y] >> x) & 1u) != 0u : (plane - P == mover);
                    ob[j] = on ? 1.f : 0.f;
                }
            }
        }
        if (kFmt == 1) {
            constexpr int kMwPad = ((kSA + 31) / 32 + 3) & ~3;      // Geometry::mw
            for (int slot = warp; slot < m; slot += T / 32) {
                uint2 *out = reinterpret_cast<uint2 *>(reinterpret_cast<uint32_t *>(a.mask) + (base + slot) * a.mask_stride);
                const uint32_t *r = (s_ctl[slot] & kCtlFresh) ? s_first : s_mask + slot * kSRS;
#pragma unroll
                for (int w = lane; w < kMwPad / 2; w += 32)          // rows are 16 B aligned (mask_words is a multiple of 4)
                    out[w] = make_uint2(2 * w < kSMW ? r[2 * w] : 0u, 2 * w + 1 < kSMW ? r[2 * w + 1] : 0u);
            }
        } else if (kFmt == 2) {
            constexpr int kChunks = (kSA + BLK_ROW_ALIGN - 1) / BLK_ROW_ALIGN * BLK_ROW_ALIGN / 16;   // Geometry::mask_bytes / 16
            const unsigned char *lutb = reinterpret_cast<const unsigned char *>(s_lut);
            for (int slot = warp; slot < m; slot += T / 32) {
                unsigned char *out = reinterpret_cast<unsigned char *>(a.mask) + (base + slot) * a.mask_stride;
                const uint32_t *r = (s_ctl[slot] & kCtlFresh) ? s_first : s_mask + slot * kSRS;
#pragma unroll
                for (int c = lane; c < kChunks; c += 32) {
                    const int w = c >> 1;
                    const uint32_t bits = w < kSMW ? (r[w] >> (16 * (c & 1))) & 0xffffu : 0u;
                    const uint2 lo = *reinterpret_cast<const uint2 *>(lutb + ((bits << 3) & 0x7f8u));
                    const uint2 hi = *reinterpret_cast<const uint2 *>(lutb + ((bits >> 5) & 0x7f8u));
                    BLK_STORE16(reinterpret_cast<uint4 *>(out + 16 * c), make_uint4(lo.x, lo.y, hi.x, hi.y));
                }
            }
        } else if (kFmt == 3) {
            // caller buffer with any base / row stride (e.g. a contiguous bool [n, A]): byte stores up to the first 16 B
            // boundary and after the last one, chunk-aligned 16 B stores in between (the 16 mask bits of a chunk are
            // cut out of two neighbouring words with a funnel shift); bytes of neighbouring rows are never touched
            const unsigned char *lutb = reinterpret_cast<const unsigned char *>(s_lut);
            for (int slot = warp; slot < m; slot += T / 32) {
                unsigned char *urow = reinterpret_cast<unsigned char *>(a.mask) + (base + slot) * a.mask_stride;
                const uint32_t *r = (s_ctl[slot] & kCtlFresh) ? s_first : s_mask + slot * kSRS;
                const int head = static_cast<int>((16u - (reinterpret_cast<uintptr_t>(urow) & 15u)) & 15u);
                if (lane < head) urow[lane] = static_cast<unsigned char>((r[0] >> lane) & 1u);
                const int nchunks = (kSA - head) >> 4;
                for (int c = lane; c < nchunks; c += 32) {
                    const int bit0 = head + 16 * c, w = bit0 >> 5;
                    const uint32_t bits = __funnelshift_r(r[w], w + 1 < kSMW ? r[w + 1] : 0u, bit0 & 31) & 0xffffu;
                    const uint2 lo = *reinterpret_cast<const uint2 *>(lutb + ((bits << 3) & 0x7f8u));
                    const uint2 hi = *reinterpret_cast<const uint2 *>(lutb + ((bits >> 5) & 0x7f8u));
                    BLK_STORE16(reinterpret_cast<uint4 *>(urow + bit0), make_uint4(lo.x, lo.y, hi.x, hi.y));
                }
                const int b = head + 16 * nchunks + lane;
                if (b < kSA) urow[b] = static_cast<unsigned char>((r[b >> 5] >> (b & 31)) & 1u);
            }
        } else if (kFmt == 4) {
            // sparse form: the ascending legal ids as uint16 (ppo/trainer.py:385 reads these lists); lane l's word precedes
            // lane l + 1's, so an exclusive scan of the per-lane popcounts gives every lane its slots
            for (int slot = warp; slot < m; slot += T / 32) {
                uint16_t *irow = reinterpret_cast<uint16_t *>(a.mask) + (base + slot) * a.mask_stride;
                const uint32_t *r = (s_ctl[slot] & kCtlFresh) ? s_first : s_mask + slot * kSRS;
                int ibase = 0;
                for (int w0 = 0; w0 < kSMW; w0 += 32) {
                    const int w = w0 + lane;
                    uint32_t word = w < kSMW ? r[w] : 0u;
                    const int pc = __popc(word);
                    const int incl = warp_incl_scan(pc, lane);
                    int pos = ibase + incl - pc;
                    while (word) {
                        const int id = (w << 5) + __ffs(word) - 1;
                        word &= word - 1;
                        if (pos < a.mask_stride) irow[pos] = static_cast<uint16_t>(id);
                        ++pos;
                    }
                    ibase += __shfl_sync(kAllLanes, incl, 31);
                }
            }
        }
        __syncthreads();
    }
}


// ---------------------------------------------------------------------------------------------
// playouts on small boards: one THREAD plays one game to the end (or to `stop_player`'s turn); state in registers,
// the mover's mask words in the thread's shared-memory row (the k-th legal action is looked up there).  Same Philox
// stream, same "k-th legal action in ascending id order" rule as rollout_kernel, so the two produce identical games.
// ---------------------------------------------------------------------------------------------
constexpr int kRT = 256;           // playouts per block

template <int kN, int kP>
__global__ void __launch_bounds__(kRT, 2) small_rollout_kernel(const SmallRollParams rp) {
    static_assert(kN == kSN, "one translation unit per board size");
    constexpr int N = kSN, P = kP;
    constexpr int SW = P * N + P + 4, kMeta = P * N + P;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *s_mask = reinterpret_cast<uint32_t *>(smem_raw);            // [kRT][kSRS]
    uint16_t *s_ck = reinterpret_cast<uint16_t *>(s_mask + kRT * kSRS);   // [kRT][4]
    int32_t *s_obase = reinterpret_cast<int32_t *>(s_ck + 4 * kRT);       // [92]
    uint32_t *s_oinfo = reinterpret_cast<uint32_t *>(s_obase + 92);       // [92]
    uint64_t *s_ocells = reinterpret_cast<uint64_t *>(s_oinfo + 92);      // [92]
    const blk_rollout_args &a = rp.a;
    const int tid = threadIdx.x;
    for (int i = tid; i < 92; i += kRT) {
        s_obase[i] = reinterpret_cast<const int32_t *>(rp.tables + rp.t.off_obase)[i];
        s_oinfo[i] = i < kOrients ? reinterpret_cast<const uint32_t *>(rp.tables + rp.t.off_oinfo)[i] : 0u;
        s_ocells[i] = i < kOrients ? rp.ocells64[i] : 0ull;
    }
    __syncthreads();
    uint32_t *row = s_mask + tid * kSRS;
    uint16_t *ck = s_ck + 4 * tid;
    const int64_t total = a.n_roots * a.per_root;
    for (int64_t gid = static_cast<int64_t>(blockIdx.x) * kRT + tid; gid < total; gid += static_cast<int64_t>(gridDim.x) * kRT) {
        const int64_t root = gid / a.per_root;
        const uint32_t *src = a.roots + root * SW;
        uint64_t own[P];
        uint32_t inv[P];
#pragma unroll
        for (int q = 0; q < P; ++q) {
            uint32_t rows[N];
#pragma unroll
            for (int y = 0; y < N; ++y) rows[y] = __ldg(src + q * N + y);
            own[q] = rows_to_board(rows);
            inv[q] = __ldg(src + P * N + q);
        }
        uint32_t meta = __ldg(src + kMeta), game = __ldg(src + kMeta + 1);
        uint32_t sc01 = __ldg(src + kMeta + 2), sc23 = __ldg(src + kMeta + 3);
        const uint32_t key0 = static_cast<uint32_t>(a.seed);
        const uint32_t key1 = static_cast<uint32_t>(a.seed >> 32) ^ (a.rollout_id_base + static_cast<uint32_t>(gid));
        int nply = 0, rnd_block = -1;
        uint4 rnd = make_uint4(0u, 0u, 0u, 0u);
        bool over = (meta >> 4) & 1u;
        int cand = static_cast<int>(meta & 15u);
        cand = cand == 0 ? P - 1 : cand - 1;          // the root's mover is evaluated first
        int tries = 1;
        uint32_t stuck = 0u;
        while (!over && nply <= 21 * P) {          // (bounded: an unreachable input state cannot spin the kernel)
            cand = (cand + 1 == P) ? 0 : cand + 1;
            int cnt = 0;
            if (!((stuck >> cand) & 1u)) {
                uint64_t occ = 0ull, mine = own[0];
                uint32_t invc = inv[0];
#pragma unroll
                for (int q = 0; q < P; ++q) {
                    occ |= own[q];
                    if (q == cand) { mine = own[q]; invc = inv[q]; }
                }
                uint64_t fr, dg;
                small_prep<P>(mine, occ, invc == kFullInv, cand, fr, dg);
                if ((fr & dg) != 0ull) cnt = small_eval(fr, dg, invc, row, ck);
                // a player without a move never gets one back: it is not evaluated again in this playout
                if (cnt == 0) stuck |= 1u << cand;
            }
            if (cnt == 0) {
                if (--tries > 0) continue;
                meta |= 1u << 4;
                over = true;
                break;
            }
            meta = (meta & ~15u) | static_cast<uint32_t>(cand);
            if (cand == a.stop_player) break;              // caller's turn: hand the state back
            const uint32_t ply = meta >> 16;
            if (static_cast<int>(ply >> 2) != rnd_block) {
                rnd = philox4(ply >> 2, game, 1u, 0u, key0, key1);
                rnd_block = static_cast<int>(ply >> 2);
            }
            int k = static_cast<int>(__umulhi(philox_word(rnd, ply), static_cast<uint32_t>(cnt)));
            int w = 0;
#pragma unroll
            for (int j = kSCk - 1; j >= 0; --j) {
                const int cj = ck[j];
                if (w == 0 && k >= cj) { w = 16 * (j + 1); k -= cj; }
            }
            uint32_t word = row[w];
            for (;;) {
                const int c = __popc(word);
                if (k < c) break;
                k -= c;
                word = row[++w];
            }
            const int act = (w << 5) + kth_set_bit(word, k);
            if (a.action_log != nullptr && nply < a.log_stride - 1) a.action_log[gid * a.log_stride + nply] = static_cast<uint16_t>(act);
            // decode and place (the action is legal by construction)
            int lo = 0, hi = kOrients - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (s_obase[mid] <= act) lo = mid; else hi = mid - 1;
            }
            const uint32_t oi = s_oinfo[lo];
            const int piece = oi & 31, wd = (oi >> 12) & 15, ncells = (oi >> 16) & 15;
            const int W = N + 1 - wd;
            const int rem = act - s_obase[lo];
            const int ay = rem / W, ax = rem - ay * W;
            const uint64_t pm = s_ocells[lo] << (8 * ay + ax);
#pragma unroll
            for (int q = 0; q < P; ++q)
                if (q == cand) { own[q] |= pm; inv[q] &= ~(1u << piece); }
            uint32_t &sc = (cand < 2) ? sc01 : sc23;
            const int sh = 16 * (cand & 1);
            sc = (sc & ~(0xffffu << sh)) | (((((sc >> sh) & 0xffffu) + ncells) & 0xffffu) << sh);
            meta = (meta & ~(1u << (8 + cand))) | ((piece == 0 ? 1u : 0u) << (8 + cand));
            meta += 1u << 16;
            ++nply;
            tries = P;
        }
        // results: final scores (R10), winners, 3 / 1 / -1 vector (blokus_wrapper.py:177-185)
        int fs[P], best = -32768, nbest = 0;
#pragma unroll
        for (int q = 0; q < P; ++q) {
            const uint32_t packed = q < 2 ? sc01 : sc23;
            int v = static_cast<int>(static_cast<int16_t>((packed >> (16 * (q & 1))) & 0xffffu));
            if (rp.g.score_rule == 1 && inv[q] == 0u) v += 15 + (((meta >> (8 + q)) & 1u) ? 5 : 0);
            fs[q] = v;
            if (v > best) { best = v; nbest = 1; } else if (v == best) ++nbest;
        }
        uint32_t win = 0u;
#pragma unroll
        for (int q = 0; q < P; ++q) {
            const bool mine = fs[q] == best;
            if (mine) win |= 1u << q;
            if (a.final_scores != nullptr) a.final_scores[gid * P + q] = static_cast<int16_t>(fs[q]);
            if (a.value_sum != nullptr && over) atomicAdd(a.value_sum + root * P + q, mine ? (nbest == 1 ? 3.f : 1.f) : -1.f);
        }
        if (a.state_out != nullptr) {
            uint32_t *dst = a.state_out + gid * SW;
#pragma unroll
            for (int q = 0; q < P; ++q) {
                uint32_t rows[N];
                board_to_rows(own[q], rows);
#pragma unroll
                for (int y = 0; y < N; ++y) dst[q * N + y] = rows[y];
                dst[P * N + q] = inv[q];
            }
            dst[kMeta] = meta; dst[kMeta + 1] = game; dst[kMeta + 2] = sc01; dst[kMeta + 3] = sc23;
        }
        if (a.winners != nullptr) a.winners[gid] = over ? static_cast<uint8_t>(win) : 0;
        if (a.plies != nullptr) a.plies[gid] = nply;
        if (a.action_log != nullptr) a.action_log[gid * a.log_stride + min(nply, a.log_stride - 1)] = 0xFFFFu;
    }
}

static int small_roll_smem_bytes() { return 4 * kRT * kSRS + 8 * kRT + 4 * 92 + 4 * 92 + 8 * 92 + 16; }

static int small_smem_bytes(int P) {
    const int T = P == 2 ? kST2 : kST4;
    const int SWP = (P * kSN + P + 4) | 1;
    return 4 * T * kSRS + 4 * T * SWP + 4 * T + 4 * T + 8 * T + 8 * T + 4 * T + 16 + 2048 + 4 * 92 + 4 * 92 + 8 * 92 + 4 * kSMW + 8 + 16;
}

#define BLK_SCAT_(n) kernels_small_##n
#define BLK_SCAT(n) BLK_SCAT_(n)
SmallKernelSet BLK_SCAT(BLK_SMALL_N)() {
    SmallKernelSet k;
    k.num_actions = kSA;
    k.smem[0] = small_smem_bytes(2); k.smem[1] = small_smem_bytes(4);
    k.threads[0] = kST2; k.threads[1] = kST4;
    k.step[0][0][0] = small_step_kernel<kSN, 2, 0, false>; k.step[0][0][1] = small_step_kernel<kSN, 2, 0, true>;
    k.step[0][1][0] = small_step_kernel<kSN, 2, 1, false>; k.step[0][1][1] = small_step_kernel<kSN, 2, 1, true>;
    k.step[0][2][0] = small_step_kernel<kSN, 2, 2, false>; k.step[0][2][1] = small_step_kernel<kSN, 2, 2, true>;
    k.step[1][0][0] = small_step_kernel<kSN, 4, 0, false>; k.step[1][0][1] = small_step_kernel<kSN, 4, 0, true>;
    k.step[1][1][0] = small_step_kernel<kSN, 4, 1, false>; k.step[1][1][1] = small_step_kernel<kSN, 4, 1, true>;
    k.step[1][2][0] = small_step_kernel<kSN, 4, 2, false>; k.step[1][2][1] = small_step_kernel<kSN, 4, 2, true>;
    k.step[0][3][0] = small_step_kernel<kSN, 2, 3, false>; k.step[0][3][1] = small_step_kernel<kSN, 2, 3, true>;
    k.step[1][3][0] = small_step_kernel<kSN, 4, 3, false>; k.step[1][3][1] = small_step_kernel<kSN, 4, 3, true>;
    k.step[0][4][0] = small_step_kernel<kSN, 2, 4, false>; k.step[0][4][1] = small_step_kernel<kSN, 2, 4, true>;
    k.step[1][4][0] = small_step_kernel<kSN, 4, 4, false>; k.step[1][4][1] = small_step_kernel<kSN, 4, 4, true>;
    k.rollout[0] = small_rollout_kernel<kSN, 2>; k.rollout[1] = small_rollout_kernel<kSN, 4>;
    k.roll_smem = small_roll_smem_bytes(); k.roll_threads = kRT;
    return k;
}

}  // namespace blk
