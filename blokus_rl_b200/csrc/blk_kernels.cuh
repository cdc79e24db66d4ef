// blk_kernels.cuh -- device side of the B200 (sm_100a) Blokus engine: state/field helpers, the step / legal-mask
// kernel and the rollout kernel, as templates over <board size, players, mask format, sampler>.
//
// Mapping: ONE WARP PER ENV.  Lane y holds row y of every player's bitboard (bit x = column x), so the whole board
// state lives in 4 registers per lane.  Legality is evaluated bit-parallel across the anchor columns: lane = anchor
// row, rows lane..lane+4 of the "free and not edge-adjacent to own colour" and "diagonal contact / start corner"
// boards are kept pre-shifted in registers, and each of the 91 piece orientations is 4-5 LOP3s with immediate cell
// offsets (blk_orient.inc, generated).  The resulting (orientation, anchor-row) *fields* are staged in shared
// memory, re-assembled into the action-id-ordered bit mask by a table-driven gather (tables staged once per block
// with a 1-D TMA bulk copy) and streamed to HBM as 128-bit stores (bit-packed or one byte per action).
//
// Every (N, P) specialisation is instantiated in its own translation unit (blk_inst.cu compiled with
// -DBLK_INST_N / -DBLK_INST_P); blk_engine.cu holds the host side and the C ABI.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/blokus_b200.h"

namespace blk {

// ---- compile-time knobs; the defaults are the measured optima on B200 ----
constexpr int kMaxN = 20;
constexpr int kPieces = 21;
constexpr int kOrients = 91;
constexpr int kSumH = 246;  // sum of bounding-box heights over the 91 orientations
#ifndef BLK_WARPS
#define BLK_WARPS 8
#define BLK_MIN_BLOCKS 3
#endif
constexpr int kWarps = BLK_WARPS;   // warps (= envs in flight) per block of the step kernel
#ifndef BLK_ROLL_WARPS
#define BLK_ROLL_WARPS 7
#define BLK_ROLL_MIN_BLOCKS 4
#endif
constexpr int kRollWarps = BLK_ROLL_WARPS;   // playouts in flight per block of the rollout kernel (slim tables: 4 blocks/SM)
constexpr uint32_t kFullInv = (1u << kPieces) - 1u;
constexpr uint32_t kAllLanes = 0xffffffffu;
#ifndef BLK_ST_POLICY
#define BLK_ST_POLICY 1
#endif
#if BLK_ST_POLICY == 0
#define BLK_STORE16(p, v) (*(p) = (v))
#elif BLK_ST_POLICY == 1
#define BLK_STORE16(p, v) __stcs((p), (v))
#elif BLK_ST_POLICY == 2
#define BLK_STORE16(p, v) __stwt((p), (v))
#else
#define BLK_STORE16(p, v) __stcg((p), (v))
#endif
#ifndef BLK_GUARDED_FIELD_STORES
#define BLK_GUARDED_FIELD_STORES 0   // 1: bounds-tested field stores in eval_fields (the round-1 form); 0: unconditional, in program order
#endif
#ifndef BLK_EMIT_UNROLL
#define BLK_EMIT_UNROLL 6
#endif
#ifndef BLK_ROW_ALIGN
#define BLK_ROW_ALIGN 128   // byte-mask rows start on 128 B lines: every 512 B warp store is line-aligned (+3.5 % measured)
#endif
constexpr int kQueueSlots = 4096;   // work-queue counter pairs: one per stream / per graph capture that ever launched on the engine
constexpr int kEmitUnroll = BLK_EMIT_UNROLL;   // passes of the emit loop unrolled together (ILP vs I-cache)
constexpr int kOffLut = 0, kOffWdesc = 2048;   // fixed offsets inside the table blob (see TableLayout)

// Device-resident constant tables, one blob, staged into shared memory by TMA (offsets in bytes).
struct TableLayout {
    int off_obase;   // int32[92]   first action id of each orientation (+ sentinel A)
    int off_oinfo;   // uint32[92]  piece | h<<8 | w<<12 | ncells<<16
    int off_ocells;  // uint32[92]  footprint as 5 row patterns of 5 bits: bits [5r, 5r+5) = columns of row r
    int off_foff;    // uint16[nf+1] bit offset (= first action id) of each field, sentinel 0xFFFF
    int off_wsrc;    // uint16[mw]  first field intersecting mask word g
    int off_fbase;   // uint16[92]  first field index of each orientation
    int off_f2o;     // uint8[nf]   orientation of each field
    int off_obslut;  // float4[16]  4 occupancy bits -> 4 floats (fused observation output of the step kernel)
    int bytes;       // multiple of 16
    int roll_begin;  // the rollout kernel stages only [roll_begin, bytes): oinfo, ocells, foff, fbase, f2o
    // Two tables sit at FIXED offsets so their shared-memory addresses are immediates in the unrolled emit loop:
    //   kOffLut   = 0     uint2[256]        byte -> 8 bytes of 0/1 (bit i -> byte i)
    //   kOffWdesc = 2048  uint2[32*rounds]  gather descriptor of mask word g (valid when <= 3 or <= 5 fields meet a word):
    //                     .x = byte offset of the first field in the staging area
    //                          (Geometry::gather == 5: | left shift of field 3 << 16 | left shift of field 4 << 24)
    //                     .y = right shift of field 0 | left shift of field 1 << 8 | left shift of field 2 << 16
};

struct Geometry {
    int N, P, A, nf, mw, mask_bytes, sw, score_rule;
    uint32_t full;      // (1 << N) - 1
    int warp_smem;      // bytes of per-warp scratch (fields + per-word popcounts)
    int fld_words;      // >= nf + 3; words nf..nf+2 stay zero (gather padding)
    int rounds;         // ceil(mw / 32): warp-wide passes over the mask words
    int gather;         // 3 / 5: every mask word gathers from at most that many fields (3 at N = 20, 5 at N = 12..14), so the
                        // branch-free descriptors apply; 0: the generic loop
};

// The geometries with their own instantiation carry these as immediates (checked against the tables at blk_create).
template <int kN> __device__ __forceinline__ int geo_nf(const Geometry &g) { return kN == 20 ? 1665 : kN == 14 ? 1119 : g.nf; }
template <int kN> __device__ __forceinline__ int geo_fld_words(const Geometry &g) { return kN == 20 ? 1668 : kN == 14 ? 1152 : g.fld_words; }
template <int kN> __device__ __forceinline__ int geo_mw(const Geometry &g) { return kN == 20 ? 952 : kN == 14 ? 432 : g.mw; }
template <int kN> __device__ __forceinline__ int geo_rounds(const Geometry &g) { return kN == 20 ? 30 : kN == 14 ? 14 : g.rounds; }
template <int kN> __device__ __forceinline__ int geo_gather(const Geometry &g) { return kN == 20 ? 3 : kN == 14 ? 5 : g.gather; }
// contiguous fields per lane when the legal set is read straight from the fields: 52 (+1 for lane 31) at N = 20,
// 36 at N = 14 (32 x 36 = 1152 = the padded staging area), otherwise ceil(nf / 32)
template <int kN> __device__ __forceinline__ int fields_per_lane(int nf) { return kN == 20 ? 53 : kN == 14 ? 36 : (nf + 31) >> 5; }

struct KParams {
    blk_step_args a;
    const unsigned char *tables;
    TableLayout t;
    Geometry g;
    unsigned long long *queue;   // [0] next env ticket, [1] blocks finished (self-resetting, one slot per launch in flight)
};

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// One elected thread issues a 1-D TMA bulk copy global -> shared and every thread waits on the mbarrier.
__device__ __forceinline__ void tma_load_tables(unsigned char *dst, const unsigned char *src, int bytes,
                                                uint64_t *bar) {
    const uint32_t bar_a = smem_u32(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                smem_u32(dst)),
            "l"(src), "r"(bytes), "r"(bar_a)
            : "memory");
    }
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar_a)
            : "memory");
    }
}

__device__ __forceinline__ uint32_t sel4(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, int q) {
    return q == 0 ? a0 : (q == 1 ? a1 : (q == 2 ? a2 : a3));
}

// Philox-4x32-10 (Salmon et al. SC'11); identical to oracle/blokus_oracle.c:orc_philox.
__device__ __forceinline__ uint4 philox4(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
// The sampler's draw for a state at `ply`: one Philox block serves four consecutive plies
// (counter = (ply >> 2, game, stream, 0), word ply & 3), so a rollout recomputes it every fourth ply only.
__device__ __forceinline__ uint32_t philox_word(const uint4 &b, uint32_t ply) {
    const uint32_t i = ply & 3u;
    return i == 0 ? b.x : (i == 1 ? b.y : (i == 2 ? b.z : b.w));
}

// k-th (0-based) set bit of w; requires popc(w) > k.  (__fns is emulated with ~500 instructions.)
__device__ __forceinline__ int kth_set_bit(uint32_t w, int k) {
    int pos = 0;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const int c = __popc((w >> pos) & ((1u << s) - 1u));
        if (k >= c) { k -= c; pos += s; }
    }
    return pos;
}
// The same for a WARP-UNIFORM (w, k), all lanes calling: lane i tests bit i (a binary search is 5 dependent rounds of
// shift / mask / popc / compare / select; this is one round plus a ballot).
__device__ __forceinline__ int kth_set_bit_warp(uint32_t w, int k, int lane) {
    const bool hit = ((w >> lane) & 1u) && __popc(w & ((1u << lane) - 1u)) == k;
    return __ffs(__ballot_sync(kAllLanes, hit)) - 1;
}

// PTX shl clamps shift amounts above 31 to 32 (result 0); C's << would be undefined there.
__device__ __forceinline__ uint32_t shl_clamp(uint32_t v, uint32_t sh) {
    uint32_t r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(sh));
    return r;
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(kAllLanes, v, d);
    return v;
}

// Board dimensions as seen by the device helpers.  In the <N, P> specialisations (20x20, 14x14, 7x7) N, P and `full`
// are compile-time constants after inlining; the <0, 0> kernels take them from Geometry at run time.
struct Dims {
    int N, P, A, score_rule;
    uint32_t full;
};
template <int kN, int kP>
__device__ __forceinline__ Dims make_dims(const Geometry &g) {
    Dims d;
    d.N = kN ? kN : g.N;
    d.P = kP ? kP : g.P;
    d.A = g.A;
    d.score_rule = g.score_rule;
    d.full = kN ? ((1u << kN) - 1u) : g.full;
    return d;
}

// Dynamic work distribution: warps draw env indices from a global ticket counter instead of a fixed stride, so
// nobody idles in the last wave (65,536 envs over 3,552 resident warps is 18.45 each) and uneven envs
// (skipped players, finished games) even out.  The last block to finish resets the counters for the next launch.
// ticket_issue / ticket_get split the fetch: the atomic is issued early (lane 0 keeps the result) and broadcast only
// where the index is needed, so its round trip to L2 overlaps a whole env instead of stalling the warp.
__device__ __forceinline__ unsigned long long ticket_issue(unsigned long long *queue, int lane) {
    return lane == 0 ? atomicAdd(queue, 1ULL) : 0ULL;
}
__device__ __forceinline__ int64_t ticket_get(unsigned long long t) {
    return static_cast<int64_t>(__shfl_sync(kAllLanes, t, 0));
}
__device__ __forceinline__ int64_t next_ticket(unsigned long long *queue, int lane) {
    unsigned long long t = 0;
    if (lane == 0) t = atomicAdd(queue, 1ULL);
    return static_cast<int64_t>(__shfl_sync(kAllLanes, t, 0));
}
__device__ __forceinline__ void queue_release(unsigned long long *queue) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long done = atomicAdd(queue + 1, 1ULL);
        if (done == gridDim.x - 1) { queue[0] = 0ULL; queue[1] = 0ULL; }
    }
}

// Per-warp view of one env held in registers (lane y = board row y).
struct EnvRegs {
    uint32_t own0, own1, own2, own3;   // this lane's row of each player's bitboard
    uint32_t inv0, inv1, inv2, inv3;   // inventories (warp-uniform)
    uint32_t sc01, sc23;               // int16 x 4 placed-squares scores (warp-uniform)
    uint32_t meta, game;               // warp-uniform
};

// this lane's share of one env state as it sits in HBM: its row of each bitboard + one tail word
struct EnvRaw {
    uint32_t r0, r1, r2, r3, tail;
};

__device__ __forceinline__ EnvRaw env_fetch(const uint32_t *s, const Dims &g, int lane) {
    const int N = g.N, P = g.P;
    const bool in = lane < N;
    EnvRaw w;
    w.r0 = in ? __ldg(s + lane) : 0u;
    w.r1 = in ? __ldg(s + N + lane) : 0u;
    w.r2 = (in && P > 2) ? __ldg(s + 2 * N + lane) : 0u;
    w.r3 = (in && P > 2) ? __ldg(s + 3 * N + lane) : 0u;
    w.tail = lane < P + 4 ? __ldg(s + P * N + lane) : 0u;
    return w;
}

__device__ __forceinline__ void env_unpack(EnvRegs &e, const EnvRaw &w, const Dims &g) {
    const int P = g.P;
    e.own0 = w.r0; e.own1 = w.r1; e.own2 = w.r2; e.own3 = w.r3;
    e.inv0 = __shfl_sync(kAllLanes, w.tail, 0);
    e.inv1 = __shfl_sync(kAllLanes, w.tail, 1);
    e.inv2 = P > 2 ? __shfl_sync(kAllLanes, w.tail, 2) : 0u;
    e.inv3 = P > 2 ? __shfl_sync(kAllLanes, w.tail, 3) : 0u;
    e.meta = __shfl_sync(kAllLanes, w.tail, P);
    e.game = __shfl_sync(kAllLanes, w.tail, P + 1);
    e.sc01 = __shfl_sync(kAllLanes, w.tail, P + 2);
    e.sc23 = __shfl_sync(kAllLanes, w.tail, P + 3);
}

__device__ __forceinline__ void env_load(EnvRegs &e, const uint32_t *s, const Dims &g, int lane) {
    env_unpack(e, env_fetch(s, g, lane), g);
}

__device__ __forceinline__ void env_store(const EnvRegs &e, uint32_t *s, const Dims &g, int lane) {
    const int N = g.N, P = g.P;
    if (lane < N) {
        s[lane] = e.own0;
        s[N + lane] = e.own1;
        if (P > 2) {
            s[2 * N + lane] = e.own2;
            s[3 * N + lane] = e.own3;
        }
    }
    if (lane < P + 4) {
        const int t = lane - P;
        uint32_t v = sel4(e.inv0, e.inv1, e.inv2, e.inv3, lane);
        if (t == 0) v = e.meta;
        if (t == 1) v = e.game;
        if (t == 2) v = e.sc01;
        if (t == 3) v = e.sc23;
        s[P * N + lane] = v;
    }
}

__device__ __forceinline__ void env_fresh(EnvRegs &e, const Dims &g, uint32_t game) {
    e.own0 = e.own1 = e.own2 = e.own3 = 0u;
    e.inv0 = e.inv1 = kFullInv;
    e.inv2 = e.inv3 = g.P > 2 ? kFullInv : 0u;
    e.sc01 = e.sc23 = 0u;
    e.meta = 0u;
    e.game = game;
}

// rows of player q needed for legality: this lane's row of "free" and "diag/corner" boards
__device__ __forceinline__ void prep_rows(const EnvRegs &e, int q, const Dims &g, int lane, uint32_t &fr0,
                                          uint32_t &dg0) {
    const uint32_t o = sel4(e.own0, e.own1, e.own2, e.own3, q);
    const uint32_t occ = e.own0 | e.own1 | e.own2 | e.own3;
    uint32_t up = __shfl_up_sync(kAllLanes, o, 1);
    uint32_t dn = __shfl_down_sync(kAllLanes, o, 1);
    if (lane == 0) up = 0u;
    if (lane == 31) dn = 0u;
    const uint32_t ud = up | dn;
    const uint32_t adj = ud | (o << 1) | (o >> 1);
    const bool first = sel4(e.inv0, e.inv1, e.inv2, e.inv3, q) == kFullInv;
    const int n1 = g.N - 1;
    const int cy = (g.P == 2) ? (q ? n1 : 0) : ((q & 2) ? n1 : 0);
    const int cx = (g.P == 2) ? (q ? n1 : 0) : ((q & 1) ? n1 : 0);
    const bool in = lane < g.N;
    fr0 = in ? (~(occ | adj) & g.full) : 0u;
    const uint32_t diag = ((ud << 1) | (ud >> 1)) & g.full;
    dg0 = in ? (first ? (lane == cy ? (1u << cx) : 0u) : diag) : 0u;
}

// final-rule score of player q (R10); lastmono/inv decide the optional bonus
__device__ __forceinline__ int final_score(const EnvRegs &e, int q, const Dims &g) {
    const uint32_t packed = (q < 2) ? e.sc01 : e.sc23;
    int s = static_cast<int>(static_cast<int16_t>((packed >> (16 * (q & 1))) & 0xffffu));
    if (g.score_rule == 1 && sel4(e.inv0, e.inv1, e.inv2, e.inv3, q) == 0u)
        s += 15 + (((e.meta >> (8 + q)) & 1u) ? 5 : 0);
    return s;
}

// All 91 orientations against rows lane..lane+4 of the free / diagonal boards.  Stages one field per
// (orientation, anchor row = lane) at fld[o*(N+1) - hsum(o) + lane] and returns this lane's OR of them.
// Every (dy, dx) a 5-cell piece can reach satisfies dy + dx <= 4: 15 shifted copies of each board.
template <bool kStage, bool kAny = true>
__device__ __forceinline__ uint32_t eval_fields(uint32_t fr0, uint32_t dg0, uint32_t invc, uint32_t *fld, int N,
                                                int lane) {
    uint32_t fs[5][5], ds[5][5];
#pragma unroll
    for (int r = 0; r < 5; ++r) {  // lanes >= N hold 0 and N <= 20, so out-of-range source lanes read 0
        const uint32_t f = r ? __shfl_down_sync(kAllLanes, fr0, r) : fr0;
        const uint32_t d = r ? __shfl_down_sync(kAllLanes, dg0, r) : dg0;
#pragma unroll
        for (int x = 0; x < 5; ++x) {
            fs[r][x] = (r + x <= 4) ? (f >> x) : 0u;
            ds[r][x] = (r + x <= 4) ? (d >> x) : 0u;
        }
    }
    // No bounds test on the stores below.  A lane whose anchor row lets the footprint stick out at the bottom
    // (lane > N - h) ANDs in a row >= N, which is 0, so it computes f_ = 0 -- and its slot fld[o * (N + 1) - hsum + lane]
    // lies behind orientation o's own rows: in the slots of the FOLLOWING orientations, whose own lanes store there later
    // (used pieces store their zeros just the same), or, behind the last orientation, in the zero padding / the 32
    // scratch words after the fields.  A warp's shared-memory stores execute in program order, and `volatile` keeps the
    // compiler from reordering them, so the later, real value always wins.  With a bounds test per store the compiler
    // turned the last orientation of every piece into a divergent branch: a compare plus a BSSY / BSYNC / BRA triple per
    // piece, 90+ warp instructions per evaluation.
#if BLK_GUARDED_FIELD_STORES
    uint32_t *fldp = fld + lane;
    bool ok[6];
#pragma unroll
    for (int h = 1; h <= 5; ++h) ok[h] = lane <= N - h;
#define BLK_FLD_ST(idx, v, h) if (ok[h]) fldp[idx] = (v)
#else
    volatile uint32_t *fldp = fld + lane;
#define BLK_FLD_ST(idx, v, h) fldp[idx] = (v)
#endif
    // the inventory is the same in every lane, but only a value that comes out of a warp-wide reduction is KNOWN to be
    // uniform: REDUX leaves it in a uniform register, and the 21 per-piece branches below become plain uniform branches
    // without a reconvergence pair (BSSY / BSYNC) around each
    invc = __reduce_or_sync(kAllLanes, invc);
    uint32_t anyacc = 0u;
    const int np1 = N + 1;
    // AND / OR over the cells of five base trominoes, shared by the 75 footprints that contain one of them
    uint32_t base_and[5], base_or[5];
#define BLK_BASE(b, y0, x0, y1, x1, y2, x2)                  \
    base_and[b] = fs[y0][x0] & fs[y1][x1] & fs[y2][x2];       \
    base_or[b] = ds[y0][x0] | ds[y1][x1] | ds[y2][x2];
#define BLK_ORIENT_B(o, p, h, w, n, hsum, b, y0, x0, y1, x1)                                            \
    {                                                                                                  \
        const uint32_t f_ = (base_and[b] & fs[y0][x0] & fs[y1][x1]) & (base_or[b] | ds[y0][x0] | ds[y1][x1]); \
        if (kAny) anyacc |= f_;                                                                        \
        if (kStage) BLK_FLD_ST((o) * np1 - (hsum), f_, h);                                             \
    }
#define BLK_PIECE_BEGIN(p) if ((invc >> (p)) & 1u) {
#define BLK_ORIENT(o, p, h, w, n, hsum, y0, x0, y1, x1, y2, x2, y3, x3, y4, x4)                        \
    {                                                                                                  \
        const uint32_t f_ = (fs[y0][x0] & fs[y1][x1] & fs[y2][x2] & fs[y3][x3] & fs[y4][x4]) &         \
                            (ds[y0][x0] | ds[y1][x1] | ds[y2][x2] | ds[y3][x3] | ds[y4][x4]);         \
        if (kAny) anyacc |= f_;                                                                        \
        if (kStage) BLK_FLD_ST((o) * np1 - (hsum), f_, h);                                             \
    }
#define BLK_PIECE_ELSE(p) \
    }                     \
    else if (kStage) {
#define BLK_ZERO(o, h, hsum) BLK_FLD_ST((o) * np1 - (hsum), 0u, h);
#define BLK_PIECE_END(p) }
#include "blk_orient.inc"
#undef BLK_BASE
#undef BLK_ORIENT_B
#undef BLK_PIECE_BEGIN
#undef BLK_ORIENT
#undef BLK_PIECE_ELSE
#undef BLK_ZERO
#undef BLK_PIECE_END
#undef BLK_FLD_ST
    return anyacc;
}

// mask word g (bits 32g..32g+31 of the action-id-ordered mask) gathered from the staged fields: generic form
__device__ __forceinline__ uint32_t assemble_word(int g, const uint32_t *fld, const uint16_t *foff,
                                                  const uint16_t *wsrc) {
    uint32_t word = 0u;
    int i = wsrc[g];
    const int bit0 = g << 5;
    while (true) {
        const int off = static_cast<int>(foff[i]) - bit0;
        if (off >= 32) break;
        const uint32_t v = fld[i];
        word |= off >= 0 ? (v << off) : (v >> (-off));
        ++i;
    }
    return word;
}

// ... and the branch-free form used when at most three fields meet a word (N = 20: fields are 16-20 bits wide)
__device__ __forceinline__ uint32_t assemble_word3(int g, const uint32_t *fld, const uint2 *wdesc) {
    const uint2 d = wdesc[g];
    const uint32_t *p = reinterpret_cast<const uint32_t *>(reinterpret_cast<const unsigned char *>(fld) + d.x);
    return __funnelshift_r(p[0], 0u, d.y) | shl_clamp(p[1], __byte_perm(d.y, 0u, 0x4441u)) |
           shl_clamp(p[2], __byte_perm(d.y, 0u, 0x4442u));
}

// ... and its five-field form (N = 12..14: fields are 8-14 bits wide); the two extra shifts ride in the upper half of .x
__device__ __forceinline__ uint32_t assemble_word5(int g, const uint32_t *fld, const uint2 *wdesc) {
    const uint2 d = wdesc[g];
    const uint32_t *p = reinterpret_cast<const uint32_t *>(reinterpret_cast<const unsigned char *>(fld) + (d.x & 0xffffu));
    return __funnelshift_r(p[0], 0u, d.y) | shl_clamp(p[1], __byte_perm(d.y, 0u, 0x4441u)) |
           shl_clamp(p[2], __byte_perm(d.y, 0u, 0x4442u)) | shl_clamp(p[3], __byte_perm(d.x, 0u, 0x4442u)) |
           shl_clamp(p[4], __byte_perm(d.x, 0u, 0x4443u));
}

struct SmemTables {
    const int32_t *obase;
    const uint32_t *oinfo;
    const uint32_t *ocells;
    const uint16_t *foff;
    const uint16_t *wsrc;
    const uint16_t *fbase;
    const uint8_t *f2o;
    const uint2 *wdesc;
    const uint2 *lut;
    const float4 *obslut;
};
__device__ __forceinline__ SmemTables make_tables(const unsigned char *tab, const TableLayout &t, int shift = 0) {
    tab -= shift;   // `tab` holds the blob from byte `shift` on (rollout kernel): tables before it must not be touched
    SmemTables tb;
    tb.obase = reinterpret_cast<const int32_t *>(tab + t.off_obase);
    tb.oinfo = reinterpret_cast<const uint32_t *>(tab + t.off_oinfo);
    tb.ocells = reinterpret_cast<const uint32_t *>(tab + t.off_ocells);
    tb.foff = reinterpret_cast<const uint16_t *>(tab + t.off_foff);
    tb.wsrc = reinterpret_cast<const uint16_t *>(tab + t.off_wsrc);
    tb.fbase = reinterpret_cast<const uint16_t *>(tab + t.off_fbase);
    tb.f2o = tab + t.off_f2o;
    tb.wdesc = reinterpret_cast<const uint2 *>(tab + kOffWdesc);
    tb.lut = reinterpret_cast<const uint2 *>(tab + kOffLut);
    tb.obslut = reinterpret_cast<const float4 *>(tab + t.off_obslut);
    return tb;
}

// decode an action id into this lane's row bits of the footprint; returns false when out of range
__device__ __forceinline__ bool decode_action(int act, const SmemTables &tb, const Dims &g, int lane,
                                              uint32_t &pm, int &piece, int &ncells) {
    pm = 0u; piece = 0; ncells = 0;
    if (act < 0 || act >= g.A) return false;
    int o = -1;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        const int oo = lane + 32 * t;
        const bool hit = oo < kOrients && tb.obase[oo] <= act && act < tb.obase[oo + 1];
        const uint32_t b = __ballot_sync(kAllLanes, hit);
        if (b) o = 32 * t + __ffs(b) - 1;
    }
    const uint32_t oi = tb.oinfo[o];
    const uint32_t cells = tb.ocells[o];
    piece = oi & 31;
    const int w = (oi >> 12) & 15;
    ncells = (oi >> 16) & 15;
    const int W = g.N + 1 - w;
    const int rem = act - tb.obase[o];
    const int ay = rem / W;
    const int ax = rem - ay * W;
    const int d = lane - ay;                                   // this lane's row of the footprint, if any
    if (d >= 0 && d < 5) pm = ((cells >> (5 * d)) & 31u) << ax;
    return true;
}

// the same from a (field, bit) pair, as the rollout sampler finds it: no search, no division
__device__ __forceinline__ void decode_field(int fsel, int bit, const SmemTables &tb, int lane, uint32_t &pm,
                                             int &piece, int &ncells) {
    const int o = tb.f2o[fsel];
    const int ay = fsel - static_cast<int>(tb.fbase[o]);
    const uint32_t oi = tb.oinfo[o], cells = tb.ocells[o];
    piece = oi & 31;
    ncells = (oi >> 16) & 15;
    const int d = lane - ay;
    pm = (d >= 0 && d < 5) ? ((cells >> (5 * d)) & 31u) << bit : 0u;
}

__device__ __forceinline__ void apply_placement(EnvRegs &e, int p, uint32_t pm, int piece, int ncells) {
    const uint32_t clr = ~(1u << piece);
    if (p == 0) { e.own0 |= pm; e.inv0 &= clr; }
    if (p == 1) { e.own1 |= pm; e.inv1 &= clr; }
    if (p == 2) { e.own2 |= pm; e.inv2 &= clr; }
    if (p == 3) { e.own3 |= pm; e.inv3 &= clr; }
    uint32_t &sc = (p < 2) ? e.sc01 : e.sc23;
    const int sh = 16 * (p & 1);
    const uint32_t cur = (sc >> sh) & 0xffffu;
    sc = (sc & ~(0xffffu << sh)) | (((cur + ncells) & 0xffffu) << sh);
    uint32_t m = e.meta;
    m = (m & ~(1u << (8 + p))) | ((piece == 0 ? 1u : 0u) << (8 + p));   // lastmono
    m += 1u << 16;                                                         // ply
    e.meta = m;
}

// winners bitmask + value of lane q (<P): 3 sole winner, 1 tied winner, -1 otherwise (blokus_wrapper.py:177-185)
__device__ __forceinline__ float terminal_value(const EnvRegs &e, const Dims &g, int lane, int &my_score) {
    my_score = lane < g.P ? final_score(e, lane, g) : -32768;
    int best = my_score;
#pragma unroll
    for (int d = 1; d < 4; d <<= 1) best = max(best, __shfl_xor_sync(kAllLanes, best, d));
    const uint32_t win = __ballot_sync(kAllLanes, lane < g.P && my_score == best);
    const bool mine = (win >> lane) & 1u;
    return mine ? (__popc(win) == 1 ? 3.f : 1.f) : -1.f;
}

// canonical_board of the env held in `e` (blokus_wrapper.py:144-146; R13): planes 0..P-1 = occupancy of player p,
// planes P..2P-1 = all ones for the side to move.  Lane y owns row y, so a plane is streamed by fetching the
// row of each output element with one shuffle; at N = 20 four cells go through a 16-entry float4 LUT and every
// warp store is 512 contiguous bytes (st.global.cs.v4), otherwise one float per lane (128 B per warp store).
template <int kN, int kP>
__device__ __forceinline__ void emit_observation(const EnvRegs &e, float *__restrict__ ob, const float4 *obslut,
                                                 const Dims &g, int lane) {
    const int N = g.N, P = g.P;
    const int mover = static_cast<int>(e.meta & 15u);
    if (kN == 20) {
        float4 *o4 = reinterpret_cast<float4 *>(ob);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            if (p < P) {
                const uint32_t own = p == 0 ? e.own0 : (p == 1 ? e.own1 : (p == 2 ? e.own2 : e.own3));
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int w = it * 32 + lane;              // quad inside the plane: 100 quads of 4 cells
                    const int y = w / 5, x0 = (w - 5 * y) * 4;
                    const uint32_t bits = __shfl_sync(kAllLanes, own, y & 31);
                    if (w < 100) __stcs(o4 + p * 100 + w, obslut[(bits >> x0) & 15u]);
                }
            }
        }
        const float4 ones = obslut[15], zeros = obslut[0];
        for (int j = lane; j < P * 100; j += 32) __stcs(o4 + P * 100 + j, (j / 100 == mover) ? ones : zeros);
    } else {
        const int nn = N * N;
        for (int p = 0; p < P; ++p) {
            const uint32_t own = sel4(e.own0, e.own1, e.own2, e.own3, p);
            for (int c0 = 0; c0 < nn; c0 += 32) {
                const int c = c0 + lane;
                const int y = c / N, x = c - y * N;
                const uint32_t bits = __shfl_sync(kAllLanes, own, y & 31);
                if (c < nn) ob[p * nn + c] = ((bits >> x) & 1u) ? 1.f : 0.f;
            }
        }
        for (int j = lane; j < P * nn; j += 32) ob[P * nn + j] = (j / nn == mover) ? 1.f : 0.f;
    }
}

// ---- the legal set read straight from the staged fields (field order is action-id order) ----
// Every lane owns a contiguous chunk of fields (52 at N = 20, lane 31 one more); returns this lane's number of legal actions.
template <int kN>
__device__ __forceinline__ int count_field_chunk(const uint32_t *fld, int nf, int per, int lane) {
    int mine = 0;
    if (kN == 20) {                 // 1665 fields = 32 x 52 (+1 for lane 31): 13 conflict-free LDS.128 per lane
        const uint4 *f4 = reinterpret_cast<const uint4 *>(fld) + 13 * lane;
#pragma unroll
        for (int j = 0; j < 13; ++j) {
            const uint4 x = f4[j];
            mine += __popc(x.x) + __popc(x.y) + __popc(x.z) + __popc(x.w);
        }
        if (lane == 31) mine += __popc(fld[1664]);
    } else if (kN == 14) {          // 1119 fields in a zero-padded area of 32 x 36 words: 9 conflict-free LDS.128 per lane
        const uint4 *f4 = reinterpret_cast<const uint4 *>(fld) + 9 * lane;
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            const uint4 x = f4[j];
            mine += __popc(x.x) + __popc(x.y) + __popc(x.z) + __popc(x.w);
        }
    } else {
        for (int j = 0; j < per; ++j) { const int i = lane * per + j; if (i < nf) mine += __popc(fld[i]); }
    }
    return mine;
}
// The k-th legal action (0-based, ascending id) given every lane's chunk count `mine` and their inclusive scan `incl`:
// returns the field, `bit` = the anchor column inside it.  First level: which lane's chunk; second level: which field of
// that chunk (at most 53: two halves of 32 whose popcounts ride through ONE warp scan as a packed 16 + 16 bit pair).
template <int kN>
__device__ __forceinline__ int kth_legal_field(const uint32_t *fld, int k, int mine, int incl, int nf, int per, int lane, int &bit) {
    const int L = __ffs(__ballot_sync(kAllLanes, k < incl)) - 1;
    k -= __shfl_sync(kAllLanes, incl - mine, L);
    const int chunk = kN == 20 ? 52 : per;
    const int chunk_len = kN == 20 ? (L == 31 ? 53 : 52) : per;
    const int i_lo = L * chunk + lane, i_hi = i_lo + 32;
    const int c_lo = (lane < chunk_len && i_lo < nf) ? __popc(fld[i_lo]) : 0;
    const int c_hi = (lane + 32 < chunk_len && i_hi < nf) ? __popc(fld[i_hi]) : 0;
    const int inc2 = warp_incl_scan(c_lo | (c_hi << 16), lane);
    const int tot_lo = __shfl_sync(kAllLanes, inc2, 31) & 0xffff;
    const bool upper = k >= tot_lo;
    if (upper) k -= tot_lo;
    const int c = upper ? c_hi : c_lo;
    const int inc = upper ? (inc2 >> 16) : (inc2 & 0xffff);
    const int J = __ffs(__ballot_sync(kAllLanes, k < inc)) - 1;
    const int kk = k - __shfl_sync(kAllLanes, inc - c, J);
    const int fsel = L * chunk + (upper ? 32 : 0) + J;
    bit = kth_set_bit_warp(fld[fsel], kk, lane);
    return fsel;
}

// ---------------------------------------------------------------------------------------------
// step / legal-mask kernel
// ---------------------------------------------------------------------------------------------
// kFmt: 0 = no mask output, 1 = bit-packed, 2 = bytes through 16 B vector stores, 3 = bytes into an unaligned buffer,
//       4 = ascending list of legal ids (uint16).
template <int kN, int kP, int kFmt, bool kSample>
__global__ void __launch_bounds__(kWarps * 32, BLK_MIN_BLOCKS) step_kernel(const KParams kp) {
    extern __shared__ __align__(128) unsigned char smem[];
    const Geometry &gg = kp.g;
    const Dims g = make_dims<kN, kP>(gg);
    const blk_step_args &a = kp.a;
    unsigned char *tab = smem;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + kp.t.bytes);
    unsigned char *scratch = smem + kp.t.bytes + 16;
    tma_load_tables(tab, kp.tables, kp.t.bytes, bar);
    const SmemTables tb = make_tables(tab, kp.t);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int N = g.N, P = g.P;
    const int sw = P * N + P + 4;
    const int fld_words = geo_fld_words<kN>(gg);
    const int mw = geo_mw<kN>(gg);
    const int rounds = geo_rounds<kN>(gg);
    const int mask_bytes = kN == 20 ? (30433 + BLK_ROW_ALIGN - 1) / BLK_ROW_ALIGN * BLK_ROW_ALIGN
                         : kN == 14 ? (13729 + BLK_ROW_ALIGN - 1) / BLK_ROW_ALIGN * BLK_ROW_ALIGN : gg.mask_bytes;
    const int gather = geo_gather<kN>(gg);
    uint32_t *fld = reinterpret_cast<uint32_t *>(scratch + static_cast<size_t>(warp) * gg.warp_smem);
    // (the 32 words behind the fields are scratch: lanes >= N of the unconditional field stores land there)
    const int64_t n = a.n;
    const int64_t mstride = a.mask_stride;
    const bool want_count = a.legal_count != nullptr;
    const uint32_t rot_lo = (lane & 1) ? 19u : 3u, rot_hi = (lane & 1) ? 11u : 27u;   // 16-bit half -> LUT byte offsets
    const int src_lo = lane >> 1, src_hi = 16 + (lane >> 1);
    const unsigned char *lutb = reinterpret_cast<const unsigned char *>(tb.lut);
    const uint2 *wdl = tb.wdesc + lane;
    for (int i = geo_nf<kN>(gg) + lane; i < fld_words; i += 32) fld[i] = 0u;   // gather padding stays zero

    // software pipeline over this warp's envs: the next env's 352 B and action are fetched while the current one
    // is processed (a warp handles its envs serially; without this every env starts with an exposed HBM round trip)
    int64_t env = next_ticket(kp.queue, lane);
    int64_t env_next = next_ticket(kp.queue, lane);
    unsigned long long pending = 0ULL;
    EnvRaw raw_next = {};
    int act_next = BLK_ACTION_NONE;
    if (env < n) {
        raw_next = env_fetch(a.state_in + (a.state_index != nullptr ? __ldg(a.state_index + env) : env) * sw, g, lane);
        if (a.action != nullptr) act_next = __ldg(a.action + env);
    }
    for (; env < n; env = env_next, env_next = ticket_get(pending)) {
        EnvRegs e;
        env_unpack(e, raw_next, g);
        const int act = act_next;
        if (env_next < n) {
            raw_next = env_fetch(a.state_in + (a.state_index != nullptr ? __ldg(a.state_index + env_next) : env_next) * sw, g, lane);
            if (a.action != nullptr) act_next = __ldg(a.action + env_next);
        }
        pending = ticket_issue(kp.queue, lane);             // broadcast by the loop increment, one env later
        const bool was_done = (e.meta >> 4) & 1u;
        const int mover = e.meta & 15u;
        uint32_t flags = 0u;
        bool moved = false;

        if (act != BLK_ACTION_NONE) {
            uint32_t pm; int piece, ncells;
            bool legal = !was_done && decode_action(act, tb, g, lane, pm, piece, ncells);
            if (legal) {
                uint32_t fr0, dg0;
                prep_rows(e, mover, g, lane, fr0, dg0);
                const bool avail = (sel4(e.inv0, e.inv1, e.inv2, e.inv3, mover) >> piece) & 1u;
                const bool bad = __any_sync(kAllLanes, (pm & ~fr0) != 0u);
                const bool touch = __any_sync(kAllLanes, (pm & dg0) != 0u);
                legal = avail && !bad && touch;
            }
            if (legal) { apply_placement(e, mover, pm, piece, ncells); moved = true; }
            else flags |= BLK_FLAG_ILLEGAL;
        }

        // ---- next mover (R8 auto-skip), terminal detection (R9), optional auto-reset ----
        bool have = false, ended = false;
        float tval = 0.f;
        int fscore = 0;
        if (lane < P) fscore = final_score(e, lane, g);
        if (was_done) {
            ended = true;
            tval = terminal_value(e, g, lane, fscore);
        } else {
            int cand = moved ? mover : (mover == 0 ? P - 1 : mover - 1);
            int tries = moved ? P : 1;
            bool did_reset = false;
#pragma unroll 1
            while (true) {
                cand = (cand + 1 == P) ? 0 : cand + 1;
                uint32_t fr0, dg0;
                prep_rows(e, cand, g, lane, fr0, dg0);
                const uint32_t acc = eval_fields<true>(fr0, dg0, sel4(e.inv0, e.inv1, e.inv2, e.inv3, cand), fld, N, lane);
                if (__any_sync(kAllLanes, acc != 0u)) { have = true; break; }
                if (--tries > 0) continue;
                if (!moved) break;                      // mask-only call on a state whose mover is stuck
                ended = true;                           // nobody can move: the game is over
                tval = terminal_value(e, g, lane, fscore);
                if ((a.options & BLK_OPT_AUTO_RESET) && !did_reset) {
                    env_fresh(e, g, e.game + 1u);
                    did_reset = true; cand = P - 1; tries = 1;
                    continue;
                }
                e.meta |= 1u << 4;                      // done; mover stays = last mover
                break;
            }
            if (have) e.meta = (e.meta & ~15u) | static_cast<uint32_t>(cand);
        }
        if (ended) flags |= BLK_FLAG_DONE;
        if (!have) {                                   // terminal (or stuck) state: empty mask
            for (int i = lane; i < fld_words; i += 32) fld[i] = 0u;
        }
        __syncwarp();

        // ---- gather the action-id-ordered mask from the staged fields and stream it out ----
        int cnt = 0;
        int idx_pick = -1;
        int fmine = 0, fincl = 0;                      // this lane's legal count over its chunk of fields, and the warp scan of it
        if (kFmt == 4) {
            // Index list: the ids come straight out of the fields (field order is id order), no mask words are built.
            // Every lane counts its contiguous chunk of fields, a scan gives it its first slot, and it writes its ids there.
            // (The mask-word route -- 30 passes of gather + scan + per-lane bit loops -- cost 4,500 warp instructions per env.)
            const int nf_ = geo_nf<kN>(gg);
            const int per = fields_per_lane<kN>(nf_);
            const int mine = count_field_chunk<kN>(fld, nf_, per, lane);
            const int incl = warp_incl_scan(mine, lane);
            cnt = __shfl_sync(kAllLanes, incl, 31);
            uint16_t *irow = reinterpret_cast<uint16_t *>(a.mask) + env * mstride;
            int64_t room = mstride;
            if (a.csr_cursor != nullptr) {
                // compact form: this env's ids go behind whatever the envs before it (in launch order) asked for
                unsigned long long base = 0ULL;
                if (lane == 0) {
                    base = atomicAdd(a.csr_cursor, static_cast<unsigned long long>(cnt));
                    a.csr_offset[env] = static_cast<int64_t>(base);
                }
                base = __shfl_sync(kAllLanes, base, 0);
                irow = reinterpret_cast<uint16_t *>(a.mask) + base;
                room = static_cast<int64_t>(base) + cnt <= mstride ? cnt : 0;      // all of the env's ids or none
            }
            int pos = incl - mine;
            const int chunk = kN == 20 ? 52 : per;
            const int i0 = lane * chunk;
            const int len = kN == 20 ? (lane == 31 ? 53 : 52) : per;
            if (mine > 0) {
                for (int t = 0; t < len; ++t) {
                    const int i = i0 + t;
                    uint32_t w = i < nf_ ? fld[i] : 0u;
                    if (w) {
                        const int base = tb.foff[i];
                        do {
                            if (pos < room) irow[pos] = static_cast<uint16_t>(base + __ffs(w) - 1);
                            ++pos;
                            w &= w - 1;
                        } while (w);
                    }
                }
            }
            __syncwarp();
            if (cnt > room) flags |= BLK_FLAG_TRUNCATED;
            if (kSample && cnt > 0) {
                const uint32_t ply = e.meta >> 16;
                const uint32_t u = philox_word(philox4(ply >> 2, e.game, 0u, 0u, static_cast<uint32_t>(a.seed),
                                                       static_cast<uint32_t>(a.seed >> 32) ^ (a.env_id_base + static_cast<uint32_t>(env))), ply);
                int bit;
                const int fsel = kth_legal_field<kN>(fld, static_cast<int>(__umulhi(u, static_cast<uint32_t>(cnt))), mine, incl, nf_, per, lane, bit);
                idx_pick = static_cast<int>(tb.foff[fsel]) + bit;
            }
        } else if (kFmt != 0 || kSample || want_count) {
            // The sampler (and a count without a mask) reads the legal set straight from the fields, like the index-list format:
            // one counting pass + one warp scan instead of POPC + REDUX + STS in every emit pass, and the emit loop below is left
            // without a cross-lane operation (the REDUX latency sat on its critical path).
            constexpr bool kFromFields = kSample || kFmt == 0;
            if (kFromFields) {
                const int nf_ = geo_nf<kN>(gg);
                const int per = fields_per_lane<kN>(nf_);
                fmine = count_field_chunk<kN>(fld, nf_, per, lane);
                fincl = warp_incl_scan(fmine, lane);
                cnt = __shfl_sync(kAllLanes, fincl, 31);
            }
            unsigned char *row = reinterpret_cast<unsigned char *>(a.mask) + env * mstride + 16 * lane;
            uint32_t *wrow = reinterpret_cast<uint32_t *>(a.mask) + env * mstride + lane;
            unsigned char *urow = reinterpret_cast<unsigned char *>(a.mask) + env * mstride;     // kFmt == 3 only
            const int ush = static_cast<int>(reinterpret_cast<uintptr_t>(urow) & 15);
            uint32_t uprev = 0u;
#pragma unroll(kEmitUnroll)
            for (int r = 0; r < (kFmt == 0 && kFromFields ? 0 : rounds); ++r) {
                uint32_t word;
                if (gather == 3) word = assemble_word3(r << 5, fld, wdl);
                else if (gather == 5) word = assemble_word5(r << 5, fld, wdl);
                else word = ((r << 5) + lane) < mw ? assemble_word((r << 5) + lane, fld, tb.foff, tb.wsrc) : 0u;
                if (!kFromFields) cnt += __popc(word);
                if (kFmt == 4) {
                    // (handled above, straight from the fields)
                } else if (kFmt == 1) {
                    if (r < rounds - 1 || (r << 5) + lane < mw) wrow[r << 5] = word;
                } else if (kFmt == 2) {
                    // 32 words -> 1024 bytes; each lane expands 16 bits through the byte LUT (two 8-byte entries) and
                    // writes 16 B, so one warp store covers 512 contiguous bytes
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t w2 = __shfl_sync(kAllLanes, word, h ? src_hi : src_lo);
                        const uint2 lo = *reinterpret_cast<const uint2 *>(lutb + (__funnelshift_l(w2, w2, rot_lo) & 0x7f8u));
                        const uint2 hi = *reinterpret_cast<const uint2 *>(lutb + (__funnelshift_l(w2, w2, rot_hi) & 0x7f8u));
                        const int boff = (r << 10) + 512 * h;
                        if (r < rounds - 1 || boff + 16 * lane < mask_bytes)
                            BLK_STORE16(reinterpret_cast<uint4 *>(row + boff), make_uint4(lo.x, lo.y, hi.x, hi.y));
                    }
                } else if (kFmt == 3) {
                    // Caller buffer with any base / row stride (e.g. a contiguous bool [n, A]): the row starts `ush`
                    // bytes into a 16 B chunk.  Still one 16 B store per lane on chunk-aligned addresses: chunk c of
                    // this pass holds row bytes [1024 r + 16 c - ush, +16), i.e. 16 mask bits that straddle at most two
                    // words (the word before this pass comes from `uprev`).  Only the chunks that overlap the row ends
                    // fall back to byte stores, so neighbouring rows (other warps) are never touched.
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int c = 32 * h + lane;
                        const int rb = 16 * c - ush;                          // bit offset inside this pass, >= -15
                        const int wi = rb >> 5;                               // -1 .. 31
                        const uint32_t lo_w = __shfl_sync(kAllLanes, word, wi & 31);
                        const uint32_t hi_w = __shfl_sync(kAllLanes, word, (wi + 1) & 31);
                        const uint32_t bits = __funnelshift_r(wi < 0 ? uprev : lo_w, hi_w, rb & 31) & 0xffffu;
                        const uint2 lo = *reinterpret_cast<const uint2 *>(lutb + ((bits << 3) & 0x7f8u));
                        const uint2 hi = *reinterpret_cast<const uint2 *>(lutb + ((bits >> 5) & 0x7f8u));
                        const int p0 = (r << 10) + rb;                        // row position of the chunk's first byte
                        unsigned char *dst = urow + p0;                       // 16 B aligned by construction
                        if (p0 >= 0 && p0 + 16 <= g.A) {
                            BLK_STORE16(reinterpret_cast<uint4 *>(dst), make_uint4(lo.x, lo.y, hi.x, hi.y));
                        } else if (p0 + 16 > 0 && p0 < g.A) {                 // first / last chunk of the row
                            const uint32_t v[4] = {lo.x, lo.y, hi.x, hi.y};
                            for (int b = 0; b < 16; ++b)
                                if (p0 + b >= 0 && p0 + b < g.A) dst[b] = static_cast<unsigned char>((v[b >> 2] >> (8 * (b & 3))) & 1u);
                        }
                    }
                    uprev = __shfl_sync(kAllLanes, word, 31);
                }
            }
            if (!kFromFields) cnt = warp_sum(cnt);
        }
        if (want_count && lane == 0) a.legal_count[env] = cnt;

        // ---- uniform random legal action for the new mover: k = mulhi(u32, n), k-th set bit ascending ----
        if (kSample) {
            int pick = idx_pick;
            if (kFmt != 4 && cnt > 0) {
                const uint32_t ply = e.meta >> 16;
                const uint32_t u = philox_word(philox4(ply >> 2, e.game, 0u, 0u, static_cast<uint32_t>(a.seed),
                                                       static_cast<uint32_t>(a.seed >> 32) ^ (a.env_id_base + static_cast<uint32_t>(env))), ply);
                const int nf_ = geo_nf<kN>(gg);
                int bit;
                const int fsel = kth_legal_field<kN>(fld, static_cast<int>(__umulhi(u, static_cast<uint32_t>(cnt))), fmine, fincl, nf_,
                                                     fields_per_lane<kN>(nf_), lane, bit);
                pick = static_cast<int>(tb.foff[fsel]) + bit;
            }
            if (lane == 0) a.next_action[env] = pick;
        }

        // ---- per-step outputs and state write-back ----
        if (lane < P) {
            if (a.terminal != nullptr) a.terminal[env * P + lane] = ended ? tval : 0.f;
            if (a.scores != nullptr) a.scores[env * P + lane] = static_cast<int16_t>(fscore);
        }
        if (a.flags != nullptr && lane == 0) a.flags[env] = static_cast<uint8_t>(flags);
        if (a.state_out != nullptr) env_store(e, a.state_out + env * sw, g, lane);
        if (a.obs != nullptr) emit_observation<kN, kP>(e, a.obs + env * (2 * P * N * N), tb.obslut, g, lane);
        __syncwarp();
    }
    queue_release(kp.queue);
}

// ---------------------------------------------------------------------------------------------
// rollout kernel: one warp plays one game to the end, state in registers, fields in shared memory
// ---------------------------------------------------------------------------------------------
struct RParams {
    blk_rollout_args a;
    const unsigned char *tables;
    TableLayout t;
    Geometry g;
    unsigned long long *queue;
};

// One uniform-random playout by one warp: from the state in `e` until the game is over (or until it is `stop_player`'s
// turn).  Sampler: Philox block (ply >> 2, game, stream, 0) under key (key0, key1), word ply & 3, k = mulhi(u, n_legal),
// k-th legal action in ascending id order.  Returns the plies played; `over` says whether the game ended.
template <int kN, int kP>
__device__ __forceinline__ int playout_game(EnvRegs &e, const SmemTables &tb, const Geometry &gg, const Dims &g, uint32_t *fld,
                                            int lane, uint32_t key0, uint32_t key1, uint32_t stream, int stop_player,
                                            uint16_t *log, int log_cap, bool &over) {
    const int N = g.N, P = g.P;
    const int nf = geo_nf<kN>(gg);
    const int per = fields_per_lane<kN>(nf);   // contiguous fields per lane for the k-th-bit search (53 at N = 20)
    int nply = 0;
    uint4 rnd = make_uint4(0u, 0u, 0u, 0u);
    int rnd_block = -1;
    over = (e.meta >> 4) & 1u;
    // the root's mover is evaluated first; afterwards every player gets a try after each placement (R8)
    int cand = static_cast<int>(e.meta & 15u);
    cand = cand == 0 ? P - 1 : cand - 1;
    int tries = 1;
    uint32_t stuck = 0u;
    // (a game has at most 21 placements per player: the bound only matters if the state handed in was not a reachable one,
    // and then it keeps the kernel from spinning)
#pragma unroll 1
    while (!over && nply <= 21 * P) {
        cand = (cand + 1 == P) ? 0 : cand + 1;
        bool has = false;
        int mine = 0, incl = 0;
        if (!((stuck >> cand) & 1u)) {
            uint32_t fr0, dg0;
            prep_rows(e, cand, g, lane, fr0, dg0);
            eval_fields<true, false>(fr0, dg0, sel4(e.inv0, e.inv1, e.inv2, e.inv3, cand), fld, N, lane);
            __syncwarp();
            // count legal actions: lane sums popcounts over its contiguous chunk of fields (field order = id order);
            // "has a move" falls out of the count, so the fields are not OR-reduced separately
            mine = count_field_chunk<kN>(fld, nf, per, lane);
            incl = warp_incl_scan(mine, lane);
            has = __shfl_sync(kAllLanes, incl, 31) > 0;
            // a player without a move never gets one back (others only take cells away, and it places nothing
            // itself), so it is not evaluated again for the rest of the playout
            if (!has) stuck |= 1u << cand;
        }
        if (!has) {
            if (--tries > 0) continue;
            e.meta |= 1u << 4;
            over = true;
            break;
        }
        e.meta = (e.meta & ~15u) | static_cast<uint32_t>(cand);
        if (cand == stop_player) break;              // caller's turn: hand the state back
        const int cnt = __shfl_sync(kAllLanes, incl, 31);
        const uint32_t ply = e.meta >> 16;
        if (static_cast<int>(ply >> 2) != rnd_block) {
            rnd = philox4(ply >> 2, e.game, stream, 0u, key0, key1);
            rnd_block = static_cast<int>(ply >> 2);
        }
        const int k = static_cast<int>(__umulhi(philox_word(rnd, ply), static_cast<uint32_t>(cnt)));
        int bit;
        const int fsel = kth_legal_field<kN>(fld, k, mine, incl, nf, per, lane, bit);
        if (log != nullptr && lane == 0 && nply < log_cap - 1) log[nply] = static_cast<uint16_t>(tb.foff[fsel] + bit);
        uint32_t pm; int piece, ncells;
        decode_field(fsel, bit, tb, lane, pm, piece, ncells);
        apply_placement(e, cand, pm, piece, ncells);
        ++nply;
        tries = P;
        __syncwarp();
    }
    return nply;
}

template <int kN, int kP>
__global__ void __launch_bounds__(kRollWarps * 32, BLK_ROLL_MIN_BLOCKS) rollout_kernel(const RParams rp) {
    extern __shared__ __align__(128) unsigned char smem[];
    const Geometry &gg = rp.g;
    const Dims g = make_dims<kN, kP>(gg);
    const blk_rollout_args &a = rp.a;
    unsigned char *tab = smem;
    const int tab_bytes = rp.t.bytes - rp.t.roll_begin;          // only the tables a playout needs (~6 KB of ~19 KB)
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + tab_bytes);
    unsigned char *scratch = smem + tab_bytes + 16;
    tma_load_tables(tab, rp.tables + rp.t.roll_begin, tab_bytes, bar);
    const SmemTables tb = make_tables(tab, rp.t, rp.t.roll_begin);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t *fld = reinterpret_cast<uint32_t *>(scratch + static_cast<size_t>(warp) * gg.warp_smem);
    const int P = g.P;
    const int sw = P * g.N + P + 4;
    const int64_t total = a.n_roots * a.per_root;
    // the chunked readers of the staged fields (count_field_chunk at N = 14) read the padding behind the last field: it
    // must hold zeros, whatever the previous kernel left in this shared memory
    for (int i = geo_nf<kN>(gg) + lane; i < geo_fld_words<kN>(gg); i += 32) fld[i] = 0u;
    __syncwarp();

    for (int64_t gid = next_ticket(rp.queue, lane); gid < total; gid = next_ticket(rp.queue, lane)) {
        const int64_t root = gid / a.per_root;
        EnvRegs e;
        env_load(e, a.roots + root * sw, g, lane);
        const uint32_t key0 = static_cast<uint32_t>(a.seed);
        const uint32_t key1 = static_cast<uint32_t>(a.seed >> 32) ^ (a.rollout_id_base + static_cast<uint32_t>(gid));
        bool over;
        const int nply = playout_game<kN, kP>(e, tb, gg, g, fld, lane, key0, key1, 1u, a.stop_player,
                                              a.action_log != nullptr ? a.action_log + gid * a.log_stride : nullptr, a.log_stride, over);
        int fscore;
        const float tval = terminal_value(e, g, lane, fscore);
        const uint32_t win = __ballot_sync(kAllLanes, lane < P && tval > 0.f);
        if (lane < P) {
            if (a.final_scores != nullptr) a.final_scores[gid * P + lane] = static_cast<int16_t>(fscore);
            if (a.value_sum != nullptr && over) atomicAdd(a.value_sum + root * P + lane, tval);
        }
        if (a.state_out != nullptr) env_store(e, a.state_out + gid * sw, g, lane);
        if (lane == 0) {
            if (a.winners != nullptr) a.winners[gid] = over ? static_cast<uint8_t>(win) : 0;
            if (a.plies != nullptr) a.plies[gid] = nply;
            if (a.action_log != nullptr) a.action_log[gid * a.log_stride + min(nply, a.log_stride - 1)] = 0xFFFFu;
        }
        __syncwarp();
    }
    queue_release(rp.queue);
}

// ---------------------------------------------------------------------------------------------
// kernel sets: one per (N, P) specialisation, defined by blk_inst.cu
// ---------------------------------------------------------------------------------------------
using StepFn = void (*)(const KParams);
using RolloutFn = void (*)(const RParams);
struct KernelSet {
    StepFn step[5][2];      // [mask format variant: none, bits, bytes (vector stores), bytes (unaligned), ids][sampler]
    RolloutFn rollout;
};
template <int kN, int kP>
inline KernelSet make_kernel_set() {
    KernelSet k;
    k.step[0][0] = step_kernel<kN, kP, 0, false>; k.step[0][1] = step_kernel<kN, kP, 0, true>;
    k.step[1][0] = step_kernel<kN, kP, 1, false>; k.step[1][1] = step_kernel<kN, kP, 1, true>;
    k.step[2][0] = step_kernel<kN, kP, 2, false>; k.step[2][1] = step_kernel<kN, kP, 2, true>;
    k.step[3][0] = step_kernel<kN, kP, 3, false>; k.step[3][1] = step_kernel<kN, kP, 3, true>;
    k.step[4][0] = step_kernel<kN, kP, 4, false>; k.step[4][1] = step_kernel<kN, kP, 4, true>;
    k.rollout = rollout_kernel<kN, kP>;
    return k;
}
// Small boards (N <= 7): thread-per-env kernels on 64-bit bitboards, one translation unit per N (blk_small.cu)
struct SmallParams {
    blk_step_args a;
    const unsigned char *tables;
    TableLayout t;
    Geometry g;
    const uint64_t *ocells64;     // [92] footprint of each orientation, row stride 8
    const uint32_t *first_mask;   // legal mask words of the fresh board (player 0's first move)
    int first_count;
};
struct SmallRollParams {
    blk_rollout_args a;
    const unsigned char *tables;
    TableLayout t;
    Geometry g;
    const uint64_t *ocells64;
};
using SmallStepFn = void (*)(const SmallParams);
using SmallRollFn = void (*)(const SmallRollParams);
struct SmallKernelSet {
    SmallStepFn step[2][5][2];    // [P == 4][mask format: none, bits, bytes (aligned rows), bytes (any rows), ids][sampler]
    SmallRollFn rollout[2];       // [P == 4] thread-per-playout
    int smem[2], threads[2];
    int roll_smem, roll_threads;
    int num_actions;
};
SmallKernelSet kernels_small_5();
SmallKernelSet kernels_small_6();
SmallKernelSet kernels_small_7();
KernelSet kernels_20_4();   // the headline geometry: every address and trip count is a compile-time constant
KernelSet kernels_20_2();
KernelSet kernels_14_4();
KernelSet kernels_14_2();   // Blokus Duo board
KernelSet kernels_7_2();    // the reference's PPO config (config/ppo_blokus_7x7.yml)
KernelSet kernels_0_0();    // runtime dimensions: any 5 <= N <= 20, P in {2, 4}

}  // namespace blk
