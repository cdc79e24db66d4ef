// blk_engine.cu -- B200 (sm_100a) batched Blokus environment engine: kernels + C ABI.
//
// Hot path (reference call sites in blokus_rl/colossumrl/blokus_wrapper.py):
//   a1 new_state      -> reset_kernel            a2 next_state   -> step_kernel<true>
//   a3 valid_actions  -> step_kernel<false/true> a4 get_winners  -> step_kernel / ended_kernel
//   a5 canonical_board-> observe_kernel          a6 board_contents -> contents_kernel
//   a7 action table   -> build_tables() (host)   rollouts (new)  -> rollout_kernel
//
// Mapping: ONE WARP PER ENV.  Lane y holds row y of every player's bitboard (bit x = column x), so
// the whole board state lives in 4 registers per lane.  Legality is evaluated bit-parallel across
// the 20 anchor columns: lane = anchor row, `fr[r]`/`dg[r]` are rows lane..lane+4 of the "free and
// not edge-adjacent to own colour" and "diagonal contact / start corner" boards, and each of the 91
// piece orientations is 4-5 LOP3s with immediate cell offsets (blk_orient.inc, generated).  The
// resulting (orientation, anchor-row) *fields* are staged in shared memory, re-assembled into the
// action-id-ordered bit mask by a table-driven gather (tables staged once per block with a 1-D TMA
// bulk copy), and streamed to HBM as 128-bit stores (bit-packed or one byte per action).
//
// No tensor cores: nothing here is a contraction.  No CPU fallback: every entry point needs the GPU.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/blokus_b200.h"

namespace {

constexpr int kMaxN = 20;
constexpr int kPieces = 21;
constexpr int kOrients = 91;
constexpr int kSumH = 246;  // sum of bounding-box heights over the 91 orientations
#ifndef BLK_WARPS
#define BLK_WARPS 8
#define BLK_MIN_BLOCKS 3
#endif
constexpr int kWarps = BLK_WARPS;   // warps (= envs in flight) per block
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
#ifndef BLK_EMIT_UNROLL
#define BLK_EMIT_UNROLL 6
#endif
#ifndef BLK_ROW_ALIGN
#define BLK_ROW_ALIGN 128   // byte-mask rows start on 128 B lines: every 512 B warp store is line-aligned (+3.5 % measured)
#endif
constexpr int kQueueSlots = 64;   // work-queue counters for up to 64 launches of one engine in flight at once
constexpr int kEmitUnroll = BLK_EMIT_UNROLL;   // passes of the emit loop unrolled together (ILP vs I-cache)
constexpr int kOffLut = 0, kOffWdesc = 2048;   // fixed offsets inside the table blob (see TableLayout)

// ---------------------------------------------------------------------------------------------
// host-side orientation metadata (same generated list the kernels unroll)
// ---------------------------------------------------------------------------------------------
struct OrientRow {
    int8_t piece, local, h, w, n;
    int8_t yx[10];
};
#define BLK_PIECE_BEGIN(p)
#define BLK_ORIENT(...)
#define BLK_PIECE_ELSE(p)
#define BLK_ZERO(...)
#define BLK_PIECE_END(p)
#define BLK_ORIENT_ROW(o, p, l, h, w, n, y0, x0, y1, x1, y2, x2, y3, x3, y4, x4) \
    {p, l, h, w, n, {y0, x0, y1, x1, y2, x2, y3, x3, y4, x4}},
const OrientRow kOrient[kOrients] = {
#include "blk_orient.inc"
};
#undef BLK_PIECE_BEGIN
#undef BLK_ORIENT
#undef BLK_PIECE_ELSE
#undef BLK_ZERO
#undef BLK_PIECE_END
#undef BLK_ORIENT_ROW

thread_local std::string g_err;
int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail(BLK_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));          \
    } while (0)

// Device-resident constant tables, one blob, staged into shared memory by TMA (offsets in bytes).
struct TableLayout {
    int off_obase;   // int32[92]   first action id of each orientation (+ sentinel A)
    int off_oinfo;   // uint32[92]  piece | h<<8 | w<<12 | ncells<<16
    int off_ocells;  // uint32[92]  5 x (dy:3, dx:3)
    int off_foff;    // uint16[nf+1] bit offset (= first action id) of each field, sentinel 0xFFFF
    int off_wsrc;    // uint16[mw]  first field intersecting mask word g
    int off_fbase;   // uint16[92]  first field index of each orientation
    int off_f2o;     // uint8[nf]   orientation of each field
    int bytes;       // multiple of 16
    // Two tables sit at FIXED offsets so their shared-memory addresses are immediates in the unrolled emit loop:
    //   kOffLut   = 0     uint2[256]        byte -> 8 bytes of 0/1 (bit i -> byte i)
    //   kOffWdesc = 2048  uint2[32*rounds]  gather descriptor of mask word g (valid when <= 3 fields meet a word):
    //                     .x = byte offset of the first field in the staging area
    //                     .y = right shift of field 0 | left shift of field 1 << 8 | left shift of field 2 << 16
};

struct Geometry {
    int N, P, A, nf, mw, mask_bytes, sw, score_rule;
    uint32_t full;      // (1 << N) - 1
    int warp_smem;      // bytes of per-warp scratch (fields + per-word popcounts)
    int fld_words;      // >= nf + 3; words nf..nf+2 stay zero (gather padding)
    int rounds;         // ceil(mw / 32): warp-wide passes over the mask words
    int fast3;          // 1 when every mask word gathers from <= 3 fields (true at N = 20)
};

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

// Board dimensions as seen by the device helpers.  The kernels are instantiated for <N=20, P=4> (every member a
// compile-time constant after inlining) and for <0, 0> (runtime values from Geometry).
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
template <bool kStage>
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
    bool ok[6];
#pragma unroll
    for (int h = 1; h <= 5; ++h) ok[h] = lane <= N - h;
    uint32_t anyacc = 0u;
    const int np1 = N + 1;
    uint32_t *fldp = fld + lane;
#define BLK_PIECE_BEGIN(p) if ((invc >> (p)) & 1u) {
#define BLK_ORIENT(o, p, h, w, n, hsum, y0, x0, y1, x1, y2, x2, y3, x3, y4, x4)                        \
    {                                                                                                  \
        const uint32_t f_ = (fs[y0][x0] & fs[y1][x1] & fs[y2][x2] & fs[y3][x3] & fs[y4][x4]) &         \
                            (ds[y0][x0] | ds[y1][x1] | ds[y2][x2] | ds[y3][x3] | ds[y4][x4]);         \
        anyacc |= f_;                                                                                  \
        if (kStage && ok[h]) fldp[(o) * np1 - (hsum)] = f_;                                            \
    }
#define BLK_PIECE_ELSE(p) \
    }                     \
    else if (kStage) {
#define BLK_ZERO(o, h, hsum) \
    if (ok[h]) fldp[(o) * np1 - (hsum)] = 0u;
#define BLK_PIECE_END(p) }
#include "blk_orient.inc"
#undef BLK_PIECE_BEGIN
#undef BLK_ORIENT
#undef BLK_PIECE_ELSE
#undef BLK_ZERO
#undef BLK_PIECE_END
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
    return (p[0] >> (d.y & 31u)) | shl_clamp(p[1], __byte_perm(d.y, 0u, 0x4441u)) |
           shl_clamp(p[2], __byte_perm(d.y, 0u, 0x4442u));
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
};
__device__ __forceinline__ SmemTables make_tables(const unsigned char *tab, const TableLayout &t) {
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
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        const int dy = (cells >> (6 * c)) & 7, dx = (cells >> (6 * c + 3)) & 7;
        if (c < ncells && ay + dy == lane) pm |= 1u << (ax + dx);
    }
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
    pm = 0u;
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        const int dy = (cells >> (6 * c)) & 7, dx = (cells >> (6 * c + 3)) & 7;
        if (c < ncells && ay + dy == lane) pm |= 1u << (bit + dx);
    }
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

// ---------------------------------------------------------------------------------------------
// step / legal-mask kernel
// ---------------------------------------------------------------------------------------------
// kFmt: 0 = no mask output, 1 = bit-packed, 2 = bytes through 16 B vector stores, 3 = bytes into an unaligned buffer.
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
    const int fld_words = kN == 20 ? 1668 : gg.fld_words;
    const int mw = kN == 20 ? 952 : gg.mw;
    const int rounds = kN == 20 ? 30 : gg.rounds;
    const int mask_bytes = kN == 20 ? (30433 + BLK_ROW_ALIGN - 1) / BLK_ROW_ALIGN * BLK_ROW_ALIGN : gg.mask_bytes;
    const bool fast3 = kN == 20 ? true : (gg.fast3 != 0);
    uint32_t *fld = reinterpret_cast<uint32_t *>(scratch + static_cast<size_t>(warp) * gg.warp_smem);
    uint32_t *tots = fld + fld_words;                                  // 32 per-pass popcount totals (sampler)
    const int64_t n = a.n;
    const int64_t mstride = a.mask_stride;
    const bool want_count = a.legal_count != nullptr;
    const uint32_t rot_lo = (lane & 1) ? 19u : 3u, rot_hi = (lane & 1) ? 11u : 27u;   // 16-bit half -> LUT byte offsets
    const int src_lo = lane >> 1, src_hi = 16 + (lane >> 1);
    const unsigned char *lutb = reinterpret_cast<const unsigned char *>(tb.lut);
    const uint2 *wdl = tb.wdesc + lane;
    for (int i = (kN == 20 ? 1665 : gg.nf) + lane; i < fld_words; i += 32) fld[i] = 0u;   // gather padding stays zero

    // software pipeline over this warp's envs: the next env's 352 B and action are fetched while the current one
    // is processed (a warp handles its envs serially; without this every env starts with an exposed HBM round trip)
    int64_t env = next_ticket(kp.queue, lane);
    int64_t env_next = next_ticket(kp.queue, lane), env_after = 0;
    EnvRaw raw_next = {};
    int act_next = BLK_ACTION_NONE;
    if (env < n) {
        raw_next = env_fetch(a.state_in + env * sw, g, lane);
        if (a.action != nullptr) act_next = __ldg(a.action + env);
    }
    for (; env < n; env = env_next, env_next = env_after) {
        EnvRegs e;
        env_unpack(e, raw_next, g);
        const int act = act_next;
        if (env_next < n) {
            raw_next = env_fetch(a.state_in + env_next * sw, g, lane);
            if (a.action != nullptr) act_next = __ldg(a.action + env_next);
        }
        env_after = next_ticket(kp.queue, lane);
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
        if (kFmt != 0 || kSample || want_count) {
            unsigned char *row = reinterpret_cast<unsigned char *>(a.mask) + env * mstride + 16 * lane;
            uint32_t *wrow = reinterpret_cast<uint32_t *>(a.mask) + env * mstride + lane;
#pragma unroll(kEmitUnroll)
            for (int r = 0; r < (kN == 20 ? 30 : rounds); ++r) {
                uint32_t word;
                if (fast3) word = assemble_word3(r << 5, fld, wdl);
                else word = ((r << 5) + lane) < mw ? assemble_word((r << 5) + lane, fld, tb.foff, tb.wsrc) : 0u;
                const int pc = __popc(word);
                if (kSample) {
                    const int tot = __reduce_add_sync(kAllLanes, pc);
                    if (lane == 0) tots[r] = tot;
                    cnt += tot;
                } else {
                    cnt += pc;
                }
                if (kFmt == 1) {
                    if (r < (kN == 20 ? 29 : rounds - 1) || (r << 5) + lane < mw) wrow[r << 5] = word;
                } else if (kFmt == 2) {
                    // 32 words -> 1024 bytes; each lane expands 16 bits through the byte LUT (two 8-byte entries) and
                    // writes 16 B, so one warp store covers 512 contiguous bytes
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t w2 = __shfl_sync(kAllLanes, word, h ? src_hi : src_lo);
                        const uint2 lo = *reinterpret_cast<const uint2 *>(lutb + (__funnelshift_l(w2, w2, rot_lo) & 0x7f8u));
                        const uint2 hi = *reinterpret_cast<const uint2 *>(lutb + (__funnelshift_l(w2, w2, rot_hi) & 0x7f8u));
                        const int boff = (r << 10) + 512 * h;
                        if (r < (kN == 20 ? 29 : rounds - 1) || boff + 16 * lane < mask_bytes)
                            BLK_STORE16(reinterpret_cast<uint4 *>(row + boff), make_uint4(lo.x, lo.y, hi.x, hi.y));
                    }
                } else if (kFmt == 3) {  // unaligned caller buffer: correct but slow byte stores
                    unsigned char *urow = reinterpret_cast<unsigned char *>(a.mask) + env * mstride;
                    for (int b = 0; b < 32; ++b) {
                        const int idx = (((r << 5) + lane) << 5) + b;
                        if (idx < g.A) urow[idx] = static_cast<unsigned char>((word >> b) & 1u);
                    }
                }
            }
            if (!kSample) cnt = warp_sum(cnt);
        }
        if (want_count && lane == 0) a.legal_count[env] = cnt;

        // ---- uniform random legal action for the new mover: k = mulhi(u32, n), k-th set bit ascending ----
        if (kSample) {
            int pick = -1;
            if (cnt > 0) {
                __syncwarp();
                const uint32_t ply = e.meta >> 16;
                const uint32_t u = philox_word(philox4(ply >> 2, e.game, 0u, 0u, static_cast<uint32_t>(a.seed),
                                                       static_cast<uint32_t>(a.seed >> 32) ^ (a.env_id_base + static_cast<uint32_t>(env))), ply);
                int k = static_cast<int>(__umulhi(u, static_cast<uint32_t>(cnt)));
                // level 1: which pass of 32 words
                const int tot = lane < rounds ? static_cast<int>(tots[lane]) : 0;
                const int incl = warp_incl_scan(tot, lane);
                const int R = __ffs(__ballot_sync(kAllLanes, k < incl)) - 1;
                k -= __shfl_sync(kAllLanes, incl - tot, R);
                // level 2: which word of that pass (re-gathered: cheaper than keeping 952 popcounts around)
                const int gi = (R << 5) + lane;
                uint32_t word;
                if (fast3) word = assemble_word3(gi, fld, tb.wdesc);
                else word = gi < mw ? assemble_word(gi, fld, tb.foff, tb.wsrc) : 0u;
                const int c2 = __popc(word);
                const int incl2 = warp_incl_scan(c2, lane);
                const int J = __ffs(__ballot_sync(kAllLanes, k < incl2)) - 1;
                k -= __shfl_sync(kAllLanes, incl2 - c2, J);
                const uint32_t wsel = __shfl_sync(kAllLanes, word, J);
                pick = (((R << 5) + J) << 5) + kth_set_bit(wsel, k);
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
        __syncwarp();
    }
    queue_release(kp.queue);
}

// ---------------------------------------------------------------------------------------------
// small streaming kernels
// ---------------------------------------------------------------------------------------------
__global__ void reset_kernel(uint32_t *state, int64_t n, Geometry g) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n * g.sw) return;
    const int w = static_cast<int>(i % g.sw);
    const int b = g.P * g.N;
    state[i] = (w >= b && w < b + g.P) ? kFullInv : 0u;
}

template <int VEC>
__global__ void observe_kernel(const uint32_t *__restrict__ state, float *__restrict__ obs, int64_t n, Geometry g) {
    // R13: planes 0..P-1 occupancy of player i, planes P..2P-1 all-ones for the mover
    const int nn = g.N * g.N, per_env = 2 * g.P * nn;
    const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t e0 = t * VEC;
    if (e0 >= n * per_env) return;
    const int64_t env = e0 / per_env;
    const int r = static_cast<int>(e0 - env * per_env);
    const int plane = r / nn, cell = r - plane * nn;
    const int y = cell / g.N, x = cell - y * g.N;
    const uint32_t *s = state + env * g.sw;
    float v[VEC];
    if (plane < g.P) {
        // VEC consecutive cells may straddle a row boundary: fetch per element
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            int yy = y, xx = x + k;
            if (xx >= g.N) { xx -= g.N; yy += 1; }
            v[k] = ((__ldg(s + plane * g.N + yy) >> xx) & 1u) ? 1.f : 0.f;
        }
    } else {
        const int mover = __ldg(s + g.P * g.N + g.P) & 15u;
        const float f = (plane - g.P == mover) ? 1.f : 0.f;
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[k] = f;
    }
    if (VEC == 4) {
        __stcs(reinterpret_cast<float4 *>(obs + e0), make_float4(v[0], v[1], v[2], v[3]));
    } else {
#pragma unroll
        for (int k = 0; k < VEC; ++k) obs[e0 + k] = v[k];
    }
}

__global__ void contents_kernel(const uint32_t *__restrict__ state, uint8_t *__restrict__ board, int64_t n, Geometry g) {
    const int nn = g.N * g.N;
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n * nn) return;
    const int64_t env = i / nn;
    const int cell = static_cast<int>(i - env * nn);
    const int y = cell / g.N, x = cell - y * g.N;
    const uint32_t *s = state + env * g.sw;
    uint8_t c = 0;
    for (int q = 0; q < g.P; ++q)
        if ((__ldg(s + q * g.N + y) >> x) & 1u) c = static_cast<uint8_t>(q + 1);
    board[i] = c;
}

__global__ void ended_kernel(const uint32_t *__restrict__ state, uint8_t *flags, float *terminal, int16_t *scores,
                             int64_t n, Geometry g) {
    const int64_t env = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (env >= n) return;
    const uint32_t *s = state + env * g.sw + g.P * g.N;
    const uint32_t meta = s[g.P];
    const bool done = (meta >> 4) & 1u;
    int sc[4], best = -32768, nbest = 0;
    for (int q = 0; q < g.P; ++q) {
        int v = static_cast<int16_t>((s[g.P + 2 + (q >> 1)] >> (16 * (q & 1))) & 0xffffu);
        if (g.score_rule == 1 && s[q] == 0u) v += 15 + (((meta >> (8 + q)) & 1u) ? 5 : 0);
        sc[q] = v;
        if (v > best) { best = v; nbest = 1; } else if (v == best) ++nbest;
    }
    if (flags) flags[env] = done ? BLK_FLAG_DONE : 0;
    for (int q = 0; q < g.P; ++q) {
        if (scores) scores[env * g.P + q] = static_cast<int16_t>(sc[q]);
        if (terminal) terminal[env * g.P + q] = !done ? 0.f : (sc[q] == best ? (nbest == 1 ? 3.f : 1.f) : -1.f);
    }
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

template <int kN, int kP>
__global__ void __launch_bounds__(kWarps * 32, BLK_MIN_BLOCKS) rollout_kernel(const RParams rp) {
    extern __shared__ __align__(128) unsigned char smem[];
    const Geometry &gg = rp.g;
    const Dims g = make_dims<kN, kP>(gg);
    const blk_rollout_args &a = rp.a;
    unsigned char *tab = smem;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + rp.t.bytes);
    unsigned char *scratch = smem + rp.t.bytes + 16;
    tma_load_tables(tab, rp.tables, rp.t.bytes, bar);
    const SmemTables tb = make_tables(tab, rp.t);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t *fld = reinterpret_cast<uint32_t *>(scratch + static_cast<size_t>(warp) * gg.warp_smem);
    const int N = g.N, P = g.P;
    const int sw = P * N + P + 4;
    const int nf = kN == 20 ? 1665 : gg.nf;
    const int64_t total = a.n_roots * a.per_root;
    const int per = (nf + 31) >> 5;   // contiguous fields per lane for the k-th-bit search (53 at N = 20)

    for (int64_t gid = next_ticket(rp.queue, lane); gid < total; gid = next_ticket(rp.queue, lane)) {
        const int64_t root = gid / a.per_root;
        EnvRegs e;
        env_load(e, a.roots + root * sw, g, lane);
        const uint32_t key0 = static_cast<uint32_t>(a.seed);
        const uint32_t key1 = static_cast<uint32_t>(a.seed >> 32) ^ (a.rollout_id_base + static_cast<uint32_t>(gid));
        int nply = 0;
        uint4 rnd = make_uint4(0u, 0u, 0u, 0u);
        int rnd_block = -1;
        bool over = (e.meta >> 4) & 1u;
        // the root's mover is evaluated first; afterwards every player gets a try after each placement (R8)
        int cand = static_cast<int>(e.meta & 15u);
        cand = cand == 0 ? P - 1 : cand - 1;
        int tries = 1;
#pragma unroll 1
        while (!over) {
            cand = (cand + 1 == P) ? 0 : cand + 1;
            uint32_t fr0, dg0;
            prep_rows(e, cand, g, lane, fr0, dg0);
            const uint32_t acc = eval_fields<true>(fr0, dg0, sel4(e.inv0, e.inv1, e.inv2, e.inv3, cand), fld, N, lane);
            if (!__any_sync(kAllLanes, acc != 0u)) {
                if (--tries > 0) continue;
                e.meta |= 1u << 4;
                over = true;
                break;
            }
            __syncwarp();
            e.meta = (e.meta & ~15u) | static_cast<uint32_t>(cand);
            // count legal actions: lane sums popcounts over its contiguous chunk of fields (field order = id order)
            int mine = 0;
            if (kN == 20) {                     // 1665 fields = 32 x 52 (+1 for lane 31): 13 conflict-free LDS.128 per lane
                const uint4 *f4 = reinterpret_cast<const uint4 *>(fld) + 13 * lane;
#pragma unroll
                for (int j = 0; j < 13; ++j) {
                    const uint4 x = f4[j];
                    mine += __popc(x.x) + __popc(x.y) + __popc(x.z) + __popc(x.w);
                }
                if (lane == 31) mine += __popc(fld[1664]);
            } else {
                for (int j = 0; j < per; ++j) { const int i = lane * per + j; if (i < nf) mine += __popc(fld[i]); }
            }
            const int incl = warp_incl_scan(mine, lane);
            const int cnt = __shfl_sync(kAllLanes, incl, 31);
            const uint32_t ply = e.meta >> 16;
            if (static_cast<int>(ply >> 2) != rnd_block) {
                rnd = philox4(ply >> 2, e.game, 1u, 0u, key0, key1);
                rnd_block = static_cast<int>(ply >> 2);
            }
            int k = static_cast<int>(__umulhi(philox_word(rnd, ply), static_cast<uint32_t>(cnt)));
            const int L = __ffs(__ballot_sync(kAllLanes, k < incl)) - 1;
            k -= __shfl_sync(kAllLanes, incl - mine, L);
            // second level: the chunk of lane L, 32 fields at a time
            const int chunk = kN == 20 ? 52 : per;
            const int chunk_len = kN == 20 ? (L == 31 ? 53 : 52) : per;
            int fsel = -1, kk = 0;
            for (int half = 0; half * 32 < chunk_len; ++half) {
                const int j = half * 32 + lane;
                const int i = L * chunk + j;
                const int c = (j < chunk_len && i < nf) ? __popc(fld[i]) : 0;
                const int inc2 = warp_incl_scan(c, lane);
                const uint32_t b = __ballot_sync(kAllLanes, k < inc2);
                if (b) {
                    const int J = __ffs(b) - 1;
                    kk = k - __shfl_sync(kAllLanes, inc2 - c, J);
                    fsel = L * chunk + half * 32 + J;
                    break;
                }
                k -= __shfl_sync(kAllLanes, inc2, 31);
            }
            const int bit = kth_set_bit(fld[fsel], kk);
            if (a.action_log != nullptr && lane == 0 && nply < a.log_stride - 1)
                a.action_log[gid * a.log_stride + nply] = static_cast<uint16_t>(tb.foff[fsel] + bit);
            uint32_t pm; int piece, ncells;
            decode_field(fsel, bit, tb, lane, pm, piece, ncells);
            apply_placement(e, cand, pm, piece, ncells);
            ++nply;
            tries = P;
            __syncwarp();
        }
        int fscore;
        const float tval = terminal_value(e, g, lane, fscore);
        const uint32_t win = __ballot_sync(kAllLanes, lane < P && tval > 0.f);
        if (lane < P) {
            a.final_scores[gid * P + lane] = static_cast<int16_t>(fscore);
            if (a.value_sum != nullptr) atomicAdd(a.value_sum + root * P + lane, tval);
        }
        if (lane == 0) {
            if (a.winners != nullptr) a.winners[gid] = static_cast<uint8_t>(win);
            if (a.plies != nullptr) a.plies[gid] = nply;
            if (a.action_log != nullptr) a.action_log[gid * a.log_stride + min(nply, a.log_stride - 1)] = 0xFFFFu;
        }
        __syncwarp();
    }
    queue_release(rp.queue);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// engine object and C ABI
// ---------------------------------------------------------------------------------------------
struct blk_engine {
    blk_config cfg;
    Geometry g;
    TableLayout t;
    unsigned char *d_tables = nullptr;
    unsigned long long *d_queue = nullptr;   // kQueueSlots x {ticket, finished}; launch i uses slot i % kQueueSlots
    unsigned launch_seq = 0;
    int sm_count = 0;
    int step_smem = 0;
    int step_blocks_per_sm = 0, rollout_blocks_per_sm = 0;
    bool special = false;
    void (*step_fn[4][2])(const KParams) = {};      // [mask format variant][sampler]
    void (*rollout_fn)(const RParams) = nullptr;
    std::vector<int32_t> obase;        // host copies for blk_action_to_cells
    std::vector<int16_t> act_o, act_y, act_x;
};

namespace {

int align16(int x) { return (x + 15) & ~15; }

int build_tables(blk_engine *h) {
    const int N = h->cfg.board_size, P = h->cfg.num_players;
    Geometry &g = h->g;
    g.N = N; g.P = P; g.score_rule = h->cfg.score_rule;
    g.full = (1u << N) - 1u;
    g.sw = P * N + P + 4;
    // action ids: (piece, orientation, anchor y, anchor x); fields: (orientation, anchor y)
    h->obase.assign(kOrients + 1, 0);
    std::vector<uint16_t> foff;
    int A = 0;
    for (int o = 0; o < kOrients; ++o) {
        h->obase[o] = A;
        const int R = N - kOrient[o].h + 1, W = N - kOrient[o].w + 1;
        for (int y = 0; y < R; ++y) {
            foff.push_back(static_cast<uint16_t>(A));
            for (int x = 0; x < W; ++x) {
                h->act_o.push_back(static_cast<int16_t>(o));
                h->act_y.push_back(static_cast<int16_t>(y));
                h->act_x.push_back(static_cast<int16_t>(x));
                ++A;
            }
        }
    }
    h->obase[kOrients] = A;
    if (A >= 0xFFFF - 64) return fail(BLK_ERR_ARG, "action space too large for 16-bit field offsets");
    g.A = A;
    g.nf = static_cast<int>(foff.size());
    if (g.nf != kOrients * (N + 1) - kSumH) return fail(BLK_ERR_ARG, "internal: field count mismatch");
    foff.push_back(0xFFFF);  // sentinel
    g.mw = ((A + 31) / 32 + 3) & ~3;
    g.mask_bytes = (A + BLK_ROW_ALIGN - 1) / BLK_ROW_ALIGN * BLK_ROW_ALIGN;
    std::vector<uint16_t> wsrc(g.mw, static_cast<uint16_t>(g.nf));
    {
        int f = 0;
        for (int w = 0; w < g.mw; ++w) {
            // first field whose bit range [foff[f], foff[f+1]) reaches past bit 32w
            while (f < g.nf && static_cast<int>(foff[f + 1] == 0xFFFF ? A : foff[f + 1]) <= 32 * w) ++f;
            wsrc[w] = static_cast<uint16_t>(f);
        }
    }
    TableLayout &t = h->t;
    g.rounds = (g.mw + 31) / 32;
    int off = kOffWdesc + 8 * 32 * g.rounds;               // LUT at 0, gather descriptors at 2048
    t.off_obase = off;  off = align16(off + 4 * (kOrients + 1));
    t.off_oinfo = off;  off = align16(off + 4 * (kOrients + 1));
    t.off_ocells = off; off = align16(off + 4 * (kOrients + 1));
    t.off_foff = off;   off = align16(off + 2 * (g.nf + 1));
    t.off_wsrc = off;   off = align16(off + 2 * g.mw);
    t.off_fbase = off;  off = align16(off + 2 * (kOrients + 1));
    t.off_f2o = off;    off = align16(off + g.nf);
    t.bytes = off;
    std::vector<unsigned char> blob(off, 0);
    memcpy(blob.data() + t.off_obase, h->obase.data(), 4 * (kOrients + 1));
    for (int o = 0; o < kOrients; ++o) {
        const OrientRow &r = kOrient[o];
        uint32_t info = static_cast<uint32_t>(r.piece) | (static_cast<uint32_t>(r.h) << 8) |
                        (static_cast<uint32_t>(r.w) << 12) | (static_cast<uint32_t>(r.n) << 16);
        uint32_t cells = 0;
        for (int c = 0; c < 5; ++c)
            cells |= (static_cast<uint32_t>(r.yx[2 * c]) | (static_cast<uint32_t>(r.yx[2 * c + 1]) << 3)) << (6 * c);
        memcpy(blob.data() + t.off_oinfo + 4 * o, &info, 4);
        memcpy(blob.data() + t.off_ocells + 4 * o, &cells, 4);
    }
    memcpy(blob.data() + t.off_foff, foff.data(), 2 * foff.size());
    memcpy(blob.data() + t.off_wsrc, wsrc.data(), 2 * wsrc.size());
    {
        int f = 0;
        for (int o = 0; o <= kOrients; ++o) {
            const uint16_t fb = static_cast<uint16_t>(f);
            memcpy(blob.data() + t.off_fbase + 2 * o, &fb, 2);
            if (o == kOrients) break;
            for (int y = 0; y < N - kOrient[o].h + 1; ++y) blob[t.off_f2o + f++] = static_cast<unsigned char>(o);
        }
    }
    // gather descriptors: word g = (fld[s] >> r0) | (fld[s+1] << s1) | (fld[s+2] << s2); shifts >= 32 give 0
    g.fast3 = 1;
    std::vector<uint32_t> wdesc(2 * 32 * g.rounds);
    for (int w = 0; w < 32 * g.rounds; ++w) { wdesc[2 * w] = 4u * g.nf; wdesc[2 * w + 1] = (63u << 8) | (63u << 16); }
    for (int w = 0; w < g.mw && g.fast3; ++w) {
        const int s0 = wsrc[w];
        if (s0 >= g.nf) continue;                       // padding word: gathers the zero slots behind the fields
        auto start = [&](int f) { return f < g.nf ? static_cast<int>(foff[f]) : (1 << 20); };   // past the last field: never
        if (start(s0 + 3) < 32 * w + 32) { g.fast3 = 0; break; }   // a 4th field reaches into this word
        const int r0 = 32 * w - start(s0);
        const int s1 = start(s0 + 1) - 32 * w, s2 = start(s0 + 2) - 32 * w;
        if (r0 < 0 || r0 > 31) { g.fast3 = 0; break; }
        wdesc[2 * w] = 4u * static_cast<uint32_t>(s0);
        wdesc[2 * w + 1] = static_cast<uint32_t>(r0) | (static_cast<uint32_t>(s1 > 63 ? 63 : s1) << 8) |
                           (static_cast<uint32_t>(s2 > 63 ? 63 : s2) << 16);
    }
    memcpy(blob.data() + kOffWdesc, wdesc.data(), 4 * wdesc.size());
    for (int b = 0; b < 256; ++b) {
        uint32_t lo = 0, hi = 0;
        for (int i = 0; i < 4; ++i) {
            lo |= static_cast<uint32_t>((b >> i) & 1) << (8 * i);
            hi |= static_cast<uint32_t>((b >> (4 + i)) & 1) << (8 * i);
        }
        memcpy(blob.data() + kOffLut + 8 * b, &lo, 4);
        memcpy(blob.data() + kOffLut + 8 * b + 4, &hi, 4);
    }
    g.fld_words = (g.nf + 3 + 3) & ~3;                   // >= nf + 3 zero slots for the gather
    g.warp_smem = align16(4 * g.fld_words + 4 * 32);      // fields + 32 per-pass popcount totals (sampler)
    if (N == 20 && (g.A != 30433 || g.nf != 1665 || g.mw != 952 || g.fld_words != 1668 || !g.fast3))
        return fail(BLK_ERR_ARG, "internal: N=20 constants in the specialised kernels are stale");
    CUDA_TRY(cudaMalloc(&h->d_tables, off));
    CUDA_TRY(cudaMemcpy(h->d_tables, blob.data(), off, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&h->d_queue, sizeof(unsigned long long) * 2 * kQueueSlots));
    CUDA_TRY(cudaMemset(h->d_queue, 0, sizeof(unsigned long long) * 2 * kQueueSlots));
    return BLK_OK;
}

int grid_for(int64_t units, int per_block, int sm_count, int blocks_per_sm) {
    const int64_t need = (units + per_block - 1) / per_block;
    const int64_t cap = static_cast<int64_t>(sm_count) * blocks_per_sm;
    return static_cast<int>(need < cap ? (need > 0 ? need : 1) : cap);
}

}  // namespace

extern "C" {

const char *blk_last_error(void) { return g_err.c_str(); }
int blk_abi_version(void) { return BLK_ABI_VERSION; }

int blk_create(const blk_config *cfg, blk_engine **out) {
    if (!cfg || !out) return fail(BLK_ERR_ARG, "null argument");
    *out = nullptr;
    if (cfg->board_size < 5 || cfg->board_size > kMaxN) return fail(BLK_ERR_ARG, "board_size must be in 5..20");
    if (cfg->num_players != 2 && cfg->num_players != 4) return fail(BLK_ERR_ARG, "num_players must be 2 or 4");
    if (cfg->score_rule != 0 && cfg->score_rule != 1) return fail(BLK_ERR_ARG, "score_rule must be 0 or 1");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(BLK_ERR_NODEV, "no CUDA device: this engine has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(BLK_ERR_ARG, "device ordinal out of range");
    CUDA_TRY(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) return fail(BLK_ERR_NODEV, "device is not sm_100 (B200); kernels are built for sm_100a only");
    blk_engine *h = new blk_engine();
    h->cfg = *cfg;
    h->sm_count = prop.multiProcessorCount;
    int rc = build_tables(h);
    if (rc != BLK_OK) { blk_destroy(h); return rc; }
    h->step_smem = h->t.bytes + 16 + kWarps * h->g.warp_smem;
    // specialised <20,4> kernels for the headline geometry, runtime-dimension <0,0> kernels for everything else
    h->special = cfg->board_size == 20 && cfg->num_players == 4;
#define BLK_STEP_VARIANTS(NN, PP)                                                                     \
    do {                                                                                              \
        h->step_fn[0][0] = step_kernel<NN, PP, 0, false>; h->step_fn[0][1] = step_kernel<NN, PP, 0, true>; \
        h->step_fn[1][0] = step_kernel<NN, PP, 1, false>; h->step_fn[1][1] = step_kernel<NN, PP, 1, true>; \
        h->step_fn[2][0] = step_kernel<NN, PP, 2, false>; h->step_fn[2][1] = step_kernel<NN, PP, 2, true>; \
        h->step_fn[3][0] = step_kernel<NN, PP, 3, false>; h->step_fn[3][1] = step_kernel<NN, PP, 3, true>; \
    } while (0)
    if (h->special) BLK_STEP_VARIANTS(20, 4); else BLK_STEP_VARIANTS(0, 0);
#undef BLK_STEP_VARIANTS
    h->rollout_fn = h->special ? rollout_kernel<20, 4> : rollout_kernel<0, 0>;
    // the attribute is per function, not per engine: only ever raise it (engines of several board sizes coexist)
    static int s_max_smem[16][2] = {};
    int &cur_max = s_max_smem[cfg->device & 15][h->special ? 1 : 0];
    if (h->step_smem > cur_max) {
        cudaError_t err = cudaFuncSetAttribute(h->rollout_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, h->step_smem);
        for (int f = 0; f < 4 && err == cudaSuccess; ++f)
            for (int sm = 0; sm < 2 && err == cudaSuccess; ++sm)
                err = cudaFuncSetAttribute(h->step_fn[f][sm], cudaFuncAttributeMaxDynamicSharedMemorySize, h->step_smem);
        if (err != cudaSuccess) { blk_destroy(h); return fail(BLK_ERR_CUDA, "cudaFuncSetAttribute(smem) failed"); }
        cur_max = h->step_smem;
    }
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->step_blocks_per_sm, h->step_fn[2][1], kWarps * 32, h->step_smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->rollout_blocks_per_sm, h->rollout_fn, kWarps * 32, h->step_smem);
    if (h->step_blocks_per_sm < 1 || h->rollout_blocks_per_sm < 1) { blk_destroy(h); return fail(BLK_ERR_CUDA, "kernel does not fit on an SM"); }
    *out = h;
    return BLK_OK;
}

void blk_destroy(blk_engine *h) {
    if (!h) return;
    if (h->d_tables) cudaFree(h->d_tables);
    if (h->d_queue) cudaFree(h->d_queue);
    delete h;
}

int blk_get_info(const blk_engine *h, blk_info *out) {
    if (!h || !out) return fail(BLK_ERR_ARG, "null argument");
    out->abi_version = BLK_ABI_VERSION;
    out->board_size = h->g.N; out->num_players = h->g.P; out->num_actions = h->g.A;
    out->num_pieces = kPieces; out->num_orients = kOrients; out->num_fields = h->g.nf;
    out->state_words = h->g.sw; out->mask_words = h->g.mw; out->mask_bytes = h->g.mask_bytes;
    out->sm_count = h->sm_count;
    return BLK_OK;
}

int blk_action_to_cells(const blk_engine *h, int32_t action, int32_t meta[4], uint8_t cells_yx[10], int32_t *ncells) {
    if (!h) return fail(BLK_ERR_ARG, "null engine");
    if (action < 0 || action >= h->g.A) return fail(BLK_ERR_ARG, "action id out of range");
    const int o = h->act_o[action];
    const OrientRow &r = kOrient[o];
    if (meta) { meta[0] = r.piece; meta[1] = r.local; meta[2] = h->act_y[action]; meta[3] = h->act_x[action]; }
    if (cells_yx)
        for (int c = 0; c < r.n; ++c) {
            cells_yx[2 * c] = static_cast<uint8_t>(h->act_y[action] + r.yx[2 * c]);
            cells_yx[2 * c + 1] = static_cast<uint8_t>(h->act_x[action] + r.yx[2 * c + 1]);
        }
    if (ncells) *ncells = r.n;
    return BLK_OK;
}

int blk_reset(blk_engine *h, uint32_t *state, int64_t n, void *stream) {
    if (!h || (!state && n > 0) || n < 0) return fail(BLK_ERR_ARG, "bad argument to blk_reset");
    if (n == 0) return BLK_OK;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const int64_t words = n * h->g.sw;
    reset_kernel<<<static_cast<unsigned>((words + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(state, n, h->g);
    CUDA_TRY(cudaGetLastError());
    return BLK_OK;
}

int blk_step(blk_engine *h, const blk_step_args *args, void *stream) {
    if (!h || !args) return fail(BLK_ERR_ARG, "null argument");
    if (args->n < 0 || (args->n > 0 && !args->state_in)) return fail(BLK_ERR_ARG, "bad n / state_in");
    if (args->n == 0) return BLK_OK;
    if (args->action && !args->state_out) return fail(BLK_ERR_ARG, "stepping needs state_out");
    if (args->mask_format != BLK_MASK_NONE) {
        if (!args->mask) return fail(BLK_ERR_ARG, "mask_format set but mask is NULL");
        if (args->mask_format == BLK_MASK_BITS && args->mask_stride < h->g.mw) return fail(BLK_ERR_ARG, "mask_stride < mask_words");
        if (args->mask_format == BLK_MASK_BYTES && args->mask_stride < h->g.A) return fail(BLK_ERR_ARG, "mask_stride < num_actions");
        if (args->mask_format != BLK_MASK_BITS && args->mask_format != BLK_MASK_BYTES) return fail(BLK_ERR_ARG, "unknown mask_format");
    }
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    KParams kp;
    kp.a = *args; kp.tables = h->d_tables; kp.t = h->t; kp.g = h->g;
    kp.queue = h->d_queue + 2 * (h->launch_seq++ % kQueueSlots);
    const int grid = grid_for(args->n, kWarps, h->sm_count, h->step_blocks_per_sm);
    int variant = args->mask_format;                  // 0 none, 1 bits, 2 bytes (vector stores), 3 bytes (unaligned buffer)
    if (variant == BLK_MASK_BYTES &&
        ((args->mask_stride & 15) != 0 || (reinterpret_cast<uintptr_t>(args->mask) & 15) != 0 || args->mask_stride < h->g.mask_bytes))
        variant = 3;
    h->step_fn[variant][args->next_action != nullptr ? 1 : 0]<<<grid, kWarps * 32, h->step_smem, static_cast<cudaStream_t>(stream)>>>(kp);
    CUDA_TRY(cudaGetLastError());
    return BLK_OK;
}

int blk_observe(blk_engine *h, const uint32_t *state, float *obs, int64_t n, void *stream) {
    if (!h || n < 0 || (n > 0 && (!state || !obs))) return fail(BLK_ERR_ARG, "bad argument to blk_observe");
    if (n == 0) return BLK_OK;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const int64_t elems = n * 2 * h->g.P * h->g.N * h->g.N;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if ((h->g.N * h->g.N) % 4 == 0 && (reinterpret_cast<uintptr_t>(obs) & 15) == 0) {
        const int64_t thr = elems / 4;
        observe_kernel<4><<<static_cast<unsigned>((thr + 255) / 256), 256, 0, st>>>(state, obs, n, h->g);
    } else {
        observe_kernel<1><<<static_cast<unsigned>((elems + 255) / 256), 256, 0, st>>>(state, obs, n, h->g);
    }
    CUDA_TRY(cudaGetLastError());
    return BLK_OK;
}

int blk_board_contents(blk_engine *h, const uint32_t *state, uint8_t *board, int64_t n, void *stream) {
    if (!h || n < 0 || (n > 0 && (!state || !board))) return fail(BLK_ERR_ARG, "bad argument to blk_board_contents");
    if (n == 0) return BLK_OK;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const int64_t elems = n * h->g.N * h->g.N;
    contents_kernel<<<static_cast<unsigned>((elems + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(state, board, n, h->g);
    CUDA_TRY(cudaGetLastError());
    return BLK_OK;
}

int blk_game_ended(blk_engine *h, const uint32_t *state, uint8_t *flags, float *terminal, int16_t *scores, int64_t n,
                   void *stream) {
    if (!h || n < 0 || (n > 0 && !state)) return fail(BLK_ERR_ARG, "bad argument to blk_game_ended");
    if (n == 0) return BLK_OK;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    ended_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(state, flags, terminal, scores, n, h->g);
    CUDA_TRY(cudaGetLastError());
    return BLK_OK;
}

int blk_rollout(blk_engine *h, const blk_rollout_args *args, void *stream) {
    if (!h || !args) return fail(BLK_ERR_ARG, "null argument");
    if (args->n_roots < 0 || args->per_root < 0) return fail(BLK_ERR_ARG, "negative count");
    if (args->n_roots == 0 || args->per_root == 0) return BLK_OK;
    if (!args->roots || !args->final_scores) return fail(BLK_ERR_ARG, "roots / final_scores are required");
    if (args->action_log && args->log_stride < 4 * kPieces + 1) return fail(BLK_ERR_ARG, "log_stride must be >= 85");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    RParams rp;
    rp.a = *args; rp.tables = h->d_tables; rp.t = h->t; rp.g = h->g;
    rp.queue = h->d_queue + 2 * (h->launch_seq++ % kQueueSlots);
    const int grid = grid_for(args->n_roots * args->per_root, kWarps, h->sm_count, h->rollout_blocks_per_sm);
    h->rollout_fn<<<grid, kWarps * 32, h->step_smem, static_cast<cudaStream_t>(stream)>>>(rp);
    CUDA_TRY(cudaGetLastError());
    return BLK_OK;
}

}  // extern "C"
