// blk_engine.cu -- host side of the B200 (sm_100a) batched Blokus environment engine: tables, launch logic, C ABI,
// and the small streaming kernels.  The step / rollout kernels live in blk_kernels.cuh (instantiated per geometry
// by blk_inst.cu).
//
// Hot path (reference call sites in blokus_rl/colossumrl/blokus_wrapper.py):
//   a1 new_state       -> reset_kernel              a2 next_state     -> step_kernel
//   a3 valid_actions   -> step_kernel               a4 get_winners    -> step_kernel / ended_kernel
//   a5 canonical_board -> observe_kernel            a6 board_contents -> contents_kernel
//   a7 action table    -> build_tables() (host)     rollouts (new)    -> rollout_kernel
//
// No tensor cores: nothing here is a contraction.  No CPU fallback: every entry point needs the GPU.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "blk_kernels.cuh"
#include "blk_search.cuh"

using namespace blk;

namespace {

// ---------------------------------------------------------------------------------------------
// host-side orientation metadata (same generated list the kernels unroll)
// ---------------------------------------------------------------------------------------------
struct OrientRow {
    int8_t piece, local, h, w, n;
    int8_t yx[10];
};
#define BLK_BASE(...)
#define BLK_ORIENT_B(...)
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
#undef BLK_BASE
#undef BLK_ORIENT_B
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
// Entry points run on the engine's device and leave the caller's current device as they found it (a process that
// drives engines on several GPUs, or torch with another current device, must not be switched under its feet).
struct DeviceGuard {
    int prev = -1, dev;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int d) : dev(d) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        if (prev >= 0 && prev != dev) cudaSetDevice(prev);
    }
};

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail(BLK_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));          \
    } while (0)

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

// Warp-per-env streaming variants for board sizes with N*N % 16 == 0 (20x20: every index division is by a constant).
// The env's row words sit in shared memory; a 16-entry LUT turns 4 occupancy bits into one float4.
template <int kN, int kP>
__global__ void __launch_bounds__(256) observe_rows_kernel(const uint32_t *__restrict__ state, float *__restrict__ obs,
                                                          int64_t n) {
    constexpr int kQuadsPerPlane = kN * kN / 4, kQuadsPerRow = kN / 4, kQuads = 2 * kP * kQuadsPerPlane;
    constexpr int kSw = kP * kN + kP + 4;
    __shared__ float4 s_lut[16];
    __shared__ uint32_t s_rows[8][kP * kN + 1];
    if (threadIdx.x < 16)
        s_lut[threadIdx.x] = make_float4(threadIdx.x & 1 ? 1.f : 0.f, threadIdx.x & 2 ? 1.f : 0.f,
                                         threadIdx.x & 4 ? 1.f : 0.f, threadIdx.x & 8 ? 1.f : 0.f);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t *rows = s_rows[warp];
    for (int64_t env = static_cast<int64_t>(blockIdx.x) * 8 + warp; env < n; env += static_cast<int64_t>(gridDim.x) * 8) {
        const uint32_t *s = state + env * kSw;
        for (int i = lane; i < kP * kN; i += 32) rows[i] = __ldg(s + i);
        if (lane == 0) rows[kP * kN] = __ldg(s + kP * kN + kP) & 15u;          // mover
        __syncwarp();
        const int mover = rows[kP * kN];
        float4 *out = reinterpret_cast<float4 *>(obs + env * (4 * kQuads));
#pragma unroll 5
        for (int j = lane; j < kQuads; j += 32) {
            const int plane = j / kQuadsPerPlane, w = j - plane * kQuadsPerPlane;
            const int y = w / kQuadsPerRow, x0 = (w - y * kQuadsPerRow) * 4;
            float4 v;
            if (plane < kP) v = s_lut[(rows[plane * kN + y] >> x0) & 15u];
            else v = s_lut[(plane - kP == mover) ? 15 : 0];
            __stcs(out + j, v);
        }
        __syncwarp();
    }
}

template <int kN, int kP>
__global__ void __launch_bounds__(256) contents_rows_kernel(const uint32_t *__restrict__ state, uint8_t *__restrict__ board,
                                                           int64_t n) {
    constexpr int kChunks = kN * kN / 16, kSw = kP * kN + kP + 4;
    __shared__ uint32_t s_rows[8][kP * kN];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t *rows = s_rows[warp];
    for (int64_t env = static_cast<int64_t>(blockIdx.x) * 8 + warp; env < n; env += static_cast<int64_t>(gridDim.x) * 8) {
        const uint32_t *s = state + env * kSw;
        for (int i = lane; i < kP * kN; i += 32) rows[i] = __ldg(s + i);
        __syncwarp();
        if (lane < kChunks) {
            uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int cell = 16 * lane + i, y = cell / kN, x = cell - y * kN;
                uint32_t c = 0u;
#pragma unroll
                for (int q = 0; q < kP; ++q) c |= ((rows[q * kN + y] >> x) & 1u) * static_cast<uint32_t>(q + 1);
                w[i >> 2] |= c << (8 * (i & 3));
            }
            reinterpret_cast<uint4 *>(board + env * (kN * kN))[lane] = make_uint4(w[0], w[1], w[2], w[3]);
        }
        __syncwarp();
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

}  // namespace

// ---------------------------------------------------------------------------------------------
// engine object and C ABI
// ---------------------------------------------------------------------------------------------
struct blk_engine {
    blk_config cfg;
    Geometry g;
    TableLayout t;
    unsigned char *d_tables = nullptr;
    // Work-queue counters: kQueueSlots x {ticket, finished}, self-resetting at the end of every launch.  Two launches may
    // share a slot only if they cannot overlap, so a slot belongs to ONE stream (launches of a stream run in order) or
    // to ONE stream capture (a captured launch keeps its slot for every replay of the graph, wherever that is
    // launched; eager launches never get that slot).  `mu` guards the map: entry points may be called from several
    // host threads.
    unsigned long long *d_queue = nullptr;
    std::mutex mu;
    std::unordered_map<unsigned long long, int> slot_of;
    int sm_count = 0;
    int step_smem = 0, roll_smem = 0;
    int step_blocks_per_sm = 0, rollout_blocks_per_sm = 0;
    bool special = false;
    KernelSet ks = {};                               // step[mask format variant][sampler], rollout
    SearchKernelSet sk = {};                         // fused PUCT search / reroot (blk_search.cuh)
    int search_smem_per_warp = 0;
    bool small = false;                              // N <= 7: thread-per-env kernels (blk_small.cu) for the common formats
    SmallKernelSet sks = {};
    unsigned char *d_small = nullptr;                // ocells64[92] | first_mask[mw]
    int first_count = 0, small_blocks_per_sm = 0, small_roll_blocks_per_sm = 0;
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
    t.off_wsrc = off;   off = align16(off + 2 * g.mw);
    t.off_obslut = off; off = align16(off + 16 * 16);
    t.roll_begin = off;                                      // everything from here on is what the rollout kernel stages
    t.off_oinfo = off;  off = align16(off + 4 * (kOrients + 1));
    t.off_ocells = off; off = align16(off + 4 * (kOrients + 1));
    t.off_foff = off;   off = align16(off + 2 * (g.nf + 1));
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
        for (int c = 0; c < r.n; ++c) cells |= 1u << (5 * r.yx[2 * c] + r.yx[2 * c + 1]);   // row patterns, 5 bits per row
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
    // gather descriptors: word g = (fld[s] >> r0) | (fld[s+1] << s1) | (fld[s+2] << s2) [| (fld[s+3] << s3) | (fld[s+4] << s4)];
    // shifts >= 32 give 0.  Three fields per word suffice at N >= 15 (fields >= 11 bits wide), five at N >= 12.
    std::vector<uint32_t> wdesc(2 * 32 * g.rounds);
    auto start = [&](int f) { return f < g.nf ? static_cast<int>(foff[f]) : (1 << 20); };   // past the last field: never
    auto clamp63 = [](int v) { return static_cast<uint32_t>(v > 63 ? 63 : v); };
    g.gather = 0;
    for (int reach : {3, 5}) {
        bool fits = true;
        for (int w = 0; w < g.mw && fits; ++w) {
            const int s0 = wsrc[w];
            if (s0 >= g.nf) continue;                   // padding word: gathers the zero slots behind the fields
            const int r0 = 32 * w - start(s0);
            fits = start(s0 + reach) >= 32 * w + 32 && r0 >= 0 && r0 <= 31;   // no further field reaches into this word
        }
        if (fits) { g.gather = reach; break; }
    }
    for (int w = 0; w < 32 * g.rounds; ++w) {
        wdesc[2 * w] = 4u * g.nf | (g.gather == 5 ? (63u << 16) | (63u << 24) : 0u);
        wdesc[2 * w + 1] = (63u << 8) | (63u << 16);
    }
    for (int w = 0; w < g.mw && g.gather; ++w) {
        const int s0 = wsrc[w];
        if (s0 >= g.nf) continue;
        const int b = 32 * w;
        wdesc[2 * w] = 4u * static_cast<uint32_t>(s0);
        if (g.gather == 5) wdesc[2 * w] |= (clamp63(start(s0 + 3) - b) << 16) | (clamp63(start(s0 + 4) - b) << 24);
        wdesc[2 * w + 1] = static_cast<uint32_t>(b - start(s0)) | (clamp63(start(s0 + 1) - b) << 8) | (clamp63(start(s0 + 2) - b) << 16);
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
    for (int b = 0; b < 16; ++b) {
        const float q[4] = {b & 1 ? 1.f : 0.f, b & 2 ? 1.f : 0.f, b & 4 ? 1.f : 0.f, b & 8 ? 1.f : 0.f};
        memcpy(blob.data() + t.off_obslut + 16 * b, q, 16);
    }
    g.fld_words = (g.nf + (g.gather == 5 ? 5 : 3) + 3) & ~3;   // >= nf + 3 (or 5) zero slots for the gather
    if (N == 14) g.fld_words = 1152;                     // 32 lanes x 36 fields: the chunked readers use 16 B loads
    g.warp_smem = align16(4 * g.fld_words + 4 * 32);      // fields + 32 per-pass popcount totals (sampler)
    if ((N == 20 && (g.A != 30433 || g.nf != 1665 || g.mw != 952 || g.rounds != 30 || g.fld_words != 1668 || g.gather != 3)) ||
        (N == 14 && (g.A != 13729 || g.nf != 1119 || g.mw != 432 || g.rounds != 14 || g.fld_words != 1152 || g.gather != 5)))
        return fail(BLK_ERR_ARG, "internal: the geometry constants in the specialised kernels are stale");
    CUDA_TRY(cudaMalloc(&h->d_tables, off));
    CUDA_TRY(cudaMemcpy(h->d_tables, blob.data(), off, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&h->d_queue, sizeof(unsigned long long) * 2 * kQueueSlots));
    CUDA_TRY(cudaMemset(h->d_queue, 0, sizeof(unsigned long long) * 2 * kQueueSlots));
    return BLK_OK;
}

// The work-queue slot of a launch on `st` (see blk_engine::d_queue).
int queue_slot(blk_engine *h, cudaStream_t st, unsigned long long **out) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    unsigned long long cap_id = 0;
    if (cudaStreamGetCaptureInfo(st, &cs, &cap_id) != cudaSuccess) {
        cudaGetLastError();
        return fail(BLK_ERR_CUDA, "cudaStreamGetCaptureInfo failed (bad stream handle?)");
    }
    // keys: streams are pointers (low bits 0, or the small special handles 0/1/2); captures get an odd key that mixes
    // the capture id with the stream, so forked capture streams of one capture do not share a slot either
    unsigned long long key = static_cast<unsigned long long>(reinterpret_cast<uintptr_t>(st)) << 1;
    if (cs == cudaStreamCaptureStatusActive) key = ((cap_id * 0x9E3779B97F4A7C15ULL) ^ key) | 1ULL;
    else if (cs != cudaStreamCaptureStatusNone) return fail(BLK_ERR_CUDA, "stream capture was invalidated");
    std::lock_guard<std::mutex> lock(h->mu);
    auto it = h->slot_of.find(key);
    int slot;
    if (it != h->slot_of.end()) {
        slot = it->second;
    } else {
        if (static_cast<int>(h->slot_of.size()) >= kQueueSlots)
            return fail(BLK_ERR_ARG, "too many distinct streams / graph captures on one engine (work-queue slots exhausted)");
        slot = static_cast<int>(h->slot_of.size());
        h->slot_of.emplace(key, slot);
    }
    *out = h->d_queue + 2 * slot;
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
    DeviceGuard guard(cfg->device);
    CUDA_TRY(guard.err);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) return fail(BLK_ERR_NODEV, "device is not sm_100 (B200); kernels are built for sm_100a only");
    blk_engine *h = new blk_engine();
    h->cfg = *cfg;
    h->sm_count = prop.multiProcessorCount;
    int rc = build_tables(h);
    if (rc != BLK_OK) { blk_destroy(h); return rc; }
    h->step_smem = h->t.bytes + 16 + kWarps * h->g.warp_smem;
    h->roll_smem = (h->t.bytes - h->t.roll_begin) + 16 + kRollWarps * h->g.warp_smem;
    // specialised kernels for the geometries the reference's configs name, runtime-dimension kernels otherwise
    const int N = cfg->board_size, P = cfg->num_players;
    int geom = 0;
    if (N == 20 && P == 4) { h->ks = kernels_20_4(); h->sk = search_kernels_20_4(); geom = 1; }
    else if (N == 20 && P == 2) { h->ks = kernels_20_2(); h->sk = search_kernels_20_2(); geom = 2; }
    else if (N == 14 && P == 4) { h->ks = kernels_14_4(); h->sk = search_kernels_14_4(); geom = 3; }
    else if (N == 14 && P == 2) { h->ks = kernels_14_2(); h->sk = search_kernels_14_2(); geom = 4; }
    else if (N == 7 && P == 2) { h->ks = kernels_7_2(); h->sk = search_kernels_7_2(); geom = 5; }
    else { h->ks = kernels_0_0(); h->sk = search_kernels_0_0(); }
    h->search_smem_per_warp = search_warp_bytes(h->g.warp_smem);
    h->special = geom != 0;
    // the attribute is per function, not per engine: only ever raise it (engines of several board sizes coexist)
    static int s_max_smem[16][6] = {};
    int &cur_max = s_max_smem[cfg->device & 15][geom];
    if (h->step_smem > cur_max) {
        cudaError_t err = cudaFuncSetAttribute(h->ks.rollout, cudaFuncAttributeMaxDynamicSharedMemorySize, h->roll_smem);
        for (int f = 0; f < 5 && err == cudaSuccess; ++f)
            for (int sm = 0; sm < 2 && err == cudaSuccess; ++sm)
                err = cudaFuncSetAttribute(h->ks.step[f][sm], cudaFuncAttributeMaxDynamicSharedMemorySize, h->step_smem);
        for (int k = 0; k < 2 && err == cudaSuccess; ++k)
            err = cudaFuncSetAttribute(h->sk.search[k], cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       h->t.bytes + 16 + (k ? 16 : kSearchTrees) * h->search_smem_per_warp);
        if (err != cudaSuccess) { blk_destroy(h); return fail(BLK_ERR_CUDA, "cudaFuncSetAttribute(smem) failed"); }
        cur_max = h->step_smem;
    }
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->step_blocks_per_sm, h->ks.step[2][1], kWarps * 32, h->step_smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->rollout_blocks_per_sm, h->ks.rollout, kRollWarps * 32, h->roll_smem);
    if (h->step_blocks_per_sm < 1 || h->rollout_blocks_per_sm < 1) { blk_destroy(h); return fail(BLK_ERR_CUDA, "kernel does not fit on an SM"); }
    if (N <= 7) {
        h->sks = N == 5 ? kernels_small_5() : (N == 6 ? kernels_small_6() : kernels_small_7());
        if (h->sks.num_actions != h->g.A) { blk_destroy(h); return fail(BLK_ERR_ARG, "internal: small-board tables are stale"); }
        // footprints with row stride 8, and the (constant) legal mask of the fresh board: every footprint that
        // covers player 0's start corner (0, 0)
        std::vector<unsigned char> blob(8 * 92 + 4 * h->g.mw, 0);
        uint64_t *oc = reinterpret_cast<uint64_t *>(blob.data());
        uint32_t *fm = reinterpret_cast<uint32_t *>(blob.data() + 8 * 92);
        for (int o = 0; o < kOrients; ++o)
            for (int c = 0; c < kOrient[o].n; ++c) oc[o] |= 1ull << (8 * kOrient[o].yx[2 * c] + kOrient[o].yx[2 * c + 1]);
        for (int act = 0; act < h->g.A; ++act) {
            if (h->act_y[act] != 0 || h->act_x[act] != 0) continue;
            const OrientRow &r = kOrient[h->act_o[act]];
            bool corner = false;
            for (int c = 0; c < r.n; ++c) corner |= r.yx[2 * c] == 0 && r.yx[2 * c + 1] == 0;
            if (corner) { fm[act >> 5] |= 1u << (act & 31); ++h->first_count; }
        }
        if (cudaMalloc(&h->d_small, blob.size()) != cudaSuccess ||
            cudaMemcpy(h->d_small, blob.data(), blob.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
            blk_destroy(h); return fail(BLK_ERR_CUDA, "small-board table upload failed");
        }
        const int pi = P == 4 ? 1 : 0;
        for (int f = 0; f < 5; ++f)
            for (int sm = 0; sm < 2; ++sm)
                if (cudaFuncSetAttribute(h->sks.step[pi][f][sm], cudaFuncAttributeMaxDynamicSharedMemorySize, h->sks.smem[pi]) != cudaSuccess) {
                    blk_destroy(h); return fail(BLK_ERR_CUDA, "cudaFuncSetAttribute(smem) failed for the small-board kernels");
                }
        if (cudaFuncSetAttribute(h->sks.rollout[pi], cudaFuncAttributeMaxDynamicSharedMemorySize, h->sks.roll_smem) != cudaSuccess) {
            blk_destroy(h); return fail(BLK_ERR_CUDA, "cudaFuncSetAttribute(smem) failed for the small-board playout kernel");
        }
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->small_blocks_per_sm, h->sks.step[pi][2][1], h->sks.threads[pi], h->sks.smem[pi]);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->small_roll_blocks_per_sm, h->sks.rollout[pi], h->sks.roll_threads, h->sks.roll_smem);
        h->small = h->small_blocks_per_sm >= 1 && h->small_roll_blocks_per_sm >= 1;
    }
    *out = h;
    return BLK_OK;
}

void blk_destroy(blk_engine *h) {
    if (!h) return;
    DeviceGuard guard(h->cfg.device);
    if (h->d_tables) cudaFree(h->d_tables);
    if (h->d_queue) cudaFree(h->d_queue);
    if (h->d_small) cudaFree(h->d_small);
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
    DeviceGuard guard(h->cfg.device);
    CUDA_TRY(guard.err);
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
        if (args->mask_format != BLK_MASK_BITS && args->mask_format != BLK_MASK_BYTES && args->mask_format != BLK_MASK_INDICES)
            return fail(BLK_ERR_ARG, "unknown mask_format");
        if (args->mask_format == BLK_MASK_INDICES && (!args->legal_count || args->mask_stride < 1))
            return fail(BLK_ERR_ARG, "BLK_MASK_INDICES needs legal_count and a positive mask_stride");
        if ((args->csr_cursor != nullptr) != (args->csr_offset != nullptr) ||
            (args->csr_cursor != nullptr && args->mask_format != BLK_MASK_INDICES))
            return fail(BLK_ERR_ARG, "csr_cursor and csr_offset go together and belong to BLK_MASK_INDICES");
    }
    if (args->obs && (reinterpret_cast<uintptr_t>(args->obs) & 15) != 0) return fail(BLK_ERR_ARG, "obs must be 16 B aligned");
    if (args->state_index && (!args->state_out || args->state_out == args->state_in))
        return fail(BLK_ERR_ARG, "state_index needs a separate state_out");
    DeviceGuard guard(h->cfg.device);
    CUDA_TRY(guard.err);
    KParams kp;
    kp.a = *args; kp.tables = h->d_tables; kp.t = h->t; kp.g = h->g;
    const int grid = grid_for(args->n, kWarps, h->sm_count, h->step_blocks_per_sm);
    int variant = args->mask_format;                  // 0 none, 1 bits, 2 bytes (vector stores), 3 bytes (unaligned buffer), 4 ids
    if (variant == BLK_MASK_INDICES) variant = 4;
    if (variant == BLK_MASK_BYTES &&
        ((args->mask_stride & 15) != 0 || (reinterpret_cast<uintptr_t>(args->mask) & 15) != 0 || args->mask_stride < h->g.mask_bytes))
        variant = 3;
    const bool bits_rows_8b = variant != BLK_MASK_BITS ||
                              ((args->mask_stride & 1) == 0 && (reinterpret_cast<uintptr_t>(args->mask) & 7) == 0);
    if (h->small && bits_rows_8b && !(args->options & BLK_OPT_WARP_KERNELS) && args->csr_cursor == nullptr) {   // (compact index lists: warp kernels)
        // N <= 7: one env per thread on 64-bit bitboards (blk_small.cu); the other formats stay on the warp-per-env kernel
        SmallParams sp;
        sp.a = *args; sp.tables = h->d_tables; sp.t = h->t; sp.g = h->g;
        sp.ocells64 = reinterpret_cast<const uint64_t *>(h->d_small);
        sp.first_mask = reinterpret_cast<const uint32_t *>(h->d_small + 8 * 92);
        sp.first_count = h->first_count;
        const int pi = h->g.P == 4 ? 1 : 0;
        const int sgrid = grid_for(args->n, h->sks.threads[pi], h->sm_count, h->small_blocks_per_sm);
        h->sks.step[pi][variant][args->next_action != nullptr ? 1 : 0]<<<sgrid, h->sks.threads[pi], h->sks.smem[pi], static_cast<cudaStream_t>(stream)>>>(sp);
        CUDA_TRY(cudaGetLastError());
        return BLK_OK;
    }
    {
        const int rc = queue_slot(h, static_cast<cudaStream_t>(stream), &kp.queue);
        if (rc != BLK_OK) return rc;
    }
    h->ks.step[variant][args->next_action != nullptr ? 1 : 0]<<<grid, kWarps * 32, h->step_smem, static_cast<cudaStream_t>(stream)>>>(kp);
    CUDA_TRY(cudaGetLastError());
    return BLK_OK;
}

int blk_observe(blk_engine *h, const uint32_t *state, float *obs, int64_t n, void *stream) {
    if (!h || n < 0 || (n > 0 && (!state || !obs))) return fail(BLK_ERR_ARG, "bad argument to blk_observe");
    if (n == 0) return BLK_OK;
    DeviceGuard guard(h->cfg.device);
    CUDA_TRY(guard.err);
    const int64_t elems = n * 2 * h->g.P * h->g.N * h->g.N;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool aligned = (reinterpret_cast<uintptr_t>(obs) & 15) == 0;
    const int wgrid = static_cast<int>(std::min<int64_t>((n + 7) / 8, static_cast<int64_t>(h->sm_count) * 8));
    if (aligned && h->g.N == 20 && h->g.P == 4) {
        observe_rows_kernel<20, 4><<<wgrid, 256, 0, st>>>(state, obs, n);
    } else if (aligned && h->g.N == 20 && h->g.P == 2) {
        observe_rows_kernel<20, 2><<<wgrid, 256, 0, st>>>(state, obs, n);
    } else if ((h->g.N * h->g.N) % 4 == 0 && aligned) {
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
    DeviceGuard guard(h->cfg.device);
    CUDA_TRY(guard.err);
    const int64_t elems = n * h->g.N * h->g.N;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int wgrid = static_cast<int>(std::min<int64_t>((n + 7) / 8, static_cast<int64_t>(h->sm_count) * 8));
    const bool aligned = (reinterpret_cast<uintptr_t>(board) & 15) == 0;
    if (aligned && h->g.N == 20 && h->g.P == 4) contents_rows_kernel<20, 4><<<wgrid, 256, 0, st>>>(state, board, n);
    else if (aligned && h->g.N == 20 && h->g.P == 2) contents_rows_kernel<20, 2><<<wgrid, 256, 0, st>>>(state, board, n);
    else contents_kernel<<<static_cast<unsigned>((elems + 255) / 256), 256, 0, st>>>(state, board, n, h->g);
    CUDA_TRY(cudaGetLastError());
    return BLK_OK;
}

int blk_game_ended(blk_engine *h, const uint32_t *state, uint8_t *flags, float *terminal, int16_t *scores, int64_t n,
                   void *stream) {
    if (!h || n < 0 || (n > 0 && !state)) return fail(BLK_ERR_ARG, "bad argument to blk_game_ended");
    if (n == 0) return BLK_OK;
    DeviceGuard guard(h->cfg.device);
    CUDA_TRY(guard.err);
    ended_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(state, flags, terminal, scores, n, h->g);
    CUDA_TRY(cudaGetLastError());
    return BLK_OK;
}

int blk_rollout(blk_engine *h, const blk_rollout_args *args, void *stream) {
    if (!h || !args) return fail(BLK_ERR_ARG, "null argument");
    if (args->n_roots < 0 || args->per_root < 0) return fail(BLK_ERR_ARG, "negative count");
    if (args->n_roots == 0 || args->per_root == 0) return BLK_OK;
    if (!args->roots) return fail(BLK_ERR_ARG, "roots are required");
    if (args->stop_player < -1 || args->stop_player >= h->g.P) return fail(BLK_ERR_ARG, "stop_player out of range");
    if (args->action_log && args->log_stride < 4 * kPieces + 1) return fail(BLK_ERR_ARG, "log_stride must be >= 85");
    DeviceGuard guard(h->cfg.device);
    CUDA_TRY(guard.err);
    if (h->small && !(args->options & BLK_OPT_WARP_KERNELS)) {
        // N <= 7: one playout per thread on 64-bit bitboards (blk_small.cu); identical games, ~an order of magnitude faster
        SmallRollParams sp;
        sp.a = *args; sp.tables = h->d_tables; sp.t = h->t; sp.g = h->g;
        sp.ocells64 = reinterpret_cast<const uint64_t *>(h->d_small);
        const int pi = h->g.P == 4 ? 1 : 0;
        const int sgrid = grid_for(args->n_roots * args->per_root, h->sks.roll_threads, h->sm_count, h->small_roll_blocks_per_sm);
        h->sks.rollout[pi]<<<sgrid, h->sks.roll_threads, h->sks.roll_smem, static_cast<cudaStream_t>(stream)>>>(sp);
        CUDA_TRY(cudaGetLastError());
        return BLK_OK;
    }
    RParams rp;
    rp.a = *args; rp.tables = h->d_tables; rp.t = h->t; rp.g = h->g;
    {
        const int rc = queue_slot(h, static_cast<cudaStream_t>(stream), &rp.queue);
        if (rc != BLK_OK) return rc;
    }
    const int grid = grid_for(args->n_roots * args->per_root, kRollWarps, h->sm_count, h->rollout_blocks_per_sm);
    h->ks.rollout<<<grid, kRollWarps * 32, h->roll_smem, static_cast<cudaStream_t>(stream)>>>(rp);
    CUDA_TRY(cudaGetLastError());
    return BLK_OK;
}

// ---- fused PUCT search (blk_search.cuh) ----
namespace {
int check_search_args(const blk_engine *h, const blk_puct_forest *f, const blk_puct_search_args *a) {
    if (!h || !f || !a) return fail(BLK_ERR_ARG, "null argument");
    if (f->num_players != h->g.P || f->num_actions != h->g.A) return fail(BLK_ERR_ARG, "forest geometry differs from the engine's");
    if (!a->pool) return fail(BLK_ERR_ARG, "the state pool is required");
    if (!f->hash_table || !f->node_hash || !f->node_tree || !f->node_front || f->hash_capacity < 2 || (f->hash_capacity & (f->hash_capacity - 1)))
        return fail(BLK_ERR_ARG, "the board-keyed node table (hash_table, node_hash, node_tree, node_front, power-of-two hash_capacity) is required");
    if (f->hash_capacity < f->node_capacity) return fail(BLK_ERR_ARG, "hash_capacity must be >= node_capacity");
    return BLK_OK;
}
}  // namespace

int blk_puct_search(blk_engine *h, const blk_puct_forest *f, const blk_puct_search_args *a, void *stream) {
    int rc = check_search_args(h, f, a);
    if (rc != BLK_OK) return rc;
    if (f->num_trees <= 0 || a->num_sims <= 0) return BLK_OK;
    const int wpt = a->warps_per_tree <= 1 ? 1 : a->warps_per_tree;
    if (wpt > 16) return fail(BLK_ERR_ARG, "warps_per_tree must be <= 16");
    if (wpt > 1 && !f->edge_vl) return fail(BLK_ERR_ARG, "leaf-parallel search needs edge_vl");
    if (a->playouts_per_leaf < 0) return fail(BLK_ERR_ARG, "playouts_per_leaf must be >= 0");
    DeviceGuard guard(h->cfg.device);
    CUDA_TRY(guard.err);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SearchParams sp;
    sp.f = *f; sp.a = *a; sp.tables = h->d_tables; sp.t = h->t; sp.g = h->g; sp.reroot_states = nullptr;
    if (wpt == 1) {
        const int grid = (f->num_trees + kSearchTrees - 1) / kSearchTrees;
        h->sk.search[0]<<<grid, kSearchTrees * 32, h->t.bytes + 16 + kSearchTrees * h->search_smem_per_warp, st>>>(sp);
    } else {
        h->sk.search[1]<<<f->num_trees, wpt * 32, h->t.bytes + 16 + wpt * h->search_smem_per_warp, st>>>(sp);
    }
    CUDA_TRY(cudaGetLastError());
    return BLK_OK;
}

int blk_puct_reroot(blk_engine *h, const blk_puct_forest *f, const blk_puct_search_args *a, const uint32_t *states, void *stream) {
    int rc = check_search_args(h, f, a);
    if (rc != BLK_OK) return rc;
    if (!states) return fail(BLK_ERR_ARG, "states are required");
    if (f->num_trees <= 0) return BLK_OK;
    DeviceGuard guard(h->cfg.device);
    CUDA_TRY(guard.err);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SearchParams sp;
    sp.f = *f; sp.a = *a; sp.tables = h->d_tables; sp.t = h->t; sp.g = h->g; sp.reroot_states = states;
    h->sk.reroot<<<(f->num_trees + kSearchTrees - 1) / kSearchTrees, kSearchTrees * 32, 0, st>>>(sp);
    CUDA_TRY(cudaGetLastError());
    return BLK_OK;
}

}  // extern "C"
