// blk_puct.cu -- device-resident PUCT forest: select / expand / backup for B independent searches in lockstep.
//
// GPU sibling of blokus_rl/alphazero/mcts.py.  The arithmetic is the reference's, in
// float64 with explicitly rounded operations (no FMA contraction), so visit counts and argmax choices equal the
// reference's whenever its priors/values are float64 (tests/test_gpu_puct.py checks this against the golden vectors
// produced by the unmodified reference file):
//   U = cpuct * P * sqrt(sum(N) + 1e-6) / (1 + N), argmax(Q + U) with the FIRST maximum       (mcts.py:42-46)
//   only the root level sees cpuct, deeper levels use 1                                        (mcts.py:50-52)
//   backup indexes the value of the player to move AFTER the action: Q <- (N*Q + v)/(N+1)      (mcts.py:47,53-56)
//   terminal states are never expanded; every visit returns their 3/1/-1 vector                (mcts.py:60-62)
// One difference is deliberate: nodes are keyed by their path, not by hash(board cells) (mcts.py:37), so two move
// orders reaching the same board are two nodes here.  The host-side BatchedMCTS keeps the reference's keying.
//
// One warp per tree for select and expand, one thread per tree for backup.  The env transition of the opened edge
// and the legal mask of the new state come from blk_step (the caller runs it between select and expand), the
// priors/values from the caller's evaluator; nothing here synchronises with the host.
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "../../include/blokus_b200.h"

namespace {

constexpr uint32_t kAll = 0xffffffffu;
constexpr int kWarpsPerBlock = 4;
constexpr int kSelUnroll = 4;      // edges per lane whose N, Q (and P) loads are in flight together in puct_select_kernel
constexpr int kMaxIds = 1024;      // legal ids compacted per sweep of puct_expand_kernel (20x20 positions have <= ~800)

// A fused expansion (ExpandArgs::fuse_backup) cannot move the pool-slot counter itself -- warps of the same launch still
// read it -- so it leaves a note and the next kernel that runs before an expansion applies it.
__device__ __forceinline__ void apply_pending_slot_advance(const blk_puct_forest &f) {
    if (blockIdx.x == 0 && threadIdx.x == 0 && f.counters[5]) { f.counters[4] += f.num_trees; f.counters[5] = 0; }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32) puct_select_kernel(blk_puct_forest f, double cpuct, int eps_fix) {
    apply_pending_slot_advance(f);
    const int t = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= f.num_trees) return;
    int node = f.root[t];
    int len = 0, status = BLK_PUCT_TERMINAL, leaf_edge = -1, src = -1, act = BLK_ACTION_NONE;
    for (int depth = 0;; ++depth) {
        if (f.node_terminal[node]) { status = BLK_PUCT_TERMINAL; break; }
        const int e0 = f.node_edge0[node];
        if (e0 < 0) { status = BLK_PUCT_NEED_EVAL; src = f.node_state[node]; break; }
        if (len == f.max_depth) { if (lane == 0) f.counters[2] = 1; status = BLK_PUCT_TERMINAL; break; }
        const int n = f.node_nedge[node];
        const bool uniform = f.node_uniform[node] != 0;           // P = 1/n for every edge: not stored per edge
        const double pu = __ddiv_rn(1.0, static_cast<double>(n));
        const double s = f.node_sum_n[node];           // sum of the edges' visit counts, maintained by the backup
        const double c = depth == 0 ? cpuct : 1.0;
        const double sq = __dsqrt_rn(__dadd_rn(s, (eps_fix || depth > 0) ? 1e-6 : 0.0));
        double best = -1.0e300;
        int besti = 0x7fffffff;
        for (int i0 = lane; i0 < n; i0 += 32 * kSelUnroll) {    // kSelUnroll edges per lane in flight, then the arithmetic
            double ep[kSelUnroll], en[kSelUnroll], eq[kSelUnroll];
#pragma unroll
            for (int k = 0; k < kSelUnroll; ++k) {
                const int i = i0 + 32 * k;
                const bool in = i < n;
                ep[k] = uniform ? pu : (in ? __ldg(f.edge_p + e0 + i) : 0.0);
                en[k] = in ? __ldg(f.edge_n + e0 + i) : 0.0;
                eq[k] = in ? __ldg(f.edge_q + e0 + i) : 0.0;
            }
#pragma unroll
            for (int k = 0; k < kSelUnroll; ++k) {
                const int i = i0 + 32 * k;
                const double u = __ddiv_rn(__dmul_rn(__dmul_rn(c, ep[k]), sq), __dadd_rn(1.0, en[k]));
                const double sc = __dadd_rn(eq[k], u);
                if (i < n && sc > best) { best = sc; besti = i; }   // ascending i per lane: keeps the first maximum
            }
        }
#pragma unroll
        for (int d = 16; d; d >>= 1) {
            const double ob = __shfl_xor_sync(kAll, best, d);
            const int oi = __shfl_xor_sync(kAll, besti, d);
            if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
        }
        const int e = e0 + besti;
        if (lane == 0) {
            f.path[static_cast<int64_t>(t) * f.max_depth + len] = e;
            f.path_node[static_cast<int64_t>(t) * f.max_depth + len] = node;
        }
        ++len;
        const int child = f.edge_child[e];
        if (child < 0) { status = BLK_PUCT_NEED_STEP; leaf_edge = e; src = f.node_state[node]; act = f.edge_action[e]; break; }
        node = child;
    }
    if (lane == 0) {
        f.status[t] = status; f.leaf_node[t] = node; f.leaf_edge[t] = leaf_edge; f.path_len[t] = len;
        // trees that need no transition still go through blk_step (action NONE) so the whole batch stays one launch
        f.src_slot[t] = src >= 0 ? src : f.node_state[node];
        f.step_action[t] = act;
    }
}

// 8 mask bytes (0/1 each) -> 8 bits
__device__ __forceinline__ uint32_t pack8(uint64_t b) { return static_cast<uint32_t>((b * 0x0102040810204080ULL) >> 56); }

struct ExpandArgs {
    blk_puct_forest f;
    int32_t new_slot_base, state_words, meta_word, attach_only;
    const uint32_t *new_states;   // [B][state_words] states written by blk_step for this simulation
    uint32_t *pool;               // state pool the new states are filed into (slot base + t), or NULL
    const uint8_t *mask;          // [B][mask_stride] byte masks of those states, or bit-packed rows (mask_bits)
    int32_t mask_bits;            // 1: `mask` holds bit-packed rows of mask_stride_words uint32 words
    int32_t mask_stride_words;
    const uint8_t *flags;         // [B]
    const float *terminal;        // [B][P]
    const void *prior;            // [B][prior_stride] float32 / float64, or NULL for the uniform prior
    int32_t prior_dtype;          // 0 uniform, 1 float32, 2 float64
    int64_t prior_stride;
    const double *value;          // [B][P]
    int32_t fuse_backup;          // 1: the expansion warp also walks the path back (no separate blk_puct_backup launch)
};

// expansion of one tree by one warp (mcts.py:59-71); every exit leaves the simulation's score vector in f.scores
__device__ __forceinline__ void expand_tree(const ExpandArgs &a, uint32_t *words, uint16_t *ids, int t, int lane) {
    const blk_puct_forest &f = a.f;
    const int P = f.num_players;
    const int st = f.status[t];
    double *score = f.scores + static_cast<int64_t>(t) * P;
    if (st == BLK_PUCT_TERMINAL) {
        if (lane < P) score[lane] = f.node_term_value[static_cast<int64_t>(f.leaf_node[t]) * P + lane];
        return;
    }
    int target = f.leaf_node[t];
    // pool slot of this tree's new state: explicit base, or the device-side counter (CUDA-graph friendly)
    const int slot = (a.new_slot_base >= 0 ? a.new_slot_base : f.counters[4]) + t;
    if (slot >= f.node_capacity) { if (lane == 0) f.counters[2] = 1; if (lane < P) score[lane] = 0.0; return; }
    if (a.pool != nullptr) {
        uint32_t sv[4];                                     // state_words <= 88 at 20x20/4p: at most 3 words per lane, loads first
#pragma unroll
        for (int k = 0; k < 4; ++k) { const int i = lane + 32 * k; sv[k] = i < a.state_words ? __ldg(a.new_states + static_cast<int64_t>(t) * a.state_words + i) : 0u; }
#pragma unroll
        for (int k = 0; k < 4; ++k) { const int i = lane + 32 * k; if (i < a.state_words) a.pool[static_cast<int64_t>(slot) * a.state_words + i] = sv[k]; }
    }
    if (st == BLK_PUCT_NEED_STEP) {
        const uint8_t fl = a.flags[t];
        int node = 0;
        if (lane == 0) {
            if (fl & BLK_FLAG_ILLEGAL) f.counters[3] = 1;
            node = atomicAdd(&f.counters[0], 1);
        }
        node = __shfl_sync(kAll, node, 0);
        if (node >= f.node_capacity) { if (lane == 0) f.counters[2] = 1; if (lane < P) score[lane] = 0.0; return; }
        const bool done = fl & BLK_FLAG_DONE;
        if (lane == 0) {
            f.node_state[node] = slot;
            f.node_mover[node] = static_cast<int8_t>(a.new_states[static_cast<int64_t>(t) * a.state_words + a.meta_word] & 15u);
            f.node_edge0[node] = -1;
            f.node_nedge[node] = 0;
            f.node_sum_n[node] = 0.0;
            f.node_terminal[node] = done ? 1 : 0;
            f.edge_child[f.leaf_edge[t]] = node;
            if (a.attach_only) f.root[t] = node;
        }
        if (lane < P) f.node_term_value[static_cast<int64_t>(node) * P + lane] = done ? static_cast<double>(a.terminal[t * P + lane]) : 0.0;
        if (a.attach_only) return;
        if (done) { if (lane < P) score[lane] = static_cast<double>(a.terminal[t * P + lane]); return; }
        target = node;
    } else if (a.attach_only) {
        return;
    }
    // ---- expand `target`: one edge per legal action, ascending ids (np.where order, mcts.py:64) ----
    const int64_t mstride = a.mask_bits ? a.mask_stride_words : f.mask_stride;
    const uint8_t *row = a.mask + static_cast<int64_t>(t) * mstride * (a.mask_bits ? 4 : 1);
    const int nwords = (f.num_actions + 31) >> 5;
    const int rounds = (nwords + 31) >> 5;                       // <= 32 words per lane
    if (a.mask_bits) {
        // bit-packed rows: independent coalesced loads, ten in flight per lane (a plain loop waits out the full memory
        // latency once per 128 B)
        const uint32_t *rw = reinterpret_cast<const uint32_t *>(row);
        for (int r0 = 0; r0 < rounds; r0 += 10) {
            uint32_t v[10];
#pragma unroll
            for (int k = 0; k < 10; ++k) { const int g = ((r0 + k) << 5) + lane; v[k] = (r0 + k < rounds && g < nwords) ? __ldg(rw + g) : 0u; }
#pragma unroll
            for (int k = 0; k < 10; ++k) if (r0 + k < rounds) words[((r0 + k) << 5) + lane] = v[k];
        }
    }
    for (int r = 0; r < (a.mask_bits ? 0 : rounds); ++r) {
        const int g = (r << 5) + lane;
        uint32_t w = 0u;
        if (g < nwords) {
            const uint64_t *p8 = reinterpret_cast<const uint64_t *>(row + 32 * g);      // rows are 128 B aligned and padded
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int byte0 = 32 * g + 8 * k;
                uint64_t b = byte0 + 8 <= f.mask_stride ? p8[k] : 0ULL;
                if (byte0 + 8 > f.num_actions) {                                        // drop padding bytes past A
                    const int keep = f.num_actions - byte0;
                    b = keep <= 0 ? 0ULL : (b & ((1ULL << (8 * keep)) - 1ULL));
                }
                w |= pack8(b) << (8 * k);
            }
        }
        words[g] = w;
    }
    __syncwarp();
    int mine = 0;
    for (int j = 0; j < rounds; ++j) { const int g = lane * rounds + j; if (g < nwords) mine += __popc(words[g]); }
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(kAll, incl, d); if (lane >= d) incl += v; }
    const int n = __shfl_sync(kAll, incl, 31);
    int e0 = 0;
    if (lane == 0) e0 = atomicAdd(&f.counters[1], n);
    e0 = __shfl_sync(kAll, e0, 0);
    if (e0 + n > f.edge_capacity) { if (lane == 0) f.counters[2] = 1; if (lane < P) score[lane] = 0.0; return; }
    if (lane == 0) { f.node_edge0[target] = e0; f.node_nedge[target] = n; f.node_sum_n[target] = 0.0; f.node_uniform[target] = a.prior_dtype == 0; }
    // Two phases so that the edge arrays are written with coalesced warp stores: (1) every lane compacts the ids of its
    // contiguous chunk of mask words into shared memory at its prefix position (ascending id order falls out of the
    // chunk order), (2) the warp sweeps the n edges 32 at a time.  Writing edges straight from the per-lane chunks put
    // 32 scattered addresses into every store instruction and left lanes with empty chunks idle (it was 60-70 % of a
    // simulation at B >= 4,096; 301 -> 120 us per launch at B = 16,384).  One bump-counter atomic per tree is not the
    // cost: claiming node / edge space once per block of 8 trees instead measured slower (137 us).
    const double uni = n > 0 ? __ddiv_rn(1.0, static_cast<double>(n)) : 0.0;
    int done_e = 0;                                        // edges already written (n exceeds kMaxIds only in theory)
    const int skip = incl - mine;                          // this lane's first position in id order
    if (n <= kMaxIds) {
        // the usual case: everything fits one sweep, the compaction loop is as short as it gets (it was 60 % of the
        // kernel's instructions with the window test inside)
        uint16_t *dst = ids + skip;
        const int g0 = lane * rounds, g1 = min(g0 + rounds, nwords);
        for (int g = g0; g < g1; ++g) {
            uint32_t w = words[g];
            const int base = g << 5;
            while (w) {
                *dst++ = static_cast<uint16_t>(base + __ffs(w) - 1);
                w &= w - 1;
            }
        }
        __syncwarp();
        for (int i = lane; i < n; i += 32) {
            const int id = ids[i];
            const int e = e0 + i;
            f.edge_action[e] = id; f.edge_child[e] = -1; f.edge_n[e] = 0.0; f.edge_q[e] = 0.0;
            if (a.prior_dtype == 1) f.edge_p[e] = static_cast<double>(reinterpret_cast<const float *>(a.prior)[t * a.prior_stride + id]);
            else if (a.prior_dtype == 2) f.edge_p[e] = reinterpret_cast<const double *>(a.prior)[t * a.prior_stride + id];
        }
        done_e = n;
    }
    while (done_e < n) {
        int pos = skip;
        for (int j = 0; j < rounds; ++j) {
            const int g = lane * rounds + j;
            if (g >= nwords) break;
            uint32_t w = words[g];
            while (w) {
                const int id = (g << 5) + __ffs(w) - 1;
                w &= w - 1;
                const int rel = pos - done_e;
                if (rel >= 0 && rel < kMaxIds) ids[rel] = static_cast<uint16_t>(id);
                ++pos;
            }
        }
        __syncwarp();
        const int m = min(n - done_e, kMaxIds);
        for (int i = lane; i < m; i += 32) {
            const int id = ids[i];
            const int e = e0 + done_e + i;
            double p = uni;
            if (a.prior_dtype == 1) p = static_cast<double>(reinterpret_cast<const float *>(a.prior)[t * a.prior_stride + id]);
            else if (a.prior_dtype == 2) p = reinterpret_cast<const double *>(a.prior)[t * a.prior_stride + id];
            f.edge_action[e] = id; f.edge_child[e] = -1; f.edge_n[e] = 0.0; f.edge_q[e] = 0.0;
            if (a.prior_dtype != 0) f.edge_p[e] = p;
        }
        __syncwarp();
        done_e += m;
    }
    if (lane < P) score[lane] = a.value != nullptr ? a.value[t * P + lane] : 0.0;
}

// running-mean backup along the recorded path (mcts.py:53-57); the path's edges belong to distinct nodes, so the
// levels are independent: lane d takes level d
__device__ __forceinline__ void backup_tree(const blk_puct_forest &f, int t, int lane) {
    const double *score = f.scores + static_cast<int64_t>(t) * f.num_players;
    const int32_t *path = f.path + static_cast<int64_t>(t) * f.max_depth;
    const int32_t *pnode = f.path_node + static_cast<int64_t>(t) * f.max_depth;
    const int len = f.path_len[t];
    for (int d = lane; d < len; d += 32) {
        const int e = path[d];
        const int child = f.edge_child[e];
        if (child < 0) continue;                                   // capacity overflow: flagged in counters[2]
        const double val = score[f.node_mover[child]];
        const double n = f.edge_n[e], q = f.edge_q[e];
        f.edge_q[e] = __ddiv_rn(__dadd_rn(__dmul_rn(n, q), val), __dadd_rn(n, 1.0));
        f.edge_n[e] = __dadd_rn(n, 1.0);
        f.node_sum_n[pnode[d]] = __dadd_rn(f.node_sum_n[pnode[d]], 1.0);     // integer-valued: exact
    }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32) puct_expand_kernel(ExpandArgs a) {
    __shared__ uint32_t s_words[kWarpsPerBlock][1024];
    __shared__ uint16_t s_ids[kWarpsPerBlock][kMaxIds];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * kWarpsPerBlock + warp;
    if (t >= a.f.num_trees) return;
    expand_tree(a, s_words[warp], s_ids[warp], t, lane);
    if (a.fuse_backup) {
        // the same warp wrote the score vector and the new child link: order them before the path walk reads them
        __syncwarp();
        backup_tree(a.f, t, lane);
        // this simulation's B pool slots are taken; the counter itself moves in the next select / advance launch, when no
        // warp of this launch can still be reading it
        if (t == 0 && lane == 0) a.f.counters[5] = 1;
    }
}

__global__ void puct_backup_kernel(blk_puct_forest f) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= f.num_trees) return;
    const double *score = f.scores + static_cast<int64_t>(t) * f.num_players;
    const int32_t *path = f.path + static_cast<int64_t>(t) * f.max_depth;
    const int32_t *pnode = f.path_node + static_cast<int64_t>(t) * f.max_depth;
    for (int d = f.path_len[t] - 1; d >= 0; --d) {
        const int e = path[d];
        const int child = f.edge_child[e];
        if (child < 0) continue;                                   // capacity overflow: flagged in counters[2]
        const double val = score[f.node_mover[child]];
        const double n = f.edge_n[e], q = f.edge_q[e];
        f.edge_q[e] = __ddiv_rn(__dadd_rn(__dmul_rn(n, q), val), __dadd_rn(n, 1.0));
        f.edge_n[e] = __dadd_rn(n, 1.0);
        f.node_sum_n[pnode[d]] = __dadd_rn(f.node_sum_n[pnode[d]], 1.0);     // integer-valued: exact
    }
    if (t == 0) f.counters[4] += f.num_trees;      // this simulation's B pool slots are taken (expand has finished)
}

// root <- child of the root edge carrying `action`; children that do not exist yet are requested as NEED_STEP
__global__ void puct_advance_kernel(blk_puct_forest f, const int32_t *actions) {
    apply_pending_slot_advance(f);
    const int t = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= f.num_trees) return;
    const int node = f.root[t];
    const int act = actions[t];
    const int e0 = f.node_edge0[node], n = f.node_nedge[node];
    int found = -1;
    for (int i = lane; i < n && e0 >= 0; i += 32) if (f.edge_action[e0 + i] == act) found = e0 + i;
#pragma unroll
    for (int d = 16; d; d >>= 1) found = max(found, __shfl_xor_sync(kAll, found, d));
    if (lane != 0) return;
    f.path_len[t] = 0;
    f.leaf_node[t] = node;
    f.src_slot[t] = f.node_state[node];
    if (act < 0 || found < 0) {                                   // tree stays where it is (finished game / unknown action)
        f.status[t] = BLK_PUCT_TERMINAL; f.leaf_edge[t] = -1; f.step_action[t] = BLK_ACTION_NONE;
        if (act >= 0) f.counters[3] = 1;
        return;
    }
    const int child = f.edge_child[found];
    if (child >= 0) { f.root[t] = child; f.status[t] = BLK_PUCT_TERMINAL; f.leaf_edge[t] = -1; f.step_action[t] = BLK_ACTION_NONE; }
    else { f.status[t] = BLK_PUCT_NEED_STEP; f.leaf_edge[t] = found; f.step_action[t] = act; }
}

// most visited root action per tree (first maximum, players/mcts_player.py:19-20); -1 when the root has no edges
__global__ void puct_best_kernel(blk_puct_forest f, int32_t *best_action, double *best_visits) {
    const int t = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= f.num_trees) return;
    const int node = f.root[t];
    const int e0 = f.node_edge0[node], n = e0 < 0 ? 0 : f.node_nedge[node];
    double best = -1.0;
    int besti = 0x7fffffff;
    for (int i = lane; i < n; i += 32) {
        const double v = f.edge_n[e0 + i];
        if (v > best) { best = v; besti = i; }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        const double ob = __shfl_xor_sync(kAll, best, d);
        const int oi = __shfl_xor_sync(kAll, besti, d);
        if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
    if (lane == 0) {
        best_action[t] = n > 0 ? f.edge_action[e0 + besti] : -1;
        if (best_visits != nullptr) best_visits[t] = n > 0 ? best : 0.0;
    }
}

thread_local std::string g_puct_err;
int puct_fail(const char *msg) { g_puct_err = msg; return BLK_ERR_ARG; }
int puct_launch_check() {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_puct_err = cudaGetErrorString(e); return BLK_ERR_CUDA; }
    return BLK_OK;
}

}  // namespace

extern "C" {

const char *blk_puct_last_error(void) { return g_puct_err.c_str(); }

int blk_puct_select(const blk_puct_forest *f, double cpuct, int32_t epsilon_fix, void *stream) {
    if (!f || f->num_trees < 0) return puct_fail("bad forest");
    if (f->num_trees == 0) return BLK_OK;
    const int grid = (f->num_trees + kWarpsPerBlock - 1) / kWarpsPerBlock;
    puct_select_kernel<<<grid, kWarpsPerBlock * 32, 0, static_cast<cudaStream_t>(stream)>>>(*f, cpuct, epsilon_fix);
    return puct_launch_check();
}

int blk_puct_expand(const blk_puct_forest *f, const blk_puct_expand_args *x, void *stream) {
    if (!f || !x) return puct_fail("null argument");
    if (f->num_trees == 0) return BLK_OK;
    if ((f->num_actions + 31) / 32 > 1024) return puct_fail("action space too large for the expand kernel");
    if (f->mask_stride % 8 != 0) return puct_fail("mask_stride must be a multiple of 8");
    ExpandArgs a;
    a.f = *f; a.new_slot_base = x->new_slot_base; a.state_words = x->state_words; a.meta_word = x->meta_word;
    a.attach_only = x->attach_only; a.new_states = x->new_states; a.pool = x->pool; a.mask = x->mask; a.flags = x->flags;
    a.mask_bits = x->mask_bits; a.mask_stride_words = x->mask_stride_words;
    a.terminal = x->terminal; a.prior = x->prior; a.prior_dtype = x->prior_dtype; a.prior_stride = x->prior_stride;
    a.value = x->value; a.fuse_backup = x->fuse_backup;
    const int grid = (f->num_trees + kWarpsPerBlock - 1) / kWarpsPerBlock;
    puct_expand_kernel<<<grid, kWarpsPerBlock * 32, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return puct_launch_check();
}

int blk_puct_backup(const blk_puct_forest *f, void *stream) {
    if (!f) return puct_fail("null argument");
    if (f->num_trees == 0) return BLK_OK;
    puct_backup_kernel<<<(f->num_trees + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(*f);
    return puct_launch_check();
}

int blk_puct_best(const blk_puct_forest *f, int32_t *best_action, double *best_visits, void *stream) {
    if (!f || !best_action) return puct_fail("null argument");
    if (f->num_trees == 0) return BLK_OK;
    const int grid = (f->num_trees + kWarpsPerBlock - 1) / kWarpsPerBlock;
    puct_best_kernel<<<grid, kWarpsPerBlock * 32, 0, static_cast<cudaStream_t>(stream)>>>(*f, best_action, best_visits);
    return puct_launch_check();
}

int blk_puct_advance(const blk_puct_forest *f, const int32_t *actions, void *stream) {
    if (!f || !actions) return puct_fail("null argument");
    if (f->num_trees == 0) return BLK_OK;
    const int grid = (f->num_trees + kWarpsPerBlock - 1) / kWarpsPerBlock;
    puct_advance_kernel<<<grid, kWarpsPerBlock * 32, 0, static_cast<cudaStream_t>(stream)>>>(*f, actions);
    return puct_launch_check();
}

}  // extern "C"
