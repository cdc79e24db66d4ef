// blk_search.cuh -- whole PUCT simulations inside ONE kernel (blk_puct_search) and board-keyed tree reuse (blk_puct_reroot).
//
// The lockstep forest (blk_puct.cu) spends three launches per simulation (select -> blk_step -> expand + backup), which
// is the right shape when a torch net has to evaluate the leaves in between.  The reference's "mcts" arena player needs no
// net: its DumbNet is the uniform prior with zero value (blokus_rl/models/dumbnet.py:14-21, compare_arena.py:87-95,
// players/mcts_player.py:15-22 -- ONE tree, 10-200 simulations per move).  Here a warp owns a tree and runs the whole
// simulation loop of blokus_rl/alphazero/mcts.py:13-71 on its own: descent (float64 UCB, first maximum), the env
// transition of the opened edge with the warp-per-env machinery of the step kernel (board rows in registers, fields in
// shared memory), expansion of the new node straight from the staged fields (no mask ever touches HBM), backup.  No
// launch per level, no host in the loop: a single search stops being launch-latency-bound.
//
// Nodes are keyed by BOARD CELLS like the reference's dict (mcts.py:37, blokus_wrapper.py:217-218): an open-addressing
// table maps hash(tree, board rows) -> node, verified word by word.  A simulation that opens an edge onto a board the
// tree already knows links the edge to that node and CONTINUES its descent there, exactly what the reference's recursion
// does when `hashed_s in self.tree`.
//
// warps_per_tree == 1 reproduces the reference's visit counts, Q values and per-simulation score vectors (tests:
// golden vectors of the unmodified reference file).  warps_per_tree > 1 is leaf-parallel search with virtual loss: several
// warps descend the same tree at once (claims by atomicCAS; an edge keeps the SUM of its backed-up values instead of their mean,
// so a backup is atomic adds, no lock) -- faster for ONE tree, not the reference's visit order.
#pragma once
#include "blk_kernels.cuh"

namespace blk {

constexpr int kSearchTrees = 4;        // trees per block when every tree has one warp
constexpr int kSearchMaxIds = 1024;    // legal ids compacted per sweep (20x20 positions have <= ~800)
constexpr int kSearchMaxDepth = 96;    // >= 4 * 21 placements + 1
// per-warp shared memory: fields (+ scratch) | compacted ids | path: edge and node (int each) | the simulation's score vector.
// At 20x20 that is 9,680 B per tree: 4 trees + the 19 KB of tables = 57.7 KB per block, FOUR blocks per SM.  (Carrying the
// edges' N / Q and the nodes' sum N along the path in shared memory, so that the backup needs no global load, was measured:
// +6 % for one tree, but 12.4 KB per tree leaves three blocks per SM and 4,096 trees ran 23 % slower.)
__host__ __device__ constexpr int search_warp_bytes(int warp_smem) { return warp_smem + 2 * kSearchMaxIds + 8 * kSearchMaxDepth + 64; }

struct SearchParams {
    blk_puct_forest f;
    blk_puct_search_args a;
    const unsigned char *tables;
    TableLayout t;
    Geometry g;
    const uint32_t *reroot_states;      // blk_puct_reroot only
};

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xFF51AFD7ED558CCDULL; x ^= x >> 33; x *= 0xC4CEB9FE1A85EC53ULL; x ^= x >> 33;
    return x;
}
__device__ __forceinline__ uint64_t shfl_xor64(uint64_t v, int d) {
    const uint32_t lo = __shfl_xor_sync(kAllLanes, static_cast<uint32_t>(v), d);
    const uint32_t hi = __shfl_xor_sync(kAllLanes, static_cast<uint32_t>(v >> 32), d);
    return (static_cast<uint64_t>(hi) << 32) | lo;
}
// hash of (tree, board cells): every lane mixes its row of the four bitboards with a lane-dependent constant
__device__ __forceinline__ uint64_t board_hash(const EnvRegs &e, int tree, const Dims &g, int lane) {
    uint64_t h = 0ULL;
    if (lane < g.N) {
        const uint64_t a = static_cast<uint64_t>(e.own0) | (static_cast<uint64_t>(e.own1) << 32);
        const uint64_t b = static_cast<uint64_t>(e.own2) | (static_cast<uint64_t>(e.own3) << 32);
        h = mix64(a + 0x9E3779B97F4A7C15ULL * static_cast<uint64_t>(2 * lane + 1)) ^
            mix64(b ^ (0xC2B2AE3D27D4EB4FULL * static_cast<uint64_t>(2 * lane + 3)));
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) h ^= shfl_xor64(h, d);
    return mix64(h ^ (static_cast<uint64_t>(static_cast<uint32_t>(tree)) * 0xD6E8FEB86659FD93ULL));
}

__device__ __forceinline__ int ld_volatile(const int32_t *p) { return *reinterpret_cast<const volatile int32_t *>(p); }
__device__ __forceinline__ double ld_volatile(const double *p) { return *reinterpret_cast<const volatile double *>(p); }

// Node filed under this board in tree `tree`, or -1.  With new_node >= 0 the node is inserted when the board is unknown
// (returns -1), unless another warp of the same tree files the same board first (returns that node).
__device__ __forceinline__ int table_find_or_insert(const blk_puct_forest &f, const uint32_t *pool, uint64_t h, int tree,
                                                    const EnvRegs &e, const Dims &g, int lane, int new_node) {
    const uint32_t cap_mask = static_cast<uint32_t>(f.hash_capacity - 1);
    uint32_t slot = static_cast<uint32_t>(h >> 17) & cap_mask;
    const int N = g.N, P = g.P, sw = P * N + P + 4;
    for (int probe = 0; probe < f.hash_capacity; ++probe) {
        int v = 0;
        if (lane == 0) v = ld_volatile(f.hash_table + slot);
        v = __shfl_sync(kAllLanes, v, 0);
        if (v == 0) {
            if (new_node < 0) return -1;
            int old = 0;
            if (lane == 0) old = atomicCAS(f.hash_table + slot, 0, new_node + 1);
            old = __shfl_sync(kAllLanes, old, 0);
            if (old == 0) return -1;                                   // filed
            v = old;                                                   // somebody took the slot meanwhile: look at it
        }
        const int cand = v - 1;
        if (cand != new_node && f.node_hash[cand] == h && f.node_tree[cand] == tree) {
            const uint32_t *s = pool + static_cast<int64_t>(f.node_state[cand]) * sw;
            bool eq = true;
            if (lane < N) {
                eq = s[lane] == e.own0 && s[N + lane] == e.own1;
                if (P > 2) eq = eq && s[2 * N + lane] == e.own2 && s[3 * N + lane] == e.own3;
            }
            if (__all_sync(kAllLanes, eq)) return cand;
        }
        slot = (slot + 1) & cap_mask;
    }
    if (lane == 0) f.counters[2] = 1;                                  // table full
    return -1;
}

// New node for the state in `e` (pool slot + node arrays); not yet in the table, not yet expanded.  -1 = out of capacity.
__device__ __forceinline__ int node_create(const blk_puct_forest &f, uint32_t *pool, const EnvRegs &e, uint64_t h, int tree,
                                           bool done, float tval, const Dims &g, int lane) {
    // a forest searched by these kernels keeps node i's state in pool slot i: one bump counter, one atomic
    int node = 0;
    if (lane == 0) node = atomicAdd(&f.counters[0], 1);
    node = __shfl_sync(kAllLanes, node, 0);
    const int slot = node;
    if (node >= f.node_capacity) { if (lane == 0) f.counters[2] = 1; return -1; }
    const int P = g.P, sw = P * g.N + P + 4;
    env_store(e, pool + static_cast<int64_t>(slot) * sw, g, lane);
    if (lane < P) f.node_term_value[static_cast<int64_t>(node) * P + lane] = done ? static_cast<double>(tval) : 0.0;
    if (lane == 0) {
        f.node_state[node] = slot;
        f.node_mover[node] = static_cast<int8_t>(e.meta & 15u);
        f.node_edge0[node] = -1;
        f.node_nedge[node] = 0;
        f.node_sum_n[node] = 0.0;
        f.node_terminal[node] = done ? 1 : 0;
        f.node_uniform[node] = 1;
        f.node_front[node] = 0;
        f.node_hash[node] = h;
        f.node_tree[node] = tree;
    }
    __syncwarp();
    return node;
}

// One edge per legal action of the mover whose fields are staged in `fld` (ascending ids: field order is id order).
// Publishes node_edge0 last.  Returns false when the edge arrays are full.
template <int kN, bool kFence>
__device__ __forceinline__ bool expand_from_fields(const blk_puct_forest &f, int node, const uint32_t *fld, uint16_t *ids,
                                                   const SmemTables &tb, int nf, int lane) {
    const int per = fields_per_lane<kN>(nf);
    const int mine = count_field_chunk<kN>(fld, nf, per, lane);       // lane l owns fields [l * chunk, + len): 13 LDS.128 at N = 20
    const int chunk = kN == 20 ? 52 : per;
    const int i0 = lane * chunk;
    const int len = kN == 20 ? (lane == 31 ? 53 : 52) : max(0, min(per, nf - i0));
    const int incl = warp_incl_scan(mine, lane);
    const int n = __shfl_sync(kAllLanes, incl, 31);
    int e0 = 0;
    if (lane == 0) e0 = atomicAdd(&f.counters[1], n);
    e0 = __shfl_sync(kAllLanes, e0, 0);
    if (e0 + n > f.edge_capacity) { if (lane == 0) f.counters[2] = 1; return false; }
    const int skip = incl - mine;                          // this lane's first position in id order
    for (int done_e = 0; done_e < n; done_e += kSearchMaxIds) {
        int pos = skip - done_e;
        // ids of one field, appended at `pos` (window [0, kSearchMaxIds) of this sweep)
#define BLK_EMIT_FIELD(w_, i_)                                                                        \
        {                                                                                             \
            uint32_t w = (w_);                                                                        \
            if (w) {                                                                                  \
                const int base = tb.foff[i_];                                                         \
                do {                                                                                  \
                    if (pos >= 0 && pos < kSearchMaxIds) ids[pos] = static_cast<uint16_t>(base + __ffs(w) - 1); \
                    ++pos;                                                                            \
                    w &= w - 1;                                                                       \
                } while (w);                                                                          \
            }                                                                                         \
        }
        if (mine > 0) {
            if (kN == 20) {                                 // four fields per load; most quads of a mid-game position are empty
                const uint4 *f4 = reinterpret_cast<const uint4 *>(fld) + 13 * lane;
#pragma unroll 1
                for (int j = 0; j < 13; ++j) {
                    const uint4 x = f4[j];
                    if ((x.x | x.y | x.z | x.w) != 0u) {
                        BLK_EMIT_FIELD(x.x, i0 + 4 * j) BLK_EMIT_FIELD(x.y, i0 + 4 * j + 1)
                        BLK_EMIT_FIELD(x.z, i0 + 4 * j + 2) BLK_EMIT_FIELD(x.w, i0 + 4 * j + 3)
                    }
                }
                if (lane == 31) BLK_EMIT_FIELD(fld[1664], 1664)
            } else {
                for (int t = 0; t < len; ++t) BLK_EMIT_FIELD(fld[i0 + t], i0 + t)
            }
        }
#undef BLK_EMIT_FIELD
        __syncwarp();
        const int m = min(n - done_e, kSearchMaxIds);
        for (int i = lane; i < m; i += 32) {
            const int e = e0 + done_e + i;
            f.edge_action[e] = ids[i]; f.edge_n[e] = 0.0; f.edge_q[e] = 0.0; f.edge_child[e] = -1;
        }
        __syncwarp();
    }
    if (lane == 0) { f.node_nedge[node] = n; f.node_sum_n[node] = 0.0; }
    if (kFence) __threadfence();
    if (lane == 0) *reinterpret_cast<volatile int32_t *>(f.node_edge0 + node) = e0;
    __syncwarp();
    return true;
}

// Next mover of `e` after a placement by `mover` (R8 auto-skip, R9 game over); leaves the new mover's fields in `fld`.
template <int kN, int kP>
__device__ __forceinline__ bool resolve_next_mover(EnvRegs &e, int mover, const Dims &g, uint32_t *fld, int lane) {
    int cand = mover;
#pragma unroll 1
    for (int tries = g.P; tries > 0; --tries) {
        cand = (cand + 1 == g.P) ? 0 : cand + 1;
        uint32_t fr0, dg0;
        prep_rows(e, cand, g, lane, fr0, dg0);
        const uint32_t acc = eval_fields<true>(fr0, dg0, sel4(e.inv0, e.inv1, e.inv2, e.inv3, cand), fld, g.N, lane);
        if (__any_sync(kAllLanes, acc != 0u)) {
            e.meta = (e.meta & ~15u) | static_cast<uint32_t>(cand);
            __syncwarp();
            return true;
        }
    }
    e.meta |= 1u << 4;                                     // nobody can move: done, mover stays = last mover
    return false;
}

// Leaf value: zeros (DumbNet) or the mean 3/1/-1 vector of k uniform-random playouts from the node's state.
template <int kN, int kP>
__device__ __forceinline__ double leaf_value(const SearchParams &sp, const EnvRegs &e0, int node, const SmemTables &tb,
                                             const Dims &g, uint32_t *fld, int lane) {
    const int k = sp.a.playouts_per_leaf;
    if (k <= 0) return 0.0;
    float sum = 0.f;
    for (int j = 0; j < k; ++j) {
        EnvRegs e = e0;
        bool over;
        const uint32_t game = static_cast<uint32_t>(node) * static_cast<uint32_t>(k) + static_cast<uint32_t>(j);
        playout_game<kN, kP>(e, tb, sp.g, g, fld, lane, static_cast<uint32_t>(sp.a.seed),
                             static_cast<uint32_t>(sp.a.seed >> 32) ^ game, 2u, -1, nullptr, 0, over);
        int fs;
        sum += terminal_value(e, g, lane, fs);
    }
    return static_cast<double>(sum / static_cast<float>(k));
}

// ---------------------------------------------------------------------------------------------------------------------
// blk_puct_search
// ---------------------------------------------------------------------------------------------------------------------
// Selection cost.  With the uniform prior every edge that was never selected has the same score c * P * sqrt(..) / 1, and
// the first maximum among them is the one with the lowest index -- so the edges selected so far always form a PREFIX
// [0, front) of the node's edge list (by induction: a never-selected edge can only win as the lowest-indexed one).  The
// argmax therefore runs over the front + 1 edges [0, min(front + 1, n)) instead of all 58-760: exact, and a node deep in
// the tree costs one pass over a handful of edges.
template <int kN, int kP, bool kParallel>
__global__ void __launch_bounds__(kParallel ? 512 : kSearchTrees * 32) puct_search_kernel(const SearchParams sp) {
    extern __shared__ __align__(128) unsigned char smem[];
    const Geometry &gg = sp.g;
    const Dims g = make_dims<kN, kP>(gg);
    const blk_puct_forest &f = sp.f;
    unsigned char *tab = smem;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + sp.t.bytes);
    int *blk_ctl = reinterpret_cast<int *>(smem + sp.t.bytes + 8);        // [1] simulations handed out
    unsigned char *scratch = smem + sp.t.bytes + 16;
    if (kParallel && threadIdx.x == 0) { blk_ctl[0] = 0; blk_ctl[1] = 0; }
    tma_load_tables(tab, sp.tables, sp.t.bytes, bar);
    const SmemTables tb = make_tables(tab, sp.t);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int N = g.N, P = g.P, sw = P * N + P + 4;
    const int nf = geo_nf<kN>(gg);
    const int fld_words = geo_fld_words<kN>(gg);
    const int per_warp = search_warp_bytes(gg.warp_smem);
    unsigned char *mine_smem = scratch + static_cast<size_t>(warp) * per_warp;
    uint32_t *fld = reinterpret_cast<uint32_t *>(mine_smem);
    uint16_t *ids = reinterpret_cast<uint16_t *>(mine_smem + gg.warp_smem);
    int *spath = reinterpret_cast<int *>(mine_smem + gg.warp_smem + 2 * kSearchMaxIds);          // edges of the path
    int *snode = spath + kSearchMaxDepth;                                                         // node of each path edge
    double *sscore = reinterpret_cast<double *>(mine_smem + gg.warp_smem + 2 * kSearchMaxIds + 8 * kSearchMaxDepth);
    for (int i = nf + lane; i < fld_words; i += 32) fld[i] = 0u;
    const int t = kParallel ? static_cast<int>(blockIdx.x) : static_cast<int>(blockIdx.x) * kSearchTrees + warp;
    if (t >= f.num_trees) return;
    uint32_t *pool = sp.a.pool;
    const double vloss = sp.a.virtual_loss != 0.0 ? sp.a.virtual_loss : 1.0;
    const int max_depth = min(f.max_depth, kSearchMaxDepth);
    // one warp per tree: nobody else touches this tree, plain loads and stores in program order are enough; several
    // warps per tree: shared words are read volatile and published behind a fence
#define BLK_LD(p) (kParallel ? ld_volatile(p) : *(p))
#define BLK_PUBLISH(p, v)                                             \
    do {                                                              \
        if (kParallel) __threadfence();                               \
        if (lane == 0) *reinterpret_cast<volatile int32_t *>(p) = (v); \
        __syncwarp();                                                 \
    } while (0)

#pragma unroll 1
    for (int sim = 0;; ++sim) {
        if (kParallel) {
            int mine_sim = 0;
            if (lane == 0) mine_sim = atomicAdd(&blk_ctl[1], 1);
            if (__shfl_sync(kAllLanes, mine_sim, 0) >= sp.a.num_sims) break;
        } else if (sim >= sp.a.num_sims) {
            break;
        }
        int node = BLK_LD(f.root + t);
        int len = 0, last_child = -1;
        double my_score = 0.0;                              // lane q < P: score of player q for this simulation
#pragma unroll 1
        for (int depth = 0;; ++depth) {
            // the node's header: independent loads, issued together (one round trip to L2, not five)
            const int term = f.node_terminal[node];
            int e0 = BLK_LD(f.node_edge0 + node);
            const int n = f.node_nedge[node];
            const int front = BLK_LD(f.node_front + node);
            const double s = BLK_LD(f.node_sum_n + node);
            const int state_slot = f.node_state[node];
            if (lane < 3)                                   // the state is wanted only if an edge gets opened here: pull its
                asm volatile("prefetch.global.L1 [%0];" ::"l"(pool + static_cast<int64_t>(node) * sw + 32 * lane));   // lines into L1 meanwhile
            if (term) {                                     // terminal states are never expanded (mcts.py:60-62)
                if (lane < P) my_score = f.node_term_value[static_cast<int64_t>(node) * P + lane];
                break;
            }
            if (e0 < 0) {
                // unexpanded node (a fresh root): expand it in place
                if (kParallel) {
                    int old = -1;
                    if (lane == 0) old = (e0 == -1) ? atomicCAS(f.node_edge0 + node, -1, -2) : e0;
                    old = __shfl_sync(kAllLanes, old, 0);
                    if (old != -1) {                        // somebody else is expanding it: wait until it has edges (or was
                        while (ld_volatile(f.node_edge0 + node) == -2) __nanosleep(100);   // handed back unexpanded: -1)
                        __threadfence();
                        --depth;
                        continue;
                    }
                }
                EnvRegs e;
                env_load(e, pool + static_cast<int64_t>(state_slot) * sw, g, lane);
                uint32_t fr0, dg0;
                const int mv = static_cast<int>(e.meta & 15u);
                prep_rows(e, mv, g, lane, fr0, dg0);
                eval_fields<true, false>(fr0, dg0, sel4(e.inv0, e.inv1, e.inv2, e.inv3, mv), fld, N, lane);
                __syncwarp();
                if (!expand_from_fields<kN, kParallel>(f, node, fld, ids, tb, nf, lane)) {
                    // edge arrays full (flagged in counters[2]): hand the node back unexpanded so that nobody waits for it
                    if (kParallel) BLK_PUBLISH(f.node_edge0 + node, -1);
                    break;
                }
                my_score = leaf_value<kN, kP>(sp, e, node, tb, g, fld, lane);
                break;
            }
            if (len >= max_depth) { if (lane == 0) f.counters[2] = 1; break; }
            if (kParallel && n <= 0) {                      // edge0 was already visible, the edge count not yet: read the header again
                __threadfence();
                --depth;
                continue;
            }
            // ---- selection: U = c * P * sqrt(sum N + eps) / (1 + N), first maximum of Q + U (mcts.py:42-46) ----
            const double pu = __ddiv_rn(1.0, static_cast<double>(n));
            const double c = depth == 0 ? sp.a.cpuct : 1.0;
            const double sq = __dsqrt_rn(__dadd_rn(s, (sp.a.epsilon_fix || depth > 0) ? 1e-6 : 0.0));
            const double cps = __dmul_rn(__dmul_rn(c, pu), sq);          // (c * P) * sqrt(..): the reference's order
            const int m = min(n, front + 1);                // edges >= front were never selected: all tie, the first one wins
            double best = -1.0e300;
            int besti = 0x7fffffff;
            for (int i0 = lane; i0 < m; i0 += 128) {
                double en[4], eq[4];
                int vl[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int i = i0 + 32 * k;
                    const bool in = i < m;
                    en[k] = in ? BLK_LD(f.edge_n + e0 + i) : 0.0;
                    eq[k] = in ? BLK_LD(f.edge_q + e0 + i) : 0.0;
                    vl[k] = (kParallel && in) ? ld_volatile(f.edge_vl + e0 + i) : 0;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int i = i0 + 32 * k;
                    double nn = en[k], qq = eq[k];
                    if (kParallel) {
                        // leaf-parallel: edge_q holds the SUM of the backed-up values (so that a backup is two atomic adds, no
                        // lock); pending visits (virtual loss) count as visits that lost `vloss`
                        const double v = static_cast<double>(vl[k]);
                        const double den = nn + v;
                        qq = den > 0.0 ? (qq - v * vloss) / den : 0.0;
                        nn = den;
                    }
                    const double u = __ddiv_rn(cps, __dadd_rn(1.0, nn));
                    const double sc = __dadd_rn(qq, u);
                    if (i < m && sc > best) { best = sc; besti = i; }   // ascending i per lane: keeps the first maximum
                }
            }
#pragma unroll
            for (int d = 16; d; d >>= 1) {
                const double ob = __shfl_xor_sync(kAllLanes, best, d);
                const int oi = __shfl_xor_sync(kAllLanes, besti, d);
                if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
            }
            const int e = e0 + besti;
            int child = BLK_LD(f.edge_child + e);
            const int act = f.edge_action[e];               // needed when the edge is still closed: same round trip
            if (lane == 0) {
                spath[len] = e; snode[len] = node;
                if (besti >= front) {                       // the front edge was taken: the prefix grows
                    if (kParallel) atomicMax(f.node_front + node, besti + 1);
                    else f.node_front[node] = besti + 1;
                }
                if (kParallel) atomicAdd(f.edge_vl + e, 1);
            }
            ++len;
            if (child == -1) {
                if (kParallel) {
                    int old = 0;
                    if (lane == 0) old = atomicCAS(f.edge_child + e, -1, -2);
                    child = __shfl_sync(kAllLanes, old, 0);
                    if (child == -1) child = -3;            // ours to open
                } else {
                    child = -3;
                }
            }
            if (child == -2) {                              // another warp is opening this edge
                while ((child = ld_volatile(f.edge_child + e)) == -2) __nanosleep(100);
                __threadfence();
                if (child < 0) {                            // the opener gave up (node pool exhausted: flagged): so does this simulation
                    if (lane == 0) atomicSub(f.edge_vl + e, 1);
                    --len;
                    break;
                }
            }
            if (child >= 0) { node = last_child = child; continue; }
            // ---- open the edge: env transition of the parent's state under the edge's action ----
            EnvRegs es;
            env_load(es, pool + static_cast<int64_t>(state_slot) * sw, g, lane);
            const int mover = static_cast<int>(es.meta & 15u);
            uint32_t pm; int piece, ncells;
            bool legal = decode_action(act, tb, g, lane, pm, piece, ncells);
            if (legal) {
                uint32_t fr0, dg0;
                prep_rows(es, mover, g, lane, fr0, dg0);
                const bool avail = (sel4(es.inv0, es.inv1, es.inv2, es.inv3, mover) >> piece) & 1u;
                const bool bad = __any_sync(kAllLanes, (pm & ~fr0) != 0u);
                const bool touch = __any_sync(kAllLanes, (pm & dg0) != 0u);
                legal = avail && !bad && touch;
            }
            int fresh = -1, known = -1;
            bool alive = false;
            float tval = 0.f;
            if (legal) {
                apply_placement(es, mover, pm, piece, ncells);
                alive = resolve_next_mover<kN, kP>(es, mover, g, fld, lane);
                int fscore;
                if (!alive) tval = terminal_value(es, g, lane, fscore);
                const uint64_t h = board_hash(es, t, g, lane);
                known = table_find_or_insert(f, pool, h, t, es, g, lane, -1);
                if (known < 0) {
                    fresh = node_create(f, pool, es, h, t, !alive, tval, g, lane);
                    if (fresh >= 0) {
                        if (kParallel) {
                            if (lane == 0 && alive) f.node_edge0[fresh] = -2;        // ours to expand
                            __threadfence();
                        }
                        known = table_find_or_insert(f, pool, h, t, es, g, lane, fresh);   // -1: filed; >= 0: an equal board won the race
                    }
                }
            } else if (lane == 0) {
                f.counters[3] = 1;                          // cannot happen for edges built from a legal mask: flag it
            }
            if (known < 0 && fresh < 0) {                   // illegal action or node pool exhausted (flagged): drop the edge
                if (lane == 0) {
                    *reinterpret_cast<volatile int32_t *>(f.edge_child + e) = -1;
                    if (kParallel) atomicSub(f.edge_vl + e, 1);
                }
                --len;
                break;
            }
            if (known >= 0) {
                // the tree already has this board (another move order reached it): share its node and go on from there
                BLK_PUBLISH(f.edge_child + e, known);
                node = last_child = known;
                continue;
            }
            last_child = fresh;
            if (!alive) {
                BLK_PUBLISH(f.edge_child + e, fresh);
                if (lane < P) my_score = static_cast<double>(tval);
                break;
            }
            const bool ok = expand_from_fields<kN, kParallel>(f, fresh, fld, ids, tb, nf, lane);
            if (kParallel && !ok) BLK_PUBLISH(f.node_edge0 + fresh, -1);     // edge arrays full: nobody may wait for these edges
            BLK_PUBLISH(f.edge_child + e, fresh);
            if (ok) my_score = leaf_value<kN, kP>(sp, es, fresh, tb, g, fld, lane);
            break;
        }
        // ---- backup: Q <- (N * Q + v) / (N + 1) with v = the score of the player to move AFTER the edge (mcts.py:53-56) ----
        if (lane < P) sscore[lane] = my_score;
        __syncwarp();
        for (int d = lane; d < len; d += 32) {
            const int e = spath[d];
            const int child = d + 1 < len ? snode[d + 1] : last_child;   // the node the edge led to in THIS simulation
            if (kParallel) {
                // several warps back up through the same edges: sums, not means, so that atomic adds suffice
                const double val = sscore[f.node_mover[child]];
                atomicAdd(f.edge_q + e, val);
                atomicAdd(f.edge_n + e, 1.0);
                atomicAdd(f.node_sum_n + snode[d], 1.0);
                atomicSub(f.edge_vl + e, 1);
            } else {
                const double val = sscore[f.node_mover[child]];
                const double nn = f.edge_n[e], qq = f.edge_q[e];
                const double sn = f.node_sum_n[snode[d]];
                f.edge_q[e] = __ddiv_rn(__dadd_rn(__dmul_rn(nn, qq), val), __dadd_rn(nn, 1.0));
                f.edge_n[e] = __dadd_rn(nn, 1.0);
                f.node_sum_n[snode[d]] = __dadd_rn(sn, 1.0);              // integer-valued: exact
            }
        }
        __syncwarp();
        if (lane < P) f.scores[static_cast<int64_t>(t) * P + lane] = my_score;
        __syncwarp();
    }
#undef BLK_LD
#undef BLK_PUBLISH
}

// blk_puct_reroot: root of tree t <- the node filed under the board of states[t], or a fresh unexpanded node
template <int kN, int kP>
__global__ void __launch_bounds__(kSearchTrees * 32) puct_reroot_kernel(const SearchParams sp) {
    const Dims g = make_dims<kN, kP>(sp.g);
    const blk_puct_forest &f = sp.f;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = static_cast<int>(blockIdx.x) * kSearchTrees + warp;
    if (t >= f.num_trees) return;
    const int sw = g.P * g.N + g.P + 4;
    EnvRegs e;
    env_load(e, sp.reroot_states + static_cast<int64_t>(t) * sw, g, lane);
    const uint64_t h = board_hash(e, t, g, lane);
    int node = table_find_or_insert(f, sp.a.pool, h, t, e, g, lane, -1);
    if (node < 0) {
        const bool done = (e.meta >> 4) & 1u;
        int fscore;
        const float tval = done ? terminal_value(e, g, lane, fscore) : 0.f;
        node = node_create(f, sp.a.pool, e, h, t, done, tval, g, lane);
        if (node < 0) return;
        __threadfence();
        table_find_or_insert(f, sp.a.pool, h, t, e, g, lane, node);
    }
    if (lane == 0) { f.root[t] = node; f.path_len[t] = 0; f.status[t] = BLK_PUCT_TERMINAL; }
}

using SearchFn = void (*)(const SearchParams);
struct SearchKernelSet {
    SearchFn search[2];     // [0] one warp per tree (the reference's visit order), [1] leaf-parallel (a block per tree)
    SearchFn reroot;
};
template <int kN, int kP>
inline SearchKernelSet make_search_set() {
    SearchKernelSet k;
    k.search[0] = puct_search_kernel<kN, kP, false>;
    k.search[1] = puct_search_kernel<kN, kP, true>;
    k.reroot = puct_reroot_kernel<kN, kP>;
    return k;
}
SearchKernelSet search_kernels_20_4();
SearchKernelSet search_kernels_20_2();
SearchKernelSet search_kernels_14_4();
SearchKernelSet search_kernels_14_2();
SearchKernelSet search_kernels_7_2();
SearchKernelSet search_kernels_0_0();

}  // namespace blk
