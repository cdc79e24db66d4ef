/*
 * blokus_b200.h -- C ABI of the B200-native batched Blokus environment engine.
 *
 * The reference (KubiakJakub01/Blokus-RL) has no FFI for this path: its env engine is the Python /
 * Cython package colosseumrl.envs.blokus, called through duck-typed Python at
 * blokus_rl/colossumrl/blokus_wrapper.py.  Each entry point below names the reference call it
 * replaces; INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *  - every function returns 0 on success, a negative blk_status otherwise; blk_last_error() gives a
 *    thread-local message.  No exception crosses the ABI.
 *  - the caller (PyTorch) owns every device buffer; the engine owns only its constant tables.
 *  - all work is enqueued asynchronously on the cudaStream_t passed as `void* stream` (NULL = legacy
 *    default stream); nothing synchronises.
 *  - an engine handle belongs to one device; its entry points restore the caller's current CUDA device before they
 *    return and may be called from several host threads.  Launches on different streams may overlap; a launch captured
 *    into a CUDA graph keeps working for every replay of that graph (do not run two instantiations of the SAME
 *    captured graph concurrently).
 *  - there is no CPU fallback: a missing/failed CUDA device is an error.
 *
 * State format (uint32 words per env, `state_words` = P*N + P + 4; 88 words = 352 B at 20x20/4p):
 *    [q*N + y]         row y of player q, bit x = column x
 *    [P*N + q]         inventory of player q, bit i = piece i still in hand
 *    [P*N + P]         meta: mover (bits 0-3) | done << 4 | lastmono(q) << (8+q) | ply << 16
 *    [P*N + P + 1]     game counter (number of auto-resets so far)
 *    [P*N + P + 2..3]  placed-squares score of players 0..3 as int16, little endian
 */
#ifndef BLOKUS_B200_H
#define BLOKUS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ABI history: 1 = first release of this round; 2 = blk_step_args grew `obs` (fused observation output) and `state_index`
 * (step states out of a pool), blk_rollout_args grew `options`, BLK_OPT_WARP_KERNELS, blk_puct_forest grew `node_uniform`
 * and a 6th counter, blk_puct_expand_args grew `fuse_backup`.  All additions are trailing fields: zero-initialised structs keep
 * their version-1 meaning.
 * 3 = blk_step_args grew csr_cursor / csr_offset (compact index lists); blk_puct_forest grew the board-keyed node table (hash_table, hash_capacity, node_hash, node_tree), edge_vl and node_front;
 * blk_puct_search / blk_puct_reroot (whole simulations inside one kernel); work-queue slots are per stream / per graph capture;
 * every entry point restores the caller's current device. */
#define BLK_ABI_VERSION 3

typedef enum {
    BLK_OK = 0,
    BLK_ERR_ARG = -1,      /* bad argument / unsupported configuration */
    BLK_ERR_CUDA = -2,     /* CUDA runtime error (message has the detail) */
    BLK_ERR_NODEV = -3     /* no usable sm_100 device */
} blk_status;

/* BLK_MASK_INDICES: the sparse form of the same mask -- the ascending legal action ids as uint16 (the list the
 * reference's PPO path reads per env: envs.get_attr("ai_possible_indexes"), blokus_rl/ppo/trainer.py:385);
 * legal_count[i] says how many entries of row i are valid (it is required in this format). */
typedef enum { BLK_MASK_NONE = 0, BLK_MASK_BITS = 1, BLK_MASK_BYTES = 2, BLK_MASK_INDICES = 3 } blk_mask_format;

/* score_rule: 0 = squares placed (default); 1 = squares placed + 15 when all 21 pieces are placed,
 * + 5 more when the monomino went last (SURVEY.md Appendix A, R10 is OPEN in the reference). */
typedef struct {
    int32_t board_size;    /* N: 5..20 */
    int32_t num_players;   /* P: 2 or 4 */
    int32_t score_rule;
    int32_t device;        /* CUDA device ordinal */
} blk_config;

typedef struct {
    int32_t abi_version;
    int32_t board_size, num_players;
    int32_t num_actions;   /* A: 30433 at N=20, 2522 at N=7 */
    int32_t num_pieces;    /* 21 */
    int32_t num_orients;   /* 91 */
    int32_t num_fields;    /* (orientation, anchor row) pairs: 1665 at N=20 */
    int32_t state_words;   /* uint32 words per env state */
    int32_t mask_words;    /* uint32 words of a bit-packed mask row, padded to 4 (952 at N=20) */
    int32_t mask_bytes;    /* bytes of a byte-mask row padded to 128 (30464 at N=20) */
    int32_t sm_count;
} blk_info;

typedef struct blk_engine blk_engine;

/* BLK_OPT_WARP_KERNELS: use the warp-per-env kernels even where a board size has thread-per-env kernels (N <= 7);
 * same results, slower -- kept for cross-checking the two implementations against each other. */
enum { BLK_OPT_AUTO_RESET = 1, BLK_OPT_WARP_KERNELS = 2 };
#define BLK_ACTION_NONE (-1)   /* per-env "do not move": the env only gets its mask / status refreshed */
enum { BLK_FLAG_DONE = 1, BLK_FLAG_ILLEGAL = 2, BLK_FLAG_TRUNCATED = 4 };

/* Arguments of blk_step().  Every pointer is a DEVICE pointer; nullable ones are marked. */
typedef struct {
    int64_t n;                  /* number of envs */
    const uint32_t *state_in;   /* [n][state_words] */
    uint32_t *state_out;        /* [n][state_words]; may alias state_in (in-place) */
    const int32_t *action;      /* [n] action ids (BLK_ACTION_NONE = skip this env); NULL = no env moves: mask only */
    void *mask;                 /* nullable; next mover's legal mask */
    int32_t mask_format;        /* blk_mask_format */
    int64_t mask_stride;        /* row stride: uint32 words for BITS (>= mask_words); bytes for BYTES (>= num_actions;
                                   fastest when a multiple of 16 with a 16 B aligned base); uint16 entries for INDICES
                                   (rows with more legal actions than that are truncated and flagged) */
    int32_t *legal_count;       /* nullable [n]; number of legal actions of the next mover */
    float *terminal;            /* nullable [n][P]; 3 / 1 / -1 terminal vector when the step ended the game, else 0 */
    uint8_t *flags;             /* nullable [n]; BLK_FLAG_* */
    int16_t *scores;            /* nullable [n][P]; scores after the step (final rule applied; before auto-reset) */
    int32_t *next_action;       /* nullable [n]; uniform random legal action of the next mover, -1 if none */
    uint64_t seed;              /* Philox-4x32-10 key = (seed_lo, seed_hi ^ (env_id_base + i)), ctr = (ply >> 2, game, 0, 0), word ply & 3 */
    uint32_t env_id_base;       /* global id of env 0 of this batch (multi-GPU sharding keeps results partition-invariant) */
    uint32_t options;           /* BLK_OPT_* */
    float *obs;                 /* nullable [n][2P][N][N] float32, 16 B aligned: canonical_board of the RESULTING state
                                   (blokus_wrapper.py:144-146), written by the same kernel -- leaf expansion for the
                                   policy/value net is then one launch: new state + legal mask + observation */
    const int32_t *state_index; /* nullable [n]: env i reads state_in[state_index[i]] instead of state_in[i] (a search tree
                                   stepping states out of its node pool: no gather pass; state_out must not alias state_in) */
    /* BLK_MASK_INDICES, compact form (both NULL = padded rows).  With csr_cursor set, `mask` is ONE flat uint16 array of
     * mask_stride entries shared by all envs: env i's ids occupy [csr_offset[i], csr_offset[i] + legal_count[i]) -- as
     * many bytes as there are legal moves (~350 B per env at 20x20 instead of a 2 KB row), which is what a host-side policy
     * wants to copy back.  The caller zeroes *csr_cursor before the call; afterwards it holds the entries requested (envs land
     * in launch order, not env order).  An env whose ids do not fit any more is flagged BLK_FLAG_TRUNCATED and writes none. */
    unsigned long long *csr_cursor;   /* nullable device counter */
    int64_t *csr_offset;              /* [n], required with csr_cursor */
} blk_step_args;

/* Arguments of blk_rollout(): uniform-random playouts to the end of the game, one warp per game. */
typedef struct {
    int64_t n_roots;
    const uint32_t *roots;      /* [n_roots][state_words] */
    int32_t per_root;           /* playouts per root */
    uint64_t seed;              /* key = (seed_lo, seed_hi ^ game_index), game_index = rollout_id_base + root*per_root + j; ctr = (ply >> 2, game, 1, 0), word ply & 3 */
    uint32_t rollout_id_base;
    int16_t *final_scores;      /* nullable [n_roots*per_root][P] */
    uint8_t *winners;           /* nullable [n_roots*per_root] bitmask of winners */
    float *value_sum;           /* nullable [n_roots][P]; sum over playouts of the 3/1/-1 terminal vector (atomicAdd; caller zeroes) */
    uint16_t *action_log;       /* nullable [n_roots*per_root][log_stride]; 0xFFFF terminated */
    int32_t log_stride;         /* >= 4*21+1 when action_log != NULL */
    int32_t *plies;             /* nullable [n_roots*per_root] plies played */
    int32_t stop_player;        /* -1: play to the end of the game; q: stop as soon as it is player q's turn (random-bot
                                   opponents inside a gym step: docs/README.md:47-51 of the reference) */
    uint32_t *state_out;        /* nullable [n_roots*per_root][state_words]: the state each playout stopped in */
    uint32_t options;           /* BLK_OPT_WARP_KERNELS or 0 */
} blk_rollout_args;

const char *blk_last_error(void);
int blk_abi_version(void);

/* Engine lifetime.  Replaces BlokusEnvironment() at blokus_wrapper.py:42 and the action-table
 * construction at blokus_wrapper.py:281-324 (tables are built here, deterministically). */
int blk_create(const blk_config *cfg, blk_engine **out);
void blk_destroy(blk_engine *h);
int blk_get_info(const blk_engine *h, blk_info *out);

/* Action table (host side, no GPU work): piece, orientation-within-piece, anchor y, anchor x and the
 * covered cells (y0,x0,y1,x1,...; `ncells` pairs) of one action id.  Footprint-level view of the ids
 * built at blokus_wrapper.py:300-316. */
int blk_action_to_cells(const blk_engine *h, int32_t action, int32_t meta[4], uint8_t cells_yx[10], int32_t *ncells);

/* new_state: blokus_wrapper.py:80-87 (env.new_state()).  Writes n fresh states (game counter 0). */
int blk_reset(blk_engine *h, uint32_t *state, int64_t n, void *stream);

/* next_state + valid_actions + get_winners fused: blokus_wrapper.py:89-106, 108-132, 164-186.
 * With args->action == NULL it is valid_actions only (blokus_wrapper.py:122-124). */
int blk_step(blk_engine *h, const blk_step_args *args, void *stream);

/* canonical_board: blokus_wrapper.py:144-146 -> float32 [n][2P][N][N] (models/blokus_nnet.py:99). */
int blk_observe(blk_engine *h, const uint32_t *state, float *obs, int64_t n, void *stream);

/* board_contents: blokus_wrapper.py:208-218, 259-266 -> uint8 [n][N][N], 0 empty, 1..P colour. */
int blk_board_contents(blk_engine *h, const uint32_t *state, uint8_t *board, int64_t n, void *stream);

/* Terminal status of existing states (get_winners without stepping): flags/terminal/scores as in blk_step. */
int blk_game_ended(blk_engine *h, const uint32_t *state, uint8_t *flags, float *terminal, int16_t *scores,
                   int64_t n, void *stream);

/* Batched uniform-random playouts (north_star kernel family 4; new capability behind the Player interface,
 * blokus_rl/players/random_player.py:11-17 repeated to the end of the game). */
int blk_rollout(blk_engine *h, const blk_rollout_args *args, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Device-resident PUCT forest (blokus_rl/alphazero/mcts.py:13-71 for B searches in lockstep).  All arrays are
 * caller-owned DEVICE memory; one simulation = blk_puct_select -> blk_step on the requested transitions ->
 * evaluator -> blk_puct_expand -> blk_puct_backup, without any host synchronisation.
 * --------------------------------------------------------------------------------------------------------- */
enum { BLK_PUCT_TERMINAL = 0, BLK_PUCT_NEED_EVAL = 1, BLK_PUCT_NEED_STEP = 2 };

typedef struct {
    int32_t num_trees, num_players, num_actions;
    int32_t mask_stride;          /* bytes per byte-mask row (engine mask_bytes) */
    int32_t node_capacity, edge_capacity, max_depth;
    int32_t *node_edge0;          /* [nodes] first edge, -1 = not expanded */
    int32_t *node_nedge;          /* [nodes] */
    int32_t *node_state;          /* [nodes] slot of the node's state in the caller's state pool */
    int8_t *node_mover;           /* [nodes] player to move */
    int8_t *node_terminal;        /* [nodes] 1 = the game is over here */
    double *node_term_value;      /* [nodes][P] 3/1/-1 vector of terminal nodes */
    int32_t *edge_action;         /* [edges] action id (ascending within a node) */
    int32_t *edge_child;          /* [edges] child node, -1 = not opened */
    double *edge_n, *edge_q, *edge_p;   /* [edges] visit count, running mean, prior (mcts.py:67-70) */
    int32_t *root;                /* [B] */
    int32_t *path;                /* [B][max_depth] edges of the current simulation */
    int32_t *path_len;            /* [B] */
    int32_t *status;              /* [B] BLK_PUCT_* of the current simulation */
    int32_t *leaf_node;           /* [B] node reached, or parent of the edge to open */
    int32_t *leaf_edge;           /* [B] edge to open or -1 */
    int32_t *src_slot;            /* [B] state slot blk_step must read for this tree */
    int32_t *step_action;         /* [B] action for blk_step (BLK_ACTION_NONE: evaluate the state as it is) */
    double *scores;               /* [B][P] score vector of the current simulation */
    int32_t *counters;            /* [6] nodes used, edges used, capacity overflow flag, illegal-action flag, pool slots used,
                                     pending pool-slot advance (fused backup) */
    double *node_sum_n;           /* [nodes] sum of the node's edge visit counts (the N.sum() of mcts.py:43) */
    int32_t *path_node;           /* [B][max_depth] node of each path edge */
    int8_t *node_uniform;         /* [nodes] 1 = expanded with the uniform prior: P = 1/nedge for every edge, edge_p is not
                                     written (saves a quarter of the expansion's stores and a third of the selection's loads) */
    /* Board-keyed node table (blk_puct_search / blk_puct_reroot; NULL for the lockstep kernels, which key nodes by path):
     * the reference files a node under hash(board cells) only (mcts.py:37, blokus_wrapper.py:217-218), so two move orders
     * that reach the same board share one node.  Open addressing; an entry is node + 1, 0 = empty. */
    int32_t *hash_table;          /* [hash_capacity], zeroed by the caller */
    int32_t hash_capacity;        /* power of two, >= 2 * node_capacity */
    uint64_t *node_hash;          /* [nodes] 64-bit hash of (tree, board rows) */
    int32_t *node_tree;           /* [nodes] tree the node belongs to (equal boards of different trees stay apart) */
    int32_t *edge_vl;             /* [edges] virtual-loss counters, zeroed by the caller (leaf-parallel search only) */
    int32_t *node_front;          /* [nodes] blk_puct_search: edges [0, front) of the node have been selected at least once (with
                                     the uniform prior the selected edges always form a prefix; see csrc/blk_search.cuh) */
} blk_puct_forest;

typedef struct {
    int32_t new_slot_base;        /* pool slot of tree 0's new state (tree t uses base + t); -1: take counters[4], which
                                     blk_puct_backup advances by B (keeps every argument constant: CUDA-graph capture) */
    int32_t state_words, meta_word;   /* engine state_words and the index of the meta word (P*N + P) */
    int32_t attach_only;          /* 1: only create the requested child nodes and make them the roots (blk_puct_advance) */
    const uint32_t *new_states;   /* [B][state_words] */
    uint32_t *pool;               /* nullable: state pool; the new states are copied to slots base + t */
    const uint8_t *mask;          /* [B][mask_stride] byte masks, or bit-packed rows when mask_bits = 1 */
    int32_t mask_bits;            /* 1: mask holds uint32 words, mask_stride_words per row (uniform prior: no net needs bytes) */
    int32_t mask_stride_words;
    const uint8_t *flags;         /* [B] */
    const float *terminal;        /* [B][P] */
    const void *prior;            /* [B][prior_stride] or NULL */
    int32_t prior_dtype;          /* 0 uniform over the legal actions (DumbNet), 1 float32, 2 float64 */
    int64_t prior_stride;
    const double *value;          /* [B][P] or NULL (zeros) */
    int32_t fuse_backup;          /* 1: the expansion also does the backup of blk_puct_backup (one launch less per simulation);
                                     counters[] then needs 6 entries ([5] = pool slots of the last simulation still to be counted) */
} blk_puct_expand_args;

/* Whole simulations inside ONE kernel (no launch per tree level, no host in the loop): for the evaluators that need no
 * network -- the uniform prior with zero value of the reference's "mcts" arena player (models/dumbnet.py:14-21,
 * compare_arena.py:87-95), optionally with the mean of `playouts_per_leaf` uniform-random playouts as the value. */
typedef struct {
    int32_t num_sims;             /* simulations per tree in this launch */
    double cpuct;                 /* root-level exploration constant (deeper levels use 1: mcts.py:50-52) */
    int32_t epsilon_fix;
    uint32_t *pool;               /* [node_capacity][state_words] state pool (node_state indexes it) */
    int32_t warps_per_tree;       /* 1: simulations of a tree run one after the other, exactly as the reference's (visit
                                     counts, Q and per-simulation scores are reproduced).  > 1 (<= 16): that many warps
                                     search the same tree at once with virtual loss -- leaf-parallel, NOT the reference's
                                     visit order; needs edge_vl.  In this mode edge_q holds the SUM of the values backed up
                                     through the edge (Q = edge_q / edge_n), so that a backup is atomic adds */
    int32_t playouts_per_leaf;    /* 0: leaf value = 0 (DumbNet); k > 0: mean 3/1/-1 vector of k uniform-random playouts */
    uint64_t seed;                /* playout RNG key (stream 2) */
    double virtual_loss;          /* value a pending visit is counted as for its edge (leaf-parallel); 0 -> 1.0 */
} blk_puct_search_args;

const char *blk_puct_last_error(void);
/* MCTS.simulate, selection half: mcts.py:39-52 for every tree, down to a leaf / unopened edge / terminal node. */
int blk_puct_select(const blk_puct_forest *f, double cpuct, int32_t epsilon_fix, void *stream);
/* MCTS.simulate, expansion half: mcts.py:59-71 (winners / valid moves / nn.predict results become nodes + edges). */
int blk_puct_expand(const blk_puct_forest *f, const blk_puct_expand_args *args, void *stream);
/* MCTS.simulate, backup half: mcts.py:53-57 along the recorded path. */
int blk_puct_backup(const blk_puct_forest *f, void *stream);
/* MCTSPlayer's choice, players/mcts_player.py:19-20: most visited root action per tree (first maximum); device arrays [B]. */
int blk_puct_best(const blk_puct_forest *f, int32_t *best_action, double *best_visits, void *stream);
/* After a real move: the child under `actions[t]` becomes the root (tree reuse, players/mcts_player.py:15-22). */
int blk_puct_advance(const blk_puct_forest *f, const int32_t *actions, void *stream);

/* MCTS.simulate x num_sims for every tree, selection, env transition, legal-move expansion and backup fused in one
 * kernel (one warp, or `warps_per_tree` warps, per tree); nodes are keyed by board cells like the reference's dict
 * (mcts.py:37), a simulation that reaches a known board continues its descent there.  Needs the engine (env tables).
 * A forest searched this way is built by blk_puct_reroot / blk_puct_search only (node i keeps its state in pool slot i;
 * counters[0] = nodes, counters[1] = edges; zero the counters and hash_table to start over); do not mix it with the
 * lockstep kernels above. */
int blk_puct_search(blk_engine *h, const blk_puct_forest *f, const blk_puct_search_args *args, void *stream);
/* Tree reuse the way the reference gets it (its dict outlives the move: players/mcts_player.py:15-25): the root of tree t
 * becomes the node filed under the board of states[t] if the tree has one, else a fresh unexpanded node. */
int blk_puct_reroot(blk_engine *h, const blk_puct_forest *f, const blk_puct_search_args *args, const uint32_t *states, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* BLOKUS_B200_H */
