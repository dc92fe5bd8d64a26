/* include/scopa_b200.h -- C ABI of libscopa_b200.so (sm_100a CUDA implementation of the Miniscopa
 * hot path: env transitions + CFR / MCCFR / SDCFR traversals).
 *
 * The reference (rug-marl-group2/scopa) has no FFI or plugin interface: its boundary is a Python
 * class API.  Each entry point below names the reference code it replaces (paths relative to
 * /root/reference/); the Python drop-in classes in scopa_b200/ bind these symbols with ctypes and
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain C types only; `d_` pointers are DEVICE pointers, `h_` pointers are HOST pointers
 *     (pinned memory makes the copies asynchronous but is not required);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); device-pointer
 *     entry points only enqueue work and return; host-pointer (`*_host`) entry points copy in,
 *     launch, copy out and synchronise before returning;
 *   - every function returns 0 on success and a negative ms_status on failure; the message is in
 *     ms_last_error() (thread-local).  No CPU fallback exists: without a CUDA device every compute
 *     entry point fails with MS_ERR_CUDA;
 *   - caller owns all I/O buffers; handles are freed with the matching *_destroy.
 */
#ifndef SCOPA_B200_H
#define SCOPA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MS_ABI_VERSION 1

typedef enum {
    MS_OK = 0,
    MS_ERR_CUDA = -1,        /* CUDA runtime error (incl. no device) */
    MS_ERR_ARG = -2,         /* bad argument */
    MS_ERR_CAPACITY = -3,    /* tree / table capacity exceeded */
    MS_ERR_STATE = -4        /* handle used in the wrong state */
} ms_status;

/* 16-byte packed game state; bit layout in scopa_b200/csrc/ms_state.cuh (also DESIGN.md). */
typedef struct { uint32_t hands, table, captures, meta; } ms_state;

int ms_abi_version(void);
const char* ms_last_error(void);
/* number of CUDA kernels this library has launched so far in this process (all streams) */
uint64_t ms_launch_count(void);

/* ---------------------------------------------------------------------------------- env ------
 * ms_deal_from_seeds: MiniScopaEnv.reset(seed) for n games = MiniDeck(seed) shuffle with CPython's
 *   MT19937 random.seed/random.shuffle, first 8 cards dealt 4+4, empty table
 *   (src/envs/mini_scopa_game.py:15-33, :56-64, :131-138; seed 0 means 42, :132).
 *   d_hand_order[g]: nibble i = i-th dealt card of player 0 (i<4) / player 1 (i>=4): the order in
 *   which legal_actions() and the infoset string list the hand. */
int ms_deal_from_seeds(const int64_t* d_seeds, int64_t n, ms_state* d_states, uint32_t* d_hand_order,
                       void* stream);

/* ms_deck_from_seeds: the whole shuffled deck MiniDeck(seed).cards (src/envs/mini_scopa_game.py:25-28):
 *   d_deck[g] nibble i = card id at position i (the deal takes positions 0-7).  Unlike
 *   ms_deal_from_seeds no seed substitution happens here (MiniDeck(0) really seeds with 0). */
int ms_deck_from_seeds(const int64_t* d_seeds, int64_t n, uint64_t* d_deck, void* stream);

/* test hook: the same result through the kernel's rarely-taken slow path (full MT19937 state) */
int ms_debug_deal_slow_path(const int64_t* d_seeds, int64_t n, ms_state* d_states, uint32_t* d_hand_order,
                            void* stream);

/* measurement hook: atomic-add throughput of this GPU on a table the size of the MCCFR delta table, pseudo-random
 * addresses, all SMs busy: h_out[0] = shared-memory fp64 atomicAdd/s, [1] = shared-memory u32 atomicAdd/s,
 * [2] = global (L2-resident) fp64 RED/s.  The atomic roofline SURVEY 8(d) asks for. */
int ms_debug_atomic_peaks(double h_out[3], void* stream);

/* tuning hook: games per pipeline stage of the *_host rollout entry points (H2D / kernels / D2H of consecutive
 * stages overlap).  games <= 0 restores the default.  Returns the value now in effect. */
int64_t ms_debug_set_host_chunk(int64_t games);
/* size of the pipeline stage that starts at game `lo` of an n-game *_host rollout call: full stages of the chunk size between
 * a quarter-size first and last stage (so the GPU neither idles through a full copy in nor ends on a full copy out); every
 * stage but the last is a multiple of 128 games.  0 when lo is outside [0, n).  No device needed. */
int64_t ms_debug_host_stage_size(int64_t lo, int64_t n);

/* ms_step: MiniScopaEnv.step(action) on n independent states in place (src/envs/mini_scopa_game.py:140-167
 *   incl. play_card :93-104, card_in_table :66-91, evaluate_game :106-114).  d_rewards ([n][2] f32,
 *   may be NULL) receives the terminal rewards (0,0 while running); d_done ([n] u8, may be NULL)
 *   the terminal flag.  Illegal action = pass; step on a terminal state = no-op. */
int ms_step(ms_state* d_states, const uint8_t* d_actions, float* d_rewards, uint8_t* d_done, int64_t n,
            void* stream);

/* ms_legal_actions: MiniScopaState.legal_actions(player) (src/envs/openspiel_mini_scopa.py:22-47);
 *   player = -1 means each state's current player.  d_mask ([n] u16, bit = action id),
 *   d_ordered ([n][4] u8, hand order, 0xFF padded) and d_count ([n] u8) may each be NULL.
 *   d_capture ([n][4] u8, may be NULL): for the k-th legal action, the table-position mask that
 *   playing it would capture (0 = the card would be placed) -- the legal capture moves. */
int ms_legal_actions(const ms_state* d_states, const uint32_t* d_hand_order, int player, uint16_t* d_mask,
                     uint8_t* d_ordered, uint8_t* d_count, uint8_t* d_capture, int64_t n, void* stream);

/* ms_capture: MiniScopaGame.card_in_table(card) for card id d_cards[g] against state g's table
 *   (src/envs/mini_scopa_game.py:66-91) -> table-position mask (0 = no capture). */
int ms_capture(const ms_state* d_states, const uint8_t* d_cards, uint8_t* d_table_pos_mask, int64_t n,
               void* stream);

/* ms_infoset_keys: packed form of information_state_string(player) (openspiel_mini_scopa.py:86-95);
 *   player = -1 means current player; terminal states give key 0xFFFFFFFFFFFFFFFF ("TERMINAL"). */
int ms_infoset_keys(const ms_state* d_states, int player, uint64_t* d_keys, int64_t n, void* stream);

/* ms_rollout_random: config 2 of BASELINE.json -- n concurrent random-policy games played to the
 *   end in one launch (reset already done: d_states/d_hand_order from ms_deal_from_seeds).  At each
 *   ply the mover plays legal[mulhi32(x0, n_legal)] with x = Philox4x32-10(key = philox_seed,
 *   ctr = (game id lo, hi, ply, "ROLL")), game id = game_offset + index.  Outputs (each may be
 *   NULL): d_actions [n][8] u8, d_rewards [n][2] f32, d_final [n] final states. */
int ms_rollout_random(const ms_state* d_states, const uint32_t* d_hand_order, int64_t n, uint64_t philox_seed,
                      uint64_t game_offset, uint8_t* d_actions, float* d_rewards, ms_state* d_final,
                      void* stream);

/* Host-buffer forms (end-to-end: H2D, kernels, D2H, synchronise). */
int ms_deal_from_seeds_host(const int64_t* h_seeds, int64_t n, ms_state* h_states, uint32_t* h_hand_order);
int ms_step_host(ms_state* h_states, const uint8_t* h_actions, float* h_rewards, uint8_t* h_done, int64_t n);
int ms_legal_actions_host(const ms_state* h_states, const uint32_t* h_hand_order, int player, uint16_t* h_mask,
                          uint8_t* h_ordered, uint8_t* h_count, uint8_t* h_capture, int64_t n);
int ms_infoset_keys_host(const ms_state* h_states, int player, uint64_t* h_keys, int64_t n);
/* seeds in -> actions and rewards out: reset + 8 steps per game, all on the device */
int ms_rollout_random_host(const int64_t* h_seeds, int64_t n, uint64_t philox_seed, uint64_t game_offset,
                           uint8_t* h_actions, float* h_rewards);

/* ------------------------------------------------------------------------------ solvers ------
 * ms_solver: one deal (root state) + its game tree + the slot-aligned infoset table, all resident
 * in device memory.  Creation enumerates the tree on the device (level-synchronous expansion with
 * the same step() as above) and assigns each infoset a dense slot in breadth-first first-occurrence
 * order, which is identical on every GPU -> the regret/strategy arrays of different ranks are
 * slot-aligned and can be all-reduced as they are.
 *   replaces: the Python dicts CFRTrainer.info_set_map (src/algorithms/vanilla_cfr.py:49-54) and
 *   MCCFRTrainer.info_sets (src/algorithms/mc_cfr.py:28-35). */
typedef struct ms_solver ms_solver;

int ms_solver_create(const ms_state* h_root, uint32_t hand_order, ms_solver** out);
void ms_solver_destroy(ms_solver* s);
int ms_solver_reset(ms_solver* s, void* stream);                 /* zero regrets / strategy sums; MS_ERR_STATE while attached to peers */
int ms_solver_counts(const ms_solver* s, int32_t* n_nodes, int32_t* n_slots, int32_t* n_levels);
/* tree export (host buffers sized n_nodes): packed state, parent index (-1 root), level, slot
 * (-1 for terminals), first-child index, child count */
int ms_solver_export_tree(const ms_solver* s, ms_state* h_states, int32_t* h_parent, uint8_t* h_level,
                          int32_t* h_slot, int32_t* h_child_begin, uint8_t* h_nchild);
/* table export (host buffers sized n_slots): key, n_legal, legal ids in hand order [.][4],
 * regret [.][4] f64, strategy [.][4] f64, touched flag (MCCFR creates nodes on first touch) */
int ms_solver_export_table(const ms_solver* s, uint64_t* h_keys, uint8_t* h_nlegal, uint8_t* h_legal,
                           double* h_regret, double* h_strategy, uint8_t* h_touched, void* stream);
int ms_solver_import_table(ms_solver* s, const double* h_regret, const double* h_strategy, void* stream);
/* device pointers to the slot-aligned arrays for collectives: regret / strategy are [n_slots][4] f64
 * (*n_table doubles each); delta is ONE contiguous buffer of *n_delta = 6 * n_slots doubles:
 * [n_slots][4] regret deltas, [n_slots] update counts (the strategy delta of a batch is count * sigma
 * because sigma is frozen for the batch) and [n_slots] first-touch marks (non-zero = some traversal of the
 * batch created the InfoNode, mc_cfr.py:52; summed like the rest, so every rank's `touched` flags agree)
 * -> one all-reduce per iteration. */
int ms_solver_device_ptrs(ms_solver* s, double** d_regret, double** d_strategy, double** d_delta, size_t* n_table,
                          size_t* n_delta);

/* ms_cfr_iterate: `iters` x (traverser 0, traverser 1) of CFRTrainer._cfr_recursive from the root
 *   (src/algorithms/vanilla_cfr.py:56-99, :105-110), float64, order-exact (per-visit sigma refresh).
 * ms_cfr_traverse: one call of _cfr_recursive(root, player, reach_p0, reach_p1); *h_value receives
 *   the returned node utility (synchronises). */
int ms_cfr_iterate(ms_solver* s, int32_t iters, void* stream);
int ms_cfr_traverse(ms_solver* s, int32_t player, double reach_p0, double reach_p1, double* h_value, void* stream);
/* ms_cfr_iterate_many: `iters` CFR iterations on n independent solvers (deals) in ONE launch, one CTA per deal
 *   (throughput mode: the per-deal sweep is latency bound, 148 of them run side by side). */
int ms_cfr_iterate_many(ms_solver* const* solvers, int32_t n_solvers, int32_t iters, void* stream);

/* ms_mccfr_inplace: `iters` reference iterations of MCCFRTrainer.iteration()
 *   (src/algorithms/mc_cfr.py:37-92) with in-place table updates after every node, sampling from the
 *   Philox "MCCF" stream with traversal id first_iter + i (DESIGN.md "Random streams").
 * ms_mccfr_inplace_many: the reference's experiment protocol (src/experiments/run_mccfr_experiment.py:195-202:
 *   independent runs of MCCFRTrainer from an empty table) in ONE launch: n_runs tables, one warp per run, run r
 *   advancing `iters` iterations on the stream of philox seed philox_seed0 + r -- bit-identical to a solo
 *   ms_mccfr_inplace(s, iters, philox_seed0 + r, first_iter) on a table holding run r's rows.  The caller owns the
 *   device buffers d_regret / d_strategy [n_runs][n_slots][4] f64 and d_touched [n_runs][n_slots] u8 (zero them for a
 *   fresh start; call again to continue); the solver's own table is not touched.
 * ms_mccfr_batch: n_trav independent traversals for `player` (0, 1, or 2 = both) against the
 *   current (frozen) table, traversal ids first_trav..; regret / strategy deltas are accumulated
 *   into the delta arrays (not applied).  On the tree of a fresh deal (every MiniScopaEnv.reset) this is the
 *   static-shape kernel on the sequential Philox stream (DESIGN.md sections 5, 7); other roots use mode 4.
 * ms_mccfr_apply: table += delta; delta = 0 (call after the all-reduce of the delta arrays). */
int ms_mccfr_inplace(ms_solver* s, int64_t iters, uint64_t philox_seed, uint64_t first_iter, void* stream);
int ms_mccfr_inplace_many(ms_solver* s, int32_t n_runs, int64_t iters, uint64_t philox_seed0, uint64_t first_iter,
                          double* d_regret, double* d_strategy, uint8_t* d_touched, void* stream);
int ms_mccfr_batch(ms_solver* s, int32_t player, int64_t n_trav, uint64_t philox_seed, uint64_t first_trav,
                   void* stream);
int ms_mccfr_apply(ms_solver* s, void* stream);
/* ms_mccfr_batch_mode: like ms_mccfr_batch with a choice of estimator: mode 0 = the reference's estimator
 *   (mc_cfr.py:37-86), 1 = external sampling, 2 = outcome sampling (epsilon 0.6) -- the textbook estimators the
 *   reference does not implement (update rules as published by Lanctot et al. 2009); 3 = the reference's estimator
 *   re-stepping the env at every node; 4 = the reference's estimator on the generic tree-walking kernel (any root;
 *   call-indexed Philox stream).  Same table, same delta buffer, same apply step. */
int ms_mccfr_batch_mode(ms_solver* s, int32_t mode, int32_t player, int64_t n_trav, uint64_t philox_seed,
                        uint64_t first_trav, void* stream);
/* Peer-memory exchange (one process per GPU, NVLink / NVSwitch): instead of a library all-reduce + ms_mccfr_apply,
 * every rank maps the other ranks' inboxes (CUDA IPC) and ONE kernel per rank pushes its deltas into every inbox,
 * meets the others at a flag barrier, sums its own inbox (in rank order: replicas stay bit-identical) and updates its table.
 *   ms_solver_ipc_export: 64-byte cudaIpcMemHandle_t of this solver's device block + byte offsets of its two inboxes
 *     (one per iteration parity) and of its flag array; exchange them between ranks (any out-of-band all-gather of 88
 *     bytes per rank);
 *   ms_solver_ipc_attach: `handles` = world x 64 bytes, `offsets` = world x 3 u64, in rank order (at most 8 ranks);
 *   ms_mccfr_apply_peers: replaces {all-reduce, ms_mccfr_apply} after ms_mccfr_batch; inboxes are double buffered by
 *     iteration parity, so one cross-GPU barrier per iteration suffices.  Every rank must call it once per iteration.
 *     The barrier is bounded: a rank that has waited 2 s for a peer sets the solver's error word and leaves its table
 *     unchanged from then on (no hang);
 *   ms_solver_peer_error: synchronises the stream and reads that word: *h_err = 0 and MS_OK, or 1 + the rank that
 *     did not arrive and MS_ERR_STATE;
 *   ms_mccfr_batch_peers: ms_mccfr_batch + ms_mccfr_apply_peers as ONE launch per iteration: the last CTA to finish its
 *     traversals performs the exchange and the table update (fused compute + exchange; fresh-deal roots, otherwise it
 *     falls back to the two launches). */
int ms_solver_ipc_export(ms_solver* s, void* handle64, uint64_t offsets[3]);
int ms_solver_ipc_attach(ms_solver* s, int32_t rank, int32_t world, const void* handles, const uint64_t* offsets);
int ms_mccfr_apply_peers(ms_solver* s, void* stream);
int ms_mccfr_batch_peers(ms_solver* s, int32_t player, int64_t n_trav, uint64_t philox_seed, uint64_t first_trav,
                         void* stream);
int ms_solver_peer_error(ms_solver* s, uint32_t* h_err, void* stream);
/* counters accumulated by the MCCFR kernels since the last reset: [0] traverser-node updates,
 * [1] node visits (_sample calls), [2] env steps */
int ms_solver_counters(ms_solver* s, uint64_t h_out[3], int reset, void* stream);

/* ms_best_response: restated open_spiel best response against the table's average policy
 *   (vanilla_cfr.py:112-118 calls exploitability.exploitability).  policy_kind 0 = LearnedCFRPolicy
 *   (strategy_sum normalised if > 0, vanilla_cfr.py:32-39), 1 = ScopaLearnedPolicy (> 1e-12 and
 *   touched, mc_cfr.py:118-130), 2 = uniform.  h_br_values[2]; exploitability = sum / 2. */
int ms_best_response(ms_solver* s, int32_t policy_kind, double h_br_values[2], void* stream);

/* ms_solver_policy: the table's average policy per slot -> d_policy [n_slots][4] f64 (probabilities over the
 *   legal actions in hand order; policy_kind as in ms_best_response).
 * ms_eval_policies: batched evaluate_agent (vanilla_cfr.py:157-216, mc_cfr.py:146-206): n_games episodes from
 *   the root, seat 0 acting with d_policy_seat0 and seat 1 with d_policy_seat1 (same [n_slots][4] layout),
 *   sampled from the Philox "EVAL" stream with episode ids first_game.. .  Outputs (may be NULL):
 *   d_reward0 [n] f32 = player 0's terminal reward, d_scopas [n][2] u8. */
int ms_solver_policy(ms_solver* s, int32_t policy_kind, double* d_policy, void* stream);
int ms_eval_policies(ms_solver* s, const double* d_policy_seat0, const double* d_policy_seat1, int64_t n_games,
                     uint64_t philox_seed, uint64_t first_game, float* d_reward0, uint8_t* d_scopas, void* stream);

/* ------------------------------------------------------------------ team Miniscopa (2v2) ------
 * The 4-player variant (src/envs/team_mini_scopa_game.py:44-243): all 16 cards dealt 4x4, 16 plies, teams {0,1}
 * vs {2,3}, last-capturer sweep, team scoring.  32-byte packed state (layout in scopa_b200/csrc/ms_team.cu).
 *   ms_team_deal_from_seeds: TeamMiniScopaEnv.reset(seed) (:167-169 -> TeamMiniScopaGame.reset :62-70; seed 0
 *     means 42); d_hand_order[g] = the shuffled deck, nibble 4p+i = i-th card dealt to player p.
 *   ms_team_step: TeamMiniScopaEnv.step (:171-205); d_rewards [n][4] f32 (team rewards t0,t0,t1,t1), d_done [n] u8.
 *   ms_team_rollout_random: 16 uniform-random legal plies per game in one launch (Philox "TEAM" stream: ctr =
 *     (game id lo, hi, ply/4, tag), word ply%4); d_actions [n][16] u8. */
typedef struct { uint32_t w[8]; } ms_team_state;
int ms_team_deal_from_seeds(const int64_t* d_seeds, int64_t n, ms_team_state* d_states, uint64_t* d_hand_order, void* stream);
int ms_team_step(ms_team_state* d_states, const uint8_t* d_actions, float* d_rewards, uint8_t* d_done, int64_t n, void* stream);
int ms_team_rollout_random(const ms_team_state* d_states, const uint64_t* d_hand_order, int64_t n, uint64_t philox_seed,
                           uint64_t game_offset, uint8_t* d_actions, float* d_rewards, ms_team_state* d_final, void* stream);
int ms_team_deal_from_seeds_host(const int64_t* h_seeds, int64_t n, ms_team_state* h_states, uint64_t* h_hand_order);
int ms_team_step_host(ms_team_state* h_states, const uint8_t* h_actions, float* h_rewards, uint8_t* h_done, int64_t n);

/* ------------------------------------------------------------------------ 40-card Scopa ------
 * FullScopaEnv, two players (src/envs/full_scopa_game.py:21-342).  Card id = action id =
 * suit_idx * 10 + rank - 1 (suits denari, coppe, spade, bastoni; :262-266).  32-byte packed state (bit layout in
 * scopa_b200/csrc/ms_full.cu) + the shuffled deck beside it (four 64-bit words, ten 6-bit card ids each: the
 * table is deck[0..3], round r deals deck[4 + 6 r + 3 p + i] to player p).
 *   ms_full_deal_from_seeds: FullScopaEnv.reset(seed) (:243-250, :68-86; seed 0 means 42) = FullDeck(seed) shuffle with
 *     CPython's MT19937 (:32-35), four cards to the table, three to each player.
 *   ms_full_deck_from_seeds: FullDeck(seed).cards only (no seed substitution); slow_path != 0 forces the kernel's
 *     rarely-taken full-state path (test hook).
 *   ms_full_step: FullScopaEnv.step(action) (:252-296 incl. play_card :130-158, find_capture_combinations :101-128,
 *     deal_new_round :92-99, evaluate_game :174-226, the 200-step safety limit).  A card the mover does not hold is a
 *     silent pass; ids above 39 (IndexError in the reference) are passes too.  d_rewards [n][2] f32, d_done [n] u8.
 *   ms_full_legal_actions: FullScopaState.legal_actions(player) (src/envs/openspiel_full_scopa.py:22-41), player -1 =
 *     current: d_ordered [n][3] u8 (hand order, 0xFF padded), d_count [n] u8.
 *   ms_full_rollout_random: n random-policy games played to the end (36 plies) in one launch; at each ply the mover
 *     plays legal[mulhi32(x, n_legal)], x = Philox4x32-10(key = philox_seed, ctr = (game id lo, hi, ply / 4, "FULL")),
 *     word ply % 4.  d_actions [n][36] u8, d_rewards [n][2] f32, d_final [n] (each may be NULL).
 *   ms_full_table_overflow: 1 if any game so far held more than 16 cards on the table (the packed state's limit;
 *     200 k random games peak at 11) -- such a game's state is invalid. */
typedef struct { uint32_t w[8]; } ms_full_state;
typedef struct { uint64_t w[4]; } ms_full_deck;
int ms_full_deal_from_seeds(const int64_t* d_seeds, int64_t n, ms_full_state* d_states, ms_full_deck* d_decks, void* stream);
int ms_full_deck_from_seeds(const int64_t* d_seeds, int64_t n, ms_full_deck* d_decks, int slow_path, void* stream);
int ms_full_step(ms_full_state* d_states, const ms_full_deck* d_decks, const uint8_t* d_actions, float* d_rewards,
                 uint8_t* d_done, int64_t n, void* stream);
int ms_full_legal_actions(const ms_full_state* d_states, const ms_full_deck* d_decks, int player, uint8_t* d_ordered,
                          uint8_t* d_count, int64_t n, void* stream);
int ms_full_rollout_random(const ms_full_state* d_states, const ms_full_deck* d_decks, int64_t n, uint64_t philox_seed,
                           uint64_t game_offset, uint8_t* d_actions, float* d_rewards, ms_full_state* d_final, void* stream);
int ms_full_table_overflow(int* h_flag, void* stream);
int ms_full_deal_from_seeds_host(const int64_t* h_seeds, int64_t n, ms_full_state* h_states, ms_full_deck* h_decks);
int ms_full_step_host(ms_full_state* h_states, const ms_full_deck* h_decks, const uint8_t* h_actions, float* h_rewards,
                      uint8_t* h_done, int64_t n);
int ms_full_rollout_random_host(const int64_t* h_seeds, int64_t n, uint64_t philox_seed, uint64_t game_offset,
                                uint8_t* h_actions, float* h_rewards);
/* FullScopaGame.evaluate_game() on its own (:174-226): last-capturer sweep + scoring of what each state holds now;
 * marks the states terminal.  h_rewards [n][2] f32 and h_detail [n][8] i32 (cards, denari, primiera sum, score of
 * player 0 / 1, interleaved: c0 c1 d0 d1 p0 p1 s0 s1) may be NULL. */
int ms_full_evaluate_host(ms_full_state* h_states, float* h_rewards, int32_t* h_detail, int64_t n);

/* -------------------------------------------------------------------------------- SDCFR ------
 * Advantage network = FlexibleNet mlp 34 -> 128 -> 64 -> 16 with ReLU (src/algorithms/deep_cfr/nets.py:151-235,
 * :296-331; deep_cfr.py:24-52).  A net is passed as ONE fp32 blob of 13776 floats in nn.Linear order:
 * w1[128][34] b1[128] w2[64][128] b2[64] w3[16][64] b3[16].
 * precision 0 = fp32 on CUDA cores in the oracle's summation order (parity path);
 * precision 1 = bf16 operands / fp32 accumulation on the tensor cores (tcgen05.mma, TMEM accumulators).
 *
 * ms_mlp_forward: batched AdvantageNetwork.get_advantages (deep_cfr.py:54-68) + positive_regret_policy
 *   (nets.py:93-101): d_feat [n][34], d_mask [n][16] -> d_adv [n][16] (masked: adv*m - 1e6*(1-m)),
 *   d_pol [n][16]; either output may be NULL.
 * ms_sdcfr_traverse: n_trav independent calls of DeepCFR._external_sampling_cfr(root, player)
 *   (deep_cfr.py:284-365) advanced level by level (the frontier of a level is one batched inference).
 *   Opponent actions are sampled from the Philox "SDCF" stream with traversal ids first_trav.. .
 *   Emits ms_sdcfr_samples_per_traversal(player) (= 41) training samples per traversal, what
 *   AdvantageNetwork.add_experience stores (:70-75): d_feat [.][34], d_target [.][16] (regrets / max|regret|),
 *   d_mask [.][16]; d_root_value [n_trav] (may be NULL) = the returned root values.
 *   d_workspace: at least ms_sdcfr_workspace_bytes(n_trav) bytes of device memory. */
int ms_sdcfr_samples_per_traversal(int player);
size_t ms_sdcfr_workspace_bytes(int64_t n_trav);
int ms_mlp_forward(const float* d_net, int precision, const float* d_feat, const float* d_mask, float* d_adv,
                   float* d_pol, int64_t n, void* stream);
int ms_sdcfr_traverse(const ms_state* h_root, uint32_t hand_order, int player, const float* d_net0, const float* d_net1,
                      int precision, int64_t n_trav, uint64_t philox_seed, uint64_t first_trav, void* d_workspace,
                      size_t workspace_bytes, float* d_feat, float* d_target, float* d_mask, float* d_root_value,
                      void* stream);
/* ms_sdcfr_infer_states: the inference half of one traversal level on its own (sd_level_mlp_kernel): n packed states, all
 *   with `player_to_move` to move -> d_raw [n][16], the advantage net's outputs before masking (features of the mover's
 *   view, deep_cfr.py:213-282).  precision 0 = fp32 CUDA cores (bit-identical to ms_mlp_forward precision 0 on the same
 *   features), 1 = tcgen05 (bit-identical to ms_mlp_forward precision 1: same operands, same MMAs).  Used by the tests to
 *   pin the level kernel to ms_mlp_forward. */
int ms_sdcfr_infer_states(const ms_state* d_states, int64_t n, int player_to_move, const float* d_net, int precision,
                          float* d_raw, void* stream);

/* ms_sdcfr_train: `epochs` optimiser steps of AdvantageNetwork.train (deep_cfr.py:77-110) in ONE launch: per epoch
 *   gather the minibatch d_idx[epoch][0..batch) (rows of the replay buffer d_feat [n_rows][34], d_target / d_mask
 *   [n_rows][16]; batch <= 128, rows distinct = random.sample), forward, loss = MSELoss(pred * mask, target * mask),
 *   backward, clip_grad_norm_(max_norm), Adam(lr, (beta1, beta2), eps) without weight decay or amsgrad.  d_net (the
 *   fp32 blob above) and Adam's d_adam_m / d_adam_v [13776] are updated in place; steps_done = optimiser steps taken
 *   before this call (bias correction).  d_loss [epochs] receives each minibatch's loss; an epoch whose indices are
 *   out of range writes NaN there and changes nothing.  d_workspace: ms_sdcfr_train_workspace_bytes() bytes.
 *   fp32 throughout; weights, minibatch and activations stay in the shared memory of one SM across the epochs. */
size_t ms_sdcfr_train_workspace_bytes(void);
int ms_sdcfr_train(float* d_net, float* d_adam_m, float* d_adam_v, int64_t steps_done, const float* d_feat,
                   const float* d_target, const float* d_mask, int64_t n_rows, const int32_t* d_idx, int32_t batch,
                   int32_t epochs, double lr, double beta1, double beta2, double eps, double max_norm, float* d_loss,
                   void* d_workspace, size_t workspace_bytes, void* stream);

/* ms_sdcfr_train_cluster: the same optimiser steps as ms_sdcfr_train (same arguments; d_workspace is not used) on a
 *   thread-block cluster of 8 CTAs = 8 SMs: CTA c runs forward / backward on minibatch rows 16c..16c+15 with its own
 *   copy of the weights, the eight partial gradients are added through distributed shared memory, CTA c applies Adam
 *   to parameter slice c and stores the new values into all eight weight copies; three cluster barriers per step.
 *   Gradients are summed in a different order than in ms_sdcfr_train, so the two agree to fp32 rounding, not bitwise. */
int ms_sdcfr_train_cluster(float* d_net, float* d_adam_m, float* d_adam_v, int64_t steps_done, const float* d_feat,
                           const float* d_target, const float* d_mask, int64_t n_rows, const int32_t* d_idx, int32_t batch,
                           int32_t epochs, double lr, double beta1, double beta2, double eps, double max_norm,
                           float* d_loss, void* d_workspace, size_t workspace_bytes, void* stream);

/* ms_sdcfr_sample_rows: the minibatches of AdvantageNetwork.train (random.sample(self.buffer, batch_size) once per
 *   epoch, deep_cfr.py:88) for all epochs of a call in one launch: d_idx [epochs][batch] receives, per epoch, `batch`
 *   DISTINCT rows of [0, n_rows), every such batch equally likely, as a function of (seed, first_epoch + epoch, n_rows,
 *   batch) alone: Philox4x32-10 with key = seed, ctr = (step lo, step hi, position | attempt << 8, "SDTR"),
 *   row = mulhi32(x0, n_rows), positions that collide with a lower position draw again.  first_epoch = optimiser steps
 *   done before the call.  batch <= 128, batch <= n_rows < 2^31.  Feed d_idx to ms_sdcfr_train / _cluster. */
int ms_sdcfr_sample_rows(int32_t* d_idx, int32_t batch, int32_t epochs, int64_t n_rows, uint64_t seed, uint64_t first_epoch,
                         void* stream);

/* ms_sdcfr_average_policy: StrategyBuffer.get_average_policy (deep_cfr.py:136-160) for n_rows states and ALL n_nets
 *   stored strategy nets at once: d_policy[row] = sum over k (ascending) of positive_regret_policy(net_k(d_feat[row]),
 *   d_mask[row]) * d_weights[k], positive_regret_policy = relu(adv) * mask / max(sum, 1e-8) (nets.py:93-101).
 *   d_nets [n_nets][13776] fp32 blobs; d_weights [n_nets] = weight_k / total_weight as fp32 (the reference multiplies
 *   a float32 array by that Python float); d_feat [n_rows][34], d_mask [n_rows][16], d_policy [n_rows][16].
 *   Two launches: one CTA per net (net in shared memory, rows in chunks of 64), then the sum over nets in k order.
 *   d_workspace: ms_sdcfr_average_policy_workspace_bytes(n_nets, n_rows) bytes.  The empty buffer (uniform policy,
 *   :138-141) is the caller's case: n_nets must be >= 1. */
size_t ms_sdcfr_average_policy_workspace_bytes(int32_t n_nets, int64_t n_rows);
int ms_sdcfr_average_policy(const float* d_nets, const float* d_weights, int32_t n_nets, const float* d_feat,
                            const float* d_mask, int64_t n_rows, float* d_policy, void* d_workspace,
                            size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------- multi-deal MCCFR ------
 * The reference solves one fixed deal (MCCFRTrainer(game), game.new_initial_state() = seed 42,
 * src/algorithms/mc_cfr.py:88-92).  These entry points run the same _sample estimator (:37-86) on a game whose
 * root is a uniform chance node over n_deals deals (MiniScopaEnv.reset(seed_d)), with ONE infoset table for all
 * deals in device memory: open addressing on the 64-bit infoset key, 128 bytes per infoset, nodes created
 * on first touch like mc_cfr.py:32-35.  Infosets are merged by information content (player, hand set, ordered
 * table).  Table columns are indexed like the reference's arrays (by card id), compacted to the cards in hand:
 * column = rank of the card id inside the hand mask.  No reference solver does this; with
 * n_deals = 1 it is ms_mccfr_batch + ms_mccfr_apply on the same Philox streams.
 *
 * ms_md_create: d_seeds = n_deals device int64 seeds (dealt with ms_deal_from_seeds); capacity = 2^log2_capacity.
 * ms_md_mccfr_batch: traversals first_trav .. first_trav+n_trav-1; traversal t plays deal
 *   mulhi32(x0, n_deals), x = Philox4x32-10(key = philox_seed, ctr = (t lo, t hi, 0, "DEAL")), for player
 *   0 / 1 / both (2), sampling like ms_mccfr_batch ("MCCF"+player stream).  Strategies are frozen for the launch;
 *   deltas go to the table's delta columns with fp64 atomics and are folded in by ms_md_apply.
 * ms_md_counters: [0] updates [1] node visits [2] env steps [3] infosets in the table [4] error flag; returns
 *   MS_ERR_CAPACITY when an insert found the table full (synchronises the stream).
 * ms_md_export: every occupied slot, in table order (unordered): d_keys [max_n], d_regret / d_strategy [max_n][4].
 * ms_md_lookup: d_keys [n] -> d_regret / d_strategy [n][4] (zeros when absent), d_found [n]; outputs may be NULL. */
typedef struct ms_mdsolver ms_mdsolver;
int ms_md_create(const int64_t* d_seeds, int64_t n_deals, int32_t log2_capacity, void* stream, ms_mdsolver** out);
void ms_md_destroy(ms_mdsolver* s);
int ms_md_reset(ms_mdsolver* s, void* stream);
int ms_md_info(const ms_mdsolver* s, int64_t* n_deals, int64_t* capacity, int64_t* table_bytes);
int ms_md_mccfr_batch(ms_mdsolver* s, int32_t player, int64_t n_trav, uint64_t philox_seed, uint64_t first_trav,
                      void* stream);
/* ms_md_mccfr_blocked: the deal-blocked form.  Visit v (first_visit .. first_visit+n_visits-1) plays deal
 *   mulhi32(x0, n_deals), x = Philox4x32-10(key = philox_seed, ctr = (v lo, v hi, 1, "DEAL")), with pairs_per_visit
 *   traversals whose global ids are v * pairs_per_visit + i; one CTA stages the deal (enumerated tree, frozen strategies
 *   of its infosets, private delta tables) in shared memory, runs the visit on chip like ms_mccfr_batch, and writes the
 *   deltas back to the table once.  The first call describes every deal's tree and creates all its infosets with more
 *   than one action in the table (so ms_md_export then lists untouched ones too, with zeros).  Same estimator, streams
 *   and ms_md_apply as ms_md_mccfr_batch; with n_deals = 1 it equals ms_mccfr_batch on the same traversal ids. */
int ms_md_mccfr_blocked(ms_mdsolver* s, int32_t player, int64_t first_visit, int64_t n_visits, int32_t pairs_per_visit,
                        uint64_t philox_seed, void* stream);
int ms_md_apply(ms_mdsolver* s, void* stream);
int ms_md_counters(ms_mdsolver* s, uint64_t h_out[5], int reset, void* stream);
int ms_md_export(ms_mdsolver* s, uint64_t* d_keys, double* d_regret, double* d_strategy, int64_t max_n, int64_t* h_n,
                 void* stream);
int ms_md_lookup(ms_mdsolver* s, const uint64_t* d_keys, int64_t n, double* d_regret, double* d_strategy,
                 uint8_t* d_found, void* stream);
/* Table sharded over the GPUs of one box (last sentence: "infosets are shared across deals ... shard by
 * hash(key) % G with an all-to-all of deltas").  One process per GPU creates its own ms_mdsolver with the SAME seeds;
 * 2^log2_capacity is then the capacity of one SHARD.  Infoset `key` lives on rank mulhi32(hi32(key * 0xD6E8FEB86659FD93), G),
 * in that rank's shard; every rank maps every shard (CUDA IPC over NVLink / NVSwitch).  The "all-to-all" is not a
 * separate collective: ms_md_mccfr_blocked gathers the frozen regrets of a deal's infosets straight from the owners'
 * shards and sends its deltas to them as fp64 RED.ADDs through peer memory, inside the traversal kernel.  One iteration:
 *     ms_md_mccfr_blocked(this rank's visits) ; ms_md_peer_barrier ; ms_md_apply (own shard) ; ms_md_peer_barrier.
 * Visit ids are global, so the union of all ranks' visits -- and the table -- does not depend on G (up to the order
 * of the fp64 additions).
 * ms_md_ipc_export: two 64-byte IPC handles (table shard, dirty bitmap + barrier flags) into handles128.
 * ms_md_ipc_attach: handles = world x 128 bytes in rank order; before the first traversal, on a fresh table.  Afterwards
 *   ms_md_reset is refused (MS_ERR_STATE), ms_md_export lists THIS rank's shard (max_n = 0: count only, pointers may be
 *   NULL), ms_md_lookup reads any rank's shard, ms_md_counters[3] counts the infosets this rank created.
 * ms_md_peer_barrier: stream-ordered barrier across the attached ranks (a no-op when not attached); a peer that does
 *   not arrive within 2 s sets a sticky error word instead of hanging the GPU: ms_md_peer_error returns MS_ERR_STATE
 *   and *h_err = 1 + that rank (synchronises). */
int ms_md_ipc_export(ms_mdsolver* s, void* handles128);
int ms_md_ipc_attach(ms_mdsolver* s, int32_t rank, int32_t world, const void* handles);
int ms_md_peer_barrier(ms_mdsolver* s, void* stream);
int ms_md_peer_error(ms_mdsolver* s, uint32_t* h_err, void* stream);
/* measurement hook: random-access ceilings of the table's pattern over a zeroed buffer of 2^log2_lines 128-byte
 * lines, 148 x 768 threads: h_out[0] dependent 64-byte reads/s (one in flight per thread), [1] independent 64-byte
 * reads/s (8 in flight per thread), [2] random lines/s receiving four fp64 RED.ADDs */
int ms_debug_random_access_peaks(int32_t log2_lines, double h_out[3], void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SCOPA_B200_H */
