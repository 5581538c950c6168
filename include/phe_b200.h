/*
 * phe_b200.h -- C ABI of the B200 (sm_100a) hand-evaluation / equity library.
 *
 * Drop-in boundary for the one data-parallel hot path of
 * weipeilun/texasholdempocker-rl (paths below are relative to that tree):
 *   evaluate_cards            env/game.py:10,675   (third-party phevaluator)
 *   GameEnv.get_winner_set_v2 env/game.py:663-681
 *   generate_game_result      env/game.py:551-558
 *   simulate_processes        env/workflow.py:102-227 (equity plans, config/default.yml:73-133)
 *   GameEnv.new_random        env/game.py:119-163  (uniform re-deal of hidden cards)
 *
 * Conventions
 *   - plain C, no torch / C++ types; every entry point returns 0 or a negative
 *     PHE_E* status, never throws, never exits; phe_strerror() names a status.
 *   - "device" entry points take DEVICE pointers owned by the caller and launch
 *     asynchronously on `stream` (a cudaStream_t passed as void*, NULL = legacy
 *     default stream) of the CURRENT device; nothing is synchronised or freed.
 *   - "*_host" entry points take HOST pointers, stage through library-owned
 *     pinned/device scratch, and return after the result is in the host buffer.
 *   - the library owns only read-only lookup tables, built once per device.
 *   - re-entrant; callable from many threads; CUDA is initialised lazily (safe
 *     to load before fork()).
 *
 * Data formats
 *   card id      : figure*4 + decor, figure 0 ('2') .. 12 ('A'), decor 0 C 1 D 2 H 3 S
 *                  (utils/pokerhandevaluator_utils.py:3-4, env/constants.py:14-36).
 *   card set     : uint64, bit 16*decor + figure  (four 16-bit suit lanes).
 *   rank         : uint16, 1 (royal flush) .. 7462 (7-5-4-3-2 off-suit); LOWER wins
 *                  (env/game.py:676-680).
 *   equity state : two uint64 words {known, hero}: known = hero | revealed board
 *                  (board size 0, 3, 4 or 5); hero = the two hole cards; bits 61-63
 *                  of `hero` (bit 60 is the ace of spades) optionally hold a per-state
 *                  opponent count 1..7 (0 = use the call's n_opp).
 *   counts       : uint64[3] per state = {win, tie, loss}; tie = hero shares the
 *                  minimum rank with at least one opponent (env/game.py:551-558).
 */
#ifndef PHE_B200_H
#define PHE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PHE_OK            0
#define PHE_EINVAL       -1   /* bad argument (size, layout, null pointer) */
#define PHE_ECUDA        -2   /* CUDA runtime error; see phe_last_cuda_error() */
#define PHE_ENODEVICE    -3   /* no usable sm_100 device */
#define PHE_EPLAN        -4   /* equity plan not expressible (see phe_equity_plan) */
#define PHE_ENOMEM       -5

#define PHE_LAYOUT_IDS    0   /* uint8[n][7] card ids */
#define PHE_LAYOUT_MASKS  1   /* uint64[n] card sets  */

#define PHE_NUM_RANKS     7462
#define PHE_NUM_HANDS_7   133784560ull   /* C(52,7) */
#define PHE_MAX_PLAYERS   9    /* hero + 8 opponents: 5 + 2*8 = 21 dealt cards at most */
#define PHE_MAX_PLAN_STEPS 8   /* rows of an equity plan */

const char *phe_strerror(int status);
const char *phe_last_cuda_error(void);       /* thread-local text of the last PHE_ECUDA */
int phe_version(void);

/* Build (once) and upload the lookup tables for `device`, make it current. */
int phe_init(int device);

/* ---- evaluate_cards (env/game.py:675), batched ------------------------- */
/* n seven-card hands -> n ranks.  Duplicate / out-of-range cards are the
 * caller's error (checked in the Python layer, as phevaluator does). */
int phe_eval7(const void *hands, int layout, int64_t n, uint16_t *ranks_out, void *stream);
int phe_eval7_host(const void *hands, int layout, int64_t n, uint16_t *ranks_out);

/* ---- exhaustive C(52,7) ------------------------------------------------- */
/* Evaluates the 7-card combinations with colex index in [comb_begin, comb_end)
 * (index = sum_i C(c_i, i+1) over sorted ids c_0 < ... < c_6) and ADDS to
 * hist9[9] (SF, quads, full house, flush, straight, trips, two pair, pair,
 * high card) and hist7463[7463] (bin = rank; bin 0 stays 0).  Either pointer
 * may be NULL.  Shards of the index range may be summed in any order. */
int phe_enum7(uint64_t *hist9, uint64_t *hist7463, uint64_t comb_begin, uint64_t comb_end, void *stream);
int phe_enum7_host(uint64_t *hist9, uint64_t *hist7463, uint64_t comb_begin, uint64_t comb_end);
/* Same range, but only part `part` of `nparts` interleaved parts: the range is cut into tiles (1-4 work units of at
 * most 512 hands and 4 enumeration blocks each, a static table of the colex axis) and tile t belongs to part
 * t % nparts.  Cost per hand varies along the axis (runs of short blocks), so interleaved parts cost the same while
 * contiguous halves do not; the parts' histograms add up to phe_enum7's exactly.  This is how the range is sharded
 * over GPUs. */
int phe_enum7_part(uint64_t *hist9, uint64_t *hist7463, uint64_t comb_begin, uint64_t comb_end,
                   uint32_t part, uint32_t nparts, void *stream);

/* Fused enumeration + all-reduce over peer memory (one kernel per GPU, no separate collective).  Rank `rank` of
 * `world` (<= 8) enumerates its interleaved part of the range; every CTA adds its histogram into a per-launch u32
 * array with one TMA bulk reduction, and the last CTA to finish pushes that 30 KB array (a whole C(52,7) bin fits
 * 32 bits) into its slot of EVERY rank's workspace with peer stores over NVLink, raises a system-scope flag on each
 * peer, waits for all peers' flags and sums the slots into out_hist7472 (device pointer, local: 9 category bins followed
 * by the 7463 rank bins, uint64).  peer_ws[r] is rank r's workspace (device pointer valid in this process:
 * CUDA IPC / torch symmetric memory), each phe_enum7_allreduce_workspace() bytes, zeroed once before first use.
 * The launch count (epoch) lives in the workspace, so the call may be captured in a CUDA graph and replayed; all
 * ranks must issue the same sequence of calls.  A peer that does not show up within ~2 s does not hang the kernel: the
 * output then holds a PARTIAL sum and the launch number is recorded in a status word of the workspace, which
 * phe_enum7_allreduce_status reads back (synchronises `stream`): *launches_out = launches completed on this rank,
 * *failed_launch_out = number of the last launch whose exchange timed out (0 = none so far). */
uint64_t phe_enum7_allreduce_workspace(void);
int phe_enum7_allreduce(uint64_t *out_hist7472, uint64_t comb_begin, uint64_t comb_end, int rank, int world,
                        void *const *peer_ws, void *stream);
/* The same through NVLink SHARP (NVLS): multicast_ws is the MULTICAST mapping of the ranks' workspaces (CUDA multicast
 * object / torch symmetric memory `multicast_ptr`).  Each rank adds its payload with multimem.red.add -- the switch
 * applies one reduction to every rank's copy, so nothing is stored `world` times and no slots are summed -- and signals
 * every rank with ONE multimem.st.  multicast_ws == NULL is phe_enum7_allreduce.  All ranks must use the same variant. */
int phe_enum7_allreduce_nvls(uint64_t *out_hist7472, uint64_t comb_begin, uint64_t comb_end, int rank, int world,
                             void *const *peer_ws, void *multicast_ws, void *stream);
int phe_enum7_allreduce_status(const void *my_ws, void *stream, uint32_t *launches_out, uint32_t *failed_launch_out);

/* ---- Monte-Carlo equity ------------------------------------------------- */
/* For each state: deal the missing board cards and two cards to each of n_opp
 * opponents, `rollouts` times, global sample indices sample_offset ..
 * sample_offset+rollouts-1, and ADD {win, tie, loss} to wtl[3*state].  The deal
 * is a pure function of (seed, state index, sample index): Philox4x32-10, key =
 * seed, counter = (sample lo, sample hi, state_index_base + state, block);
 * results do not depend on grid shape or on how samples are sharded. */
int phe_equity_mc(const uint64_t *states, int64_t n_states, int n_opp, uint64_t rollouts,
                  uint64_t sample_offset, uint64_t seed, uint32_t state_index_base,
                  uint64_t *wtl, void *stream);
int phe_equity_mc_host(const uint64_t *states, int64_t n_states, int n_opp, uint64_t rollouts,
                       uint64_t sample_offset, uint64_t seed, uint32_t state_index_base,
                       uint64_t *wtl);

/* ---- exact-plan equity (simulate_processes, env/workflow.py:150-199) ---- */
/* Heads-up.  exist[n_tasks][8]: the task's known cards as ids in deal order
 * (hero 2, then flop 3, turn 1, river 1 as far as revealed), n_exist of them
 * valid (2, 5, 6 or 7; same for the whole call).  plan[n_steps][3] (HOST
 * pointer) = (gen_num, num_random, is_segment_end) rows of
 * simulation_recurrent_params; every step but the last must have gen_num 1,
 * and n_exist + sum(gen_num) must be 9, else PHE_EPLAN.  ADDS {win, tie, loss}
 * to wtl[3*task]; num_estimation = win+tie+loss, game_result_sum = win - loss.
 * The random step draws from the Philox stream with stream id
 * stream_base + task and sample index leaf*num_random + draw. */
int phe_equity_plan(const uint8_t *exist, int n_exist, int64_t n_tasks, const int32_t *plan,
                    int n_steps, uint64_t seed, uint32_t stream_base, uint64_t *wtl, void *stream);
int phe_equity_plan_host(const uint8_t *exist, int n_exist, int64_t n_tasks, const int32_t *plan,
                         int n_steps, uint64_t seed, uint32_t stream_base, uint64_t *wtl);
/* number of deals the plan generates for n_exist known cards (0 if invalid) */
uint64_t phe_plan_num_deals(int n_exist, const int32_t *plan, int n_steps);

/* ---- showdown (get_winner_set_v2, env/game.py:663-681) ------------------ */
/* boards[n_games] (5-card sets), holes[n_games][n_players] (2-card sets),
 * live[n_games] = bit p set if player p takes part.  winners_out[g] = bitmask
 * of the live players holding the minimum rank; a game with one live player
 * returns that player without evaluation (env/game.py:664-665).  ranks_out
 * (optional) [n_games][n_players], 0 for players not evaluated. */
int phe_showdown(const uint64_t *boards, const uint64_t *holes, const uint32_t *live,
                 int64_t n_games, int n_players, uint32_t *winners_out, uint16_t *ranks_out, void *stream);
int phe_showdown_host(const uint64_t *boards, const uint64_t *holes, const uint32_t *live,
                      int64_t n_games, int n_players, uint32_t *winners_out, uint16_t *ranks_out);

/* ---- determinisation (GameEnv.new_random, env/game.py:119-163) ---------- */
/* For each of n deals: draw `need` (1..21) distinct cards uniformly from the cards not
 * in known[i] (same Philox stream as phe_equity_mc: stream id
 * stream_base + i, sample index sample_offset) and write their ids to
 * cards_out[i*need ..]: the single card first when `need` is odd, then need/2
 * unordered pairs, each as (lower id, higher id).  Any value-independent
 * assignment of these slots to roles is a uniform deal. */
int phe_deal(const uint64_t *known, int64_t n, int need, uint64_t seed, uint64_t sample_offset,
             uint32_t stream_base, uint8_t *cards_out, void *stream);
int phe_deal_host(const uint64_t *known, int64_t n, int need, uint64_t seed, uint64_t sample_offset,
                  uint32_t stream_base, uint8_t *cards_out);
/* Same deal routine, rows numbered along the SAMPLE axis: row i is sample (sample_offset + i) of the one stream
 * `stream_id`.  This is the form a caller with one logical stream (a game, a search tree) wants: n re-deals of
 * GameEnv.new_random (env/game.py:119-163) are n consecutive samples, and two stream ids never share a deal. */
int phe_deal_samples(const uint64_t *known, int64_t n, int need, uint64_t seed, uint64_t sample_offset,
                     uint32_t stream_id, uint8_t *cards_out, void *stream);
int phe_deal_samples_host(const uint64_t *known, int64_t n, int need, uint64_t seed, uint64_t sample_offset,
                          uint32_t stream_id, uint8_t *cards_out);

/* ---- terminal settlement (finish_game / _share_value, env/game.py:410-441, 560-632, 683-734) ---- */
/* Per game: boards[g] (5-card set), holes[g][P] (2-card sets, every seat), actions[g][R][P] = last action of each
 * betting round (0 None, 1 FOLD, 2 any other), values[g][R][P] = chips put in during the round, n_rounds[g] <= R,
 * button[g] = seat index, onboard[g] = bit p set if seat p is still ONBOARD (odd-chip rule).  Outputs per seat: chips
 * won, game result and pure card result in {-1 LOSE, 0 EVEN, 1 WIN}; paid_out[g] (optional) = seats present in the
 * reference's winner_value_dict.  P <= 9, R <= 4. */
int phe_settle(const uint64_t *boards, const uint64_t *holes, const int8_t *actions, const int32_t *values,
               const int32_t *n_rounds, const int32_t *button, const uint32_t *onboard, int64_t n_games,
               int n_players, int max_rounds, int small_blind, int32_t *won_out, int8_t *result_out,
               int8_t *card_result_out, uint32_t *paid_out, void *stream);
int phe_settle_host(const uint64_t *boards, const uint64_t *holes, const int8_t *actions, const int32_t *values,
                    const int32_t *n_rounds, const int32_t *button, const uint32_t *onboard, int64_t n_games,
                    int n_players, int max_rounds, int small_blind, int32_t *won_out, int8_t *result_out,
                    int8_t *card_result_out, uint32_t *paid_out);

/* ---- policy / value network glue (PolicyValueNet: rl/models.py:12-115, 402-481) ---------------------------- */
#define PHE_DTYPE_F32   0
#define PHE_DTYPE_BF16  1
/* out[r] = LayerNorm(x[r] + residual[r]) * gamma + beta over rows of d contiguous elements (residual may be NULL);
 * the post-LN steps of nn.TransformerEncoderLayer as built at rl/models.py:446-447.  Statistics in fp32, biased
 * variance, eps inside the square root.  Pointers 16-byte aligned, d a multiple of 16 bytes, d <= 2048 (bf16) /
 * 1024 (f32). */
int phe_nn_add_layernorm(const void *x, const void *residual, const void *gamma, const void *beta, void *out,
                         int64_t rows, int d, float eps, int dtype, void *stream);
/* Encoder input of EnvEmbedding.forward (rl/models.py:100-114) with the sqrt(d_model) scaling of rl/models.py:462-466
 * folded into the tables: out[b][s] = starters[s] for s < n_starters, else
 * field_tab[obs[b][f] + field_start[f]] (de elements) followed by pos_tab[f] (dp elements), f = s - n_starters.
 * obs int32[batch][n_fields]; out [batch][n_starters + n_fields][de + dp].  Table indices are clamped to
 * [0, n_rows) for memory safety only. */
int phe_nn_embed_obs(const int32_t *obs, const int32_t *field_start, const void *field_tab, const void *pos_tab,
                     const void *starters, void *out, int64_t batch, int n_fields, int n_starters, int de, int dp,
                     int n_rows, int dtype, void *stream);

/* ---- observation encoder (Env.get_obs, env/env.py:416-581) -------------------------------------------------- */
#define PHE_OBS_CUTTERS      16   /* cutter instances of env/env.py:52-81, in the order listed below */
#define PHE_OBS_MAX_THR      20   /* longest threshold list (CUTTER_SELF_DEFAULT_LIST has 19) */
#define PHE_OBS_CARD_NONE    52   /* card slot not dealt yet   -> (figure 14, decor 5), env/env.py:709 */
#define PHE_OBS_CARD_HIDDEN  53   /* dealt but not observable  -> (figure 13, decor 4), env/env.py:717-719 */
/* n observations -> obs_out int32[n][F], F = phe_obs_num_fields(n_bins, n_other, n_hist) (93 for the reference's
 * heads-up configuration: n_bins 10, n_other 1, n_hist 24), field order of Env._map_obs_idx_to_model_index
 * (env/env.py:383-408).  All inputs are device pointers:
 *   round_role int32[n][2]   current round, acting player's position relative to the button (env/env.py:441-443)
 *   cards      uint8[n][7]   hand 2, flop 3, turn, river as ids, or PHE_OBS_CARD_NONE / PHE_OBS_CARD_HIDDEN
 *   values     int64[n][5]   game_init_value, pot_value, acting player's init / left / history-bet value
 *   bins       int64[n][n_bins]  infoset.bin_bet_value_list (-1 = bin not available)
 *   others     int64[n][n_other][3]  status, value left, init value of the other players by relative position
 *   history    uint8[n][n_hist]  acting player's, then the opponent's action bins, padded with num_action_bins
 *   cut_table  { double thr[16][20]; int32 len[16]; }: threshold lists of the cutters, in this order:
 *              0 init/game_init, 1 left/game_init, 2 left/init, 3 bet/pot, 4-6 bet/{game_init, init, left},
 *              7-9 pot/{game_init, init, left}, 10 action/pot, 11-13 action/{game_init, init, left} (10-13 have the
 *              extra invalid bin), 14 other left/init, 15 other init/acting init.
 * A ratio is an IEEE double division of the two integers, compared as tools/biner.py:8-27 does. */
int phe_obs_num_fields(int n_bins, int n_other, int n_hist);
uint64_t phe_obs_cut_table_bytes(void);
int phe_encode_obs(const int32_t *round_role, const uint8_t *cards, const int64_t *values, const int64_t *bins,
                   const int64_t *others, const uint8_t *history, const void *cut_table, int64_t n, int n_bins,
                   int n_other, int n_hist, int32_t *obs_out, void *stream);

/* softmax(Q K^T / sqrt(head_dim)) V per (batch, head) for sequences of at most 112 tokens (the model has 98), bf16:
 * the self-attention of nn.TransformerEncoderLayer (rl/models.py:446-447).  q [batch][s_q][heads*head_dim] and
 * k, v [batch][s_kv][heads*head_dim] are addressed with element strides (batch stride, row stride), so the packed
 * output of the fused in-projection can be used in place; out is contiguous [batch][s_q][heads*head_dim].
 * head_dim 16 or 64; pointers and strides multiples of 16 bytes. */
int phe_nn_attention(const void *q, const void *k, const void *v, void *out, int64_t batch, int heads, int s_q,
                     int s_kv, int head_dim, int64_t q_batch_stride, int64_t q_row_stride, int64_t kv_batch_stride,
                     int64_t kv_row_stride, int dtype, void *stream);

/* What SingleThreadMCTS.predict consumes (rl/MCTS.py:464-473) from the encoder rows of tokens 0 and 1:
 * out[b][0..n_actions) = softmax(action_prob_dense(h[b][0])), out[b][n_actions] = reward_value_dense(h[b][1]), fp32.
 * h [batch][2][row_stride >= d]; weights as nn.Linear stores them (w_policy [n_actions][d], w_value [1][d]);
 * n_actions <= 32. */
int phe_nn_policy_value_heads(const void *h, const void *w_policy, const void *b_policy, const void *w_value,
                              const void *b_value, float *out, int64_t batch, int d, int n_actions,
                              int64_t row_stride, int dtype, void *stream);

/* number of kernels this library has launched in this process (bench.py "gpu_launches") */
uint64_t phe_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* PHE_B200_H */
