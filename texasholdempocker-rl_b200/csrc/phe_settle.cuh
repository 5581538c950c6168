// phe_settle.cuh -- terminal settlement of one game: the integer logic of the reference's
// GameEnv.finish_game / _share_value / share_pot_for_winner / share_value_to_winner
// (env/game.py:410-441, 560-632, 702-734) on seat-indexed arrays and winner bitmasks.
//
// The data-parallel part (one 7-card evaluation per player) is done by the caller; this routine only compares
// ranks.  Semantics kept from the reference: all-in layers of a betting round are resolved in ascending stake
// (ties in seat order); a seat whose action in a round is None still contends (the reference tests `!= FOLD`);
// the game result is the one written by the last showdown a seat won; chips are floored to small-blind multiples
// and the odd chips go to the button if it won, else to the next ONBOARD seat after the button if that won.
// Two deliberate fixes for tables of more than two seats, where the reference itself cannot finish: (1) where its
// odd-chip loop would not terminate (env/game.py:731-733) the chips go to the first winner clockwise from the button;
// (2) a final pot left without a contender (every seat still in the hand was settled as an all-in layer; the reference
// divides by an empty winner set, env/game.py:721) goes to the best hand among the seats that have not folded.
#pragma once
#include "phe_core.cuh"

namespace phe {

constexpr int SETTLE_MAX_PLAYERS = 9;
constexpr int SETTLE_MAX_ROUNDS = 4;
constexpr int ACT_NONE = 0, ACT_FOLD = 1, ACT_LIVE = 2;
constexpr int RES_LOSE = -1, RES_EVEN = 0, RES_WIN = 1;

PHE_HD uint32_t settle_winners(const uint32_t* ranks, uint32_t contenders, int P)
{
    if (popc32(contenders) <= 1) return contenders;   // env/game.py:664-665: a lone contender wins unevaluated
    uint32_t best = 0xFFFFFFFFu, w = 0;
    for (int p = 0; p < P; p++)
        if (contenders >> p & 1) {
            if (ranks[p] < best) { best = ranks[p]; w = 0; }
            if (ranks[p] == best) w |= 1u << p;
        }
    return w;
}

PHE_HD void settle_mark(uint32_t winners, int8_t* result, int P)
{
    const int8_t r = popc32(winners) > 1 ? (int8_t)RES_EVEN : (int8_t)RES_WIN;
    for (int p = 0; p < P; p++)
        if (winners >> p & 1) result[p] = r;
}

PHE_HD void settle_pay(int64_t total, uint32_t winners, int32_t* won, uint32_t* paid, int small_blind, int button, uint32_t onboard, int P)
{
    if (total <= 0 || !winners) return;
    const int n = popc32(winners);
    const int64_t share = (total / ((int64_t)n * small_blind)) * small_blind;
    for (int p = 0; p < P; p++)
        if (winners >> p & 1) won[p] += (int32_t)share;
    *paid |= winners;
    const int64_t left = total - share * n;
    if (left > 0) {
        int who = button;   // callers pass a seat index in [0, P) (k_settle reduces it)
        if (!(winners >> who & 1)) {
            int q = -1;
            for (int k = 1; k < P && q < 0; k++) {
                const int s = (button + k) % P;
                if (onboard >> s & 1) q = s;
            }
            if (q >= 0 && (winners >> q & 1)) who = q;
            else
                for (int k = 1; k <= P; k++) {
                    const int s = (button + k) % P;
                    if (winners >> s & 1) { who = s; break; }
                }
        }
        won[who] += (int32_t)left;
    }
}

// act / val are [R][P] row-major.  Outputs won[P] (chips), result[P] and card[P] in {-1, 0, 1}, *paid = seats that
// appear in the reference's winner_value_dict.
PHE_HD void settle_game(int P, int R, int button, int small_blind, uint32_t onboard, const uint32_t* ranks, const int8_t* act,
                        const int32_t* val, int32_t* won, int8_t* result, int8_t* card, uint32_t* paid)
{
    const uint32_t all = (1u << P) - 1u;
    for (int p = 0; p < P; p++) { won[p] = 0; result[p] = (int8_t)RES_LOSE; }
    *paid = 0;
    int64_t pot = 0;
    uint32_t all_shared = 0;
    for (int r = 0; r < R; r++) {
        const int8_t* a = act + r * P;
        const int32_t* v = val + r * P;
        int32_t spare[SETTLE_MAX_PLAYERS];
        int32_t top = 0;
        uint32_t not_folded = 0;
        for (int p = 0; p < P; p++) {
            spare[p] = v[p];
            top = v[p] > top ? v[p] : top;
            if (a[p] != ACT_FOLD) not_folded |= 1u << p;
        }
        uint32_t allin = 0;
        for (int p = 0; p < P; p++)
            if (a[p] == ACT_LIVE && v[p] < top) allin |= 1u << p;
        uint32_t present = all, round_shared = 0;
        while (allin) {
            int s = -1;
            for (int p = 0; p < P; p++)   // ascending stake, ties in seat order (stable sort of the reference)
                if ((allin >> p & 1) && (s < 0 || v[p] < v[s])) s = p;
            allin &= ~(1u << s);
            if ((present >> s & 1) && spare[s] > 0) {
                const uint32_t w = settle_winners(ranks, not_folded & ~round_shared, P);
                settle_mark(w, result, P);
                if (w >> s & 1) {
                    const int32_t stake = spare[s];
                    int64_t total = pot;
                    for (int p = 0; p < P; p++)
                        if (present >> p & 1) {
                            if (spare[p] > stake) { spare[p] -= stake; total += stake; }
                            else { total += spare[p]; present &= ~(1u << p); }
                        }
                    settle_pay(total, w, won, paid, small_blind, button, onboard, P);
                    pot = 0;
                } else {
                    pot += spare[s];
                    present &= ~(1u << s);
                }
            }
            round_shared |= 1u << s;
            all_shared |= 1u << s;
        }
        for (int p = 0; p < P; p++)
            if (present >> p & 1) pot += spare[p];
    }
    if (pot > 0 && R > 0) {
        const int8_t* a = act + (R - 1) * P;
        uint32_t cont = 0;
        for (int p = 0; p < P; p++)
            if (a[p] != ACT_FOLD && !(all_shared >> p & 1)) cont |= 1u << p;
        if (!cont)   // every seat still in the hand was settled as an all-in layer: the reference divides by zero (env/game.py:721)
            for (int p = 0; p < P; p++)
                if (a[p] != ACT_FOLD) cont |= 1u << p;
        const uint32_t w = settle_winners(ranks, cont, P);
        settle_mark(w, result, P);
        settle_pay(pot, w, won, paid, small_blind, button, onboard, P);
    }
    const uint32_t cw = settle_winners(ranks, all, P);
    const int8_t cr = popc32(cw) == 1 ? (int8_t)RES_WIN : (int8_t)RES_EVEN;
    for (int p = 0; p < P; p++) card[p] = (cw >> p & 1) ? cr : (int8_t)RES_LOSE;
}

}  // namespace phe
