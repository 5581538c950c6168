// phe_plan.h -- host-side parsing of an equity plan into a PlanDesc.
//
// Plan rows are (gen_num, num_random, is_segment_end) exactly as in the reference's
// simulation_recurrent_params (config/default.yml:73-133) and are interpreted the way
// simulate_processes.generate_and_calculate_recurrent does (env/workflow.py:150-199):
//   * gen_num == 1 rows enumerate one card each in increasing deck index; consecutive rows up to
//     and including the first one with is_segment_end form one unordered combination (:172-177);
//     a new segment restarts at deck index 0, so two single-row segments give ORDERED pairs;
//   * a row with gen_num > 1 draws num_random random completions of gen_num cards for every
//     enumeration leaf and ends the recursion (:181-195), so it must be the last row;
//   * the deal must reach 9 cards (hero 2, board 5, villain 2; :131-135).
#pragma once
#include "phe_core.cuh"

namespace phe {

inline uint64_t binom_u64(int n, int k)
{
    if (k < 0 || n < k) return 0;
    uint64_t r = 1;
    for (int i = 1; i <= k; i++) r = r * (uint64_t)(n - k + i) / (uint64_t)i;
    return r;
}

// returns false if the plan cannot be expressed
inline bool parse_plan(int n_exist, const int32_t* plan, int n_steps, PlanDesc* pd)
{
    if (!plan || n_steps < 1 || n_steps > 8) return false;
    if (!(n_exist == 2 || n_exist == 5 || n_exist == 6 || n_exist == 7)) return false;
    PlanDesc d{};
    d.n_exist = n_exist;
    d.n_leaves = 1;
    int cards = n_exist, seg_cards = 0;
    for (int s = 0; s < n_steps; s++) {
        const int gen = plan[3 * s], nrand = plan[3 * s + 1], seg_end = plan[3 * s + 2];
        if (gen < 1) return false;
        if (gen == 1) {
            seg_cards++;
            d.n_enum++;
            const bool last = (s == n_steps - 1);
            if (seg_end || last) {
                if (d.n_seg >= 7) return false;
                const uint64_t cnt = binom_u64(52 - cards, seg_cards);
                if (cnt == 0 || cnt > 0xFFFFFFFFull) return false;
                d.seg_k[d.n_seg] = seg_cards;
                d.seg_count[d.n_seg] = (uint32_t)cnt;
                d.n_seg++;
                if (d.n_leaves > (1ull << 62) / cnt) return false;
                d.n_leaves *= cnt;
                cards += seg_cards;
                seg_cards = 0;
            }
        } else {
            if (s != n_steps - 1 || seg_cards != 0) return false;   // random step must be last and start a segment
            if (nrand < 1 || gen > 7) return false;
            // the reference refuses plans asking for more draws than combinations exist (:187-197)
            if (binom_u64(52 - cards, gen) < (uint64_t)nrand) return false;
            d.rnd_k = gen;
            d.rnd_n = nrand;
            cards += gen;
        }
    }
    if (cards != 9) return false;
    d.pair_tail = (d.rnd_k == 0 && d.n_seg >= 1 && d.seg_k[d.n_seg - 1] == 2) ? 1 : 0;
    d.n_prefix = d.pair_tail ? d.n_leaves / d.seg_count[d.n_seg - 1] : 0;
    *pd = d;
    return true;
}

inline uint64_t plan_num_deals(const PlanDesc& pd) { return pd.n_leaves * (uint64_t)(pd.rnd_k ? pd.rnd_n : 1); }

}  // namespace phe
