// phe_hostcheck.cpp -- TEST-ONLY host instantiation of the device arithmetic in phe_core.cuh.
//
// Built into libphe_hostcheck.so and loaded only by tests/ (CPU, "not gpu"): it lets the exact
// evaluator / dealing / unranking code the sm_100a kernels execute be compared with the oracle in
// a container without a GPU.  It is not linked into libphe_b200.so and is not a CPU fallback:
// the product library has no host compute path.
#include <cstring>
#include <vector>

#include "phe_core.cuh"
#include "phe_enum.cuh"
#include "phe_plan.h"
#include "phe_settle.cuh"
#include "phe_tables.h"

using namespace phe;

static std::vector<uint16_t> g_image;
static HashParams g_hp;
static uint32_t g_binom[53 * 8];
static std::vector<unsigned char> g_enum_image;
static std::vector<uint64_t> g_trimask;
static std::vector<uint64_t> g_dealtab;

extern "C" {

int hc_init(void)
{
    if (!g_image.empty()) return 0;
    g_image.resize(TAB_U16);
    if (!build_table_image(g_image.data(), &g_hp)) { g_image.clear(); return -1; }
    build_binom(g_binom);
    g_enum_image.resize(ENUM_IMG_BYTES);
    g_trimask.resize(N_TRIPLES);
    g_dealtab.resize(DEALTAB_ENTRIES);
    build_dealtab(g_dealtab.data());
    build_enum_image(g_enum_image.data(), g_trimask.data());
    return 0;
}

void hc_get_image(uint16_t* out, uint32_t* hp4)
{
    std::memcpy(out, g_image.data(), TAB_BYTES);
    hp4[0] = g_hp.a1; hp4[1] = g_hp.b1; hp4[2] = g_hp.a2; hp4[3] = g_hp.b2;
}

int hc_num_keys(void) { return (int)all_noflush_keys().size(); }

void hc_eval7_masks(const uint64_t* m, int64_t n, uint16_t* out)
{
    GlobalTab tab{g_image.data()};
    for (int64_t i = 0; i < n; i++) out[i] = (uint16_t)eval7(m[i], tab, g_hp);
}

void hc_deal(uint64_t known, uint64_t seed, uint64_t sample, uint32_t stream, int need, uint8_t* ids)
{
    deal_ids(GlobalDealTab{g_dealtab.data()}, known, seed, sample, stream, need, ids);
}

void hc_equity_mc(uint64_t known, uint64_t hero, int n_opp, uint64_t rollouts, uint64_t offset, uint64_t seed,
                  uint32_t stream, uint64_t wtl[3])
{
    GlobalTab tab{g_image.data()};
    const int nboard = __builtin_popcountll(known & ~hero);
    LocalSlots scratch;
    const GlobalDealTab dt{g_dealtab.data()};
    for (uint64_t s = 0; s < rollouts; s++) wtl[mc_rollout(tab, g_hp, dt, known, hero, nboard, n_opp, seed, offset + s, stream, scratch)]++;
}

void hc_colex_unrank7(uint64_t idx, int32_t* c)
{
    Binom C{g_binom};
    int cc[7];
    colex_unrank7(idx, C, cc);
    for (int i = 0; i < 7; i++) c[i] = cc[i];
}

int hc_equity_plan(const uint8_t* exist, int n_exist, const int32_t* plan, int n_steps, uint64_t seed, uint32_t stream,
                   uint64_t out[4])
{
    PlanDesc pd;
    if (!parse_plan(n_exist, plan, n_steps, &pd)) return -4;
    GlobalTab tab{g_image.data()};
    Binom C{g_binom};
    uint64_t known = 0;
    for (int i = 0; i < n_exist; i++) known |= id_bit(exist[i]);
    out[0] = out[1] = out[2] = out[3] = 0;
    for (uint64_t leaf = 0; leaf < pd.n_leaves; leaf++) {
        uint8_t cards[9];
        std::memcpy(cards, exist, n_exist);
        plan_unrank_leaf(pd, leaf, C, cards);
        const int base = pd.n_exist + pd.n_enum;
        if (pd.rnd_k == 0) {
            out[plan_leaf_result(tab, g_hp, cards)]++; out[3]++;
        } else {
            uint64_t kn = known;
            for (int i = pd.n_exist; i < base; i++) kn |= id_bit(cards[i]);
            for (int j = 0; j < pd.rnd_n; j++) {
                deal_ids(GlobalDealTab{g_dealtab.data()}, kn, seed, leaf * (uint64_t)pd.rnd_n + (uint64_t)j, stream, pd.rnd_k, cards + base);
                out[plan_leaf_result(tab, g_hp, cards)]++; out[3]++;
            }
        }
    }
    return 0;
}

// serial emulation of the enumeration kernel's warp loop over the colex range [begin, end)
void hc_enum7(uint64_t begin, uint64_t end, uint64_t* hist7463)
{
    if (begin >= end) return;
    Binom C{g_binom};
    EnumImage img{g_enum_image.data(), g_trimask.data()};
    int c[7];
    colex_unrank7(begin, C, c);
    int c3 = c[3], c4 = c[4], c5 = c[5], c6 = c[6];
    uint64_t remaining = end - begin;
    uint32_t t_start = C(c[2], 3) + C(c[1], 2) + (uint32_t)c[0];
    for (;;) {
        EnumBlock b;
        enum_block_setup(c3, c4, c5, c6, C, b);
        uint64_t stop = b.ntriples;
        if (stop > t_start + remaining) stop = t_start + remaining;
        for (uint32_t t = t_start; t < stop; t++) hist7463[enum_hand_rank(img, b, t)]++;
        remaining -= stop - t_start;
        if (!remaining) break;
        t_start = 0;
        enum_block_next(c3, c4, c5, c6);
    }
}

// host instantiation of settle_game with ranks from the host-instantiated evaluator
void hc_settle(uint64_t board, const uint64_t* holes, int P, int R, int button, int small_blind, uint32_t onboard, const int8_t* act,
               const int32_t* val, int32_t* won, int8_t* result, int8_t* card, uint32_t* paid)
{
    GlobalTab tab{g_image.data()};
    uint32_t ranks[SETTLE_MAX_PLAYERS];
    for (int p = 0; p < P; p++) ranks[p] = eval7(board | holes[p], tab, g_hp);
    settle_game(P, R, button, small_blind, onboard, ranks, act, val, won, result, card, paid);
}

// the product's Philox with the key derived on the fly (SeedKey) and with the twenty round keys precomputed (RoundKeys: what
// the Monte-Carlo kernel takes as a launch parameter); out[0..4) and out[4..8) must agree with each other and with the oracle
void hc_philox(const uint32_t ctr[4], uint64_t seed, uint32_t out[8])
{
    philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], (uint32_t)seed, (uint32_t)(seed >> 32), out);
    const RoundKeys rk = round_keys(seed);
    philox4x32_10_k(ctr[0], ctr[1], ctr[2], ctr[3], rk, out + 4);
}

// work-unit boundaries of the enumeration kernel (build_enum_units); returns the count, copies at most `cap` entries
int64_t hc_enum_units(uint32_t max_hands, uint32_t max_blocks, uint32_t align, uint32_t* out, uint64_t* starts_out, int64_t cap)
{
    std::vector<uint32_t> u;
    std::vector<uint64_t> st;
    build_enum_units(u, st, max_hands, max_blocks, align);
    for (int64_t i = 0; i < (int64_t)u.size() && i < cap; i++) {
        out[i] = u[(size_t)i];
        starts_out[i] = st[(size_t)i];
    }
    return (int64_t)u.size();
}

uint64_t hc_plan_num_deals(int n_exist, const int32_t* plan, int n_steps)
{
    PlanDesc pd;
    return parse_plan(n_exist, plan, n_steps, &pd) ? plan_num_deals(pd) : 0;
}

}  // extern "C"
