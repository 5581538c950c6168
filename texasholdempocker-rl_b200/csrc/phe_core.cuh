// phe_core.cuh -- evaluator, dealing stream and enumeration arithmetic shared by all kernels.
//
// Everything here is __host__ __device__ so that the *same* code the sm_100a kernels run can
// be instantiated on the host by csrc/phe_hostcheck.cu (a test-only library) and compared
// with the oracle without a GPU.  The product library never runs these on the host.
//
// Replaces, for the hot path only:
//   phevaluator.evaluate_cards (7 cards)         reference call site env/game.py:675
//   random.sample based dealing                  env/workflow.py:190, env/game.py:139-140
//   recursive plan enumeration                   env/workflow.py:150-199
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PHE_HD __host__ __device__ __forceinline__
#else
#define PHE_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define PHE_UNROLL _Pragma("unroll")
#else
#define PHE_UNROLL
#endif

namespace phe {

// ---- lookup table image (uint16 units) --------------------------------------------------
// [0, 65536)            NOFLUSH : rank of a non-flush 7-card rank multiset, perfect-hash indexed
// [65536, 73728)        DISP    : per-bucket displacement of the perfect hash
// [73728, 81920)        FLUSH   : rank of the best flush in a 13-bit suit lane (5..7 bits set)
constexpr uint32_t NOFLUSH_SIZE = 65536;
constexpr uint32_t DISP_SIZE = 8192;
constexpr uint32_t FLUSH_SIZE = 8192;
constexpr uint32_t OFF_DISP = NOFLUSH_SIZE;
constexpr uint32_t OFF_FLUSH = NOFLUSH_SIZE + DISP_SIZE;
constexpr uint32_t TAB_U16 = NOFLUSH_SIZE + DISP_SIZE + FLUSH_SIZE;   // 81920 -> 163,840 bytes
constexpr uint32_t TAB_BYTES = TAB_U16 * 2;

// multipliers of the two hash functions (found by the table builder, verified collision free)
struct HashParams { uint32_t a1, b1, a2, b2; };

constexpr uint64_t CARD_BITS = 0x1FFF1FFF1FFF1FFFull;

PHE_HD int popc32(uint32_t x)
{
#ifdef __CUDA_ARCH__
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

PHE_HD uint32_t umulhi32(uint32_t a, uint32_t b)
{
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

// card id (figure*4 + decor) -> bit position 16*decor + figure
PHE_HD uint32_t id_to_pos(uint32_t id) { return ((id & 3u) << 4) | (id >> 2); }
PHE_HD uint32_t pos_to_id(uint32_t pos) { return ((pos & 15u) << 2) | (pos >> 4); }
PHE_HD uint64_t pos_bit(uint32_t pos) { return 1ull << pos; }
PHE_HD uint64_t id_bit(uint32_t id) { return 1ull << id_to_pos(id); }

// ---- table accessors ----------------------------------------------------------------------
struct GlobalTab {
    const uint16_t* p;
    PHE_HD uint32_t ld(uint32_t idx) const
    {
#ifdef __CUDA_ARCH__
        return __ldg(p + idx);
#else
        return p[idx];
#endif
    }
};

#ifdef __CUDACC__
struct SharedTab {
    uint32_t base;   // shared-window byte address of the table image
    __device__ __forceinline__ uint32_t ld(uint32_t idx) const
    {
        uint16_t v;
        asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(base + idx * 2u));
        return v;
    }
};
#endif

// ---- 7-card evaluator on a card-set word -----------------------------------------------------
// m must have exactly 7 card bits set.  Branch free: the flush and the non-flush index are both
// formed and one of them selects the single final table read.
template <class Tab>
PHE_HD uint32_t eval7(uint64_t m, const Tab& tab, const HashParams& hp)
{
    const uint32_t lo = (uint32_t)m, hi = (uint32_t)(m >> 32);   // lanes {D,C} and {S,H}
    // rank multiplicities, bit sliced: count(r) = b0 + 2*b1 + 4*b2
    const uint32_t x = lo ^ hi, a = lo & hi;
    const uint32_t x1 = x >> 16, a1 = a >> 16;
    const uint32_t b0 = x ^ x1;                       // low 16 bits valid
    const uint32_t cy = x & x1;
    const uint32_t b1 = a ^ a1 ^ cy;
    const uint32_t b2 = ((a & a1) | (a & cy) | (a1 & cy)) & 0xFFFFu;
    const uint32_t key = (b0 & 0xFFFFu) | (b1 << 16);
    const uint32_t h1 = key * hp.a1 + b2 * hp.b1;
    const uint32_t h2 = key * hp.a2 + b2 * hp.b2;
    const uint32_t d = tab.ld(OFF_DISP + (h1 >> 19));
    const uint32_t inf = ((h2 >> 16) + d) & 0xFFFFu;
    // flush: with 7 cards at most one suit lane holds >= 5
    const int plo = popc32(lo);
    const uint32_t w = plo >= 5 ? lo : hi;
    const uint32_t wl = w & 0xFFFFu;
    const int pl = popc32(wl), pw = popc32(w);
    const uint32_t lane = pl >= 5 ? wl : (w >> 16);
    const bool fl = (pl >= 5) | (pw - pl >= 5);
    return tab.ld(fl ? OFF_FLUSH + lane : inf);
}

PHE_HD int category_of(uint32_t rank)
{
    // last rank of SF, quads, full house, flush, straight, trips, two pair, pair
    return (rank > 10) + (rank > 166) + (rank > 322) + (rank > 1599) + (rank > 1609) + (rank > 2467) +
           (rank > 3325) + (rank > 6185);
}

// ---- Philox4x32-10 ---------------------------------------------------------------------------
PHE_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                          uint32_t out[4])
{
PHE_UNROLL
    for (int r = 0; r < 10; r++) {
        const uint32_t h0 = umulhi32(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = umulhi32(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ k0; c1 = l1;
        c2 = h0 ^ c3 ^ k1; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// ---- dealing -----------------------------------------------------------------------------------
// Draws `need` (<= 21) distinct cards outside `known` for (seed, stream, sample); exactly uniform over the
// unknown cards, a pure function of its arguments.  Philox block b = philox(counter (sample lo, sample hi,
// stream, b), key seed).  Every 32-bit output word w (words 0..3 of block 0, then block 1, ...) carries five
// draws: suit_j = (w >> 2j) & 3 from the low ten bits, and v = w >> 10 gives rank_j as the base-13 digits of v,
// least significant first.  A word with v >= 11 * 13^5 is skipped whole (2.6 %), which makes the digits exactly
// uniform.  Draw j addresses card-set bit 16*suit_j + rank_j and is accepted iff that bit is clear in the running
// used set; accepted positions are stored in order.  Words are consumed until `need` cards are accepted (or 64
// blocks have been used, which only malformed input can reach).
constexpr int MAX_DEAL = 21;
constexpr uint32_t DEAL_V_LIMIT = 11u * 371293u;   // 11 * 13^5 = 4,084,223 <= 2^22

struct LocalCards {          // host / small kernels: accepted positions in a thread-local array
    uint8_t p[MAX_DEAL];
    PHE_HD void put(int k, uint32_t pos) { p[k] = (uint8_t)pos; }
    PHE_HD uint32_t get(int k) const { return p[k]; }
};

template <class Cards>
PHE_HD void deal_cards(uint64_t known, uint64_t seed, uint64_t sample, uint32_t stream, int need, Cards& out)
{
    uint32_t used_lo = (uint32_t)known, used_hi = (uint32_t)(known >> 32);
    int got = 0;
    // 64 blocks = 1280 draws: never reached with a valid state (>= 31 of 52 cards free); bounds the loop on garbage
    for (uint32_t blk = 0; got < need && blk < 64u; blk++) {
        uint32_t rnd[4];
        philox4x32_10((uint32_t)sample, (uint32_t)(sample >> 32), stream, blk, (uint32_t)seed, (uint32_t)(seed >> 32), rnd);
PHE_UNROLL
        for (int wi = 0; wi < 4; wi++) {
            if (got >= need) break;
            const uint32_t w = rnd[wi];
            uint32_t v = w >> 10;
            if (v >= DEAL_V_LIMIT) continue;
PHE_UNROLL
            for (int j = 0; j < 5; j++) {
                const uint32_t q = v / 13u;
                const uint32_t rank = v - q * 13u;
                v = q;
                const uint32_t suit = (w >> (2 * j)) & 3u;
                const uint32_t bit = 1u << (((suit & 1u) << 4) | rank);
                const uint32_t word = (suit & 2u) ? used_hi : used_lo;
                if (!(word & bit) && got < need) {
                    if (suit & 2u) used_hi |= bit; else used_lo |= bit;
                    out.put(got, (suit << 4) | rank);
                    got++;
                }
            }
        }
    }
}

// one Monte-Carlo rollout (env/workflow.py:130-148 generalised to n_opp opponents):
// returns 0 win, 1 tie, 2 loss for the hero.  `cards` is scratch for the dealt positions.
template <class Tab, class Cards>
PHE_HD int mc_rollout(const Tab& tab, const HashParams& hp, uint64_t known, uint64_t hero, int nboard, int n_opp,
                      uint64_t seed, uint64_t sample, uint32_t stream, Cards& cards)
{
    const int need = (5 - nboard) + 2 * n_opp;
    deal_cards(known, seed, sample, stream, need, cards);
    uint64_t board = known & ~hero;
    int k = 0;
    for (; k < 5 - nboard; k++) board |= pos_bit(cards.get(k));
    const uint32_t hr = eval7(board | hero, tab, hp);
    uint32_t best = 0xFFFFu;
    for (int o = 0; o < n_opp; o++) {
        const uint64_t hole = pos_bit(cards.get(k)) | pos_bit(cards.get(k + 1));
        k += 2;
        const uint32_t r = eval7(board | hole, tab, hp);
        best = r < best ? r : best;
    }
    return hr < best ? 0 : (hr == best ? 1 : 2);
}

// ---- combinations ---------------------------------------------------------------------------
// binom[n*8 + k] = C(n, k), n <= 52, k <= 7
struct Binom {
    const uint32_t* t;
    PHE_HD uint32_t operator()(int n, int k) const { return (k < 0 || n < k) ? 0u : t[n * 8 + k]; }
};

// colex unrank of a 7-subset of {0..51}: idx = sum_i C(c_i, i+1), c_0 < ... < c_6
PHE_HD void colex_unrank7(uint64_t idx, const Binom& C, int c[7])
{
    uint32_t rem = (uint32_t)idx;
PHE_UNROLL
    for (int i = 6; i >= 0; i--) {
        int v = i;
        while (v + 1 <= 52 && C(v + 1, i + 1) <= rem) v++;
        c[i] = v;
        rem -= C(v, i + 1);
    }
}

// lexicographic unrank of a k-subset (k <= 7) of {0..n-1}: first element slowest, increasing
PHE_HD void lex_unrank(uint64_t idx, int n, int k, const Binom& C, int out[7])
{
    int start = 0;
    for (int i = 0; i < k; i++) {
        int v = start;
        // number of subsets whose i-th element is v: C(n-1-v, k-1-i)
        for (;;) {
            const uint64_t cnt = C(n - 1 - v, k - 1 - i);
            if (idx < cnt) break;
            idx -= cnt; v++;
        }
        out[i] = v;
        start = v + 1;
    }
}

// ---- exact plan (simulate_processes) ------------------------------------------------------
// A plan is a list of enumeration segments (unordered combinations in increasing card id,
// env/workflow.py:163-180) followed by an optional random step (:181-195).
struct PlanDesc {
    int n_exist;            // known cards (2 hero + board)
    int n_seg;              // enumeration segments
    int seg_k[7];           // cards per segment
    uint32_t seg_count[7];  // combinations per segment = C(52 - n_exist - cards before, seg_k)
    int n_enum;             // sum seg_k
    int rnd_k;              // cards in the random step (0 = none)
    int rnd_n;              // draws per enumeration leaf
    uint64_t n_leaves;      // product of seg_count
    // exhaustive plans whose last segment is the villain's two cards: the leaves of the segments before it
    // ("prefix": board completions) are walked by warps, the C(.,2) villain hands by lanes over a pair table
    int pair_tail;          // 1 if the fast path applies
    uint64_t n_prefix;      // product of seg_count[0 .. n_seg-2]
};

constexpr uint32_t N_PAIRS = 1326;   // C(52,2): pair index q = C(c1,2) + c0, c0 < c1

// exist: the task's known ids (deal order).  Fills cards[n_exist .. n_exist+n_enum) with the leaf's
// enumerated ids in visiting order.  leaf numbering = reference visiting order.
PHE_HD void plan_unrank_leaf(const PlanDesc& pd, uint64_t leaf, const Binom& C, uint8_t cards[9])
{
    uint64_t usedids = 0;   // bit = card id
    for (int i = 0; i < pd.n_exist; i++) usedids |= 1ull << cards[i];
    // mixed radix, last segment fastest
    uint32_t segidx[7];
    for (int s = pd.n_seg - 1; s >= 0; s--) { segidx[s] = (uint32_t)(leaf % pd.seg_count[s]); leaf /= pd.seg_count[s]; }
    int nc = pd.n_exist;
    for (int s = 0; s < pd.n_seg; s++) {
        const int navail = 52 - nc;
        int sel[7];
        lex_unrank(segidx[s], navail, pd.seg_k[s], C, sel);
        // map "j-th free id" -> id, increasing
        int j = 0, t = 0;
        uint64_t add = 0;
        for (int id = 0; id < 52 && t < pd.seg_k[s]; id++) {
            if (usedids >> id & 1) continue;
            if (j == sel[t]) { cards[nc + t] = (uint8_t)id; add |= 1ull << id; t++; }
            j++;
        }
        usedids |= add;
        nc += pd.seg_k[s];
    }
}

// evaluates one heads-up deal given the 9 ids in the layout of env/workflow.py:131-135
template <class Tab>
PHE_HD int plan_leaf_result(const Tab& tab, const HashParams& hp, const uint8_t cards[9])
{
    const uint64_t board = id_bit(cards[2]) | id_bit(cards[3]) | id_bit(cards[4]) | id_bit(cards[5]) | id_bit(cards[6]);
    const uint32_t hr = eval7(board | id_bit(cards[0]) | id_bit(cards[1]), tab, hp);
    const uint32_t vr = eval7(board | id_bit(cards[7]) | id_bit(cards[8]), tab, hp);
    return hr < vr ? 0 : (hr == vr ? 1 : 2);
}

}  // namespace phe
