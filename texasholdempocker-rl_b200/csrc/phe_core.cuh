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

// multipliers of the two hash functions (found by the table builder, verified collision free).  `two` is the constant 2,
// kept where the compiler cannot see it: table byte addresses are formed as idx * two + base, one IMAD on the FMA pipe
// instead of a LEA / IADD3 on the logic pipe that bounds these kernels.
struct HashParams { uint32_t a1, b1, a2, b2, two, one; };   // `one` = 1, opaque like `two`: see or_disjoint
// Range reductions are multiply-high (umulhi(h, n), Lemire) so that they run on the FMA pipe: the kernels are bound by
// the logic (alu) pipe.  Ranges that are not powers of two keep the compiler from turning them back into shifts.
constexpr uint32_t HASH_BUCKETS = 8191;        // bucket = umulhi(h1, 8191) in [0, 8191)
constexpr uint32_t HASH_DISP_MUL = 0x9E3779B1u; // slot = (h2 + DISP[bucket] * HASH_DISP_MUL) >> 16: the displacement re-seeds
                                                // the second hash, so no wrap-around mask is needed

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

// a | b for words WITHOUT common bits.  With -DPHE_PIPE_BALANCE it becomes one multiply-add on the FMA pipe (b * one + a,
// `one` = HashParams::one, the constant 1 read from constant memory so that the compiler cannot fold it back) instead of a
// LOP3 on the logic pipe.  Measured on the B200 (tools/lab/mc_variants.py, profiles/mc_variants_r02.jsonl): a LOSS in the
// Monte-Carlo kernel (3.65e10 against 3.77e10 rollouts/s) -- the dealing loop is a dependent chain through `used`, and the
// IMAD's latency costs more than the logic-pipe slot it frees -- so it is off by default and kept for the record.
PHE_HD uint32_t or_disjoint(uint32_t a, uint32_t b, uint32_t one)
{
#if defined(__CUDA_ARCH__) && defined(PHE_PIPE_BALANCE)
    return b * one + a;
#else
    (void)one;
    return a | b;
#endif
}

// the same for the two ORs of board | hole cards in front of every evaluation: these are NOT on the dealing loop's dependent
// chain, so the FMA form was judged on its own (-DPHE_FMA_EVAL_OR): 3.990e10 against 4.023e10 rollouts/s, and the cursor add
// alone (-DPHE_FMA_CURSOR) 3.942e10 -- a multiply by a register-held 1 is not a free way off the logic pipe; only the table
// ADDRESSES (idx * two + base, q * eight + base), where the multiply replaces a shift-add, pay.  Both stay off.
PHE_HD uint32_t or_disjoint_eval(uint32_t a, uint32_t b, uint32_t one)
{
#if defined(__CUDA_ARCH__) && (defined(PHE_PIPE_BALANCE) || defined(PHE_FMA_EVAL_OR))
    return b * one + a;
#else
    (void)one;
    return a | b;
#endif
}

// a + b as b * one + a: the same trick for a plain addition (cursor bookkeeping of the dealing loop)
PHE_HD uint32_t add_fma(uint32_t a, uint32_t b, uint32_t one)
{
#if defined(__CUDA_ARCH__) && (defined(PHE_PIPE_BALANCE) || defined(PHE_FMA_CURSOR))
    return b * one + a;
#else
    (void)one;
    return a + b;
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
    PHE_HD uint32_t ld_disp(uint32_t bucket, const HashParams&) const { return ld(OFF_DISP + bucket); }
    // the single final read: FLUSH[lane] for a flush, NOFLUSH[slot] otherwise
    PHE_HD uint32_t ld_rank(bool fl, uint32_t lane, uint32_t slot, const HashParams&) const { return ld(fl ? OFF_FLUSH + lane : slot); }
};

#ifdef __CUDACC__
struct SharedTab {
    uint32_t base;   // shared-window byte address of the table image
    __device__ __forceinline__ uint32_t lds(uint32_t a) const
    {
        uint16_t v;
        asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
        return v;
    }
    __device__ __forceinline__ uint32_t ld(uint32_t idx) const { return lds(base + idx * 2u); }
    __device__ __forceinline__ uint32_t ld_disp(uint32_t bucket, const HashParams& hp) const { return lds(bucket * hp.two + (base + 2u * OFF_DISP)); }
    __device__ __forceinline__ uint32_t ld_rank(bool fl, uint32_t lane, uint32_t slot, const HashParams& hp) const
    {
        const uint32_t af = lane * hp.two + (base + 2u * OFF_FLUSH), an = slot * hp.two + base;
        return lds(fl ? af : an);
    }
};
#endif

// (p ^ q) & 0xFFFF, p ^ q ^ r and majority(p, q, r) as single LOP3s: left to itself the compiler re-associates the
// bit-sliced adder below into 10 logic ops instead of 8
PHE_HD uint32_t xor_low16(uint32_t p, uint32_t q)
{
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, 0xFFFF, 0x28;" : "=r"(r) : "r"(p), "r"(q));
    return r;
#else
    return (p ^ q) & 0xFFFFu;
#endif
}
PHE_HD uint32_t xor3(uint32_t p, uint32_t q, uint32_t r)
{
#ifdef __CUDA_ARCH__
    uint32_t o;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(o) : "r"(p), "r"(q), "r"(r));
    return o;
#else
    return p ^ q ^ r;
#endif
}
PHE_HD uint32_t maj3(uint32_t p, uint32_t q, uint32_t r)
{
#ifdef __CUDA_ARCH__
    uint32_t o;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(o) : "r"(p), "r"(q), "r"(r));
    return o;
#else
    return (p & q) | (p & r) | (q & r);
#endif
}

// ---- 7-card evaluator on a card-set word -----------------------------------------------------
// m must have exactly 7 card bits set.  Branch free: the flush and the non-flush index are both
// formed and one of them selects the single final table read.
// perfect-hash slot of the rank multiset of the card set (lo, hi): the non-flush index
template <class Tab>
PHE_HD uint32_t noflush_slot(uint32_t lo, uint32_t hi, const Tab& tab, const HashParams& hp)
{
    // rank multiplicities, bit sliced: count(r) = b0 + 2*b1 + 4*b2
    const uint32_t x = lo ^ hi, a = lo & hi;
    const uint32_t x1 = x >> 16, a1 = a >> 16;
    const uint32_t b0 = xor_low16(x, x1);
    const uint32_t cy = x & x1;
    const uint32_t b1 = xor3(a, a1, cy);               // upper half is garbage, shifted out below
    const uint32_t b2 = maj3(a, a1, cy);               // a1 and cy have clean upper halves, so has b2
    const uint32_t key = b1 * 65536u + b0;
    const uint32_t h1 = key * hp.a1 + b2 * hp.b1;
    const uint32_t h2 = key * hp.a2 + b2 * hp.b2;
    const uint32_t d = tab.ld_disp(umulhi32(h1, HASH_BUCKETS), hp);
    return (h2 + d * HASH_DISP_MUL) >> 16;
}

template <class Tab>
PHE_HD uint32_t eval7(uint64_t m, const Tab& tab, const HashParams& hp)
{
    const uint32_t lo = (uint32_t)m, hi = (uint32_t)(m >> 32);   // lanes {D,C} and {S,H}
    const uint32_t inf = noflush_slot(lo, hi, tab, hp);
    // flush: with 7 cards at most one suit lane holds >= 5
    const int plo = popc32(lo);
    const uint32_t w = plo >= 5 ? lo : hi;
    const uint32_t wl = w & 0xFFFFu;
    const int pl = popc32(wl), pw = popc32(w);
    const uint32_t lane = pl >= 5 ? wl : (w >> 16);
    const bool fl = (pl >= 5) | (pw - pl >= 5);
    return tab.ld_rank(fl, lane, inf, hp);
}

// ---- several hands on ONE board (a Monte-Carlo rollout, a showdown) ---------------------------------------------------
// A flush needs three board cards of its suit, and five board cards hold at most one such suit: it is found once per
// board; per hand the test is then "board bits of that suit | the hand's bits of that suit has >= 5 bits" -- select,
// shift, and-or, one POPC -- instead of the three POPCs and the selects of the general form.
struct BoardFlush {
    uint32_t lane;   // the board's rank bits in the candidate suit (0 if no suit holds three)
    uint32_t sh;     // 0 / 16: position of that suit lane in its word
    uint32_t pm;     // 0xFFFF if a candidate suit exists, else 0 (no hand can make a flush)
    bool hi;         // candidate suit lives in the high word
};

PHE_HD BoardFlush board_flush(uint32_t blo, uint32_t bhi)
{
    const int c0 = popc32(blo & 0xFFFFu), c1 = popc32(blo) - c0, c2 = popc32(bhi & 0xFFFFu), c3 = popc32(bhi) - c2;
    BoardFlush f;
    f.hi = (c2 >= 3) | (c3 >= 3);
    const bool up = f.hi ? (c3 >= 3) : (c1 >= 3);
    const bool any = f.hi | (c0 >= 3) | (c1 >= 3);
    f.sh = up ? 16u : 0u;
    f.pm = any ? 0xFFFFu : 0u;
    f.lane = ((f.hi ? bhi : blo) >> f.sh) & f.pm;
    return f;
}

// rank of board (5 cards: blo, bhi) + hole (2 cards, disjoint from the board)
template <class Tab>
PHE_HD uint32_t eval7_board(uint32_t blo, uint32_t bhi, const BoardFlush& bf, uint64_t hole, const Tab& tab, const HashParams& hp)
{
    const uint32_t plo = (uint32_t)hole, phi = (uint32_t)(hole >> 32);
    const uint32_t inf = noflush_slot(or_disjoint_eval(blo, plo, hp.one), or_disjoint_eval(bhi, phi, hp.one), tab, hp);
    const uint32_t lane = (((bf.hi ? phi : plo) >> bf.sh) & bf.pm) | bf.lane;
    return tab.ld_rank(popc32(lane) >= 5, lane, inf, hp);
}

PHE_HD int category_of(uint32_t rank)
{
    // last rank of SF, quads, full house, flush, straight, trips, two pair, pair
    return (rank > 10) + (rank > 166) + (rank > 322) + (rank > 1599) + (rank > 1609) + (rank > 2467) +
           (rank > 3325) + (rank > 6185);
}

// ---- Philox4x32-10 ---------------------------------------------------------------------------
// The key of round r is (k0 + r * 0x9E3779B9, k1 + r * 0xBB67AE85).  The key is the seed, a launch constant, so the hot
// kernels take the twenty round keys as a kernel parameter (RoundKeys: constant-bank operands of the XORs) instead of
// re-deriving them with 18 additions in every block; SeedKey derives them on the fly (host code, cold kernels).
struct SeedKey {
    uint32_t lo, hi;
    PHE_HD uint32_t k0(int r) const { return lo + (uint32_t)r * 0x9E3779B9u; }
    PHE_HD uint32_t k1(int r) const { return hi + (uint32_t)r * 0xBB67AE85u; }
};
struct RoundKeys {
    uint32_t k[20];
    PHE_HD uint32_t k0(int r) const { return k[2 * r]; }
    PHE_HD uint32_t k1(int r) const { return k[2 * r + 1]; }
};
PHE_HD SeedKey seed_key(uint64_t seed) { return SeedKey{(uint32_t)seed, (uint32_t)(seed >> 32)}; }
inline RoundKeys round_keys(uint64_t seed)
{
    RoundKeys rk;
    const SeedKey sk = seed_key(seed);
    for (int r = 0; r < 10; r++) { rk.k[2 * r] = sk.k0(r); rk.k[2 * r + 1] = sk.k1(r); }
    return rk;
}

template <class Key>
PHE_HD void philox4x32_10_k(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const Key& key, uint32_t out[4])
{
PHE_UNROLL
    for (int r = 0; r < 10; r++) {
        const uint32_t h0 = umulhi32(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = umulhi32(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ key.k0(r); c1 = l1;
        c2 = h0 ^ c3 ^ key.k1(r); c3 = l0;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

PHE_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                          uint32_t out[4])
{
    philox4x32_10_k(c0, c1, c2, c3, SeedKey{k0, k1}, out);
}

// ---- dealing -----------------------------------------------------------------------------------
// A deal is a sequence of SLOTS: at most one single card, then pairs.  Drawing an unordered pair uniformly among the
// free pairs, one pair after another, is the same distribution as dealing the cards one by one -- and a hold'em deal
// only ever needs unordered groups: the missing board cards (as pairs plus one single when their number is odd) and
// two hole cards per opponent.  One pair costs one 16-bit draw and one table read instead of two card draws.
//
// Stream (a pure function of seed, stream id, sample index; shared bit for bit with the oracle):
//   Philox block b = philox4x32_10(counter (sample lo, sample hi, stream, b), key seed); its four words are cut into
//   eight 16-bit draws, low half first.  DEALTAB[0..1326) holds the card sets of all pairs (index C(c1,2)+c0, c0<c1),
//   DEALTAB[1326..1378) those of the 52 single cards.
//   * single slot (only if asked for): draws 0 and 1 of block 0 are reserved for it; a draw r is valid if r < 65520
//     (= 1260*52) and proposes card r % 52; the first valid proposal whose card is free is taken.  If neither works
//     (about 1 %) further draws come from blocks 0x80000000, 0x80000001, ...
//   * pair slots, in order: draws continue after the reserved word (or from draw 0 if there is no single); r is valid
//     if r < 64974 (= 49*1326) and proposes pair r % 1326, accepted iff both cards are free.
//   Every accepted slot marks its cards used.  Slot roles are the caller's: hold'em uses single + first pairs = board
//   completion, remaining pairs = opponents in order.
constexpr int MAX_DEAL = 21;                 // cards: 5 board + 2 * 8 opponents
constexpr int MAX_PAIR_SLOTS = 10;           // 2 board pairs + 8 opponents
constexpr uint32_t N_PAIRS = 1326;           // C(52,2): pair index q = C(c1,2) + c0, c0 < c1
constexpr uint32_t DEALTAB_ENTRIES = N_PAIRS + 52;
constexpr uint32_t DEAL_PAIR_LIMIT = 49u * 1326u;    // 64,974
constexpr uint32_t DEAL_SINGLE_LIMIT = 1260u * 52u;  // 65,520

// r % 1326 and r % 52 for r < 65536 by multiply-shift (exact for every 16-bit r; checked exhaustively in the tests)
PHE_HD uint32_t mod1326_u16(uint32_t r) { return r - umulhi32(r, 3239054u) * 1326u; }   // floor(r / 1326) = (r * 3239054) >> 32 for r < 2^16: two FMA-pipe instructions
PHE_HD uint32_t mod52_u16(uint32_t r) { return r - ((r * 40330u) >> 21) * 52u; }

struct GlobalDealTab {          // host / small kernels
    const uint64_t* t;
    PHE_HD uint64_t mask(uint32_t q) const { return t[q]; }
};

struct LocalSlots {             // host / small kernels: accepted slots kept in thread-local storage
    uint64_t single;
    uint64_t pair[MAX_PAIR_SLOTS + 1];      // one spare row: see deal_slots
    PHE_HD void put_single(uint64_t m) { single = m; }
    PHE_HD void put_pair(int k, uint32_t, uint64_t m) { pair[k] = m; }
    PHE_HD uint64_t get_single() const { return single; }
    template <class DealTab>
    PHE_HD uint64_t get_pair(int k, const DealTab&) const { return pair[k]; }
    // cursor over the pair slots: an index here, a shared-memory address in the kernels' slot store
    typedef int cursor;
    PHE_HD cursor cursor_at(int k) const { return k; }
    PHE_HD cursor cursor_next(cursor c, uint32_t) const { return c + 1; }
    PHE_HD int cursor_index(cursor c) const { return c; }
    PHE_HD void put_at(cursor c, uint32_t q, uint64_t m) { put_pair(c, q, m); }
};

// returns the number of pair slots accepted (== n_pairs unless the input was malformed)
template <class DealTab, class Slots, class Key>
PHE_HD int deal_slots_k(const DealTab& dt, uint64_t known, const Key& key, uint64_t sample, uint32_t stream, bool want_single, int n_pairs,
                        Slots& out, uint32_t one = 1u)
{
    uint32_t used_lo = (uint32_t)known, used_hi = (uint32_t)(known >> 32);
    const uint32_t s_lo = (uint32_t)sample, s_hi = (uint32_t)(sample >> 32);
    uint32_t rnd[4];
    philox4x32_10_k(s_lo, s_hi, stream, 0u, key, rnd);
    if (want_single) {
        bool have = false;
PHE_UNROLL
        for (int h = 0; h < 2; h++) {
            const uint32_t r = (rnd[0] >> (16 * h)) & 0xFFFFu;
            const uint64_t m = dt.mask(N_PAIRS + mod52_u16(r));
            if (!have && r < DEAL_SINGLE_LIMIT && !(((uint32_t)m & used_lo) | ((uint32_t)(m >> 32) & used_hi))) {
                have = true; used_lo |= (uint32_t)m; used_hi |= (uint32_t)(m >> 32); out.put_single(m);
            }
        }
        for (uint32_t t = 0; !have && t < 64u; t++) {      // rare overflow path
            uint32_t ex[4];
            philox4x32_10_k(s_lo, s_hi, stream, 0x80000000u + t, key, ex);
            for (int d = 0; d < 8 && !have; d++) {
                const uint32_t r = (ex[d >> 1] >> (16 * (d & 1))) & 0xFFFFu;
                const uint64_t m = dt.mask(N_PAIRS + mod52_u16(r));
                if (r < DEAL_SINGLE_LIMIT && !(((uint32_t)m & used_lo) | ((uint32_t)(m >> 32) & used_hi))) {
                    have = true; used_lo |= (uint32_t)m; used_hi |= (uint32_t)(m >> 32); out.put_single(m);
                }
            }
        }
    }
    // the number of accepted pairs is kept as a cursor into the slot store (an address in the kernels): the accept path
    // is one store and one add, with no index-to-address multiply.  The accept test does not look at the cursor: the word
    // loop below stops as soon as the slots are full, so at most ONE further pair (second half of the last word) can be
    // accepted after that -- it lands in the spare row n_pairs of the slot store and is never read.  The marks of an
    // accepted pair are added, not OR-ed (its cards are free, no bit is shared): a multiply-add on the FMA pipe.
    typename Slots::cursor cur = out.cursor_at(0);
    const typename Slots::cursor end = out.cursor_at(n_pairs);
    int first_word = want_single ? 1 : 0;
    // 64 blocks = 512 draws: never reached with a valid state; bounds the loop on garbage
    for (uint32_t blk = 0; cur < end && blk < 64u; blk++) {
        if (blk) philox4x32_10_k(s_lo, s_hi, stream, blk, key, rnd);
PHE_UNROLL
        for (int wi = 0; wi < 4; wi++) {
            if (wi < first_word) continue;
            if (!(cur < end)) break;
PHE_UNROLL
            for (int h = 0; h < 2; h++) {
                const uint32_t r = (rnd[wi] >> (16 * h)) & 0xFFFFu;
                const uint32_t q = mod1326_u16(r);
                const uint64_t m = dt.mask(q);
                const uint32_t m_lo = (uint32_t)m, m_hi = (uint32_t)(m >> 32);
                if (r < DEAL_PAIR_LIMIT && !((m_lo & used_lo) | (m_hi & used_hi))) {
                    used_lo = or_disjoint(used_lo, m_lo, one);
                    used_hi = or_disjoint(used_hi, m_hi, one);
                    out.put_at(cur, q, m);
                    cur = out.cursor_next(cur, one);
                }
            }
        }
        first_word = 0;
    }
    const int got = out.cursor_index(cur);
    return got < n_pairs ? got : n_pairs;
}

template <class DealTab, class Slots>
PHE_HD int deal_slots(const DealTab& dt, uint64_t known, uint64_t seed, uint64_t sample, uint32_t stream, bool want_single, int n_pairs,
                      Slots& out, uint32_t one = 1u)
{
    return deal_slots_k(dt, known, seed_key(seed), sample, stream, want_single, n_pairs, out, one);
}

// one Monte-Carlo rollout (env/workflow.py:130-148 generalised to n_opp opponents):
// returns 0 win, 1 tie, 2 loss for the hero.  `slots` is the scratch for the accepted deal slots.
template <class Tab, class DealTab, class Slots, class Key>
PHE_HD int mc_rollout_k(const Tab& tab, const HashParams& hp, const DealTab& dt, uint64_t known, uint64_t hero, int nboard, int n_opp,
                        const Key& key, uint64_t sample, uint32_t stream, Slots& slots)
{
    const int nbm = 5 - nboard, nbp = nbm >> 1;
    deal_slots_k(dt, known, key, sample, stream, (nbm & 1) != 0, nbp + n_opp, slots, hp.one);
    uint64_t board = known & ~hero;
    if (nbm & 1) board |= slots.get_single();
    for (int k = 0; k < nbp; k++) board |= slots.get_pair(k, dt);
    const uint32_t blo = (uint32_t)board, bhi = (uint32_t)(board >> 32);
#ifndef PHE_MC_GENERIC_EVAL
    const BoardFlush bf = board_flush(blo, bhi);        // the board is common to all n_opp + 1 hands of the rollout
    const uint32_t hr = eval7_board(blo, bhi, bf, hero, tab, hp);
    uint32_t best = 0xFFFFu;
    for (int o = 0; o < n_opp; o++) {
        const uint32_t r = eval7_board(blo, bhi, bf, slots.get_pair(nbp + o, dt), tab, hp);
        best = r < best ? r : best;
    }
#else
    const uint32_t hr = eval7(board | hero, tab, hp);
    uint32_t best = 0xFFFFu;
    for (int o = 0; o < n_opp; o++) {
        const uint32_t r = eval7(board | slots.get_pair(nbp + o, dt), tab, hp);
        best = r < best ? r : best;
    }
#endif
    return hr < best ? 0 : (hr == best ? 1 : 2);
}

template <class Tab, class DealTab, class Slots>
PHE_HD int mc_rollout(const Tab& tab, const HashParams& hp, const DealTab& dt, uint64_t known, uint64_t hero, int nboard, int n_opp,
                      uint64_t seed, uint64_t sample, uint32_t stream, Slots& slots)
{
    return mc_rollout_k(tab, hp, dt, known, hero, nboard, n_opp, seed_key(seed), sample, stream, slots);
}

// `need` cards as ids, for phe_deal and the random step of equity plans.  Slot order = single (if need is odd), then
// need/2 pairs; ids come out as [single, pair 0 low, pair 0 high, pair 1 low, ...], so the last two ids are always one
// unordered pair (the villain's hand in a plan) and everything before it is an unordered group (the board completion).
PHE_HD uint32_t lowest_pos(uint64_t m)
{
#ifdef __CUDA_ARCH__
    return m ? (uint32_t)__ffsll((long long)m) - 1u : 0u;
#else
    return m ? (uint32_t)__builtin_ctzll(m) : 0u;
#endif
}

template <class DealTab>
PHE_HD void deal_ids(const DealTab& dt, uint64_t known, uint64_t seed, uint64_t sample, uint32_t stream, int need, uint8_t* ids)
{
    LocalSlots sl;
    sl.single = 0;
    for (int k = 0; k <= MAX_PAIR_SLOTS; k++) sl.pair[k] = 0;
    const int np = need >> 1, odd = need & 1;
    deal_slots(dt, known, seed, sample, stream, odd != 0, np, sl);
    if (odd) ids[0] = (uint8_t)pos_to_id(lowest_pos(sl.single));
    for (int k = 0; k < np; k++) {
        const uint64_t m = sl.pair[k];
        ids[odd + 2 * k] = (uint8_t)pos_to_id(lowest_pos(m));
        ids[odd + 2 * k + 1] = (uint8_t)pos_to_id(lowest_pos(m & (m - 1)));
        if (ids[odd + 2 * k] > ids[odd + 2 * k + 1]) {
            const uint8_t t = ids[odd + 2 * k]; ids[odd + 2 * k] = ids[odd + 2 * k + 1]; ids[odd + 2 * k + 1] = t;
        }
    }
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
