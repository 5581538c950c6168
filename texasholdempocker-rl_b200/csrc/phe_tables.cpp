// phe_tables.cpp -- builds the lookup-table image the kernels stage into shared memory.
//
// Product code (not the oracle): ranks are computed analytically from the canonical
// 7462-class numbering that phevaluator.evaluate_cards returns (reference call site
// env/game.py:675; category blocks SF 1-10, quads 11-166, full house 167-322, flush 323-1599,
// straight 1600-1609, trips 1610-2467, two pair 2468-3325, pair 3326-6185, high card 6186-7462).
// tests/ compare the result with the independent sort-based oracle on every table entry.
//
// Image layout (uint16): NOFLUSH[65536] | DISP[8192] | FLUSH[8192]  (see phe_core.cuh).
//   NOFLUSH is addressed by a two-level perfect hash (hash-and-displace) of the bit-sliced rank
//   multiplicities (b0, b1, b2) the evaluator derives from a card-set word.
#include "phe_tables.h"
#include "phe_enum.cuh"

#include <algorithm>
#include <cstring>
#include <vector>

namespace phe {

static uint32_t binom_u32(int n, int k)
{
    if (k < 0 || n < k) return 0;
    uint64_t r = 1;
    for (int i = 1; i <= k; i++) r = r * (uint64_t)(n - k + i) / (uint64_t)i;
    return (uint32_t)r;
}

void build_binom(uint32_t out[53 * 8])
{
    for (int n = 0; n <= 52; n++)
        for (int k = 0; k < 8; k++) out[n * 8 + k] = binom_u32(n, k);
}

// highest card of a straight contained in rank set m (wheel = 3), or -1
static int straight_high(uint32_t m)
{
    for (int h = 12; h >= 4; h--)
        if (((m >> (h - 4)) & 0x1F) == 0x1F) return h;
    if ((m & 0x100F) == 0x100F) return 3;
    return -1;
}

// keep the `k` highest set bits of m
static uint32_t top_bits(uint32_t m, int k)
{
    uint32_t out = 0;
    for (int r = 12; r >= 0 && k > 0; r--)
        if (m >> r & 1) { out |= 1u << r; k--; }
    return out;
}

// strongest-first index of a k-subset (as a bit set over n ranks) among all C(n,k) subsets,
// ordered by comparing the highest element first
static uint32_t subset_index_desc(uint32_t m, int n, int k)
{
    uint32_t colex = 0;
    int i = 0;
    for (int r = 0; r < n; r++)
        if (m >> r & 1) { i++; colex += binom_u32(r, i); }
    return binom_u32(n, k) - 1 - colex;
}

// remove bit positions listed in `drop` (a bit set) and close the gaps
static uint32_t squeeze(uint32_t m, uint32_t drop)
{
    uint32_t out = 0; int o = 0;
    for (int r = 0; r < 13; r++) {
        if (drop >> r & 1) continue;
        if (m >> r & 1) out |= 1u << o;
        o++;
    }
    return out;
}

static int highest_bit(uint32_t m) { return m ? 31 - __builtin_clz(m) : -1; }

struct FiveIndex {
    // for 5-bit rank sets: index among the 1277 non-straight sets, strongest first
    uint16_t idx[8192];
    FiveIndex()
    {
        std::memset(idx, 0xFF, sizeof idx);
        std::vector<uint32_t> sets;
        for (uint32_t m = 0; m < 8192; m++)
            if (__builtin_popcount(m) == 5) sets.push_back(m);
        std::sort(sets.begin(), sets.end(), [](uint32_t a, uint32_t b) { return a > b; });   // numeric desc == highest card first
        uint16_t run = 0;
        for (uint32_t m : sets)
            if (straight_high(m) < 0) idx[m] = run++;
    }
};

static const FiveIndex& five_index()
{
    static const FiveIndex f;
    return f;
}

uint16_t flush_rank(uint32_t lane)
{
    const int h = straight_high(lane);
    if (h >= 0) return (uint16_t)(1 + (12 - h));                       // straight flush
    return (uint16_t)(323 + five_index().idx[top_bits(lane, 5)]);      // plain flush
}

// rank of a non-flush 7-card hand from its rank multiplicities
uint16_t noflush_rank(const uint8_t cnt[13])
{
    uint32_t any = 0, ge2 = 0, ge3 = 0, eq4 = 0;
    for (int r = 0; r < 13; r++) {
        if (cnt[r] >= 1) any |= 1u << r;
        if (cnt[r] >= 2) ge2 |= 1u << r;
        if (cnt[r] >= 3) ge3 |= 1u << r;
        if (cnt[r] == 4) eq4 |= 1u << r;
    }
    if (eq4) {
        const int q = highest_bit(eq4);
        const uint32_t rest = any & ~(1u << q);
        const uint32_t kick = squeeze(1u << highest_bit(rest), 1u << q);
        return (uint16_t)(11 + (12 - q) * 12 + (11 - highest_bit(kick)));
    }
    if (ge3) {
        const int t = highest_bit(ge3);
        const uint32_t pairs = ge2 & ~(1u << t);
        if (pairs) {
            const uint32_t p = squeeze(1u << highest_bit(pairs), 1u << t);
            return (uint16_t)(167 + (12 - t) * 12 + (11 - highest_bit(p)));
        }
    }
    const int sh = straight_high(any);
    if (sh >= 0) return (uint16_t)(1600 + (12 - sh));
    if (ge3) {
        const int t = highest_bit(ge3);
        const uint32_t kick = squeeze(top_bits(any & ~(1u << t), 2), 1u << t);
        return (uint16_t)(1610 + (12 - t) * 66 + subset_index_desc(kick, 12, 2));
    }
    if (__builtin_popcount(ge2) >= 2) {
        const uint32_t two = top_bits(ge2, 2);
        const uint32_t kick = squeeze(1u << highest_bit(any & ~two), two);
        return (uint16_t)(2468 + subset_index_desc(two, 13, 2) * 11 + (10 - highest_bit(kick)));
    }
    if (ge2) {
        const int p = highest_bit(ge2);
        const uint32_t kick = squeeze(top_bits(any & ~(1u << p), 3), 1u << p);
        return (uint16_t)(3326 + (12 - p) * 220 + subset_index_desc(kick, 12, 3));
    }
    return (uint16_t)(6186 + five_index().idx[top_bits(any, 5)]);
}

static void enum_counts(uint8_t cnt[13], int pos, int left, std::vector<NoflushKey>& out)
{
    if (pos == 13) {
        if (left) return;
        NoflushKey k; k.b0 = k.b1 = k.b2 = 0;
        for (int r = 0; r < 13; r++) {
            if (cnt[r] & 1) k.b0 |= 1u << r;
            if (cnt[r] & 2) k.b1 |= 1u << r;
            if (cnt[r] & 4) k.b2 |= 1u << r;
        }
        k.rank = noflush_rank(cnt);
        out.push_back(k);
        return;
    }
    for (int d = 0; d <= 4 && d <= left; d++) { cnt[pos] = (uint8_t)d; enum_counts(cnt, pos + 1, left - d, out); }
    cnt[pos] = 0;
}

std::vector<NoflushKey> all_noflush_keys()
{
    std::vector<NoflushKey> v;
    uint8_t cnt[13] = {0};
    enum_counts(cnt, 0, 7, v);
    return v;
}

static inline uint32_t key_of(const NoflushKey& k) { return (k.b0 & 0xFFFFu) | (k.b1 << 16); }

// hash-and-displace: bucket = umulhi(h1, HASH_BUCKETS); slot = (h2 + DISP[bucket] * HASH_DISP_MUL) >> 16
static bool try_build(const std::vector<NoflushKey>& keys, const HashParams& hp, uint16_t* image)
{
    const size_t n = keys.size();
    std::vector<uint32_t> bucket(n), h2(n);
    for (size_t i = 0; i < n; i++) {
        const uint32_t key = key_of(keys[i]);
        bucket[i] = umulhi32(key * hp.a1 + keys[i].b2 * hp.b1, HASH_BUCKETS);
        h2[i] = key * hp.a2 + keys[i].b2 * hp.b2;
    }
    // the pair (bucket, h2) must identify a key
    {
        std::vector<uint64_t> id(n);
        for (size_t i = 0; i < n; i++) id[i] = ((uint64_t)bucket[i] << 32) | h2[i];
        std::sort(id.begin(), id.end());
        if (std::adjacent_find(id.begin(), id.end()) != id.end()) return false;
    }
    std::vector<std::vector<uint32_t>> members(DISP_SIZE);
    for (size_t i = 0; i < n; i++) members[bucket[i]].push_back((uint32_t)i);
    std::vector<uint32_t> order(DISP_SIZE);
    for (uint32_t b = 0; b < DISP_SIZE; b++) order[b] = b;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return members[x].size() > members[y].size(); });
    std::vector<uint8_t> taken(NOFLUSH_SIZE, 0);
    uint16_t* noflush = image;
    uint16_t* disp = image + OFF_DISP;
    std::memset(noflush, 0, NOFLUSH_SIZE * sizeof(uint16_t));
    std::memset(disp, 0, DISP_SIZE * sizeof(uint16_t));
    std::vector<uint32_t> slots;
    for (uint32_t b : order) {
        const auto& mb = members[b];
        if (mb.empty()) break;
        bool placed = false;
        for (uint32_t d = 0; d < 65536u && !placed; d++) {
            slots.clear();
            bool ok = true;
            for (uint32_t i : mb) {
                const uint32_t s = (h2[i] + d * HASH_DISP_MUL) >> 16;
                if (taken[s]) { ok = false; break; }
                slots.push_back(s);
            }
            if (!ok) continue;
            std::sort(slots.begin(), slots.end());
            if (std::adjacent_find(slots.begin(), slots.end()) != slots.end()) continue;
            for (uint32_t i : mb) {
                const uint32_t s = (h2[i] + d * HASH_DISP_MUL) >> 16;
                taken[s] = 1;
                noflush[s] = keys[i].rank;
            }
            disp[b] = (uint16_t)d;
            placed = true;
        }
        if (!placed) return false;
    }
    return true;
}

bool build_table_image(uint16_t* image, HashParams* hp_out)
{
    const std::vector<NoflushKey> keys = all_noflush_keys();
    if (keys.size() != 49205) return false;
    uint16_t* flush = image + OFF_FLUSH;
    for (uint32_t m = 0; m < FLUSH_SIZE; m++) {
        const int pc = __builtin_popcount(m);
        flush[m] = (pc >= 5 && pc <= 7) ? flush_rank(m) : (uint16_t)0;
    }
    // deterministic multiplier search (first candidate normally succeeds)
    uint64_t s = 0x9E3779B97F4A7C15ull;
    for (int attempt = 0; attempt < 64; attempt++) {
        uint32_t v[4];
        for (int i = 0; i < 4; i++) {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            v[i] = (uint32_t)(s >> 32) | 1u;
        }
        HashParams hp{v[0], v[1], v[2], v[3], 2u, 1u};
        if (try_build(keys, hp, image)) { *hp_out = hp; return true; }
    }
    return false;
}

void build_dealtab(uint64_t out[DEALTAB_ENTRIES])
{
    for (int c1 = 1; c1 < 52; c1++)
        for (int c0 = 0; c0 < c1; c0++) out[c1 * (c1 - 1) / 2 + c0] = id_bit(c0) | id_bit(c1);
    for (int c = 0; c < 52; c++) out[N_PAIRS + c] = id_bit(c);
}

// ---- enumeration image (phe_enum.cuh) --------------------------------------------------------
static void fill_msrank(uint8_t cnt[13], int rank_from, int placed, uint32_t idx, uint16_t* ms)
{
    // places cards in non-decreasing rank order; idx accumulates C(r + i - 1, i) for the i-th card
    if (placed == 7) {
        bool ok = true;
        for (int r = 0; r < 13; r++) ok = ok && cnt[r] <= 4;
        ms[idx] = ok ? noflush_rank(cnt) : (uint16_t)0;
        return;
    }
    for (int r = rank_from; r < 13; r++) {
        cnt[r]++;
        fill_msrank(cnt, r, placed + 1, idx + binom_u32(r + placed, placed + 1), ms);
        cnt[r]--;
    }
}

bool build_enum_image(unsigned char* image, uint64_t* trimask)
{
    std::memset(image, 0, ENUM_IMG_BYTES);
    uint16_t* ms = reinterpret_cast<uint16_t*>(image + EOFF_MSRANK);
    uint8_t cnt[13] = {0};
    fill_msrank(cnt, 0, 0, 0, ms);
    uint16_t* fl = reinterpret_cast<uint16_t*>(image + EOFF_FLUSH);
    for (uint32_t m = 0; m < FLUSH_SIZE; m++) {
        const int pc = __builtin_popcount(m);
        fl[m] = (pc >= 5 && pc <= 7) ? flush_rank(m) : (uint16_t)0;
    }
    uint32_t* tri = reinterpret_cast<uint32_t*>(image + EOFF_TRI);
    for (int c2 = 2; c2 < 49; c2++)
        for (int c1 = 1; c1 < c2; c1++)
            for (int c0 = 0; c0 < c1; c0++) {
                const uint32_t t = binom_u32(c2, 3) + binom_u32(c1, 2) + (uint32_t)c0;
                const uint32_t r0 = c0 >> 2, r1 = c1 >> 2, r2 = c2 >> 2;
                const uint32_t part = binom_u32(r0, 1) + binom_u32(r1 + 1, 2) + binom_u32(r2 + 2, 3);
                const uint32_t suits = (1u << (4 * (c0 & 3))) + (1u << (4 * (c1 & 3))) + (1u << (4 * (c2 & 3)));
                tri[t] = ((2u * part) << 16) | suits;
                trimask[t] = id_bit(c0) | id_bit(c1) | id_bit(c2);
            }
    return true;
}

void build_enum_blocks(std::vector<EnumBlockRec>& out)
{
    uint32_t binom[53 * 8];
    build_binom(binom);
    const Binom C{binom};
    out.clear();
    out.reserve(N_ENUM_BLOCKS);
    for (int c6 = 6; c6 < 52; c6++)
        for (int c5 = 5; c5 < c6; c5++)
            for (int c4 = 4; c4 < c5; c4++)
                for (int c3 = 3; c3 < c4; c3++) {
                    EnumBlock b;
                    enum_block_setup(c3, c4, c5, c6, C, b);
                    out.push_back(EnumBlockRec{(uint32_t)b.s3, (uint32_t)(b.s3 >> 32), b.bidx3 | b.ntriples << 16, b.kbias});
                }
}

void build_enum_units(std::vector<uint32_t>& out, std::vector<uint64_t>& starts, uint32_t max_hands, uint32_t max_blocks, uint32_t align)
{
    if (align == 0) align = 1;
    out.clear();
    out.push_back(0);
    uint32_t unit_start = 0, blocks_in = 0, pos = 0;
    for (int c6 = 6; c6 < 52; c6++)
        for (int c5 = 5; c5 < c6; c5++)
            for (int c4 = 4; c4 < c5; c4++)
                for (int c3 = 3; c3 < c4; c3++) {
                    const uint32_t block_end = pos + binom_u32(c3, 3);
                    const uint32_t block_start = pos;
                    for (;;) {
                        uint32_t room = max_hands - (pos - unit_start);
                        if (block_end - pos >= room) {            // the unit fills up inside (or exactly at the end of) this block
                            // cut on a multiple of `align` hands from the block's start when that still advances: a tile that
                            // resumes the block there starts on a whole trip of the kernel's 4 x 32-hand loop, and only the
                            // block's last segment is left with a remainder of single trips
                            const uint32_t off = pos + room - block_start, aligned = off - off % align;
                            if (pos + room < block_end) {
                                if (aligned > pos - block_start) room = block_start + aligned - pos;
                                else if (pos > unit_start) {      // no aligned cut within reach: close the unit in front of the block
                                    out.push_back(pos);
                                    unit_start = pos;
                                    blocks_in = 0;
                                    continue;
                                }
                            }
                            pos += room;
                            out.push_back(pos);
                            unit_start = pos;
                            blocks_in = 0;
                            if (pos == block_end) break;
                        } else {                                  // the block ends inside the unit
                            pos = block_end;
                            if (++blocks_in >= max_blocks) {
                                out.push_back(pos);
                                unit_start = pos;
                                blocks_in = 0;
                            }
                            break;
                        }
                    }
                }
    if (out.back() != pos) out.push_back(pos);
    // second walk: the block and the triple offset at which every unit starts
    starts.assign(out.size(), 0);
    size_t u = 0;
    uint32_t block_start = 0;
    uint64_t block_id = 0;
    for (int c6 = 6; c6 < 52; c6++)
        for (int c5 = 5; c5 < c6; c5++)
            for (int c4 = 4; c4 < c5; c4++)
                for (int c3 = 3; c3 < c4; c3++, block_id++) {
                    const uint32_t block_end = block_start + binom_u32(c3, 3);
                    for (; u + 1 < out.size() && out[u] < block_end; u++) starts[u] = block_id | (uint64_t)(out[u] - block_start) << 18;
                    block_start = block_end;
                }
}

}  // namespace phe
