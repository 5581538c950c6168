// phe_enum.cuh -- arithmetic of the exhaustive C(52,7) enumeration kernel (BASELINE config 2).
//
// Hands are visited in colex order of their sorted ids c0 < ... < c6, so ranks arrive sorted and the
// rank multiset of a hand has the dense index  M = sum_{i=1..7} C(r_i + i - 1, i)  (r_1 <= ... <= r_7),
// a bijection onto [0, C(19,7)) = [0, 50388).  A warp works on one (c3..c6) suffix at a time ("block"):
// the block fixes  sum_{i>=4}  and the suit counts of four cards, and the C(c3,3) triples (c0,c1,c2) below it
// are consecutive colex hand indices t = C(c2,3) + C(c1,2) + c0 that the 32 lanes stride over.  Per hand: one
// read of the triple table, one add, one read of MSRANK.  A hand holds a flush iff some suit nibble of
// (block counts + triple counts) reaches 5; only then (3 % of hands) the triple's card set is fetched from
// the L2-resident TRIMASK table and FLUSH[suit lane] replaces the MSRANK read.
//
// Shared-memory image (bytes):  MSRANK u16[50388] | pad | FLUSH u16[8192] | TRI u32[18424]
//   TRI[t] = (2 * multiset index part of the triple) << 16 | per-suit card counts of the triple (4 nibbles)
// Global:  TRIMASK u64[18424] = card set of the triple.
#pragma once
#include "phe_core.cuh"

namespace phe {

constexpr uint32_t MS_ENTRIES = 50388;          // C(19,7)
constexpr uint32_t N_TRIPLES = 18424;           // C(49,3): c2 <= 48
constexpr uint32_t EOFF_MSRANK = 0;
constexpr uint32_t EOFF_FLUSH = 100784;         // 50388*2 rounded up to 16
constexpr uint32_t EOFF_TRI = EOFF_FLUSH + 16384;             // 117,168
constexpr uint32_t ENUM_IMG_BYTES = EOFF_TRI + N_TRIPLES * 4; // 190,864 (multiple of 16)
constexpr uint32_t TRIMASK_BYTES = N_TRIPLES * 8;

constexpr uint32_t N_ENUM_BLOCKS = 211876;      // C(49,4): suffixes 3 <= c3 < c4 < c5 < c6 <= 51, id = colex rank of (c3-3, .., c6-3)

// one enumeration block as the kernel reads it (16 bytes, table in global memory / L2; built from enum_block_setup)
struct EnumBlockRec {
    uint32_t s3lo, s3hi;   // card set of c3..c6
    uint32_t idx_n;        // bidx3 (16 bits: < 50388) | ntriples << 16 (<= C(48,3) = 17296)
    uint32_t kbias;
};

struct EnumBlock {
    uint64_t s3;        // card set of c3..c6
    uint32_t bidx3;     // sum_{i=4..7} C(r_i + i - 1, i)
    uint32_t kbias;     // per-suit card counts of c3..c6 (one nibble per suit) + 3 per nibble
    uint32_t ntriples;  // C(c3, 3)
};

PHE_HD void enum_block_setup(int c3, int c4, int c5, int c6, const Binom& C, EnumBlock& b)
{
    b.s3 = id_bit(c3) | id_bit(c4) | id_bit(c5) | id_bit(c6);
    b.bidx3 = C((c3 >> 2) + 3, 4) + C((c4 >> 2) + 4, 5) + C((c5 >> 2) + 5, 6) + C((c6 >> 2) + 6, 7);
    b.kbias = 0x3333u + (1u << (4 * (c3 & 3))) + (1u << (4 * (c4 & 3))) + (1u << (4 * (c5 & 3))) + (1u << (4 * (c6 & 3)));
    b.ntriples = C(c3, 3);
}

// colex successor of the suffix (c3..c6); c3 runs fastest
PHE_HD void enum_block_next(int& c3, int& c4, int& c5, int& c6)
{
    ++c3;
    if (c3 == c4) {
        c3 = 3; ++c4;
        if (c4 == c5) {
            c4 = 4; ++c5;
            if (c5 == c6) { c5 = 5; ++c6; }
        }
    }
}

// suit (0..3) whose nibble has bit 3 set in x = (kbias + triple counts) & 0x8888; x != 0
PHE_HD uint32_t flush_suit_of(uint32_t x) { return x & 0x0008u ? 0u : (x & 0x0080u ? 1u : (x & 0x0800u ? 2u : 3u)); }

// reference (host / test) form of the per-hand rank; the kernel inlines the same steps
struct EnumImage {
    const unsigned char* p;        // shared-memory image
    const uint64_t* trimask;       // global TRIMASK
    PHE_HD uint32_t msrank_at(uint32_t byte_off) const { return *reinterpret_cast<const uint16_t*>(p + EOFF_MSRANK + byte_off); }
    PHE_HD uint32_t flush(uint32_t lane) const { return reinterpret_cast<const uint16_t*>(p + EOFF_FLUSH)[lane]; }
    PHE_HD uint32_t tri(uint32_t t) const { return reinterpret_cast<const uint32_t*>(p + EOFF_TRI)[t]; }
};

PHE_HD uint32_t enum_hand_rank(const EnumImage& img, const EnumBlock& b, uint32_t t)
{
    const uint32_t e = img.tri(t);
    const uint32_t x = (e + b.kbias) & 0x8888u;
    if (x) {
        const uint32_t s = flush_suit_of(x);
        const uint64_t m = b.s3 | img.trimask[t];
        return img.flush((uint32_t)(m >> (16 * s)) & 0x1FFFu);
    }
    return img.msrank_at(2u * b.bidx3 + (e >> 16));
}

}  // namespace phe
