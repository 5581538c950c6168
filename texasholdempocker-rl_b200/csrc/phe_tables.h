// phe_tables.h -- host-side construction of the evaluator's lookup-table image.
#pragma once
#include <stdint.h>

#include <vector>

#include "phe_core.cuh"
#include "phe_enum.cuh"

namespace phe {

struct NoflushKey { uint32_t b0, b1, b2; uint16_t rank; };

// C(n,k) for n <= 52, k <= 7, row-major [53][8]
void build_binom(uint32_t out[53 * 8]);

// rank of the best flush / straight flush inside one 13-bit suit lane with 5..7 bits set
uint16_t flush_rank(uint32_t lane);
// rank of a 7-card hand without a flush, from its 13 rank multiplicities (sum 7, each <= 4)
uint16_t noflush_rank(const uint8_t cnt[13]);
// all 49205 rank multisets with their bit-sliced keys and ranks
std::vector<NoflushKey> all_noflush_keys();
// fills image[TAB_U16] and the hash multipliers; false if no perfect hash was found
bool build_table_image(uint16_t* image, HashParams* hp_out);

// card sets of all C(52,2) two-card hands (index C(c1,2) + c0) followed by the 52 single cards (index 1326 + id)
void build_dealtab(uint64_t out[DEALTAB_ENTRIES]);

// fills image[ENUM_IMG_BYTES] and trimask[N_TRIPLES] (layouts in phe_enum.cuh)
bool build_enum_image(unsigned char* image, uint64_t* trimask);

// records of all N_ENUM_BLOCKS enumeration blocks in colex order (phe_enum.cuh)
void build_enum_blocks(std::vector<EnumBlockRec>& out);

// Work units of the exhaustive enumeration on the colex axis of C(52,7): unit u covers [out[u], out[u+1]).  A unit ends
// after `max_hands` hands or after `max_blocks` enumeration blocks (a block = one suffix (c3..c6), C(c3,3) hands), whichever
// comes first (a cut inside a block falls on a multiple of `align` hands from the block's start) -- runs of tiny blocks (wherever c4 restarts) cost a warp a serial set-up each, so they are cut by count.
// out.front() == 0, out.back() == C(52,7).  starts[u] packs where unit u begins, so that the kernel need not unrank the index:
// block id | t << 18 with t = C(c2,3) + C(c1,2) + c0 the colex index of the triple inside its block
// (starts.size() == out.size(); the entry of the end sentinel is 0).
void build_enum_units(std::vector<uint32_t>& out, std::vector<uint64_t>& starts, uint32_t max_hands, uint32_t max_blocks, uint32_t align = 1);

}  // namespace phe
