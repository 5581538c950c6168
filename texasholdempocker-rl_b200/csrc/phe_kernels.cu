// phe_kernels.cu -- sm_100a kernels and the C ABI declared in include/phe_b200.h.
//
// All kernels are integer / shared-memory-lookup work (no tensor cores): card sets are 64-bit
// words, the rank-table images (160 KB generic, 190 KB enumeration) are staged into shared memory
// once per CTA with a TMA bulk copy (cp.async.bulk + mbarrier), grids are persistent (one
// 1024-thread CTA per SM), counts are reduced with redux/shuffles before any atomic, and the
// enumeration's histogram all-reduce is fused into its kernel over peer memory.  There is no host
// compute path: every entry point either launches a kernel or returns an error.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/phe_b200.h"
#include "phe_core.cuh"
#include "phe_enum.cuh"
#include "phe_plan.h"
#include "phe_settle.cuh"
#include "phe_tables.h"

using namespace phe;

// ------------------------------------------------------------------------------------------------
// device-side constants
// ------------------------------------------------------------------------------------------------
__constant__ HashParams c_hp;
__constant__ uint32_t c_binom[53 * 8];

static constexpr int kCtaThreads = 1024;
static constexpr int kHistBins = PHE_NUM_RANKS + 1;                 // 7463, bin = rank
static constexpr int kHistBinsPad = 7464;
static constexpr uint32_t kSmemTab = TAB_BYTES;                       // 163,840
static constexpr uint32_t kEnumHistOff = ENUM_IMG_BYTES;                     // 190,864
static constexpr uint32_t kEnumH9Off = kEnumHistOff + kHistBinsPad * 4;       // 220,720
static constexpr uint32_t kEnumBarOff = kEnumH9Off + 9 * 8;                   // 220,792
static constexpr uint32_t kEnumBinOff = kEnumBarOff + 16;                      // u32[8][64]: C(v, k) at [k*64 + v]
static constexpr uint32_t kSmemEnum = kEnumBinOff + 8 * 64 * 4;
static constexpr uint32_t kSmemTabOnly = kSmemTab + 16;
static constexpr uint32_t kSmemPairs = kSmemTabOnly + N_PAIRS * 8;
static constexpr uint32_t kSmemEvalIds = kSmemTabOnly + (kCtaThreads / 32) * 896;      // + per-warp transpose strips of the id layout
static constexpr uint32_t kDealTabBytes = DEALTAB_ENTRIES * 8;                 // 11,024
#ifndef PHE_MC_THREADS
#define PHE_MC_THREADS 1024
#endif
#ifdef PHE_MC_CLUSTER1
#define PHE_MC_CLUSTER_ATTR __cluster_dims__(1, 1, 1)
#else
#define PHE_MC_CLUSTER_ATTR
#endif
static constexpr int kMcThreads = PHE_MC_THREADS;                      // CTA size of k_equity_mc_smem (one CTA per SM)
static constexpr uint32_t kMcSlotRows = MAX_PAIR_SLOTS + 1;            // + the spare row of deal_slots
static constexpr uint32_t kSmemMc = kSmemTabOnly + kDealTabBytes + kMcSlotRows * 4 * kMcThreads;   // + deal table + slot scratch

// ------------------------------------------------------------------------------------------------
// table staging: one TMA bulk copy of the whole image, completion on an mbarrier
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_image(unsigned char* s_dst, const void* __restrict__ g_src, uint32_t bytes, uint64_t* s_bar)
{
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(s_bar);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_dst);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        constexpr uint32_t kChunk = 32768;
        const unsigned char* src = reinterpret_cast<const unsigned char*>(g_src);
#pragma unroll 1
        for (uint32_t off = 0; off < bytes; off += kChunk) {
            const uint32_t n = bytes - off < kChunk ? bytes - off : kChunk;   // multiples of 16
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + off),
                         "l"(src + off), "r"(n), "r"(bar)
                         : "memory");
        }
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(0u)
            : "memory");
    }
}

__device__ __forceinline__ void stage_tables(unsigned char* s_tab, const uint16_t* __restrict__ g_tab, uint64_t* s_bar)
{
    stage_image(s_tab, g_tab, TAB_BYTES, s_bar);
}

__device__ __forceinline__ uint32_t lds_u16(uint32_t a)
{
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void red_inc_shared(uint32_t a)
{
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a) : "memory");
}

__device__ __forceinline__ uint64_t ld_nc_u64(const uint64_t* p)
{
    uint64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ uint64_t hand_from_ids(const uint8_t* p)
{
    uint64_t m = 0;
#pragma unroll
    for (int i = 0; i < 7; i++) m |= id_bit(__ldg(p + i));
    return m;
}

// ------------------------------------------------------------------------------------------------
// K1: batched eval7 (evaluate_cards, env/game.py:675)
// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch: the kernel is launched with programmaticStreamSerialization, so in a stream of batches
// its CTAs become resident as the previous grid's CTAs retire and stage the table image while that grid drains;
// griddepcontrol.wait (a no-op without a programmatic predecessor) orders every access to caller memory after the
// previous grid has completed.
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <int LAYOUT>
__global__ void __launch_bounds__(kCtaThreads, 1)
k_eval7_smem(const uint16_t* __restrict__ g_tab, const void* __restrict__ hands, int64_t n, uint16_t* __restrict__ out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    stage_tables(smem, g_tab, reinterpret_cast<uint64_t*>(smem + kSmemTab));
    const SharedTab tab{(uint32_t)__cvta_generic_to_shared(smem)};
    const HashParams hp = c_hp;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    grid_dependency_wait();
    if (LAYOUT == PHE_LAYOUT_MASKS) {
        const uint64_t* h = static_cast<const uint64_t*>(hands);
        const bool vec = ((reinterpret_cast<uintptr_t>(h) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 7) == 0);
        const int64_t n4 = vec ? (n >> 2) : 0;
        const ulonglong2* hv = reinterpret_cast<const ulonglong2*>(h);
        // four hands per thread per trip: two 128-bit loads, one 64-bit store.  The loads of trip i+1 are issued before
        // trip i is evaluated (one CTA of 1024 threads per SM holds only 32 KB in flight per trip: not enough to cover
        // the HBM latency at full bandwidth).
        int64_t q = gtid;
        ulonglong2 v0 = make_ulonglong2(0, 0), v1 = v0;
        if (q < n4) { v0 = __ldcs(hv + 2 * q); v1 = __ldcs(hv + 2 * q + 1); }
        grid_launch_dependents();
        for (; q < n4; q += nthr) {
            const int64_t qn = q + nthr;
            ulonglong2 w0 = v0, w1 = v1;
            if (qn < n4) { w0 = __ldcs(hv + 2 * qn); w1 = __ldcs(hv + 2 * qn + 1); }
            const uint32_t r0 = eval7(v0.x, tab, hp), r1 = eval7(v0.y, tab, hp);
            const uint32_t r2 = eval7(v1.x, tab, hp), r3 = eval7(v1.y, tab, hp);
            uint2 o;
            o.x = r0 | (r1 << 16);
            o.y = r2 | (r3 << 16);
            __stcs(reinterpret_cast<uint2*>(out) + q, o);
            v0 = w0; v1 = w1;
        }
        for (int64_t i = n4 * 4 + gtid; i < n; i += nthr) out[i] = (uint16_t)eval7(ld_nc_u64(h + i), tab, hp);
    } else {
        grid_launch_dependents();
        const uint8_t* h = static_cast<const uint8_t*>(hands);
        // id layout: a warp takes 128 consecutive hands = 896 bytes = 224 words.  The words are loaded coalesced (7 per
        // lane), transposed through a per-warp shared-memory strip (lane L then owns words 7L..7L+6 = its four hands;
        // stride 7 is conflict free), and the 28 ids become four card sets with packed byte arithmetic:
        // pos = 16*(id&3) + (id>>2) for four ids per word at once, bit = 1 << pos by two clamping 32-bit shifts
        // (pos for the low word, pos ^ 32 for the high word: the half that does not hold the bit shifts out to 0).
        const bool vec = ((reinterpret_cast<uintptr_t>(h) & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 7) == 0);
        const int64_t nchunk = vec ? (n >> 7) : 0;
        const uint32_t* hw = reinterpret_cast<const uint32_t*>(h);
        const uint32_t lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
        uint32_t* stg = reinterpret_cast<uint32_t*>(smem + kSmemTabOnly) + wic * 224u;
        const int64_t nwarps = nthr >> 5;
        int64_t c = gtid >> 5;
        uint32_t ld[7];
        if (c < nchunk) {
#pragma unroll
            for (int j = 0; j < 7; j++) ld[j] = __ldcs(hw + c * 224 + lane + 32 * j);
        }
        for (; c < nchunk; c += nwarps) {
#pragma unroll
            for (int j = 0; j < 7; j++) stg[lane + 32 * j] = ld[j];
            __syncwarp();
            uint32_t w[7];
#pragma unroll
            for (int j = 0; j < 7; j++) w[j] = stg[7 * lane + j];
            __syncwarp();
            const int64_t cn = c + nwarps;
            if (cn < nchunk) {
#pragma unroll
                for (int j = 0; j < 7; j++) ld[j] = __ldcs(hw + cn * 224 + lane + 32 * j);
            }
            uint32_t plo[7], phi[7];
#pragma unroll
            for (int j = 0; j < 7; j++) {
                plo[j] = ((w[j] << 4) & 0x30303030u) | ((w[j] >> 2) & 0x0F0F0F0Fu);
                phi[j] = plo[j] ^ 0x20202020u;
            }
            uint32_t r[4];
#pragma unroll
            for (int hnd = 0; hnd < 4; hnd++) {
                uint32_t lo = 0, hi = 0;
#pragma unroll
                for (int k = 0; k < 7; k++) {
                    const int bi = 7 * hnd + k;
                    const uint32_t sel = 0x4440u | (uint32_t)(bi & 3);              // byte (bi & 3) of the word, zero extended
                    uint32_t sl, sh, bl, bh;
                    asm("prmt.b32 %0, %1, 0, %2;" : "=r"(sl) : "r"(plo[bi >> 2]), "r"(sel));
                    asm("prmt.b32 %0, %1, 0, %2;" : "=r"(sh) : "r"(phi[bi >> 2]), "r"(sel));
                    asm("shl.b32 %0, 1, %1;" : "=r"(bl) : "r"(sl));                  // clamps: shift >= 32 gives 0
                    asm("shl.b32 %0, 1, %1;" : "=r"(bh) : "r"(sh));
                    lo |= bl;
                    hi |= bh;
                }
                r[hnd] = eval7(((uint64_t)hi << 32) | lo, tab, hp);
            }
            uint2 o;
            o.x = r[0] | (r[1] << 16);
            o.y = r[2] | (r[3] << 16);
            __stcs(reinterpret_cast<uint2*>(out) + (c * 32 + lane), o);
        }
        for (int64_t i = nchunk * 128 + gtid; i < n; i += nthr) out[i] = (uint16_t)eval7(hand_from_ids(h + 7 * i), tab, hp);
    }
}

template <int LAYOUT>
__global__ void __launch_bounds__(256)
k_eval7_gmem(const uint16_t* __restrict__ g_tab, const void* __restrict__ hands, int64_t n, uint16_t* __restrict__ out)
{
    const GlobalTab tab{g_tab};
    const HashParams hp = c_hp;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t m = LAYOUT == PHE_LAYOUT_MASKS ? static_cast<const uint64_t*>(hands)[i]
                                                  : hand_from_ids(static_cast<const uint8_t*>(hands) + 7 * i);
    out[i] = (uint16_t)eval7(m, tab, hp);
}

// ------------------------------------------------------------------------------------------------
// K2: exhaustive C(52,7) over a colex index range (arithmetic and table layout: phe_enum.cuh).
// A warp walks tiles of consecutive colex hand indices block by block: a block is one (c3..c6) suffix whose
// C(c3,3) triples are consecutive hand indices; the 32 lanes stride over the triples.  Everything outside
// the triple loop is warp-uniform; the only divergent code is the flush path (3 % of hands).
// ------------------------------------------------------------------------------------------------
// FLUSH-table address of the hand (block s3, triple t) whose flush suit is flagged in x (one bit of 0x8888)
__device__ __forceinline__ uint32_t enum_flush_addr(uint32_t x, uint32_t t, uint32_t s3lo, uint32_t s3hi, uint32_t a_fl,
                                                    const uint64_t* __restrict__ g_trimask)
{
    const uint2 m = __ldg(reinterpret_cast<const uint2*>(g_trimask) + t);
    uint32_t pos;                                                // bit 3, 7, 11 or 15: the flagged nibble (bfind = FLO, no 31 - clz round trip)
    asm("bfind.u32 %0, %1;" : "=r"(pos) : "r"(x));
    const uint32_t sel = (pos >> 2) * 0x22u + 0x10u;             // PRMT selector: bytes 2s and 2s+1 of {hi:lo}, s = suit 0..3
    uint32_t two_lanes, addr;                                    // prmt directly: __byte_perm also masks the selector
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(two_lanes) : "r"(s3lo | m.x), "r"(s3hi | m.y), "r"(sel));
    asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(addr) : "r"(two_lanes & 0x1FFFu), "r"(a_fl));
    return addr;
}

// ---- fused all-reduce epilogue (phe_enum7_allreduce) --------------------------------------------------------
// In fused mode every CTA adds its 30 KB shared-memory histogram (u32 bins; a whole C(52,7) bin fits 32 bits) into a
// per-launch u32 array in global memory with ONE TMA bulk reduction (cp.reduce.async.bulk .add.u32: the adds are performed
// by the L2, no thread touches a bin).  The last CTA of the launch pushes that array to the peers, sums what the peers
// pushed, and clears it for the next launch.  (The plain kernel keeps per-CTA 64-bit REDs straight into the caller's
// histograms: with no exchange to feed, a second level only adds a serial pass -- measured 2 us slower.)
//
// Workspace of one rank (all ranks use the same layout; peers write into it through NVLink):
//   slots  u32[2][8][7464]   payload (bins, one pad word) of source rank s for launches of parity p
//   flags  u32[8]            flags[s] = number of the last launch whose contribution from rank s is complete
//   acc    u32[2][7464]      NVLS mode only: the sum itself, by launch parity (every rank multimem.red-adds its payload into
//                            the multicast mapping of this array, i.e. into all ranks' copies at once; no slots are used)
//   epoch  u32, failed u32   launch number (starts at 0 -> first launch is 1); launch number of a timed-out exchange
constexpr int kAllBins = 9 + kHistBins;                       // 7472 (layout of the u64 result: hist9 then hist7463)
static_assert(kAllBins == 7472, "phe_enum7_allreduce writes out_hist7472");
constexpr int kMaxWorld = 8;
constexpr int kSlotWords = kHistBinsPad;                      // 7464 u32 = 1866 uint4
constexpr size_t kWsSlots = 0;
constexpr size_t kWsFlags = kWsSlots + (size_t)2 * kMaxWorld * kSlotWords * 4;
constexpr size_t kWsAcc = kWsFlags + 64;
constexpr size_t kWsEpoch = kWsAcc + (size_t)2 * kSlotWords * 4;
constexpr size_t kWsBytes = kWsEpoch + 128;                   // epoch, failed, then (PHE_ENUM_PROF) eight %globaltimer stamps
static_assert(kSlotWords % 4 == 0 && kWsFlags % 16 == 0 && kWsAcc % 16 == 0, "uint4 access to the slots");

struct FusedArgs {
    int rank, world;                       // world == 0: plain kernel, results go to hist9 / hist7463
    unsigned char* ws[kMaxWorld];          // every rank's workspace
    unsigned long long* out;               // local all-reduced histogram [7472]
    unsigned char* mc;                     // NVLS: multicast mapping of the workspaces (one store / reduction reaches every rank), or null
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void multimem_red_add_u32(void* mc_addr, uint32_t v)
{
    asm volatile("multimem.red.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(mc_addr), "r"(v) : "memory");
}
__device__ __forceinline__ void multimem_st_release_u32(void* mc_addr, uint32_t v)
{
    asm volatile("multimem.st.release.sys.global.u32 [%0], %1;" ::"l"(mc_addr), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#ifdef PHE_ENUM_PROF
#define PHE_ENUM_STAMP(ws, k) do { if (threadIdx.x == 0) reinterpret_cast<unsigned long long*>((ws) + kWsEpoch + 16)[k] = globaltimer_ns(); } while (0)
__device__ unsigned long long g_enum_prof[256 * 4];   // per CTA of the last launch: start, tables staged, loop done, flush done
#define PHE_ENUM_CTA_STAMP(k) do { if (threadIdx.x == 0) g_enum_prof[blockIdx.x * 4 + (k)] = globaltimer_ns(); } while (0)
#else
#define PHE_ENUM_STAMP(ws, k) do { } while (0)
#define PHE_ENUM_CTA_STAMP(k) do { } while (0)
#endif

constexpr int kHistVec = kHistBinsPad / 4;                              // 1866 uint4 groups of bins
constexpr int kVecPer = (kHistVec + kCtaThreads - 1) / kCtaThreads;     // 2 per thread

// Executed by the last CTA: push this rank's u32 payload into its slot on every rank, signal, wait for the peers' slots,
// sum.  Ordering follows the usual one-thread pattern: the payload stores of all threads are ordered before the flag by
// bar.sync + st.release.sys of the signalling threads (release is cumulative), and the peers' payloads are ordered after
// ld.acquire.sys + bar.sync on the reading side; no other thread issues a system-scope fence.
__device__ __noinline__ void enum_allreduce_epilogue(const FusedArgs& fa, uint32_t* g_h32, uint32_t* s_h9)
{
    unsigned char* my = fa.ws[fa.rank];
    uint32_t* epoch_p = reinterpret_cast<uint32_t*>(my + kWsEpoch);
    const uint32_t e = *reinterpret_cast<volatile uint32_t*>(epoch_p) + 1u;
    const size_t slot_off = kWsSlots + ((size_t)(e & 1u) * kMaxWorld + (size_t)fa.rank) * kSlotWords * 4;
    const int world = fa.world;
    unsigned long long* const out = fa.out;
    PHE_ENUM_STAMP(my, 1);
    // 1. payload -> every rank; g_h32 cleared for the next launch.  Loads first, then the stores.
    //    Peer-store mode: into slot (e, rank) of every rank's workspace.  NVLS mode: multimem.red.add into the multicast
    //    mapping of acc[e & 1] -- the switch delivers the addition to all ranks' copies, so one reduction per non-zero bin
    //    replaces `world` stores and nobody has to sum slots afterwards; acc[(e + 1) & 1] of THIS rank is cleared here,
    //    before the signal: peers add to it only in launch e + 1, which they start after they have seen this launch's flag.
    //    (Also tried: leaving the payload at home and pulling the sum with multimem.ld_reduce.add.u64 on bin pairs -- no push
    //    at all, but the 3732 eight-byte reductions per rank take 9.4 us against 3.0 + 3.8 us for this push + local read.)
    const bool nvls = fa.mc != nullptr;
    const size_t acc_off = kWsAcc + (size_t)(e & 1u) * kSlotWords * 4;
    {
        uint4 v[kVecPer];
#pragma unroll
        for (int k = 0; k < kVecPer; k++) {
            const int g = threadIdx.x + k * kCtaThreads;
            v[k] = g < kHistVec ? __ldcg(reinterpret_cast<const uint4*>(g_h32) + g) : make_uint4(0, 0, 0, 0);
        }
        if (nvls) {
            uint32_t* dst = reinterpret_cast<uint32_t*>(fa.mc + acc_off);
            uint4* next = reinterpret_cast<uint4*>(my + kWsAcc + (size_t)((e + 1u) & 1u) * kSlotWords * 4);
#pragma unroll
            for (int k = 0; k < kVecPer; k++) {
                const int g = threadIdx.x + k * kCtaThreads;
                if (g >= kHistVec) continue;
                if (v[k].x) multimem_red_add_u32(dst + 4 * g + 0, v[k].x);
                if (v[k].y) multimem_red_add_u32(dst + 4 * g + 1, v[k].y);
                if (v[k].z) multimem_red_add_u32(dst + 4 * g + 2, v[k].z);
                if (v[k].w) multimem_red_add_u32(dst + 4 * g + 3, v[k].w);
                next[g] = make_uint4(0, 0, 0, 0);
            }
        } else {
#pragma unroll 1   // once-per-launch code runs from a cold instruction cache: compact beats unrolled here
            for (int p = 0; p < world; p++) {
                uint4* dst = reinterpret_cast<uint4*>(fa.ws[p] + slot_off);
#pragma unroll
                for (int k = 0; k < kVecPer; k++) {
                    const int g = threadIdx.x + k * kCtaThreads;
                    if (g < kHistVec) dst[g] = v[k];
                }
            }
        }
#pragma unroll
        for (int k = 0; k < kVecPer; k++) {
            const int g = threadIdx.x + k * kCtaThreads;
            if (g < kHistVec) reinterpret_cast<uint4*>(g_h32)[g] = make_uint4(0, 0, 0, 0);
        }
    }
    __syncthreads();
    PHE_ENUM_STAMP(my, 2);
    // 2. tell every rank that slot (e, rank) is complete; 3. wait until all of this rank's slots of launch e are
    if ((int)threadIdx.x < world) {
        if (!nvls) st_release_sys(reinterpret_cast<uint32_t*>(fa.ws[threadIdx.x] + kWsFlags) + fa.rank, e);
        else if (threadIdx.x == 0) multimem_st_release_u32(reinterpret_cast<uint32_t*>(fa.mc + kWsFlags) + fa.rank, e);   // one store, every rank's flags[rank]
        const uint32_t* mine = reinterpret_cast<const uint32_t*>(my + kWsFlags) + threadIdx.x;
        const long long t0 = clock64();
        bool ok = true;
        while ((int32_t)(ld_acquire_sys(mine) - e) < 0) {
            if (clock64() - t0 > 4000000000ll) { ok = false; break; }   // ~2 s: a peer never launched
            __nanosleep(64);
        }
        // a late peer leaves a PARTIAL sum: recorded per launch in a status word of its own (epoch_p[1] = the launch
        // number that failed; never a histogram bin), read by phe_enum7_allreduce_status / FusedEnum7.check()
        if (!ok) atomicExch(epoch_p + 1, e);
    }
    __syncthreads();
    PHE_ENUM_STAMP(my, 3);
    // 4. sum the slots (all loads of a group in flight together), widen to the u64 result: hist9 at [0, 9), bins at [9, 7472).
    //    The nine category counts are segment sums of the summed bins: bins are ordered by category, so a warp's 128
    //    consecutive bins nearly always lie in one category -> one warp reduction and one shared atomic per warp.
    const unsigned char* base = nvls ? my + acc_off : my + kWsSlots + (size_t)(e & 1u) * kMaxWorld * kSlotWords * 4;
    const int n_slots = nvls ? 1 : world;                              // NVLS: acc already holds the sum of all ranks
    unsigned long long sum[kVecPer][4];
#pragma unroll
    for (int k = 0; k < kVecPer; k++) sum[k][0] = sum[k][1] = sum[k][2] = sum[k][3] = 0;
#pragma unroll 4
    for (int s = 0; s < n_slots; s++) {                                // slot loop outside: the loads of both groups fly together
        const uint4* slot = reinterpret_cast<const uint4*>(base + (size_t)s * kSlotWords * 4);
#pragma unroll
        for (int k = 0; k < kVecPer; k++) {
            const int g = threadIdx.x + k * kCtaThreads;
            if (g < kHistVec) {
                const uint4 w = __ldcg(slot + g);
                sum[k][0] += w.x; sum[k][1] += w.y; sum[k][2] += w.z; sum[k][3] += w.w;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kVecPer; k++) {
        const int g = threadIdx.x + k * kCtaThreads;                  // whole warps take part in the reduction below
        if (g < kHistVec) {
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (4 * g + j < kHistBins) out[9 + 4 * g + j] = sum[k][j];
        }
        const int b0 = min(4 * g, kHistBins - 1), b3 = min(4 * g + 3, kHistBins - 1);
        const int ca = category_of((uint32_t)b0), cb = category_of((uint32_t)b3);
        const int c_first = __shfl_sync(0xFFFFFFFFu, ca, 0);
        if (__all_sync(0xFFFFFFFFu, ca == c_first && cb == c_first)) {
            const uint32_t tot = __reduce_add_sync(0xFFFFFFFFu, (uint32_t)(sum[k][0] + sum[k][1] + sum[k][2] + sum[k][3]));   // < 2^32: at most C(52,7)
            if ((threadIdx.x & 31) == 0 && tot) atomicAdd(s_h9 + c_first, tot);
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (sum[k][j]) atomicAdd(s_h9 + category_of((uint32_t)min(4 * g + j, kHistBins - 1)), (uint32_t)sum[k][j]);
        }
    }
    PHE_ENUM_STAMP(my, 5);
    __syncthreads();
    PHE_ENUM_STAMP(my, 6);
    if (threadIdx.x < 9) out[threadIdx.x] = s_h9[threadIdx.x];
    if (threadIdx.x == 0) *epoch_p = e;
    PHE_ENUM_STAMP(my, 4);
}

__global__ void __launch_bounds__(kCtaThreads, 1)
k_enum7(const unsigned char* __restrict__ g_img, const uint64_t* __restrict__ g_trimask, unsigned long long* __restrict__ hist9,
        unsigned long long* __restrict__ hist7463, uint64_t begin, uint64_t end, const uint32_t* __restrict__ g_units, const uint64_t* __restrict__ g_unit_starts, const uint4* __restrict__ g_blocks, uint32_t unit_first,
        uint32_t n_units, uint32_t units_per_tile, uint32_t my_tiles, uint32_t n_static, unsigned int* __restrict__ tile_counter, uint32_t* __restrict__ g_h32, uint32_t part, uint32_t nparts, const __grid_constant__ FusedArgs fa)
{
    extern __shared__ __align__(128) unsigned char smem[];
#ifdef PHE_ENUM_PROF
    if (fa.world > 0 && blockIdx.x == 0) PHE_ENUM_STAMP(fa.ws[fa.rank], 0);
    PHE_ENUM_CTA_STAMP(0);
#endif
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(smem + kEnumHistOff);
    uint32_t* s_h9 = reinterpret_cast<uint32_t*>(smem + kEnumH9Off);   // per-CTA category counts fit 32 bits
    for (int i = threadIdx.x; i < kHistBinsPad; i += blockDim.x) s_hist[i] = 0;
    if (threadIdx.x < 9) s_h9[threadIdx.x] = 0;
    uint32_t* s_bin = reinterpret_cast<uint32_t*>(smem + kEnumBinOff);
    if (threadIdx.x < 512) {
        const int k = threadIdx.x >> 6, v = threadIdx.x & 63;
        s_bin[threadIdx.x] = v <= 52 ? c_binom[v * 8 + k] : 0xFFFFFFFFu;
    }
    stage_image(smem, g_img, ENUM_IMG_BYTES, reinterpret_cast<uint64_t*>(smem + kEnumBarOff));
    __syncthreads();
    PHE_ENUM_CTA_STAMP(1);

    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t a_tri = sb + EOFF_TRI, a_ms = sb + EOFF_MSRANK, a_fl = sb + EOFF_FLUSH, a_hist = sb + kEnumHistOff;
    const uint32_t lane = threadIdx.x & 31;
    const Binom C{c_binom};

    // Tiles: a tile = `units_per_tile` consecutive work units of the colex axis (g_units: at most 512 hands or 4 enumeration
    // blocks each, see build_enum_units).  Cutting by block count as well as by hands bounds the serial block set-ups one warp
    // can be handed (runs of tiny blocks occur wherever c4 restarts), which is what kept single warps busy long after the rest
    // of the launch had drained.
    const uint32_t a_bin = sb + kEnumBinOff;
    // A launch owns the tiles == part (mod nparts); its j-th tile goes to warp j (mod W) for j < n_static (no atomic, W = warps
    // of the grid: up to six rounds), the rest through the counter.  One 4736-way contended counter for EVERY tile was the
    // limiter of short launches (a sharded part hands out 1.4 tiles/ns from one address, and each claim is a serial L2 round
    // trip in front of the tile); tiles cost about the same by construction, and the dynamic rest evens out the finish line.
    const uint32_t n_warps = gridDim.x * (kCtaThreads / 32);
    uint32_t j = blockIdx.x * (kCtaThreads / 32) + (threadIdx.x >> 5);
    bool dynamic = false;
    for (;; j += n_warps) {
        if (!dynamic && j >= n_static) dynamic = true;
        if (dynamic) {   // (claiming one tile ahead hides the round trip but reserves a tile behind a busy warp: the finish line got 1-2 us worse)
            uint32_t got = 0;
            if (lane == 0) got = atomicAdd(tile_counter, 1u);
            j = n_static + __shfl_sync(0xFFFFFFFFu, got, 0);
        }
        if (j >= my_tiles) break;
        const uint32_t tile = j * nparts + part;
        const uint32_t u0 = tile * units_per_tile;
        const uint32_t u1 = min(u0 + units_per_tile, n_units);
        uint64_t wb = __ldg(g_units + unit_first + u0), we = __ldg(g_units + unit_first + u1);
        const uint64_t packed = __ldg(g_unit_starts + unit_first + u0);
        if (we > end) we = end;                                      // the last unit of a sub-range is clipped
        uint32_t blk, t_start;
        if (wb >= begin) {                                           // the table knows where the unit starts
            blk = (uint32_t)packed & 0x3FFFFu;
            t_start = (uint32_t)(packed >> 18);
        } else {                                                     // first unit of a sub-range that begins inside it
            wb = begin;
            // warp-cooperative colex unrank: lane v (and v + 32) tests C(v, k) <= rem; the count of hits - 1 is c_i
            int c[7];
            uint32_t rem = (uint32_t)wb;
#pragma unroll
            for (int i = 6; i >= 0; i--) {
                const uint32_t b0 = lds_u32(a_bin + 4u * ((uint32_t)(i + 1) * 64u + lane));
                const uint32_t b1 = lds_u32(a_bin + 4u * ((uint32_t)(i + 1) * 64u + lane + 32u));
                const int hits = __popc(__ballot_sync(0xFFFFFFFFu, b0 <= rem)) + __popc(__ballot_sync(0xFFFFFFFFu, lane + 32u <= 51u && b1 <= rem));
                c[i] = hits - 1;
                rem -= lds_u32(a_bin + 4u * ((uint32_t)(i + 1) * 64u + (uint32_t)c[i]));
            }
            blk = C(c[3] - 3, 1) + C(c[4] - 3, 2) + C(c[5] - 3, 3) + C(c[6] - 3, 4);   // colex rank of the shifted suffix
            t_start = C(c[2], 3) + C(c[1], 2) + (uint32_t)c[0];
        }
        if (wb >= we) continue;
        uint32_t remaining = (uint32_t)(we - wb);                    // a tile is at most 4 units = 2048 hands: 32-bit bookkeeping
        // block records come from the L2-resident table (one 16-byte load instead of eight binomial look-ups and the
        // suit arithmetic of enum_block_setup: block set-up was a fifth of all executed instructions); the next block's
        // record is requested before this block's hands are processed
        uint4 rec = __ldg(g_blocks + blk);
        for (;;) {
            const uint4 rec_next = __ldg(g_blocks + min(blk + 1u, N_ENUM_BLOCKS - 1u));
            uint32_t stop = rec.z >> 16;                              // ntriples
            stop = min(stop, t_start + remaining);                   // t_start < 17296, remaining <= 2048: no overflow
            const uint32_t msb = a_ms + 2u * (rec.z & 0xFFFFu);
            const uint32_t s3lo = rec.x, s3hi = rec.y, kbias = rec.w;
            // warp-uniform trip counts; kU hands per lane per trip keep kU independent LDS -> (LDG) -> LDS -> ATOMS
            // chains in flight, which is what hides the L2 round trip of the flush path
#ifndef PHE_ENUM_KU
#define PHE_ENUM_KU 4
#endif
            constexpr int kU = PHE_ENUM_KU;
            uint32_t tb = t_start;
            for (; tb + 32 * kU <= stop; tb += 32 * kU) {
                const uint32_t t0 = tb + lane;
                uint32_t e[kU], a[kU], x[kU], any = 0;
#pragma unroll
                for (int u = 0; u < kU; u++) e[u] = lds_u32(a_tri + 4u * t0 + 128u * u);
#pragma unroll
                for (int u = 0; u < kU; u++) {
                    a[u] = msb + (e[u] >> 16);
                    x[u] = (e[u] + kbias) & 0x8888u;
                    any |= x[u];
                }
                if (__any_sync(0xFFFFFFFFu, any != 0u)) {   // some lane holds a flush (3 % of hands; taken by 98 % of the trips, yet dropping the vote costs 7 %)
                    // two phases: first every flagged hand requests its triple's card set, then the addresses are formed -- the L2
                    // round trips of the trip's flush hands overlap instead of queueing behind each other inside four separate
                    // branch regions (long-scoreboard was the top stall; +1.5 % measured)
                    uint2 tri_set[kU];
#pragma unroll
                    for (int u = 0; u < kU; u++) {
                        tri_set[u] = make_uint2(0u, 0u);
                        if (x[u]) tri_set[u] = __ldg(reinterpret_cast<const uint2*>(g_trimask) + (t0 + 32u * u));
                    }
#pragma unroll
                    for (int u = 0; u < kU; u++)
                        if (x[u]) {
                            const uint32_t w = x[u] > 0xFFu ? (s3hi | tri_set[u].y) : (s3lo | tri_set[u].x);   // suits 2, 3 live in the high word
                            const uint32_t lane13 = ((x[u] & 0x8080u) ? w >> 16 : w) & 0x1FFFu;               // odd suits in the high half-word
                            a[u] = a_fl + 2u * lane13;
                        }
                }
                uint32_t r[kU];
#pragma unroll
                for (int u = 0; u < kU; u++) r[u] = lds_u16(a[u]);
#pragma unroll
                for (int u = 0; u < kU; u++) red_inc_shared(a_hist + 4u * r[u]);
            }
            for (; tb < stop; tb += 32) {   // at most kU partial trips
                const uint32_t t0 = tb + lane;
                if (t0 < stop) {
                    const uint32_t e0 = lds_u32(a_tri + 4u * t0);
                    uint32_t a0 = msb + (e0 >> 16);
                    const uint32_t x0 = (e0 + kbias) & 0x8888u;
                    if (x0) a0 = enum_flush_addr(x0, t0, s3lo, s3hi, a_fl, g_trimask);
                    red_inc_shared(a_hist + 4u * lds_u16(a0));
                }
            }
            remaining -= stop - t_start;
            if (!remaining) break;
            t_start = 0;
            blk++;
            rec = rec_next;
        }
    }
    __syncthreads();
    PHE_ENUM_CTA_STAMP(2);
    __shared__ uint32_t s_last;
    if (fa.world > 0) {
        // fused mode: one TMA bulk reduction of the whole shared histogram into the launch's u32 array (adds done by the L2)
        if (threadIdx.x == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the shared atomics above -> visible to the TMA read
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.u32 [%0], [%1], %2;" ::"l"(g_h32), "r"(a_hist), "r"((uint32_t)(kHistBinsPad * 4))
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // reads AND global adds of the group complete
            asm volatile("fence.proxy.async;" ::: "memory");
        }
    } else {
        // plain mode: every CTA starts at a different bin so the 64-bit REDs do not pile up on one L2 line.  Category counts:
        // bins are ordered by category, so the 32 consecutive bins of a warp nearly always share one -> a warp reduction
        // and ONE shared atomic (32 lanes adding to the same shared word serialise)
        const int rot = (int)((blockIdx.x * 397u) % (uint32_t)kHistBins);
        for (int k0 = 0; k0 < kHistBins; k0 += blockDim.x) {           // whole warps iterate: the reduction is warp-wide
            const int k = k0 + threadIdx.x;
            int i = k + rot;
            if (i >= kHistBins) i -= kHistBins;
            const uint32_t v = k < kHistBins ? s_hist[i] : 0u;
            if (v && hist7463) atomicAdd(hist7463 + i, (unsigned long long)v);
            if (hist9) {
                const int cat = category_of((uint32_t)min(i, kHistBins - 1));
                const int c_first = __shfl_sync(0xFFFFFFFFu, cat, 0);
                if (__all_sync(0xFFFFFFFFu, cat == c_first || v == 0u)) {
                    const uint32_t tot = __reduce_add_sync(0xFFFFFFFFu, v);   // per-CTA counts fit 32 bits
                    if (lane == 0 && tot) atomicAdd(s_h9 + c_first, tot);
                } else if (v) atomicAdd(s_h9 + cat, v);                 // native 32-bit shared atomic (the 64-bit form is a CAS loop)
            }
        }
        __syncthreads();
        if (hist9 && threadIdx.x < 9 && s_h9[threadIdx.x]) atomicAdd(hist9 + threadIdx.x, (unsigned long long)s_h9[threadIdx.x]);
        __threadfence();
        __syncthreads();
    }
    PHE_ENUM_CTA_STAMP(3);
    // completion ticket: the last CTA re-arms the tile counter for the next launch (no memset node in front of the
    // kernel) and, in fused mode, owns the all-reduce
    if (threadIdx.x == 0) {
        __threadfence();
        const bool last = atomicAdd(tile_counter + 1, 1u) == gridDim.x - 1;
        if (last) {
            tile_counter[0] = 0;
            tile_counter[1] = 0;
            __threadfence();                                              // acquire side: the other CTAs' flushes, for the whole CTA via the barrier
            asm volatile("fence.proxy.async;" ::: "memory");
        }
        s_last = last ? 1u : 0u;
    }
    __syncthreads();
    if (s_last && fa.world > 0) {
        if (threadIdx.x < 9) s_h9[threadIdx.x] = 0;
        __syncthreads();
        enum_allreduce_epilogue(fa, g_h32, s_h9);
    }
}

// ------------------------------------------------------------------------------------------------
// Monte-Carlo kernels: the deal table (pairs + singles) in shared memory and the per-thread slot scratch -- row k of a
// [MAX_PAIR_SLOTS][blockDim] u32 matrix holds the k-th accepted pair index (threads of a warp touch 32 consecutive
// words of one row: conflict free; 32-bit cells avoid the 16-bit conversions around the store); the single card set stays in a register.
struct SharedDealTab {
    uint32_t base;
    uint32_t eight;    // 8, derived from the opaque c_hp.two: the entry address is q * eight + base, an IMAD on the FMA pipe
                       // (the alu pipe bounds the kernel and a visible 8 becomes an LEA there: 23 of them per warp-rollout)
    __device__ __forceinline__ uint64_t mask(uint32_t q) const
    {
        uint64_t v;
        asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(q * eight + base));
        return v;
    }
};

struct SmemSlots {
    uint32_t base;     // shared address of this thread's word in row 0
    uint32_t stride;   // bytes per row = 4 * blockDim.x
    uint64_t single;
    __device__ __forceinline__ void put_single(uint64_t m) { single = m; }
    __device__ __forceinline__ uint64_t get_single() const { return single; }
    template <class DealTab>
    __device__ __forceinline__ uint64_t get_pair(int k, const DealTab& dt) const
    {
        uint32_t q;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(q) : "r"(base + (uint32_t)k * stride) : "memory");
        return dt.mask(q);
    }
    typedef uint32_t cursor;      // shared address of the next free cell of this thread's column
    __device__ __forceinline__ cursor cursor_at(int k) const { return base + (uint32_t)k * stride; }
    __device__ __forceinline__ cursor cursor_next(cursor c, uint32_t one) const { return add_fma(c, stride, one); }
    __device__ __forceinline__ int cursor_index(cursor c) const { return (int)((c - base) / stride); }
    __device__ __forceinline__ void put_at(cursor c, uint32_t q, uint64_t)
    {
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(c), "r"(q) : "memory");
    }
};

// ------------------------------------------------------------------------------------------------
// K3: Monte-Carlo equity.  One warp works on one (state, sample chunk) item at a time; lanes take
// samples chunk_begin + lane + 32 i.  The deal is a pure function of the global sample index.
// ------------------------------------------------------------------------------------------------
template <class Tab, class DealTab, class Key>
__device__ __forceinline__ void equity_mc_body(const Tab& tab, const DealTab& dt, const ulonglong2* __restrict__ states, int64_t n_states,
                                               int n_opp_default, uint64_t rollouts, uint64_t sample_offset, const Key& key,
                                               uint32_t state_base, unsigned long long* __restrict__ wtl, uint32_t chunk,
                                               SmemSlots slots, unsigned int* __restrict__ item_counter)
{
    const HashParams hp = c_hp;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warps_per_cta = blockDim.x >> 5;
    const uint64_t nwarps = (uint64_t)gridDim.x * warps_per_cta;
    const uint64_t chunks_per_state = (rollouts + chunk - 1) / chunk;
    const uint64_t items = (uint64_t)n_states * chunks_per_state;
    // Items cost differently (a pre-flop six-max deal places 15 cards, a river deal 10) and SMs do not run at one speed:
    // with a counter the warps pull items as they finish (the static round-robin left SMs idle for up to 15 % of the
    // kernel).  Counts are integer sums, so the schedule does not change the result.
    uint64_t item = (uint64_t)blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
    for (;; item += nwarps) {
        if (item_counter) {
            uint32_t nxt = 0;
            if (lane == 0) nxt = atomicAdd(item_counter, 1u);
            item = __shfl_sync(0xFFFFFFFFu, nxt, 0);
        }
        if (item >= items) break;
        // consecutive items belong to DIFFERENT states: warps that run side by side then work on different streets, whose
        // dealing / evaluation mix differs -- a batch of few states with many chunks each would otherwise run one state at a time
        const uint64_t ch = item / (uint64_t)n_states, s = item - ch * (uint64_t)n_states;
        const ulonglong2 st = states[s];
        const uint64_t known = st.x & CARD_BITS, hero = st.y & CARD_BITS;
        const int over = (int)(st.y >> 61);
        const int n_opp = over ? over : n_opp_default;
        const int nboard = __popcll(known & ~hero);
        const uint64_t b = ch * chunk;
        uint64_t e = b + chunk;
        if (e > rollouts) e = rollouts;
        uint32_t w = 0, t = 0, l = 0;
        for (uint64_t i = b + lane; i < e; i += 32) {
            const int res = mc_rollout_k(tab, hp, dt, known, hero, nboard, n_opp, key, sample_offset + i, state_base + (uint32_t)s, slots);
            w += (res == 0); t += (res == 1); l += (res == 2);
        }
        w = __reduce_add_sync(0xFFFFFFFFu, w);
        t = __reduce_add_sync(0xFFFFFFFFu, t);
        l = __reduce_add_sync(0xFFFFFFFFu, l);
        if (lane == 0) {
            if (w) atomicAdd(wtl + 3 * s + 0, (unsigned long long)w);
            if (t) atomicAdd(wtl + 3 * s + 1, (unsigned long long)t);
            if (l) atomicAdd(wtl + 3 * s + 2, (unsigned long long)l);
        }
    }
}

__global__ void __launch_bounds__(kMcThreads, 1) PHE_MC_CLUSTER_ATTR
k_equity_mc_smem(const uint16_t* __restrict__ g_tab, const uint64_t* __restrict__ g_dealtab, const ulonglong2* __restrict__ states,
                 int64_t n_states, int n_opp, uint64_t rollouts, uint64_t sample_offset, const RoundKeys rkeys, uint32_t state_base,
                 unsigned long long* __restrict__ wtl, uint32_t chunk, unsigned int* __restrict__ counters)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* s_deal = reinterpret_cast<uint64_t*>(smem + kSmemTabOnly);
    for (uint32_t i = threadIdx.x; i < DEALTAB_ENTRIES; i += blockDim.x) s_deal[i] = g_dealtab[i];
    stage_tables(smem, g_tab, reinterpret_cast<uint64_t*>(smem + kSmemTab));   // includes a __syncthreads
    __syncthreads();
    uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem);
#ifndef PHE_MC_REMAT_BASE
    // always + 0, but opaque to ptxas: the shared-window base then stays in a register instead of being rebuilt
    // (S2R CgaCtaId, MOV, VIADD, LEA) in front of every table access of the dealing loop
    sb += (uint32_t)(rollouts >> 63);
#endif
    const SharedTab tab{sb};
    const SharedDealTab dt{sb + kSmemTabOnly, c_hp.two * c_hp.two * c_hp.two};
    const SmemSlots slots{sb + kSmemTabOnly + kDealTabBytes + 4u * threadIdx.x, 4u * blockDim.x, 0ull};
    equity_mc_body(tab, dt, states, n_states, n_opp, rollouts, sample_offset, rkeys, state_base, wtl, chunk, slots, counters);
    // completion ticket: the last CTA re-arms the {item counter, ticket} pair for the next launch that draws it
    __shared__ uint32_t s_last;
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(counters + 1, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        counters[0] = 0;
        counters[1] = 0;
    }
}

#ifdef PHE_MC_POOL
// ------------------------------------------------------------------------------------------------
// K3, experiment of round 2 (compiled with -DPHE_MC_POOL only): Monte-Carlo equity with dealing decoupled from evaluation.
//
// In the kernel above a warp deals 32 samples in lock step and waits for its slowest lane: rejection sampling needs a
// different number of draws per sample, so the dealing loop -- 60 % of the instructions -- ran at ~20 of 32 lanes (ncu:
// 16 draw slots executed per rollout against 9.3 needed, 2.5 Philox blocks against 1.7).  Here every lane walks its OWN
// sequence of samples (lane, lane + 32, ... of the item's chunk): one ROUND = one Philox block + its eight draws for
// whatever sample the lane is working on, all 32 lanes busy.  A lane that completes a deal publishes the shared-memory
// column holding its accepted slots on the warp's READY ring, takes a free column and starts its next sample in the
// following round; whenever 32 deals are ready the warp evaluates them together, lane j taking the j-th ready column --
// evaluation always runs on 32 lanes, whoever dealt the hands.  The deal of (state, sample) and therefore every count
// is unchanged (same stream contract as deal_slots / the oracle; tests/test_gpu_parity.py passes with either kernel).
//
// Outcome on the B200 (profiles/mc_streets_r02.jsonl, profiles/ncu_mc_r02_pool.json): 30.1 of 32 lanes active instead of
// 25.4 and 9 % fewer instructions, but every lane now reads the pair table for all eight draws of a round: 25 % more
// shared-memory wavefronts, 87 % of the wavefront rate -- the bound moved from the issue slots to shared memory.  Equal
// to the lock-step kernel on the mixed six-max batch of BASELINE config 3 (3.894e10 against 3.886e10 rollouts/s), 8 %
// faster on river-only six-max batches, 7-17 % slower on the other streets and 12 % slower heads-up.  The lock-step
// kernel therefore stays the product path.
//
// Per warp: 64 columns x 12 rows of u16 cells (rows 0..9 pair slots, row n_pairs doubles as the spare row of the
// clamped cursor, row 11 the single card) = [row][64] so that any 32 distinct columns are bank-conflict free, plus a
// 64-entry ring of column ids split into a free and a ready segment: 32 columns are always held by the lanes.
// ------------------------------------------------------------------------------------------------
static constexpr uint32_t kPoolRows = MAX_PAIR_SLOTS + 2;
static constexpr uint32_t kPoolRowBytes = 128;                         // 64 columns x u16
static constexpr uint32_t kPoolSingleRow = kPoolRows - 1;
static constexpr uint32_t kPoolWarpBytes = kPoolRows * kPoolRowBytes + 64;   // + ring of 64 column ids (u8)
static constexpr uint32_t kSmemMcPool = kSmemTabOnly + kDealTabBytes + (kCtaThreads / 32) * kPoolWarpBytes;
static_assert(kSmemMcPool <= 232448, "Monte-Carlo pool kernel: shared memory budget");

__device__ __forceinline__ uint32_t lds_u8(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((uint16_t)v) : "memory"); }
__device__ __forceinline__ uint32_t lds_u16v(uint32_t a)
{
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
    return v;
}

__global__ void __launch_bounds__(kCtaThreads, 1)
k_equity_mc_pool(const uint16_t* __restrict__ g_tab, const uint64_t* __restrict__ g_dealtab, const ulonglong2* __restrict__ states,
                 int64_t n_states, int n_opp_default, uint64_t rollouts, uint64_t sample_offset, uint64_t seed, uint32_t state_base,
                 unsigned long long* __restrict__ wtl, uint32_t chunk, unsigned int* __restrict__ counters)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* s_deal = reinterpret_cast<uint64_t*>(smem + kSmemTabOnly);
    for (uint32_t i = threadIdx.x; i < DEALTAB_ENTRIES; i += blockDim.x) s_deal[i] = g_dealtab[i];
    stage_tables(smem, g_tab, reinterpret_cast<uint64_t*>(smem + kSmemTab));   // includes a __syncthreads
    uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem);
    sb += (uint32_t)(rollouts >> 63);        // + 0, opaque to ptxas: keeps the shared-window base in a register (see k_equity_mc_smem)
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t wcells = sb + kSmemTabOnly + kDealTabBytes + warp * kPoolWarpBytes;
    const uint32_t ring = wcells + kPoolRows * kPoolRowBytes;
    sts_u8(ring + lane, 32u + lane);         // free segment: columns 32..63; lane holds column `lane`
    __syncthreads();
    const HashParams hp = c_hp;
    const SharedTab tab{sb};
    const SharedDealTab dt{sb + kSmemTabOnly, c_hp.two * c_hp.two * c_hp.two};
    const uint32_t k_lo = (uint32_t)seed, k_hi = (uint32_t)(seed >> 32);
    uint32_t col = lane;
    uint32_t rf = 0, rr = 32, rt = 32;       // ring cursors: free [rf, rr), ready [rr, rt)  (mod 64)

    const uint64_t chunks_per_state = (rollouts + chunk - 1) / chunk;
    const uint64_t items = (uint64_t)n_states * chunks_per_state;
    for (;;) {
        uint32_t nxt = 0;
        if (lane == 0) nxt = atomicAdd(counters, 1u);
        const uint64_t item = __shfl_sync(0xFFFFFFFFu, nxt, 0);
        if (item >= items) break;
        const uint64_t ch = item / (uint64_t)n_states, s = item - ch * (uint64_t)n_states;      // interleaved over the states (see equity_mc_body)
        const ulonglong2 st = states[s];
        const uint64_t known = st.x & CARD_BITS, hero = st.y & CARD_BITS;
        const int over = (int)(st.y >> 61);
        const int n_opp = over ? over : n_opp_default;
        const int nbm = 5 - __popcll(known & ~hero), nbp = nbm >> 1;
        const bool want_single = (nbm & 1) != 0;
        const uint32_t n_pairs = (uint32_t)(nbp + n_opp);
        const uint64_t b = ch * chunk;
        const uint32_t limit = (uint32_t)((b + chunk > rollouts ? rollouts : b + chunk) - b);
        const uint64_t base_sample = sample_offset + b;
        const uint32_t stream = state_base + (uint32_t)s;
        const uint32_t known_lo = (uint32_t)known, known_hi = (uint32_t)(known >> 32);
        const uint32_t kb_lo = (uint32_t)(known & ~hero), kb_hi = (uint32_t)((known & ~hero) >> 32);

        // river states: the board is complete, so its flush context and the hero's rank are the same for every rollout
        BoardFlush bf_river = board_flush(kb_lo, kb_hi);
        const uint32_t hr_river = nbm == 0 ? eval7_board(kb_lo, kb_hi, bf_river, hero, tab, hp) : 0u;
        // evaluation of n ready deals: lane j takes the j-th ready column
        uint32_t acc = 0;                    // three 10-bit counters (win, tie, loss): a lane evaluates <= chunk / 32 + 1 <= 257 deals per item
        auto evaluate = [&](uint32_t first, uint32_t n) {
            if (lane < n) {
                const uint32_t cb = wcells + 2u * lds_u8(ring + ((first + lane) & 63u));
                uint32_t blo = kb_lo, bhi = kb_hi;
                if (want_single) {
                    const uint64_t m = dt.mask(lds_u16v(cb + kPoolSingleRow * kPoolRowBytes));
                    blo |= (uint32_t)m; bhi |= (uint32_t)(m >> 32);
                }
                for (int k = 0; k < nbp; k++) {
                    const uint64_t m = dt.mask(lds_u16v(cb + (uint32_t)k * kPoolRowBytes));
                    blo |= (uint32_t)m; bhi |= (uint32_t)(m >> 32);
                }
                BoardFlush bf = bf_river;
                uint32_t hr = hr_river;
                if (nbm != 0) {                          // uniform per item
                    bf = board_flush(blo, bhi);
                    hr = eval7_board(blo, bhi, bf, hero, tab, hp);
                }
                uint32_t best = 0xFFFFu;
                for (int o = 0; o < n_opp; o++) {
                    const uint32_t r = eval7_board(blo, bhi, bf, dt.mask(lds_u16v(cb + (uint32_t)(nbp + o) * kPoolRowBytes)), tab, hp);
                    best = r < best ? r : best;
                }
                acc += hr < best ? 1u : (hr == best ? (1u << 10) : (1u << 20));
            }
            __syncwarp();
        };

        // per-lane dealing state
        uint32_t rel = lane;                 // sample index within the item
        bool live = rel < limit;
        uint32_t blk = 0;
        uint32_t used_lo = live ? known_lo : 0xFFFFFFFFu, used_hi = live ? known_hi : 0xFFFFFFFFu;   // all ones: nothing is ever accepted
        uint32_t cur = wcells + 2u * col;
        uint32_t end = cur + n_pairs * kPoolRowBytes;
#ifdef PHE_POOL_SKIP_READS
        if (!live) cur = end;
#endif

        while (__any_sync(0xFFFFFFFFu, live)) {
            const uint64_t sample = base_sample + rel;
            uint32_t rnd[4];
            philox4x32_10((uint32_t)sample, (uint32_t)(sample >> 32), stream, blk, k_lo, k_hi, rnd);
            const bool first = blk == 0;
            if (want_single && first && live) {
                // the single card: draws 0 and 1 of block 0 are reserved for it; overflow blocks 0x80000000 + t in the ~1 % case
                bool have = false;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t r = (rnd[0] >> (16 * h)) & 0xFFFFu;
                    const uint32_t c = N_PAIRS + mod52_u16(r);
                    const uint64_t m = dt.mask(c);
                    if (!have && r < DEAL_SINGLE_LIMIT && !(((uint32_t)m & used_lo) | ((uint32_t)(m >> 32) & used_hi))) {
                        have = true; used_lo |= (uint32_t)m; used_hi |= (uint32_t)(m >> 32);
                        sts_u16(wcells + 2u * col + kPoolSingleRow * kPoolRowBytes, c);
                    }
                }
                for (uint32_t t = 0; !have && t < 64u; t++) {
                    uint32_t ex[4];
                    philox4x32_10((uint32_t)sample, (uint32_t)(sample >> 32), stream, 0x80000000u + t, k_lo, k_hi, ex);
                    for (int d = 0; d < 8 && !have; d++) {
                        const uint32_t r = (ex[d >> 1] >> (16 * (d & 1))) & 0xFFFFu;
                        const uint32_t c = N_PAIRS + mod52_u16(r);
                        const uint64_t m = dt.mask(c);
                        if (r < DEAL_SINGLE_LIMIT && !(((uint32_t)m & used_lo) | ((uint32_t)(m >> 32) & used_hi))) {
                            have = true; used_lo |= (uint32_t)m; used_hi |= (uint32_t)(m >> 32);
                            sts_u16(wcells + 2u * col + kPoolSingleRow * kPoolRowBytes, c);
                        }
                    }
                }
            }
            const bool skip01 = want_single && first;      // the reserved word of block 0
#pragma unroll
            for (int wi = 0; wi < 4; wi++) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t r = (rnd[wi] >> (16 * h)) & 0xFFFFu;
                    const uint32_t q = mod1326_u16(r);
#ifndef PHE_POOL_SKIP_READS
                    const uint64_t m = dt.mask(q);
                    const uint32_t m_lo = (uint32_t)m, m_hi = (uint32_t)(m >> 32);
                    bool ok = r < DEAL_PAIR_LIMIT && !((m_lo & used_lo) | (m_hi & used_hi));
                    if (wi == 0) ok = ok && !skip01;
                    if (ok) {
                        used_lo |= m_lo; used_hi |= m_hi;
                        sts_u16(cur, q);
                        cur = cur + kPoolRowBytes < end ? cur + kPoolRowBytes : end;     // full: later accepts of this round land in the spare row
                    }
#else
                    // -DPHE_POOL_SKIP_READS: a lane whose slots are full (or that has no sample left: cur is parked on end) skips
                    // the table read.  Fewer shared-memory wavefronts (2.37e9 against 2.68e9 per 4.1e8 rollouts) but more
                    // instructions (9.61e9 against 8.90e9): 3.84e10 against 3.91e10 rollouts/s on the B200 -- kept for the record
                    bool need = cur != end;
                    if (wi == 0) need = need && !skip01;
                    uint32_t m_lo = 0xFFFFFFFFu, m_hi = 0xFFFFFFFFu;
                    if (need) {
                        const uint64_t m = dt.mask(q);
                        m_lo = (uint32_t)m; m_hi = (uint32_t)(m >> 32);
                    }
                    if (r < DEAL_PAIR_LIMIT && !((m_lo & used_lo) | (m_hi & used_hi))) {
                        used_lo |= m_lo; used_hi |= m_hi;
                        sts_u16(cur, q);
                        cur += kPoolRowBytes;
                    }
#endif
                }
            }
            const bool done = live && (cur == end || blk == 63u);      // 64 blocks are never reached by a valid state; bounds garbage
            const uint32_t dm = __ballot_sync(0xFFFFFFFFu, done);
            blk++;
            if (dm) {
                const uint32_t rank = __popc(dm & lt_mask), cnt = __popc(dm);
                if (done) sts_u8(ring + ((rt + rank) & 63u), col);
                rt += cnt;
                __syncwarp();
                while (rt - rr >= 32u) { evaluate(rr, 32u); rr += 32u; }
                if (done) {
                    col = lds_u8(ring + ((rf + rank) & 63u));
                    rel += 32u;
                    live = rel < limit;
                    blk = 0;
                    used_lo = live ? known_lo : 0xFFFFFFFFu;
                    used_hi = live ? known_hi : 0xFFFFFFFFu;
                    cur = wcells + 2u * col;
                    end = cur + n_pairs * kPoolRowBytes;
#ifdef PHE_POOL_SKIP_READS
                    if (!live) cur = end;
#endif
                }
                rf += cnt;
                __syncwarp();
            }
        }
        if (rt != rr) { evaluate(rr, rt - rr); rr = rt; }
        const uint32_t w = __reduce_add_sync(0xFFFFFFFFu, acc & 1023u), t = __reduce_add_sync(0xFFFFFFFFu, (acc >> 10) & 1023u),
                       l = __reduce_add_sync(0xFFFFFFFFu, acc >> 20);
        if (lane == 0) {
            if (w) atomicAdd(wtl + 3 * s + 0, (unsigned long long)w);
            if (t) atomicAdd(wtl + 3 * s + 1, (unsigned long long)t);
            if (l) atomicAdd(wtl + 3 * s + 2, (unsigned long long)l);
        }
    }
    // completion ticket: the last CTA re-arms the {item counter, ticket} pair for the next launch that draws it
    __shared__ uint32_t s_last;
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(counters + 1, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        counters[0] = 0;
        counters[1] = 0;
    }
}

#endif   // PHE_MC_POOL

__global__ void __launch_bounds__(256)
k_equity_mc_gmem(const uint16_t* __restrict__ g_tab, const uint64_t* __restrict__ g_dealtab, const ulonglong2* __restrict__ states,
                 int64_t n_states, int n_opp, uint64_t rollouts, uint64_t sample_offset, uint64_t seed, uint32_t state_base,
                 unsigned long long* __restrict__ wtl, uint32_t chunk)
{
    __shared__ uint32_t s_slots[kMcSlotRows * 256];
    const GlobalTab tab{g_tab};
    const GlobalDealTab dt{g_dealtab};
    const SmemSlots slots{(uint32_t)__cvta_generic_to_shared(s_slots) + 4u * threadIdx.x, 4u * blockDim.x, 0ull};
    equity_mc_body(tab, dt, states, n_states, n_opp, rollouts, sample_offset, seed_key(seed), state_base, wtl, chunk, slots, nullptr);
}

// ------------------------------------------------------------------------------------------------
// K3b: exact-plan equity (simulate_processes).  One thread per enumeration leaf.
// ------------------------------------------------------------------------------------------------
template <class Tab>
__device__ __forceinline__ void equity_plan_body(const Tab& tab, const GlobalDealTab dt, const uint8_t* __restrict__ exist, int64_t n_tasks,
                                                 const PlanDesc& pd, uint64_t seed, uint32_t stream_base,
                                                 unsigned long long* __restrict__ wtl)
{
    const HashParams hp = c_hp;
    const Binom C{c_binom};
    const uint64_t total = (uint64_t)n_tasks * pd.n_leaves;
    const uint64_t nthr = (uint64_t)gridDim.x * blockDim.x;
    const int base = pd.n_exist + pd.n_enum;
    for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += nthr) {
        const uint64_t task = idx / pd.n_leaves, leaf = idx - task * pd.n_leaves;
        uint8_t cards[9];
        uint64_t known = 0;
        for (int i = 0; i < pd.n_exist; i++) {
            cards[i] = __ldg(exist + task * 8 + i);
            known |= id_bit(cards[i]);
        }
        plan_unrank_leaf(pd, leaf, C, cards);
        uint32_t w = 0, t = 0, l = 0;
        if (pd.rnd_k == 0) {
            const int res = plan_leaf_result(tab, hp, cards);
            w = (res == 0); t = (res == 1); l = (res == 2);
        } else {
            // random step (env/workflow.py:181-195): rnd_k cards = the board cards still missing + the villain's two.
            // Slots: [single if the missing board count is odd][board pairs][villain pair]; card sets are used directly
            // and the hero is ranked once per leaf when the enumeration has already completed the board.
            for (int i = pd.n_exist; i < base; i++) known |= id_bit(cards[i]);
            const uint64_t hero = id_bit(cards[0]) | id_bit(cards[1]);
            const uint64_t board0 = known & ~hero;
            const int nbm = pd.rnd_k - 2, nbp = nbm >> 1;
            const uint32_t hr0 = nbm == 0 ? eval7(board0 | hero, tab, hp) : 0u;
            LocalSlots sl;
            for (int j = 0; j < pd.rnd_n; j++) {
                deal_slots(dt, known, seed, leaf * (uint64_t)pd.rnd_n + (uint64_t)j, stream_base + (uint32_t)task, (nbm & 1) != 0, nbp + 1, sl);
                uint64_t board = board0;
                if (nbm & 1) board |= sl.single;
                for (int k = 0; k < nbp; k++) board |= sl.pair[k];
                const uint32_t hr = nbm == 0 ? hr0 : eval7(board | hero, tab, hp);
                const uint32_t vr = eval7(board | sl.pair[nbp], tab, hp);
                w += hr < vr; t += hr == vr; l += hr > vr;
            }
        }
        // aggregate lanes that work on the same task before touching global memory
        const unsigned act = __activemask();
        const unsigned same = __match_any_sync(act, (unsigned long long)task);
        const int leader = __ffs(same) - 1;
        for (int off = 16; off > 0; off >>= 1) {
            const uint32_t w2 = __shfl_down_sync(act, w, off), t2 = __shfl_down_sync(act, t, off), l2 = __shfl_down_sync(act, l, off);
            const int src = (threadIdx.x & 31) + off;
            if (src < 32 && (same >> src & 1)) { w += w2; t += t2; l += l2; }
        }
        if ((int)(threadIdx.x & 31) == leader) {
            if (w) atomicAdd(wtl + 3 * task + 0, (unsigned long long)w);
            if (t) atomicAdd(wtl + 3 * task + 1, (unsigned long long)t);
            if (l) atomicAdd(wtl + 3 * task + 2, (unsigned long long)l);
        }
    }
}

__global__ void __launch_bounds__(kCtaThreads, 1)
k_equity_plan_smem(const uint16_t* __restrict__ g_tab, const uint64_t* __restrict__ g_dealtab, const uint8_t* __restrict__ exist, int64_t n_tasks,
                   PlanDesc pd, uint64_t seed, uint32_t stream_base, unsigned long long* __restrict__ wtl)
{
    extern __shared__ __align__(128) unsigned char smem[];
    stage_tables(smem, g_tab, reinterpret_cast<uint64_t*>(smem + kSmemTab));
    const SharedTab tab{(uint32_t)__cvta_generic_to_shared(smem)};
    equity_plan_body(tab, GlobalDealTab{g_dealtab}, exist, n_tasks, pd, seed, stream_base, wtl);
}

__global__ void __launch_bounds__(256)
k_equity_plan_gmem(const uint16_t* __restrict__ g_tab, const uint64_t* __restrict__ g_dealtab, const uint8_t* __restrict__ exist, int64_t n_tasks,
                   PlanDesc pd, uint64_t seed, uint32_t stream_base, unsigned long long* __restrict__ wtl)
{
    const GlobalTab tab{g_tab};
    equity_plan_body(tab, GlobalDealTab{g_dealtab}, exist, n_tasks, pd, seed, stream_base, wtl);
}

// ------------------------------------------------------------------------------------------------
// K3c: exhaustive plans ending in the villain's two cards (default rounds 2 and 3).  One warp per (task, board
// completion): the prefix leaf is unranked once, the hero is ranked once, and the 32 lanes stride over the pair
// table, skipping hands that touch a known card.  Visiting order differs from the reference's, counts do not.
// ------------------------------------------------------------------------------------------------
template <class Tab>
__device__ __forceinline__ void equity_plan_pairs_body(const Tab& tab, const uint64_t* __restrict__ pairmask, const uint8_t* __restrict__ exist,
                                                       int64_t n_tasks, const PlanDesc& pd, unsigned long long* __restrict__ wtl)
{
    const HashParams hp = c_hp;
    const Binom C{c_binom};
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warps_per_cta = blockDim.x >> 5;
    const uint64_t nwarps = (uint64_t)gridDim.x * warps_per_cta;
    const uint64_t items = (uint64_t)n_tasks * pd.n_prefix;
    PlanDesc prefix = pd;           // the plan without its last segment
    prefix.n_seg = pd.n_seg - 1;
    prefix.n_enum = pd.n_enum - 2;
    for (uint64_t item = (uint64_t)blockIdx.x * warps_per_cta + (threadIdx.x >> 5); item < items; item += nwarps) {
        const uint64_t task = item / pd.n_prefix, leaf = item - task * pd.n_prefix;
        uint8_t cards[9];
        for (int i = 0; i < pd.n_exist; i++) cards[i] = __ldg(exist + task * 8 + i);
        plan_unrank_leaf(prefix, leaf, C, cards);
        const uint64_t hero = id_bit(cards[0]) | id_bit(cards[1]);
        const uint64_t board = id_bit(cards[2]) | id_bit(cards[3]) | id_bit(cards[4]) | id_bit(cards[5]) | id_bit(cards[6]);
        const uint64_t known = hero | board;
        const uint32_t hr = eval7(board | hero, tab, hp);
        uint32_t w = 0, t = 0, l = 0;
        for (uint32_t q = lane; q < N_PAIRS; q += 32) {
            const uint64_t pm = pairmask[q];
            if (pm & known) continue;
            const uint32_t vr = eval7(board | pm, tab, hp);
            w += hr < vr; t += hr == vr; l += hr > vr;
        }
        w = __reduce_add_sync(0xFFFFFFFFu, w);
        t = __reduce_add_sync(0xFFFFFFFFu, t);
        l = __reduce_add_sync(0xFFFFFFFFu, l);
        if (lane == 0) {
            if (w) atomicAdd(wtl + 3 * task + 0, (unsigned long long)w);
            if (t) atomicAdd(wtl + 3 * task + 1, (unsigned long long)t);
            if (l) atomicAdd(wtl + 3 * task + 2, (unsigned long long)l);
        }
    }
}

__global__ void __launch_bounds__(kCtaThreads, 1)
k_equity_plan_pairs_smem(const uint16_t* __restrict__ g_tab, const uint64_t* __restrict__ g_pairmask, const uint8_t* __restrict__ exist,
                         int64_t n_tasks, PlanDesc pd, unsigned long long* __restrict__ wtl)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* s_pairs = reinterpret_cast<uint64_t*>(smem + kSmemTabOnly);
    for (uint32_t i = threadIdx.x; i < N_PAIRS; i += blockDim.x) s_pairs[i] = g_pairmask[i];
    stage_tables(smem, g_tab, reinterpret_cast<uint64_t*>(smem + kSmemTab));   // includes a __syncthreads
    __syncthreads();
    const SharedTab tab{(uint32_t)__cvta_generic_to_shared(smem)};
    equity_plan_pairs_body(tab, s_pairs, exist, n_tasks, pd, wtl);
}

__global__ void __launch_bounds__(256)
k_equity_plan_pairs_gmem(const uint16_t* __restrict__ g_tab, const uint64_t* __restrict__ g_pairmask, const uint8_t* __restrict__ exist,
                         int64_t n_tasks, PlanDesc pd, unsigned long long* __restrict__ wtl)
{
    const GlobalTab tab{g_tab};
    equity_plan_pairs_body(tab, g_pairmask, exist, n_tasks, pd, wtl);
}

// ------------------------------------------------------------------------------------------------
// K4: showdown (get_winner_set_v2, env/game.py:663-681).  One thread per game.
// ------------------------------------------------------------------------------------------------
template <class Tab>
__device__ __forceinline__ void showdown_body(const Tab& tab, const uint64_t* __restrict__ boards, const uint64_t* __restrict__ holes,
                                              const uint32_t* __restrict__ live, int64_t n_games, int P,
                                              uint32_t* __restrict__ winners, uint16_t* __restrict__ ranks)
{
    const HashParams hp = c_hp;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_games; g += nthr) {
        const uint32_t lv = live ? live[g] : ((1u << P) - 1u);
        const bool solo = __popc(lv) == 1;   // env/game.py:664-665: a lone player wins unevaluated
        const uint64_t board = boards[g];
        uint32_t best = 0xFFFFu, win = 0;
        for (int p = 0; p < P; p++) {
            uint32_t r = 0;
            if ((lv >> p & 1) && !solo) {
                r = eval7(board | holes[g * P + p], tab, hp);
                if (r <= best) {
                    if (r < best) { best = r; win = 0; }
                    win |= 1u << p;
                }
            }
            if (ranks) ranks[g * P + p] = (uint16_t)r;
        }
        winners[g] = solo ? lv : win;
    }
}

__global__ void __launch_bounds__(kCtaThreads, 1)
k_showdown_smem(const uint16_t* __restrict__ g_tab, const uint64_t* __restrict__ boards, const uint64_t* __restrict__ holes,
                const uint32_t* __restrict__ live, int64_t n_games, int P, uint32_t* __restrict__ winners, uint16_t* __restrict__ ranks)
{
    extern __shared__ __align__(128) unsigned char smem[];
    stage_tables(smem, g_tab, reinterpret_cast<uint64_t*>(smem + kSmemTab));
    const SharedTab tab{(uint32_t)__cvta_generic_to_shared(smem)};
    showdown_body(tab, boards, holes, live, n_games, P, winners, ranks);
}

__global__ void __launch_bounds__(256)
k_showdown_gmem(const uint16_t* __restrict__ g_tab, const uint64_t* __restrict__ boards, const uint64_t* __restrict__ holes,
                const uint32_t* __restrict__ live, int64_t n_games, int P, uint32_t* __restrict__ winners, uint16_t* __restrict__ ranks)
{
    const GlobalTab tab{g_tab};
    showdown_body(tab, boards, holes, live, n_games, P, winners, ranks);
}

// ------------------------------------------------------------------------------------------------
// K5: determinisation deals (GameEnv.new_random, env/game.py:119-163)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_deal(const uint64_t* __restrict__ g_dealtab, const uint64_t* __restrict__ known, int64_t n, int need, uint64_t seed, uint64_t sample,
       uint32_t stream_base, int by_sample, uint8_t* __restrict__ cards_out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t ids[MAX_DEAL];
    // row i is either stream (stream_base + i) at one sample index, or sample (sample + i) of one stream
    deal_ids(GlobalDealTab{g_dealtab}, known[i] & CARD_BITS, seed, by_sample ? sample + (uint64_t)i : sample,
             by_sample ? stream_base : stream_base + (uint32_t)i, need, ids);
    for (int k = 0; k < need; k++) cards_out[i * need + k] = ids[k];
}

// ------------------------------------------------------------------------------------------------
// K6: terminal settlement (finish_game / _share_value).  One thread per game: P evaluations, then integer logic.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_settle(const uint16_t* __restrict__ g_tab, const uint64_t* __restrict__ boards, const uint64_t* __restrict__ holes,
         const int8_t* __restrict__ actions, const int32_t* __restrict__ values, const int32_t* __restrict__ n_rounds,
         const int32_t* __restrict__ button, const uint32_t* __restrict__ onboard, int64_t n_games, int P, int R, int small_blind,
         int32_t* __restrict__ won_out, int8_t* __restrict__ result_out, int8_t* __restrict__ card_out, uint32_t* __restrict__ paid_out)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_games) return;
    const GlobalTab tab{g_tab};
    const HashParams hp = c_hp;
    const uint64_t board = boards[g];
    uint32_t ranks[SETTLE_MAX_PLAYERS];
    for (int p = 0; p < P; p++) ranks[p] = eval7(board | holes[g * P + p], tab, hp);
    int8_t act[SETTLE_MAX_ROUNDS * SETTLE_MAX_PLAYERS];
    int32_t val[SETTLE_MAX_ROUNDS * SETTLE_MAX_PLAYERS];
    int nr = n_rounds[g];
    nr = nr < 0 ? 0 : (nr > R ? R : nr);
    for (int i = 0; i < nr * P; i++) {
        act[i] = actions[g * R * P + i];
        val[i] = values[g * R * P + i];
    }
    int32_t won[SETTLE_MAX_PLAYERS];
    int8_t res[SETTLE_MAX_PLAYERS], card[SETTLE_MAX_PLAYERS];
    uint32_t paid;
    int bt = button[g];
    bt = bt < 0 || bt >= P ? 0 : bt;      // caller memory is not trusted with an array index (the host layer rejects such rows)
    settle_game(P, nr, bt, small_blind, onboard[g], ranks, act, val, won, res, card, &paid);
    for (int p = 0; p < P; p++) {
        won_out[g * P + p] = won[p];
        result_out[g * P + p] = res[p];
        card_out[g * P + p] = card[p];
    }
    if (paid_out) paid_out[g] = paid;
}

// ------------------------------------------------------------------------------------------------
// host side: per-device context, error plumbing, launches
// ------------------------------------------------------------------------------------------------
namespace {

constexpr int kMaxDevices = 64;
constexpr uint32_t kCounterRing = 256;
constexpr uint32_t kEnumRing = 64;      // per-launch u32 histograms of the enumeration (30 KB each), one per in-flight launch

struct DevCtx {
    std::once_flag once;
    int status = PHE_ENODEVICE;
    uint16_t* d_image = nullptr;
    unsigned char* d_enum = nullptr;
    uint64_t* d_trimask = nullptr;
    uint32_t* d_units = nullptr;          // work-unit boundaries of the enumeration (build_enum_units)
    uint64_t* d_unit_starts = nullptr;    // packed block / triple at which every unit starts
    EnumBlockRec* d_blocks = nullptr;     // the 211,876 enumeration blocks as the kernel reads them (3.4 MB, L2 resident)
    uint64_t* d_pairmask = nullptr;
    unsigned int* d_counters = nullptr;   // ring of {tile counter, completion ticket} pairs, one per in-flight enumeration launch
    uint32_t* d_h32 = nullptr;            // ring of u32[kHistBinsPad] flush targets of k_enum7 (zero at rest, cleared by the kernel's last CTA)
    std::atomic<uint32_t> next_counter{0};
    int sm_count = 0;
};

DevCtx g_ctx[kMaxDevices];
std::atomic<uint64_t> g_launches{0};
thread_local char t_cuda_err[256] = "";

std::once_flag g_image_once;
std::vector<uint16_t> g_image;
std::vector<unsigned char> g_enum_image;
std::vector<uint64_t> g_trimask;
std::vector<uint32_t> g_enum_units;
std::vector<uint64_t> g_enum_unit_starts;
std::vector<EnumBlockRec> g_enum_blocks;
constexpr uint32_t kEnumUnitHands = 512, kEnumUnitBlocks = 4;
HashParams g_hp;
uint32_t g_binom[53 * 8];
bool g_image_ok = false;

int cuda_fail(cudaError_t e, const char* what)
{
    snprintf(t_cuda_err, sizeof t_cuda_err, "%s: %s", what, cudaGetErrorString(e));
    return PHE_ECUDA;
}

#define PHE_CUDA(call)                                      \
    do {                                                    \
        cudaError_t e__ = (call);                           \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

void init_device(int device)
{
    DevCtx& cx = g_ctx[device];
    std::call_once(g_image_once, [] {
        g_image.resize(TAB_U16);
        g_image_ok = build_table_image(g_image.data(), &g_hp);
        g_enum_image.resize(ENUM_IMG_BYTES);
        g_trimask.resize(N_TRIPLES);
        g_image_ok = g_image_ok && build_enum_image(g_enum_image.data(), g_trimask.data());
#ifndef PHE_ENUM_UNIT_ALIGN
#define PHE_ENUM_UNIT_ALIGN 128
#endif
        build_enum_units(g_enum_units, g_enum_unit_starts, kEnumUnitHands, kEnumUnitBlocks, PHE_ENUM_UNIT_ALIGN);
        build_enum_blocks(g_enum_blocks);
        build_binom(g_binom);
    });
    if (!g_image_ok) { cx.status = PHE_EINVAL; return; }
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { cx.status = cuda_fail(e, "cudaGetDeviceProperties"); return; }
    if (prop.major != 10) {
        snprintf(t_cuda_err, sizeof t_cuda_err, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        cx.status = PHE_ENODEVICE;
        return;
    }
    cx.sm_count = prop.multiProcessorCount;
    auto chk = [&](cudaError_t err, const char* what) {
        if (err != cudaSuccess && cx.status == PHE_OK) cx.status = cuda_fail(err, what);
    };
    cx.status = PHE_OK;
    chk(cudaSetDevice(device), "cudaSetDevice");
    chk(cudaMalloc(&cx.d_image, TAB_BYTES), "cudaMalloc(tables)");
    chk(cudaMemcpy(cx.d_image, g_image.data(), TAB_BYTES, cudaMemcpyHostToDevice), "cudaMemcpy(tables)");
    chk(cudaMalloc(&cx.d_enum, ENUM_IMG_BYTES), "cudaMalloc(enum tables)");
    chk(cudaMemcpy(cx.d_enum, g_enum_image.data(), ENUM_IMG_BYTES, cudaMemcpyHostToDevice), "cudaMemcpy(enum tables)");
    chk(cudaMalloc(&cx.d_trimask, TRIMASK_BYTES), "cudaMalloc(trimask)");
    chk(cudaMemcpy(cx.d_trimask, g_trimask.data(), TRIMASK_BYTES, cudaMemcpyHostToDevice), "cudaMemcpy(trimask)");
    chk(cudaMalloc(&cx.d_units, g_enum_units.size() * sizeof(uint32_t)), "cudaMalloc(enum units)");
    chk(cudaMemcpy(cx.d_units, g_enum_units.data(), g_enum_units.size() * sizeof(uint32_t), cudaMemcpyHostToDevice), "cudaMemcpy(enum units)");
    chk(cudaMalloc(&cx.d_unit_starts, g_enum_unit_starts.size() * sizeof(uint64_t)), "cudaMalloc(enum unit starts)");
    chk(cudaMemcpy(cx.d_unit_starts, g_enum_unit_starts.data(), g_enum_unit_starts.size() * sizeof(uint64_t), cudaMemcpyHostToDevice), "cudaMemcpy(enum unit starts)");
    chk(cudaMalloc(&cx.d_blocks, g_enum_blocks.size() * sizeof(EnumBlockRec)), "cudaMalloc(enum blocks)");
    chk(cudaMemcpy(cx.d_blocks, g_enum_blocks.data(), g_enum_blocks.size() * sizeof(EnumBlockRec), cudaMemcpyHostToDevice), "cudaMemcpy(enum blocks)");
    chk(cudaMalloc(&cx.d_counters, 2 * kCounterRing * sizeof(unsigned int)), "cudaMalloc(counters)");
    chk(cudaMemset(cx.d_counters, 0, 2 * kCounterRing * sizeof(unsigned int)), "cudaMemset(counters)");
    chk(cudaMalloc(&cx.d_h32, (size_t)kEnumRing * kHistBinsPad * sizeof(uint32_t)), "cudaMalloc(h32)");
    chk(cudaMemset(cx.d_h32, 0, (size_t)kEnumRing * kHistBinsPad * sizeof(uint32_t)), "cudaMemset(h32)");
    {
        std::vector<uint64_t> pm(DEALTAB_ENTRIES);   // pairs, then singles; the plan fast path reads the pair prefix
        build_dealtab(pm.data());
        chk(cudaMalloc(&cx.d_pairmask, DEALTAB_ENTRIES * 8), "cudaMalloc(dealtab)");
        chk(cudaMemcpy(cx.d_pairmask, pm.data(), DEALTAB_ENTRIES * 8, cudaMemcpyHostToDevice), "cudaMemcpy(dealtab)");
    }
    chk(cudaFuncSetAttribute(k_equity_plan_pairs_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemPairs), "attr equity_plan_pairs");
    chk(cudaMemcpyToSymbol(c_hp, &g_hp, sizeof g_hp), "cudaMemcpyToSymbol(c_hp)");
    chk(cudaMemcpyToSymbol(c_binom, g_binom, sizeof g_binom), "cudaMemcpyToSymbol(c_binom)");
    chk(cudaFuncSetAttribute(k_eval7_smem<PHE_LAYOUT_IDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemEvalIds), "attr eval7 ids");
    chk(cudaFuncSetAttribute(k_eval7_smem<PHE_LAYOUT_MASKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTabOnly), "attr eval7 masks");
    chk(cudaFuncSetAttribute(k_enum7, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemEnum), "attr enum7");
    chk(cudaFuncSetAttribute(k_equity_mc_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMc), "attr equity_mc");
#ifdef PHE_MC_POOL
    chk(cudaFuncSetAttribute(k_equity_mc_pool, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMcPool), "attr equity_mc_pool");
#endif
    chk(cudaFuncSetAttribute(k_equity_plan_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTabOnly), "attr equity_plan");
    chk(cudaFuncSetAttribute(k_showdown_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTabOnly), "attr showdown");
}

// context of the current device, initialising it on first use
int current_ctx(DevCtx** out)
{
    int dev = 0;
    PHE_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return PHE_ENODEVICE;
    DevCtx& cx = g_ctx[dev];
    std::call_once(cx.once, init_device, dev);
    if (cx.status != PHE_OK) return cx.status;
    *out = &cx;
    return PHE_OK;
}

int after_launch(const char* what)
{
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, what);
    return PHE_OK;
}

// growable device scratch for the *_host entry points (one per host thread)
struct Scratch {
    void* p = nullptr;
    size_t cap = 0;
    int dev = -1;
    int ensure(size_t bytes)
    {
        int d = 0;
        PHE_CUDA(cudaGetDevice(&d));
        if (p && d == dev && cap >= bytes) return PHE_OK;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes < 4096 ? 4096 : bytes;
        PHE_CUDA(cudaMalloc(&p, want));
        cap = want; dev = d;
        return PHE_OK;
    }
};
thread_local Scratch t_in, t_out;

// Small host-buffer calls of the store-only kernels (the drop-in single-hand / single-game paths; kernels that accumulate
// with atomics keep their outputs in device memory) skip the staging copies: the
// arguments are memcpy'd into a per-thread pinned buffer that the kernel reads and writes through UVA, so a call is one
// launch and one stream synchronisation.
constexpr size_t kPinnedBytes = 64 * 1024;
struct Pinned {
    unsigned char* p = nullptr;
    int get(unsigned char** out)
    {
        if (!p) PHE_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&p), kPinnedBytes, cudaHostAllocDefault));
        *out = p;
        return PHE_OK;
    }
};
thread_local Pinned t_pin;

// The *_host entry points run on a non-blocking stream owned by the calling host thread (one per thread and device) and
// synchronise that stream only: game-loop threads neither serialise on the legacy NULL stream nor implicitly wait for the
// predictor's or anybody else's blocking streams (SURVEY 8b "Threading").
struct HostStream {
    cudaStream_t s = nullptr;
    int dev = -1;
    int get(cudaStream_t* out)
    {
        int d = 0;
        PHE_CUDA(cudaGetDevice(&d));
        if (!s || d != dev) {
            if (s) cudaStreamDestroy(s);
            s = nullptr;
            PHE_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
            dev = d;
        }
        *out = s;
        return PHE_OK;
    }
};
thread_local HostStream t_hstream;

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

// launch accounting for the other translation units of the library (phe_nn.cu)
void phe_internal_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

extern "C" {

const char* phe_strerror(int status)
{
    switch (status) {
        case PHE_OK: return "ok";
        case PHE_EINVAL: return "invalid argument";
        case PHE_ECUDA: return "CUDA error";
        case PHE_ENODEVICE: return "no usable sm_100 device";
        case PHE_EPLAN: return "equity plan not expressible";
        case PHE_ENOMEM: return "out of memory";
        default: return "unknown status";
    }
}

const char* phe_last_cuda_error(void) { return t_cuda_err; }
int phe_version(void) { return 100; }
uint64_t phe_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int phe_init(int device)
{
    if (device < 0 || device >= kMaxDevices) return PHE_EINVAL;
    int n = 0;
    PHE_CUDA(cudaGetDeviceCount(&n));
    if (device >= n) return PHE_ENODEVICE;
    PHE_CUDA(cudaSetDevice(device));
    DevCtx* cx = nullptr;
    return current_ctx(&cx);
}

// ---- eval7 -------------------------------------------------------------------------------------
static constexpr int64_t kSmallBatch = 1 << 16;   // below this the table image is read through L1/L2

int phe_eval7(const void* hands, int layout, int64_t n, uint16_t* out, void* stream)
{
    if (n < 0 || (layout != PHE_LAYOUT_IDS && layout != PHE_LAYOUT_MASKS)) return PHE_EINVAL;
    if (n == 0) return PHE_OK;
    if (!hands || !out) return PHE_EINVAL;
    DevCtx* cx = nullptr;
    int rc = current_ctx(&cx);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n < kSmallBatch) {
        const unsigned blocks = (unsigned)((n + 255) / 256);
        if (layout == PHE_LAYOUT_IDS) k_eval7_gmem<PHE_LAYOUT_IDS><<<blocks, 256, 0, st>>>(cx->d_image, hands, n, out);
        else k_eval7_gmem<PHE_LAYOUT_MASKS><<<blocks, 256, 0, st>>>(cx->d_image, hands, n, out);
    } else {
        // programmatic dependent launch (see k_eval7_smem): overlaps this grid's table staging with the previous grid's tail
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)cx->sm_count);
        cfg.blockDim = dim3(kCtaThreads);
        cfg.dynamicSmemBytes = layout == PHE_LAYOUT_IDS ? kSmemEvalIds : kSmemTabOnly;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        const uint16_t* img = cx->d_image;
        cudaError_t e = layout == PHE_LAYOUT_IDS ? cudaLaunchKernelEx(&cfg, k_eval7_smem<PHE_LAYOUT_IDS>, img, hands, n, out)
                                                 : cudaLaunchKernelEx(&cfg, k_eval7_smem<PHE_LAYOUT_MASKS>, img, hands, n, out);
        if (e != cudaSuccess) return cuda_fail(e, "eval7 launch");
    }
    return after_launch("eval7 launch");
}

int phe_eval7_host(const void* hands, int layout, int64_t n, uint16_t* out)
{
    if (n < 0 || (layout != PHE_LAYOUT_IDS && layout != PHE_LAYOUT_MASKS)) return PHE_EINVAL;
    if (n == 0) return PHE_OK;
    if (!hands || !out) return PHE_EINVAL;
    const size_t in_bytes = (size_t)n * (layout == PHE_LAYOUT_IDS ? 7 : 8);
    int rc;
    cudaStream_t hs;
    if ((rc = t_hstream.get(&hs))) return rc;
    if (align256(in_bytes) + (size_t)n * 2 <= kPinnedBytes) {
        unsigned char* pin;
        if ((rc = t_pin.get(&pin))) return rc;
        std::memcpy(pin, hands, in_bytes);
        uint16_t* res = reinterpret_cast<uint16_t*>(pin + align256(in_bytes));
        if ((rc = phe_eval7(pin, layout, n, res, hs))) return rc;
        PHE_CUDA(cudaStreamSynchronize(hs));
        std::memcpy(out, res, (size_t)n * 2);
        return PHE_OK;
    }
    if ((rc = t_in.ensure(in_bytes)) || (rc = t_out.ensure((size_t)n * 2))) return rc;
    PHE_CUDA(cudaMemcpyAsync(t_in.p, hands, in_bytes, cudaMemcpyHostToDevice, hs));
    if ((rc = phe_eval7(t_in.p, layout, n, static_cast<uint16_t*>(t_out.p), hs))) return rc;
    PHE_CUDA(cudaMemcpyAsync(out, t_out.p, (size_t)n * 2, cudaMemcpyDeviceToHost, hs));
    PHE_CUDA(cudaStreamSynchronize(hs));
    return PHE_OK;
}

// ---- enum7 -------------------------------------------------------------------------------------
static int launch_enum7(uint64_t* hist9, uint64_t* hist7463, uint64_t comb_begin, uint64_t comb_end, uint32_t part, uint32_t nparts,
                       const FusedArgs& fa, void* stream)
{
    if (comb_begin > comb_end || comb_end > PHE_NUM_HANDS_7 || nparts == 0 || part >= nparts) return PHE_EINVAL;
    if (fa.world == 0 && (comb_begin == comb_end || (!hist9 && !hist7463))) return PHE_OK;
    DevCtx* cx = nullptr;
    int rc = current_ctx(&cx);
    if (rc) return rc;
    const uint64_t total = comb_end - comb_begin;
    // work units [unit_first, unit_first + n_units) of the static unit table overlap the range; a tile is 1..4 of them: about
    // 8 tiles per resident warp of one part, but never more than 4 units (2048 hands: tail balance)
    const std::vector<uint32_t>& units = g_enum_units;
    uint64_t unit_first = 0, n_units = 0;
    if (total) {
        unit_first = (uint64_t)(std::upper_bound(units.begin(), units.end(), (uint32_t)comb_begin) - units.begin()) - 1;
        const uint64_t unit_last = (uint64_t)(std::lower_bound(units.begin(), units.end(), (uint32_t)comb_end) - units.begin()) - 1;
        n_units = unit_last - unit_first + 1;
    }
#ifndef PHE_ENUM_TILES_PER_WARP
#define PHE_ENUM_TILES_PER_WARP 8
#endif
    uint64_t per_tile = (total / ((uint64_t)nparts * cx->sm_count * (kCtaThreads / 32) * PHE_ENUM_TILES_PER_WARP) + kEnumUnitHands / 2) / kEnumUnitHands;   // rounded
    if (per_tile < 1) per_tile = 1;
#ifndef PHE_ENUM_TILE_UNITS
#define PHE_ENUM_TILE_UNITS 4
#endif
    if (per_tile > PHE_ENUM_TILE_UNITS) per_tile = PHE_ENUM_TILE_UNITS;
    const uint64_t ntiles = (n_units + per_tile - 1) / per_tile;
    const uint64_t mine = ntiles > part ? (ntiles - part + nparts - 1) / nparts : 0;
    if (mine == 0 && fa.world == 0) return PHE_OK;
    uint64_t ctas = (mine + kCtaThreads / 32 - 1) / (kCtaThreads / 32);
    if (ctas == 0) ctas = 1;   // fused mode: a rank without tiles still takes part in the all-reduce
    if (ctas > (uint64_t)cx->sm_count) ctas = cx->sm_count;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // {tile counter, completion ticket}: zero at rest, re-armed by the kernel's last CTA
    const uint32_t seq = cx->next_counter.fetch_add(1, std::memory_order_relaxed);
    unsigned int* counter = cx->d_counters + 2 * (seq % kCounterRing);
    uint32_t* h32 = cx->d_h32 + (size_t)(seq % kEnumRing) * kHistBinsPad;
    // static share of this part's tiles: whole rounds over the grid's warps, at most 7/8 of the tiles and at most
    // kStaticRoundsMax rounds (work pinned to a warp cannot be levelled afterwards, so long launches stay mostly dynamic)
    const uint64_t n_warps = ctas * (kCtaThreads / 32);
    uint64_t rounds = (mine - (mine >> 3)) / n_warps;
    constexpr uint64_t kStaticRoundsMax = 6;   // swept 0..8 on a B200: 6 is best or tied for the whole range, halves and eighths
    if (rounds > kStaticRoundsMax) rounds = kStaticRoundsMax;
    const uint64_t n_static = rounds * n_warps;
    k_enum7<<<(unsigned)ctas, kCtaThreads, kSmemEnum, st>>>(
        cx->d_enum, cx->d_trimask, reinterpret_cast<unsigned long long*>(hist9), reinterpret_cast<unsigned long long*>(hist7463), comb_begin, comb_end,
        cx->d_units, cx->d_unit_starts, reinterpret_cast<const uint4*>(cx->d_blocks), (uint32_t)unit_first, (uint32_t)n_units, (uint32_t)per_tile, (uint32_t)mine, (uint32_t)n_static, counter, h32, part, nparts, fa);
    return after_launch("enum7 launch");
}

int phe_enum7_part(uint64_t* hist9, uint64_t* hist7463, uint64_t comb_begin, uint64_t comb_end, uint32_t part, uint32_t nparts,
                   void* stream)
{
    FusedArgs fa{};
    return launch_enum7(hist9, hist7463, comb_begin, comb_end, part, nparts, fa, stream);
}

uint64_t phe_enum7_allreduce_workspace(void) { return kWsBytes; }

#ifdef PHE_ENUM_PROF
extern "C" int phe_debug_enum_prof(unsigned long long* host_out_1024)   // lab builds only (tools/lab/enum_stamps.py)
{
    PHE_CUDA(cudaDeviceSynchronize());
    PHE_CUDA(cudaMemcpyFromSymbol(host_out_1024, g_enum_prof, sizeof(unsigned long long) * 1024));
    return PHE_OK;
}
#endif

int phe_enum7_allreduce_nvls(uint64_t* out_hist7472, uint64_t comb_begin, uint64_t comb_end, int rank, int world, void* const* peer_ws,
                             void* multicast_ws, void* stream)
{
    if (!out_hist7472 || !peer_ws || world < 1 || world > kMaxWorld || rank < 0 || rank >= world) return PHE_EINVAL;
    FusedArgs fa{};
    fa.rank = rank;
    fa.world = world;
    for (int r = 0; r < world; r++) {
        if (!peer_ws[r]) return PHE_EINVAL;
        fa.ws[r] = static_cast<unsigned char*>(peer_ws[r]);
    }
    fa.out = reinterpret_cast<unsigned long long*>(out_hist7472);
    fa.mc = static_cast<unsigned char*>(multicast_ws);
    return launch_enum7(nullptr, nullptr, comb_begin, comb_end, (uint32_t)rank, (uint32_t)world, fa, stream);
}

int phe_enum7_allreduce(uint64_t* out_hist7472, uint64_t comb_begin, uint64_t comb_end, int rank, int world, void* const* peer_ws,
                        void* stream)
{
    return phe_enum7_allreduce_nvls(out_hist7472, comb_begin, comb_end, rank, world, peer_ws, nullptr, stream);
}

int phe_enum7_allreduce_status(const void* my_ws, void* stream, uint32_t* launches_out, uint32_t* failed_launch_out)
{
    if (!my_ws || !launches_out || !failed_launch_out) return PHE_EINVAL;
    uint32_t words[2] = {0, 0};    // {epoch = launches completed, launch number of the last exchange that timed out (0 = none)}
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    PHE_CUDA(cudaMemcpyAsync(words, static_cast<const unsigned char*>(my_ws) + kWsEpoch, sizeof words, cudaMemcpyDeviceToHost, st));
    PHE_CUDA(cudaStreamSynchronize(st));
    *launches_out = words[0];
    *failed_launch_out = words[1];
    return PHE_OK;
}

int phe_enum7(uint64_t* hist9, uint64_t* hist7463, uint64_t comb_begin, uint64_t comb_end, void* stream)
{
    return phe_enum7_part(hist9, hist7463, comb_begin, comb_end, 0, 1, stream);
}

int phe_enum7_host(uint64_t* hist9, uint64_t* hist7463, uint64_t comb_begin, uint64_t comb_end)
{
    if (comb_begin > comb_end || comb_end > PHE_NUM_HANDS_7) return PHE_EINVAL;
    int rc;
    cudaStream_t hs;
    if ((rc = t_hstream.get(&hs))) return rc;
    const size_t bytes = (size_t)(9 + kHistBins) * 8;
    if ((rc = t_out.ensure(bytes))) return rc;
    uint64_t* d9 = static_cast<uint64_t*>(t_out.p);
    uint64_t* d7 = d9 + 9;
    PHE_CUDA(cudaMemsetAsync(t_out.p, 0, bytes, hs));
    if ((rc = phe_enum7(d9, d7, comb_begin, comb_end, hs))) return rc;
    // the 59,776-byte result comes back through the per-thread pinned buffer: a pageable destination would add a
    // staging copy inside the driver to a call that is only ~0.2 ms long
    static_assert((9 + kHistBins) * 8 <= kPinnedBytes, "histograms fit the pinned buffer");
    unsigned char* pin;
    if ((rc = t_pin.get(&pin))) return rc;
    PHE_CUDA(cudaMemcpyAsync(pin, t_out.p, bytes, cudaMemcpyDeviceToHost, hs));
    PHE_CUDA(cudaStreamSynchronize(hs));
    const uint64_t* tmp = reinterpret_cast<const uint64_t*>(pin);
    if (hist9) for (int i = 0; i < 9; i++) hist9[i] += tmp[i];
    if (hist7463) for (int i = 0; i < kHistBins; i++) hist7463[i] += tmp[9 + i];
    return PHE_OK;
}

// ---- equity_mc ---------------------------------------------------------------------------------
int phe_equity_mc(const uint64_t* states, int64_t n_states, int n_opp, uint64_t rollouts, uint64_t sample_offset, uint64_t seed,
                  uint32_t state_index_base, uint64_t* wtl, void* stream)
{
    if (n_states < 0 || n_opp < 1 || n_opp > PHE_MAX_PLAYERS - 1) return PHE_EINVAL;
    if (n_states == 0 || rollouts == 0) return PHE_OK;
    if (!states || !wtl) return PHE_EINVAL;
    DevCtx* cx = nullptr;
    int rc = current_ctx(&cx);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // samples per warp item: aim at >= 4 items per resident warp, 32..8192, multiple of 32
    const double total = (double)n_states * (double)rollouts;
    const uint64_t warps = (uint64_t)cx->sm_count * (kMcThreads / 32);
    uint64_t chunk = (uint64_t)(total / (double)(warps * 4));
    chunk = (chunk + 31) & ~(uint64_t)31;
    if (chunk < 32) chunk = 32;
    if (chunk > 8192) chunk = 8192;
    if (chunk > ((rollouts + 31) & ~(uint64_t)31)) chunk = (rollouts + 31) & ~(uint64_t)31;
    while ((uint64_t)n_states * ((rollouts + chunk - 1) / chunk) > (1ull << 31) && chunk < (1ull << 31)) chunk *= 2;   // 32-bit item counter
    const uint64_t items = (uint64_t)n_states * ((rollouts + chunk - 1) / chunk);
    const auto* sp = reinterpret_cast<const ulonglong2*>(states);
    auto* wp = reinterpret_cast<unsigned long long*>(wtl);
    if (total < 1.0e6) {
        uint64_t blocks = (items + 7) / 8;   // 8 warps per 256-thread block
        if (blocks > (uint64_t)cx->sm_count * 8) blocks = (uint64_t)cx->sm_count * 8;
        k_equity_mc_gmem<<<(unsigned)blocks, 256, 0, st>>>(cx->d_image, cx->d_pairmask, sp, n_states, n_opp, rollouts, sample_offset, seed, state_index_base, wp, (uint32_t)chunk);
    } else {
        uint64_t ctas = (items + kMcThreads / 32 - 1) / (kMcThreads / 32);
        if (ctas > (uint64_t)cx->sm_count) ctas = cx->sm_count;
        unsigned int* counters = cx->d_counters + 2 * (cx->next_counter.fetch_add(1, std::memory_order_relaxed) % kCounterRing);
#ifdef PHE_MC_POOL
        k_equity_mc_pool<<<(unsigned)ctas, kCtaThreads, kSmemMcPool, st>>>(cx->d_image, cx->d_pairmask, sp, n_states, n_opp, rollouts, sample_offset, seed, state_index_base, wp, (uint32_t)chunk, counters);
#else
        k_equity_mc_smem<<<(unsigned)ctas, kMcThreads, kSmemMc, st>>>(cx->d_image, cx->d_pairmask, sp, n_states, n_opp, rollouts, sample_offset, round_keys(seed), state_index_base, wp, (uint32_t)chunk, counters);
#endif
    }
    return after_launch("equity_mc launch");
}

int phe_equity_mc_host(const uint64_t* states, int64_t n_states, int n_opp, uint64_t rollouts, uint64_t sample_offset, uint64_t seed,
                       uint32_t state_index_base, uint64_t* wtl)
{
    if (n_states < 0) return PHE_EINVAL;
    if (n_states == 0) return PHE_OK;
    if (!states || !wtl) return PHE_EINVAL;
    int rc;
    cudaStream_t hs;
    if ((rc = t_hstream.get(&hs))) return rc;
    const size_t in_bytes = (size_t)n_states * 16, out_bytes = (size_t)n_states * 24;
    if ((rc = t_in.ensure(in_bytes)) || (rc = t_out.ensure(out_bytes))) return rc;
    PHE_CUDA(cudaMemcpyAsync(t_in.p, states, in_bytes, cudaMemcpyHostToDevice, hs));
    PHE_CUDA(cudaMemsetAsync(t_out.p, 0, out_bytes, hs));
    if ((rc = phe_equity_mc(static_cast<const uint64_t*>(t_in.p), n_states, n_opp, rollouts, sample_offset, seed, state_index_base,
                            static_cast<uint64_t*>(t_out.p), hs)))
        return rc;
    std::vector<uint64_t> tmp((size_t)n_states * 3);
    PHE_CUDA(cudaMemcpyAsync(tmp.data(), t_out.p, out_bytes, cudaMemcpyDeviceToHost, hs));
    PHE_CUDA(cudaStreamSynchronize(hs));
    for (size_t i = 0; i < tmp.size(); i++) wtl[i] += tmp[i];
    return PHE_OK;
}

// ---- equity_plan -------------------------------------------------------------------------------
uint64_t phe_plan_num_deals(int n_exist, const int32_t* plan, int n_steps)
{
    PlanDesc pd;
    return parse_plan(n_exist, plan, n_steps, &pd) ? plan_num_deals(pd) : 0;
}

int phe_equity_plan(const uint8_t* exist, int n_exist, int64_t n_tasks, const int32_t* plan, int n_steps, uint64_t seed,
                    uint32_t stream_base, uint64_t* wtl, void* stream)
{
    if (n_tasks < 0) return PHE_EINVAL;
    PlanDesc pd;
    if (!parse_plan(n_exist, plan, n_steps, &pd)) return PHE_EPLAN;
    if (n_tasks == 0) return PHE_OK;
    if (!exist || !wtl) return PHE_EINVAL;
    DevCtx* cx = nullptr;
    int rc = current_ctx(&cx);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const double deals = (double)n_tasks * (double)plan_num_deals(pd);
    const uint64_t threads = (uint64_t)n_tasks * pd.n_leaves;
    auto* wp = reinterpret_cast<unsigned long long*>(wtl);
    if (pd.pair_tail) {
        const uint64_t items = (uint64_t)n_tasks * pd.n_prefix;   // one warp each
        if (deals < 2.0e6) {
            uint64_t blocks = (items + 7) / 8;
            if (blocks > (uint64_t)cx->sm_count * 8) blocks = (uint64_t)cx->sm_count * 8;
            k_equity_plan_pairs_gmem<<<(unsigned)blocks, 256, 0, st>>>(cx->d_image, cx->d_pairmask, exist, n_tasks, pd, wp);
        } else {
            uint64_t ctas = (items + 31) / 32;
            if (ctas > (uint64_t)cx->sm_count) ctas = cx->sm_count;
            k_equity_plan_pairs_smem<<<(unsigned)ctas, kCtaThreads, kSmemPairs, st>>>(cx->d_image, cx->d_pairmask, exist, n_tasks, pd, wp);
        }
        return after_launch("equity_plan (pair tail) launch");
    }
    if (deals < 2.0e6) {
        uint64_t blocks = (threads + 255) / 256;
        if (blocks > (uint64_t)cx->sm_count * 8) blocks = (uint64_t)cx->sm_count * 8;
        k_equity_plan_gmem<<<(unsigned)blocks, 256, 0, st>>>(cx->d_image, cx->d_pairmask, exist, n_tasks, pd, seed, stream_base, wp);
    } else {
        uint64_t ctas = (threads + kCtaThreads - 1) / kCtaThreads;
        if (ctas > (uint64_t)cx->sm_count) ctas = cx->sm_count;
        k_equity_plan_smem<<<(unsigned)ctas, kCtaThreads, kSmemTabOnly, st>>>(cx->d_image, cx->d_pairmask, exist, n_tasks, pd, seed, stream_base, wp);
    }
    return after_launch("equity_plan launch");
}

int phe_equity_plan_host(const uint8_t* exist, int n_exist, int64_t n_tasks, const int32_t* plan, int n_steps, uint64_t seed,
                         uint32_t stream_base, uint64_t* wtl)
{
    if (n_tasks < 0) return PHE_EINVAL;
    PlanDesc pd;
    if (!parse_plan(n_exist, plan, n_steps, &pd)) return PHE_EPLAN;
    if (n_tasks == 0) return PHE_OK;
    if (!exist || !wtl) return PHE_EINVAL;
    int rc;
    cudaStream_t hs;
    if ((rc = t_hstream.get(&hs))) return rc;
    const size_t in_bytes = (size_t)n_tasks * 8, out_bytes = (size_t)n_tasks * 24;
    if ((rc = t_in.ensure(in_bytes)) || (rc = t_out.ensure(out_bytes))) return rc;
    PHE_CUDA(cudaMemcpyAsync(t_in.p, exist, in_bytes, cudaMemcpyHostToDevice, hs));
    PHE_CUDA(cudaMemsetAsync(t_out.p, 0, out_bytes, hs));
    if ((rc = phe_equity_plan(static_cast<const uint8_t*>(t_in.p), n_exist, n_tasks, plan, n_steps, seed, stream_base,
                              static_cast<uint64_t*>(t_out.p), hs)))
        return rc;
    std::vector<uint64_t> tmp((size_t)n_tasks * 3);
    PHE_CUDA(cudaMemcpyAsync(tmp.data(), t_out.p, out_bytes, cudaMemcpyDeviceToHost, hs));
    PHE_CUDA(cudaStreamSynchronize(hs));
    for (size_t i = 0; i < tmp.size(); i++) wtl[i] += tmp[i];
    return PHE_OK;
}

// ---- showdown ----------------------------------------------------------------------------------
int phe_showdown(const uint64_t* boards, const uint64_t* holes, const uint32_t* live, int64_t n_games, int n_players,
                 uint32_t* winners_out, uint16_t* ranks_out, void* stream)
{
    if (n_games < 0 || n_players < 1 || n_players > PHE_MAX_PLAYERS) return PHE_EINVAL;
    if (n_games == 0) return PHE_OK;
    if (!boards || !holes || !winners_out) return PHE_EINVAL;
    DevCtx* cx = nullptr;
    int rc = current_ctx(&cx);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n_games * n_players < kSmallBatch) {
        const unsigned blocks = (unsigned)((n_games + 255) / 256);
        k_showdown_gmem<<<blocks, 256, 0, st>>>(cx->d_image, boards, holes, live, n_games, n_players, winners_out, ranks_out);
    } else {
        uint64_t ctas = ((uint64_t)n_games + kCtaThreads - 1) / kCtaThreads;
        if (ctas > (uint64_t)cx->sm_count) ctas = cx->sm_count;
        k_showdown_smem<<<(unsigned)ctas, kCtaThreads, kSmemTabOnly, st>>>(cx->d_image, boards, holes, live, n_games, n_players, winners_out, ranks_out);
    }
    return after_launch("showdown launch");
}

int phe_showdown_host(const uint64_t* boards, const uint64_t* holes, const uint32_t* live, int64_t n_games, int n_players,
                      uint32_t* winners_out, uint16_t* ranks_out)
{
    if (n_games < 0 || n_players < 1 || n_players > PHE_MAX_PLAYERS) return PHE_EINVAL;
    if (n_games == 0) return PHE_OK;
    if (!boards || !holes || !winners_out) return PHE_EINVAL;
    const size_t G = (size_t)n_games, P = (size_t)n_players;
    const size_t o_boards = 0, o_holes = align256(G * 8), o_live = o_holes + align256(G * P * 8);
    const size_t in_bytes = o_live + align256(G * 4);
    const size_t o_win = 0, o_ranks = align256(G * 4), out_bytes = o_ranks + align256(G * P * 2);
    int rc;
    cudaStream_t hs;
    if ((rc = t_hstream.get(&hs))) return rc;
    if (in_bytes + out_bytes <= kPinnedBytes) {
        unsigned char* pin;
        if ((rc = t_pin.get(&pin))) return rc;
        std::memcpy(pin + o_boards, boards, G * 8);
        std::memcpy(pin + o_holes, holes, G * P * 8);
        if (live) std::memcpy(pin + o_live, live, G * 4);
        unsigned char* po = pin + in_bytes;
        if ((rc = phe_showdown(reinterpret_cast<uint64_t*>(pin + o_boards), reinterpret_cast<uint64_t*>(pin + o_holes),
                               live ? reinterpret_cast<uint32_t*>(pin + o_live) : nullptr, n_games, n_players,
                               reinterpret_cast<uint32_t*>(po + o_win), ranks_out ? reinterpret_cast<uint16_t*>(po + o_ranks) : nullptr, hs)))
            return rc;
        PHE_CUDA(cudaStreamSynchronize(hs));
        std::memcpy(winners_out, po + o_win, G * 4);
        if (ranks_out) std::memcpy(ranks_out, po + o_ranks, G * P * 2);
        return PHE_OK;
    }
    if ((rc = t_in.ensure(in_bytes)) || (rc = t_out.ensure(out_bytes))) return rc;
    char* di = static_cast<char*>(t_in.p);
    char* dout = static_cast<char*>(t_out.p);
    PHE_CUDA(cudaMemcpyAsync(di + o_boards, boards, G * 8, cudaMemcpyHostToDevice, hs));
    PHE_CUDA(cudaMemcpyAsync(di + o_holes, holes, G * P * 8, cudaMemcpyHostToDevice, hs));
    if (live) PHE_CUDA(cudaMemcpyAsync(di + o_live, live, G * 4, cudaMemcpyHostToDevice, hs));
    if ((rc = phe_showdown(reinterpret_cast<uint64_t*>(di + o_boards), reinterpret_cast<uint64_t*>(di + o_holes),
                           live ? reinterpret_cast<uint32_t*>(di + o_live) : nullptr, n_games, n_players,
                           reinterpret_cast<uint32_t*>(dout + o_win), ranks_out ? reinterpret_cast<uint16_t*>(dout + o_ranks) : nullptr, hs)))
        return rc;
    PHE_CUDA(cudaMemcpyAsync(winners_out, dout + o_win, G * 4, cudaMemcpyDeviceToHost, hs));
    if (ranks_out) PHE_CUDA(cudaMemcpyAsync(ranks_out, dout + o_ranks, G * P * 2, cudaMemcpyDeviceToHost, hs));
    PHE_CUDA(cudaStreamSynchronize(hs));
    return PHE_OK;
}

// ---- settle ------------------------------------------------------------------------------------
int phe_settle(const uint64_t* boards, const uint64_t* holes, const int8_t* actions, const int32_t* values, const int32_t* n_rounds,
               const int32_t* button, const uint32_t* onboard, int64_t n_games, int n_players, int max_rounds, int small_blind,
               int32_t* won_out, int8_t* result_out, int8_t* card_result_out, uint32_t* paid_out, void* stream)
{
    if (n_games < 0 || n_players < 1 || n_players > SETTLE_MAX_PLAYERS || max_rounds < 1 || max_rounds > SETTLE_MAX_ROUNDS || small_blind < 1)
        return PHE_EINVAL;
    if (n_games == 0) return PHE_OK;
    if (!boards || !holes || !actions || !values || !n_rounds || !button || !onboard || !won_out || !result_out || !card_result_out)
        return PHE_EINVAL;
    DevCtx* cx = nullptr;
    int rc = current_ctx(&cx);
    if (rc) return rc;
    const unsigned blocks = (unsigned)((n_games + 127) / 128);
    k_settle<<<blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(cx->d_image, boards, holes, actions, values, n_rounds, button, onboard, n_games,
                                                                    n_players, max_rounds, small_blind, won_out, result_out, card_result_out, paid_out);
    return after_launch("settle launch");
}

int phe_settle_host(const uint64_t* boards, const uint64_t* holes, const int8_t* actions, const int32_t* values, const int32_t* n_rounds,
                    const int32_t* button, const uint32_t* onboard, int64_t n_games, int n_players, int max_rounds, int small_blind,
                    int32_t* won_out, int8_t* result_out, int8_t* card_result_out, uint32_t* paid_out)
{
    if (n_games < 0 || n_players < 1 || n_players > SETTLE_MAX_PLAYERS || max_rounds < 1 || max_rounds > SETTLE_MAX_ROUNDS) return PHE_EINVAL;
    if (n_games == 0) return PHE_OK;
    if (!boards || !holes || !actions || !values || !n_rounds || !button || !onboard || !won_out || !result_out || !card_result_out)
        return PHE_EINVAL;
    const size_t G = (size_t)n_games, P = (size_t)n_players, R = (size_t)max_rounds;
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += align256(bytes); return o; };
    const size_t o_b = take(G * 8), o_h = take(G * P * 8), o_a = take(G * R * P), o_v = take(G * R * P * 4), o_n = take(G * 4), o_bt = take(G * 4),
                 o_ob = take(G * 4);
    const size_t in_bytes = off;
    off = 0;
    const size_t o_w = take(G * P * 4), o_r = take(G * P), o_c = take(G * P), o_p = take(G * 4);
    const size_t out_bytes = off;
    int rc;
    cudaStream_t hs;
    if ((rc = t_hstream.get(&hs))) return rc;
    if ((rc = t_in.ensure(in_bytes)) || (rc = t_out.ensure(out_bytes))) return rc;
    char* di = static_cast<char*>(t_in.p);
    char* dout = static_cast<char*>(t_out.p);
    PHE_CUDA(cudaMemcpyAsync(di + o_b, boards, G * 8, cudaMemcpyHostToDevice, hs));
    PHE_CUDA(cudaMemcpyAsync(di + o_h, holes, G * P * 8, cudaMemcpyHostToDevice, hs));
    PHE_CUDA(cudaMemcpyAsync(di + o_a, actions, G * R * P, cudaMemcpyHostToDevice, hs));
    PHE_CUDA(cudaMemcpyAsync(di + o_v, values, G * R * P * 4, cudaMemcpyHostToDevice, hs));
    PHE_CUDA(cudaMemcpyAsync(di + o_n, n_rounds, G * 4, cudaMemcpyHostToDevice, hs));
    PHE_CUDA(cudaMemcpyAsync(di + o_bt, button, G * 4, cudaMemcpyHostToDevice, hs));
    PHE_CUDA(cudaMemcpyAsync(di + o_ob, onboard, G * 4, cudaMemcpyHostToDevice, hs));
    if ((rc = phe_settle(reinterpret_cast<uint64_t*>(di + o_b), reinterpret_cast<uint64_t*>(di + o_h), reinterpret_cast<int8_t*>(di + o_a),
                         reinterpret_cast<int32_t*>(di + o_v), reinterpret_cast<int32_t*>(di + o_n), reinterpret_cast<int32_t*>(di + o_bt),
                         reinterpret_cast<uint32_t*>(di + o_ob), n_games, n_players, max_rounds, small_blind, reinterpret_cast<int32_t*>(dout + o_w),
                         reinterpret_cast<int8_t*>(dout + o_r), reinterpret_cast<int8_t*>(dout + o_c), reinterpret_cast<uint32_t*>(dout + o_p), hs)))
        return rc;
    PHE_CUDA(cudaMemcpyAsync(won_out, dout + o_w, G * P * 4, cudaMemcpyDeviceToHost, hs));
    PHE_CUDA(cudaMemcpyAsync(result_out, dout + o_r, G * P, cudaMemcpyDeviceToHost, hs));
    PHE_CUDA(cudaMemcpyAsync(card_result_out, dout + o_c, G * P, cudaMemcpyDeviceToHost, hs));
    if (paid_out) PHE_CUDA(cudaMemcpyAsync(paid_out, dout + o_p, G * 4, cudaMemcpyDeviceToHost, hs));
    PHE_CUDA(cudaStreamSynchronize(hs));
    return PHE_OK;
}

// ---- deal --------------------------------------------------------------------------------------
static int launch_deal(const uint64_t* known, int64_t n, int need, uint64_t seed, uint64_t sample_offset, uint32_t stream_base, int by_sample,
                       uint8_t* cards_out, void* stream)
{
    if (n < 0 || need < 1 || need > MAX_DEAL) return PHE_EINVAL;
    if (n == 0) return PHE_OK;
    if (!known || !cards_out) return PHE_EINVAL;
    DevCtx* cx = nullptr;
    int rc = current_ctx(&cx);
    if (rc) return rc;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    k_deal<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(cx->d_pairmask, known, n, need, seed, sample_offset, stream_base, by_sample, cards_out);
    return after_launch("deal launch");
}

static int deal_host(const uint64_t* known, int64_t n, int need, uint64_t seed, uint64_t sample_offset, uint32_t stream_base, int by_sample,
                     uint8_t* cards_out)
{
    if (n < 0 || need < 1 || need > MAX_DEAL) return PHE_EINVAL;
    if (n == 0) return PHE_OK;
    if (!known || !cards_out) return PHE_EINVAL;
    int rc;
    cudaStream_t hs;
    if ((rc = t_hstream.get(&hs))) return rc;
    if ((rc = t_in.ensure((size_t)n * 8)) || (rc = t_out.ensure((size_t)n * need))) return rc;
    PHE_CUDA(cudaMemcpyAsync(t_in.p, known, (size_t)n * 8, cudaMemcpyHostToDevice, hs));
    if ((rc = launch_deal(static_cast<const uint64_t*>(t_in.p), n, need, seed, sample_offset, stream_base, by_sample, static_cast<uint8_t*>(t_out.p), hs)))
        return rc;
    PHE_CUDA(cudaMemcpyAsync(cards_out, t_out.p, (size_t)n * need, cudaMemcpyDeviceToHost, hs));
    PHE_CUDA(cudaStreamSynchronize(hs));
    return PHE_OK;
}

int phe_deal(const uint64_t* known, int64_t n, int need, uint64_t seed, uint64_t sample_offset, uint32_t stream_base, uint8_t* cards_out, void* stream)
{
    return launch_deal(known, n, need, seed, sample_offset, stream_base, 0, cards_out, stream);
}

int phe_deal_host(const uint64_t* known, int64_t n, int need, uint64_t seed, uint64_t sample_offset, uint32_t stream_base, uint8_t* cards_out)
{
    return deal_host(known, n, need, seed, sample_offset, stream_base, 0, cards_out);
}

int phe_deal_samples(const uint64_t* known, int64_t n, int need, uint64_t seed, uint64_t sample_offset, uint32_t stream_id, uint8_t* cards_out, void* stream)
{
    return launch_deal(known, n, need, seed, sample_offset, stream_id, 1, cards_out, stream);
}

int phe_deal_samples_host(const uint64_t* known, int64_t n, int need, uint64_t seed, uint64_t sample_offset, uint32_t stream_id, uint8_t* cards_out)
{
    return deal_host(known, n, need, seed, sample_offset, stream_id, 1, cards_out);
}

}  // extern "C"
