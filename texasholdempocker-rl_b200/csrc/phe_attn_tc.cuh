// phe_attn_tc.cuh -- self-attention of the leaf network on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// softmax(Q K^T / sqrt(D)) V for one (batch, head) pair at a time: <= 112 keys, <= 128 query rows, head dim 64, bf16 --
// the shape of rl/models.py:446-447 with config/default.yml (98 tokens, 10 heads of 64).  Replaces the mma.sync kernel
// (k_attention<64> in phe_nn.cu) for that shape.
//
//   S = Q K^T   one UMMA M=128 x N=112 x K=64 (four tcgen05.mma, K = 16 each): Q and K tiles in shared memory in the
//               canonical K-major 128-byte-swizzle layout, the fp32 score tile in TENSOR MEMORY (lane = query row)
//   softmax     tcgen05.ld: thread r of a 128-thread warpgroup receives the whole row r (112 scores) -- row maximum and
//               row sum need no shuffles; the probabilities are packed to bf16 and written back over the score tile
//               (tcgen05.st, 56 columns)
//   O = P V     seven tcgen05.mma with the A operand (P) read from tensor memory and V from shared memory as an
//               MN-major operand (keys x 64 is already "K x N with N contiguous": no transpose anywhere); O (64 fp32
//               columns) lands in the upper half of the dead score tile
//   epilogue    tcgen05.ld O, scale by 1 / row sum, 128 bytes of bf16 per query row straight to global memory
//
// A CTA holds four independent warpgroups (512 threads; one (batch, head) pair each in flight, 128 TMEM columns each = the
// whole tensor memory of the SM, one 44 KB shared-memory stage each).  The single thread that issues a warpgroup's MMAs is
// its thread 0; completion comes back through an mbarrier (tcgen05.commit).  A stage is refilled as early as its operands
// are dead: the Q and K tiles of the warpgroup's NEXT pair are requested as soon as S = Q K^T has completed (they land
// under the softmax, P V and the epilogue), the V tile as soon as P V has completed.  Tiles are loaded by TMA
// (cp.async.bulk.tensor.4d, 128-byte swizzle, one instruction per tile issued by the warpgroup's thread 0, completion on
// an mbarrier): the three 4-D tensor maps (d, token, head, batch) describe the strided views of the packed QKV projection
// as they are, and rows beyond the sequence are zero-filled by the out-of-bounds rule -- nothing is copied, padded or
// zeroed by threads.  (A first version issued 16-byte cp.async copies from the compute threads: they stalled 4,300 cycles
// per pair behind the LSU queue, the kernel stayed at the mma.sync kernel's 167 us.)  The softmax walks the score row
// twice in 16-column pieces (maximum, then exponentials) instead of holding 112 scores: 128 registers per thread.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace phe_attn_tc {

#ifdef PHE_ATTN_TC_PROF
#define TCP_T(i) do { if (blockIdx.x == 0 && tid == 0) { const long long n_ = clock64(); prof[i] += n_ - last; last = n_; } } while (0)
#else
#define TCP_T(i) do {} while (0)
#endif

constexpr int kD = 64;
constexpr int kM = 128;                 // query rows of the MMA (TMEM lanes); rows >= s_q are zero
constexpr int kN = 112;                 // key columns; rows >= s_kv of the K / V tiles are zero
constexpr int kWG = 4;                  // warpgroups per CTA
constexpr int kStages = 1;              // shared-memory stages per warpgroup
constexpr int kThreads = kWG * 128;
constexpr uint32_t kQBytes = kM * 128, kKBytes = kN * 128, kVBytes = kN * 128;
constexpr uint32_t kStageBytes = kQBytes + kKBytes + kVBytes;              // 45,056 = 44 x 1024
constexpr uint32_t kTilesBytes = kWG * kStages * kStageBytes;              // 180,224
constexpr uint32_t kSmemBytes = kTilesBytes + 1024 + 128;                  // + alignment slack + barriers / TMEM address
constexpr uint32_t kTmemCols = 128 * kWG;                                  // 128 per warpgroup: S / P in [0, 112), O in [64, 128)
constexpr uint32_t kOCol = 64;

// instruction descriptors (cute::UMMA::InstrDescriptor): c = f32, a = b = bf16, M = 128
constexpr uint32_t kIdescS = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);                 // N = 112, A and B K-major
constexpr uint32_t kIdescO = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(kD >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);   // N = 64, B MN-major

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): 128-byte swizzle, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; spins++) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (spins > (1u << 26)) __trap();      // an MMA that never completes must not hang the device
    }
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// one (rows x 64) bf16 tile of (batch b, head h): coordinates (d, token, head, batch)
__device__ __forceinline__ void tma_tile(uint32_t dst, const CUtensorMap* map, int h, int b, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(0), "r"(0), "r"(h), "r"(b), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void wg_sync(int wg) { asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 16 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
        "%23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2f(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi)
{
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}

__global__ void __launch_bounds__(kThreads, 1)
k_attention_tc(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
               __nv_bfloat16* __restrict__ out, int total_bh, int heads, int s_q, int s_kv, float scale_log2e)
{
    extern __shared__ unsigned char attn_tc_smem_raw[];
    const uint32_t raw = (uint32_t)__cvta_generic_to_shared(attn_tc_smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;                  // the swizzle pattern is relative to 1024-byte boundaries
    const uint32_t ctl = tiles + kTilesBytes;                        // per warpgroup three mbarriers (MMA, Q+K landed, V landed) at 32 g; [96, 100): TMEM base address
    const int tid = threadIdx.x, wg = tid >> 7, t = tid & 127, warp = tid >> 5;

    if (tid == 0) {
        for (int g = 0; g < 3 * kWG; g++) mbar_init(ctl + 8u * g, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ctl + 96), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(ctl + 96) : "memory");
    const uint32_t tm_wg = tmem_base + (uint32_t)wg * 128u;                              // this warpgroup's 128 columns
    const uint32_t tm_row = tm_wg + ((uint32_t)((warp & 3) * 32) << 16);                 // + this warp's lane quarter
    const uint32_t bar = ctl + 24u * (uint32_t)wg, bar_qk = bar + 8u, bar_v = bar + 16u;
    const uint32_t wg_tiles = tiles + (uint32_t)wg * kStages * kStageBytes;
#ifdef PHE_ATTN_TC_PROF
    long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0}, last = clock64();
#endif
    uint32_t phase = 0;
    const uint32_t sq = wg_tiles, sk = sq + kQBytes, sv = sk + kKBytes;

    // issued by the warpgroup's thread 0 only
    auto load_qk = [&](int bh) {
        const int b = bh / heads, h = bh - b * heads;
        mbar_expect_tx(bar_qk, kQBytes + kKBytes);
        tma_tile(sq, &map_q, h, b, bar_qk);
        tma_tile(sk, &map_k, h, b, bar_qk);
    };
    auto load_v = [&](int bh) {
        const int b = bh / heads, h = bh - b * heads;
        mbar_expect_tx(bar_v, kVBytes);
        tma_tile(sv, &map_v, h, b, bar_v);
    };
    uint32_t ph_qk = 0, ph_v = 0;

    const int step = (int)gridDim.x * kWG;
    int bh = (int)blockIdx.x * kWG + wg;
    if (bh < total_bh && t == 0) { load_qk(bh); load_v(bh); }
    for (; bh < total_bh; bh += step) {
        const bool more = bh + step < total_bh;
        TCP_T(0);
        __syncwarp();
        tc_fence_before();
        wg_sync(wg);                             // every thread has read the previous pair's O: the score tile may be overwritten
        TCP_T(1);
        if (t == 0) {
            mbar_wait(bar_qk, ph_qk);
            ph_qk ^= 1;
            tc_fence_after();
            const uint64_t da = smem_desc(sq), db = smem_desc(sk);
#pragma unroll
            for (int ks = 0; ks < kD / 16; ks++) mma_ss(tm_wg, da + (uint64_t)(ks * 2), db + (uint64_t)(ks * 2), kIdescS, ks > 0);     // + 32 bytes per K step
            tc_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        TCP_T(2);
        if (more && t == 0) load_qk(bh + step);  // Q and K are dead: the next pair's tiles land under the softmax, P V and the epilogue
        __syncwarp();
        tc_fence_after();
        TCP_T(3);
        // ---- softmax of row t: the row sits in this thread's TMEM lane.  Two walks (maximum, then exponentials) in 16-column
        // pieces; a piece is either all real keys (no per-element test), the boundary piece, or all padding (skipped: its
        // probabilities are zeros).  Pieces of 32 columns with the next load in flight were slower (register pressure). ----
        const int full = s_kv >> 4;               // pieces [0, full) hold only real keys
        float m = -INFINITY;
#pragma unroll
        for (int j = 0; j < kN / 16; j++) {
            if (16 * j < s_kv) {                  // uniform
                uint32_t sr[16];
                tmem_ld16(tm_row + 16u * j, sr);
                tmem_wait_ld();
                if (j < full) {
#pragma unroll
                    for (int i = 0; i < 16; i++) m = fmaxf(m, __uint_as_float(sr[i]));
                } else {
#pragma unroll
                    for (int i = 0; i < 16; i++)
                        if (16 * j + i < s_kv) m = fmaxf(m, __uint_as_float(sr[i]));
                }
            }
        }
        const float c0 = m * scale_log2e;
        float l = 0.f;
#pragma unroll
        for (int j = 0; j < kN / 16; j++) {
            uint32_t p[8];
            if (16 * j < s_kv) {
                uint32_t sr[16];
                tmem_ld16(tm_row + 16u * j, sr);
                tmem_wait_ld();
                if (j < full) {
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        const float e0 = ex2f(fmaf(__uint_as_float(sr[i]), scale_log2e, -c0)), e1 = ex2f(fmaf(__uint_as_float(sr[i + 1]), scale_log2e, -c0));
                        l += e0 + e1;
                        p[i >> 1] = pack2(e0, e1);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        const float e0 = 16 * j + i < s_kv ? ex2f(fmaf(__uint_as_float(sr[i]), scale_log2e, -c0)) : 0.f;
                        const float e1 = 16 * j + i + 1 < s_kv ? ex2f(fmaf(__uint_as_float(sr[i + 1]), scale_log2e, -c0)) : 0.f;
                        l += e0 + e1;
                        p[i >> 1] = pack2(e0, e1);
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; i++) p[i] = 0u;
            }
            tmem_st8(tm_row + 8u * j, p);        // P columns [8j, 8j + 8) only cover score columns that were already read
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        TCP_T(4);
        tc_fence_before();
        wg_sync(wg);
        if (t == 0) {
            mbar_wait(bar_v, ph_v);
            ph_v ^= 1;
            tc_fence_after();
            const uint64_t dv = smem_desc(sv);
#pragma unroll
            for (int ks = 0; ks < kN / 16; ks++)
                mma_ts(tm_wg + kOCol, tm_wg + 8u * ks, dv + (uint64_t)(ks * (2048 >> 4)), kIdescO, ks > 0);      // 16 keys = 2048 bytes of V per K step
            tc_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        TCP_T(5);
        if (more && t == 0) load_v(bh + step);   // V is dead
        __syncwarp();
        tc_fence_after();
        const float inv = 1.f / l;
        const int b = bh / heads, h = bh - b * heads;
        uint4* dst = reinterpret_cast<uint4*>(out + (((int64_t)b * s_q + t) * heads + h) * kD);
#pragma unroll
        for (int j = 0; j < kD / 16; j++) {
            uint32_t o[16];
            tmem_ld16(tm_row + kOCol + 16u * j, o);
            tmem_wait_ld();
            if (t < s_q) {
                uint4 o0, o1;
                o0.x = pack2(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
                o0.y = pack2(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
                o0.z = pack2(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
                o0.w = pack2(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
                o1.x = pack2(__uint_as_float(o[8]) * inv, __uint_as_float(o[9]) * inv);
                o1.y = pack2(__uint_as_float(o[10]) * inv, __uint_as_float(o[11]) * inv);
                o1.z = pack2(__uint_as_float(o[12]) * inv, __uint_as_float(o[13]) * inv);
                o1.w = pack2(__uint_as_float(o[14]) * inv, __uint_as_float(o[15]) * inv);
                dst[2 * j] = o0;
                dst[2 * j + 1] = o1;
            }
        }
        TCP_T(6);
        __syncwarp();
        tc_fence_before();      // the next pair's first MMA overwrites the score tile: ordered by the barrier at the top of the loop
    }
#ifdef PHE_ATTN_TC_PROF
    if (blockIdx.x == 0 && tid == 0)
        printf("cycles: other %lld | load wait + barrier %lld | MMA1 %lld | issue QK loads %lld | softmax %lld | barrier + MMA2 %lld | V loads + epilogue %lld\n", prof[0], prof[1],
               prof[2], prof[3], prof[4], prof[5], prof[6]);
#endif
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
}

// ---- host side: tensor maps of the strided Q / K / V views and the launch ------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult res;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &res) != cudaSuccess || res != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// view[b][token][h][d]: element strides batch_stride / row_stride / 64 / 1; box = rows x 64, rows >= tokens come back as zeros
inline bool make_map(CUtensorMap* map, const __nv_bfloat16* base, int64_t batch, int heads, int tokens, int64_t batch_stride, int64_t row_stride, int rows)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)kD, (cuuint64_t)tokens, (cuuint64_t)heads, (cuuint64_t)batch};
    const cuuint64_t strides[3] = {(cuuint64_t)row_stride * 2, (cuuint64_t)kD * 2, (cuuint64_t)batch_stride * 2};
    const cuuint32_t box[4] = {(cuuint32_t)kD, (cuuint32_t)rows, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline cudaError_t launch(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, __nv_bfloat16* out, int64_t batch, int heads, int s_q,
                          int s_kv, int64_t q_bs, int64_t q_rs, int64_t kv_bs, int64_t kv_rs, float scale_log2e, cudaStream_t st)
{
    static bool attr_done[64] = {};
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (!attr_done[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(k_attention_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) return e;
        attr_done[dev & 63] = true;
    }
    CUtensorMap mq, mk, mv;
    if (!make_map(&mq, q, batch, heads, s_q, q_bs, q_rs, kM) || !make_map(&mk, k, batch, heads, s_kv, kv_bs, kv_rs, kN) ||
        !make_map(&mv, v, batch, heads, s_kv, kv_bs, kv_rs, kN))
        return cudaErrorInvalidValue;
    const int total = (int)(batch * heads);
    const int need = (total + kWG - 1) / kWG;
    const unsigned blocks = (unsigned)(need < sms ? need : sms);
    k_attention_tc<<<blocks, kThreads, kSmemBytes, st>>>(mq, mk, mv, out, total, heads, s_q, s_kv, scale_log2e);
    return cudaGetLastError();
}

}  // namespace phe_attn_tc
