// phe_nn.cu -- memory-bound glue kernels of the policy / value network, hand-written for sm_100a.
// The GEMMs and the attention of PolicyValueNet are library calls (cuBLASLt / cuDNN through PyTorch); what PyTorch
// leaves on the table are the HBM-bound passes between them, which these two kernels replace:
//
//   phe_nn_add_layernorm   y = LayerNorm(x + residual) * gamma + beta      (nn.TransformerEncoderLayer post-LN,
//                          rl/models.py:446-447).  One warp per row, the row stays in registers between the two
//                          reductions: 3 row-passes of traffic (read x, read residual, write y) instead of the 5 of
//                          add + layer_norm; ATen's vectorized_layer_norm_kernel alone ran at 1.3 TB/s on [100352, 640].
//   phe_nn_embed_obs       encoder input of EnvEmbedding (rl/models.py:100-114) + the sqrt(d) scaling (:462-466):
//                          out[b, s, :] = starters[s]                                   s <  n_starters
//                                       = field_tab[obs[b, f] + start[f]] | pos_tab[f]   f = s - n_starters
//                          with the scaling folded into the tables: one gather-copy instead of gather, mul, two cats.
//
// Algorithmic HBM bytes: add_layernorm 3 * d * sizeof(T) per row (2 without residual); embed_obs d * sizeof(T) per token
// written + 4 B per field read.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <mutex>

#include "../../include/phe_b200.h"
#include "phe_attn_tc.cuh"

void phe_internal_count_launch();          // phe_kernels.cu

namespace {

int nn_after_launch()
{
    phe_internal_count_launch();
    return cudaGetLastError() == cudaSuccess ? PHE_OK : PHE_ECUDA;
}

template <class T> struct Vec;            // 16-byte vectors
template <> struct Vec<float> {
    static constexpr int N = 4;
    __device__ static void load(const float* p, float (&f)[4]) { const float4 v = *reinterpret_cast<const float4*>(p); f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
    __device__ static void load_stream(const float* p, float (&f)[4]) { const float4 v = __ldcs(reinterpret_cast<const float4*>(p)); f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
    __device__ static void store(float* p, const float (&f)[4]) { *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]); }
};
template <> struct Vec<__nv_bfloat16> {
    static constexpr int N = 8;
    __device__ static void unpack(const uint4& v, float (&f)[8])
    {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; i++) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u); }
    }
    __device__ static void load(const __nv_bfloat16* p, float (&f)[8]) { unpack(*reinterpret_cast<const uint4*>(p), f); }
    __device__ static void load_stream(const __nv_bfloat16* p, float (&f)[8]) { unpack(__ldcs(reinterpret_cast<const uint4*>(p)), f); }
    __device__ static void store(__nv_bfloat16* p, const float (&f)[8])
    {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
            w[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
};

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// VPL = 16-byte vectors per lane (the row has at most 32 * VPL of them); rows are contiguous (stride d)
template <class T, int VPL>
__global__ void __launch_bounds__(256)
k_add_layernorm(const T* __restrict__ x, const T* __restrict__ res, const T* __restrict__ gamma, const T* __restrict__ beta,
                T* __restrict__ out, int64_t rows, int d, float eps)
{
    constexpr int N = Vec<T>::N;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int nvec = d / N;
    for (int64_t row = warp; row < rows; row += nwarps) {
        const T* xr = x + row * d;
        float v[VPL][N];
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < VPL; k++) {
            const int j = lane + 32 * k;
            if (j < nvec) {
                Vec<T>::load_stream(xr + j * N, v[k]);
                if (res) {
                    float r[N];
                    Vec<T>::load_stream(res + row * d + j * N, r);
#pragma unroll
                    for (int i = 0; i < N; i++) v[k][i] += r[i];
                }
#pragma unroll
                for (int i = 0; i < N; i++) s += v[k][i];
            }
        }
        const float mean = warp_sum(s) / (float)d;
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < VPL; k++) {
            if (lane + 32 * k < nvec) {
#pragma unroll
                for (int i = 0; i < N; i++) { const float c = v[k][i] - mean; q += c * c; }
            }
        }
        const float rstd = rsqrtf(warp_sum(q) / (float)d + eps);
#pragma unroll
        for (int k = 0; k < VPL; k++) {
            const int j = lane + 32 * k;
            if (j < nvec) {
                float g[N], b[N], o[N];
                Vec<T>::load(gamma + j * N, g);
                Vec<T>::load(beta + j * N, b);
#pragma unroll
                for (int i = 0; i < N; i++) o[i] = (v[k][i] - mean) * rstd * g[i] + b[i];
                Vec<T>::store(out + row * d + j * N, o);
            }
        }
    }
}

template <class T>
int launch_add_layernorm(const void* x, const void* res, const void* gamma, const void* beta, void* out, int64_t rows, int d, float eps,
                         cudaStream_t st)
{
    constexpr int N = Vec<T>::N;
    if (d % N) return PHE_EINVAL;
    const int nvec = d / N, vpl = (nvec + 31) / 32;
    int sms = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = (rows + 7) / 8;
    const unsigned blocks = (unsigned)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);      // 8 CTAs of 8 warps per SM, grid-stride
    const T *xp = static_cast<const T*>(x), *rp = static_cast<const T*>(res), *gp = static_cast<const T*>(gamma), *bp = static_cast<const T*>(beta);
    T* op = static_cast<T*>(out);
    if (vpl <= 1) k_add_layernorm<T, 1><<<blocks, 256, 0, st>>>(xp, rp, gp, bp, op, rows, d, eps);
    else if (vpl <= 2) k_add_layernorm<T, 2><<<blocks, 256, 0, st>>>(xp, rp, gp, bp, op, rows, d, eps);
    else if (vpl <= 3) k_add_layernorm<T, 3><<<blocks, 256, 0, st>>>(xp, rp, gp, bp, op, rows, d, eps);
    else if (vpl <= 4) k_add_layernorm<T, 4><<<blocks, 256, 0, st>>>(xp, rp, gp, bp, op, rows, d, eps);
    else if (vpl <= 8) k_add_layernorm<T, 8><<<blocks, 256, 0, st>>>(xp, rp, gp, bp, op, rows, d, eps);
    else return PHE_EINVAL;                                                                           // d > 2048 (bf16) / 1024 (f32)
    return nn_after_launch();
}

// one warp per token; rows of the three tables are multiples of 16 bytes
template <class T>
__global__ void __launch_bounds__(256)
k_embed_obs(const int32_t* __restrict__ obs, const int32_t* __restrict__ field_start, const T* __restrict__ field_tab,
            const T* __restrict__ pos_tab, const T* __restrict__ starters, T* __restrict__ out, int64_t batch, int n_fields,
            int n_starters, int de, int dp, int n_rows)
{
    constexpr int N = 16 / (int)sizeof(T);
    const int lane = threadIdx.x & 31;
    const int seq = n_starters + n_fields, d = de + dp;
    const int64_t ntok = batch * seq;
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int nve = de / N, nvd = d / N;
    for (int64_t tok = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tok < ntok; tok += nwarps) {
        const int64_t b = tok / seq;
        const int s = (int)(tok - b * seq);
        uint4* dst = reinterpret_cast<uint4*>(out + tok * d);
        if (s < n_starters) {
            const uint4* src = reinterpret_cast<const uint4*>(starters + (int64_t)s * d);
            for (int j = lane; j < nvd; j += 32) __stcs(dst + j, __ldg(src + j));
        } else {
            const int f = s - n_starters;
            int idx = __ldg(obs + b * n_fields + f) + __ldg(field_start + f);
            idx = idx < 0 ? 0 : (idx >= n_rows ? n_rows - 1 : idx);            // memory safety; vocabulary checks are the caller's
            const uint4* se = reinterpret_cast<const uint4*>(field_tab + (int64_t)idx * de);
            const uint4* sp = reinterpret_cast<const uint4*>(pos_tab + (int64_t)f * dp);
            for (int j = lane; j < nvd; j += 32) __stcs(dst + j, j < nve ? __ldg(se + j) : __ldg(sp + (j - nve)));
        }
    }
}

template <class T>
int launch_embed_obs(const int32_t* obs, const int32_t* field_start, const void* field_tab, const void* pos_tab, const void* starters,
                     void* out, int64_t batch, int n_fields, int n_starters, int de, int dp, int n_rows, cudaStream_t st)
{
    constexpr int N = 16 / (int)sizeof(T);
    if (de % N || dp % N) return PHE_EINVAL;
    const int64_t ntok = batch * (n_starters + n_fields);
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = (ntok + 7) / 8;
    const unsigned blocks = (unsigned)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);
    k_embed_obs<T><<<blocks, 256, 0, st>>>(obs, field_start, static_cast<const T*>(field_tab), static_cast<const T*>(pos_tab),
                                           static_cast<const T*>(starters), static_cast<T*>(out), batch, n_fields, n_starters, de, dp, n_rows);
    return nn_after_launch();
}

}  // namespace

extern "C" {

int phe_nn_add_layernorm(const void* x, const void* residual, const void* gamma, const void* beta, void* out, int64_t rows, int d,
                         float eps, int dtype, void* stream)
{
    if (rows < 0 || d <= 0 || (dtype != PHE_DTYPE_F32 && dtype != PHE_DTYPE_BF16)) return PHE_EINVAL;
    if (rows == 0) return PHE_OK;
    if (!x || !gamma || !beta || !out) return PHE_EINVAL;
    const uintptr_t al = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(residual) | reinterpret_cast<uintptr_t>(gamma) |
                         reinterpret_cast<uintptr_t>(beta) | reinterpret_cast<uintptr_t>(out);
    if (al & 15) return PHE_EINVAL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return dtype == PHE_DTYPE_BF16 ? launch_add_layernorm<__nv_bfloat16>(x, residual, gamma, beta, out, rows, d, eps, st)
                                   : launch_add_layernorm<float>(x, residual, gamma, beta, out, rows, d, eps, st);
}

int phe_nn_embed_obs(const int32_t* obs, const int32_t* field_start, const void* field_tab, const void* pos_tab, const void* starters,
                     void* out, int64_t batch, int n_fields, int n_starters, int de, int dp, int n_rows, int dtype, void* stream)
{
    if (batch < 0 || n_fields <= 0 || n_starters < 0 || de <= 0 || dp < 0 || n_rows <= 0 || (dtype != PHE_DTYPE_F32 && dtype != PHE_DTYPE_BF16))
        return PHE_EINVAL;
    if (batch == 0) return PHE_OK;
    if (!obs || !field_start || !field_tab || !out || (dp > 0 && !pos_tab) || (n_starters > 0 && !starters)) return PHE_EINVAL;
    const uintptr_t al = reinterpret_cast<uintptr_t>(field_tab) | reinterpret_cast<uintptr_t>(pos_tab) | reinterpret_cast<uintptr_t>(starters) |
                         reinterpret_cast<uintptr_t>(out);
    if (al & 15) return PHE_EINVAL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return dtype == PHE_DTYPE_BF16 ? launch_embed_obs<__nv_bfloat16>(obs, field_start, field_tab, pos_tab, starters, out, batch, n_fields, n_starters, de, dp, n_rows, st)
                                   : launch_embed_obs<float>(obs, field_start, field_tab, pos_tab, starters, out, batch, n_fields, n_starters, de, dp, n_rows, st);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// phe_nn_attention: softmax(Q K^T / sqrt(D)) V for the short sequences of this model (98 tokens, <= 112), bf16.
// One CTA of four warps per (batch, head): K, V and Q of that head are staged once in shared memory (<= 48 KB), each
// warp takes 16-row query tiles, keeps the whole 16 x 112 score tile in registers (no online-softmax rescaling is
// needed at this length), and feeds the probabilities straight back as the A operand of the P V product
// (mma.sync.m16n8k16, ldmatrix / ldmatrix.trans operands).  The generic flash kernels of cuDNN / FA2 run this shape at
// 2.6x its memory floor; here the traffic is the algorithmic one: Q, K, V read once, O written once
// ((2 S_q + 2 S_kv) * D * 2 bytes per head).
// ---------------------------------------------------------------------------------------------------------------
namespace {

constexpr int kAttnMaxKeys = 112;       // 7 k-steps of 16 keys
constexpr int kAttnWarps = 8;          // 98 query rows = 7 tiles of 16: one round, 7 of the 8 warps busy in the MMA phase

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ float ex2(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi)
{
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// D = head dimension (16 or 64).  Rows of the shared tiles are padded by 8 elements (16 bytes): the 8 row addresses
// of an ldmatrix then fall into distinct bank groups.  Persistent CTAs (two per SM) walk the (batch, head) pairs with a
// two-stage shared-memory pipeline: the cp.async copies of the next pair are in flight while the current one is in the
// MMA phase -- with one-shot CTAs the load and MMA phases of the four resident CTAs alternated and the kernel ran at
// half of the memory rate.
#ifndef PHE_ATTN_STAGES
#define PHE_ATTN_STAGES 2      // 2: persistent CTAs, next pair prefetched; 1: one CTA per pair (grid = pairs)
#endif

template <int D>
struct AttnSmem {
    static constexpr int LD = D + 8;
    static constexpr int kTile = kAttnMaxKeys * LD;            // elements of one K / V / Q tile
    static constexpr uint32_t kBytes = PHE_ATTN_STAGES * 3u * kTile * 2u;   // stages of {K, V, Q}
};

#ifndef PHE_ATTN_GRID_PER_SM
#define PHE_ATTN_GRID_PER_SM 2
#endif
#ifndef PHE_ATTN_MIN_CTAS
#define PHE_ATTN_MIN_CTAS 2
#endif

template <int D>
__global__ void __launch_bounds__(kAttnWarps * 32, PHE_ATTN_MIN_CTAS)
k_attention(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
            __nv_bfloat16* __restrict__ out, int total_bh, int heads, int s_q, int s_kv, int64_t q_bs, int64_t q_rs, int64_t kv_bs,
            int64_t kv_rs, float scale_log2e)
{
    constexpr int LD = AttnSmem<D>::LD;
    constexpr int kTile = AttnSmem<D>::kTile;
    constexpr int VPR = D / 8;                 // 16-byte vectors per row
    extern __shared__ __align__(16) unsigned char attn_smem[];
    __nv_bfloat16* sm = reinterpret_cast<__nv_bfloat16*>(attn_smem);          // stage st: K at (3 st) kTile, V at (3 st + 1) kTile, Q at (3 st + 2) kTile
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q_tiles = (s_q + 15) >> 4;
    // rows beyond the sequence stay zero for the whole launch (the copies below never touch them)
    for (int i = tid; i < PHE_ATTN_STAGES * kAttnMaxKeys * VPR; i += blockDim.x) {
        const int st = i / (kAttnMaxKeys * VPR), j = i - st * (kAttnMaxKeys * VPR);
        const int r = j / VPR, c = j - r * VPR;
        __nv_bfloat16* base = sm + 3 * st * kTile + r * LD + c * 8;
        if (r >= s_kv) {
            *reinterpret_cast<uint4*>(base) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(base + kTile) = make_uint4(0, 0, 0, 0);
        }
        if (r >= s_q) *reinterpret_cast<uint4*>(base + 2 * kTile) = make_uint4(0, 0, 0, 0);
    }
    auto issue = [&](int bh, int st) {
        const int b = bh / heads, h = bh - b * heads;
        const __nv_bfloat16* kb = k + (int64_t)b * kv_bs + (int64_t)h * D;
        const __nv_bfloat16* vb = v + (int64_t)b * kv_bs + (int64_t)h * D;
        const __nv_bfloat16* qb = q + (int64_t)b * q_bs + (int64_t)h * D;
        __nv_bfloat16* sk = sm + 3 * st * kTile;
        for (int i = tid; i < s_kv * VPR; i += blockDim.x) {
            const int r = i / VPR, c = i - r * VPR;
            cp_async16(sk + r * LD + c * 8, reinterpret_cast<const uint4*>(kb + (int64_t)r * kv_rs) + c);
            cp_async16(sk + kTile + r * LD + c * 8, reinterpret_cast<const uint4*>(vb + (int64_t)r * kv_rs) + c);
        }
        for (int i = tid; i < s_q * VPR; i += blockDim.x) {
            const int r = i / VPR, c = i - r * VPR;
            cp_async16(sk + 2 * kTile + r * LD + c * 8, reinterpret_cast<const uint4*>(qb + (int64_t)r * q_rs) + c);
        }
    };
    const int g = lane >> 2, t = lane & 3;
    constexpr int NT = kAttnMaxKeys / 8;       // 14 key tiles of 8
    constexpr int KS = D / 16;                 // k-steps of Q K^T
    constexpr int DT = D / 8;
    const int nt_live = (s_kv + 7) >> 3;       // key tiles that hold at least one real key
    int bh = blockIdx.x;
    if (bh < total_bh) issue(bh, 0);
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int it = 0; bh < total_bh; bh += gridDim.x, it++) {
        const int st = it & 1;
        // One barrier per pair: the copies of this stage were issued an iteration ago (a whole MMA phase to land), and
        // once every warp is past the barrier nobody still reads the other stage, so the next pair's copies go there.
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (bh + (int)gridDim.x < total_bh) issue(bh + gridDim.x, st ^ 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const uint32_t a_k = (uint32_t)__cvta_generic_to_shared(sm + 3 * st * kTile), a_v = a_k + 2u * kTile, a_q = a_k + 4u * kTile;
        const int b = bh / heads, h = bh - b * heads;
        // 98 rows are 7 tiles for 8 warps: the idle slot rotates (per iteration and per CTA) so that the tensor pipe of
        // every scheduler gets the same share -- with a fixed map the scheduler of warp 7 ran at half load
        for (int qt = (warp + it + (int)blockIdx.x) & (kAttnWarps - 1); qt < q_tiles; qt += kAttnWarps) {
            // ---- S = Q K^T ----
            uint32_t aq[KS][4];
#pragma unroll
            for (int ks = 0; ks < KS; ks++)
                ldmatrix_x4(aq[ks], a_q + 2u * (uint32_t)((qt * 16 + (lane & 15)) * LD + ks * 16 + (lane >> 4) * 8));
            float s[NT][4];
#pragma unroll
            for (int nt = 0; nt < NT; nt++) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
            // k-step outermost: consecutive HMMAs write different accumulators (14 independent chains instead of 4-deep
            // dependent ones; the kernel was stalled on HMMA latency with only four warps per scheduler)
#pragma unroll
            for (int ks = 0; ks < KS; ks++) {
#pragma unroll
                for (int nt = 0; nt < NT; nt += 2) {
                    if (nt < nt_live) {
                        // one ldmatrix.x4 = the B fragments of two key tiles: (keys nt*8.., dims lo / hi), (keys (nt+1)*8.., dims lo / hi)
                        uint32_t bk[4];
                        ldmatrix_x4(bk, a_k + 2u * (uint32_t)((nt * 8 + (lane & 7) + ((lane >> 4) & 1) * 8) * LD + ks * 16 + ((lane >> 3) & 1) * 8));
                        const uint32_t b0[2] = {bk[0], bk[1]}, b1[2] = {bk[2], bk[3]};
                        mma_bf16(s[nt], aq[ks], b0);
                        mma_bf16(s[nt + 1], aq[ks], b1);
                    }
                }
            }
            // ---- softmax over the keys (rows g and g + 8 of the tile; a row lives in the four lanes of a quad) ----
            // Padded key columns (index >= s_kv; K = V = 0) are masked out of both the row maximum and the row sum, so a
            // row whose real logits are all strongly negative keeps its full precision (a closed-form correction of the
            // padded terms cancels catastrophically there).  Column of s[nt][j]: nt * 8 + 2 t + (j & 1).
            const int lim = s_kv - 2 * t;              // s[nt][0|2] is real iff nt * 8 < lim, s[nt][1|3] iff nt * 8 + 1 < lim
            float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < NT; nt++) {
                if (nt * 8 < lim) { m0 = fmaxf(m0, s[nt][0]); m1 = fmaxf(m1, s[nt][2]); }
                if (nt * 8 + 1 < lim) { m0 = fmaxf(m0, s[nt][1]); m1 = fmaxf(m1, s[nt][3]); }
            }
            m0 = fmaxf(m0, __shfl_xor_sync(0xFFFFFFFFu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xFFFFFFFFu, m0, 2));
            m1 = fmaxf(m1, __shfl_xor_sync(0xFFFFFFFFu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xFFFFFFFFu, m1, 2));
            const float c0 = m0 * scale_log2e, c1 = m1 * scale_log2e;       // exp2(s * scale - max * scale): one FFMA + one MUFU per score
            float l0 = 0.f, l1 = 0.f;
            uint32_t p[NT][2];                       // probabilities as bf16 pairs: [nt][0] row g, [nt][1] row g + 8
#pragma unroll
            for (int nt = 0; nt < NT; nt++) {
                const bool ra = nt * 8 < lim, rb = nt * 8 + 1 < lim;
                const float e0 = ra ? ex2(fmaf(s[nt][0], scale_log2e, -c0)) : 0.f, e1 = rb ? ex2(fmaf(s[nt][1], scale_log2e, -c0)) : 0.f;
                const float e2 = ra ? ex2(fmaf(s[nt][2], scale_log2e, -c1)) : 0.f, e3 = rb ? ex2(fmaf(s[nt][3], scale_log2e, -c1)) : 0.f;
                l0 += e0 + e1;
                l1 += e2 + e3;
                p[nt][0] = pack_bf16(e0, e1);
                p[nt][1] = pack_bf16(e2, e3);
            }
            l0 += __shfl_xor_sync(0xFFFFFFFFu, l0, 1); l0 += __shfl_xor_sync(0xFFFFFFFFu, l0, 2);
            l1 += __shfl_xor_sync(0xFFFFFFFFu, l1, 1); l1 += __shfl_xor_sync(0xFFFFFFFFu, l1, 2);
            const float inv0 = 1.f / l0, inv1 = 1.f / l1;             // >= 1: the row maximum is a real key
            // ---- O = P V: the score accumulators of key tiles (2j, 2j + 1) are exactly the A fragment of k-step j ----
            float o[DT][4];
#pragma unroll
            for (int dt = 0; dt < DT; dt++) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f;
#pragma unroll
            for (int j = 0; j < NT / 2; j++) {
                if (2 * j < nt_live) {
                    const uint32_t ap[4] = {p[2 * j][0], p[2 * j][1], p[2 * j + 1][0], p[2 * j + 1][1]};
#pragma unroll
                    for (int dt = 0; dt < DT; dt += 2) {
                        // transposed x4: (keys lo / hi, dims dt*8..) and (keys lo / hi, dims (dt+1)*8..)
                        uint32_t bv[4];
                        ldmatrix_x4_trans(bv, a_v + 2u * (uint32_t)((j * 16 + (lane & 15)) * LD + dt * 8 + ((lane >> 4) & 1) * 8));
                        const uint32_t b0[2] = {bv[0], bv[1]}, b1[2] = {bv[2], bv[3]};
                        mma_bf16(o[dt], ap, b0);
                        mma_bf16(o[dt + 1], ap, b1);
                    }
                }
            }
            const int r0 = qt * 16 + g, r1 = r0 + 8;
            __nv_bfloat16* ob = out + ((int64_t)b * s_q) * heads * D + (int64_t)h * D;
#pragma unroll
            for (int dt = 0; dt < DT; dt++) {
                if (r0 < s_q) *reinterpret_cast<uint32_t*>(ob + (int64_t)r0 * heads * D + dt * 8 + 2 * t) = pack_bf16(o[dt][0] * inv0, o[dt][1] * inv0);
                if (r1 < s_q) *reinterpret_cast<uint32_t*>(ob + (int64_t)r1 * heads * D + dt * 8 + 2 * t) = pack_bf16(o[dt][2] * inv1, o[dt][3] * inv1);
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// phe_nn_policy_value_heads: what SingleThreadMCTS.predict reads (rl/MCTS.py:464-473) from encoder rows 0 and 1:
// out[b] = { softmax(W_p h[b][0] + b_p) (A values), w_v . h[b][1] + b_v } as fp32.  One warp per observation; replaces
// two GEMV launches, the softmax, two casts and two strided copies (eight ~3 us kernels in the small-batch regime the
// reference runs in).
// ---------------------------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(128)
k_policy_value_heads(const T* __restrict__ h, const T* __restrict__ wp, const T* __restrict__ bp, const T* __restrict__ wv,
                     const T* __restrict__ bv, float* __restrict__ out, int64_t batch, int d, int n_actions, int64_t row_stride)
{
    constexpr int N = Vec<T>::N;
    const int lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= batch) return;
    const T* h0 = h + b * 2 * row_stride;
    const T* h1 = h0 + row_stride;
    const int nvec = d / N;
    float logit = -INFINITY;                 // lane a keeps logit a (n_actions <= 32)
    for (int a = 0; a <= n_actions; a++) {   // a == n_actions: the value head on row 1
        const T* w = a < n_actions ? wp + (int64_t)a * d : wv;
        const T* x = a < n_actions ? h0 : h1;
        float acc = 0.f;
        for (int j = lane; j < nvec; j += 32) {
            float xv[N], wvv[N];
            Vec<T>::load(x + j * N, xv);
            Vec<T>::load(w + j * N, wvv);
#pragma unroll
            for (int i = 0; i < N; i++) acc += xv[i] * wvv[i];
        }
        acc = warp_sum(acc);
        if (a < n_actions) {
            float bias[1];
            bias[0] = sizeof(T) == 2 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(bp)[a]) : reinterpret_cast<const float*>(bp)[a];
            // the reference's Linear output is rounded to the compute dtype before the fp32 softmax
            float z = acc + bias[0];
            if (sizeof(T) == 2) z = __bfloat162float(__float2bfloat16_rn(z));
            if (lane == a) logit = z;
        } else if (lane == 0) {
            float z = acc + (sizeof(T) == 2 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(bv)[0]) : reinterpret_cast<const float*>(bv)[0]);
            if (sizeof(T) == 2) z = __bfloat162float(__float2bfloat16_rn(z));
            out[b * (n_actions + 1) + n_actions] = z;
        }
    }
    float m = logit;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
    const float e = lane < n_actions ? expf(logit - m) : 0.f;
    const float ssum = warp_sum(e);
    if (lane < n_actions) out[b * (n_actions + 1) + lane] = e / ssum;
}

}  // namespace

extern "C" int phe_nn_policy_value_heads(const void* h, const void* w_policy, const void* b_policy, const void* w_value, const void* b_value,
                                         float* out, int64_t batch, int d, int n_actions, int64_t row_stride, int dtype, void* stream)
{
    if (batch < 0 || d <= 0 || n_actions <= 0 || n_actions > 32 || (dtype != PHE_DTYPE_F32 && dtype != PHE_DTYPE_BF16)) return PHE_EINVAL;
    if (batch == 0) return PHE_OK;
    if (!h || !w_policy || !b_policy || !w_value || !b_value || !out) return PHE_EINVAL;
    const int esz = dtype == PHE_DTYPE_BF16 ? 2 : 4;
    const uintptr_t al = reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(w_policy) | reinterpret_cast<uintptr_t>(w_value);
    if ((al & 15) || (d * esz) % 16 || (row_stride * esz) % 16) return PHE_EINVAL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned blocks = (unsigned)((batch + 3) / 4);
    if (dtype == PHE_DTYPE_BF16)
        k_policy_value_heads<__nv_bfloat16><<<blocks, 128, 0, st>>>(static_cast<const __nv_bfloat16*>(h), static_cast<const __nv_bfloat16*>(w_policy),
                                                                     static_cast<const __nv_bfloat16*>(b_policy), static_cast<const __nv_bfloat16*>(w_value),
                                                                     static_cast<const __nv_bfloat16*>(b_value), out, batch, d, n_actions, row_stride);
    else
        k_policy_value_heads<float><<<blocks, 128, 0, st>>>(static_cast<const float*>(h), static_cast<const float*>(w_policy), static_cast<const float*>(b_policy),
                                                             static_cast<const float*>(w_value), static_cast<const float*>(b_value), out, batch, d, n_actions, row_stride);
    return nn_after_launch();
}

extern "C" int phe_nn_attention(const void* q, const void* k, const void* v, void* out, int64_t batch, int heads, int s_q, int s_kv,
                                int head_dim, int64_t q_batch_stride, int64_t q_row_stride, int64_t kv_batch_stride, int64_t kv_row_stride,
                                int dtype, void* stream)
{
    if (batch < 0 || heads <= 0 || s_q <= 0 || s_kv <= 0 || s_q > kAttnMaxKeys || s_kv > kAttnMaxKeys || dtype != PHE_DTYPE_BF16) return PHE_EINVAL;
    if (head_dim != 16 && head_dim != 64) return PHE_EINVAL;
    if (batch == 0) return PHE_OK;
    if (!q || !k || !v || !out) return PHE_EINVAL;
    if (batch * heads > 0x7FFFFFFFll) return PHE_EINVAL;
    const uintptr_t al = reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(out);
    if ((al & 15) || ((q_batch_stride | q_row_stride | kv_batch_stride | kv_row_stride) & 7)) return PHE_EINVAL;      // 16-byte rows
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float scale_log2e = 1.4426950408889634f / sqrtf((float)head_dim);
    const int total = (int)(batch * heads);
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned blocks = (unsigned)(total < PHE_ATTN_GRID_PER_SM * sms ? total : PHE_ATTN_GRID_PER_SM * sms);
    static std::once_flag attr_once[16];
    std::call_once(attr_once[dev & 15], [] {
        cudaFuncSetAttribute(k_attention<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem<64>::kBytes);
        cudaFuncSetAttribute(k_attention<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem<16>::kBytes);
    });
    const auto *qp = static_cast<const __nv_bfloat16*>(q), *kp = static_cast<const __nv_bfloat16*>(k), *vp = static_cast<const __nv_bfloat16*>(v);
    auto* op = static_cast<__nv_bfloat16*>(out);
#ifndef PHE_ATTN_LEGACY
    if (head_dim == 64) {
        // head dim 64 (config/default.yml): the tcgen05 / TMEM kernel (csrc/phe_attn_tc.cuh)
        const cudaError_t e = phe_attn_tc::launch(qp, kp, vp, op, batch, heads, s_q, s_kv, q_batch_stride, q_row_stride, kv_batch_stride, kv_row_stride,
                                                  scale_log2e, st);
        if (e != cudaSuccess) { cudaGetLastError(); return PHE_ECUDA; }
        phe_internal_count_launch();
        return PHE_OK;
    }
#endif
    if (head_dim == 64)
        k_attention<64><<<blocks, kAttnWarps * 32, AttnSmem<64>::kBytes, st>>>(qp, kp, vp, op, total, heads, s_q, s_kv, q_batch_stride, q_row_stride, kv_batch_stride, kv_row_stride, scale_log2e);
    else
        k_attention<16><<<blocks, kAttnWarps * 32, AttnSmem<16>::kBytes, st>>>(qp, kp, vp, op, total, heads, s_q, s_kv, q_batch_stride, q_row_stride, kv_batch_stride, kv_row_stride, scale_log2e);
    return nn_after_launch();
}
