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

#include "../../include/phe_b200.h"

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
