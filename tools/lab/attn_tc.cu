// attn_tc.cu -- standalone lab for the tcgen05 / TMEM attention kernel (head dim 64, <= 112 keys, bf16).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/lab/attn_tc tools/lab/attn_tc.cu
//   tools/lab/attn_tc [batch] [heads] [seq]          (GPU box; prints max error against an fp32 reference kernel and the time)
//
// The kernel under test is the one compiled into the library (csrc/phe_attn_tc.cuh); this file only adds a naive
// reference and a driver, so that descriptor / layout mistakes are found without the Python stack.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../texasholdempocker-rl_b200/csrc/phe_attn_tc.cuh"

__global__ void k_ref(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, float* out, int B, int H, int S, int D, int64_t bs, int64_t rs)
{
    const int bh = blockIdx.x, b = bh / H, h = bh % H, r = threadIdx.x;
    if (r >= S) return;
    const __nv_bfloat16* qr = q + b * bs + (int64_t)r * rs + h * D;
    float sc[128], m = -INFINITY;
    for (int j = 0; j < S; j++) {
        const __nv_bfloat16* kr = k + b * bs + (int64_t)j * rs + h * D;
        float a = 0;
        for (int d = 0; d < D; d++) a += __bfloat162float(qr[d]) * __bfloat162float(kr[d]);
        sc[j] = a / sqrtf((float)D);
        m = fmaxf(m, sc[j]);
    }
    float l = 0;
    for (int j = 0; j < S; j++) { sc[j] = expf(sc[j] - m); l += sc[j]; }
    for (int d = 0; d < D; d++) {
        float o = 0;
        for (int j = 0; j < S; j++) o += sc[j] * __bfloat162float(v[b * bs + (int64_t)j * rs + h * D + d]);
        out[((int64_t)(b * S + r) * H + h) * D + d] = o / l;
    }
}

int main(int argc, char** argv)
{
    const int B = argc > 1 ? atoi(argv[1]) : 64, H = argc > 2 ? atoi(argv[2]) : 10, S = argc > 3 ? atoi(argv[3]) : 98, D = 64;
    const int64_t rs = 3LL * H * D, bs = (int64_t)S * rs;
    const size_t n = (size_t)B * bs;
    std::vector<__nv_bfloat16> h(n);
    srand(1);
    for (size_t i = 0; i < n; i++) h[i] = __float2bfloat16(((rand() % 2001) - 1000) / 500.0f);
    __nv_bfloat16 *d_qkv, *d_out;
    float* d_ref;
    cudaMalloc(&d_qkv, n * 2);
    cudaMalloc(&d_out, (size_t)B * S * H * D * 2);
    cudaMalloc(&d_ref, (size_t)B * S * H * D * 4);
    cudaMemcpy(d_qkv, h.data(), n * 2, cudaMemcpyHostToDevice);
    cudaMemset(d_out, 0, (size_t)B * S * H * D * 2);
    const __nv_bfloat16 *q = d_qkv, *k = d_qkv + H * D, *v = d_qkv + 2 * H * D;
    k_ref<<<B * H, 128>>>(q, k, v, d_ref, B, H, S, D, bs, rs);
    cudaError_t e = phe_attn_tc::launch(q, k, v, d_out, B, H, S, S, bs, rs, bs, rs, 1.4426950408889634f / sqrtf((float)D), 0);
    printf("launch: %s\n", cudaGetErrorString(e));
    e = cudaDeviceSynchronize();
    printf("sync: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<__nv_bfloat16> o((size_t)B * S * H * D);
    std::vector<float> r((size_t)B * S * H * D);
    cudaMemcpy(o.data(), d_out, o.size() * 2, cudaMemcpyDeviceToHost);
    cudaMemcpy(r.data(), d_ref, r.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    size_t worst = 0;
    for (size_t i = 0; i < o.size(); i++) {
        const double d = fabs((double)__bfloat162float(o[i]) - r[i]);
        if (!(d <= maxerr)) { maxerr = d; worst = i; }
    }
    printf("B=%d H=%d S=%d: max abs err %.4g at %zu (got %.4f want %.4f)\n", B, H, S, maxerr, worst, __bfloat162float(o[worst]), r[worst]);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; i++) phe_attn_tc::launch(q, k, v, d_out, B, H, S, S, bs, rs, bs, rs, 1.4426950408889634f / sqrtf((float)D), 0);
    cudaEventRecord(e0);
    for (int i = 0; i < 20; i++) phe_attn_tc::launch(q, k, v, d_out, B, H, S, S, bs, rs, bs, rs, 1.4426950408889634f / sqrtf((float)D), 0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("time %.1f us per call (%.0f GB/s algorithmic)\n", ms / 20 * 1e3, (double)B * H * S * D * 2 * 4 / (ms / 20 * 1e-3) / 1e9);
    return maxerr < 0.05 ? 0 : 2;
}
