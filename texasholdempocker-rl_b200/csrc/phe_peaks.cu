// phe_peaks.cu -- on-box microbenchmarks for the roofline denominators of this integer / shared-memory path
// (SURVEY 8d asks for them to be MEASURED): alu-pipe stream (LOP3+SHF), balanced alu+fma stream (LOP3+IMAD),
// POPC, conflict-free LDS, random u16 LDS into a 128 KB table, random ATOMS.POPC.INC into 7463 bins.
// Prints one JSON object.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o phe_peaks phe_peaks.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int kThreads = 1024;
constexpr int kChains = 8;
constexpr int kUnroll = 16;

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) k_peak(uint32_t* out, int iters, uint32_t seed)
{
    extern __shared__ uint32_t sm[];
    const uint32_t tid = threadIdx.x;
    uint32_t v[kChains];
#pragma unroll
    for (int c = 0; c < kChains; c++) v[c] = seed * (c + 1) + tid * 2654435761u;
    if (MODE >= 3) {
        for (uint32_t i = tid; i < 40960; i += kThreads) sm[i] = i * 2654435761u + seed;   // 160 KB
        __syncthreads();
    }
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
    const uint32_t k = seed | 1u;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
#pragma unroll
            for (int c = 0; c < kChains; c++) {
                if (MODE == 0) {   // 2 alu-pipe ops
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[c]) : "r"(k), "r"(tid));
                    asm volatile("shf.l.wrap.b32 %0, %0, %0, 7;" : "+r"(v[c]));   // ptxas turns add.u32 into IMAD.IADD (fma pipe), SHF stays on alu
                } else if (MODE == 1) {   // 1 alu-pipe + 1 fma-pipe op
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[c]) : "r"(k), "r"(tid));
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[c]) : "r"(k), "r"(tid));
                } else if (MODE == 2) {   // popc + add
                    uint32_t p;
                    asm volatile("popc.b32 %0, %1;" : "=r"(p) : "r"(v[c]));
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(v[c]) : "r"(p));
                } else if (MODE == 3) {   // conflict-free LDS.32: lane-consecutive words
                    uint32_t a = sbase + ((((v[c] >> 8) & 0x3FFu) * 32u + (tid & 31u)) << 2), r;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a));
                    v[c] += r;
                } else if (MODE == 4) {   // random LDS.U16 in a 128 KB window
                    uint32_t a = sbase + ((v[c] >> 15) & 0x1FFFEu), r;
                    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(r) : "r"(a));
                    v[c] = v[c] * 1664525u + r;
                } else if (MODE == 5) {   // random shared atomic increment over 7463 words
                    uint32_t a = sbase + 131072u + ((__umulhi(v[c], 7463u)) << 2);
                    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a) : "memory");
                    v[c] = v[c] * 1664525u + 1013904223u;
                }
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < kChains; c++) acc ^= v[c];
    if (acc == 0x12345678u) out[blockIdx.x * kThreads + tid] = acc;
}

template <int MODE>
double run(int sms, int iters, uint32_t* d_out, float* ms_out)
{
    const size_t smem = MODE >= 3 ? 163840 + 32768 : 0;
    CK(cudaFuncSetAttribute(k_peak<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int w = 0; w < 2; w++) k_peak<MODE><<<sms, kThreads, smem>>>(d_out, iters, 12345u + w);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        CK(cudaEventRecord(a));
        k_peak<MODE><<<sms, kThreads, smem>>>(d_out, iters, 777u + r);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    *ms_out = best;
    const double inner = (double)sms * kThreads * (double)iters * kUnroll * kChains;
    return inner / (best * 1e-3);   // inner-loop bodies per second (lane granularity)
}

int main()
{
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    uint32_t* d_out;
    CK(cudaMalloc(&d_out, (size_t)sms * kThreads * 4));
    float ms[6];
    const double r0 = run<0>(sms, 2000, d_out, &ms[0]);
    const double r1 = run<1>(sms, 2000, d_out, &ms[1]);
    const double r2 = run<2>(sms, 1000, d_out, &ms[2]);
    const double r3 = run<3>(sms, 500, d_out, &ms[3]);
    const double r4 = run<4>(sms, 500, d_out, &ms[4]);
    const double r5 = run<5>(sms, 200, d_out, &ms[5]);
    int clk = 0;
    CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d, "
           "\"alu_lane_ops_per_s\": %.4e, \"balanced_lane_ops_per_s\": %.4e, \"popc_add_pairs_per_s\": %.4e, "
           "\"lds32_conflict_free_per_s\": %.4e, \"lds16_random_128k_per_s\": %.4e, \"atoms_inc_random_7463_per_s\": %.4e, "
           "\"ms\": [%.3f, %.3f, %.3f, %.3f, %.3f, %.3f], "
           "\"note\": \"lane-ops/s: alu = LOP3+SHF (2 ops/body), balanced = LOP3+IMAD (2 ops/body); lookups and atomics are per lane\"}\n",
           p.name, sms, clk, r0 * 2, r1 * 2, r2, r3, r4, r5, ms[0], ms[1], ms[2], ms[3], ms[4], ms[5]);
    return 0;
}
