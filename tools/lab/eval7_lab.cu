// eval7_lab.cu -- A/B bench of batched-evaluator kernel variants (tools/lab; not part of the product library).
// Builds the product's table image, evaluates 2^25 random card sets with each variant, checks every output against
// the baseline variant and times it with CUDA events (256 MB L2 flush between launches).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o eval7_lab eval7_lab.cu ../../texasholdempocker-rl_b200/csrc/phe_tables.cpp
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../texasholdempocker-rl_b200/csrc/phe_core.cuh"
#include "../../texasholdempocker-rl_b200/csrc/phe_tables.h"

using namespace phe;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

struct LabConst { HashParams hp; uint32_t two16, two17; };
__constant__ LabConst c_lab;
__device__ uint64_t* g_stamps = nullptr;

static constexpr int kThreads = 1024;
static constexpr uint32_t kSmem = TAB_BYTES + 16;

__device__ __forceinline__ void stage(unsigned char* s_dst, const void* g_src, uint32_t bytes, uint64_t* s_bar)
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
        const unsigned char* src = reinterpret_cast<const unsigned char*>(g_src);
        for (uint32_t off = 0; off < bytes; off += 32768) {
            const uint32_t n = bytes - off < 32768 ? bytes - off : 32768;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + off),
                         "l"(src + off), "r"(n), "r"(bar) : "memory");
        }
    }
}
// cluster variant: every CTA of the cluster copies its 1/CL share of the image into ALL CTAs' shared memory (TMA multicast),
// each CTA's mbarrier expects the whole image
template <int CL>
__device__ __forceinline__ void stage_mc(unsigned char* s_dst, const void* g_src, uint32_t bytes, uint64_t* s_bar)
{
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(s_bar);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_dst);
    uint32_t crank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    }
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x == 0) {
        const unsigned char* src = reinterpret_cast<const unsigned char*>(g_src);
        constexpr uint32_t kChunk = 8192;
        const uint16_t mask = (uint16_t)((1u << CL) - 1u);
        for (uint32_t off = crank * kChunk; off < bytes; off += CL * kChunk) {
            const uint32_t n = bytes - off < kChunk ? bytes - off : kChunk;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst + off),
                         "l"(src + off), "r"(n), "r"(bar), "h"(mask) : "memory");
        }
    }
}
__device__ __forceinline__ void stage_wait(uint64_t* s_bar)
{
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(s_bar);
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(0u) : "memory");
    }
}

// EV 0: product evaluator.  EV 1: constant shifts of the bit-slicing moved to the FMA pipe (IMAD.HI by a constant the
// compiler cannot see).  EV 2: flush by two sentinel FLUSH reads + min (table with 0xFFFF for lanes without a flush).
template <int EV>
__device__ __forceinline__ uint32_t ev7(uint64_t m, const SharedTab& tab, const LabConst& lc)
{
    if (EV == 0) return eval7(m, tab, lc.hp);
    if (EV == 4) {
        // hand-scheduled slicing: 8 logic ops, key merge and table addresses on the FMA pipe
        const HashParams& hp = lc.hp;
        const uint32_t lo = (uint32_t)m, hi = (uint32_t)(m >> 32);
        const uint32_t x = lo ^ hi, a = lo & hi;
        const uint32_t x1 = x >> 16, a1 = a >> 16;
        uint32_t b0, cy, b1, b2;
        asm("lop3.b32 %0, %1, %2, 0xFFFF, 0x28;" : "=r"(b0) : "r"(x), "r"(x1));          // (x ^ x1) & 0xFFFF
        cy = x & x1;
        asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(b1) : "r"(a), "r"(a1), "r"(cy));      // a ^ a1 ^ cy
        asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(b2) : "r"(a), "r"(a1), "r"(cy));      // majority
        const uint32_t key = b1 * 65536u + b0;
        const uint32_t h1 = key * hp.a1 + b2 * hp.b1;
        const uint32_t h2 = key * hp.a2 + b2 * hp.b2;
        uint32_t ad;
        ad = umulhi32(h1, HASH_BUCKETS) * lc.two17 + (tab.base + 2u * OFF_DISP);
        uint16_t dv;
        asm("ld.shared.u16 %0, [%1];" : "=h"(dv) : "r"(ad));
        const uint32_t inf = (h2 + (uint32_t)dv * HASH_DISP_MUL) >> 16;
        const int plo = popc32(lo);
        const uint32_t w = plo >= 5 ? lo : hi;
        const uint32_t wl = w & 0xFFFFu;
        const int pl = popc32(wl), pw = popc32(w);
        const uint32_t lane = pl >= 5 ? wl : (w >> 16);
        const bool fl = (pl >= 5) | (pw - pl >= 5);
        const uint32_t af = lane * lc.two17 + (tab.base + 2u * OFF_FLUSH), an = inf * lc.two17 + tab.base;
        const uint32_t ar = fl ? af : an;
        uint16_t rv;
        asm("ld.shared.u16 %0, [%1];" : "=h"(rv) : "r"(ar));
        return rv;
    }
    if (EV == 3) return (uint32_t)__popcll(m * 0x9E3779B97F4A7C15ull) + 1u;   // memory-streaming ceiling of this loop shape
    const HashParams& hp = lc.hp;
    const uint32_t lo = (uint32_t)m, hi = (uint32_t)(m >> 32);
    const uint32_t x = lo ^ hi, a = lo & hi;
    const uint32_t x1 = EV == 1 ? __umulhi(x, lc.two16) : x >> 16, a1 = EV == 1 ? __umulhi(a, lc.two16) : a >> 16;
    const uint32_t b0 = x ^ x1;
    const uint32_t cy = x & x1;
    const uint32_t b1 = a ^ a1 ^ cy;
    const uint32_t b2 = (a & a1) | (a & cy) | (a1 & cy);
    const uint32_t key = (b0 & 0xFFFFu) | (b1 << 16);
    const uint32_t h1 = key * hp.a1 + b2 * hp.b1;
    const uint32_t h2 = key * hp.a2 + b2 * hp.b2;
    const uint32_t d = tab.ld(OFF_DISP + umulhi32(h1, HASH_BUCKETS));
    const uint32_t inf = (h2 + d * HASH_DISP_MUL) >> 16;
    const int plo = popc32(lo);
    const uint32_t w = plo >= 5 ? lo : hi;
    if (EV == 2) {
        const uint32_t f0 = tab.ld(OFF_FLUSH + (w & 0xFFFFu)), f1 = tab.ld(OFF_FLUSH + (w >> 16));
        const uint32_t nf = tab.ld(inf);
        return min(min(f0, f1), nf);
    }
    const uint32_t wl = w & 0xFFFFu;
    const int pl = popc32(wl), pw = popc32(w);
    const uint32_t lane = pl >= 5 ? wl : (EV == 1 ? __umulhi(w, lc.two16) : (w >> 16));
    const bool fl = (pl >= 5) | (pw - pl >= 5);
    return tab.ld(fl ? OFF_FLUSH + lane : inf);
}

// HPT hands per thread per trip (4 or 8); PF: 0 = load, evaluate, store; 1 = next trip's loads are issued before this
// trip is evaluated (and the first trip's before the table image has landed)
template <int HPT, int PF, int EV, int CL = 1>
__global__ void __launch_bounds__(kThreads, 1)
k_lab(const uint16_t* __restrict__ g_tab, const uint64_t* __restrict__ h, int64_t n, uint16_t* __restrict__ out)
{
    uint64_t t0 = 0, t1 = 0;
    if (g_stamps && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + TAB_BYTES);
    constexpr int NV = HPT / 2;                      // 128-bit loads per trip
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t ntrip = n / HPT;                   // lab: n is a multiple of HPT
    const ulonglong2* hv = reinterpret_cast<const ulonglong2*>(h);
    ulonglong2 cur[NV], nxt[NV];
    int64_t q = gtid;
    if (PF == 1) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (PF == 1 && q < ntrip) {
#pragma unroll
        for (int j = 0; j < NV; j++) cur[j] = __ldcs(hv + NV * q + j);
    }
    if (CL == 1) stage(smem, g_tab, TAB_BYTES, bar); else stage_mc<CL>(smem, g_tab, TAB_BYTES, bar);
    stage_wait(bar);
    if (!PF) asm volatile("griddepcontrol.wait;" ::: "memory");
    const SharedTab tab{(uint32_t)__cvta_generic_to_shared(smem)};
    const LabConst lc = c_lab;
    if (PF == 2) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        if (q < ntrip) {
#pragma unroll
            for (int j = 0; j < NV; j++) cur[j] = __ldcs(hv + NV * q + j);
        }
    }
    if (g_stamps && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    for (; q < ntrip; q += nthr) {
        if (!PF) {
#pragma unroll
            for (int j = 0; j < NV; j++) cur[j] = __ldcs(hv + NV * q + j);
        } else {
            const int64_t qn = q + nthr < ntrip ? q + nthr : q;      // clamp: the last prefetch re-reads a line already in flight
#pragma unroll
            for (int j = 0; j < NV; j++) nxt[j] = __ldcs(hv + NV * qn + j);
        }
        uint32_t r[HPT];
#pragma unroll
        for (int j = 0; j < NV; j++) {
            r[2 * j] = ev7<EV>(cur[j].x, tab, lc);
            r[2 * j + 1] = ev7<EV>(cur[j].y, tab, lc);
        }
        if (HPT == 4) {
            uint2 o; o.x = r[0] | (r[1] << 16); o.y = r[2] | (r[3] << 16);
            __stcs(reinterpret_cast<uint2*>(out) + q, o);
        } else {
            uint4 o; o.x = r[0] | (r[1] << 16); o.y = r[2] | (r[3] << 16); o.z = r[4] | (r[5] << 16); o.w = r[6] | (r[7] << 16);
            __stcs(reinterpret_cast<uint4*>(out) + q, o);
        }
        if (PF) {
#pragma unroll
            for (int j = 0; j < NV; j++) cur[j] = nxt[j];
        }
    }
    if (CL > 1) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (g_stamps) {
        __syncthreads();
        if (threadIdx.x == 0) {
            uint64_t t2;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t2));
            g_stamps[3 * blockIdx.x] = t0; g_stamps[3 * blockIdx.x + 1] = t1; g_stamps[3 * blockIdx.x + 2] = t2;
        }
    }
}

__global__ void k_gen(uint64_t* h, int64_t n, uint64_t seed)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t s = seed + (uint64_t)i * 0x9E3779B97F4A7C15ull, m = 0;
    int k = 0;
    while (k < 7) {
        s ^= s >> 33; s *= 0xFF51AFD7ED558CCDull; s ^= s >> 29; s *= 0xC4CEB9FE1A85EC53ull; s ^= s >> 32;
        const uint32_t c = (uint32_t)(s >> 40) % 52u;
        const uint64_t b = id_bit(c);
        if (!(m & b)) { m |= b; k++; }
    }
    h[i] = m;
}

struct Variant { const char* name; void (*fn)(const uint16_t*, const uint64_t*, int64_t, uint16_t*); int ev; int cl; };

int main(int argc, char** argv)
{
    const int64_t n = argc > 1 ? atoll(argv[1]) : (1ll << 25);
    const int reps = argc > 2 ? atoi(argv[2]) : 15;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    std::vector<uint16_t> image(TAB_U16), image2;
    LabConst lc;
    if (!build_table_image(image.data(), &lc.hp)) { printf("table build failed\n"); return 1; }
    lc.two16 = 65536u; lc.two17 = 2u;
    image2 = image;
    for (uint32_t i = 0; i < FLUSH_SIZE; i++) if (image2[OFF_FLUSH + i] == 0) image2[OFF_FLUSH + i] = 0xFFFFu;
    CK(cudaMemcpyToSymbol(c_lab, &lc, sizeof lc));
    uint16_t *d_tab, *d_tab2, *d_out, *d_ref;
    uint64_t* d_h;
    unsigned char* d_flush;
    const size_t flush_bytes = 256u << 20;
    CK(cudaMalloc(&d_tab, TAB_BYTES)); CK(cudaMalloc(&d_tab2, TAB_BYTES));
    CK(cudaMemcpy(d_tab, image.data(), TAB_BYTES, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_tab2, image2.data(), TAB_BYTES, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_h, n * 8)); CK(cudaMalloc(&d_out, n * 2)); CK(cudaMalloc(&d_ref, n * 2)); CK(cudaMalloc(&d_flush, flush_bytes));
    k_gen<<<(unsigned)((n + 255) / 256), 256>>>(d_h, n, 20261020ull);
    CK(cudaDeviceSynchronize());

#define V(H, P, E) Variant{"hpt" #H "_pf" #P "_ev" #E, k_lab<H, P, E>, E, 1}
#define VC(H, P, E, C) Variant{"hpt" #H "_pf" #P "_ev" #E "_cl" #C, k_lab<H, P, E, C>, E, C}
    std::vector<Variant> vs = {V(4, 0, 0), V(4, 1, 0), V(4, 2, 0), V(4, 1, 4), V(4, 2, 4), V(4, 1, 3)};
    for (auto& v : vs) CK(cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    int pdl = 0;
    auto launch = [&](const Variant& v, const uint16_t* tab, uint16_t* dst) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(sms); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kSmem; cfg.stream = 0;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = v.cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = pdl ? 2 : 1;
        CK(cudaLaunchKernelEx(&cfg, v.fn, tab, (const uint64_t*)d_h, n, dst));
    };
    for (auto& v : vs) if (v.cl > 1) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(sms); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kSmem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = v.cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int nc = 0;
        CK(cudaOccupancyMaxActiveClusters(&nc, v.fn, &cfg));
        printf("%s: max active clusters %d (x%d = %d CTAs of %d SMs)\n", v.name, nc, v.cl, nc * v.cl, sms);
    }
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    std::vector<uint16_t> href(n), hout(n);
    for (size_t vi = 0; vi < vs.size(); vi++) {
        auto& v = vs[vi];
        const uint16_t* tab = v.ev == 2 ? d_tab2 : d_tab;
        uint16_t* dst = vi == 0 ? d_ref : d_out;
        std::vector<float> ms;
        for (int r = 0; r < reps + 3; r++) {
            CK(cudaMemsetAsync(d_flush, r, flush_bytes));
            CK(cudaEventRecord(e0));
            launch(v, tab, dst);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float t; CK(cudaEventElapsedTime(&t, e0, e1));
            if (r >= 3) ms.push_back(t);
        }
        if (vi < 2) {   // per-CTA stamps of one extra launch
            uint64_t* d_st; CK(cudaMalloc(&d_st, sms * 24));
            CK(cudaMemcpyToSymbol(g_stamps, &d_st, sizeof d_st));
            CK(cudaMemsetAsync(d_flush, 7, flush_bytes));
            launch(v, tab, dst);
            CK(cudaDeviceSynchronize());
            std::vector<uint64_t> st(sms * 3);
            CK(cudaMemcpy(st.data(), d_st, sms * 24, cudaMemcpyDeviceToHost));
            uint64_t* nul = nullptr; CK(cudaMemcpyToSymbol(g_stamps, &nul, sizeof nul));
            uint64_t first = ~0ull, last_start = 0, last_stage = 0, first_end = ~0ull, last_end = 0; double stage_sum = 0;
            for (int b = 0; b < sms; b++) {
                first = std::min(first, st[3 * b]); last_start = std::max(last_start, st[3 * b]);
                last_stage = std::max(last_stage, st[3 * b + 1]); first_end = std::min(first_end, st[3 * b + 2]); last_end = std::max(last_end, st[3 * b + 2]);
                stage_sum += (double)(st[3 * b + 1] - st[3 * b]);
            }
            printf("   stamps (ns from first CTA start): last start %llu, mean stage %.0f, last stage done %llu, first end %llu, last end %llu\n",
                   (unsigned long long)(last_start - first), stage_sum / sms, (unsigned long long)(last_stage - first),
                   (unsigned long long)(first_end - first), (unsigned long long)(last_end - first));
            cudaFree(d_st);
        }
        // back-to-back stream of 20 launches (input > L2), with and without programmatic dependent launch
        for (pdl = 0; pdl < 2; pdl++) {
            float best = 1e9f;
            for (int r = 0; r < 5; r++) {
                CK(cudaEventRecord(e0));
                for (int k = 0; k < 20; k++) launch(v, tab, dst);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float t; CK(cudaEventElapsedTime(&t, e0, e1));
                best = std::min(best, t / 20);
            }
            printf("   stream of 20%s: %.2f us/launch  %.0f GB/s\n", pdl ? " (PDL)" : "", best * 1e3, n * 10.0 / (best * 1e-3) / 1e9);
        }
        pdl = 0;
        std::sort(ms.begin(), ms.end());
        const float med = ms[ms.size() / 2];
        long bad = 0;
        if (vi == 0) CK(cudaMemcpy(href.data(), d_ref, n * 2, cudaMemcpyDeviceToHost));
        else {
            CK(cudaMemcpy(hout.data(), d_out, n * 2, cudaMemcpyDeviceToHost));
            for (int64_t i = 0; i < n; i++) bad += hout[i] != href[i];
            if (v.ev == 3) bad = -1;
        }
        printf("%-14s median %.2f us  min %.2f us  %.3e evals/s  %.0f GB/s  mismatches %ld\n", v.name, med * 1e3, ms[0] * 1e3,
               n / (med * 1e-3), n * 10.0 / (med * 1e-3) / 1e9, bad);
    }
    return 0;
}
