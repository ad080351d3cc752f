// tc_probe.cu -- stand-alone probe of tcgen05 kind::tf32 shared-memory descriptor conventions (dev tool).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o /tmp/tc_probe tools/tc_probe.cu && /tmp/tc_probe
// Runs D[128 x N] = A[128 x 48] . B for several (major, LBO, SBO, K-step) hypotheses and prints the max error
// against a CPU reference, for the forward (B = W^T, K-major) and the backward (B = W, MN-major) GEMM.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout = 0) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46) |
           ((uint64_t)layout << 61);
}

// wmode 0: no swizzle W4[k/4][n];  wmode 1: SWIZZLE_32B blocks [k/8][n/8] of 256 B = 8 n-rows x 8 k (32 B rows)
struct Hyp { int b_major; uint32_t lbo, sbo, kstep; int n; int wmode; uint32_t layout; };

__global__ void probe(const float* A, const float* W, float* D, Hyp h) {
    extern __shared__ __align__(1024) float sm[];
    float4* A4 = reinterpret_cast<float4*>(sm);             // [12][128]
    float* Wt = sm + 12 * 128 * 4;   /* 24576 B offset: 1024-aligned */                          // W4[k/4][n]: ((k>>2)*48 + n)*4 + (k&3)
    uint64_t* bar = reinterpret_cast<uint64_t*>(Wt + 48 * 48);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int m = threadIdx.x;
    for (int c = 0; c < 12; ++c) A4[c * 128 + m] = make_float4(A[m * 48 + 4 * c], A[m * 48 + 4 * c + 1], A[m * 48 + 4 * c + 2], A[m * 48 + 4 * c + 3]);
    for (int i = m; i < 48 * 48; i += 128) {
        int n = i / 48, k = i % 48;
        if (h.wmode == 0) Wt[((k >> 2) * 48 + n) * 4 + (k & 3)] = W[i];
        else {
            uint32_t off = ((k >> 3) * 6 + (n >> 3)) * 256 + (n & 7) * 32 + ((k >> 2) & 1) * 16 + (k & 3) * 4;
            off ^= ((off >> 7) & 1) << 4;   // Swizzle<1,4,3>
            Wt[off >> 2] = W[i];
        }
    }
    if (m == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (m < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(slot)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (m == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)h.b_major << 16) | ((uint32_t)(h.n >> 3) << 17) | ((128u >> 4) << 24);
        for (int ks = 0; ks < 6; ++ks) {
            uint64_t ad = smem_desc(smem_addr(A4) + ks * 4096, 2048, 128);
            uint64_t bd = smem_desc(smem_addr(Wt) + ks * h.kstep, h.lbo, h.sbo, h.layout);
            uint32_t acc = ks ? 1u : 0u;
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
    }
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_addr(bar)), "r"(0) : "memory");
    } while (!done);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem + ((uint32_t)(m & ~31) << 16);
    for (int j = 0; j < h.n / 16; ++j) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr + 16 * j) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) D[m * 48 + 16 * j + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (m < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64));
}

// A operand from TMEM (tcgen05.mma with [a_tmem]): A[m][k] is written by thread m into lane m, column a_col + k
__global__ void probe_ts(const float* A, const float* W, float* D, int n) {
    extern __shared__ __align__(1024) float sm[];
    float* Wt = sm;                                         // K-major no-swizzle W4[k/4][n]
    uint64_t* bar = reinterpret_cast<uint64_t*>(Wt + 48 * 48);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int m = threadIdx.x;
    for (int i = m; i < 48 * 48; i += 128) { int nn = i / 48, k = i % 48; Wt[((k >> 2) * 48 + nn) * 4 + (k & 3)] = W[i]; }
    if (m == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (m < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(slot)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    const uint32_t lane_base = tmem + ((uint32_t)(m & ~31) << 16);
    const uint32_t a_col = 64;
    for (int j = 0; j < 3; ++j) {
        uint32_t r[16];
        for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(A[m * 48 + 16 * j + i]);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                     :: "r"(lane_base + a_col + 16 * j), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (m == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
        for (int ks = 0; ks < 6; ++ks) {
            uint64_t bd = smem_desc(smem_addr(Wt) + ks * 1536, 768, 128);
            uint32_t acc = ks ? 1u : 0u;
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem), "r"(tmem + a_col + 8 * ks), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
    }
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_addr(bar)), "r"(0) : "memory");
    } while (!done);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int j = 0; j < n / 16; ++j) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(lane_base + 16 * j) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) D[m * 48 + 16 * j + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (m < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

int main() {
    float *A = (float*)malloc(128 * 48 * 4), *W = (float*)malloc(48 * 48 * 4), *D = (float*)malloc(128 * 48 * 4);
    srand(1);
    for (int i = 0; i < 128 * 48; ++i) A[i] = (rand() % 2001 - 1000) / 1000.f;
    for (int i = 0; i < 48 * 48; ++i) W[i] = (rand() % 2001 - 1000) / 1000.f;
    float *dA, *dW, *dD;
    cudaMalloc(&dA, 128 * 48 * 4); cudaMalloc(&dW, 48 * 48 * 4); cudaMalloc(&dD, 128 * 48 * 4);
    cudaMemcpy(dA, A, 128 * 48 * 4, cudaMemcpyHostToDevice); cudaMemcpy(dW, W, 48 * 48 * 4, cudaMemcpyHostToDevice);
    const size_t smem = (12 * 128 * 4 + 48 * 48 + 8) * 4;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    Hyp hyps[] = {
        {0, 768, 128, 1536, 48, 0, 0},    // forward, K-major no swizzle (known good)
        {1, 128, 768, 128, 48, 0, 0},     // backward MN-major no swizzle (gives zeros)
        {0, 16, 256, 1536, 48, 1, 6},     // forward, K-major SWIZZLE_32B over the block layout
        {1, 1536, 256, 256, 48, 1, 6},    // backward, MN-major SWIZZLE_32B over the SAME block layout
        {1, 256, 1536, 256, 48, 1, 6},    // same with LBO/SBO swapped
        {1, 1536, 256, 256, 16, 1, 6},    // N = 16
        {1, 256, 1536, 256, 16, 1, 6},
    };
    for (auto& h : hyps) {
        cudaMemset(dD, 0, 128 * 48 * 4);
        probe<<<1, 128, smem>>>(dA, dW, dD, h);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(D, dD, 128 * 48 * 4, cudaMemcpyDeviceToHost);
        double err = 0, mag = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < h.n; ++n) {
                double ref = 0;
                for (int k = 0; k < 48; ++k) ref += (double)A[m * 48 + k] * (h.b_major ? W[k * 48 + n] : W[n * 48 + k]);
                err = fmax(err, fabs(ref - D[m * 48 + n]));
                mag = fmax(mag, fabs(ref));
            }
        printf("wmode=%d layout=%u b_major=%d lbo=%u sbo=%u kstep=%u N=%d : cuda=%s max_err=%.3e (max |ref| %.2f) D[0][0..3]=%.4f %.4f %.4f %.4f\n", h.wmode, h.layout, h.b_major, h.lbo, h.sbo,
               h.kstep, h.n, cudaGetErrorString(e), err, mag, D[0], D[1], D[2], D[3]);
    }
    {
        cudaMemset(dD, 0, 128 * 48 * 4);
        probe_ts<<<1, 128, (48 * 48 + 8) * 4>>>(dA, dW, dD, 48);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(D, dD, 128 * 48 * 4, cudaMemcpyDeviceToHost);
        double err = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 48; ++n) {
                double ref = 0;
                for (int k = 0; k < 48; ++k) ref += (double)A[m * 48 + k] * W[n * 48 + k];
                err = fmax(err, fabs(ref - D[m * 48 + n]));
            }
        printf("A-from-TMEM forward: cuda=%s max_err=%.3e D[0][0..3]=%.4f %.4f %.4f %.4f\n", cudaGetErrorString(e), err, D[0], D[1], D[2], D[3]);
    }
    return 0;
}
