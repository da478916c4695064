// hankel_tf32.cu -- probe for the round-1 review's stretch item: can the
// exponential-window sum  S[m,n] = sum_k X[s + m + k] W[k,n]  run on the 5th-gen tensor cores?
//
// The A operand is a HANKEL matrix.  In the no-swizzle K-major canonical layout of a tcgen05 shared-
// memory descriptor -- ((8,m),2):((1,SBO),LBO) in 16-byte units -- a row is 16 bytes = 4 TF32 values and
// the 8 rows of a core matrix are 16 bytes apart, so with LBO = 16 B and SBO = 128 B the descriptor
// laid over the plain array X reads A[i,k] = X[4 i + k]: the map rows m = 4 i + r of row class r, with
// NO expansion of the Hankel matrix in shared memory (class r uses a copy of X shifted by r values).
// Precision: 3xTF32 (X = Xhi + Xlo, W = Whi + Wlo; Xhi Whi + Xhi Wlo + Xlo Whi), FP32 accumulation in TMEM.
//
// This program checks (1) that the descriptor trick computes the right thing, (2) the accuracy against
// an FP64 reference next to a sequential FP32 sum (what the FFMA2 kernel does), (3) the MMA issue
// rate for the shape the map kernel would use (M = 128, N = 64, K = 8).
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hankel_tf32 hankel_tf32.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define M_ROWS 128
#ifndef N_COLS
#define N_COLS 64
#endif
#define NB (N_COLS / 8)
#define STR2(x) #x
#define STR(x) STR2(x)
#define KC 32  // k per staged chunk (4 MMA K-steps of 8)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// no-swizzle K-major descriptor: start, LBO, SBO in bytes (multiples of 16)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version 1 (sm_100)
    return d;                // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE
}

// instruction descriptor, kind::tf32, D = F32, A and B K-major
__host__ __device__ constexpr uint32_t make_idesc(uint32_t M, uint32_t N) {
    return (1u << 4) /*D F32*/ | (2u << 7) /*A TF32*/ | (2u << 10) /*B TF32*/ | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\tbra WAIT_LOOP;\n\tWAIT_DONE:\n\t}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// split mode 0: truncation (what the hardware does to a raw FP32 operand); 1: round to nearest (cvt.rna.tf32.f32),
// the low part rounded as well, so that the hardware's own truncation of the operands changes nothing
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ float tf32_hi(float x, int rna) {
    return rna ? tf32_rna(x) : __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
}
__device__ __forceinline__ float tf32_lo(float x, float hi, int rna) { return rna ? tf32_rna(x - hi) : x - hi; }

// One CTA, 128 threads.  X: [Lx] floats (global), Wt: [K/KC] chunks, each tiled [kq 8][nb 8][nr 8][kk 4]
// (core matrices of 8 n x 4 k, 128 B each).  D: [128][64].  reps > 1: re-issue the same MMAs (timing).
__global__ void __launch_bounds__(128, 1)
hankel_kernel(const float *__restrict__ X, const float *__restrict__ Wt, int K, float *__restrict__ D, int reps,
              int split3, int rna, int rate_only) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int Lx = 4 * (M_ROWS - 1) + K + 16;
    float *sXhi = reinterpret_cast<float *>(smem);
    float *sXlo = sXhi + ((Lx + 31) & ~31);
    float *sWhi = sXlo + ((Lx + 31) & ~31);  // [KC*N_COLS] tiled
    float *sWlo = sWhi + KC * N_COLS;
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int i = tid; i < Lx; i += 128) {
        const float x = X[(size_t)blockIdx.x * 0 + i];
        const float hi = tf32_hi(x, rna);
        sXhi[i] = hi;
        sXlo[i] = tf32_lo(x, hi, rna);
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(N_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = make_idesc(M_ROWS, N_COLS);

    uint32_t phase = 0;
    for (int rep = 0; rep < reps; rep++) {
        for (int c = 0; c < K / KC; c++) {
            if (!rate_only || (c == 0 && rep == 0)) {
                // stage the W chunk (split into hi / lo)
                for (int i = tid; i < KC * N_COLS; i += 128) {
                    const float wv = Wt[(size_t)c * KC * N_COLS + i];
                    const float hi = tf32_hi(wv, rna);
                    sWhi[i] = hi;
                    sWlo[i] = tf32_lo(wv, hi, rna);
                }
                // generic-proxy writes -> visible to the tensor-core (async) proxy
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncthreads();
            }
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int j = 0; j < KC / 8; j++) {
                    // A: rows 16 B apart inside a core matrix (fixed), K halves 16 B apart, 8-row groups 128 B apart
                    const uint32_t a_off = (uint32_t)(c * KC + j * 8) * 4u;
                    const uint64_t a_hi = make_desc(smem_u32(sXhi) + a_off, 16, 128);
                    const uint64_t a_lo = make_desc(smem_u32(sXlo) + a_off, 16, 128);
                    // B: chunk tiled [kq][nb][8][4]: K halves NB * 128 B apart, 8-column groups 128 B apart
                    const uint32_t b_off = (uint32_t)(2 * j) * (NB * 128u);
                    const uint64_t b_hi = make_desc(smem_u32(sWhi) + b_off, NB * 128, 128);
                    const uint64_t b_lo = make_desc(smem_u32(sWlo) + b_off, NB * 128, 128);
                    const uint32_t first = (c == 0 && j == 0 && rep == 0) ? 0u : 1u;
                    mma_tf32(tmem, a_hi, b_hi, idesc, first);
                    if (split3) {
                        mma_tf32(tmem, a_hi, b_lo, idesc, 1u);
                        mma_tf32(tmem, a_lo, b_hi, idesc, 1u);
                    }
                }
                if (!rate_only || (c == K / KC - 1 && rep == reps - 1)) mma_commit(&bar);
            }
            if (!rate_only || (c == K / KC - 1 && rep == reps - 1)) {
                mbar_wait(&bar, phase);  // MMAs done: the W buffers may be overwritten
                phase ^= 1u;
            }
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: warp w reads TMEM lanes 32 w .. 32 w + 31 (= rows), 64 columns
    uint32_t v[N_COLS];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int q = 0; q < N_COLS / 16; q++) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(v[16 * q + 0]), "=r"(v[16 * q + 1]), "=r"(v[16 * q + 2]), "=r"(v[16 * q + 3]), "=r"(v[16 * q + 4]),
              "=r"(v[16 * q + 5]), "=r"(v[16 * q + 6]), "=r"(v[16 * q + 7]), "=r"(v[16 * q + 8]), "=r"(v[16 * q + 9]),
              "=r"(v[16 * q + 10]), "=r"(v[16 * q + 11]), "=r"(v[16 * q + 12]), "=r"(v[16 * q + 13]),
              "=r"(v[16 * q + 14]), "=r"(v[16 * q + 15])
            : "r"(taddr + 16 * q));
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (blockIdx.x == 0)
        for (int n = 0; n < N_COLS; n++) D[(size_t)tid * N_COLS + n] = __uint_as_float(v[n]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(N_COLS) : "memory");
}

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e = (x);                                                       \
        if (e != cudaSuccess) {                                                    \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            return 1;                                                              \
        }                                                                          \
    } while (0)

int main(int argc, char **argv) {
    const int K = argc > 1 ? atoi(argv[1]) : 2048;
    const int Lx = 4 * (M_ROWS - 1) + K + 16;
    std::vector<float> X(Lx), W((size_t)K * N_COLS), Wt((size_t)K * N_COLS);
    srand(12345);
    for (int i = 0; i < Lx; i++) X[i] = (float)((rand() / (double)RAND_MAX - 0.5) * 2.0 + (i % 7 == 0 ? 0.3 : 0.0));
    for (int k = 0; k < K; k++)
        for (int n = 0; n < N_COLS; n++) W[(size_t)k * N_COLS + n] = (float)exp(-(double)k / (200.0 + 30.0 * n));
    // tile W per chunk: [kq][nb][nr][kk]
    for (int c = 0; c < K / KC; c++)
        for (int k = 0; k < KC; k++)
            for (int n = 0; n < N_COLS; n++) {
                const int kq = k / 4, kk = k % 4, nb = n / 8, nr = n % 8;
                Wt[(size_t)c * KC * N_COLS + ((kq * NB + nb) * 8 + nr) * 4 + kk] = W[(size_t)(c * KC + k) * N_COLS + n];
            }
    // references
    std::vector<double> ref((size_t)M_ROWS * N_COLS), mag((size_t)M_ROWS * N_COLS);
    std::vector<float> seq((size_t)M_ROWS * N_COLS);
    for (int i = 0; i < M_ROWS; i++)
        for (int n = 0; n < N_COLS; n++) {
            double s = 0, a = 0;
            float f = 0;
            for (int k = 0; k < K; k++) {
                const double t = (double)X[4 * i + k] * (double)W[(size_t)k * N_COLS + n];
                s += t;
                a += fabs(t);
                f = fmaf(X[4 * i + k], W[(size_t)k * N_COLS + n], f);
            }
            ref[(size_t)i * N_COLS + n] = s;
            mag[(size_t)i * N_COLS + n] = a;
            seq[(size_t)i * N_COLS + n] = f;
        }
    float *dX, *dW, *dD;
    CK(cudaMalloc(&dX, Lx * 4));
    CK(cudaMalloc(&dW, Wt.size() * 4));
    CK(cudaMalloc(&dD, (size_t)M_ROWS * N_COLS * 4));
    CK(cudaMemcpy(dX, X.data(), Lx * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dW, Wt.data(), Wt.size() * 4, cudaMemcpyHostToDevice));
    const size_t smem = (size_t)(2 * ((Lx + 31) & ~31) + 2 * KC * N_COLS) * 4 + 256;
    CK(cudaFuncSetAttribute(hankel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    std::vector<float> Dh((size_t)M_ROWS * N_COLS);
    for (int mode = 0; mode < 3; mode++) {
        const int split3 = mode > 0, rna = mode > 1;
        CK(cudaMemset(dD, 0, Dh.size() * 4));
        hankel_kernel<<<1, 128, smem>>>(dX, dW, K, dD, 1, split3, rna, 0);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(Dh.data(), dD, Dh.size() * 4, cudaMemcpyDeviceToHost));
        double e_abs = 0, e_mag = 0, e_seq = 0, e_seq_abs = 0;
        int bad = 0;
        for (size_t i = 0; i < Dh.size(); i++) {
            const double d = fabs((double)Dh[i] - ref[i]);
            e_abs = fmax(e_abs, d / fmax(fabs(ref[i]), 1e-30));
            e_mag = fmax(e_mag, d / mag[i]);
            e_seq = fmax(e_seq, fabs((double)seq[i] - ref[i]) / mag[i]);
            e_seq_abs = fmax(e_seq_abs, fabs((double)seq[i] - ref[i]) / fmax(fabs(ref[i]), 1e-30));
            if (d / mag[i] > 1e-2) bad++;
        }
        printf("K=%d %s: max |err|/|sum| %.3e (sequential FP32 FMA %.3e), max |err|/sum|terms| %.3e (FP32 %.3e), gross mismatches %d\n",
               K, mode == 0 ? "1xTF32" : mode == 1 ? "3xTF32 truncated split" : "3xTF32 rounded split", e_abs, e_seq_abs, e_mag,
               e_seq, bad);
    }
    // issue-rate probe: all SMs
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int rate_only = 0; rate_only <= 1; rate_only++)
        for (int split3 = 0; split3 <= 1; split3++) {
            const int reps = rate_only ? 200 : 20;
            hankel_kernel<<<prop.multiProcessorCount, 128, smem>>>(dX, dW, K, dD, 2, split3, 1, rate_only);
            CK(cudaEventRecord(e0));
            hankel_kernel<<<prop.multiProcessorCount, 128, smem>>>(dX, dW, K, dD, reps, split3, 1, rate_only);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            const double macs = (double)prop.multiProcessorCount * reps * (double)K * M_ROWS * N_COLS * (split3 ? 3 : 1);
            printf("rate %s, %s: %.3f ms, %.1f TMAC/s issued = %.1f useful TMAC/s (FFMA2 kernel: 30.5 useful TMAC/s; TF32 dense peak 565 TMAC/s)\n",
                   split3 ? "3xTF32" : "1xTF32",
                   rate_only ? "MMAs back to back on resident operands (M128 N" STR(N_COLS) " K8), one commit at the end"
                             : "SIMT staging of W + a full MMA drain per 32-k chunk",
                   ms, macs / ms * 1e-9, macs / (split3 ? 3 : 1) / ms * 1e-9);
        }
    return 0;
}
