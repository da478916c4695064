// tcw_common.cuh -- shared device/host helpers of the B200 transient F-stat map backend.
//
// Index arithmetic follows the reference kernels bit for bit
// (pyCUDAkernels/cudaTransientFstatRectWindow.cu:21-31, 54-69; ...ExpWindow.cu:27-65):
// all uint32, signed re-interpretation only for the `< 0` clamp.  The runtime division by
// TAtom is done with a Granlund-Montgomery magic multiplier that is exact for every uint32
// dividend (tests/test_host_logic.py sweeps it against Python integers).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "tcw_b200.h"

#define TCW_NCH 7          // a2, b2, ab, Fa_re, Fa_im, Fb_re, Fb_im (tcw:711-721)

// XLALFastNegExp emulation: e^{-x} from a table of `len + 1` points on [0, xmax],
// nearest point `tab[(UINT4)(x * dxinv + 0.5)]` with dxinv = len / xmax, 0 beyond xmax.  The
// geometry is a RUNTIME property of the handle (tcw_set_exp_lut): lalsuite's constants cannot be
// read in the build container, so nothing about them is compiled in.
struct ExpLut {
    const double *tab;  // device pointer, len + 1 entries
    double dxinv;       // len / xmax
    double xmax;
    uint32_t len;
    float neg_dx_log2e_hi, neg_dx_log2e_lo;  // -(xmax/len) * log2(e) split in two floats (lnBtSG pass)
    uint32_t canonical;  // 1 if tab[i] == exp(-i * xmax/len) to 1e-13: values may be recomputed instead of fetched
};

// ---------------------------------------------------------------------------------------
// exact uint32 division by a runtime-invariant divisor
//   q = (t + ((x - t) >> sh1)) >> sh2,  t = mulhi(m, x)      (Granlund & Montgomery 1994, fig 4.1)
// ---------------------------------------------------------------------------------------
struct MagicDiv {
    uint32_t m, sh1, sh2;
};

static inline MagicDiv make_magic(uint32_t d) {
    MagicDiv md;
    uint32_t l = 0;
    while (l < 32 && (1ull << l) < (uint64_t)d) l++;  // l = ceil(log2 d)
    uint64_t num = ((1ull << l) - (uint64_t)d) << 32;  // 2^32 * (2^l - d) < 2^64
    md.m = (uint32_t)(num / d + 1);
    md.sh1 = l < 1 ? l : 1;
    md.sh2 = l > 1 ? l - 1 : 0;
    return md;
}

__host__ __device__ __forceinline__ uint32_t magic_div(uint32_t x, const MagicDiv md) {
#ifdef __CUDA_ARCH__
    uint32_t t = __umulhi(md.m, x);
#else
    uint32_t t = (uint32_t)(((uint64_t)md.m * (uint64_t)x) >> 32);
#endif
    return (t + ((x - t) >> md.sh1)) >> md.sh2;
}

// per-template geometry of the merged (binned) atoms + what the index math needs
struct TplMeta {
    uint32_t t0_data;   // first merged timestamp (tcw:725)
    uint32_t numAtoms;  // merged atoms on the TAtom grid (tcw:709)
};

struct IndexGeom {
    uint32_t TAtom, TAtomHalf;
    MagicDiv md;
    uint32_t ef;  // 1 (rect: t1 = t0+tau) or 3 (exp: t1 = t0 + 3 tau)
};

// clamp as the reference does: int i_tmp = q; if (i_tmp < 0) i_tmp = 0; min(., numAtoms-1)
__host__ __device__ __forceinline__ uint32_t clamp_index(uint32_t q, uint32_t numAtoms) {
    int32_t i = (int32_t)q;
    if (i < 0) i = 0;
    uint32_t u = (uint32_t)i;
    return u >= numAtoms ? numAtoms - 1 : u;
}
__host__ __device__ __forceinline__ uint32_t index_t0(uint32_t t0_m, uint32_t t0_data,
                                                      uint32_t numAtoms, const IndexGeom g) {
    return clamp_index(magic_div(t0_m - t0_data + g.TAtomHalf, g.md), numAtoms);
}
__host__ __device__ __forceinline__ uint32_t index_t1(uint32_t t1, uint32_t t0_data,
                                                      uint32_t numAtoms, const IndexGeom g) {
    return clamp_index(magic_div(t1 - t0_data + g.TAtomHalf, g.md) - 1u, numAtoms);
}

// window range with the TRANSIENT_NONE substitution already applied
struct MapWindow {
    uint32_t type;  // TCW_WINDOW_RECT or TCW_WINDOW_EXP
    uint32_t t0, dt0, tau, dtau;
    uint32_t t0Band, tauBand;
    uint32_t N_t0, N_tau;
    uint32_t pitch;  // row pitch (floats) of the device F_mn: N_tau rounded up to 4 -> 16-byte aligned rows
};

// ---------------------------------------------------------------------------------------
// max / argmax bookkeeping: one packed 64-bit key per template,
//   (orderable float bits << 32) | (0xFFFFFFFF - flat_index)
// so atomicMax picks the larger F and, among equal F, the smaller row-major index
// (first occurrence = np.argmax order, tcw:194, 810-813).  Key 0 = "no cell exceeded -1".
// ---------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t float_orderable(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float orderable_float(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    return __uint_as_float(u);
}
__device__ __forceinline__ unsigned long long pack_key(float F, uint32_t flat) {
    return ((unsigned long long)float_orderable(F) << 32) | (unsigned long long)(0xFFFFFFFFu - flat);
}

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o);
        v = other > v ? other : v;
    }
    return v;
}

// block-wide max of a packed key, then one atomicMax per CTA (hierarchical reduction)
template <int NWARPS>
__device__ __forceinline__ void block_atomic_max_key(unsigned long long key,
                                                     unsigned long long *dst,
                                                     unsigned long long *smem_scratch) {
    key = warp_max_u64(key);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) smem_scratch[warp] = key;
    __syncthreads();
    if (warp == 0) {
        unsigned long long v = lane < NWARPS ? smem_scratch[lane] : 0ull;
        v = warp_max_u64(v);
        if (lane == 0 && v != 0ull) atomicMax(dst, v);
    }
}

// block-wide max of a packed key; the result is valid in warp 0
template <int NWARPS>
__device__ __forceinline__ unsigned long long block_max_key(unsigned long long key,
                                                            unsigned long long *smem_scratch) {
    key = warp_max_u64(key);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) smem_scratch[warp] = key;
    __syncthreads();
    unsigned long long v = 0ull;
    if (warp == 0) {
        v = lane < NWARPS ? smem_scratch[lane] : 0ull;
        v = warp_max_u64(v);
    }
    return v;
}

// ---------------------------------------------------------------------------------------
// TMA bulk copy (1-D cp.async.bulk -> SASS UBLKCP) + mbarrier helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)
                 : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_plain(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy, completion signalled on an mbarrier (bytes % 16 == 0,
// both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                         uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---------------------------------------------------------------------------------------
// F-statistic from the 7 window sums.
//
// fstat_faithful: rounding for rounding what the reference kernels compute
// (Rect.cu:98-118 == Exp.cu:109-129; double literals promote sub-expressions), with explicit
// _rn intrinsics so nvcc cannot contract anything into FMAs.  Used by the generic kernels,
// which are bit-identical to the CPU oracle.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float fstat_faithful(float Ad, float Bd, float Cd, float Fa_re,
                                                float Fa_im, float Fb_re, float Fb_im) {
    const float sumAB = __fadd_rn(Ad, Bd);
    const float diffAB = __fsub_rn(Ad, Bd);
    const double d2 = __dadd_rn((double)__fmul_rn(diffAB, diffAB),
                                __dmul_rn(__dmul_rn(4.0, (double)Cd), (double)Cd));
    const float disc = __double2float_rn(__dsqrt_rn(d2));
    const float denom = __fsub_rn(sumAB, disc);
    const float cond = (denom > 0.0f) ? __fdiv_rn(__fadd_rn(sumAB, disc), denom) : INFINITY;
    float DdInv = 0.0f;
    if (cond < 1e4f) {
        const float det = __fsub_rn(__fmul_rn(Ad, Bd), __fmul_rn(Cd, Cd));
        DdInv = __double2float_rn(__ddiv_rn(1.0, (double)det));
    }
    float F = 2.0f;
    if (DdInv > 0.0f) {
        const float fa2 = __fadd_rn(__fmul_rn(Fa_re, Fa_re), __fmul_rn(Fa_im, Fa_im));
        const float fb2 = __fadd_rn(__fmul_rn(Fb_re, Fb_re), __fmul_rn(Fb_im, Fb_im));
        const float re = __fadd_rn(__fmul_rn(Fa_re, Fb_re), __fmul_rn(Fa_im, Fb_im));
        const float s1 = __fadd_rn(__fmul_rn(Bd, fa2), __fmul_rn(Ad, fb2));
        const double u = __dsub_rn((double)s1, __dmul_rn(__dmul_rn(2.0, (double)Cd), (double)re));
        F = __double2float_rn(__dmul_rn((double)DdInv, u));
    }
    return F;
}

// fstat_fast: the same guarded formula in pure FP32 for the tiled kernels, without sqrt or
// division in the conditioning test.  With s = A+B, d = sqrt((A-B)^2 + 4C^2), det = AB - C^2:
//   s^2 - d^2 = 4 det,   cond = (s+d)/(s-d) < 1e4  <=>  d < k s,  k = (1e4-1)/(1e4+1)   (s > 0)
//                                                   <=>  det > kappa s^2,  kappa = (1-k^2)/4
// which also implies the reference's DdInv > 0 guard.  DdInv comes from one MUFU.RCP (1 ulp).
// Near the cut both this test and the reference's (s - d suffers the same cancellation) are
// fuzzy at the few-1e-4 level in cond; cells flipping across it are listed by the parity
// tests.  ~20 FP32 instructions per cell.
__device__ __forceinline__ float fstat_fast(float Ad, float Bd, float Cd, float Fa_re, float Fa_im,
                                            float Fb_re, float Fb_im) {
    constexpr float kK = 9999.0f / 10001.0f;
    constexpr float kKappa = (1.0f - kK * kK) * 0.25f;
    const float sumAB = Ad + Bd;
    const float det = fmaf(Ad, Bd, -(Cd * Cd));
    // det - kappa s^2 > 0.  (The reference also needs s > 0, which holds for any physical atoms:
    // a2, b2 >= 0; with all-zero sums margin == 0 and the fallback is taken as in the reference.)
    const float margin = fmaf(sumAB * (-kKappa), sumAB, det);
    const bool ok = margin > 0.0f;
    const float fa2 = fmaf(Fa_re, Fa_re, Fa_im * Fa_im);
    const float fb2 = fmaf(Fb_re, Fb_re, Fb_im * Fb_im);
    const float re = fmaf(Fa_re, Fb_re, Fa_im * Fb_im);
    const float num = fmaf(Bd, fa2, fmaf(Cd * re, -2.0f, Ad * fb2));
    float rdet;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rdet) : "f"(det));
    return ok ? num * rdet : 2.0f;
}

// ---------------------------------------------------------------------------------------
// Blackwell packed FP32 (add/mul/fma.rn.f32x2 -> SASS FADD2/FMUL2/FFMA2: two results per issue
// slot).  Pairs are kept in 64-bit registers through inline PTX: with the float2 intrinsics
// nvcc re-assembles loop-invariant pairs from scalar registers before every use (2 MOVs per
// packed instruction in the rect kernel's hot loop).
// ---------------------------------------------------------------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// 3-input max / NaN-propagating min (sm_100: one FMNMX3 each)
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float fmin3_nan(float a, float b, float c) {
    float r;
    asm("min.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// fstat_fast for TWO cells at once (the rectangular-window kernel is instruction-issue-bound),
// UNGUARDED: returns F = num / det and the conditioning margin det - kappa s^2 of both cells; the
// caller selects F = 2 where the margin is not positive.  The ab sums come in as Cp = -2 C:
//   det = A B - Cp^2 / 4,   num = B |Fa|^2 + A |Fb|^2 + Cp Re(Fa conj Fb)
// 15 packed instructions + 2 MUFU.RCP + 2 FMUL per pair.
struct FstatConst2 {
    f32x2 neg_kappa, neg_quarter;
};
__device__ __forceinline__ FstatConst2 fstat_const2() {
    constexpr float kK = 9999.0f / 10001.0f;
    constexpr float kKappa = (1.0f - kK * kK) * 0.25f;
    FstatConst2 k;
    k.neg_kappa = pack2(-kKappa, -kKappa);
    k.neg_quarter = pack2(-0.25f, -0.25f);
    return k;
}
// NOGUARD: the caller holds a tile-wide conditioning certificate; the margin is not computed (3 packed
// instructions less per cell pair) and M0 = M1 = 1.
template <bool NOGUARD = false>
__device__ __forceinline__ void fstat_core2(const FstatConst2 &k, f32x2 Ad, f32x2 Bd, f32x2 Cp, f32x2 Fa_re,
                                            f32x2 Fa_im, f32x2 Fb_re, f32x2 Fb_im, float &F0, float &F1,
                                            float &M0, float &M1) {
    const f32x2 det = fma2(mul2(Cp, k.neg_quarter), Cp, mul2(Ad, Bd));  // AB - C^2
    f32x2 margin = det;
    if (!NOGUARD) {
        const f32x2 sumAB = add2(Ad, Bd);
        margin = fma2(mul2(sumAB, k.neg_kappa), sumAB, det);  // det - kappa s^2
    }
    const f32x2 fa2 = fma2(Fa_re, Fa_re, mul2(Fa_im, Fa_im));
    const f32x2 fb2 = fma2(Fb_re, Fb_re, mul2(Fb_im, Fb_im));
    const f32x2 re = fma2(Fa_re, Fb_re, mul2(Fa_im, Fb_im));
    const f32x2 num = fma2(Bd, fa2, fma2(Cp, re, mul2(Ad, fb2)));
    float d0, d1, n0, n1, r0, r1;
    unpack2(det, d0, d1);
    unpack2(num, n0, n1);
    if (NOGUARD) M0 = M1 = 1.0f;
    else unpack2(margin, M0, M1);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(d1));
    F0 = n0 * r0;
    F1 = n1 * r1;
}
#endif  // __CUDACC__
