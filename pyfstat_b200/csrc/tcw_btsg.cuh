// tcw_btsg.cuh -- lnBtSG marginalisation pass and the per-template finalize kernel.
//
// Replaces the host-side numpy reductions of the reference's GPU path (a full D2H of F_mn
// followed by max/argmax, tcw:810-815, and exp-sums, tcw:210-216, 247-251, 282-286) and the
// lalpulsar calls of its CPU path (XLALComputeTransientBstat, XLALComputeTransientPosterior_t0
// / _tau, XLALFindModeOfPDF1D; tcw:577-586).
//
//   lnBtSG = ln(70/(N_t0 N_tau)) + maxF + ln sum_mn e^{-(maxF - F_mn)}
//   t0_MP  = t0  + (argmax_m sum_n e^{..} + 1/2) t0Band / N_t0      (bin centre)
//   tau_MP = tau + (argmax_n sum_m e^{..} + 1/2) tauBand / N_tau
//
// In `lal` mode every term is XLALFastNegExp(maxF - F_mn): a nearest-point table lookup whose
// argument needs the FINAL maxF, so this is a second pass after the map kernel's atomicMax has
// settled (it re-reads the float32 F_mn the map kernel left in HBM/L2 -- 4 B per cell).  All
// sums are FP64, as in lalpulsar.
#pragma once
#include "tcw_common.cuh"
#include "tcw_generic.cuh"

#define TCW_BTSG_THREADS 256
#define TCW_BTSG_WARPS (TCW_BTSG_THREADS / 32)
#define TCW_BTSG_RPW 8  // rows per warp
#define TCW_BTSG_ROWS (TCW_BTSG_WARPS * TCW_BTSG_RPW)
#define TCW_BTSG_CPL 8  // columns per lane
#define TCW_BTSG_COLS (32 * TCW_BTSG_CPL)
#define TCW_BTSG_SMEM ((TCW_LUT_LEN + 2) * 8 + TCW_BTSG_WARPS * TCW_BTSG_COLS * 8 + TCW_BTSG_WARPS * TCW_BTSG_RPW * 32 * 8)

// e^{-(maxF - F)}: XLALFastNegExp emulation from a shared-memory copy of the table, or exact
template <bool EXACT_EXP>
__device__ __forceinline__ double btsg_term(double maxF, float F, const double *__restrict__ slut) {
    const double dF = maxF - (double)F;  // >= 0
    if (EXACT_EXP) return exp(-dF);
    // LUT[(UINT4)(dF*100 + 0.5)], 0 beyond 20.  The product and the sum round separately, as in
    // lalpulsar; the truncation to UINT4 is done on the FP64 pipe (add 2^52 rounding toward zero
    // leaves the integer part in the low mantissa word) instead of an XU-pipe F2I conversion.
    const double v = __dadd_rn(__dmul_rn(dF, (double)TCW_LUT_LEN / TCW_LUT_XMAX), 0.5);  // >= 0.5
    const uint32_t i0 = (uint32_t)__double2loint(__dadd_rz(v, 4503599627370496.0));
    return dF > TCW_LUT_XMAX ? 0.0 : slut[min(i0, (uint32_t)TCW_LUT_LEN)];
}

// CTA tile: 64 rows x 256 columns; a warp owns 8 rows, a lane 2 x 4 consecutive columns read
// with 128-bit loads (the device F_mn has a row pitch that is a multiple of 4 floats).  Column
// partials stay in registers over the warp's rows, row partials are combined through shared
// memory; one FP64 atomicAdd per row / column per CTA.
// Measured alternatives (60 d rect, T=64, pass time): scalar loads 0.70 ms; 128-bit loads
// 0.66 ms; + first rows' loads issued before the table-copy barrier and the UINT4 truncation
// on the FP64 pipe 0.625 ms (this version); cp.async.4 staging 1.29 ms; persistent CTAs with
// all 16 loads of a thread issued up front (126 registers, 2 CTAs/SM) 0.95 ms; warp-private
// rings of 1-KB TMA bulk copies (band of 256 columns x 256 rows per CTA) 0.96 ms -- ~100 cycles
// of TMA service per copy, serialised per SM, 2.2 M copies; persistent warp-autonomous units
// of 32 rows x 256 columns with the column partials flushed by FP64 atomics 0.82 ms (1.7e7
// atomics).  Many short-lived CTAs at high occupancy beat every "smarter" streaming scheme
// tried here.
//
// LOCATE: the rect map kernel published max VALUES only (key index part 0); complete the key
// with the smallest flat index whose F equals the max (first occurrence, np.argmax order).
template <bool EXACT_EXP, bool LOCATE>
__global__ void __launch_bounds__(TCW_BTSG_THREADS, 4)
tcw_btsg_kernel(const float *__restrict__ Fmn, int t_base, uint32_t N_t0, uint32_t N_tau, uint32_t pitch,
                unsigned long long *__restrict__ maxkey, const double *__restrict__ lut,
                double *__restrict__ rowsum, double *__restrict__ colsum) {
    extern __shared__ __align__(16) unsigned char tcw_btsg_smem[];
    double *slut = reinterpret_cast<double *>(tcw_btsg_smem);               // [LUT_LEN + 2]
    double *scol = slut + (TCW_LUT_LEN + 2);                                // [WARPS][COLS]
    double *srow = scol + TCW_BTSG_WARPS * TCW_BTSG_COLS;                   // [WARPS][RPW][32]
    const int tz = blockIdx.z;
    const int t = t_base + tz;
    const unsigned long long key = maxkey[t];
    const float maxFf = key ? orderable_float((uint32_t)(key >> 32)) : -1.0f;
    const double maxF = (double)maxFf;
    const float *Ft = Fmn + (size_t)tz * N_t0 * pitch;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t m0 = blockIdx.y * TCW_BTSG_ROWS + warp * TCW_BTSG_RPW, n0 = blockIdx.x * TCW_BTSG_COLS;
    // this lane's columns: n0 + 128*jv + 4*lane + q, jv = 0,1, q = 0..3  (accumulator j = 4*jv + q)
    double colacc[TCW_BTSG_CPL];
#pragma unroll
    for (int j = 0; j < TCW_BTSG_CPL; j++) colacc[j] = 0.0;
    const bool full = (m0 + TCW_BTSG_RPW <= N_t0) && (n0 + TCW_BTSG_COLS <= N_tau);
    const float4 *p = reinterpret_cast<const float4 *>(Ft + (size_t)m0 * pitch + n0) + lane;
    // the first rows' loads are issued BEFORE the table copy and its barrier, so that the two
    // round trips to L2/HBM overlap (the CTA is short-lived: 64 cells per thread)
    constexpr int kPre = 4;
    float4 pre[kPre][2];
    if (full) {
#pragma unroll
        for (int i = 0; i < kPre; i++)
#pragma unroll
            for (int jv = 0; jv < 2; jv++) pre[i][jv] = __ldg(p + (size_t)i * (pitch / 4) + 32 * jv);
    }
    if (!EXACT_EXP) {
        for (int i = threadIdx.x; i <= TCW_LUT_LEN; i += TCW_BTSG_THREADS) slut[i] = __ldg(lut + i);
        __syncthreads();
    }
    if (full) {
#pragma unroll
        for (int i = 0; i < TCW_BTSG_RPW; i++) {
            float f[TCW_BTSG_CPL];
#pragma unroll
            for (int jv = 0; jv < 2; jv++) {
                const float4 v = i < kPre ? pre[i < kPre ? i : 0][jv] : __ldg(p + (size_t)i * (pitch / 4) + 32 * jv);
                f[4 * jv + 0] = v.x; f[4 * jv + 1] = v.y; f[4 * jv + 2] = v.z; f[4 * jv + 3] = v.w;
            }
            double ra = 0.0;
#pragma unroll
            for (int j = 0; j < TCW_BTSG_CPL; j++) {
                const double e = btsg_term<EXACT_EXP>(maxF, f[j], slut);
                ra += e;
                colacc[j] += e;
                if (LOCATE && f[j] == maxFf)
                    atomicMax(&maxkey[t], pack_key(f[j], (m0 + i) * N_tau + n0 + 128 * (j >> 2) + 4 * lane + (j & 3)));
            }
            srow[(warp * TCW_BTSG_RPW + i) * 32 + lane] = ra;
        }
    } else {
#pragma unroll 1
        for (int i = 0; i < TCW_BTSG_RPW; i++) {
            double ra = 0.0;
            const uint32_t m = m0 + i;
#pragma unroll
            for (int j = 0; j < TCW_BTSG_CPL; j++) {
                const uint32_t n = n0 + 128 * (j >> 2) + 4 * lane + (j & 3);
                if (m < N_t0 && n < N_tau) {
                    const float fv = __ldg(Ft + (size_t)m * pitch + n);
                    const double e = btsg_term<EXACT_EXP>(maxF, fv, slut);
                    ra += e;
                    colacc[j] += e;
                    if (LOCATE && fv == maxFf) atomicMax(&maxkey[t], pack_key(fv, m * N_tau + n));
                }
            }
            srow[(warp * TCW_BTSG_RPW + i) * 32 + lane] = ra;
        }
    }
    // column partials: scol[warp][local column], local column = 128*jv + 4*lane + q
#pragma unroll
    for (int j = 0; j < TCW_BTSG_CPL; j++)
        scol[warp * TCW_BTSG_COLS + 128 * (j >> 2) + 4 * lane + (j & 3)] = colacc[j];
    __syncthreads();
    {  // row partials: 4 lanes per row sum 8 entries each, 2 shuffles, 1 atomic
        const int r = lane >> 2, q = lane & 3;
        const double *src = srow + (warp * TCW_BTSG_RPW + r) * 32 + q * 8;
        double sacc = 0.0;
#pragma unroll
        for (int k = 0; k < 8; k++) sacc += src[k];
        sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
        sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
        const uint32_t m = m0 + r;
        if (q == 0 && m < N_t0) atomicAdd(&rowsum[(size_t)t * N_t0 + m], sacc);
    }
    {
        const uint32_t n = n0 + threadIdx.x;
        if (n < N_tau) {
            double sacc = 0.0;
#pragma unroll
            for (int wv = 0; wv < TCW_BTSG_WARPS; wv++) sacc += scol[wv * TCW_BTSG_COLS + threadIdx.x];
            atomicAdd(&colsum[(size_t)t * N_tau + n], sacc);
        }
    }
}

// (value desc, index asc) argmax + sum, one CTA per template
struct ArgMaxD {
    double v;
    uint32_t i;
};
__device__ __forceinline__ ArgMaxD argmax_better(ArgMaxD a, ArgMaxD b) {
    return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}

#define TCW_FIN_THREADS 256

__device__ __forceinline__ void block_sum_argmax(const double *__restrict__ v, uint32_t n,
                                                 double *sum_out, uint32_t *arg_out) {
    __shared__ double ssum[TCW_FIN_THREADS / 32];
    __shared__ ArgMaxD sarg[TCW_FIN_THREADS / 32];
    double s = 0.0;
    ArgMaxD a;
    a.v = -INFINITY;
    a.i = 0xFFFFFFFFu;
    for (uint32_t i = threadIdx.x; i < n; i += TCW_FIN_THREADS) {
        const double x = v[i];
        s += x;
        ArgMaxD b;
        b.v = x;
        b.i = i;
        a = argmax_better(a, b);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        ArgMaxD b;
        b.v = __shfl_xor_sync(0xffffffffu, a.v, o);
        b.i = __shfl_xor_sync(0xffffffffu, a.i, o);
        a = argmax_better(a, b);
    }
    __syncthreads();
    if (lane == 0) {
        ssum[warp] = s;
        sarg[warp] = a;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        ArgMaxD best = sarg[0];
        for (int wv = 0; wv < TCW_FIN_THREADS / 32; wv++) {
            tot += ssum[wv];
            best = argmax_better(best, sarg[wv]);
        }
        *sum_out = tot;
        *arg_out = best.i == 0xFFFFFFFFu ? 0u : best.i;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(TCW_FIN_THREADS)
tcw_finalize_kernel(const unsigned long long *__restrict__ maxkey, const uint32_t *__restrict__ flags,
                    const double *__restrict__ rowsum, const double *__restrict__ colsum,
                    const TplMeta *__restrict__ meta, MapWindow w, const MapWindow *__restrict__ wins, int none_window,
                    uint32_t TAtom,
                    int want_btsg, int allow_degenerate, uint32_t path, tcw_result *__restrict__ results) {
    const int t = blockIdx.x;
    if (wins) w = wins[t];
    __shared__ double s_tot, s_tot2;
    __shared__ uint32_t s_mMP, s_nMP;
    if (want_btsg) {
        block_sum_argmax(rowsum + (size_t)t * w.N_t0, w.N_t0, &s_tot, &s_mMP);
        block_sum_argmax(colsum + (size_t)t * w.N_tau, w.N_tau, &s_tot2, &s_nMP);
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    tcw_result r;
    r.N_t0 = w.N_t0;
    r.N_tau = w.N_tau;
    r.numAtoms = meta[t].numAtoms;
    r.t0_data = meta[t].t0_data;
    r.path = path;
    r.reserved = 0;
    // TRANSIENT_NONE: rect window spanning this template's data (tcw:742-749)
    const uint32_t w_t0 = none_window ? meta[t].t0_data : w.t0;
    const uint32_t w_tau = none_window ? meta[t].numAtoms * TAtom : w.tau;
    const unsigned long long key = maxkey[t];
    if (key) {
        r.maxF = orderable_float((uint32_t)(key >> 32));
        const uint32_t flat = 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull);
        r.m_ML = flat / w.N_tau;
        r.n_ML = flat - r.m_ML * w.N_tau;
        r.t0_ML = w_t0 + r.m_ML * w.dt0;
        r.tau_ML = w_tau + r.n_ML * w.dtau;
    } else {  // no cell exceeded the initial -1: lalpulsar leaves the calloc'ed zeros
        r.maxF = -1.0f;
        r.m_ML = r.n_ML = 0;
        r.t0_ML = r.tau_ML = 0;
    }
    r.m_MP = r.n_MP = 0;
    r.lnBtSG = r.t0_MP = r.tau_MP = __longlong_as_double(0x7ff8000000000000LL);  // NaN (tcw:142-144)
    if (want_btsg) {
        const double normBh = 70.0 / ((double)w.N_t0 * (double)w.N_tau);
        const double logBhat = (double)r.maxF + log(s_tot);
        r.lnBtSG = log(normBh) + logBhat;
        r.m_MP = s_mMP;
        r.n_MP = s_nMP;
        r.t0_MP = (double)w_t0 + ((double)s_mMP + 0.5) * ((double)w.t0Band / (double)w.N_t0);
        r.tau_MP = (double)w_tau + ((double)s_nMP + 0.5) * ((double)w.tauBand / (double)w.N_tau);
    }
    const uint32_t f = flags[t];
    r.status = TCW_OK;
    if ((f & TCW_FLAG_DEGENERATE) && !allow_degenerate) r.status = TCW_E_DEGENERATE;
    if (f & TCW_FLAG_UNSORTED) r.status = TCW_E_INVALID;
    results[t] = r;
}
