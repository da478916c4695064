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
// dynamic shared memory of the table-fetching kernel for a table of `len + 1` entries
#define TCW_BTSG_TABLE_SMEM(len) \
    (((size_t)(len) + 2) * 8 + TCW_BTSG_WARPS * TCW_BTSG_COLS * 8 + TCW_BTSG_WARPS * TCW_BTSG_RPW * 32 * 8)

// Row / column marginals are accumulated across CTAs as 64-bit FIXED-POINT integers (partial sum x
// 2^k, k chosen by the host so that a whole map cannot overflow: 2^39 -> 1.8e-12 per partial at
// 60 d): integer atomics commute, so the marginals, lnBtSG and the MP indices are bit-reproducible
// from run to run and independent of how the templates are sharded over GPUs -- FP64 atomicAdd in
// arrival order is not.
__device__ __forceinline__ unsigned long long btsg_fixed(double partial, double fx_scale) {
    return (unsigned long long)__double2ll_rn(partial * fx_scale);
}

// ---------------------------------------------------------------------------------------
// (1) Table-fetching pass: every term is read from a shared-memory copy of the table and summed
// in FP64, exactly the values lalpulsar adds.  Used when the caller asks for it
// (TCW_BTSG_TABLE) or when a non-canonical table was installed with tcw_set_exp_lut.
// ---------------------------------------------------------------------------------------
// e^{-(maxF - F)}: XLALFastNegExp emulation from a shared-memory copy of the table, or exact
template <bool EXACT_EXP>
__device__ __forceinline__ double btsg_term(double maxF, float F, const double *__restrict__ slut,
                                            const ExpLut &lut) {
    const double dF = maxF - (double)F;  // >= 0
    if (EXACT_EXP) return exp(-dF);
    // tab[(UINT4)(dF*dxinv + 0.5)], 0 beyond xmax.  The product and the sum round separately, as in
    // lalpulsar; the truncation to UINT4 is done on the FP64 pipe (add 2^52 rounding toward zero
    // leaves the integer part in the low mantissa word) instead of an XU-pipe F2I conversion.
    const double v = __dadd_rn(__dmul_rn(dF, lut.dxinv), 0.5);  // >= 0.5
    const uint32_t i0 = (uint32_t)__double2loint(__dadd_rz(v, 4503599627370496.0));
    return dF > lut.xmax ? 0.0 : slut[min(i0, lut.len)];
}

// CTA tile: 64 rows x 256 columns; a warp owns 8 rows, a lane 2 x 4 consecutive columns read
// with 128-bit loads (the device F_mn has a row pitch that is a multiple of 4 floats).  Column
// partials stay in registers over the warp's rows, row partials are combined through shared
// memory; one FP64 atomicAdd per row / column per CTA.
//
// LOCATE: the rect map kernel published max VALUES only (key index part 0); complete the key
// with the smallest flat index whose F equals the max (first occurrence, np.argmax order).
template <bool EXACT_EXP, bool LOCATE>
__global__ void __launch_bounds__(TCW_BTSG_THREADS, 2)
tcw_btsg_table_kernel(const float *__restrict__ Fmn, int t_base, uint32_t N_t0, uint32_t N_tau, uint32_t pitch,
                      uint32_t n_ct, unsigned long long *__restrict__ maxkey, const ExpLut lut,
                      double fx_scale, unsigned long long *__restrict__ rowsum,
                      unsigned long long *__restrict__ colsum) {
    extern __shared__ __align__(16) unsigned char tcw_btsg_smem[];
    double *slut = reinterpret_cast<double *>(tcw_btsg_smem);               // [len + 2]
    double *scol = slut + (EXACT_EXP ? 0 : lut.len + 2);                    // [WARPS][COLS]
    double *srow = scol + TCW_BTSG_WARPS * TCW_BTSG_COLS;                   // [WARPS][RPW][32]
    const int tz = blockIdx.z;
    const int t = t_base + tz;
    const unsigned long long key = maxkey[t];
    const float maxFf = key ? orderable_float((uint32_t)(key >> 32)) : -1.0f;
    const double maxF = (double)maxFf;
    const float *Ft = Fmn + (size_t)tz * N_t0 * pitch;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // tiles linearised in grid.x, column tile fastest
    const uint32_t tile_y = blockIdx.x / n_ct, tile_x = blockIdx.x - tile_y * n_ct;
    const uint32_t m0 = tile_y * TCW_BTSG_ROWS + warp * TCW_BTSG_RPW, n0 = tile_x * TCW_BTSG_COLS;
    // this lane's columns: n0 + 128*jv + 4*lane + q, jv = 0,1, q = 0..3  (accumulator j = 4*jv + q)
    double colacc[TCW_BTSG_CPL];
#pragma unroll
    for (int j = 0; j < TCW_BTSG_CPL; j++) colacc[j] = 0.0;
    if (!EXACT_EXP) {
        for (uint32_t i = threadIdx.x; i <= lut.len; i += TCW_BTSG_THREADS) slut[i] = __ldg(lut.tab + i);
        __syncthreads();
    }
#pragma unroll 1
    for (int i = 0; i < TCW_BTSG_RPW; i++) {
        double ra = 0.0;
        const uint32_t m = m0 + i;
#pragma unroll
        for (int j = 0; j < TCW_BTSG_CPL; j++) {
            const uint32_t n = n0 + 128 * (j >> 2) + 4 * lane + (j & 3);
            if (m < N_t0 && n < N_tau) {
                const float fv = __ldg(Ft + (size_t)m * pitch + n);
                const double e = btsg_term<EXACT_EXP>(maxF, fv, slut, lut);
                ra += e;
                colacc[j] += e;
                if (LOCATE && fv == maxFf) atomicMax(&maxkey[t], pack_key(fv, m * N_tau + n));
            }
        }
        srow[(warp * TCW_BTSG_RPW + i) * 32 + lane] = ra;
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
        if (q == 0 && m < N_t0) atomicAdd(&rowsum[(size_t)t * N_t0 + m], btsg_fixed(sacc, fx_scale));
    }
    {
        const uint32_t n = n0 + threadIdx.x;
        if (n < N_tau) {
            double sacc = 0.0;
#pragma unroll
            for (int wv = 0; wv < TCW_BTSG_WARPS; wv++) sacc += scol[wv * TCW_BTSG_COLS + threadIdx.x];
            atomicAdd(&colsum[(size_t)t * N_tau + n], btsg_fixed(sacc, fx_scale));
        }
    }
}

// ---------------------------------------------------------------------------------------
// (2) Streaming pass (default): no table in shared memory.  Round 1's pass kept a copy of the
// table per CTA (16 KB re-read from L2 by every 64-KB tile, one barrier, and 47 % of its
// shared-memory wavefronts were bank conflicts of the random FP64 lookups -- ncu r01_btsg_v4).
// For a canonical table, tab[i] = e^{-i dx}, the looked-up VALUE is recomputed instead:
//   * the table INDEX i0 = (UINT4)((maxF - F) dxinv + 0.5) stays in FP64 (exact difference of the
//     two floats; one DFMA against c0 = maxF dxinv + 0.5, then the 2^52 truncation trick), so the
//     quantisation -- the part of XLALFastNegExp that matters at the 1e-3 level -- is lalpulsar's;
//   * the value e^{-i0 dx} = 2^{i0 k}, k = -dx log2(e) carried as two floats, comes from one
//     FFMA pair and one MUFU.EX2: relative error <= ~3e-7 per term (table entries are exact to
//     1e-16), i.e. <= 3e-7 absolute on lnBtSG -- the parity bar is 1e-4;
//   * "0 beyond xmax" is one FP32 compare against the smallest float >= maxF - xmax (exact,
//     because F is a float);
//   * partial sums of 8 terms in FP32, row partials combined with a 9-shuffle butterfly, column
//     partials through 8 KB of shared memory, then FP64 atomics per row / column per CTA.
// EXACT_EXP (the reference's numpy semantics, tcw:210-216): e^{-(maxF-F)} with the difference
// rounded to FP32 (relative error <= 1.5e-6 on terms that matter).
// ---------------------------------------------------------------------------------------
struct BtsgConst {
    double c0;       // maxF * dxinv + 0.5 (two roundings, as lalpulsar forms it for dF = maxF - F ... see above)
    double dxinv;
    float thr;       // smallest float >= maxF - xmax: F < thr  <=>  maxF - F > xmax
    float maxFf;
    float k_hi, k_lo;
};

template <bool EXACT_EXP>
__device__ __forceinline__ float btsg_term_fast(float F, const BtsgConst &k) {
    if (EXACT_EXP) {
        float e;
        const float y = (k.maxFf - F) * -1.4426950408889634f;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(y));
        return e;
    }
    // v = c0 - F dxinv (one rounding instead of lalpulsar's three: the results differ only when v is
    // within 1e-16 relative of an integer, ~1e-13 per cell), truncated through the 2^52 trick
    const double v = fma(-(double)F, k.dxinv, k.c0);
    const uint32_t i0 = (uint32_t)__double2loint(__dadd_rz(v, 4503599627370496.0));
    const float i0f = __uint_as_float(0x4B000000u | i0) - 8388608.0f;  // exact for i0 < 2^23
    const float y = fmaf(i0f, k.k_lo, i0f * k.k_hi);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(y));
    return F < k.thr ? 0.0f : e;
}

template <bool EXACT_EXP, bool LOCATE>
__global__ void __launch_bounds__(TCW_BTSG_THREADS, 4)
tcw_btsg_kernel(const float *__restrict__ Fmn, int t_base, uint32_t N_t0, uint32_t N_tau, uint32_t pitch,
                uint32_t n_ct, unsigned long long *__restrict__ maxkey, const ExpLut lut,
                double fx_scale, unsigned long long *__restrict__ rowsum,
                unsigned long long *__restrict__ colsum) {
    __shared__ float scol[TCW_BTSG_WARPS][TCW_BTSG_COLS];
    const int tz = blockIdx.z;
    const int t = t_base + tz;
    const unsigned long long key = maxkey[t];
    BtsgConst k;
    k.maxFf = key ? orderable_float((uint32_t)(key >> 32)) : -1.0f;
    const double maxF = (double)k.maxFf;
    k.dxinv = lut.dxinv;
    k.c0 = __dadd_rn(__dmul_rn(maxF, lut.dxinv), 0.5);
    {
        const double thr = maxF - lut.xmax;
        float tf = __double2float_ru(thr);
        k.thr = tf;
    }
    k.k_hi = lut.neg_dx_log2e_hi;
    k.k_lo = lut.neg_dx_log2e_lo;
    const float *Ft = Fmn + (size_t)tz * N_t0 * pitch;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // tiles linearised in grid.x, column tile fastest
    const uint32_t tile_y = blockIdx.x / n_ct, tile_x = blockIdx.x - tile_y * n_ct;
    const uint32_t m0 = tile_y * TCW_BTSG_ROWS + warp * TCW_BTSG_RPW, n0 = tile_x * TCW_BTSG_COLS;
    // this lane's columns: n0 + 128*jv + 4*lane + q, jv = 0,1, q = 0..3  (accumulator j = 4*jv + q)
    float colacc[TCW_BTSG_CPL], rowacc[TCW_BTSG_RPW];
#pragma unroll
    for (int j = 0; j < TCW_BTSG_CPL; j++) colacc[j] = 0.0f;
    const bool full = (m0 + TCW_BTSG_RPW <= N_t0) && (n0 + TCW_BTSG_COLS <= N_tau);
    const float4 *p = reinterpret_cast<const float4 *>(Ft + (size_t)m0 * pitch + n0) + lane;
    const uint32_t p4 = pitch / 4;
    if (full) {
#pragma unroll
        for (int half = 0; half < 2; half++) {
            float4 v[TCW_BTSG_RPW / 2][2];
#pragma unroll
            for (int i = 0; i < TCW_BTSG_RPW / 2; i++)
#pragma unroll
                for (int jv = 0; jv < 2; jv++)
                    v[i][jv] = __ldcs(p + (size_t)(half * (TCW_BTSG_RPW / 2) + i) * p4 + 32 * jv);
#pragma unroll
            for (int i = 0; i < TCW_BTSG_RPW / 2; i++) {
                const int row = half * (TCW_BTSG_RPW / 2) + i;
                float f[TCW_BTSG_CPL];
                f[0] = v[i][0].x; f[1] = v[i][0].y; f[2] = v[i][0].z; f[3] = v[i][0].w;
                f[4] = v[i][1].x; f[5] = v[i][1].y; f[6] = v[i][1].z; f[7] = v[i][1].w;
                float ra = 0.0f;
                bool hit = false;  // one (rarely taken) branch per row instead of one per cell
#pragma unroll
                for (int j = 0; j < TCW_BTSG_CPL; j++) {
                    const float e = btsg_term_fast<EXACT_EXP>(f[j], k);
                    ra += e;
                    colacc[j] += e;
                    if (LOCATE) hit |= f[j] == k.maxFf;
                }
                if (LOCATE && hit) {
#pragma unroll
                    for (int j = 0; j < TCW_BTSG_CPL; j++)
                        if (f[j] == k.maxFf)
                            atomicMax(&maxkey[t],
                                      pack_key(f[j], (m0 + row) * N_tau + n0 + 128 * (j >> 2) + 4 * lane + (j & 3)));
                }
                rowacc[row] = ra;
            }
        }
    } else {
        // edge tiles: rows beyond N_t0 / columns beyond N_tau masked; the 128-bit loads stay inside
        // the padded row (pitch is a multiple of 4 floats)
#pragma unroll 1
        for (int i = 0; i < TCW_BTSG_RPW; i++) {
            const uint32_t m = m0 + i;
            float ra = 0.0f;
#pragma unroll
            for (int jv = 0; jv < 2; jv++) {
                const uint32_t nb = n0 + 128 * jv + 4 * lane;
                if (m < N_t0 && nb < N_tau) {
                    const float4 v4 = __ldcs(p + (size_t)i * p4 + 32 * jv);
                    const float f[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        if (nb + q < N_tau) {
                            const float e = btsg_term_fast<EXACT_EXP>(f[q], k);
                            ra += e;
                            colacc[4 * jv + q] += e;
                            if (LOCATE && f[q] == k.maxFf) atomicMax(&maxkey[t], pack_key(f[q], m * N_tau + nb + q));
                        }
                    }
                }
            }
            // dynamic index into a register array would spill: select statically
#pragma unroll
            for (int r = 0; r < TCW_BTSG_RPW; r++)
                if (r == i) rowacc[r] = ra;
        }
    }
    // column partials -> shared memory
#pragma unroll
    for (int jv = 0; jv < 2; jv++)
        *reinterpret_cast<float4 *>(&scol[warp][128 * jv + 4 * lane]) =
            make_float4(colacc[4 * jv], colacc[4 * jv + 1], colacc[4 * jv + 2], colacc[4 * jv + 3]);
    // row partials: 8 values per lane reduced over the 32 lanes by a halving butterfly
    // (4 + 2 + 1 + 1 + 1 shuffles); lane L ends up with row ((L>>4)&1)*4 + ((L>>3)&1)*2 + ((L>>2)&1)
    {
        bool up = lane & 16;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const float send = up ? rowacc[q] : rowacc[q + 4];
            const float keep = up ? rowacc[q + 4] : rowacc[q];
            rowacc[q] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
        up = lane & 8;
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const float send = up ? rowacc[q] : rowacc[q + 2];
            const float keep = up ? rowacc[q + 2] : rowacc[q];
            rowacc[q] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
        up = lane & 4;
        {
            const float send = up ? rowacc[0] : rowacc[1];
            const float keep = up ? rowacc[1] : rowacc[0];
            rowacc[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
        rowacc[0] += __shfl_xor_sync(0xffffffffu, rowacc[0], 2);
        rowacc[0] += __shfl_xor_sync(0xffffffffu, rowacc[0], 1);
        const uint32_t r = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
        const uint32_t m = m0 + r;
        if ((lane & 3) == 0 && m < N_t0) atomicAdd(&rowsum[(size_t)t * N_t0 + m], btsg_fixed((double)rowacc[0], fx_scale));
    }
    __syncthreads();
    {
        const uint32_t n = n0 + threadIdx.x;
        if (n < N_tau) {
            float sacc = 0.0f;
#pragma unroll
            for (int wv = 0; wv < TCW_BTSG_WARPS; wv++) sacc += scol[wv][threadIdx.x];
            atomicAdd(&colsum[(size_t)t * N_tau + n], btsg_fixed((double)sacc, fx_scale));
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

__device__ __forceinline__ void block_sum_argmax(const unsigned long long *__restrict__ v, uint32_t n,
                                                 double fx_inv, double *sum_out, uint32_t *arg_out) {
    __shared__ unsigned long long ssum[TCW_FIN_THREADS / 32];
    __shared__ ArgMaxD sarg[TCW_FIN_THREADS / 32];
    unsigned long long s = 0ull;  // integer total: exact, order-independent
    ArgMaxD a;
    a.v = -INFINITY;
    a.i = 0xFFFFFFFFu;
    for (uint32_t i = threadIdx.x; i < n; i += TCW_FIN_THREADS) {
        const unsigned long long xi = v[i];
        s += xi;
        ArgMaxD b;
        b.v = (double)xi;
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
        unsigned long long tot = 0ull;
        ArgMaxD best = sarg[0];
        for (int wv = 0; wv < TCW_FIN_THREADS / 32; wv++) {
            tot += ssum[wv];
            best = argmax_better(best, sarg[wv]);
        }
        *sum_out = (double)tot * fx_inv;
        *arg_out = best.i == 0xFFFFFFFFu ? 0u : best.i;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(TCW_FIN_THREADS)
tcw_finalize_kernel(const unsigned long long *__restrict__ maxkey, const uint32_t *__restrict__ flags,
                    const unsigned long long *__restrict__ rowsum, const unsigned long long *__restrict__ colsum,
                    double fx_inv,
                    const TplMeta *__restrict__ meta, MapWindow w, const MapWindow *__restrict__ wins, int none_window,
                    uint32_t TAtom,
                    int want_btsg, int allow_degenerate, uint32_t path, tcw_result *__restrict__ results) {
    const int t = blockIdx.x;
    if (wins) w = wins[t];
    __shared__ double s_tot, s_tot2;
    __shared__ uint32_t s_mMP, s_nMP;
    if (want_btsg) {
        block_sum_argmax(rowsum + (size_t)t * w.N_t0, w.N_t0, fx_inv, &s_tot, &s_mMP);
        block_sum_argmax(colsum + (size_t)t * w.N_tau, w.N_tau, fx_inv, &s_tot2, &s_nMP);
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
