// tcw_exp.cuh -- tiled exponential-window map kernel (the FP32-FMA-bound path) and its
// weight-table builder.
//
// Replaces pyCUDAkernels/cudaTransientFstatExpWindow.cu (one thread per cell, every thread
// re-reading its whole atom range from global memory and calling exp() per atom visit).
//
// When t0 - t0_data advances in whole atoms (dt0 == TAtom; host certificate in tcw_b200.cu),
// the window weight of atom i for cell (m,n) depends only on k = i - i_t0(m) and n:
//     w(k,n) = e^{-(k*TAtom + delta)/tau_n}   for 0 <= k*TAtom + delta <= 3 tau_n, k <= K_n
// (delta = offset of the first summed atom from t0; K_n = i_t1 - i_t0, Exp.cu:27-65), and not
// on the template.  It is tabulated once per window range -- in `lal` mode with bit-exact
// emulation of lalpulsar's XLALFastNegExp lookup table, evaluated in FP64 like the original --
// and the map becomes a tiled direct sum
//     S_c[m,n] = sum_k X_c[i_t0(m)+k] * w(k,n)^(p_c),   p_c = 2 for a2,b2,ab; 1 for Fa,Fb
// (Exp.cu:92-100).  No exp() in the inner loop, no recurrence.
//
// CTA tile 64 (t0) x 64 (tau) cells, 256 threads, 4x4 cells x 7 channels = 112 FP32
// accumulators per thread.  Per k step a thread issues 112 FFMA + 4 FMUL (w^2) for
// 7 LDS.32 + 1 LDS.128: rows of a thread are consecutive t0, so their atoms form a sliding
// window held in registers (one new atom per channel per step), weights are shared along m.
// Operand tiles (64 k x 64 n weights, contiguous by construction of the table; 7 x 132 atoms)
// are staged by 1-D TMA bulk copies through a 3-stage mbarrier ring.  Ragged edges: weights
// are zero beyond each column's K_n, atoms are zero-padded beyond the data end, and a tile
// stops at min(K of its last column, atoms left after its first row).
#pragma once
#include "tcw_common.cuh"
#include "tcw_generic.cuh"
#include "tcw_prep.cuh"

#define TCW_EXP_THREADS 256
#define TCW_EXP_TM 64
#define TCW_EXP_TN 64
#define TCW_EXP_RM 4
#define TCW_EXP_RN 4
#define TCW_EXP_KC 64
#define TCW_EXP_XS (TCW_EXP_TM + TCW_EXP_KC + 4)  // staged atoms per channel (132)
#define TCW_EXP_STAGES 3
#define TCW_EXP_W_BYTES (TCW_EXP_KC * TCW_EXP_TN * 4)
#define TCW_EXP_X_BYTES (TCW_NCH * TCW_EXP_XS * 4)
#define TCW_EXP_STAGE_BYTES (TCW_EXP_W_BYTES + TCW_EXP_X_BYTES)
#define TCW_EXP_SMEM (TCW_EXP_STAGES * TCW_EXP_STAGE_BYTES)

struct ExpTableGeom {
    uint32_t N_tau, n_tiles, KW;  // KW: table rows per column tile (multiple of KC)
    uint32_t tau, dtau, TAtom;
    int32_t delta;  // (t0_data + i00*TAtom) - t0, in (-TAtom/2, TAtom/2]
};

// W[nt][k][TN]: weight of relative atom k for column n = nt*TN + j.
__global__ void tcw_exp_table_kernel(float *__restrict__ W, const int32_t *__restrict__ Kn,
                                     ExpTableGeom eg, const double *__restrict__ lut, int exact) {
    const size_t total = (size_t)eg.n_tiles * eg.KW * TCW_EXP_TN;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const uint32_t j = (uint32_t)(idx % TCW_EXP_TN);
        const size_t rest = idx / TCW_EXP_TN;
        const uint32_t k = (uint32_t)(rest % eg.KW);
        const uint32_t nt = (uint32_t)(rest / eg.KW);
        const uint32_t n = nt * TCW_EXP_TN + j;
        float wv = 0.0f;
        if (n < eg.N_tau && (int32_t)k <= Kn[n]) {
            const uint32_t tau_n = eg.tau + n * eg.dtau;
            const long long t_rel = (long long)k * eg.TAtom + eg.delta;  // t_i - t0_m
            if (t_rel >= 0 && t_rel <= (long long)TCW_EXP_EFOLDING * tau_n) {
                // REAL8 x = 1.0*(t_i - t0)/tau; XLALFastNegExp(x)
                const double x = __ddiv_rn((double)t_rel, (double)tau_n);
                wv = (float)(exact ? exp(-x) : fast_neg_exp_lut(x, lut));
            }
        }
        W[idx] = wv;
    }
}

__global__ void __launch_bounds__(TCW_EXP_THREADS, 1)
tcw_exp_map_kernel(const float *__restrict__ X, uint32_t xpad, const float *__restrict__ W,
                   const int32_t *__restrict__ Kn, uint32_t KW, const TplMeta *__restrict__ meta,
                   int t_base, MapWindow w, uint32_t i00, float *__restrict__ Fmn,
                   unsigned long long *__restrict__ maxkey, uint32_t *__restrict__ flags) {
    extern __shared__ __align__(128) unsigned char tcw_exp_smem[];
    __shared__ __align__(8) uint64_t full[TCW_EXP_STAGES];
    __shared__ unsigned long long red[TCW_EXP_THREADS / 32];

    const int tz = blockIdx.z;
    const int t = t_base + tz;
    const uint32_t numAtoms = meta[t].numAtoms;
    const uint32_t nt = blockIdx.x, mt = blockIdx.y;
    const uint32_t m0 = mt * TCW_EXP_TM, n0 = nt * TCW_EXP_TN;
    const uint32_t s_base = i00 + m0;  // i_t0 of the tile's first row (dt0 == TAtom)
    const uint32_t n_last = min(n0 + TCW_EXP_TN, w.N_tau) - 1;
    const int k_end = min(Kn[n_last] + 1, (int)numAtoms - (int)s_base);
    const int nchunks = k_end > 0 ? (k_end + TCW_EXP_KC - 1) / TCW_EXP_KC : 0;
    const uint32_t off = s_base & 3u;  // bulk copies need 16-byte aligned sources
    const float *Xt = X + (size_t)t * TCW_NCH * xpad + (s_base - off);
    const float *Wt = W + (size_t)nt * KW * TCW_EXP_TN;

    const int tid = threadIdx.x;
    const int tm = tid >> 4, tn = tid & 15;

    auto issue = [&](int chunk) {
        const int s = chunk % TCW_EXP_STAGES;
        unsigned char *st = tcw_exp_smem + (size_t)s * TCW_EXP_STAGE_BYTES;
        mbar_arrive_expect_tx(&full[s], TCW_EXP_STAGE_BYTES);
        bulk_g2s(st, Wt + (size_t)chunk * TCW_EXP_KC * TCW_EXP_TN, TCW_EXP_W_BYTES, &full[s]);
#pragma unroll
        for (int c = 0; c < TCW_NCH; c++)
            bulk_g2s(st + TCW_EXP_W_BYTES + c * TCW_EXP_XS * 4,
                     Xt + (size_t)c * xpad + (size_t)chunk * TCW_EXP_KC, TCW_EXP_XS * 4, &full[s]);
    };

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < TCW_EXP_STAGES; s++) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (int c = 0; c < TCW_EXP_STAGES - 1 && c < nchunks; c++) issue(c);
    }

    float acc[TCW_NCH][TCW_EXP_RM][TCW_EXP_RN];
#pragma unroll
    for (int c = 0; c < TCW_NCH; c++)
#pragma unroll
        for (int r = 0; r < TCW_EXP_RM; r++)
#pragma unroll
            for (int j = 0; j < TCW_EXP_RN; j++) acc[c][r][j] = 0.0f;

    for (int chunk = 0; chunk < nchunks; chunk++) {
        // refill the stage consumed in the previous iteration (all threads passed its sync)
        if (tid == 0 && chunk + TCW_EXP_STAGES - 1 < nchunks) issue(chunk + TCW_EXP_STAGES - 1);
        const int s = chunk % TCW_EXP_STAGES;
        mbar_wait(&full[s], (uint32_t)((chunk / TCW_EXP_STAGES) & 1));
        const float *Ws = reinterpret_cast<const float *>(tcw_exp_smem + (size_t)s * TCW_EXP_STAGE_BYTES);
        const float *Xs = Ws + TCW_EXP_KC * TCW_EXP_TN;
        const float *xrow = Xs + off + tm * TCW_EXP_RM;  // + c*XS + k + r
        const float4 *wrow = reinterpret_cast<const float4 *>(Ws) + tn;  // + k*(TN/4)

        // sliding window of 4 consecutive atoms per channel: value with relative index q
        // lives in slot q & 3
        float xr[TCW_NCH][4];
#pragma unroll
        for (int c = 0; c < TCW_NCH; c++) {
            xr[c][0] = xrow[c * TCW_EXP_XS + 0];
            xr[c][1] = xrow[c * TCW_EXP_XS + 1];
            xr[c][2] = xrow[c * TCW_EXP_XS + 2];
        }
#pragma unroll 1
        for (int kk = 0; kk < TCW_EXP_KC; kk += 4) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int k = kk + u;
#pragma unroll
                for (int c = 0; c < TCW_NCH; c++) xr[c][(u + 3) & 3] = xrow[c * TCW_EXP_XS + k + 3];
                const float4 wv = wrow[k * (TCW_EXP_TN / 4)];
                const float w1[4] = {wv.x, wv.y, wv.z, wv.w};
                float w2[4];
#pragma unroll
                for (int j = 0; j < 4; j++) w2[j] = w1[j] * w1[j];
#pragma unroll
                for (int c = 0; c < TCW_NCH; c++)
#pragma unroll
                    for (int r = 0; r < TCW_EXP_RM; r++) {
                        const float xv = xr[c][(u + r) & 3];
#pragma unroll
                        for (int j = 0; j < TCW_EXP_RN; j++)
                            acc[c][r][j] = fmaf(xv, c < 3 ? w2[j] : w1[j], acc[c][r][j]);
                    }
            }
        }
        __syncthreads();  // everyone is done with stage s before it is refilled
    }

    // ---- fused epilogue: F, optional store, max/argmax, degenerate flag ----
    const size_t cells = (size_t)w.N_t0 * w.N_tau;
    float *Ft = Fmn ? Fmn + (size_t)tz * cells : nullptr;
    float best = -1.0f;
    uint32_t best_flat = 0;
    bool degenerate = false;
#pragma unroll
    for (int r = 0; r < TCW_EXP_RM; r++) {
        const uint32_t m = m0 + tm * TCW_EXP_RM + r;
#pragma unroll
        for (int j = 0; j < TCW_EXP_RN; j++) {
            const uint32_t n = n0 + tn * TCW_EXP_RN + j;
            if (m < w.N_t0 && n < w.N_tau) {
                const float F = fstat_fast(acc[0][r][j], acc[1][r][j], acc[2][r][j], acc[3][r][j],
                                           acc[4][r][j], acc[5][r][j], acc[6][r][j]);
                const uint32_t flat = m * w.N_tau + n;
                if (Ft) Ft[flat] = F;
                if (F > best) {
                    best = F;
                    best_flat = flat;
                }
                const int K = Kn[n];
                const uint32_t s_m = i00 + m;
                if (K >= 0 && (K == 0 || s_m == numAtoms - 1)) degenerate = true;
            }
        }
    }
    if (degenerate) atomicOr(&flags[t], TCW_FLAG_DEGENERATE);
    const unsigned long long key = best > -1.0f ? pack_key(best, best_flat) : 0ull;
    block_atomic_max_key<TCW_EXP_THREADS / 32>(key, &maxkey[t], red);
}
