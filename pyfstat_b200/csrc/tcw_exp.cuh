// tcw_exp.cuh -- tiled exponential-window map kernel (the FP32-FMA-bound path) and its
// weight-table builder.
//
// Replaces pyCUDAkernels/cudaTransientFstatExpWindow.cu (one thread per cell, every thread
// re-reading its whole atom range from global memory and calling exp() per atom visit).
//
// Row classes.  The window weight of atom i for cell (m,n) is a function of t_i - t0_m.  Rows whose
// t0_m - t0_data are congruent modulo TAtom share it: with P = TAtom / gcd(dt0 mod TAtom, TAtom)
// (P = 1 when dt0 is a whole number of atoms) the rows m = r, r + P, r + 2P, ... of class r start
// A = P dt0 / TAtom atoms apart and see the same offsets (host certificate in tcw_b200.cu: P <= 4,
// A <= 4; templates may differ in t0_data by whole atoms; rows starting beyond the data end read
// zero padding).  Within a class the weight depends only on k = i - i_t0(m) and n:
//     w(k,n) = e^{-(k*TAtom + delta)/tau_n}   for 0 <= k*TAtom + delta <= 3 tau_n, k <= K_n
// (delta = offset of the first summed atom from t0; K_n = i_t1 - i_t0, Exp.cu:27-65), and not
// on the template.  It is tabulated once per window range -- in `lal` mode with bit-exact
// emulation of lalpulsar's XLALFastNegExp lookup table, evaluated in FP64 like the original --
// and the map becomes a tiled direct sum
//     S_c[m,n] = sum_k X_c[i_t0(m)+k] * w(k,n)^(p_c),   p_c = 2 for a2,b2,ab; 1 for Fa,Fb
// (Exp.cu:92-100).  No exp() in the inner loop, no recurrence.
//
// Default CTA tile 32 (t0) x 64 (tau) cells, 128 threads, 3 CTAs/SM, 4x4 cells x 7 channels =
// 112 FP32 accumulators per thread (ExpCfg below parametrises the tile; measured in round 1:
// 64x64/256 thr 14.70 ms, 32x64/128 thr 13.77 ms, 64x32/256 thr 13.80 ms per 32 30-d templates).
// Per k step a thread issues 112 FFMA + 4 FMUL (w^2) for
// 3 LDS.128: rows of a thread are consecutive t0, so their atoms form a sliding window held in
// registers (one new atom = one 32-byte record of all 7 channels per step), weights are shared
// along m.
// Operand tiles (64 k x 64 n weights and 100 atom records, both contiguous by construction)
// are staged by 1-D TMA bulk copies through a 3-stage mbarrier ring.  Ragged edges: weights
// are zero beyond each column's K_n, atoms are zero-padded beyond the data end, and a tile
// stops at min(K of its last column, atoms left after its first row).
// Measured and rejected: warps of 8 (t0) x 4 (tau) threads, each stopping at its own 16 columns' K (saves the ~7 % of
// FMAs spent on zero weights at 30 d): -3.5 % -- the atom records of a quarter-warp then sit 128 bytes apart (2-way
// LDS.128 conflicts) and finished warps idle at the per-chunk barrier.
#pragma once
#include "tcw_common.cuh"
#include "tcw_generic.cuh"
#include "tcw_prep.cuh"

#define TCW_EXP_KC 32      // k-steps per staged chunk (64: -2% at 30 d with FFMA2)
#ifndef TCW_EXP_STAGES
#define TCW_EXP_STAGES 3    // measured: 2, 3 and 4 stages give the same time (41.2 ms / 128 30-d maps): staging is fully hidden
#endif
// canonical kernel: stages of its ring and how many chunks ahead the refills run.  Measured (B200,
// 128 x 30-d / 4 x 120-d maps): barrier per chunk + w^2 by FMUL2 41.21 / 77.86 ms; decoupled ring alone
// 43.81 / 82.66; w^2 plane alone 40.96 / 76.11; both 39.20 / 73.96 ms (the default).
#ifndef TCW_EXP_NO_DECOUPLE
#define TCW_EXP_DECOUPLE 1
#endif
#ifndef TCW_EXP_NO_W2TAB
#define TCW_EXP_W2TAB 1
#endif
#ifdef TCW_EXP_DECOUPLE
#define TCW_EXP_CSTAGES 4
#define TCW_EXP_PREFETCH 2
#else
#define TCW_EXP_CSTAGES TCW_EXP_STAGES
#define TCW_EXP_PREFETCH (TCW_EXP_STAGES - 1)
#endif
// weight planes per table row: 1 = w only (w^2 by one FMUL2 per column pair and k step), 2 = [w | w^2]
#ifdef TCW_EXP_W2TAB
#define TCW_EXP_WP 2
#else
#define TCW_EXP_WP 1
#endif
#define TCW_EXP_TNT 16     // threads along tau per CTA (fixed); a warp = 2 (t0) x 16 (tau) threads
#define TCW_EXP_AMAX 4     // max atoms between consecutive rows of a row class
#define TCW_EXP_PMAX 4     // max row classes

// Tile configuration: NT threads, RM x RN cells per thread.
//   tile = (NT/16 * RM) rows x (16 * RN) columns
template <int NT, int RM, int RN>
struct ExpCfg {
    static constexpr int kThreads = NT;
    static constexpr int kRM = RM, kRN = RN;
    static constexpr int kTM = NT / TCW_EXP_TNT * RM;
    static constexpr int kTN = TCW_EXP_TNT * RN;
    // staged atoms (32-byte records: 7 channels + pad): rows of a class are A atoms apart
    static constexpr int kXS1 = kTM + TCW_EXP_KC + 4;                             // A == 1
    static constexpr int kXS = (kTM - 1) * TCW_EXP_AMAX + 1 + TCW_EXP_KC + 4 + 3;  // A <= AMAX (multiple of 4)
    static constexpr int kWBytes = TCW_EXP_KC * kTN * 4 * TCW_EXP_WP;
    static constexpr int kXBytes = kXS * 32;
    static constexpr int kXBytes1 = kXS1 * 32;
    static constexpr int kStageBytes = kWBytes + kXBytes;    // rows A atoms apart
    static constexpr int kStageBytes1 = kWBytes + kXBytes1;  // sliding window (A == 1)
    static constexpr int kSmem = TCW_EXP_STAGES * kStageBytes;
    static constexpr int kSmem1 = TCW_EXP_STAGES * kStageBytes1;
    static constexpr int kSmemC = TCW_EXP_CSTAGES * kStageBytes1;  // canonical kernel
    static_assert(kWBytes % 16 == 0, "bulk copies need 16-byte multiples");
    static_assert(RM == 4, "the register sliding window is written for 4 rows per thread");
    static_assert(RN == 4 || RN == 2, "weights are fetched as float4 / float2");
};

struct ExpTableGeom {
    uint32_t N_tau, n_tiles, KW;  // KW: table rows per column tile (multiple of KC)
    uint32_t TN;                  // columns per tile
    uint32_t tau, dtau, TAtom;
    uint32_t P;                       // row classes
    int32_t delta[TCW_EXP_PMAX];      // per class: (t0_data + i00*TAtom) - t0_m, in (-TAtom/2, TAtom/2]
};

// geometry of the row classes for the map kernel
struct ExpClasses {
    uint32_t P, A;                    // classes; atoms between consecutive rows of a class
    uint32_t ybeg[TCW_EXP_PMAX + 1];  // first row tile (blockIdx.y) of each class
    uint32_t i00[TCW_EXP_PMAX];       // i_t0 of the class's first row, for the reference template
};

// W[class][nt][k][TN]: weight of relative atom k for column n = nt*TN + j; Kn[class][N_tau].
__global__ void tcw_exp_table_kernel(float *__restrict__ W, const int32_t *__restrict__ Kn,
                                     ExpTableGeom eg, const ExpLut lut, int exact) {
    const size_t per_class = (size_t)eg.n_tiles * eg.KW * eg.TN;
    const size_t total = per_class * eg.P;
    for (size_t idx0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx0 < total;
         idx0 += (size_t)gridDim.x * blockDim.x) {
        const uint32_t cls = (uint32_t)(idx0 / per_class);
        const size_t idx = idx0 - (size_t)cls * per_class;
        const uint32_t j = (uint32_t)(idx % eg.TN);
        const size_t rest = idx / eg.TN;
        const uint32_t k = (uint32_t)(rest % eg.KW);
        const uint32_t nt = (uint32_t)(rest / eg.KW);
        const uint32_t n = nt * eg.TN + j;
        float wv = 0.0f;
        if (n < eg.N_tau && (int32_t)k <= Kn[(size_t)cls * eg.N_tau + n]) {
            const uint32_t tau_n = eg.tau + n * eg.dtau;
            const long long t_rel = (long long)k * eg.TAtom + eg.delta[cls];  // t_i - t0_m
            if (t_rel >= 0 && t_rel <= (long long)TCW_EXP_EFOLDING * tau_n) {
                // REAL8 x = 1.0*(t_i - t0)/tau; XLALFastNegExp(x)
                const double x = __ddiv_rn((double)t_rel, (double)tau_n);
                wv = (float)(exact ? exp(-x) : fast_neg_exp_lut(x, lut));
            }
        }
        // row layout of the table: [TN values of w][TN values of w^2] when TCW_EXP_WP == 2
        const size_t row = idx0 / eg.TN;
        W[row * (eg.TN * TCW_EXP_WP) + j] = wv;
        if (TCW_EXP_WP == 2) W[row * (eg.TN * TCW_EXP_WP) + eg.TN + j] = __fmul_rn(wv, wv);
    }
}

// The canonical case -- one row class, rows one atom apart, every template starting on the same
// atom, no start beyond the data end (dt0 == TAtom grids: every BASELINE config) -- keeps round 1's
// kernel VERBATIM: the generalised kernel below has the same main loop (224 FFMA2, 18 LDS.128, 8 FMUL2
// per 4 k-steps), but nvcc schedules it differently around the extra tile geometry (131 instead of 146
// operand-reuse hints) and it measures 2.4 % slower (42.22 vs 41.20 ms per 128 x 30-d maps,
// 79.7 vs 77.9 ms per 4 x 120-d maps) whatever is done to the prologue / epilogue.
template <class Cfg>
__global__ void __launch_bounds__(Cfg::kThreads, (Cfg::kThreads == 256 ? (Cfg::kRN == 4 ? 1 : 2) : 3))
tcw_exp_map_canon_kernel(const float *__restrict__ X8, uint32_t xpad, const float *__restrict__ W,
                   const int32_t *__restrict__ Kn, uint32_t KW, const TplMeta *__restrict__ meta,
                   int t_base, MapWindow w, uint32_t i00, float *__restrict__ Fmn,
                   unsigned long long *__restrict__ maxkey, uint32_t *__restrict__ flags) {
    constexpr int TM = Cfg::kTM, TN = Cfg::kTN, RM = Cfg::kRM, RN = Cfg::kRN, XS = Cfg::kXS1;
    constexpr int NT = Cfg::kThreads;
    extern __shared__ __align__(128) unsigned char tcw_exp_smem[];
    __shared__ __align__(8) uint64_t full[TCW_EXP_CSTAGES];
#ifdef TCW_EXP_DECOUPLE
    __shared__ __align__(8) uint64_t empty[TCW_EXP_CSTAGES];
#endif
    __shared__ unsigned long long red[NT / 32];

    const int tz = blockIdx.z;
    const int t = t_base + tz;
    const uint32_t numAtoms = meta[t].numAtoms;
    const uint32_t nt = blockIdx.x, mt = blockIdx.y;
    const uint32_t m0 = mt * TM, n0 = nt * TN;
    const uint32_t s_base = i00 + m0;  // i_t0 of the tile's first row (dt0 == TAtom)
    const uint32_t n_last = min(n0 + TN, w.N_tau) - 1;
    const int k_end = min(Kn[n_last] + 1, (int)numAtoms - (int)s_base);
    const int nchunks = k_end > 0 ? (k_end + TCW_EXP_KC - 1) / TCW_EXP_KC : 0;
    const float *Xt = X8 + ((size_t)t * xpad + s_base) * 8;  // 32-byte atom records: always 16-byte aligned
    const float *Wt = W + (size_t)nt * KW * TN * TCW_EXP_WP;

    const int tid = threadIdx.x;
    const int tm = tid >> 4, tn = tid & 15;

    auto issue = [&](int chunk) {
        const int s = chunk % TCW_EXP_CSTAGES;
        unsigned char *st = tcw_exp_smem + (size_t)s * Cfg::kStageBytes1;
        mbar_arrive_expect_tx(&full[s], Cfg::kStageBytes1);
        bulk_g2s(st, Wt + (size_t)chunk * TCW_EXP_KC * TN * TCW_EXP_WP, Cfg::kWBytes, &full[s]);
        bulk_g2s(st + Cfg::kWBytes, Xt + (size_t)chunk * TCW_EXP_KC * 8, Cfg::kXBytes1, &full[s]);
    };

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < TCW_EXP_CSTAGES; s++) mbar_init(&full[s], 1);
#ifdef TCW_EXP_DECOUPLE
        for (int s = 0; s < TCW_EXP_CSTAGES; s++) mbar_init(&empty[s], NT / 32);
#endif
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (int c = 0; c < TCW_EXP_PREFETCH && c < nchunks; c++) issue(c);
    }

    // accumulators are pairs of adjacent tau columns: one FFMA2 (fma.rn.f32x2, Blackwell packed
    // FP32) updates two cells, halving the issue slots of the inner loop
    float2 acc[TCW_NCH][RM][RN / 2];
#pragma unroll
    for (int c = 0; c < TCW_NCH; c++)
#pragma unroll
        for (int r = 0; r < RM; r++)
#pragma unroll
            for (int j = 0; j < RN / 2; j++) acc[c][r][j] = make_float2(0.0f, 0.0f);

    for (int chunk = 0; chunk < nchunks; chunk++) {
#ifdef TCW_EXP_DECOUPLE
        // no CTA-wide barrier per chunk: every warp signals `empty` when it is done with a stage, and
        // the refill goes into the stage consumed TWO iterations ago (its `empty` phase has normally
        // completed long before), so the warps of a CTA may drift apart by a chunk
        if (tid == 0 && chunk + TCW_EXP_PREFETCH < nchunks) {
            const int cn = chunk + TCW_EXP_PREFETCH;  // uses the stage of chunk cn - CSTAGES
            if (cn >= TCW_EXP_CSTAGES)
                mbar_wait(&empty[cn % TCW_EXP_CSTAGES], (uint32_t)(((cn / TCW_EXP_CSTAGES) - 1) & 1));
            issue(cn);
        }
#else
        // refill the stage consumed in the previous iteration (all threads passed its sync)
        if (tid == 0 && chunk + TCW_EXP_PREFETCH < nchunks) issue(chunk + TCW_EXP_PREFETCH);
#endif
        const int s = chunk % TCW_EXP_CSTAGES;
        mbar_wait(&full[s], (uint32_t)((chunk / TCW_EXP_CSTAGES) & 1));
        const float *Ws = reinterpret_cast<const float *>(tcw_exp_smem + (size_t)s * Cfg::kStageBytes1);
        const float4 *xrow = reinterpret_cast<const float4 *>(Ws + TCW_EXP_KC * TN * TCW_EXP_WP) + tm * RM * 2;  // + 2*(k + r)
        const float *wrow = Ws + tn * RN;                                                         // + k*TN

        // sliding window of 4 consecutive atoms per channel: value with relative index q
        // lives in slot q & 3
        float xr[TCW_NCH][4];
#pragma unroll
        for (int q = 0; q < 3; q++) {
            const float4 lo = xrow[2 * q], hi = xrow[2 * q + 1];
            xr[0][q] = lo.x; xr[1][q] = lo.y; xr[2][q] = lo.z; xr[3][q] = lo.w;
            xr[4][q] = hi.x; xr[5][q] = hi.y; xr[6][q] = hi.z;
        }
#pragma unroll 1
        for (int kk = 0; kk < TCW_EXP_KC; kk += 4) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int k = kk + u;
                {  // the one new atom of this step: all 7 channels in two 128-bit loads
                    const float4 lo = xrow[2 * (k + 3)], hi = xrow[2 * (k + 3) + 1];
                    const int q = (u + 3) & 3;
                    xr[0][q] = lo.x; xr[1][q] = lo.y; xr[2][q] = lo.z; xr[3][q] = lo.w;
                    xr[4][q] = hi.x; xr[5][q] = hi.y; xr[6][q] = hi.z;
                }
                float2 w1[RN / 2], w2[RN / 2];
                if (RN == 4) {
                    const float4 wv = *reinterpret_cast<const float4 *>(wrow + k * TN * TCW_EXP_WP);
                    w1[0] = make_float2(wv.x, wv.y);
                    w1[RN / 2 - 1] = make_float2(wv.z, wv.w);
                    if (TCW_EXP_WP == 2) {
                        const float4 wq = *reinterpret_cast<const float4 *>(wrow + k * TN * TCW_EXP_WP + TN);
                        w2[0] = make_float2(wq.x, wq.y);
                        w2[RN / 2 - 1] = make_float2(wq.z, wq.w);
                    }
                } else {
                    w1[0] = *reinterpret_cast<const float2 *>(wrow + k * TN * TCW_EXP_WP);
                    if (TCW_EXP_WP == 2) w2[0] = *reinterpret_cast<const float2 *>(wrow + k * TN * TCW_EXP_WP + TN);
                }
                if (TCW_EXP_WP == 1) {
#pragma unroll
                    for (int j = 0; j < RN / 2; j++) w2[j] = __fmul2_rn(w1[j], w1[j]);
                }
#pragma unroll
                for (int c = 0; c < TCW_NCH; c++)
#pragma unroll
                    for (int r = 0; r < RM; r++) {
                        const float xv = xr[c][(u + r) & 3];
                        const float2 xx = make_float2(xv, xv);
#pragma unroll
                        for (int j = 0; j < RN / 2; j++)
                            acc[c][r][j] = __ffma2_rn(xx, c < 3 ? w2[j] : w1[j], acc[c][r][j]);
                    }
            }
        }
#ifdef TCW_EXP_DECOUPLE
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive_plain(&empty[s]);
#else
        __syncthreads();  // everyone is done with stage s before it is refilled
#endif
    }

    // ---- fused epilogue: F, optional store, max/argmax, degenerate flag ----
    float *Ft = Fmn ? Fmn + (size_t)tz * w.N_t0 * w.pitch : nullptr;
    float best = -1.0f;
    uint32_t best_flat = 0;
    bool degenerate = false;
#pragma unroll
    for (int r = 0; r < RM; r++) {
        const uint32_t m = m0 + tm * RM + r;
#pragma unroll
        for (int j = 0; j < RN; j++) {
            const uint32_t n = n0 + tn * RN + j;
            if (m < w.N_t0 && n < w.N_tau) {
#define ACC(c_) ((j & 1) ? acc[c_][r][j >> 1].y : acc[c_][r][j >> 1].x)
                const float F = fstat_fast(ACC(0), ACC(1), ACC(2), ACC(3), ACC(4), ACC(5), ACC(6));
#undef ACC
                const uint32_t flat = m * w.N_tau + n;
                if (Ft) Ft[(size_t)m * w.pitch + n] = F;
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
    block_atomic_max_key<NT / 32>(key, &maxkey[t], red);
}

// SLIDE = true: A == 1, the rows of a thread are consecutive atoms -- their 4 atoms per channel form
// a register sliding window (one new 32-byte record per k step).  SLIDE = false: rows are A <= 4
// atoms apart; every row fetches its own record per step (2 broadcast LDS.128 each).
template <class Cfg, bool SLIDE>
__global__ void __launch_bounds__(Cfg::kThreads, (Cfg::kThreads == 256 ? (Cfg::kRN == 4 ? 1 : 2) : 3))
tcw_exp_map_kernel(const float *__restrict__ X8, uint32_t xpad, const float *__restrict__ W,
                   const int32_t *__restrict__ Kn, uint32_t KW, const TplMeta *__restrict__ meta,
                   const int32_t *__restrict__ shift, int t_base, MapWindow w, ExpClasses ec,
                   float *__restrict__ Fmn, unsigned long long *__restrict__ maxkey, uint32_t *__restrict__ flags) {
    constexpr int TM = Cfg::kTM, TN = Cfg::kTN, RM = Cfg::kRM, RN = Cfg::kRN;
    constexpr int NT = Cfg::kThreads;
    extern __shared__ __align__(128) unsigned char tcw_exp_smem[];
    __shared__ __align__(8) uint64_t full[TCW_EXP_STAGES];
    __shared__ unsigned long long red[NT / 32];

    const int tz = blockIdx.z;
    const int t = t_base + tz;
    const uint32_t numAtoms = meta[t].numAtoms;
    const uint32_t nt = blockIdx.x;
    // row class of this row tile and the tile's first row within the class (selects, not indexing:
    // a dynamically indexed kernel parameter would be copied to local memory)
    uint32_t cls = 0, ybeg = 0, i00c = ec.i00[0];
#pragma unroll
    for (int c = 1; c < TCW_EXP_PMAX; c++)
        if ((uint32_t)c < ec.P && blockIdx.y >= ec.ybeg[c]) {
            cls = c;
            ybeg = ec.ybeg[c];
            i00c = ec.i00[c];
        }
    const uint32_t A = SLIDE ? 1u : ec.A;
    const uint32_t j0 = (blockIdx.y - ybeg) * TM;                // row index within the class
    const uint32_t n_rows = (w.N_t0 - cls + ec.P - 1) / ec.P;    // rows of the class
    const uint32_t n0 = nt * TN;
    const int64_t s_base64 = (int64_t)i00c + shift[t] + (int64_t)j0 * A;  // i_t0 of the tile's first row
    const uint32_t s_base = (uint32_t)s_base64;
    const uint32_t n_last = min(n0 + TN, w.N_tau) - 1;
    const int32_t *Knc = Kn + (size_t)cls * w.N_tau;
    const int k_end = (int)min((int64_t)Knc[n_last] + 1, (int64_t)numAtoms - s_base64);
    const int nchunks = k_end > 0 ? (k_end + TCW_EXP_KC - 1) / TCW_EXP_KC : 0;
    const float *Xt = X8 + ((size_t)t * xpad + s_base) * 8;  // 32-byte atom records: always 16-byte aligned
    const uint32_t n_tiles = (w.N_tau + TN - 1) / TN;
    const float *Wt = W + ((size_t)cls * n_tiles + nt) * KW * TN * TCW_EXP_WP;
    constexpr int XBYTES = SLIDE ? Cfg::kXBytes1 : Cfg::kXBytes;
    constexpr int STAGE = SLIDE ? Cfg::kStageBytes1 : Cfg::kStageBytes;

    const int tid = threadIdx.x;
    const int tm = tid >> 4, tn = tid & 15;

    auto issue = [&](int chunk) {
        const int s = chunk % TCW_EXP_STAGES;
        unsigned char *st = tcw_exp_smem + (size_t)s * STAGE;
        mbar_arrive_expect_tx(&full[s], Cfg::kWBytes + XBYTES);
        bulk_g2s(st, Wt + (size_t)chunk * TCW_EXP_KC * TN * TCW_EXP_WP, Cfg::kWBytes, &full[s]);
        bulk_g2s(st + Cfg::kWBytes, Xt + (size_t)chunk * TCW_EXP_KC * 8, XBYTES, &full[s]);
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

    // accumulators are pairs of adjacent tau columns: one FFMA2 (fma.rn.f32x2, Blackwell packed
    // FP32) updates two cells, halving the issue slots of the inner loop
    float2 acc[TCW_NCH][RM][RN / 2];
#pragma unroll
    for (int c = 0; c < TCW_NCH; c++)
#pragma unroll
        for (int r = 0; r < RM; r++)
#pragma unroll
            for (int j = 0; j < RN / 2; j++) acc[c][r][j] = make_float2(0.0f, 0.0f);

    for (int chunk = 0; chunk < nchunks; chunk++) {
        // refill the stage consumed in the previous iteration (all threads passed its sync)
        if (tid == 0 && chunk + TCW_EXP_STAGES - 1 < nchunks) issue(chunk + TCW_EXP_STAGES - 1);
        const int s = chunk % TCW_EXP_STAGES;
        mbar_wait(&full[s], (uint32_t)((chunk / TCW_EXP_STAGES) & 1));
        const float *Ws = reinterpret_cast<const float *>(tcw_exp_smem + (size_t)s * STAGE);
        const float *wrow = Ws + tn * RN;  // + k*TN

        if (SLIDE) {
            const float4 *xrow = reinterpret_cast<const float4 *>(Ws + TCW_EXP_KC * TN * TCW_EXP_WP) + tm * RM * 2;  // + 2*(k + r)
            // sliding window of 4 consecutive atoms per channel: value with relative index q
            // lives in slot q & 3
            float xr[TCW_NCH][4];
#pragma unroll
            for (int q = 0; q < 3; q++) {
                const float4 lo = xrow[2 * q], hi = xrow[2 * q + 1];
                xr[0][q] = lo.x; xr[1][q] = lo.y; xr[2][q] = lo.z; xr[3][q] = lo.w;
                xr[4][q] = hi.x; xr[5][q] = hi.y; xr[6][q] = hi.z;
            }
#pragma unroll 1
            for (int kk = 0; kk < TCW_EXP_KC; kk += 4) {
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int k = kk + u;
                    {  // the one new atom of this step: all 7 channels in two 128-bit loads
                        const float4 lo = xrow[2 * (k + 3)], hi = xrow[2 * (k + 3) + 1];
                        const int q = (u + 3) & 3;
                        xr[0][q] = lo.x; xr[1][q] = lo.y; xr[2][q] = lo.z; xr[3][q] = lo.w;
                        xr[4][q] = hi.x; xr[5][q] = hi.y; xr[6][q] = hi.z;
                    }
                    float2 w1[RN / 2], w2[RN / 2];
                    if (RN == 4) {
                        const float4 wv = *reinterpret_cast<const float4 *>(wrow + k * TN * TCW_EXP_WP);
                        w1[0] = make_float2(wv.x, wv.y);
                        w1[RN / 2 - 1] = make_float2(wv.z, wv.w);
                        if (TCW_EXP_WP == 2) {
                            const float4 wq = *reinterpret_cast<const float4 *>(wrow + k * TN * TCW_EXP_WP + TN);
                            w2[0] = make_float2(wq.x, wq.y);
                            w2[RN / 2 - 1] = make_float2(wq.z, wq.w);
                        }
                    } else {
                        w1[0] = *reinterpret_cast<const float2 *>(wrow + k * TN * TCW_EXP_WP);
                        if (TCW_EXP_WP == 2) w2[0] = *reinterpret_cast<const float2 *>(wrow + k * TN * TCW_EXP_WP + TN);
                    }
                    if (TCW_EXP_WP == 1) {
#pragma unroll
                        for (int j = 0; j < RN / 2; j++) w2[j] = __fmul2_rn(w1[j], w1[j]);
                    }
#pragma unroll
                    for (int c = 0; c < TCW_NCH; c++)
#pragma unroll
                        for (int r = 0; r < RM; r++) {
                            const float xv = xr[c][(u + r) & 3];
                            const float2 xx = make_float2(xv, xv);
#pragma unroll
                            for (int j = 0; j < RN / 2; j++)
                                acc[c][r][j] = __ffma2_rn(xx, c < 3 ? w2[j] : w1[j], acc[c][r][j]);
                        }
                }
            }
        } else {
            // rows A atoms apart: record of (row r, step k) = staged atom (tm*RM + r)*A + k
            const float4 *xbase = reinterpret_cast<const float4 *>(Ws + TCW_EXP_KC * TN * TCW_EXP_WP) + (size_t)tm * RM * A * 2;
#pragma unroll 1
            for (int k = 0; k < TCW_EXP_KC; k++) {
                float2 w1[RN / 2], w2[RN / 2];
                if (RN == 4) {
                    const float4 wv = *reinterpret_cast<const float4 *>(wrow + k * TN * TCW_EXP_WP);
                    w1[0] = make_float2(wv.x, wv.y);
                    w1[RN / 2 - 1] = make_float2(wv.z, wv.w);
                    if (TCW_EXP_WP == 2) {
                        const float4 wq = *reinterpret_cast<const float4 *>(wrow + k * TN * TCW_EXP_WP + TN);
                        w2[0] = make_float2(wq.x, wq.y);
                        w2[RN / 2 - 1] = make_float2(wq.z, wq.w);
                    }
                } else {
                    w1[0] = *reinterpret_cast<const float2 *>(wrow + k * TN * TCW_EXP_WP);
                    if (TCW_EXP_WP == 2) w2[0] = *reinterpret_cast<const float2 *>(wrow + k * TN * TCW_EXP_WP + TN);
                }
                if (TCW_EXP_WP == 1) {
#pragma unroll
                    for (int j = 0; j < RN / 2; j++) w2[j] = __fmul2_rn(w1[j], w1[j]);
                }
#pragma unroll
                for (int r = 0; r < RM; r++) {
                    const float4 lo = xbase[2 * (r * A + k)], hi = xbase[2 * (r * A + k) + 1];
                    const float xv[TCW_NCH] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z};
#pragma unroll
                    for (int c = 0; c < TCW_NCH; c++) {
                        const float2 xx = make_float2(xv[c], xv[c]);
#pragma unroll
                        for (int j = 0; j < RN / 2; j++)
                            acc[c][r][j] = __ffma2_rn(xx, c < 3 ? w2[j] : w1[j], acc[c][r][j]);
                    }
                }
            }
        }
        __syncthreads();  // everyone is done with stage s before it is refilled
    }

    // ---- fused epilogue: F, optional store, max/argmax, degenerate flag ----
    float *Ft = Fmn ? Fmn + (size_t)tz * w.N_t0 * w.pitch : nullptr;
    float best = -1.0f;
    uint32_t best_flat = 0;
    bool degenerate = false;
#pragma unroll
    for (int r = 0; r < RM; r++) {
        const uint32_t jrow = j0 + tm * RM + r;   // row within the class
        const uint32_t m = cls + ec.P * jrow;     // row of the map
#pragma unroll
        for (int j = 0; j < RN; j++) {
            const uint32_t n = n0 + tn * RN + j;
            if (jrow < n_rows && n < w.N_tau) {
#define ACC(c_) ((j & 1) ? acc[c_][r][j >> 1].y : acc[c_][r][j >> 1].x)
                const float F = fstat_fast(ACC(0), ACC(1), ACC(2), ACC(3), ACC(4), ACC(5), ACC(6));
#undef ACC
                const uint32_t flat = m * w.N_tau + n;
                if (Ft) Ft[(size_t)m * w.pitch + n] = F;
                if (F > best) {
                    best = F;
                    best_flat = flat;
                }
                // i_t1 == i_t0 after the reference's clamps: a one-atom window, or a start at / beyond
                // the last atom
                const int K = Knc[n];
                const int64_t s_m = s_base64 + (int64_t)(tm * RM + r) * A;
                if (K >= 0 && (K == 0 || s_m >= (int64_t)numAtoms - 1)) degenerate = true;
            }
        }
    }
    if (degenerate) atomicOr(&flags[t], TCW_FLAG_DEGENERATE);
    const unsigned long long key = best > -1.0f ? pack_key(best, best_flat) : 0ull;
    block_atomic_max_key<NT / 32>(key, &maxkey[t], red);
}

// variants selectable at run time (TCW_EXP_VARIANT = 0/1/2) while tuning; B (1) is the default
typedef ExpCfg<256, 4, 4> ExpCfgA;  // 64 x 64 tile, 1 CTA/SM
typedef ExpCfg<128, 4, 4> ExpCfgB;  // 32 x 64 tile, 3 CTAs/SM (default)
typedef ExpCfg<256, 4, 2> ExpCfgC;  // 64 x 32 tile, 2 CTAs/SM
