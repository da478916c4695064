// tcw_rect.cuh -- tiled rectangular-window map kernel (the HBM/output-bound path).
//
// Replaces pyCUDAkernels/cudaTransientFstatRectWindow.cu (one thread per t0 row, serial loop
// over tau with float32 running sums, uncoalesced stores).  Here every (t0,tau) cell is an
// O(1) difference of FP64 prefix sums,
//     S_c[m,n] = P_c[i_t1(m,n)+1] - P_c[i_t0(m)],
// which goes through the guarded F-stat formula in FP32.
//
// Work decomposition ("skewed" tiles).  With dt0 == dtau (the canonical grids of the
// reference's tests/examples) the cells (m, n) and (m+1, n-1) share the window END time
// t1 = t0_m + tau_n, hence the same end index.  A thread therefore owns a group of R
// consecutive rows and walks d = n + r: one end-prefix fetch from shared memory feeds R
// cells, and each row's start prefix lives in registers.  R = 1 is the plain mapping for
// dt0 != dtau.  A warp's lanes cover 32 consecutive d, so F_mn stores are coalesced rows (the
// reference kernel stores with stride N_tau).
//
// Staging.  A tile's end indices form one contiguous range (monotone under the host-side
// no-wrap certificate); that slice of the 7 FP64 prefix channels is brought into shared
// memory with 1-D TMA bulk copies (cp.async.bulk + mbarrier -> SASS UBLKCP) while the CTA
// computes the tile's end-index table (exact uint32 formulas, once per distinct end time
// instead of once per cell).
//
// Precision modes (measured: FP64->FP32 conversions issue on the XU pipe at 16/clk/SM, so 7
// conversions per cell cap the kernel at ~2 cells/clk/SM):
//   * off-diagonal tiles -- every window of the tile contains the tile's first staged end
//     index rho -- split the difference at rho:
//         S = fl32(P[e+1] - P[rho])  +  fl32(P[rho] - P[s])
//     The first term is tabulated once per tile (FP64 subtract + convert, ~0.25 per cell), the
//     second lives in registers per row, so a cell costs ONE FADD per channel.  Both terms are
//     sub-window sums of the cell's own window: for a2, b2 they are non-negative (no
//     cancellation), for the signed channels the error is that of a two-term float32
//     summation -- tighter than the reference's sequential float32 running sums.
//   * tiles touching the diagonal (short windows, where a common split point does not
//     exist) keep the FP64 difference + conversion per cell.  The launch gives the diagonal
//     its own narrow strip of d (host-chosen width DD), ~1 % of the cells, processed one
//     row at a time to keep the kernel's register footprint at 3 CTAs/SM.
//
// Fused epilogue: per-row running (max F, first d), combined per CTA and published with one
// 64-bit atomicMax per CTA; F_mn is stored only if the caller (or the lnBtSG pass) needs it.
// Tiles without a map edge run a branch-free body.
#pragma once
#include "tcw_common.cuh"
#include "tcw_prep.cuh"

#define TCW_RECT_THREADS 256
#define TCW_RECT_WARPS (TCW_RECT_THREADS / 32)
#define TCW_RECT_DT 1024   // max d values per regular tile (the launch picks DT <= this, a multiple of 32)
#define TCW_RECT_ECAP 1100 // staged end-prefix entries per channel (even)
#define TCW_RECT_G 2       // row groups per warp (a tile has 8 warps x G groups x R rows)
#define TCW_RECT_ROWS(R) (TCW_RECT_WARPS * TCW_RECT_G * (R))
#define TCW_RECT_MAXROWS (TCW_RECT_WARPS * TCW_RECT_G * 4)
#define TCW_RECT_UCAP (TCW_RECT_DT + TCW_RECT_MAXROWS)  // end-index table entries
#define TCW_RECT_SMEM_P (TCW_NCH * TCW_RECT_ECAP * 8)
#define TCW_RECT_SMEM_Q (TCW_NCH * TCW_RECT_ECAP * 4)
#define TCW_RECT_SMEM_E (TCW_RECT_UCAP * 4)
#define TCW_RECT_SMEM_S (TCW_RECT_MAXROWS * 4)
#define TCW_RECT_SMEM_R (TCW_RECT_MAXROWS * 8 * 4)
#define TCW_RECT_SMEM (TCW_RECT_SMEM_P + TCW_RECT_SMEM_E + TCW_RECT_SMEM_S + TCW_RECT_SMEM_R)

// Fast body of one warp: R rows x (32 * n_j) values of d, split-point FP32 sums.
//   CHECKED = false: every (row, d) is a valid cell (off-diagonal tiles have no degenerate cell).
//   TRACK   = false: only the running max value is kept (the lnBtSG pass re-reads F_mn and
//             locates the first cell equal to the final max), saving 2 instructions per cell.
template <int R, bool CHECKED, bool STORE, bool TRACK>
__device__ __forceinline__ void rect_rows_fp32(
    const f32x2 *__restrict__ sQ2, const uint32_t *__restrict__ sE, const f32x2 (&Rs2)[(R + 1) / 2][TCW_NCH],
    float *const (&rowp)[R], const bool (&rowok)[R], uint32_t u_off, uint32_t d0, int j_begin, int j_end,
    uint32_t lane, uint32_t N_tau, uint32_t d_total, uint32_t t1_lane, uint32_t t1_step, uint32_t a0,
    uint32_t t0_data, uint32_t numAtoms, const IndexGeom g, float (&best)[R], uint32_t (&best_d)[R]) {
    const FstatConst2 kc = fstat_const2();
#pragma unroll 4
    for (int j = j_begin; j < j_end; j++) {
        const uint32_t d = d0 + lane + 32u * j;
        if (CHECKED && d >= d_total) break;
        uint32_t idx;  // index of P[e+1] relative to the staged slice
        if (R > 1) {
            idx = sE[u_off + lane + 32u * j];
        } else {
            idx = min(index_t1(t1_lane + (uint32_t)j * t1_step, t0_data, numAtoms, g) + 1 - a0,
                      (uint32_t)(TCW_RECT_ECAP - 1));
        }
        f32x2 Q2[TCW_NCH];  // {q, q}: stored duplicated so that the packed adds need no register shuffling
#pragma unroll
        for (int c = 0; c < TCW_NCH; c++) Q2[c] = sQ2[c * TCW_RECT_ECAP + idx];
        // rows in pairs: both cells share Q (same end index), so every FP32 operation of the
        // pair is one packed instruction (FADD2 with Q broadcast, then fstat_fast2)
        float Fr[R];
        if (R % 2 == 0) {
#pragma unroll
            for (int rp = 0; rp < R / 2; rp++) {
                f32x2 S[TCW_NCH];
#pragma unroll
                for (int c = 0; c < TCW_NCH; c++) S[c] = add2(Q2[c], Rs2[rp][c]);
                fstat_fast2(kc, S[0], S[1], S[2], S[3], S[4], S[5], S[6], Fr[2 * rp], Fr[2 * rp + (R > 1 ? 1 : 0)]);
            }
        } else {
            float lo[TCW_NCH], q[TCW_NCH], hi;
#pragma unroll
            for (int c = 0; c < TCW_NCH; c++) {
                unpack2(Rs2[0][c], lo[c], hi);
                unpack2(Q2[c], q[c], hi);
            }
            Fr[0] = fstat_fast(q[0] + lo[0], q[1] + lo[1], q[2] + lo[2], q[3] + lo[3], q[4] + lo[4], q[5] + lo[5],
                               q[6] + lo[6]);
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            const float F = Fr[r];
            bool valid = true;
            if (CHECKED) valid = rowok[r] && (d - r) < N_tau;  // d - r wraps for d < r
            if (valid) {
                if (STORE) rowp[r][32 * j] = F;
                if (TRACK) {
                    if (F > best[r]) {
                        best[r] = F;
                        best_d[r] = d;
                    }
                } else {
                    best[r] = fmaxf(best[r], F);  // NaN-safe: fmaxf returns the non-NaN operand
                }
            }
        }
    }
}

// Precise body (tiles touching the diagonal, or unstaged): FP64 difference per cell, ONE row
// per call (its start prefix in registers), every cell bounds- and degeneracy-checked.  Scalar
// in/out on purpose: the hot path's per-row state must stay in registers.
struct RectRowResult {
    float best;
    uint32_t best_d;
    uint32_t degenerate;
};
template <int R, bool STAGED, bool STORE>
__device__ __noinline__ RectRowResult rect_row_fp64(
    const double *__restrict__ sP, const uint32_t *__restrict__ sE, const double *__restrict__ Pt, uint32_t ppad,
    uint32_t s_row, float *rowp, uint32_t r, uint32_t u_off, uint32_t d0, int n_j, uint32_t lane, uint32_t N_tau,
    uint32_t d_total, uint32_t t1_lane, uint32_t t1_step, uint32_t a0, uint32_t t0_data, uint32_t numAtoms,
    const IndexGeom g) {
    RectRowResult out;
    out.best = -1.0f;
    out.best_d = r;
    out.degenerate = 0;
    double Ps[TCW_NCH];
#pragma unroll
    for (int c = 0; c < TCW_NCH; c++) Ps[c] = __ldg(Pt + (size_t)c * ppad + s_row);
#pragma unroll 1
    for (int j = 0; j < n_j; j++) {
        const uint32_t d = d0 + lane + 32u * j;
        if (d >= d_total) break;
        if (d - r >= N_tau) continue;  // wraps for d < r
        uint32_t e1;  // absolute index e + 1
        if (R > 1) e1 = sE[u_off + lane + 32u * j] + a0;
        else e1 = index_t1(t1_lane + (uint32_t)j * t1_step, t0_data, numAtoms, g) + 1;
        float S[TCW_NCH];
#pragma unroll
        for (int c = 0; c < TCW_NCH; c++) {
            const double E = STAGED ? sP[c * TCW_RECT_ECAP + (e1 - a0)] : __ldg(Pt + (size_t)c * ppad + e1);
            S[c] = (float)(E - Ps[c]);
        }
        const float F = fstat_fast(S[0], S[1], S[2], S[3], S[4], S[5], S[6]);
        if (STORE) rowp[32 * j] = F;
        if (F > out.best) {
            out.best = F;
            out.best_d = d;
        }
        if (e1 - 1 == s_row) out.degenerate = 1;  // i_t1 == i_t0
    }
    return out;
}

// grid: x = d tiles: 0 = head strip [0, DD) holding the diagonal; b >= 1 = [DD + (b-1) DT, +DT),
//           DT chosen by the host so that the regular tiles divide the map evenly
//       y = tiles of 8 warps x G row groups, z = template in sub-batch
// TRACK = false (only with a following lnBtSG pass over a stored F_mn): publish max values
// only; tcw_btsg_kernel completes the key with the first flat index that attains the max.
template <int R, bool STAGED, bool TRACK>
__global__ void __launch_bounds__(TCW_RECT_THREADS, 3)
tcw_rect_map_kernel(const double *__restrict__ P, uint32_t ppad, const TplMeta *__restrict__ meta,
                    int t_base, MapWindow w, IndexGeom g, uint32_t DD, uint32_t DT, float *__restrict__ Fmn,
                    unsigned long long *__restrict__ maxkey, uint32_t *__restrict__ flags) {
    extern __shared__ __align__(16) unsigned char tcw_rect_smem[];
    unsigned char *sp = tcw_rect_smem;
    double *sP = reinterpret_cast<double *>(sp);        // [7][ECAP]  staged FP64 end prefixes
    f32x2 *sQ2 = reinterpret_cast<f32x2 *>(sp);         // [7][ECAP]  {q, q}, q = fl32(P[i] - P[rho]): IN PLACE over sP
    sp += TCW_RECT_SMEM_P;
    uint32_t *sE = reinterpret_cast<uint32_t *>(sp);    // [UCAP]     end index (rel. to slice) per u
    sp += TCW_RECT_SMEM_E;
    uint32_t *sS = reinterpret_cast<uint32_t *>(sp);    // [ROWS]     start index i_t0 per row
    sp += TCW_RECT_SMEM_S;
    float *sR = reinterpret_cast<float *>(sp);          // [ROWS/2][8][2]  fl32(P[rho] - P[s]), row pairs interleaved
    __shared__ __align__(8) uint64_t bar;
    __shared__ unsigned long long red[TCW_RECT_WARPS];

    const int tz = blockIdx.z;
    const int t = t_base + tz;
    const uint32_t numAtoms = meta[t].numAtoms;
    const uint32_t t0_data = meta[t].t0_data;
    const double *Pt = P + (size_t)t * TCW_NCH * ppad;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    constexpr uint32_t ROWS = TCW_RECT_ROWS(R);  // rows per tile
    const uint32_t d_total = w.N_tau + R - 1;
    const uint32_t m0 = blockIdx.y * ROWS;
    const uint32_t m_last = min(m0 + ROWS, w.N_t0) - 1;
    const uint32_t d0 = blockIdx.x == 0 ? 0u : DD + (blockIdx.x - 1) * DT;
    const uint32_t d_cnt = blockIdx.x == 0 ? DD : DT;
    const uint32_t d_last = min(d0 + d_cnt, d_total) - 1;
    const bool edge_rows = (d0 < (uint32_t)(R - 1)) || (m0 + ROWS > w.N_t0);  // head strip / bottom row tile

    // end time of (row group, d): rows of a group differ by dt0 == dtau (R > 1), absorbed into d
    const uint32_t t1_tile = w.t0 + w.tau + m0 * w.dt0 + d0 * w.dtau;
    const uint32_t e_lo = index_t1(t1_tile, t0_data, numAtoms, g);
    const uint32_t a0 = STAGED ? ((e_lo + 1) & ~1u) : 0u;
    uint32_t cnt = 0;
    if (STAGED) {
        const uint32_t e_hi = index_t1(t1_tile + ((m_last - m0) / R * R) * w.dt0 + (d_last - d0) * w.dtau, t0_data,
                                       numAtoms, g);
        cnt = min((e_hi + 1 - a0 + 1 + 1) & ~1u, (uint32_t)TCW_RECT_ECAP);  // even; <= ECAP by the host check
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            mbar_fence_init();
            mbar_arrive_expect_tx(&bar, TCW_NCH * cnt * (uint32_t)sizeof(double));
#pragma unroll
            for (int c = 0; c < TCW_NCH; c++)
                bulk_g2s(sP + c * TCW_RECT_ECAP, Pt + (size_t)c * ppad + a0, cnt * (uint32_t)sizeof(double), &bar);
        }
    }
    // per-tile index tables, computed cooperatively while the bulk copies are in flight:
    //   sE[u], u = (row - m0)/R*R + (d - d0): end index of t1 = t1_tile + u*dtau  (dt0 == dtau)
    //   sS[row - m0]: start index i_t0 of the row
    if (R > 1) {
        const uint32_t u_cnt = min((uint32_t)TCW_RECT_UCAP, ROWS + d_cnt);
        for (uint32_t u = threadIdx.x; u < u_cnt; u += TCW_RECT_THREADS)
            sE[u] = STAGED ? min(index_t1(t1_tile + u * w.dtau, t0_data, numAtoms, g) + 1 - a0,
                                 (uint32_t)(TCW_RECT_ECAP - 1))  // overhang entries stay in bounds
                           : index_t1(t1_tile + u * w.dtau, t0_data, numAtoms, g) + 1;
    }
    if (threadIdx.x < ROWS) {
        const uint32_t m = min(m0 + threadIdx.x, w.N_t0 - 1);
        sS[threadIdx.x] = index_t0(w.t0 + m * w.dt0, t0_data, numAtoms, g);
    }
    // start prefixes P_c[s_row] for the per-row terms: fetched now, while the bulk copies fly
    double ps_early[(ROWS * 8 + TCW_RECT_THREADS - 1) / TCW_RECT_THREADS];
#pragma unroll
    for (int k = 0; k < (int)((ROWS * 8 + TCW_RECT_THREADS - 1) / TCW_RECT_THREADS); k++) {
        const uint32_t i = threadIdx.x + k * TCW_RECT_THREADS;
        const uint32_t row = i >> 3, c = i & 7;
        ps_early[k] = 0.0;
        if (i < ROWS * 8 && c < TCW_NCH) {
            const uint32_t m = min(m0 + row, w.N_t0 - 1);
            ps_early[k] = __ldg(Pt + (size_t)c * ppad + index_t0(w.t0 + m * w.dt0, t0_data, numAtoms, g));
        }
    }
    const uint32_t s_hi = index_t0(w.t0 + m_last * w.dt0, t0_data, numAtoms, g);
    // off-diagonal: the split point rho = a0 lies strictly inside every window of the tile,
    // s < rho <= e + 1 with e > s (so no cell of the tile is degenerate): rho >= s_hi + 2
    const bool offdiag = STAGED && (a0 >= s_hi + 2);
    float *Ft = Fmn ? Fmn + (size_t)tz * w.N_t0 * w.pitch : nullptr;

    __syncthreads();  // sE, sS visible; mbarrier init visible to all waiters
    if (STAGED) mbar_wait(&bar, 0);
    if (offdiag) {
        // off-diagonal tiles no longer need the FP64 slice itself: convert it IN PLACE to
        // {q, q} pairs, q = fl32(P_c[a0+i] - P_c[rho]) (each thread rewrites only the 8-byte slots
        // it read), after everyone has fetched P[rho] and the per-row terms fl32(P[rho] - P[s])
        double pref[TCW_NCH];
#pragma unroll
        for (int c = 0; c < TCW_NCH; c++) pref[c] = sP[c * TCW_RECT_ECAP];
#pragma unroll
        for (int k = 0; k < (int)((ROWS * 8 + TCW_RECT_THREADS - 1) / TCW_RECT_THREADS); k++) {
            const uint32_t i = threadIdx.x + k * TCW_RECT_THREADS;
            const uint32_t row = i >> 3, c = i & 7;
            // stored as row PAIRS {row 2p, row 2p+1} per channel: a 64-bit load yields a packed operand
            if (i < ROWS * 8 && c < TCW_NCH)
                sR[(((row >> 1) * 8 + c) << 1) + (row & 1)] = (float)(sP[c * TCW_RECT_ECAP] - ps_early[k]);
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < cnt; i += TCW_RECT_THREADS) {
#pragma unroll
            for (int c = 0; c < TCW_NCH; c++) {
                const float q = (float)(sP[c * TCW_RECT_ECAP + i] - pref[c]);
                sQ2[c * TCW_RECT_ECAP + i] = pack2(q, q);
            }
        }
        __syncthreads();
    }

    const uint32_t t1_step = 32u * w.dtau;
    const int n_j = (int)(d_cnt / 32);
    const int j_full = w.N_tau > d0 ? (int)min((w.N_tau - d0) / 32u, (uint32_t)n_j) : 0;  // fully valid chunks
    unsigned long long key = 0ull;
    uint32_t degenerate = 0;
    // each warp walks TCW_RECT_G row groups of R rows: the tile's staging cost is shared
#pragma unroll 1
    for (uint32_t gi = 0; gi < TCW_RECT_G; gi++) {
        const uint32_t grow = (gi * TCW_RECT_WARPS + warp) * R;  // first row of the group, relative to m0
        if (m0 + grow >= w.N_t0) break;
        const uint32_t u_off = grow;
        const uint32_t t1_lane = t1_tile + grow * w.dt0 + lane * w.dtau;  // used by R == 1 only
        float best[R];
        uint32_t best_d[R];
        float *rowp[R];
        bool rowok[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const uint32_t m = m0 + grow + r;
            best[r] = -1.0f;  // maxF starts at -1, strict > (tcw:135-139)
            best_d[r] = r;
            rowok[r] = m < w.N_t0;
            const uint32_t mc = rowok[r] ? m : 0u;
            // cell (m, n = d - r) with d = d0 + lane + 32 j  ->  rowp[r][32 j]
            rowp[r] = Ft ? Ft + ((size_t)mc * w.pitch + d0 + lane) - r : nullptr;
        }
        if (offdiag) {
            f32x2 Rs2[(R + 1) / 2][TCW_NCH];  // {row 2rp, row 2rp+1} pairs of fl32(P[rho] - P[s])
#pragma unroll
            for (int rp = 0; rp < (R + 1) / 2; rp++) {
                const f32x2 *pr = reinterpret_cast<const f32x2 *>(sR) + ((grow >> 1) + rp) * 8;
#pragma unroll
                for (int c = 0; c < TCW_NCH; c++) Rs2[rp][c] = pr[c];
                if (R == 1 && (grow & 1)) {  // odd single row: its value sits in the high half
#pragma unroll
                    for (int c = 0; c < TCW_NCH; c++) {
                        float lo, hi;
                        unpack2(Rs2[rp][c], lo, hi);
                        Rs2[rp][c] = pack2(hi, hi);
                    }
                }
            }
#define RECT_FAST(CHK_, STORE_, J0_, J1_)                                                                       \
    rect_rows_fp32<R, CHK_, STORE_, TRACK>(sQ2, sE, Rs2, rowp, rowok, u_off, d0, J0_, J1_, lane, w.N_tau, d_total, \
                                           t1_lane, t1_step, a0, t0_data, numAtoms, g, best, best_d)
            if (edge_rows) {
                if (Ft) RECT_FAST(true, true, 0, n_j);
                else RECT_FAST(true, false, 0, n_j);
            } else {
                // chunks of 32 d that are valid for every lane and row run unchecked; only the
                // chunk(s) straddling the map's right edge are bounds-checked
                if (Ft) {
                    RECT_FAST(false, true, 0, j_full);
                    if (j_full < n_j) RECT_FAST(true, true, j_full, n_j);
                } else {
                    RECT_FAST(false, false, 0, j_full);
                    if (j_full < n_j) RECT_FAST(true, false, j_full, n_j);
                }
            }
#undef RECT_FAST
        } else {
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (!rowok[r]) continue;
                const uint32_t s_row = sS[grow + r];
                const RectRowResult rr =
                    Ft ? rect_row_fp64<R, STAGED, true>(sP, sE, Pt, ppad, s_row, rowp[r], r, u_off, d0, n_j, lane,
                                                        w.N_tau, d_total, t1_lane, t1_step, a0, t0_data, numAtoms, g)
                       : rect_row_fp64<R, STAGED, false>(sP, sE, Pt, ppad, s_row, rowp[r], r, u_off, d0, n_j, lane,
                                                         w.N_tau, d_total, t1_lane, t1_step, a0, t0_data, numAtoms, g);
                best[r] = rr.best;
                best_d[r] = rr.best_d;
                degenerate |= rr.degenerate;
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            if (best[r] > -1.0f) {
                // TRACK == false: index part 0 (flat = 0xFFFFFFFF), completed by the lnBtSG pass
                const uint32_t flat = TRACK ? (m0 + grow + r) * w.N_tau + (best_d[r] - r) : 0xFFFFFFFFu;
                const unsigned long long k = pack_key(best[r], flat);
                key = k > key ? k : key;
            }
        }
    }
    if (degenerate) atomicOr(&flags[t], TCW_FLAG_DEGENERATE);
    block_atomic_max_key<TCW_RECT_WARPS>(key, &maxkey[t], red);
}
