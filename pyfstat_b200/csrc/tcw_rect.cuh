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
// conversions per cell cap the kernel at ~2 cells/clk/SM).  Every cell is evaluated as
//         S = fl32(P[e+1] - P[rho])  +  fl32(P[rho] - P[s])
// with a split point rho chosen so that both terms are (close to) sub-window sums of the cell's
// own window -- for a2, b2 no cancellation, for the signed channels the error of a two-term
// float32 summation, tighter than the reference's sequential float32 running sums:
//   * off-diagonal tiles -- every window of the tile contains the tile's first staged end
//     index -- use ONE rho per tile: the first term is tabulated once per tile (FP64 subtract +
//     convert, ~0.15 per cell), the second lives in registers per row, so a cell costs ONE
//     packed FADD per channel and row pair;
//   * tiles touching the diagonal (short windows, where a tile-wide split point does not
//     exist; ~4 % of the cells) use one rho per GROUP of R rows, rho = 1 + the group's last
//     start index: the first term is built per (group, d) on the fly (7 FP64 subtracts +
//     conversions per R cells).  Windows shorter than the group's span see a negative first
//     term; its magnitude is at most R-1 row steps, so with tau_min >= dtau (host-checked,
//     otherwise R = 1) the cancellation costs a few ulp;
//   * R == 1 or unstaged diagonal tiles keep the FP64 difference + conversion per cell.
//
// The ab channel is carried as C' = -2 * sum(ab) (exact scaling applied when the FP32 terms are
// formed): det = A B - C'^2/4 and num = B|Fa|^2 + A|Fb|^2 + C' Re(Fa conj Fb) need one packed
// instruction and one constant less.
//
// Fused epilogue: the map kernel tracks max VALUES only (one 3-input FMNMX per two cells); the
// per-tile maxima go to a small table and the argmax is completed afterwards, either by the
// lnBtSG pass (which re-reads F_mn anyway) or by tcw_rect_locate_kernel, which re-evaluates
// only the tile(s) whose maximum equals the template maximum, with first-occurrence tracking.
// Conditioning guard: interior chunks run 4 x R cells per lane unguarded while folding the
// margins into one NaN-propagating 3-input minimum; a chunk with any non-positive margin is
// re-evaluated with the per-cell select (F = 2 fallback of the reference).
#pragma once
#include "tcw_common.cuh"
#include "tcw_prep.cuh"

#define TCW_RECT_THREADS 256
#ifndef TCW_RECT_MINB
#define TCW_RECT_MINB 2  // CTAs per SM the register allocation targets.  Measured (60 d, T=64, F_mn stored): R=4 at
                         // 2 CTAs/SM x 128 registers 0.64 ms; R=4 at 3 x 80 (spills) 0.67; R=2 rows per thread at
                         // 3 x 80 0.67, at 4 x 64 0.74: occupancy does not buy back the shared Q fetches
#endif
#define TCW_RECT_WARPS (TCW_RECT_THREADS / 32)
#ifndef TCW_RECT_DT
#define TCW_RECT_DT 1472   // max d values per regular tile (the launch picks DT <= this, a multiple of 32)
#define TCW_RECT_ECAP 1608 // staged end-prefix entries per channel (even)
#endif
#define TCW_RECT_GMAX 4    // max row groups per warp (a tile has 8 warps x G groups x R rows; G is a launch parameter)
#define TCW_RECT_MAXROWS (TCW_RECT_WARPS * TCW_RECT_GMAX * 4)
#define TCW_RECT_UCAP (TCW_RECT_DT + TCW_RECT_MAXROWS)  // end-index table entries
#define TCW_RECT_JB 4      // d chunks (of 32) per guarded block of the interior loop
#define TCW_RECT_SMEM_P (TCW_NCH * TCW_RECT_ECAP * 8)
#define TCW_RECT_SMEM_E (TCW_RECT_UCAP * 4)
#define TCW_RECT_SMEM_S (TCW_RECT_MAXROWS * 4)
#define TCW_RECT_SMEM_R (TCW_RECT_MAXROWS * 8 * 4)
#define TCW_RECT_SMEM_G (TCW_RECT_MAXROWS / 4 * 8 * 8)
#define TCW_RECT_SMEM (TCW_RECT_SMEM_P + TCW_RECT_SMEM_E + TCW_RECT_SMEM_S + TCW_RECT_SMEM_R + TCW_RECT_SMEM_G)

// exact power-of-two scaling of channel c when the FP32 terms are formed: ab is carried as -2 ab
__device__ __forceinline__ float rect_chan_scale(int c) { return c == 2 ? -2.0f : 1.0f; }

// The R cells that share one end index (one lane, one d): window sums S = Q + R-term per row,
// then F = num / det and the conditioning margin, both UNGUARDED (the caller applies the guard).
//   DIAG = false: Q comes from the per-tile table of {q, q} pairs;
//   DIAG = true : Q is built here from the staged FP64 slice and the group's P[rho].
template <int R, bool DIAG, bool NOGUARD = false>
__device__ __forceinline__ void rect_eval(const f32x2 *__restrict__ sQ2, const double *__restrict__ sP,
                                          const double *__restrict__ sGg, uint32_t idx,
                                          const f32x2 (&Rs2)[(R + 1) / 2][TCW_NCH], const FstatConst2 &kc,
                                          float (&F)[R], float (&mg)[R]) {
    f32x2 Q2[TCW_NCH];  // {q, q}: duplicated so that the packed adds need no register shuffling
#pragma unroll
    for (int c = 0; c < TCW_NCH; c++) {
        if (DIAG) {
            const float q = (float)(sP[c * TCW_RECT_ECAP + idx] - sGg[c]) * rect_chan_scale(c);
            Q2[c] = pack2(q, q);
        } else {
            Q2[c] = sQ2[c * TCW_RECT_ECAP + idx];
        }
    }
    // rows in pairs: both cells share Q (same end index), so every FP32 operation of the pair is
    // one packed instruction.  An odd R evaluates its last row in both halves of a pair.
#pragma unroll
    for (int rp = 0; rp < (R + 1) / 2; rp++) {
        f32x2 S[TCW_NCH];
#pragma unroll
        for (int c = 0; c < TCW_NCH; c++) S[c] = add2(Q2[c], Rs2[rp][c]);
        float f0, f1, m0, m1;
        fstat_core2<NOGUARD>(kc, S[0], S[1], S[2], S[3], S[4], S[5], S[6], f0, f1, m0, m1);
        F[2 * rp] = f0;
        mg[2 * rp] = m0;
        if (2 * rp + 1 < R) {
            F[2 * rp + 1] = f1;
            mg[2 * rp + 1] = m1;
        }
    }
}

// One cell straight from the FP64 prefixes (start prefix from global memory): the few cells of a
// diagonal group whose window ends before the group's split point.
__device__ __noinline__ float rect_cell_fp64(const double *__restrict__ sP, uint32_t idx,
                                             const double *__restrict__ Pt, uint32_t ppad, uint32_t s_row) {
    float S[TCW_NCH];
#pragma unroll
    for (int c = 0; c < TCW_NCH; c++)
        S[c] = (float)(sP[c * TCW_RECT_ECAP + idx] - __ldg(Pt + (size_t)c * ppad + s_row));
    return fstat_fast(S[0], S[1], S[2], S[3], S[4], S[5], S[6]);
}

// Per-row running state of a group.  TRACK = false keeps one max VALUE for the whole group.
template <int R>
struct RectBest {
    float v[R];
    uint32_t d[R];
};

// One chunk (32 d, one per lane) of a group with the per-cell guard, bounds checks (CHECKED),
// degenerate-cell detection (DIAG) and first-occurrence tracking (TRACK).
template <int R, bool DIAG, bool CHECKED, bool STORE, bool TRACK>
__device__ __forceinline__ void rect_chunk_careful(
    const f32x2 *__restrict__ sQ2, const double *__restrict__ sP, const double *__restrict__ sGg, uint32_t idx,
    const f32x2 (&Rs2)[(R + 1) / 2][TCW_NCH], const FstatConst2 &kc, float *const (&rowp)[R], const bool (&rowok)[R],
    const uint32_t (&srow)[R], uint32_t e_abs, uint32_t d, int j, uint32_t N_tau, const double *__restrict__ Pt,
    uint32_t ppad, RectBest<R> &best, float &vmax, uint32_t &degenerate) {
    float F[R], mg[R];
    rect_eval<R, DIAG>(sQ2, sP, sGg, idx, Rs2, kc, F, mg);
    // DIAG: a window ending before the group's split point rho = srow[R-1] + 1 would see
    // cancellation between the two terms -> evaluate those few cells exactly
    const bool before_rho = DIAG && e_abs < srow[R - 1];  // e + 1 < rho
#pragma unroll
    for (int r = 0; r < R; r++) {
        float Fr = mg[r] > 0.0f ? F[r] : 2.0f;  // NaN margin -> fallback, as the reference's cond test
        bool valid = true;
        if (CHECKED) valid = rowok[r] && (d - r) < N_tau;  // d - r wraps for d < r
        if (DIAG && before_rho && valid) Fr = rect_cell_fp64(sP, idx, Pt, ppad, srow[r]);
        if (valid) {
            if (STORE) rowp[r][32 * j] = Fr;
            if (TRACK) {
                if (Fr > best.v[r]) {
                    best.v[r] = Fr;
                    best.d[r] = d;
                }
            } else {
                vmax = fmaxf(vmax, Fr);  // NaN-safe: fmaxf returns the non-NaN operand
            }
            if (DIAG && e_abs == srow[r]) degenerate = 1;  // i_t1 == i_t0
        }
    }
}

// All chunks [j_begin, j_end) of one group.
//   CHECKED = false: every (row, d) is a valid cell; with TRACK = false the chunks run in blocks
//   of TCW_RECT_JB with the guard folded into one minimum per block (see the header).
//   NOGUARD = true (tile-wide conditioning certificate, see tcw_rect_p.cuh): the unguarded block
//   path neither computes nor tests the margins.
template <int R, bool DIAG, bool CHECKED, bool STORE, bool TRACK, bool NOGUARD = false>
__device__ __forceinline__ void rect_rows(
    const f32x2 *__restrict__ sQ2, const double *__restrict__ sP, const double *__restrict__ sGg,
    const uint32_t *__restrict__ sE, const f32x2 (&Rs2)[(R + 1) / 2][TCW_NCH], float *const (&rowp)[R],
    const bool (&rowok)[R], const uint32_t (&srow)[R], uint32_t u_off, uint32_t d0, int j_begin, int j_end,
    uint32_t lane, uint32_t N_tau, uint32_t d_total, uint32_t t1_lane, uint32_t t1_step, uint32_t a0,
    uint32_t t0_data, uint32_t numAtoms, const IndexGeom g, const double *__restrict__ Pt, uint32_t ppad,
    RectBest<R> &best, float &vmax, uint32_t &degenerate) {
    const FstatConst2 kc = fstat_const2();
    auto end_index = [&](int j) -> uint32_t {  // index of P[e+1] relative to the staged slice
        if (R > 1) return sE[u_off + lane + 32u * j];
        return min(index_t1(t1_lane + (uint32_t)j * t1_step, t0_data, numAtoms, g) + 1 - a0,
                   (uint32_t)(TCW_RECT_ECAP - 1));
    };
    // A run of chunks is "plain" if every (row, d) in it is a cell of the map whose window ends
    // at or after the group's split point (so it is neither degenerate nor in need of the exact
    // evaluation): such runs take the unguarded block path.  All conditions are warp-uniform.
    constexpr int JBX = DIAG ? 1 : TCW_RECT_JB;
    bool allrows = true;
#pragma unroll
    for (int r = 0; r < R; r++) allrows = allrows && rowok[r];
    auto plain = [&](int j) -> bool {
        if (!CHECKED) return true;
        const uint32_t d_lo = d0 + 32u * j;
        bool ok = allrows && d_lo >= (uint32_t)(R - 1) && d_lo + 32u * JBX <= N_tau;
        if (DIAG && ok) ok = sE[u_off + 32u * j] + a0 - 1u > srow[R - 1];  // lane 0's end index: the smallest
        return ok;
    };
    int j = j_begin;
#pragma unroll 1
    while (j < j_end) {
        if (!TRACK && j + JBX <= j_end && plain(j)) {
            const float vmax_in = vmax;
            float mmin = 1.0f;
#pragma unroll
            for (int u = 0; u < JBX; u++) {
                float F[R], mg[R];
                rect_eval<R, DIAG, NOGUARD>(sQ2, sP, sGg, end_index(j + u), Rs2, kc, F, mg);
#pragma unroll
                for (int r = 0; r < R; r++)
                    if (STORE) rowp[r][32 * (j + u)] = F[r];
                if (R % 2 == 0) {
#pragma unroll
                    for (int r = 0; r < R; r += 2) {
                        vmax = fmax3(vmax, F[r], F[r + 1]);
                        if (!NOGUARD) mmin = fmin3_nan(mmin, mg[r], mg[r + 1]);
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        vmax = fmaxf(vmax, F[r]);
                        mmin = fmin3_nan(mmin, mg[r], mg[r]);
                    }
                }
            }
            if (!NOGUARD && !(mmin > 0.0f)) {  // rare: some cell of the block needs the F = 2 fallback -> redo guarded
                vmax = vmax_in;
#pragma unroll 1
                for (int u = 0; u < JBX; u++) {
                    const uint32_t idx = end_index(j + u);
                    rect_chunk_careful<R, DIAG, false, STORE, false>(sQ2, sP, sGg, idx, Rs2, kc, rowp, rowok, srow,
                                                                     idx + a0 - 1u, 0u, j + u, N_tau, Pt, ppad, best,
                                                                     vmax, degenerate);
                }
            }
            j += JBX;
        } else {
            const uint32_t d = d0 + lane + 32u * j;
            if (!CHECKED || d < d_total) {
                const uint32_t idx = end_index(j);
                rect_chunk_careful<R, DIAG, CHECKED, STORE, TRACK>(sQ2, sP, sGg, idx, Rs2, kc, rowp, rowok, srow,
                                                                    idx + a0 - 1u, d, j, N_tau, Pt, ppad, best, vmax,
                                                                    degenerate);
            }
            j++;
        }
    }
}

// Precise body (R == 1 or unstaged tiles touching the diagonal): FP64 difference per cell, ONE
// row per call (its start prefix in registers), every cell bounds- and degeneracy-checked.
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

// One tile.  bx = d tile: 0 = head strip [0, DD) holding the diagonal; b >= 1 = [DD + (b-1) DT, +DT),
//                DT chosen by the host so that the regular tiles divide the map evenly
//            by = tile of 8 warps x G row groups x R rows, tz = template in sub-batch
// TRACK = false: returns the tile's max VALUE in the high half of the key (index part 0).
// TRACK = true : full key (value, first flat index); nothing is stored.
template <int R, bool STAGED, bool TRACK>
__device__ __forceinline__ unsigned long long rect_tile(
    const double *__restrict__ P, uint32_t ppad, const TplMeta *__restrict__ meta, int t, int tz, uint32_t bx,
    uint32_t by, const MapWindow &w, const IndexGeom &g, uint32_t DD, uint32_t DT, uint32_t G,
    float *__restrict__ Fmn, uint32_t *__restrict__ flags, uint32_t *__restrict__ gmax, uint32_t top,
    unsigned char *smem, uint64_t *bar, uint32_t phase, bool init_bar, unsigned long long *red) {
    unsigned char *sp = smem;
    double *sP = reinterpret_cast<double *>(sp);        // [7][ECAP]  staged FP64 end prefixes
    f32x2 *sQ2 = reinterpret_cast<f32x2 *>(sp);         // [7][ECAP]  {q, q}, q = fl32(P[i] - P[rho]): IN PLACE over sP
    sp += TCW_RECT_SMEM_P;
    uint32_t *sE = reinterpret_cast<uint32_t *>(sp);    // [UCAP]     end index (rel. to slice) per u
    sp += TCW_RECT_SMEM_E;
    uint32_t *sS = reinterpret_cast<uint32_t *>(sp);    // [ROWS]     start index i_t0 per row
    sp += TCW_RECT_SMEM_S;
    float *sR = reinterpret_cast<float *>(sp);          // [ROWS/2][8][2]  fl32(P[rho] - P[s]), row pairs interleaved
    sp += TCW_RECT_SMEM_R;
    double *sG = reinterpret_cast<double *>(sp);        // [ROWS/R][8]  P[rho] of each row group (diagonal tiles)

    const uint32_t numAtoms = meta[t].numAtoms;
    const uint32_t t0_data = meta[t].t0_data;
    const double *Pt = P + (size_t)t * TCW_NCH * ppad;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    const uint32_t ROWS = TCW_RECT_WARPS * G * R;  // rows per tile
    const uint32_t d_total = w.N_tau + R - 1;
    const uint32_t m0 = by * ROWS;
    const uint32_t m_last = min(m0 + ROWS, w.N_t0) - 1;
    const uint32_t d0 = bx == 0 ? 0u : DD + (bx - 1) * DT;
    const uint32_t d_cnt = bx == 0 ? DD : DT;
    const uint32_t d_last = min(d0 + d_cnt, d_total) - 1;

    // end time of (row group, d): rows of a group differ by dt0 == dtau (R > 1), absorbed into d
    const uint32_t t1_tile = w.t0 + w.tau + m0 * w.dt0 + d0 * w.dtau;
    const uint32_t e_lo = index_t1(t1_tile, t0_data, numAtoms, g);
    const uint32_t a0 = STAGED ? ((e_lo + 1) & ~1u) : 0u;
    uint32_t cnt = 0;
    if (STAGED) {
        const uint32_t e_hi = index_t1(t1_tile + ((m_last - m0) / R * R) * w.dt0 + (d_last - d0) * w.dtau, t0_data,
                                       numAtoms, g);
        cnt = min((e_hi + 1 - a0 + 1 + 1) & ~1u, (uint32_t)TCW_RECT_ECAP);  // even; <= ECAP by the host check
        if (threadIdx.x == 0) {  // one barrier phase per tile; initialised with the CTA's first tile
            if (init_bar) {
                mbar_init(bar, 1);
                mbar_fence_init();
            }
            mbar_arrive_expect_tx(bar, TCW_NCH * cnt * (uint32_t)sizeof(double));
#pragma unroll
            for (int c = 0; c < TCW_NCH; c++)
                bulk_g2s(sP + c * TCW_RECT_ECAP, Pt + (size_t)c * ppad + a0, cnt * (uint32_t)sizeof(double), bar);
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
    constexpr int KPS = (TCW_RECT_MAXROWS * 8 + TCW_RECT_THREADS - 1) / TCW_RECT_THREADS;
    double ps_early[KPS];
#pragma unroll
    for (int k = 0; k < KPS; k++) {
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
    const bool groupsplit = STAGED && R > 1 && !offdiag;
    float *Ft = Fmn ? Fmn + (size_t)tz * w.N_t0 * w.pitch : nullptr;

    __syncthreads();  // sE, sS visible; mbarrier init visible to all waiters
    if (groupsplit) {
        // diagonal tile: rho of a group = 1 + start index of its last row; sR = fl32(P[rho] - P[s]),
        // sG = P[rho] (global loads issued before waiting for the slice)
#pragma unroll
        for (int k = 0; k < KPS; k++) {
            const uint32_t i = threadIdx.x + k * TCW_RECT_THREADS;
            const uint32_t row = i >> 3, c = i & 7;
            if (i < ROWS * 8 && c < TCW_NCH) {
                const uint32_t rho = sS[row / R * R + R - 1] + 1;  // <= numAtoms: P has numAtoms + 1 entries
                const double prho = __ldg(Pt + (size_t)c * ppad + rho);
                sR[(((row >> 1) * 8 + c) << 1) + (row & 1)] = (float)(prho - ps_early[k]) * rect_chan_scale(c);
                if (row % R == 0) sG[(row / R) * 8 + c] = prho;
            }
        }
    }
    if (STAGED) mbar_wait(bar, phase);
    if (offdiag) {
        // off-diagonal tiles no longer need the FP64 slice itself: convert it IN PLACE to
        // {q, q} pairs, q = fl32(P_c[a0+i] - P_c[rho]) (each thread rewrites only the 8-byte slots
        // it read), after everyone has fetched P[rho] and the per-row terms fl32(P[rho] - P[s])
        double pref[TCW_NCH];
#pragma unroll
        for (int c = 0; c < TCW_NCH; c++) pref[c] = sP[c * TCW_RECT_ECAP];
#pragma unroll
        for (int k = 0; k < KPS; k++) {
            const uint32_t i = threadIdx.x + k * TCW_RECT_THREADS;
            const uint32_t row = i >> 3, c = i & 7;
            // stored as row PAIRS {row 2p, row 2p+1} per channel: a 64-bit load yields a packed operand
            if (i < ROWS * 8 && c < TCW_NCH)
                sR[(((row >> 1) * 8 + c) << 1) + (row & 1)] =
                    (float)(sP[c * TCW_RECT_ECAP] - ps_early[k]) * rect_chan_scale(c);
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < cnt; i += TCW_RECT_THREADS) {
#pragma unroll
            for (int c = 0; c < TCW_NCH; c++) {
                const float q = (float)(sP[c * TCW_RECT_ECAP + i] - pref[c]) * rect_chan_scale(c);
                sQ2[c * TCW_RECT_ECAP + i] = pack2(q, q);
            }
        }
    }
    __syncthreads();

    const uint32_t t1_step = 32u * w.dtau;
    const int n_j = (int)(d_cnt / 32);
    const int j_full = w.N_tau > d0 ? (int)min((w.N_tau - d0) / 32u, (uint32_t)n_j) : 0;  // fully valid chunks
    unsigned long long key = 0ull;
    float vmax = -1.0f;
    uint32_t degenerate = 0;
    // each warp walks G row groups of R rows: the tile's staging cost is shared.  The locate
    // pass (TRACK) visits every group slot of the tile with ALL warps, which share the chunks
    // of a group that attains the template max.
    const uint32_t n_it = TRACK ? G * TCW_RECT_WARPS : G;
#pragma unroll 1
    for (uint32_t it = 0; it < n_it; it++) {
        const uint32_t gi = TRACK ? it / TCW_RECT_WARPS : it;
        const uint32_t wsel = TRACK ? it % TCW_RECT_WARPS : warp;  // the warp that owned the group in the map pass
        const uint32_t grow = (gi * TCW_RECT_WARPS + wsel) * R;  // first row of the group, relative to m0
        // this group's max value; gmax points at the entry of the tile's first row group in the
        // table [tz][d tile][row group of the map] (independent of the row tiling of the kernel)
        uint32_t *gslot = gmax ? gmax + gi * TCW_RECT_WARPS + wsel : nullptr;
        if (m0 + grow >= w.N_t0) continue;
        if (TRACK && *gslot != top) continue;  // locate pass: only the groups that attain the template max
        float vgrp = -1.0f;  // maxF starts at -1, strict > (tcw:135-139)
        const uint32_t u_off = grow;
        const uint32_t t1_lane = t1_tile + grow * w.dt0 + lane * w.dtau;  // used by R == 1 only
        RectBest<R> best;
        float *rowp[R];
        bool rowok[R];
        uint32_t srow[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const uint32_t m = m0 + grow + r;
            best.v[r] = -1.0f;
            best.d[r] = r;
            rowok[r] = m < w.N_t0;
            const uint32_t mc = rowok[r] ? m : 0u;
            srow[r] = 0u;
            // cell (m, n = d - r) with d = d0 + lane + 32 j  ->  rowp[r][32 j]
            rowp[r] = Ft ? Ft + ((size_t)mc * w.pitch + d0 + lane) - r : nullptr;
        }
        // a group is an edge group if some (row, d) of its full chunks is not a cell of the map
        const bool edge = TRACK || (d0 < (uint32_t)(R - 1)) || (m0 + grow + R > w.N_t0);
        if (offdiag || groupsplit) {
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
            const double *sGg = sG + (grow / R) * 8;
#define RECT_ROWS(DIAG_, CHK_, STORE_, J0_, J1_)                                                                     \
    rect_rows<R, DIAG_, CHK_, STORE_, TRACK>(sQ2, sP, sGg, sE, Rs2, rowp, rowok, srow, u_off, d0, J0_, J1_, lane,   \
                                             w.N_tau, d_total, t1_lane, t1_step, a0, t0_data, numAtoms, g, Pt, ppad, \
                                             best, vgrp, degenerate)
            if (groupsplit) {
#pragma unroll
                for (int r = 0; r < R; r++) srow[r] = sS[min(grow + r, ROWS - 1)];
            }
            // map pass: one call over all chunks; locate pass: this warp's share of the chunks
            const int jstep = TRACK ? TCW_RECT_WARPS : n_j;
#pragma unroll 1
            for (int jj = TRACK ? (int)warp : 0; jj < n_j; jj += jstep) {
                const int J0 = jj, J1 = TRACK ? jj + 1 : n_j;
                if (groupsplit) {
                    if (Ft && !TRACK) RECT_ROWS(true, true, true, J0, J1);
                    else RECT_ROWS(true, true, false, J0, J1);
                } else if (edge) {
                    if (Ft && !TRACK) RECT_ROWS(false, true, true, J0, J1);
                    else RECT_ROWS(false, true, false, J0, J1);
                } else {
                    // chunks of 32 d that are valid for every lane and row run unchecked; only the
                    // chunk(s) straddling the map's right edge are bounds-checked
                    if (Ft) {
                        RECT_ROWS(false, false, true, 0, j_full);
                        if (j_full < n_j) RECT_ROWS(false, true, true, j_full, n_j);
                    } else {
                        RECT_ROWS(false, false, false, 0, j_full);
                        if (j_full < n_j) RECT_ROWS(false, true, false, j_full, n_j);
                    }
                }
            }
#undef RECT_ROWS
        } else {
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (!rowok[r] || (TRACK && warp != wsel)) continue;
                const uint32_t s_row = sS[grow + r];
                const RectRowResult rr =
                    (Ft && !TRACK)
                        ? rect_row_fp64<R, STAGED, true>(sP, sE, Pt, ppad, s_row, rowp[r], r, u_off, d0, n_j, lane,
                                                         w.N_tau, d_total, t1_lane, t1_step, a0, t0_data, numAtoms, g)
                        : rect_row_fp64<R, STAGED, false>(sP, sE, Pt, ppad, s_row, rowp[r], r, u_off, d0, n_j, lane,
                                                          w.N_tau, d_total, t1_lane, t1_step, a0, t0_data, numAtoms, g);
                best.v[r] = rr.best;
                best.d[r] = rr.best_d;
                vgrp = fmaxf(vgrp, rr.best);
                degenerate |= rr.degenerate;
            }
        }
        if (!TRACK) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) vgrp = fmaxf(vgrp, __shfl_xor_sync(0xffffffffu, vgrp, o));
            if (gslot && lane == 0) *gslot = vgrp > -1.0f ? float_orderable(vgrp) : 0u;
            vmax = fmaxf(vmax, vgrp);
        } else {
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (best.v[r] > -1.0f) {
                    const unsigned long long k =
                        pack_key(best.v[r], (m0 + grow + r) * w.N_tau + (best.d[r] - r));
                    key = k > key ? k : key;
                }
            }
        }
    }
    if (!TRACK && vmax > -1.0f) key = pack_key(vmax, 0xFFFFFFFFu);  // index part 0: completed later
    if (degenerate) atomicOr(&flags[t], TCW_FLAG_DEGENERATE);
    return block_max_key<TCW_RECT_WARPS>(key, red);
}

// Completes the argmax of template t without a stored F_mn: finds the row groups whose maximum
// equals the template maximum (normally one) and re-evaluates their tiles with first-occurrence
// tracking (identical arithmetic, nothing stored).  Run by one whole CTA; (gx, gy) = tile grid.
#define TCW_RECT_HITWORDS 256
template <int R, bool STAGED>
__device__ __forceinline__ void rect_locate(const double *__restrict__ P, uint32_t ppad,
                                            const TplMeta *__restrict__ meta, int t, int tz, const MapWindow &w,
                                            const IndexGeom &g, uint32_t DD, uint32_t DT, uint32_t G, uint32_t gx,
                                            uint32_t gy, unsigned long long *__restrict__ maxkey,
                                            uint32_t *__restrict__ groupmax, uint32_t *__restrict__ flags,
                                            unsigned char *smem, uint64_t *bar, uint32_t phase,
                                            unsigned long long *red, uint32_t *hit) {
    const uint32_t top = (uint32_t)(*(volatile unsigned long long *)&maxkey[t] >> 32);
    if (top == 0u) return;
    // table layout: [tz][d tile bx][row group of the map], n_grp = ceil(N_t0 / R) groups
    const uint32_t n_grp = (w.N_t0 + R - 1) / R;
    const uint32_t gpt = TCW_RECT_WARPS * G;  // row groups per tile of THIS pass
    const uint32_t n_tiles = gx * gy;
    uint32_t *gm = groupmax + (size_t)tz * gx * n_grp;
    for (uint32_t base = 0; base < n_tiles; base += TCW_RECT_HITWORDS * 32) {
        const uint32_t n_here = min(n_tiles - base, (uint32_t)(TCW_RECT_HITWORDS * 32));
        for (uint32_t i = threadIdx.x; i < TCW_RECT_HITWORDS; i += TCW_RECT_THREADS) hit[i] = 0u;
        __syncthreads();
        // tiles are numbered by * gx + bx; mark those holding a group that attains the template max
        for (uint32_t e = threadIdx.x; e < gx * n_grp; e += TCW_RECT_THREADS) {
            const uint32_t bx = e / n_grp, grp = e - bx * n_grp;
            const uint32_t tile = (grp / gpt) * gx + bx;
            if (tile >= base && tile < base + n_here && __ldcg(gm + e) == top)
                atomicOr(&hit[(tile - base) >> 5], 1u << ((tile - base) & 31));
        }
        __syncthreads();
        for (uint32_t wd = 0; wd < (n_here + 31) / 32; wd++) {
            uint32_t bits = hit[wd];  // uniform across the CTA
            while (bits) {
                const uint32_t tile = base + wd * 32 + (uint32_t)(__ffs(bits) - 1);
                bits &= bits - 1;
                const uint32_t bx = tile % gx, by = tile / gx;
                const unsigned long long key =
                    rect_tile<R, STAGED, true>(P, ppad, meta, t, tz, bx, by, w, g, DD, DT, G, nullptr, flags,
                                               gm + (size_t)bx * n_grp + (size_t)by * gpt, top, smem, bar, phase & 1u,
                                               phase == 0u, red);
                phase++;
                if (threadIdx.x == 0 && key != 0ull) atomicMax(&maxkey[t], key);
                __syncthreads();  // red[] and the staged slice are reused by the next tile
            }
        }
        __syncthreads();
    }
}

// grid: x = d tiles, y = row tiles, z = template in sub-batch.  Publishes the max VALUE per
// template (atomicMax on the packed key, index part 0) and, when the argmax has to be completed
// by tcw_rect_locate_kernel (groupmax != nullptr: no lnBtSG pass follows), per row group:
// one entry per (d tile, row group of the map).
// Natural dispatch order (x fastest): the cheap head-strip tiles are interleaved with the regular
// ones, which keeps the CTAs sharing an SM out of phase (one stages while the other computes).
// Measured: running all regular tiles first and the head tiles last is 7 % slower; so is running
// the locate step inside this kernel (last CTA of a template) instead of a second small launch.
template <int R, bool STAGED>
__global__ void __launch_bounds__(TCW_RECT_THREADS, TCW_RECT_MINB)
tcw_rect_map_kernel(const double *__restrict__ P, uint32_t ppad, const TplMeta *__restrict__ meta,
                    int t_base, MapWindow w, IndexGeom g, uint32_t DD, uint32_t DT, uint32_t G,
                    uint32_t gx_total, float *__restrict__ Fmn, unsigned long long *__restrict__ maxkey,
                    uint32_t *__restrict__ groupmax, uint32_t *__restrict__ flags) {
    extern __shared__ __align__(16) unsigned char tcw_rect_smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ unsigned long long red[TCW_RECT_WARPS];
    const int tz = blockIdx.z, t = t_base + tz;
    // group-max table [tz][d tile][row group of the map]; gx_total = d tiles of the whole map (this
    // launch may cover only the head strip, gridDim.x == 1, when the regular tiles run in the
    // persistent kernel of tcw_rect_p.cuh)
    const uint32_t n_grp = (w.N_t0 + R - 1) / R;
    uint32_t *gmax = groupmax ? groupmax + ((size_t)tz * gx_total + blockIdx.x) * n_grp +
                                    (size_t)blockIdx.y * (TCW_RECT_WARPS * G)
                              : nullptr;
    const unsigned long long key = rect_tile<R, STAGED, false>(P, ppad, meta, t, tz, blockIdx.x, blockIdx.y, w, g, DD,
                                                               DT, G, Fmn, flags, gmax, 0u, tcw_rect_smem, &bar, 0u,
                                                               true, red);
    if (threadIdx.x == 0 && key != 0ull) atomicMax(&maxkey[t], key);
}

// One CTA per template of the sub-batch, launched after the map kernel (stream order makes the
// template maxima and the group table final): completes the argmax, see rect_locate.
template <int R, bool STAGED>
__global__ void __launch_bounds__(TCW_RECT_THREADS, TCW_RECT_MINB)
tcw_rect_locate_kernel(const double *__restrict__ P, uint32_t ppad, const TplMeta *__restrict__ meta,
                       int t_base, MapWindow w, IndexGeom g, uint32_t DD, uint32_t DT, uint32_t G, uint32_t gx,
                       uint32_t gy, unsigned long long *__restrict__ maxkey, uint32_t *__restrict__ groupmax,
                       uint32_t *__restrict__ flags) {
    extern __shared__ __align__(16) unsigned char tcw_rect_smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ unsigned long long red[TCW_RECT_WARPS];
    __shared__ uint32_t hit[TCW_RECT_HITWORDS];  // bitmap of tiles holding a group with the top value
    const int tz = blockIdx.x, t = t_base + tz;
    rect_locate<R, STAGED>(P, ppad, meta, t, tz, w, g, DD, DT, G, gx, gy, maxkey, groupmax, flags, tcw_rect_smem, &bar,
                           0u, red, hit);
}
