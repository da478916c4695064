// tcw_rect.cuh -- tiled rectangular-window map kernel (the HBM/output-bound path).
//
// Replaces pyCUDAkernels/cudaTransientFstatRectWindow.cu (one thread per t0 row, serial loop
// over tau with float32 running sums, uncoalesced stores).  Here every (t0,tau) cell is an
// O(1) difference of FP64 prefix sums,
//     S_c[m,n] = P_c[i_t1(m,n)+1] - P_c[i_t0(m)],
// the 7 differences are rounded to FP32 once and go through the guarded F-stat formula.
//
// Work decomposition ("skewed" tiles).  With dt0 == dtau (the canonical grids of the
// reference's tests/examples) the cells (m, n) and (m+1, n-1) share the window END time
// t1 = t0_m + tau_n, hence the same end index.  A thread therefore owns a group of R
// consecutive rows and walks d = n + r: one end-prefix fetch (7 x FP64 from shared memory)
// feeds R cells, and each row's start prefix lives in registers.  R = 1 is the plain mapping
// for dt0 != dtau.  A warp's lanes cover 32 consecutive d, so F_mn stores are coalesced rows
// (the reference kernel stores with stride N_tau).
//
// Staging.  A tile's end indices form one contiguous range (monotone under the host-side
// no-wrap certificate); that slice of the 7 prefix channels is brought into shared memory
// with 1-D TMA bulk copies (cp.async.bulk + mbarrier -> SASS UBLKCP) while the CTA computes
// the tile's end-index table (exact uint32 formulas, once per distinct end time instead of
// once per cell).  STAGED = false reads the prefixes straight from global/L2 (very coarse
// dtau, where a tile's range would not fit).
//
// Fused epilogue: per-row running (max F, first d), combined per CTA and published with one
// 64-bit atomicMax per CTA; F_mn is stored only if the caller (or the lnBtSG pass) needs it.
// Interior tiles (no map edge, no diagonal) run a branch-free body.
#pragma once
#include "tcw_common.cuh"
#include "tcw_prep.cuh"

#define TCW_RECT_THREADS 256
#define TCW_RECT_WARPS (TCW_RECT_THREADS / 32)
#define TCW_RECT_DT 512     // d values per tile (16 per lane)
#define TCW_RECT_ECAP 1024  // staged end-prefix entries per channel (even)
#define TCW_RECT_UCAP (TCW_RECT_DT + 32)  // end-index table entries (R*WARPS <= 32)
#define TCW_RECT_SMEM (TCW_NCH * TCW_RECT_ECAP * 8 + TCW_RECT_UCAP * 4)

// One warp's share of a tile: R rows x DT values of d.
//   CHECKED = false: interior tile, every (row, d) is a valid cell and none is degenerate.
template <int R, bool STAGED, bool CHECKED, bool STORE>
__device__ __forceinline__ void rect_tile_rows(
    const double *__restrict__ sP, const uint32_t *__restrict__ sE, const double *__restrict__ Pt,
    uint32_t ppad, const double (&Ps)[R][TCW_NCH], const uint32_t (&s_idx)[R], float *const (&rowp)[R],
    const bool (&rowok)[R], uint32_t u_off, uint32_t d0, uint32_t lane, uint32_t N_tau, uint32_t d_total,
    uint32_t t1_lane, uint32_t t1_step, uint32_t a0, uint32_t t0_data, uint32_t numAtoms, const IndexGeom g,
    float (&best)[R], uint32_t (&best_d)[R], bool &degenerate) {
#pragma unroll 4
    for (int j = 0; j < TCW_RECT_DT / 32; j++) {
        const uint32_t d = d0 + lane + 32u * j;
        if (CHECKED && d >= d_total) break;
        uint32_t idx;  // index of P[e+1] relative to the staged slice (or absolute if !STAGED)
        if (R > 1) {
            idx = sE[u_off + lane + 32u * j];
        } else {
            const uint32_t e = index_t1(t1_lane + (uint32_t)j * t1_step, t0_data, numAtoms, g);
            idx = e + 1 - a0;
        }
        double E[TCW_NCH];
        if (STAGED) {
#pragma unroll
            for (int c = 0; c < TCW_NCH; c++) E[c] = sP[c * TCW_RECT_ECAP + idx];
        } else {
#pragma unroll
            for (int c = 0; c < TCW_NCH; c++) E[c] = __ldg(Pt + (size_t)c * ppad + idx + a0);
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            float S[TCW_NCH];
#pragma unroll
            for (int c = 0; c < TCW_NCH; c++) S[c] = (float)(E[c] - Ps[r][c]);
            const float F = fstat_fast(S[0], S[1], S[2], S[3], S[4], S[5], S[6]);
            if (CHECKED) {
                const uint32_t n = d - r;  // wraps for d < r -> fails the n < N_tau test
                if (rowok[r] && n < N_tau) {
                    if (STORE) rowp[r][32 * j] = F;
                    if (F > best[r]) {
                        best[r] = F;
                        best_d[r] = d;
                    }
                    if (idx + a0 - 1 == s_idx[r]) degenerate = true;  // i_t1 == i_t0
                }
            } else {
                if (STORE) rowp[r][32 * j] = F;
                if (F > best[r]) {
                    best[r] = F;
                    best_d[r] = d;
                }
            }
        }
    }
}

template <int R, bool STAGED>
__global__ void __launch_bounds__(TCW_RECT_THREADS, 2)
tcw_rect_map_kernel(const double *__restrict__ P, uint32_t ppad, const TplMeta *__restrict__ meta,
                    int t_base, MapWindow w, IndexGeom g, float *__restrict__ Fmn,
                    unsigned long long *__restrict__ maxkey, uint32_t *__restrict__ flags) {
    extern __shared__ __align__(16) unsigned char tcw_rect_smem[];
    double *sP = reinterpret_cast<double *>(tcw_rect_smem);                                  // [7][ECAP]
    uint32_t *sE = reinterpret_cast<uint32_t *>(tcw_rect_smem + TCW_NCH * TCW_RECT_ECAP * 8);  // [UCAP]
    __shared__ __align__(8) uint64_t bar;
    __shared__ unsigned long long red[TCW_RECT_WARPS];

    const int tz = blockIdx.z;
    const int t = t_base + tz;
    const uint32_t numAtoms = meta[t].numAtoms;
    const uint32_t t0_data = meta[t].t0_data;
    const double *Pt = P + (size_t)t * TCW_NCH * ppad;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    const uint32_t n_groups = (w.N_t0 + R - 1) / R;
    const uint32_t d_total = w.N_tau + R - 1;
    const uint32_t g0 = blockIdx.y * TCW_RECT_WARPS;
    const uint32_t d0 = blockIdx.x * TCW_RECT_DT;
    const uint32_t g_last = min(g0 + TCW_RECT_WARPS, n_groups) - 1;
    const uint32_t d_last = min(d0 + TCW_RECT_DT, d_total) - 1;

    // end time of (group, d): rows of a group differ by dt0 == dtau, absorbed into d
    const uint32_t t1_tile = w.t0 + w.tau + g0 * R * w.dt0 + d0 * w.dtau;
    const uint32_t e_lo = index_t1(t1_tile, t0_data, numAtoms, g);
    const uint32_t a0 = STAGED ? ((e_lo + 1) & ~1u) : 0u;
    if (STAGED) {
        const uint32_t e_hi = index_t1(w.t0 + w.tau + g_last * R * w.dt0 + d_last * w.dtau, t0_data, numAtoms, g);
        const uint32_t cnt = (e_hi + 1 - a0 + 1 + 1) & ~1u;  // even count, <= ECAP (host-checked)
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            mbar_fence_init();
            mbar_arrive_expect_tx(&bar, TCW_NCH * cnt * (uint32_t)sizeof(double));
#pragma unroll
            for (int c = 0; c < TCW_NCH; c++)
                bulk_g2s(sP + c * TCW_RECT_ECAP, Pt + (size_t)c * ppad + a0, cnt * (uint32_t)sizeof(double), &bar);
        }
    }
    // end-index table over u = (grp - g0)*R + (d - d0): t1 = t1_tile + u*dtau  (dt0 == dtau)
    if (R > 1) {
        for (uint32_t u = threadIdx.x; u < TCW_RECT_UCAP; u += TCW_RECT_THREADS)
            sE[u] = STAGED ? min(index_t1(t1_tile + u * w.dtau, t0_data, numAtoms, g) + 1 - a0,
                                 (uint32_t)(TCW_RECT_ECAP - 1))  // overhang entries stay in bounds
                           : index_t1(t1_tile + u * w.dtau, t0_data, numAtoms, g) + 1;
    }
    const uint32_t m_hi = min((g_last + 1) * R, w.N_t0) - 1;
    const uint32_t s_hi = index_t0(w.t0 + m_hi * w.dt0, t0_data, numAtoms, g);
    // degenerate (single-atom) cells can only occur in tiles touching the diagonal
    const bool interior = (e_lo > s_hi) && (d0 >= (uint32_t)(R - 1)) && (d0 + TCW_RECT_DT <= w.N_tau) &&
                          ((g0 + TCW_RECT_WARPS) * R <= w.N_t0);

    // this warp's row group: start prefixes in registers
    const uint32_t grp = g0 + warp;
    const size_t cells = (size_t)w.N_t0 * w.N_tau;
    float *Ft = Fmn ? Fmn + (size_t)tz * cells : nullptr;
    double Ps[R][TCW_NCH];
    uint32_t s_idx[R];
    float best[R];
    uint32_t best_d[R];
    float *rowp[R];
    bool rowok[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const uint32_t m = grp * R + r;
        best[r] = -1.0f;  // maxF starts at -1, strict > (tcw:135-139)
        best_d[r] = r;
        rowok[r] = m < w.N_t0;
        const uint32_t mc = rowok[r] ? m : 0u;
        s_idx[r] = index_t0(w.t0 + mc * w.dt0, t0_data, numAtoms, g);
#pragma unroll
        for (int c = 0; c < TCW_NCH; c++) Ps[r][c] = __ldg(Pt + (size_t)c * ppad + s_idx[r]);
        // cell (m, n = d - r) with d = d0 + lane + 32 j  ->  rowp[r][32 j]
        rowp[r] = Ft ? Ft + ((size_t)mc * w.N_tau + d0 + lane) - r : nullptr;
    }
    __syncthreads();  // sE visible; mbarrier init visible to all waiters
    if (STAGED) mbar_wait(&bar, 0);

    bool degenerate = false;
    const uint32_t u_off = warp * R;
    const uint32_t t1_lane = t1_tile + (warp * R * w.dt0) + lane * w.dtau;  // used by R == 1 only
    const uint32_t t1_step = 32u * w.dtau;
    if (grp < n_groups) {
        if (interior) {
            if (Ft)
                rect_tile_rows<R, STAGED, false, true>(sP, sE, Pt, ppad, Ps, s_idx, rowp, rowok, u_off, d0, lane,
                                                       w.N_tau, d_total, t1_lane, t1_step, a0, t0_data, numAtoms,
                                                       g, best, best_d, degenerate);
            else
                rect_tile_rows<R, STAGED, false, false>(sP, sE, Pt, ppad, Ps, s_idx, rowp, rowok, u_off, d0, lane,
                                                        w.N_tau, d_total, t1_lane, t1_step, a0, t0_data, numAtoms,
                                                        g, best, best_d, degenerate);
        } else {
            if (Ft)
                rect_tile_rows<R, STAGED, true, true>(sP, sE, Pt, ppad, Ps, s_idx, rowp, rowok, u_off, d0, lane,
                                                      w.N_tau, d_total, t1_lane, t1_step, a0, t0_data, numAtoms,
                                                      g, best, best_d, degenerate);
            else
                rect_tile_rows<R, STAGED, true, false>(sP, sE, Pt, ppad, Ps, s_idx, rowp, rowok, u_off, d0, lane,
                                                       w.N_tau, d_total, t1_lane, t1_step, a0, t0_data, numAtoms,
                                                       g, best, best_d, degenerate);
        }
    }
    if (degenerate) atomicOr(&flags[t], TCW_FLAG_DEGENERATE);

    unsigned long long key = 0ull;
#pragma unroll
    for (int r = 0; r < R; r++) {
        if (best[r] > -1.0f) {
            const uint32_t flat = (grp * R + r) * w.N_tau + (best_d[r] - r);
            const unsigned long long k = pack_key(best[r], flat);
            key = k > key ? k : key;
        }
    }
    block_atomic_max_key<TCW_RECT_WARPS>(key, &maxkey[t], red);
}
