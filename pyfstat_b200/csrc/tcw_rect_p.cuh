// tcw_rect_p.cuh -- persistent, warp-specialised rectangular-window map kernel.
//
// Why (ncu, round 1, tcw_rect_map_kernel at 64 x 60 d: profiles/r01_rect_v9_ncu_summary.txt): the hot
// loop of tcw_rect.cuh runs at ~80 % of the FP32 pipe, but the kernel as a whole at 59 %: every
// CTA is one tile, and before its loop starts it waits for the bulk copies of the end-prefix
// slice, builds the index tables and converts the slice to FP32 {q,q} pairs -- ~16 % of the warp
// samples sit in that prologue and its barriers, during which the SM runs on the other CTA's 8
// warps alone; the narrow head-strip tiles (the diagonal) are almost all prologue.  Here ONE CTA
// per SM lives for the whole launch and splits into
//   * 2 producer warps: fetch the next tile (global atomic counter), issue the 1-D TMA bulk copies
//     of its FP64 prefix slice (cp.async.bulk -> UBLKCP), build the end-index table and the
//     per-row terms, convert the slice in place -- into the OTHER of two shared-memory tile
//     buffers, and
//   * 14 consumer warps that never leave the F-stat loops (rect_rows of tcw_rect.cuh, unchanged
//     arithmetic).  The 32 row groups (4 rows each) of a 128-row tile are handed out through a
//     shared-memory counter: warp schedulers hosting a producer warp run 3 consumer warps, the
//     others 4, so equal static shares left the faster warps waiting ~10 % of the time (ncu
//     r02_rectp_v1: long-scoreboard samples on the `ready` barrier),
// handing tiles over through mbarriers (full: TMA bytes landed; ready: tables + {q,q} built;
// empty: every row group of the buffer's tile is done and every consumer warp has left it).
//
// Tile kinds.  Regular tiles (d >= DD) are off-diagonal by the host's certificate: one split
// point per tile, {q,q} table, packed-FP32 loop.  Head tiles (d < DD, holding the diagonal) keep
// the FP64 slice and use the per-group split point of tcw_rect.cuh (DIAG path).
//
// Conditioning certificate (NOGUARD).  Every window of a regular tile contains the tile's core
// range [s_hi, e_lo] and lies inside its hull [s_lo, e_hi]; the antenna-pattern matrices are sums
// of positive semi-definite per-atom terms, so M_core <= M <= M_hull in the Loewner order and
//     cond(M) <= lambda_max(M_hull) / lambda_min(M_core).
// A producer thread evaluates that bound once per tile from the FP64 prefixes; when it is below
// 2500 (the reference's cut is 1e4, Rect.cu:104-107) no cell of the tile can take the F = 2
// fallback and the consumers skip the per-cell margin (3 of 22 packed FP32 instructions per cell
// pair and the min).  Otherwise the tile runs the guarded loop of tcw_rect.cuh.
//
// Used by the host only under the same certificate as the tiled kernel plus: R = 4 (dt0 == dtau),
// staged slices, and every tile at d >= DD off-diagonal for every template of the launch
// (tcw_b200.cu: plan_rect_p).  Everything else stays in tcw_rect_map_kernel.
#pragma once
#include "tcw_rect.cuh"

#ifndef TCW_RECTP_CWARPS
// measured (60 d, T = 64, F_mn stored, map stage): 14 + 2 warps 0.614 ms, 15 + 1 warps 0.638 ms (one
// producer warp cannot keep up on the cheap head tiles); one-tile-per-CTA kernel 0.635 ms
#define TCW_RECTP_CWARPS 14
#define TCW_RECTP_PWARPS 2
#endif
#define TCW_RECTP_THREADS ((TCW_RECTP_CWARPS + TCW_RECTP_PWARPS) * 32)
#define TCW_RECTP_PTHREADS (TCW_RECTP_PWARPS * 32)
#define TCW_RECTP_ROWS 128                      // rows per tile
#define TCW_RECTP_GROUPS (TCW_RECTP_ROWS / 4)   // row groups per tile, handed out dynamically
#define TCW_RECTP_UCAP (TCW_RECT_DT + TCW_RECTP_ROWS)
#define TCW_RECTP_BUF_Q (TCW_NCH * TCW_RECT_ECAP * 8)
#define TCW_RECTP_BUF_E (TCW_RECTP_UCAP * 4)
#define TCW_RECTP_BUF_R (TCW_RECTP_ROWS * 8 * 4)
#define TCW_RECTP_BUF_S (TCW_RECTP_ROWS * 4)
#define TCW_RECTP_BUF_G (TCW_RECTP_GROUPS * 8 * 8)
#define TCW_RECTP_BUF (TCW_RECTP_BUF_Q + TCW_RECTP_BUF_E + TCW_RECTP_BUF_R + TCW_RECTP_BUF_S + TCW_RECTP_BUF_G)
#define TCW_RECTP_SMEM (2 * TCW_RECTP_BUF)
#define TCW_RECTP_COND_MAX 2500.0  // certificate bound (reference cut: 1e4)

struct RectPDesc {
    uint32_t tile;      // 0xFFFFFFFF: no more tiles
    uint32_t noguard;   // bit g: the conditioning certificate holds for row group g (regular tiles)
    uint32_t a0;        // first staged prefix index
    uint32_t next;      // next row group to hand out
    uint32_t numAtoms;  // of the tile's template (consumers need no global load per tile)
    uint32_t t0_data;
    uint32_t pad[2];
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// grid: min(#SM, tiles) CTAs.  Tiles are numbered (tz * n_gy + by) * (1 + n_reg) + bx; bx = 0 is
// the head strip [0, DD), bx >= 1 the regular tile [DD + (bx-1) DT, +DT).
__global__ void __launch_bounds__(TCW_RECTP_THREADS, 1)
tcw_rect_map_p_kernel(const double *__restrict__ P, uint32_t ppad, const TplMeta *__restrict__ meta, int t_base,
                      MapWindow w, IndexGeom g, uint32_t DD, uint32_t DT, uint32_t n_reg, uint32_t n_gy,
                      uint32_t n_tiles, uint32_t *__restrict__ tile_counter, float *__restrict__ Fmn,
                      unsigned long long *__restrict__ maxkey, uint32_t *__restrict__ groupmax,
                      uint32_t *__restrict__ flags) {
    constexpr int R = 4;
    extern __shared__ __align__(16) unsigned char rectp_smem[];
    __shared__ __align__(8) uint64_t bar_full[2], bar_ready[2], bar_empty[2];
    __shared__ RectPDesc desc[2];

    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int b = 0; b < 2; b++) {
            mbar_init(&bar_full[b], 1);
            mbar_init(&bar_ready[b], TCW_RECTP_PTHREADS);
            mbar_init(&bar_empty[b], TCW_RECTP_GROUPS + TCW_RECTP_CWARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const uint32_t d_total = w.N_tau + R - 1;
    const uint32_t n_grp = (w.N_t0 + R - 1) / R;
    const uint32_t gx_total = 1 + n_reg;

    if (warp >= TCW_RECTP_CWARPS) {
        // ================================ producers ================================
        const uint32_t pt = threadIdx.x - TCW_RECTP_CWARPS * 32;
        for (uint32_t k = 0;; k++) {
            const uint32_t b = k & 1u;
            unsigned char *buf = rectp_smem + (size_t)b * TCW_RECTP_BUF;
            double *sP = reinterpret_cast<double *>(buf);
            f32x2 *sQ2 = reinterpret_cast<f32x2 *>(buf);
            uint32_t *sE = reinterpret_cast<uint32_t *>(buf + TCW_RECTP_BUF_Q);
            float *sR = reinterpret_cast<float *>(buf + TCW_RECTP_BUF_Q + TCW_RECTP_BUF_E);
            uint32_t *sS = reinterpret_cast<uint32_t *>(buf + TCW_RECTP_BUF_Q + TCW_RECTP_BUF_E + TCW_RECTP_BUF_R);
            double *sG = reinterpret_cast<double *>(buf + TCW_RECTP_BUF_Q + TCW_RECTP_BUF_E + TCW_RECTP_BUF_R +
                                                    TCW_RECTP_BUF_S);
            if (k >= 2) mbar_wait(&bar_empty[b], ((k >> 1) - 1) & 1u);  // consumers are done with this buffer
            if (pt == 0) desc[b].tile = atomicAdd(tile_counter, 1u);
            named_bar_sync(1, TCW_RECTP_PTHREADS);
            const uint32_t tile = desc[b].tile;
            if (tile >= n_tiles) {
                named_bar_sync(1, TCW_RECTP_PTHREADS);  // everyone has read the id before it is overwritten
                if (pt == 0) desc[b].tile = 0xFFFFFFFFu;
                mbar_arrive(&bar_ready[b]);
                break;
            }
            const uint32_t bx = tile % gx_total;
            const uint32_t by = (tile / gx_total) % n_gy;
            const uint32_t tz = tile / (gx_total * n_gy);
            const bool head = bx == 0;
            const int t = t_base + (int)tz;
            const uint32_t numAtoms = meta[t].numAtoms, t0_data = meta[t].t0_data;
            const double *Pt = P + (size_t)t * TCW_NCH * ppad;
            const uint32_t m0 = by * TCW_RECTP_ROWS;
            const uint32_t m_last = min(m0 + TCW_RECTP_ROWS, w.N_t0) - 1;
            const uint32_t d0 = head ? 0u : DD + (bx - 1) * DT;
            const uint32_t d_cnt = head ? DD : DT;
            const uint32_t d_last = min(d0 + d_cnt, d_total) - 1;
            const uint32_t t1_tile = w.t0 + w.tau + m0 * w.dt0 + d0 * w.dtau;
            const uint32_t e_lo = index_t1(t1_tile, t0_data, numAtoms, g);
            const uint32_t a0 = (e_lo + 1) & ~1u;
            const uint32_t e_hi =
                index_t1(t1_tile + ((m_last - m0) / R * R) * w.dt0 + (d_last - d0) * w.dtau, t0_data, numAtoms, g);
            const uint32_t cnt = min((e_hi + 1 - a0 + 1 + 1) & ~1u, (uint32_t)TCW_RECT_ECAP);
            if (pt == 0) {
                mbar_arrive_expect_tx(&bar_full[b], TCW_NCH * cnt * (uint32_t)sizeof(double));
#pragma unroll
                for (int c = 0; c < TCW_NCH; c++)
                    bulk_g2s(sP + c * TCW_RECT_ECAP, Pt + (size_t)c * ppad + a0, cnt * (uint32_t)sizeof(double),
                             &bar_full[b]);
                desc[b].a0 = a0;
                desc[b].next = 0u;
                desc[b].numAtoms = numAtoms;
                desc[b].t0_data = t0_data;
                if (head) desc[b].noguard = 0u;
            }
            // start index and start prefixes P_c[i_t0(row)] of the tile's rows, issued while the bulk copies fly
            constexpr int KPS = (TCW_RECTP_ROWS * 8 + TCW_RECTP_PTHREADS - 1) / TCW_RECTP_PTHREADS;
            double ps[KPS];
#pragma unroll
            for (int q = 0; q < KPS; q++) {
                const uint32_t i = pt + q * TCW_RECTP_PTHREADS;
                const uint32_t row = i >> 3, c = i & 7;
                ps[q] = 0.0;
                if (i < TCW_RECTP_ROWS * 8) {
                    const uint32_t m = min(m0 + row, w.N_t0 - 1);
                    const uint32_t s = index_t0(w.t0 + m * w.dt0, t0_data, numAtoms, g);
                    if (c < TCW_NCH) ps[q] = __ldg(Pt + (size_t)c * ppad + s);
                    else sS[row] = s;
                }
            }
            // end-index table: sE[u], u = (row - m0)/R*R + (d - d0), relative to the staged slice
            for (uint32_t u = pt; u < TCW_RECTP_ROWS + d_cnt; u += TCW_RECTP_PTHREADS)
                sE[u] = min(index_t1(t1_tile + u * w.dtau, t0_data, numAtoms, g) + 1 - a0, (uint32_t)(TCW_RECT_ECAP - 1));
            if (head) {
                // diagonal strip: split point per row group, rho = 1 + start index of the group's last row;
                // sR = fl32(P[rho] - P[s]), sG = P[rho].  The FP64 slice stays as it is.
                named_bar_sync(1, TCW_RECTP_PTHREADS);  // sS complete
#pragma unroll
                for (int q = 0; q < KPS; q++) {
                    const uint32_t i = pt + q * TCW_RECTP_PTHREADS;
                    const uint32_t row = i >> 3, c = i & 7;
                    if (i < TCW_RECTP_ROWS * 8 && c < TCW_NCH) {
                        const uint32_t rho = sS[row / R * R + R - 1] + 1;  // <= numAtoms: P has numAtoms + 1 entries
                        const double prho = __ldg(Pt + (size_t)c * ppad + rho);
                        sR[(((row >> 1) * 8 + c) << 1) + (row & 1)] = (float)(prho - ps[q]) * rect_chan_scale(c);
                        if (row % R == 0) sG[(row / R) * 8 + c] = prho;
                    }
                }
                mbar_wait(&bar_full[b], (k >> 1) & 1u);
            } else {
                // conditioning certificate per row group (lane g <-> group g; 12 FP64 loads each): every
                // window of the group contains [s_hi, e_lo] and lies inside [s_lo, e_hi]
                if (pt < 32) {  // the first producer warp: lane g <-> row group g
                    static_assert(TCW_RECTP_GROUPS == 32, "one lane per row group");
                    const uint32_t r0 = m0 + pt * R;
                    bool ok = false;
                    if (r0 < w.N_t0) {
                        const uint32_t r3 = min(r0 + R - 1, w.N_t0 - 1);
                        const uint32_t s_lo = index_t0(w.t0 + r0 * w.dt0, t0_data, numAtoms, g);
                        const uint32_t s_hi = index_t0(w.t0 + r3 * w.dt0, t0_data, numAtoms, g);
                        const uint32_t ge_lo = index_t1(t1_tile + (pt * R) * w.dtau, t0_data, numAtoms, g);
                        const uint32_t ge_hi = index_t1(t1_tile + (pt * R + d_last - d0) * w.dtau, t0_data, numAtoms, g);
                        double core[3], hull[3];
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            const double *pc = Pt + (size_t)c * ppad;
                            core[c] = __ldg(pc + ge_lo + 1) - __ldg(pc + s_hi);
                            hull[c] = __ldg(pc + ge_hi + 1) - __ldg(pc + s_lo);
                        }
                        const double sc = core[0] + core[1];
                        const double dc = sqrt((core[0] - core[1]) * (core[0] - core[1]) + 4.0 * core[2] * core[2]);
                        const double sh = hull[0] + hull[1];
                        const double dh = sqrt((hull[0] - hull[1]) * (hull[0] - hull[1]) + 4.0 * hull[2] * hull[2]);
                        const double lmin = sc - dc, lmax = sh + dh;
                        ok = ge_lo >= s_hi && lmin > 0.0 && lmax < TCW_RECTP_COND_MAX * lmin;
                    }
                    const uint32_t mask = __ballot_sync(0xffffffffu, ok);
                    if (pt == 0) desc[b].noguard = mask;
                }
                mbar_wait(&bar_full[b], (k >> 1) & 1u);
                // split point rho = a0: per-row terms fl32(P[rho] - P[s]) and, in place over the FP64
                // slice, {q, q} with q = fl32(P[a0 + i] - P[rho])
                double pref[TCW_NCH];
#pragma unroll
                for (int c = 0; c < TCW_NCH; c++) pref[c] = sP[c * TCW_RECT_ECAP];
#pragma unroll
                for (int q = 0; q < KPS; q++) {
                    const uint32_t i = pt + q * TCW_RECTP_PTHREADS;
                    const uint32_t row = i >> 3, c = i & 7;
                    if (i < TCW_RECTP_ROWS * 8 && c < TCW_NCH)
                        sR[(((row >> 1) * 8 + c) << 1) + (row & 1)] =
                            (float)(sP[c * TCW_RECT_ECAP] - ps[q]) * rect_chan_scale(c);
                }
                named_bar_sync(1, TCW_RECTP_PTHREADS);  // entry 0 of every channel has been read by all
                for (uint32_t i = pt; i < cnt; i += TCW_RECTP_PTHREADS) {
#pragma unroll
                    for (int c = 0; c < TCW_NCH; c++) {
                        const float q = (float)(sP[c * TCW_RECT_ECAP + i] - pref[c]) * rect_chan_scale(c);
                        sQ2[c * TCW_RECT_ECAP + i] = pack2(q, q);
                    }
                }
            }
            mbar_arrive(&bar_ready[b]);  // release: tables, slice and descriptor are visible to the waiters
        }
        return;
    }

    // ================================ consumers ================================
    for (uint32_t k = 0;; k++) {
        const uint32_t b = k & 1u;
        unsigned char *buf = rectp_smem + (size_t)b * TCW_RECTP_BUF;
        const double *sP = reinterpret_cast<const double *>(buf);
        const f32x2 *sQ2 = reinterpret_cast<const f32x2 *>(buf);
        const uint32_t *sE = reinterpret_cast<const uint32_t *>(buf + TCW_RECTP_BUF_Q);
        const float *sR = reinterpret_cast<const float *>(buf + TCW_RECTP_BUF_Q + TCW_RECTP_BUF_E);
        const uint32_t *sS = reinterpret_cast<const uint32_t *>(buf + TCW_RECTP_BUF_Q + TCW_RECTP_BUF_E + TCW_RECTP_BUF_R);
        const double *sG = reinterpret_cast<const double *>(buf + TCW_RECTP_BUF_Q + TCW_RECTP_BUF_E + TCW_RECTP_BUF_R +
                                                            TCW_RECTP_BUF_S);
        mbar_wait(&bar_ready[b], (k >> 1) & 1u);
        const uint32_t tile = desc[b].tile;
        if (tile == 0xFFFFFFFFu) break;
        const uint32_t noguard_mask = desc[b].noguard;
        const uint32_t a0 = desc[b].a0;
        const uint32_t bx = tile % gx_total;
        const uint32_t by = (tile / gx_total) % n_gy;
        const uint32_t tz = tile / (gx_total * n_gy);
        const bool head = bx == 0;
        const int t = t_base + (int)tz;
        const uint32_t numAtoms = desc[b].numAtoms, t0_data = desc[b].t0_data;
        const double *Pt = P + (size_t)t * TCW_NCH * ppad;
        const uint32_t m0 = by * TCW_RECTP_ROWS;
        const uint32_t d0 = head ? 0u : DD + (bx - 1) * DT;
        const uint32_t t1_step = 32u * w.dtau;
        const int n_j = (int)((head ? DD : DT) / 32);
        const int j_full = w.N_tau > d0 ? (int)min((w.N_tau - d0) / 32u, (uint32_t)n_j) : 0;  // fully valid chunks
        float *Ft = Fmn ? Fmn + (size_t)tz * w.N_t0 * w.pitch : nullptr;
        uint32_t *gm = groupmax ? groupmax + ((size_t)tz * gx_total + bx) * n_grp + (size_t)by * TCW_RECTP_GROUPS : nullptr;
        uint32_t degenerate = 0;
        // the ticket of the NEXT row group is drawn before the current one is processed, so that the
        // shared-memory atomic's latency never sits between two groups
        uint32_t ticket = 0;
        if (lane == 0) ticket = atomicAdd(&desc[b].next, 1u);
#pragma unroll 1
        for (;;) {
            const uint32_t grp = __shfl_sync(0xffffffffu, ticket, 0);
            if (grp >= TCW_RECTP_GROUPS) break;
            if (lane == 0) ticket = atomicAdd(&desc[b].next, 1u);
            const bool noguard = (noguard_mask >> grp) & 1u;
            const uint32_t grow = grp * R;
            if (m0 + grow < w.N_t0) {
                float vgrp = -1.0f;
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
                    srow[r] = head ? sS[min(grow + r, (uint32_t)(TCW_RECTP_ROWS - 1))] : 0u;
                    rowp[r] = Ft ? Ft + ((size_t)(rowok[r] ? m : 0u) * w.pitch + d0 + lane) - r : nullptr;
                }
                f32x2 Rs2[2][TCW_NCH];
#pragma unroll
                for (int rp = 0; rp < 2; rp++) {
                    const f32x2 *pr = reinterpret_cast<const f32x2 *>(sR) + ((grow >> 1) + rp) * 8;
#pragma unroll
                    for (int c = 0; c < TCW_NCH; c++) Rs2[rp][c] = pr[c];
                }
                const double *sGg = sG + grp * 8;
#define RECTP_ROWS(DIAG_, CHK_, STORE_, NOG_, J0_, J1_)                                                              \
    rect_rows<R, DIAG_, CHK_, STORE_, false, NOG_>(sQ2, sP, sGg, sE, Rs2, rowp, rowok, srow, grow, d0, J0_, J1_, lane, \
                                                   w.N_tau, d_total, 0u, t1_step, a0, t0_data, numAtoms, g, Pt, ppad, \
                                                   best, vgrp, degenerate)
                if (head) {
                    if (Ft) RECTP_ROWS(true, true, true, false, 0, n_j);
                    else RECTP_ROWS(true, true, false, false, 0, n_j);
                } else {
                    const bool edge = m0 + grow + R > w.N_t0;  // d0 >= DD >= 32 > R - 1: no left edge here
#define RECTP_GROUP(STORE_, NOG_)                                                      \
    do {                                                                               \
        if (edge) {                                                                    \
            RECTP_ROWS(false, true, STORE_, NOG_, 0, n_j);                             \
        } else {                                                                       \
            RECTP_ROWS(false, false, STORE_, NOG_, 0, j_full);                         \
            if (j_full < n_j) RECTP_ROWS(false, true, STORE_, NOG_, j_full, n_j);      \
        }                                                                              \
    } while (0)
                    if (Ft) {
                        if (noguard) RECTP_GROUP(true, true);
                        else RECTP_GROUP(true, false);
                    } else {
                        if (noguard) RECTP_GROUP(false, true);
                        else RECTP_GROUP(false, false);
                    }
#undef RECTP_GROUP
                }
#undef RECTP_ROWS
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) vgrp = fmaxf(vgrp, __shfl_xor_sync(0xffffffffu, vgrp, o));
                if (lane == 0) {
                    if (gm) gm[grp] = vgrp > -1.0f ? float_orderable(vgrp) : 0u;
                    if (vgrp > -1.0f) atomicMax(&maxkey[t], pack_key(vgrp, 0xFFFFFFFFu));  // index part 0: completed later
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty[b]);  // this row group no longer reads the buffer
        }
        // every warp also checks out of the tile: the buffer (and its ticket counter) is not recycled
        // while a late warp may still draw from it
        if (lane == 0) mbar_arrive(&bar_empty[b]);
        if (__any_sync(0xffffffffu, degenerate != 0u) && lane == 0) atomicOr(&flags[t], TCW_FLAG_DEGENERATE);
    }
}
