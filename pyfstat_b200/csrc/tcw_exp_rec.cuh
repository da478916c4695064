// tcw_exp_rec.cuh -- exponential-window map as an FP64 recurrence (+ a tensor-core correction in
// `lal` lookup-table mode).  Canonical grids only (one row class, rows one atom apart, every
// template starting on the same atom, no start beyond the data end: tcw_b200.cu plan_exp `canon`).
//
// Exp.cu:82-102 sums, per cell, K ~ 3 tau / TAtom weighted atoms:  O(N_t0 N_tau K).  With rows one
// atom apart the weight of atom i for row m depends on k = i - i_t0(m) only, and for EXACT
// exponentials it is geometric in k:  w(k,n) = w0_n rho_n^(k - ka),  rho_n = e^{-TAtom/tau_n},
// k in [ka, kb_n].  Then
//     U_c[m,n] = sum_{j=0}^{L-1} X_c[s_m + ka + j] rho^(p_c j)            (L = kb - ka + 1)
//              = X_c[s_m + ka] + rho^p U_c[m+1,n] - rho^(pL) X_c[s_m + kb + 1]
// is a first-order recurrence down the rows of a column: O(1) per cell and channel instead of O(K).
// north_star allows a recurrence "only where it is shown to be numerically safe"; this one is:
//   * it is a CONTRACTION (0 < rho < 1): a rounding error made at row m is multiplied by rho at
//     every later row, so the accumulated error is bounded by eps * sum_j |X| rho^j -- the size of
//     the sum's own terms -- and does not grow with the number of rows walked;
//   * U is carried in FP64 (eps = 1.1e-16).  The atom that leaves the window was added with weight
//     1 and has been multiplied L times by rho; removing it with rho^L leaves a residue of
//     ~L eps |X| rho^L, 1e-12 relative at L = 17 000;
//   * measured (tests/test_gpu_parity.py::test_exp_recurrence_*): the FP64 recurrence agrees with
//     FP64 direct sums to 1e-13 and is CLOSER to them than the reference's sequential FP32 sums are.
// The sums then go through the same FP32 F-statistic epilogue as the other kernels.
//
// `lal` mode weighs atoms with XLALFastNegExp, a nearest-point table: w_lut(k,n) = e^{-x} (1 + d),
// |d| <= dx/2 (0.2 % at dx = 1/256), a sawtooth that no recurrence reproduces.  Split
//     sum_k X w_lut = sum_k X w_exact  +  sum_k X (w_lut - w_exact)
// The first term is the recurrence; the second is a Hankel contraction whose weights are <= 2e-3 of
// the first term's, so it needs 3 significant digits, not 7: it runs on the 5th-gen tensor cores in
// one pass with 11-bit operands (tcgen05.mma kind::f16 -- or kind::tf32 -- FP32 accumulation in
// TMEM).  The operand roundings (2^-11 each) and the tensor core's truncating accumulation (measured
// 6e-6 of sum|terms| at K = 5760, profiles/r02_tc_probe.txt) are multiplied by the 2e-3.
//
// Tensor-core kernel.  D_c[(i), n] = sum_k X_c[s0 + R i + k] V[k, n], R = elements per 16 bytes: in
// the no-swizzle K-major canonical layout a row is 16 bytes and the 8 rows of a core matrix are 16
// bytes apart, so a descriptor laid over the plain atom array IS a Hankel operand: the 8 rows of a
// core matrix are the map rows m, m + R, ..., m + 7 R (tools/tc_probe/hankel_tf32.cu).  The 8-row
// groups of an MMA are free to start anywhere (SBO), so group g takes the rows of CLASS g (m = m0 + g
// + R i'): the 64 rows of a channel are 64 CONSECUTIVE map rows, whose k ranges differ by less than
// one stage.  (The first version laid one descriptor over 64 rows of ONE class -- a band of 64 R = 512
// map rows per tile, executing 8 % more MMAs than needed at 120 d and 33 % more at 30 d because the k
// range of a tile is that of its longest row.)  Channels are stacked next to the rows: an MMA covers 3 or 4
// channels x 64 rows, each 8-row group reading its own 256-byte chunk of atoms (its rows' span + one
// stage of k; a start shifted by the class needs its own 16-byte aligned copy), the chunks 256 bytes
// apart (SBO); a prep kernel lays the atoms out in that chunked form -- per unit and atom block u:
// [group][channel][chunk] -- so a stage's atom operand is one or two contiguous bulk copies (6 KB, 2 x 4 KB).
// 128 window lengths per tile.  A tile is processed as two units -- the channels a2, b2, ab with the w^2
// table (one MMA of N = 3 x 64 rows per 32 bytes of K), then Fa | Fb with the w table (one MMA of N = 4 x 64
// rows) -- each unit in one half of TMEM.  The MMAs compute the TRANSPOSED tile (M = window lengths,
// N = channel x row), see the kernel.  Warp
// roles: TMA producer, MMA issuer, 4 epilogue warps (TMEM -> HBM scratch C); persistent CTAs, one
// per SM, static round-robin over tiles ordered by decreasing k range.
#pragma once
#include <cuda_fp16.h>

#include <type_traits>

#include "tcw_common.cuh"
#include "tcw_generic.cuh"

#define TCX_TAUS 128     // window lengths per tile (MMA N)
#define TCX_IROWS 64     // rows per channel per tile: 64 consecutive map rows
#ifndef TCX_STAGES
#define TCX_STAGES 8
#endif
#define TCX_A_BYTES 8192   // atoms of a stage: 24 (unit 0) or 32 (unit 1) chunks of 256 B
#define TCX_B_BYTES 16384  // 128 taus x 32 k x 4 B (one table)
#define TCX_STAGE_BYTES (TCX_A_BYTES + TCX_B_BYTES)
#define TCX_THREADS 192
#define TCX_SMEM (TCX_STAGES * TCX_STAGE_BYTES + 128)

__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// Operand precision of the tensor-core pass.  Both have an 11-bit significand; what differs is how many
// elements a 16-byte core-matrix row holds, i.e. the atom stride between consecutive MMA rows:
//   TF32: 4 -> 4 row classes, a core matrix spans 32 rows: groups 0-3 = classes 0-3 of rows m0 .. m0 + 31,
//              groups 4-7 the same for m0 + 32 .. m0 + 63; 32 k per 24-KB stage
//   FP16: 8 -> 8 row classes, group g = class g of rows m0 .. m0 + 63; 64 k per 24-KB stage: twice the MACs per
//         shared-memory byte, and shared-memory bandwidth is what bounds this kernel.
// FP16's range is handled by exact power-of-two scaling: atoms per template and channel group to
// [2^13, 2^14), weights by 2^12; the scale is undone on the FP32 accumulators in the epilogue.
template <bool F16>
struct TcxCfg {
    static constexpr int kRowStep = F16 ? 8 : 4;    // elements per 16 bytes = atoms between MMA rows = row classes
    static constexpr int kKC = 8 * kRowStep;        // k per stage (= atoms spanned by the rows of an 8-row group)
    static constexpr int kSpan = TCX_IROWS;         // map rows of a tile
    static constexpr int kUStep = TCX_IROWS / kKC;  // atom blocks (of kKC) per tile row block
    static constexpr int kChunk = 2 * kKC;          // elements per 256-byte chunk (8 rows' span + one stage of k)
    static constexpr int kElem = F16 ? 2 : 4;
    static constexpr size_t kTableElems = (size_t)TCX_TAUS * kKC;  // one table of one (nt, chunk): 16 KB
    // group gg (0..7) of a channel, row i' (0..7) of the group -> row of the tile, and the group's atom offset
    __host__ __device__ static constexpr int row_of(int gg, int ip) { return gg % kRowStep + kRowStep * ip + kKC * (gg / kRowStep); }
    __host__ __device__ static constexpr int atom_of(int gg) { return gg % kRowStep + kKC * (gg / kRowStep); }
};
#define TCX_VSCALE_LOG2 12
// Windows of fewer than TCX_SHORT_K + 1 atoms (the first columns of a map whose tau range starts at an atom or
// two) are not taken from this path in lookup-table mode: with three or four terms nothing averages the 2^-11
// operand roundings of the tensor-core pass (~1e-6 of the window sums), and few-atom windows have badly conditioned
// antenna-pattern matrices (cond ~ 400 at 3 atoms, two detectors) -- measured 2e-4 on F_mn there, above the 1e-4
// bar.  The generic kernel (the reference's own sequential sums, bit-identical to the oracle) computes those
// columns -- a few thousand cells per template.
#define TCX_SHORT_K 16

// Storage of the correction sums C between the tensor-core pass and the walk, per channel GROUP:
//   group A = a2, b2, ab (the antenna-pattern matrix M), group F = Fa, Fb (the data vector v), F = v^T M^-1 v.
// A relative error eps in M moves F by ~eps cond(M), the same error in v by ~eps sqrt(cond(M)) (v ~ N(0, M/2)),
// and cond reaches 1e4 on single-detector data.  FP16 (2^-11 of a sum that is <= 2e-3 of the window sum: 1e-6)
// is therefore safe for group F and not for group A: measured on H1-only 15-d maps, all-FP16 raises the largest
// F_mn error from 8.9e-5 to 1.2e-4 of the oracle's value (bar 1e-4); two-detector maps 2.6e-5 -> 3.1e-5.
// Default: A in FP32, F in FP16 -- 20 bytes per cell instead of 28.
#ifndef TCX_CA_F16
#define TCX_CA_F16 0
#endif
#ifndef TCX_CF_F16
#define TCX_CF_F16 1
#endif
struct TcxC {
    static constexpr bool kA16 = TCX_CA_F16 != 0, kF16 = TCX_CF_F16 != 0;
    static constexpr int kElemA = kA16 ? 2 : 4, kElemF = kF16 ? 2 : 4;
    static constexpr int kBytesPerCell = 3 * kElemA + 4 * kElemF;
    // a warp's row in the walk's ring: [a2 | b2 | ab | Fa_re | Fa_im | Fb_re | Fb_im] x 32 window lengths
    static constexpr int kRowA = 32 * kElemA, kRowF = 32 * kElemF, kRowBytes = 3 * kRowA + 4 * kRowF;
    static constexpr int kPiecesA = 3 * kRowA / 16, kPiecesF = 4 * kRowF / 16;
};

// per template: power-of-two scales of the two channel groups (a2,b2,ab | Fa,Fb).  scale[tz][0..1] multiplies the
// atoms into [2^13, 2^14); scale[tz][2..3] undoes it (and the weights' 2^12) on the accumulators.
__global__ void tcw_exptc_scale_kernel(const float *__restrict__ X, uint32_t xpad, const TplMeta *__restrict__ meta,
                                       int t_base, float *__restrict__ scale) {
    __shared__ float red[2][8];
    const int tz = blockIdx.x, t = t_base + tz;
    const uint32_t numAtoms = meta[t].numAtoms;
    float mx[2] = {0.0f, 0.0f};
    for (int c = 0; c < TCW_NCH; c++)
        for (uint32_t j = threadIdx.x; j < numAtoms; j += blockDim.x)
            mx[c >= 3] = fmaxf(mx[c >= 3], fabsf(__ldg(X + ((size_t)t * TCW_NCH + c) * xpad + j)));
    for (int g = 0; g < 2; g++) {
        float v = mx[g];
        for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
        if ((threadIdx.x & 31) == 0) red[g][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        float v = 0.0f;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) v = fmaxf(v, red[threadIdx.x][i]);
        int e = 0;
        if (v > 0.0f && v < INFINITY) {
            frexpf(v, &e);  // v = f 2^e, f in [0.5, 1)
            e -= 14;        // v 2^-e in [2^13, 2^14)
        }
        e = max(-100, min(100, e));
        scale[4 * tz + threadIdx.x] = ldexpf(1.0f, -e);
        scale[4 * tz + 2 + threadIdx.x] = ldexpf(1.0f, e - TCX_VSCALE_LOG2);
    }
}

// ---- atoms in chunked form, per template:
//        unit 0 (a2, b2, ab):  G0[u][gg 8][c 3][kChunk]          -- 6 KB per atom block u
//        unit 1 (Fa | Fb):     G1[p 2][u][gg 8][c' 2][kChunk]    -- 4 KB per pair and atom block
//      value = X_ch[i00 + atom_of(gg) + kKC u + e]
#define TCX_G0_BYTES 6144
#define TCX_G1_BYTES 4096
#define TCX_G_BYTES_PER_U (TCX_G0_BYTES + 2 * TCX_G1_BYTES)
template <bool F16>
__global__ void tcw_exptc_atoms_kernel(const float *__restrict__ X, uint32_t xpad, const TplMeta *__restrict__ meta,
                                       int t_base, uint32_t i00, uint32_t U, const float *__restrict__ scale,
                                       void *__restrict__ Gv) {
    using Cfg = TcxCfg<F16>;
    const int tz = blockIdx.y, t = t_base + tz;
    const uint32_t numAtoms = meta[t].numAtoms;
    const size_t n0 = (size_t)U * 24 * Cfg::kChunk, per_tpl = (size_t)U * 56 * Cfg::kChunk;
    const float s2 = scale[4 * tz], s1 = scale[4 * tz + 1];
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < per_tpl; idx += (size_t)gridDim.x * blockDim.x) {
        uint32_t e, gg, u, ch;
        if (idx < n0) {
            e = (uint32_t)(idx % Cfg::kChunk);
            size_t rest = idx / Cfg::kChunk;
            ch = (uint32_t)(rest % 3);
            rest /= 3;
            gg = (uint32_t)rest & 7u;
            u = (uint32_t)(rest >> 3);
        } else {
            const size_t i1 = idx - n0;
            e = (uint32_t)(i1 % Cfg::kChunk);
            size_t rest = i1 / Cfg::kChunk;
            const uint32_t cp = (uint32_t)rest & 1u;
            gg = ((uint32_t)rest >> 1) & 7u;
            rest >>= 4;
            u = (uint32_t)(rest % U);
            ch = 3u + 2u * (uint32_t)(rest / U) + cp;
        }
        const uint64_t j = (uint64_t)i00 + (uint32_t)Cfg::atom_of((int)gg) + (uint64_t)Cfg::kKC * u + e;
        float v = 0.0f;
        if (j < numAtoms) v = __ldg(X + ((size_t)t * TCW_NCH + ch) * xpad + j);
        v *= ch < 3 ? s2 : s1;
        if (F16) reinterpret_cast<__half *>(Gv)[(size_t)tz * per_tpl + idx] = __float2half_rn(v);
        else reinterpret_cast<float *>(Gv)[(size_t)tz * per_tpl + idx] = tf32_rna(v);
    }
}

// ---- correction weights V = w_lut - w_exact (table 0) and w_lut^2 - w_exact^2 (table 1) in the MMA's
//      canonical layout: Vt[nt][chunk][table][kq 8][ng 16][nr 8][kk kRowStep] (FP16: times 2^12) ----
template <bool F16>
__global__ void tcw_exptc_table_kernel(void *__restrict__ Vtv, const int32_t *__restrict__ Kn, uint32_t N_tau,
                                       uint32_t n_nt, uint32_t n_chunks, uint32_t tau, uint32_t dtau, uint32_t TAtom,
                                       int32_t delta, const ExpLut lut) {
    using Cfg = TcxCfg<F16>;
    constexpr uint32_t TE = (uint32_t)Cfg::kTableElems;
    const size_t total = (size_t)n_nt * n_chunks * TE;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const uint32_t off = (uint32_t)(idx % TE);
        const uint32_t kk = off % Cfg::kRowStep, nr = (off / Cfg::kRowStep) & 7u, ng = (off / (8 * Cfg::kRowStep)) & 15u,
                       kq = off / (128 * Cfg::kRowStep);
        const size_t rest = idx / TE;
        const uint32_t chunk = (uint32_t)(rest % n_chunks), nt = (uint32_t)(rest / n_chunks);
        const uint32_t k = chunk * Cfg::kKC + kq * Cfg::kRowStep + kk, n = nt * TCX_TAUS + ng * 8 + nr;
        double v1 = 0.0, v2 = 0.0;
        if (n < N_tau && (int32_t)k <= Kn[n]) {
            const uint32_t tau_n = tau + n * dtau;
            const long long t_rel = (long long)k * TAtom + delta;  // t_i - t0_m
            if (t_rel >= 0 && t_rel <= (long long)TCW_EXP_EFOLDING * tau_n) {
                const double x = __ddiv_rn((double)t_rel, (double)tau_n);
                const double wl = fast_neg_exp_lut(x, lut), we = exp(-x);
                v1 = wl - we;
                v2 = (wl - we) * (wl + we);
            }
        }
        if (F16) {
            __half *base = reinterpret_cast<__half *>(Vtv) + rest * 2 * TE;
            base[off] = __float2half_rn((float)ldexp(v1, TCX_VSCALE_LOG2));
            base[TE + off] = __float2half_rn((float)ldexp(v2, TCX_VSCALE_LOG2));
        } else {
            float *base = reinterpret_cast<float *>(Vtv) + rest * 2 * TE;
            base[off] = tf32_rna((float)v1);
            base[TE + off] = tf32_rna((float)v2);
        }
    }
}

// ---- tcgen05 helpers ----
__device__ __forceinline__ uint64_t tcx_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;  // no-swizzle K-major: start, leading (K) and stride (M/N) byte offsets in 16-byte units
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version 1 (sm_100)
    return d;
}
// D = F32, A and B K-major; operand format 2 = TF32 (kind::tf32), 0 = F16 (kind::f16)
__host__ __device__ constexpr uint32_t tcx_idesc(uint32_t M, uint32_t N, bool f16) {
    return (1u << 4) | ((f16 ? 0u : 2u) << 7) | ((f16 ? 0u : 2u) << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
template <bool F16>
__device__ __forceinline__ void tcx_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    if (F16)
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
            : "memory");
}
__device__ __forceinline__ void tcx_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tcx_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcx_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// -DTCX_TIMING (development): cycles each role spends waiting, summed over the CTAs of every launch
// (tools/tcx_timing.py reads them through tcw_debug_tcx_timing)
#ifdef TCX_TIMING
__device__ unsigned long long tcx_timing[16];
#define TCX_T0() const long long _t0 = clock64()
#define TCX_T1(acc) acc += clock64() - _t0
#else
#define TCX_T0()
#define TCX_T1(acc)
#endif

struct TcxTile {
    uint32_t tz, mb, nt;
    int nchunks;
};
// Tile j -> (template, row block, tau tile) and its number of k stages.  The two global loads a tile needs
// (the template's atom count, the longest window of the tau tile) are issued one tile ahead (tcx_tile_load)
// and consumed when the tile starts (tcx_tile_finish): an L2 round trip per tile in the single thread that
// issues the MMAs otherwise shows as 3-6 % of the pass.
struct TcxRaw {
    uint32_t numAtoms;
    int32_t kn;
};
__device__ __forceinline__ TcxRaw tcx_tile_load(uint32_t j, uint32_t n_tiles, uint32_t cnt, uint32_t n_mb, uint32_t n_nt,
                                                const MapWindow &w, const TplMeta *__restrict__ meta, int t_base,
                                                const int32_t *__restrict__ Kn) {
    TcxRaw r;
    r.numAtoms = 0;
    r.kn = -1;
    if (j < n_tiles) {
        const uint32_t tz = j % cnt, nt = n_nt - 1u - (j / cnt) / n_mb;
        const uint32_t n_last = min(nt * TCX_TAUS + TCX_TAUS, w.N_tau) - 1u;
        r.numAtoms = __ldg(&meta[t_base + tz].numAtoms);
        r.kn = __ldg(Kn + n_last);
    }
    return r;
}
template <bool F16>
__device__ __forceinline__ TcxTile tcx_tile_finish(uint32_t j, const TcxRaw &raw, uint32_t cnt, uint32_t n_mb, uint32_t n_nt,
                                                   uint32_t i00) {
    using Cfg = TcxCfg<F16>;
    TcxTile tl;
    tl.tz = j % cnt;
    const uint32_t rest = j / cnt;
    tl.mb = rest % n_mb;
    tl.nt = n_nt - 1u - rest / n_mb;  // widest windows first
    const long long s_first = (long long)i00 + (long long)tl.mb * Cfg::kSpan;  // the tile's first row has the longest k range
    const long long k_end = min((long long)raw.kn + 1, (long long)raw.numAtoms - s_first);
    tl.nchunks = k_end > 0 ? (int)((k_end + Cfg::kKC - 1) / Cfg::kKC) : 0;
    return tl;
}
#define TCX_TILE_LOOP_BEGIN                                                                        \
    TcxRaw raw_next = tcx_tile_load(blockIdx.x, n_tiles, cnt, n_mb, n_nt, w, meta, t_base, Kn);    \
    for (uint32_t j = blockIdx.x; j < n_tiles; j += gridDim.x) {                                   \
        const TcxTile tl = tcx_tile_finish<F16>(j, raw_next, cnt, n_mb, n_nt, i00);                \
        raw_next = tcx_tile_load(j + gridDim.x, n_tiles, cnt, n_mb, n_nt, w, meta, t_base, Kn);

// CA[tz][3][m][cpitch], CF[tz][4][m][cpitch]: correction sums of every cell of the sub-batch (cpitch = n_nt * 128)
// in the accumulators' own (scaled) units times the power of two `cs`, which the host derives from a bound on
// sum_k |X V| so that no FP16 value can overflow; the walk undoes the scaling.  Precision per group: TcxC.
// A tile is processed as two UNITS -- the channels a2, b2, ab with the w^2 table, then Fa_re, Fa_im, Fb_re, Fb_im
// with the w table; no operand is shared between them, so the split costs no traffic -- each accumulating into
// one half of TMEM (192 / 256 of its 256 columns): the epilogue warps drain unit u while the MMAs of unit u + 1 run.
template <bool F16>
__global__ void __launch_bounds__(TCX_THREADS, 1)
tcw_exptc_map_kernel(const void *__restrict__ Gv, uint32_t U, const void *__restrict__ Vtv, uint32_t n_chunks_tab,
                     const int32_t *__restrict__ Kn, const TplMeta *__restrict__ meta, int t_base, uint32_t cnt,
                     MapWindow w, uint32_t i00, uint32_t n_nt, uint32_t n_mb, uint32_t n_tiles,
                     float cs, unsigned char *__restrict__ CA, unsigned char *__restrict__ CF, uint32_t cpitch) {
    using Cfg = TcxCfg<F16>;
    extern __shared__ __align__(128) unsigned char tcx_smem_raw[];
    __shared__ __align__(8) uint64_t full[TCX_STAGES], empty[TCX_STAGES], tmem_full[2], tmem_empty[2];
    __shared__ uint32_t tmem_base_s;
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(tcx_smem_raw) + 127) & ~(uintptr_t)127);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned char *G = reinterpret_cast<const unsigned char *>(Gv);
    const unsigned char *Vt = reinterpret_cast<const unsigned char *>(Vtv);

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < TCX_STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
#pragma unroll
        for (int b = 0; b < 2; b++) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], 128);
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcx_fence_before();
    __syncthreads();
    tcx_fence_after();
    const uint32_t tmem = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {  // ---- TMA producer ----
            uint32_t it = 0;
            long long tp_wait = 0, tp_total = 0;
            (void)tp_wait;
            (void)tp_total;
#ifdef TCX_TIMING
            const long long tp_begin = clock64();
#endif
            TCX_TILE_LOOP_BEGIN
                // atoms of a stage (atom block u = kUStep mb + c): unit 0 one 6-KB copy, unit 1 a 4-KB copy per pair
                const unsigned char *gT = G + (size_t)tl.tz * U * TCX_G_BYTES_PER_U;
                const unsigned char *gA0 = gT + (size_t)Cfg::kUStep * tl.mb * TCX_G0_BYTES;                                // + c * 6 KB
                const unsigned char *gA1 = gT + (size_t)U * TCX_G0_BYTES + (size_t)Cfg::kUStep * tl.mb * TCX_G1_BYTES;  // + (p U + c) * 4 KB
                const unsigned char *gB = Vt + (size_t)tl.nt * n_chunks_tab * 32768;  // + c * 32 KB (+ 16 KB: w^2)
                for (int hh = 0; hh < 2; hh++)
                    for (int c = 0; c < tl.nchunks; c++, it++) {
                        const uint32_t s = it % TCX_STAGES;
                        {
                            TCX_T0();
                            mbar_wait(&empty[s], ((it / TCX_STAGES) & 1u) ^ 1u);
                            TCX_T1(tp_wait);
                        }
                        unsigned char *st = smem + (size_t)s * TCX_STAGE_BYTES;
#ifdef TCX_PROBE_NO_TMA  // timing probe: no operand traffic (the MMAs run on whatever the stage holds)
                        if (it >= TCX_STAGES) {
                            mbar_arrive_plain(&full[s]);
                            continue;
                        }
#endif
                        if (hh == 0) {
                            mbar_arrive_expect_tx(&full[s], TCX_G0_BYTES + TCX_B_BYTES);
                            bulk_g2s(st, gA0 + (size_t)c * TCX_G0_BYTES, TCX_G0_BYTES, &full[s]);
                        } else {
                            mbar_arrive_expect_tx(&full[s], 2 * TCX_G1_BYTES + TCX_B_BYTES);
#pragma unroll
                            for (int pl = 0; pl < 2; pl++)
                                bulk_g2s(st + pl * TCX_G1_BYTES, gA1 + ((size_t)pl * U + c) * TCX_G1_BYTES, TCX_G1_BYTES, &full[s]);
                        }
                        bulk_g2s(st + TCX_A_BYTES, gB + (size_t)c * 32768 + (hh == 0 ? 16384 : 0), TCX_B_BYTES, &full[s]);
                    }
            }
#ifdef TCX_TIMING
            atomicAdd(&tcx_timing[0], (unsigned long long)(clock64() - tp_begin));
            atomicAdd(&tcx_timing[1], (unsigned long long)tp_wait);
#endif
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---- MMA issuer ----
            constexpr uint32_t idesc = tcx_idesc(128, 128, F16), idesc3 = tcx_idesc(128, 192, F16), idesc4 = tcx_idesc(128, 256, F16);
            (void)idesc;
            // The MMA computes the TRANSPOSED tile: its A operand (M = 128) are the window lengths -- the weights
            // V[k, n] -- its B operand the Hankel atoms: N = 3 x 64 (a2, b2, ab: unit 0) or 4 x 64 rows (Fa_re, Fa_im,
            // Fb_re, Fb_im: unit 1) -- ONE MMA per unit and 32 bytes of K, so the weights are read from shared memory
            // once (two N = 128 MMAs for unit 1, -DTCX_UNIT1_SPLIT, read them twice: 7.65 instead of 6.77 ms per
            // 8 x 120-d maps).  TMEM lane = window length: the epilogue's stores are coalesced as they come (32 lanes = 32
            // consecutive window lengths of one row).  (The first versions stacked the channels in M, two per MMA:
            // unit 0 then needs two M128 MMAs with a quarter of their lanes idle; as N = 192 it is one MMA of 3/4 the
            // duration -- 1/8 fewer tensor-pipe cycles per tile, and this pass runs at the board's power cap.)
            const uint64_t dx = tcx_desc(0, 16, 256);      // atoms: K halves 16 B apart, 8-row groups = chunks 256 B apart
            const uint64_t dv = tcx_desc(0, 2048, 128);    // weights: [kq][ng][8][16 B]: K steps 2 KB apart, 8-column groups 128 B
            uint32_t it = 0, unit = 0;
            long long tm_dec = 0, tm_tmem = 0, tm_full = 0, tm_n = 0;
            (void)tm_dec, (void)tm_tmem, (void)tm_full, (void)tm_n;
#ifdef TCX_TIMING
            const long long tm_begin = clock64();
#endif
#ifdef TCX_TIMING
            long long _td = clock64();
#endif
            TCX_TILE_LOOP_BEGIN
#ifdef TCX_TIMING
                tm_dec += (tl.nchunks >= 0 ? clock64() : 0) - _td;
                tm_n += 12 * tl.nchunks;
#endif
                for (int hh = 0; hh < 2; hh++, unit++) {
                    const uint32_t buf = unit & 1u;
                    {
                        TCX_T0();
                        mbar_wait(&tmem_empty[buf], ((unit >> 1) & 1u) ^ 1u);
                        TCX_T1(tm_tmem);
                    }
                    tcx_fence_after();
                    const uint32_t d0 = tmem + buf * 256u;
                    for (int c = 0; c < tl.nchunks; c++, it++) {
                        const uint32_t s = it % TCX_STAGES;
                        {
                            TCX_T0();
                            mbar_wait(&full[s], (it / TCX_STAGES) & 1u);
                            TCX_T1(tm_full);
                        }
                        tcx_fence_after();
                        const uint32_t a0 = smem_u32(smem + (size_t)s * TCX_STAGE_BYTES), b0 = a0 + TCX_A_BYTES;
#pragma unroll
                        for (int q = 0; q < 4; q++) {  // 32 bytes of K per MMA: 8 TF32 / 16 FP16 values
                            const uint64_t vd = dv | (uint64_t)(((b0 + q * 4096) >> 4) & 0x3FFF);
                            const uint64_t xd0 = dx | (uint64_t)(((a0 + q * 32) >> 4) & 0x3FFF);
                            const uint32_t acc = (c > 0 || q > 0) ? 1u : 0u;
#ifdef TCX_UNIT1_SPLIT  // two N = 128 MMAs (one per pair) instead of one N = 256: the weights are read twice
                            if (hh == 0) {
                                tcx_mma<F16>(d0, vd, xd0, idesc3, acc);
                            } else {
                                const uint64_t xd1 = dx | (uint64_t)(((a0 + TCX_G1_BYTES + q * 32) >> 4) & 0x3FFF);
                                tcx_mma<F16>(d0, vd, xd0, idesc, acc);
                                tcx_mma<F16>(d0 + 128, vd, xd1, idesc, acc);
                            }
#else
                            // unit 0: N = 192 (24 chunks), unit 1: N = 256 (the two pairs' 2 x 16 chunks are adjacent)
                            tcx_mma<F16>(d0, vd, xd0, hh == 0 ? idesc3 : idesc4, acc);
#endif
                        }
                        tcx_commit(&empty[s]);  // the stage is free once these MMAs have read it
                    }
                    if (tl.nchunks > 0) tcx_commit(&tmem_full[buf]);
                    else mbar_arrive_plain(&tmem_full[buf]);
                }
#ifdef TCX_TIMING
                _td = clock64();
#endif
            }
#ifdef TCX_TIMING
            atomicAdd(&tcx_timing[2], (unsigned long long)(clock64() - tm_begin));
            atomicAdd(&tcx_timing[3], (unsigned long long)tm_dec);
            atomicAdd(&tcx_timing[4], (unsigned long long)tm_tmem);
            atomicAdd(&tcx_timing[5], (unsigned long long)tm_full);
            atomicAdd(&tcx_timing[6], (unsigned long long)tm_n);
#endif
        }
    } else {  // ---- epilogue warps: TMEM -> C ----
        // TMEM lane = window length n, column = (channel slot c', row): a thread holds ONE window length and, per
        // tcgen05.ld, 32 (slot, row) values; the 32 lanes of a warp store 32 consecutive window lengths of one
        // row -- a full 128-byte line (FP32) or 64 bytes (FP16) per instruction, no transposition needed.  The
        // load of the next 32 columns is in flight while these are stored.
        const uint32_t q = warp & 3u;  // TMEM lane quadrant this warp may read
        uint32_t unit = 0;
        long long te_wait = 0;
        (void)te_wait;
#ifdef TCX_TIMING
        const long long te_begin = clock64();
#endif
#define TCX_LD32(V, TADDR)                                                                                                   \
    asm volatile(                                                                                                            \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                            \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                            \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                            \
        : "=r"(V[0]), "=r"(V[1]), "=r"(V[2]), "=r"(V[3]), "=r"(V[4]), "=r"(V[5]), "=r"(V[6]), "=r"(V[7]), "=r"(V[8]),        \
          "=r"(V[9]), "=r"(V[10]), "=r"(V[11]), "=r"(V[12]), "=r"(V[13]), "=r"(V[14]), "=r"(V[15]), "=r"(V[16]),             \
          "=r"(V[17]), "=r"(V[18]), "=r"(V[19]), "=r"(V[20]), "=r"(V[21]), "=r"(V[22]), "=r"(V[23]), "=r"(V[24]),            \
          "=r"(V[25]), "=r"(V[26]), "=r"(V[27]), "=r"(V[28]), "=r"(V[29]), "=r"(V[30]), "=r"(V[31])                          \
        : "r"(TADDR))
        // one unit (`u0`: unit 0 = a2, b2, ab in 192 columns [gg][c 3][i']; else Fa | Fb in 256 columns
        // [Fa | Fb][gg][re | im][i']); `c16`: its group is stored as FP16; `edge`: the tile reaches beyond the last map row
        auto drain_unit = [&](auto u0, auto c16, auto edge, const TcxTile &tl, uint32_t buf) __attribute__((always_inline)) {
            constexpr bool U0 = decltype(u0)::value, C16 = decltype(c16)::value, EDGE = decltype(edge)::value;
            constexpr int ES = C16 ? 2 : 4, NPAIR = U0 ? 3 : 4;  // pairs of 32-column loads
            const bool has = tl.nchunks > 0;
            const uint32_t tbase = tmem + ((q * 32u) << 16) + buf * 256u;
            const int nch = U0 ? 3 : 4;
            const uint32_t m0 = tl.mb * Cfg::kSpan;
            const size_t rowb = (size_t)cpitch * ES, chb = (size_t)w.N_t0 * rowb;
            unsigned char *dstb = (U0 ? CA : CF) + ((size_t)tl.tz * nch * w.N_t0 + m0) * rowb +
                                  ((size_t)tl.nt * TCX_TAUS + 32u * q + lane) * ES;
            const size_t rowstep = (size_t)Cfg::kRowStep * rowb;  // rows of a group are kRowStep apart
            uint32_t va[32], vb[32];
            // columns 32 j ..: the four 8-row groups 4 j .. 4 j + 3
            auto store32 = [&](const uint32_t(&cur)[32], int j) __attribute__((always_inline)) {
#pragma unroll
                for (int h = 0; h < 4; h++) {
                    const int grp = 4 * j + h;
                    // group -> (gg, channel index within the stored group)
                    const int gg = U0 ? grp / 3 : (grp & 15) >> 1;
                    const int ci = U0 ? grp % 3 : 2 * (grp >> 4) + (grp & 1);
                    const int r0 = gg % Cfg::kRowStep + Cfg::kKC * (gg / Cfg::kRowStep);  // row_of(gg, 0)
                    unsigned char *d = dstb + (size_t)ci * chb + (size_t)r0 * rowb;
#pragma unroll
                    for (int ip = 0; ip < 8; ip++) {
                        const float val = has ? __uint_as_float(cur[8 * h + ip]) * cs : 0.0f;
                        if (!EDGE || m0 + (uint32_t)(r0 + Cfg::kRowStep * ip) < w.N_t0) {
                            if (C16) *reinterpret_cast<__half *>(d) = __float2half_rn(val);
                            else *reinterpret_cast<float *>(d) = val;
                        }
                        d += rowstep;
                        asm volatile("" : "+l"(d));  // one running pointer: keeps nvcc from materialising 64 row addresses
                    }
                }
            };
            // (a unit without MMAs -- no atom in any of its windows -- reads whatever TMEM holds and stores zeros)
            TCX_LD32(va, tbase);
#pragma unroll 1
            for (int jj = 0; jj < NPAIR; jj++) {
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                TCX_LD32(vb, tbase + (uint32_t)(32 * (2 * jj + 1)));
                store32(va, 2 * jj);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (jj + 1 < NPAIR) TCX_LD32(va, tbase + (uint32_t)(32 * (2 * jj + 2)));
                store32(vb, 2 * jj + 1);
            }
        };
        TCX_TILE_LOOP_BEGIN
            const bool edge = tl.mb * Cfg::kSpan + Cfg::kSpan > w.N_t0;
            for (int hh = 0; hh < 2; hh++, unit++) {
                const uint32_t buf = unit & 1u;
                {
                    TCX_T0();
                    mbar_wait(&tmem_full[buf], (unit >> 1) & 1u);
                    TCX_T1(te_wait);
                }
                tcx_fence_after();
                using A16 = std::integral_constant<bool, TcxC::kA16>;
                using F16c = std::integral_constant<bool, TcxC::kF16>;
                if (hh == 0) {
                    if (edge) drain_unit(std::true_type{}, A16{}, std::true_type{}, tl, buf);
                    else drain_unit(std::true_type{}, A16{}, std::false_type{}, tl, buf);
                } else {
                    if (edge) drain_unit(std::false_type{}, F16c{}, std::true_type{}, tl, buf);
                    else drain_unit(std::false_type{}, F16c{}, std::false_type{}, tl, buf);
                }
                tcx_fence_before();
                mbar_arrive_plain(&tmem_empty[buf]);
            }
        }
#undef TCX_LD32
#ifdef TCX_TIMING
        if (tid == 64) {
            atomicAdd(&tcx_timing[7], (unsigned long long)(clock64() - te_begin));
            atomicAdd(&tcx_timing[8], (unsigned long long)te_wait);
            atomicAdd(&tcx_timing[9], 1ull);
        }
#endif
    }
    tcx_fence_before();
    __syncthreads();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

// ---- the walk: a warp per (32 window lengths, row segment), rows from the segment's end down ----
// The recurrence is linear, so the rows of a column split into NSEG segments walked concurrently
// (NSEG chosen by the host so that the launch fills the GPU; 1 when the batch alone does):
// pass 1 walks every segment from a zero state (no output) and leaves its end value E_s in shared memory;
// the state entering segment s is then  sum_{s' > s} rho^(p SEG (s' - s - 1)) E_s'  (Horner, a few terms);
// pass 2 walks the segment again from that state and emits the cells.
//
// Operands of a row step, per warp -- all of them from shared memory, none held in registers across rows.
// A warp at 2-6 warps per scheduler issues nearly serially, so what counts is the number of instructions
// per step and the latency of each; the versions measured on the way (8 x 120-d maps, exact mode):
//   * atoms as 32-byte FP32 records from global memory (broadcast load of the ENTERING atom, 96-byte-stride
//     gather of the LEAVING atoms), FP32 -> FP64 conversions: 2.4 ms whatever the segmentation -- 150
//     instructions per step, 14 conversions per cell and pass on the XU pipe (ncu: XU 52 %), and the
//     compiler reusing the dead padding register of a prefetching load as a temporary (35 % of all stall
//     samples on one instruction that waits for that load);
//   * the same with the leaving atoms in a shared-memory ring: 2.4 ms; FP64 copies of the atoms (no
//     conversions) but 16 more registers for the prefetch: 2.6-3.0 ms (occupancy);
//   * this version: 1.80 ms.  The 32 leaving atoms of a row lie in a window of <= 97 consecutive atoms that
//     slides down by ONE atom per row: the warp keeps it in a ring (FP64, channel-major, 128 atoms; the
//     stride-3 reads are bank-conflict-free) and fetches 16 new atoms every 16 rows; the entering atom comes
//     from a second ring of 32 atoms (a broadcast read).  FP64 copies of the atoms
//     (tcw_exp_atoms_f64_kernel) feed DFMAs directly; the only conversions left are the 7 sums going to FP32.
// Correction sums (HAS_C; FP32 for a2, b2, ab and FP16 for Fa, Fb: TcxC) are prefetched TCX_WALK_DEPTH rows ahead
// into a third ring (a row of a warp = 3 x 128 B + 4 x 64 B = 40 16-byte pieces, one or two per lane).  Everything arrives by cp.async.cg -- one commit
// group per row, so a fixed wait_group count covers all rings -- bypassing L1.
#ifndef TCX_WALK_DEPTH
#define TCX_WALK_DEPTH 4
#endif
#define TCX_WALK_XRING 128  // atoms in the leaving-atom ring (8 blocks of 16)
template <int NSEG>
struct WalkCfg {
#ifndef TCX_WALK_CGW
#define TCX_WALK_CGW 2  // warps per CTA when the rows are not segmented
#endif
    static constexpr int kCG = NSEG >= TCX_WALK_CGW ? 1 : TCX_WALK_CGW / NSEG;  // column groups (of 32 window lengths) per CTA
    static constexpr int kWarps = NSEG * kCG;
    static constexpr int kThreads = 32 * kWarps;
    static constexpr int kWarpXBytes = TCW_NCH * (TCX_WALK_XRING + 32) * 8;  // leaving-atom ring + entering-atom ring
    static constexpr int kXRingBytes = kWarps * kWarpXBytes;                 // (also holds E between the passes)
    static constexpr int kDepth = kWarps > 4 ? 4 : TCX_WALK_DEPTH;  // rows of correction sums in flight per warp
    static constexpr int kCRingBytes = kWarps * kDepth * TcxC::kRowBytes;
    static constexpr int smem(bool has_c) { return kXRingBytes + (has_c ? kCRingBytes : 0); }
};

__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
// src_bytes = 0: the 16 destination bytes are zero-filled and the source is not read
__device__ __forceinline__ void cp_async16_zfill(void *dst_smem, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// FP64 copy of the merged atoms for the walk: Xd[t][7][xpad], channel-major and zero padded like X.
__global__ void tcw_exp_atoms_f64_kernel(const float *__restrict__ X, uint32_t xpad, int t_base, double *__restrict__ Xd) {
    const int t = t_base + blockIdx.y;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < xpad; j += gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < TCW_NCH; c++)
            Xd[((size_t)t * TCW_NCH + c) * xpad + j] = (double)__ldg(X + ((size_t)t * TCW_NCH + c) * xpad + j);
    }
}

struct WalkRingOn {
    __device__ constexpr operator bool() const { return true; }
};
struct WalkRingOff {
    __device__ constexpr operator bool() const { return false; }
};
// STEP1: rows one atom apart (rowstep == 1, every refined row is a map row): the emission test is compiled out --
// with it in the loop the canonical maps run 8-25 % slower (the F-statistic epilogue becomes conditional code).
template <bool HAS_C, int NSEG, bool STEP1>
__global__ void __launch_bounds__(WalkCfg<NSEG>::kThreads)
tcw_exp_walk_kernel(const double *__restrict__ X, uint32_t xpad,
                    const int32_t *__restrict__ Kn, const TplMeta *__restrict__ meta, int t_base, MapWindow w,
                    uint32_t i00, int32_t delta, uint32_t TAtom, int rowstep, const unsigned char *__restrict__ CA,
                    const unsigned char *__restrict__ CF, uint32_t c_rows, uint32_t cpitch,
                    const float *__restrict__ cscale, float cunshift, uint32_t n_skip, float *__restrict__ Fmn,
                    unsigned long long *__restrict__ maxkey, uint32_t *__restrict__ flags) {
    using Cfg = WalkCfg<NSEG>;
    extern __shared__ __align__(16) unsigned char walk_smem[];
    __shared__ unsigned long long red[Cfg::kWarps];
    const int tz = blockIdx.y, t = t_base + tz;
    const uint32_t numAtoms = meta[t].numAtoms;
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int seg = wi % NSEG, cg = wi / NSEG;
    double(*ringX)[TCX_WALK_XRING] = reinterpret_cast<double(*)[TCX_WALK_XRING]>(
        walk_smem + (size_t)wi * Cfg::kWarpXBytes);  // [channel][atom & 127]: atoms leaving the windows
    double(*ringX0)[32] = reinterpret_cast<double(*)[32]>(
        walk_smem + (size_t)wi * Cfg::kWarpXBytes + TCW_NCH * TCX_WALK_XRING * 8);  // [channel][atom & 31]: atoms entering
    // end values of pass 1, [channel][lane] at the start of each warp's (then idle) atom ring
    auto E = [&](int wj) { return reinterpret_cast<double(*)[32]>(walk_smem + (size_t)wj * Cfg::kWarpXBytes); };
    unsigned char *ringC = walk_smem + Cfg::kXRingBytes + (size_t)wi * Cfg::kDepth * TcxC::kRowBytes;  // [slot][channel][lane]
    // the correction sums come in the tensor-core pass's scaled units (tcw_exptc_scale_kernel, times 2^-shift)
    const float cf2 = HAS_C ? cscale[4 * tz + 2] * cunshift : 0.0f, cf1 = HAS_C ? cscale[4 * tz + 3] * cunshift : 0.0f;
    const uint32_t n = (blockIdx.x * Cfg::kCG + cg) * 32 + lane;
    // the first n_skip columns (windows of a few atoms) are computed by the generic kernel, see the host
    const bool in_range = n < w.N_tau, active = in_range && n >= n_skip;
    const double *Xs = X + (size_t)t * TCW_NCH * xpad;

    // the column's window: k in [ka, kb], weights w0 rho^(k - ka)
    const uint32_t nn = in_range ? n : w.N_tau - 1;  // (skipped columns keep their own geometry: the warp's ring span)
    const int K = Kn[nn];
    const long long tau_n = (long long)w.tau + (long long)nn * w.dtau;
    const int ka = delta < 0 ? 1 : 0;
    const long long num = (long long)TCW_EXP_EFOLDING * tau_n - delta;  // t_rel <= 3 tau
    const int kb = (int)min((long long)K, num >= 0 ? num / (long long)TAtom : -1ll);
    const int L = kb - ka + 1;
    const bool empty_win = L <= 0;
    const double inv_tau = 1.0 / (double)tau_n;
    const double rho = exp(-(double)TAtom * inv_tau), rho2 = rho * rho;
    const double w0 = empty_win ? 0.0 : exp(-((double)ka * TAtom + (double)delta) * inv_tau), w02 = w0 * w0;
    const double rhoL = empty_win ? 0.0 : exp(-(double)L * (double)TAtom * inv_tau);
    const double rhoL2 = rhoL * rhoL;
    const float w0f = (float)w0, w02f = (float)w02;  // scale of the sums, applied in FP32 after the conversion

    // leaving atom of row m: index m + off.  The warp's 32 offsets span <= 96 atoms whenever dtau <= TAtom
    // (3 dtau / TAtom per column); wider spans take the plain gather.
    const int off_own = (int)i00 + kb + 1;
    int off_min = empty_win ? 0x7fffffff : off_own, off_max = empty_win ? -0x7fffffff : off_own;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        off_min = min(off_min, __shfl_xor_sync(0xffffffffu, off_min, o));
        off_max = max(off_max, __shfl_xor_sync(0xffffffffu, off_max, o));
    }
    if (off_max < off_min) off_min = off_max = 0;  // no lane has a window
    const bool use_ring = off_max - off_min <= TCX_WALK_XRING - 32;
    const int off = empty_win ? off_min : off_own;

    // REFINED rows 0 .. R-1, one per atom; map row m is refined row rowstep * m (rowstep = dt0 / TAtom; the plan
    // guarantees that the last map row lies within the data), the others and those beyond the map give no
    // output; segment `seg` owns refined rows [lo, hi)
    const int R = (int)numAtoms - (int)i00;
    const int SEG = (R + NSEG - 1) / NSEG;
    const int lo = min(seg * SEG, R), hi = min(lo + SEG, R);

    double U[TCW_NCH];
#pragma unroll
    for (int c = 0; c < TCW_NCH; c++) U[c] = 0.0;

    float *Ft = Fmn ? Fmn + (size_t)tz * w.N_t0 * w.pitch + n : nullptr;  // + m * pitch
    float best = -1.0f;
    int best_m = 0;

    // atoms [16 b, 16 b + 16) of the 7 channels into the ring: 56 pieces of 16 bytes (2 atoms), two per lane
    // (zeros outside the padded array; the array itself is zero beyond numAtoms)
    auto fetch_block = [&](int b, bool entering) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int q = lane + 32 * h;
            if (q < 8 * TCW_NCH) {
                const int c = q >> 3;
                const int j = 16 * b + 2 * (q & 7);
                const bool ok = j >= 0 && (uint32_t)j + 2u <= xpad;
                double *dst = entering ? &ringX0[c][j & 31] : &ringX[c][j & (TCX_WALK_XRING - 1)];
                cp_async16_zfill(dst, Xs + (size_t)c * xpad + (ok ? j : 0), ok ? 16u : 0u);
            }
        }
    };
    const int off0 = (int)i00 + ka;  // entering atom of row m: index m + off0
    // everything the pass starting at row m_first needs at once: blocks from one below the lowest
    // leaving atom up to the highest
    // next rows at whose end a ring block is refilled (every 16 rows each, at their own phases): one compare per
    // row instead of two masked tests
    int evL = 0, evE = 0, ev = 0;
    auto fill_ring = [&](int m_first) {
        __syncwarp();
        if (use_ring) {
            const int jl = m_first + off_min;
            for (int b = (jl >> 4) - 1; b <= (jl + (off_max - off_min)) >> 4; b++) fetch_block(b, false);
        }
        fetch_block((m_first + off0) >> 4, true);
        fetch_block(((m_first + off0) >> 4) - 1, true);
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
        // the largest m <= m_first with ((m + o) & 15) == 15
        evL = use_ring ? m_first - (((m_first + off_min) + 1) & 15) : -0x40000000;
        evE = m_first - (((m_first + off0) + 1) & 15);
        ev = max(evL, evE);
    };
    // end of row m: the block needed 16 rows from now joins this row's commit group
    auto row_end = [&](int m, auto ring) {
        if (m == ev) {
            if (ring && m == evL) {
                fetch_block(((m + off_min) >> 4) - 1, false);
                evL -= 16;
            }
            if (m == evE) {
                fetch_block(((m + off0) >> 4) - 1, true);  // into the half the rows above have left
                evE -= 16;
            }
            ev = max(evL, evE);
        }
        cp_async_commit();
    };

    // entering atom of a row, fetched one row ahead (one broadcast 32-byte record)
    // (rows m < 0 and atoms beyond the data read the zero padding of the array: xpad >= numAtoms + 64)
    // `ring`: use_ring as a bool, or as a compile-time constant in the main loops (a warp-uniform branch per row less)
    auto step = [&](int m, auto ring) {
        double x0[TCW_NCH], x1[TCW_NCH];
        {
            const int slot0 = (m + off0) & 31;  // the same address in every lane: a broadcast read
#pragma unroll
            for (int c = 0; c < TCW_NCH; c++) x0[c] = ringX0[c][slot0];
        }
        if (ring) {
            const int slot = (m + off) & (TCX_WALK_XRING - 1);
#pragma unroll
            for (int c = 0; c < TCW_NCH; c++) x1[c] = ringX[c][slot];
        } else {
            const uint32_t j1 = min((uint32_t)(m + off), numAtoms);  // the record at numAtoms is zero padding
#pragma unroll
            for (int c = 0; c < TCW_NCH; c++) x1[c] = __ldg(Xs + (size_t)c * xpad + j1);
        }
        // U = rho^p U + (X[s + ka] - rho^(pL) X[s + kb + 1]), all in FP64.
        // (Columns with an empty window keep accumulating the entering atoms; their cells are forced to the
        // F = 2 fallback in cell().)
#pragma unroll
        for (int c = 0; c < TCW_NCH; c++) U[c] = fma(c < 3 ? rho2 : rho, U[c], fma(-(c < 3 ? rhoL2 : rhoL), x1[c], x0[c]));
    };
    // cc: the slot of the correction-sum ring holding row m
    // (running pointers instead of row * pitch products: the walk is bound by instruction issue, and the 64-bit
    // multiply-adds of the addresses were ~20 of its ~170 instructions per row)
    float *Fp = nullptr;  // &Ft[m * pitch] of the row about to be emitted
    auto cell = [&](int m, const unsigned char *cc, auto has_f) {
        float S[TCW_NCH];
#pragma unroll
        for (int c = 0; c < TCW_NCH; c++) {
            const float u = (float)U[c];
            float corr = 0.0f;
            if (HAS_C) {
                const unsigned char *pc = cc + (c < 3 ? c * TcxC::kRowA : 3 * TcxC::kRowA + (c - 3) * TcxC::kRowF);
                if (c < 3 ? TcxC::kA16 : TcxC::kF16) corr = __half2float(reinterpret_cast<const __half *>(pc)[lane]);
                else corr = reinterpret_cast<const float *>(pc)[lane];
            }
            S[c] = HAS_C ? fmaf(corr, c < 3 ? cf2 : cf1, u * (c < 3 ? w02f : w0f)) : u * (c < 3 ? w02f : w0f);
        }
        float F = fstat_fast(S[0], S[1], S[2], S[3], S[4], S[5], S[6]);
        if (empty_win) F = 2.0f;  // no atom in the window: all sums zero in the reference -> its fallback value
        if (active) {
            if (has_f) *Fp = F;
            if (F >= best) {  // rows are walked downwards: among equal F the smaller row wins (np.argmax order)
                best = F;
                best_m = m;
            }
        }
        if (has_f) Fp -= w.pitch;
    };

    if (NSEG > 1) {
        // ---- pass 1: the segment from a zero state (nobody needs the lowest segment's end value) ----
        if (seg > 0 && hi > lo) {
            fill_ring(hi - 1);
#pragma unroll 1
            for (int m = hi - 1; m >= lo; m--) {
                cp_async_wait<Cfg::kDepth - 1>();
                __syncwarp();
                step(m, use_ring);
                __syncwarp();
                row_end(m, use_ring);
            }
        }
        cp_async_wait<0>();  // no copy into the ring is in flight any more
        __syncwarp();
#pragma unroll
        for (int c = 0; c < TCW_NCH; c++) E(wi)[c][lane] = U[c];
        __syncthreads();
        // ---- state entering the segment ----
        const double rS = exp(-(double)SEG * (double)TAtom * inv_tau), rS2 = rS * rS;
#pragma unroll
        for (int c = 0; c < TCW_NCH; c++) U[c] = 0.0;
        for (int sp = NSEG - 1; sp > seg; sp--) {
#pragma unroll
            for (int c = 0; c < TCW_NCH; c++) U[c] = fma(c < 3 ? rS2 : rS, U[c], E(cg * NSEG + sp)[c][lane]);
        }
        __syncthreads();  // every warp has read the end values: the rings may be refilled
    }

    // ---- pass 2: the segment (again) from its true state, with output ----
    const int last_emit = rowstep * ((int)w.N_t0 - 1);  // the highest refined row that is a map row
    int m = hi - 1;
    if (hi > lo) {
        fill_ring(m);
    }
#pragma unroll 1
    for (; m >= lo && m > last_emit; m--) {  // rows beyond the map: no output
        cp_async_wait<Cfg::kDepth - 1>();
        __syncwarp();
        step(m, use_ring);
        __syncwarp();
        row_end(m, use_ring);
    }
    // m <= last_emit from here on: map row mr = m / rowstep is emitted when phase == 0
    int mr = STEP1 ? m : (m >= 0 ? m / rowstep : 0);
    int phase = STEP1 ? 0 : (m >= 0 ? m - mr * rowstep : 0);
    Fp = Ft + (ptrdiff_t)mr * (ptrdiff_t)w.pitch;
    if (HAS_C) {
        // a row of the warp: 16-byte pieces, group A first; piece q is fetched by lane q % 32
        const unsigned char *src[2];
        uint32_t dsto[2];
        size_t rowb[2];
#pragma unroll
        for (int hq = 0; hq < 2; hq++) {
            const int qq = lane + 32 * hq;
            const bool isA = qq < TcxC::kPiecesA;
            const int qf = qq - TcxC::kPiecesA;
            const int ppc = (isA ? TcxC::kRowA : TcxC::kRowF) / 16;  // pieces per channel
            const int ci = isA ? qq / ppc : qf / ppc, pi = isA ? qq % ppc : qf % ppc;
            const int es = isA ? TcxC::kElemA : TcxC::kElemF, nch = isA ? 3 : 4;
            src[hq] = (isA ? CA : CF) + (((size_t)tz * nch + ci) * c_rows * cpitch + (size_t)(n - lane)) * es + 16 * pi;
            rowb[hq] = (size_t)cpitch * es * (size_t)rowstep;  // the correction sums are indexed by REFINED row
            dsto[hq] = (uint32_t)(isA ? ci * TcxC::kRowA : 3 * TcxC::kRowA + ci * TcxC::kRowF) + 16u * pi;
        }
        const int mr_top = mr;
        // source pointers of the map row to fetch next, destination offset of its ring slot
        const unsigned char *s0 = src[0] + (size_t)mr_top * rowb[0], *s1 = src[1] + (size_t)mr_top * rowb[1];
        const bool two = lane + 32 < TcxC::kPiecesA + TcxC::kPiecesF;
        auto fetch_c = [&](int row, uint32_t slot_off) {  // row: map row
            if (row >= 0 && row * rowstep >= lo) {
                cp_async16(ringC + slot_off + dsto[0], s0);
                if (two) cp_async16(ringC + slot_off + dsto[1], s1);
            }
            s0 -= rowb[0];
            s1 -= rowb[1];
        };
        // (the rows of a segment are fetched kDepth EMITTED rows ahead; a copy is issued at least kDepth refined
        // rows -- commit groups -- before it is read, so the fixed wait count below covers it)
#pragma unroll 1
        for (int r = 0; r < Cfg::kDepth; r++) {
            fetch_c(mr_top - r, (uint32_t)r * TcxC::kRowBytes);
            cp_async_commit();
        }
        uint32_t slot_off = 0;
        auto main_loop = [&](auto ring, auto has_f) {
#pragma unroll 1
            for (; m >= lo; m--) {
                cp_async_wait<Cfg::kDepth - 1>();  // the oldest row in flight has landed (this lane's pieces)
                __syncwarp();                         // ... and every other lane's
                step(m, ring);
                const bool emit = STEP1 || phase == 0;
                if (emit) cell(mr, ringC + slot_off, has_f);
                __syncwarp();  // all lanes have read the slots before they are refilled
                if (emit) {
                    fetch_c(mr - Cfg::kDepth, slot_off);
                    slot_off = slot_off + TcxC::kRowBytes == (uint32_t)Cfg::kDepth * TcxC::kRowBytes ? 0u : slot_off + TcxC::kRowBytes;
                    mr--;
                    if (!STEP1) phase = rowstep;
                }
                if (!STEP1) phase--;
                row_end(m, ring);
            }
        };
        // (use_ring is warp-uniform, Ft kernel-uniform: four compiled loop bodies, no such branch per row)
        if (use_ring) {
            if (Ft) main_loop(WalkRingOn{}, WalkRingOn{});
            else main_loop(WalkRingOn{}, WalkRingOff{});
        } else {
            if (Ft) main_loop(WalkRingOff{}, WalkRingOn{});
            else main_loop(WalkRingOff{}, WalkRingOff{});
        }
    } else {
        auto main_loop = [&](auto ring, auto has_f) {
#pragma unroll 1
            for (; m >= lo; m--) {
                cp_async_wait<Cfg::kDepth - 1>();
                __syncwarp();
                step(m, ring);
                if (STEP1 || phase == 0) {
                    cell(mr, nullptr, has_f);
                    mr--;
                    if (!STEP1) phase = rowstep;
                }
                if (!STEP1) phase--;
                __syncwarp();
                row_end(m, ring);
            }
        };
        // (use_ring is warp-uniform, Ft kernel-uniform: four compiled loop bodies, no such branch per row)
        if (use_ring) {
            if (Ft) main_loop(WalkRingOn{}, WalkRingOn{});
            else main_loop(WalkRingOn{}, WalkRingOff{});
        } else {
            if (Ft) main_loop(WalkRingOff{}, WalkRingOn{});
            else main_loop(WalkRingOff{}, WalkRingOff{});
        }
    }
    cp_async_wait<0>();
    // single-atom cells (Exp.cu's i_t1 == i_t0): a column whose window holds one atom, or the row that starts
    // on the last atom
    {
        const int m_last = (int)numAtoms - 1 - (int)i00;  // refined row
        const int top = min(hi, last_emit + 1);
        const bool any_emit = top > lo && (top - 1) / rowstep * rowstep >= lo;
        const bool degenerate = active && K >= 0 && any_emit &&
                                (K == 0 || (m_last >= lo && m_last < top && m_last % rowstep == 0));
        if (degenerate) atomicOr(&flags[t], TCW_FLAG_DEGENERATE);
    }
    const unsigned long long key = (active && best > -1.0f) ? pack_key(best, (uint32_t)best_m * w.N_tau + n) : 0ull;
    block_atomic_max_key<Cfg::kWarps>(key, &maxkey[t], red);
}
