// tcw_prep.cuh -- on-device detector merge + channel transpose + FP64 prefix scan.
//
// Replaces, per template, what the reference does on the host before its kernel launch:
//   lalpulsar.mergeMultiFstatAtomsBinned(multiFstatAtoms, TAtom)        (tcw:706)
//   reshape_FstatAtomsVector + column_stack -> [numAtoms x 7] float32   (tcw:710-721)
// Merge rule (recalled from XLALmergeMultiFstatAtomsBinned): bins of width TAtom
// starting at the earliest first timestamp tMin; atom (X,i) goes to bin
// floor((t_Xi - tMin)/TAtom); the 7 quantities are summed in float32 in detector order;
// empty bins stay zero.  One CTA per template.
//
// Outputs (zero padded; X8 and P are optional -- nullptr skips them):
//   X[t][c][xpad] float32  merged channel c (SoA)  -- generic kernels, tcw_fetch_merged
//   X8[t][xpad][8] float32 merged atoms, channel-interleaved (7 + pad) -- exp window kernel
//   P[t][c][ppad] float64  exclusive prefix: P[i] = sum_{j<i} X[j], P[0] = 0, i <= numAtoms
//                          -- rect window: every (t0,tau) cell is one FP64 difference
#pragma once
#include "tcw_common.cuh"

#define TCW_PREP_THREADS 256
#define TCW_FLAG_UNSORTED 0x1u
#define TCW_FLAG_DEGENERATE 0x2u

// Round 1 ran merge + transpose + the 7 prefix scans in ONE CTA per template (47 us for 64 x 60-d
// templates -- 64 of 148 SMs busy, a 12-step dependent binary search per bin, 14 block barriers).
// Now two kernels, each with many more CTAs:
//   tcw_prep_merge_kernel  grid (bins / 256, T): one thread per TAtom bin
//   tcw_prep_scan_kernel   grid (7, T): one CTA per (channel, template), rect window only

// first atom of detector vector `a` (n atoms, non-decreasing timestamps) with timestamp >= lo_t.
// SFTs of one detector do not overlap, so atom i starts at or after bin i of its detector: the guess
// i = j is right for gap-free data and an upper bound otherwise; verified, else binary search.
__device__ __forceinline__ uint32_t prep_lower_bound(const tcw_atom *__restrict__ a, uint32_t n, uint32_t j,
                                                     uint32_t lo_t) {
    uint32_t hi = n;
    if (j < n) {
        const uint32_t tj = __ldg(&a[j].timestamp);
        const uint32_t tp = j ? __ldg(&a[j - 1].timestamp) : 0u;
        if (tj >= lo_t) {
            if (j == 0 || tp < lo_t) return j;  // the guess is the lower bound
            hi = j;                             // lower bound lies before the guess
        } else {
            uint32_t lo = j + 1;                // (overlapping atoms) lies after it
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (__ldg(&a[mid].timestamp) < lo_t) lo = mid + 1;
                else hi = mid;
            }
            return lo;
        }
    }
    uint32_t lo = 0;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(&a[mid].timestamp) < lo_t) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(TCW_PREP_THREADS)
tcw_prep_merge_kernel(const tcw_atom *__restrict__ atoms, const uint32_t *__restrict__ n_atoms,
                      const TplMeta *__restrict__ meta, int t_base, int numDet, uint32_t stride, uint32_t TAtom,
                      MagicDiv md, float *__restrict__ X, float *__restrict__ X8, uint32_t xpad,
                      uint32_t *__restrict__ flags) {
    const int t = t_base + blockIdx.y;
    const uint32_t N = meta[t].numAtoms;
    const uint32_t tMin = meta[t].t0_data;
    const uint32_t j = blockIdx.x * TCW_PREP_THREADS + threadIdx.x;  // bin
    if (j >= xpad) return;
    float *Xt = X + (size_t)t * TCW_NCH * xpad;
    const tcw_atom *At = atoms + (size_t)t * numDet * stride;
    const uint32_t *nt = n_atoms + (size_t)t * numDet;

    float s[TCW_NCH];
#pragma unroll
    for (int c = 0; c < TCW_NCH; c++) s[c] = 0.0f;
    bool unsorted = false;
    for (int Xd = 0; Xd < numDet; Xd++) {
        const tcw_atom *a = At + (size_t)Xd * stride;
        const uint32_t n = nt[Xd];
        // sortedness check (non-decreasing timestamps per detector; atoms sharing a bin are summed in
        // order, as XLALmergeMultiFstatAtomsBinned accumulates them): thread j looks at the pair (j, j+1)
        for (uint32_t i = j; i + 1 < n; i += xpad)
            unsorted |= __ldg(&a[i].timestamp) > __ldg(&a[i + 1].timestamp);
        if (j >= N) continue;
        const uint32_t lo_t = tMin + j * TAtom;
        for (uint32_t i = prep_lower_bound(a, n, j, lo_t); i < n; i++) {
            // vectorised 32-byte atom read: {ts,a2,b2,ab} {Fa_re,Fa_im,Fb_re,Fb_im}
            const uint4 q0 = __ldg(reinterpret_cast<const uint4 *>(a + i));
            if (magic_div(q0.x - tMin, md) != j) break;
            const float4 q1 = __ldg(reinterpret_cast<const float4 *>(a + i) + 1);
            s[0] = __fadd_rn(s[0], __uint_as_float(q0.y));
            s[1] = __fadd_rn(s[1], __uint_as_float(q0.z));
            s[2] = __fadd_rn(s[2], __uint_as_float(q0.w));
            s[3] = __fadd_rn(s[3], q1.x);
            s[4] = __fadd_rn(s[4], q1.y);
            s[5] = __fadd_rn(s[5], q1.z);
            s[6] = __fadd_rn(s[6], q1.w);
        }
    }
    if (unsorted) atomicOr(&flags[t], TCW_FLAG_UNSORTED);
#pragma unroll
    for (int c = 0; c < TCW_NCH; c++) Xt[(size_t)c * xpad + j] = s[c];
    // atom-interleaved copy for the exponential-window kernel: one 32-byte record per atom
    if (X8) {
        float4 *x8 = reinterpret_cast<float4 *>(X8 + ((size_t)t * xpad + j) * 8);
        x8[0] = make_float4(s[0], s[1], s[2], s[3]);
        x8[1] = make_float4(s[4], s[5], s[6], 0.0f);
    }
}

// FP64 exclusive prefix scan of one channel of one template: each thread owns a contiguous chunk,
// chunk totals are scanned with warp shuffles + one inter-warp carry through shared memory.
__global__ void __launch_bounds__(TCW_PREP_THREADS)
tcw_prep_scan_kernel(const float *__restrict__ X, uint32_t xpad, const TplMeta *__restrict__ meta, int t_base,
                     double *__restrict__ P, uint32_t ppad) {
    const int t = t_base + blockIdx.y;
    const int c = blockIdx.x;
    const uint32_t N = meta[t].numAtoms;
    const int tid = threadIdx.x;
    const float *xc = X + ((size_t)t * TCW_NCH + c) * xpad;
    double *pc = P + ((size_t)t * TCW_NCH + c) * ppad;
    __shared__ double warp_tot[TCW_PREP_THREADS / 32];
    const uint32_t L = (N + TCW_PREP_THREADS - 1) / TCW_PREP_THREADS;
    const uint32_t beg = min((uint32_t)tid * L, N);
    const uint32_t end = min(beg + L, N);
    const int lane = tid & 31, warp = tid >> 5;
    double tot = 0.0;
    for (uint32_t i = beg; i < end; i++) tot += (double)xc[i];
    double incl = tot;  // inclusive warp scan of chunk totals
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        double wv = lane < (TCW_PREP_THREADS / 32) ? warp_tot[lane] : 0.0;
        double wi = wv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double v = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += v;
        }
        if (lane < (TCW_PREP_THREADS / 32)) warp_tot[lane] = wi - wv;  // exclusive carry
    }
    __syncthreads();
    double run = warp_tot[warp] + (incl - tot);  // exclusive prefix at `beg`
    for (uint32_t i = beg; i < end; i++) {
        pc[i] = run;
        run += (double)xc[i];
    }
    if (end == N && beg < N) pc[N] = run;  // the owner of the last element closes P
    if (N == 0 && tid == 0) pc[0] = 0.0;
    // zero padding beyond N (TMA tiles may read a few entries past the end)
    for (uint32_t i = N + 1 + tid; i < ppad; i += TCW_PREP_THREADS) pc[i] = 0.0;
}
