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

#define TCW_PREP_THREADS 512
#define TCW_FLAG_UNSORTED 0x1u
#define TCW_FLAG_DEGENERATE 0x2u

__global__ void __launch_bounds__(TCW_PREP_THREADS)
tcw_prep_kernel(const tcw_atom *__restrict__ atoms, const uint32_t *__restrict__ n_atoms,
                const TplMeta *__restrict__ meta, int t_base, int numDet, uint32_t stride, uint32_t TAtom,
                MagicDiv md, float *__restrict__ X, float *__restrict__ X8, uint32_t xpad,
                double *__restrict__ P, uint32_t ppad, uint32_t *__restrict__ flags) {
    const int t = t_base + blockIdx.x;
    const uint32_t N = meta[t].numAtoms;
    const uint32_t tMin = meta[t].t0_data;
    const int tid = threadIdx.x;
    float *Xt = X + (size_t)t * TCW_NCH * xpad;
    double *Pt = P ? P + (size_t)t * TCW_NCH * ppad : nullptr;
    const tcw_atom *At = atoms + (size_t)t * numDet * stride;
    const uint32_t *nt = n_atoms + (size_t)t * numDet;

    // sortedness check (non-decreasing timestamps per detector; atoms sharing a bin are summed in
    // order, as XLALmergeMultiFstatAtomsBinned accumulates them)
    bool unsorted = false;
    for (int Xd = 0; Xd < numDet; Xd++) {
        const tcw_atom *a = At + (size_t)Xd * stride;
        const uint32_t n = nt[Xd];
        for (uint32_t i = tid; i + 1 < n; i += blockDim.x)
            unsorted |= a[i].timestamp > a[i + 1].timestamp;
    }
    if (unsorted) atomicOr(&flags[t], TCW_FLAG_UNSORTED);

    // gather per bin: binary search the first atom of each detector that falls into bin j
    for (uint32_t j = tid; j < xpad; j += blockDim.x) {
        float s[TCW_NCH];
#pragma unroll
        for (int c = 0; c < TCW_NCH; c++) s[c] = 0.0f;
        if (j < N) {
            const uint32_t lo_t = tMin + j * TAtom;
            for (int Xd = 0; Xd < numDet; Xd++) {
                const tcw_atom *a = At + (size_t)Xd * stride;
                const uint32_t n = nt[Xd];
                uint32_t lo = 0, hi = n;  // lower_bound(timestamp >= lo_t)
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (a[mid].timestamp < lo_t) lo = mid + 1;
                    else hi = mid;
                }
                for (uint32_t i = lo; i < n; i++) {
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
        }
#pragma unroll
        for (int c = 0; c < TCW_NCH; c++) Xt[(size_t)c * xpad + j] = s[c];
        // atom-interleaved copy for the exponential-window kernel: one 32-byte record per atom
        if (X8) {
            float4 *x8 = reinterpret_cast<float4 *>(X8 + ((size_t)t * xpad + j) * 8);
            x8[0] = make_float4(s[0], s[1], s[2], s[3]);
            x8[1] = make_float4(s[4], s[5], s[6], 0.0f);
        }
    }
    if (!P) return;  // prefix sums are only needed by the rectangular-window kernel
    __syncthreads();

    // FP64 exclusive prefix scan per channel: each thread owns a contiguous chunk, chunk
    // totals are scanned with warp shuffles + one inter-warp carry through shared memory.
    __shared__ double warp_tot[TCW_PREP_THREADS / 32];
    const uint32_t L = (N + blockDim.x - 1) / blockDim.x;
    const uint32_t beg = min((uint32_t)tid * L, N);
    const uint32_t end = min(beg + L, N);
    const int lane = tid & 31, warp = tid >> 5;
    for (int c = 0; c < TCW_NCH; c++) {
        const float *xc = Xt + (size_t)c * xpad;
        double *pc = Pt + (size_t)c * ppad;
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
            double w = lane < (TCW_PREP_THREADS / 32) ? warp_tot[lane] : 0.0;
            double wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double v = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += v;
            }
            if (lane < (TCW_PREP_THREADS / 32)) warp_tot[lane] = wi - w;  // exclusive carry
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
        for (uint32_t i = N + 1 + tid; i < ppad; i += blockDim.x) pc[i] = 0.0;
        __syncthreads();  // warp_tot is reused by the next channel
    }
}
