// tcw_generic.cuh -- generic map kernels: any window geometry, bit-faithful arithmetic.
//
// One thread per (t0,tau) cell.  Index ranges use the reference's uint32 formulas verbatim
// (Rect.cu:21-31, 54-69; Exp.cu:27-65) so wrap-around / clamped / unaligned windows behave
// exactly like the reference; sums are sequential float32 in atom order, which is what the
// reference's running sums produce (Rect.cu:33-40, 75-91), and every operation is an explicit
// round-to-nearest intrinsic so the results are bit-identical to the CPU oracle
// (oracle/tcw_oracle.c, semantics "lal").
//
// These kernels are the correctness path: they serve windows the tiled kernels do not certify
// (tcw_b200.cu: fast_path_certificate) and anchor the parity tests of the tiled kernels.
#pragma once
#include "tcw_common.cuh"
#include "tcw_prep.cuh"

#define TCW_GENERIC_THREADS 256

// XLALFastNegExp: nearest-point lookup in a table of e^{-x} on [0, xmax];
// 0 beyond; (libm exp for negative x never happens here: x >= 0).  Table geometry: ExpLut.
__device__ __forceinline__ double fast_neg_exp_lut(double mx, const ExpLut &lut) {
    if (mx > lut.xmax) return 0.0;
    const uint32_t i0 = __double2uint_rz(__dadd_rn(__dmul_rn(mx, lut.dxinv), 0.5));
    return __ldg(lut.tab + min(i0, lut.len));
}

template <int WTYPE, bool EXACT_EXP>
__global__ void __launch_bounds__(TCW_GENERIC_THREADS)
tcw_map_generic_kernel(const float *__restrict__ X, uint32_t xpad, const TplMeta *__restrict__ meta,
                       int t_base, MapWindow w, const MapWindow *__restrict__ wins, int none_window,
                       IndexGeom g, const ExpLut lut,
                       float *__restrict__ Fmn, unsigned long long *__restrict__ maxkey,
                       uint32_t *__restrict__ flags, uint32_t idx_ntau = 0) {
    // idx_ntau != 0: the launch covers the first w.N_tau columns of a map that has idx_ntau of them (the few-atom
    // windows of an exponential-window map whose other columns the recurrence path computes): the argmax index is
    // that of the full map
    __shared__ unsigned long long red[TCW_GENERIC_THREADS / 32];
    const int tz = blockIdx.z;
    const int t = t_base + tz;
    // per-template window ranges (all of one type and shape): the MCMC case, where every
    // walker carries its own (t0, tau) (mcmc_based_searches.py:3511-3516, core.py:1447-1449)
    if (wins) w = wins[t];
    const uint32_t numAtoms = meta[t].numAtoms;
    const uint32_t t0_data = meta[t].t0_data;
    const size_t cells = (size_t)w.N_t0 * w.N_tau;
    const size_t flat = (size_t)blockIdx.x * TCW_GENERIC_THREADS + threadIdx.x;
    unsigned long long key = 0ull;
    if (flat < cells) {
        const uint32_t m = (uint32_t)(flat / w.N_tau);
        const uint32_t n = (uint32_t)(flat - (size_t)m * w.N_tau);
        // TRANSIENT_NONE: rect window spanning this template's data, 1x1 map (tcw:742-749)
        const uint32_t t0_m = (none_window ? t0_data : w.t0) + m * w.dt0;
        const uint32_t tau_n = (none_window ? numAtoms * g.TAtom : w.tau) + n * w.dtau;
        const uint32_t t1 = t0_m + g.ef * tau_n;
        const uint32_t i_t0 = index_t0(t0_m, t0_data, numAtoms, g);
        const uint32_t i_t1 = index_t1(t1, t0_data, numAtoms, g);
        if (i_t1 == i_t0) atomicOr(&flags[t], TCW_FLAG_DEGENERATE);
        const float *x = X + (size_t)t * TCW_NCH * xpad;
        float S[TCW_NCH];
#pragma unroll
        for (int c = 0; c < TCW_NCH; c++) S[c] = 0.0f;
        if (WTYPE == TCW_WINDOW_RECT) {
            for (uint32_t i = i_t0; i <= i_t1; i++) {
#pragma unroll
                for (int c = 0; c < TCW_NCH; c++) S[c] = __fadd_rn(S[c], __ldg(x + (size_t)c * xpad + i));
            }
        } else {
            for (uint32_t i = i_t0; i <= i_t1; i++) {
                const uint32_t t_i = t0_data + i * g.TAtom;  // Exp.cu:84
                double win = 0.0;
                if (t_i >= t0_m && t_i <= t1) {
                    const double xx = __ddiv_rn((double)(t_i - t0_m), (double)tau_n);
                    win = EXACT_EXP ? exp(-xx) : fast_neg_exp_lut(xx, lut);
                }
                const double win2 = __dmul_rn(win, win);
#pragma unroll
                for (int c = 0; c < TCW_NCH; c++) {
                    const double a = (double)__ldg(x + (size_t)c * xpad + i);
                    // REAL4 accumulator += REAL4 atom * REAL8 window, evaluated in double
                    S[c] = __double2float_rn(__dadd_rn((double)S[c], __dmul_rn(a, c < 3 ? win2 : win)));
                }
            }
        }
        const float F = fstat_faithful(S[0], S[1], S[2], S[3], S[4], S[5], S[6]);
        if (Fmn) Fmn[(size_t)tz * w.N_t0 * w.pitch + (size_t)m * w.pitch + n] = F;
        // maxF starts at -1, strict > (tcw:135-139)
        if (F > -1.0f) key = pack_key(F, idx_ntau ? m * idx_ntau + n : (uint32_t)flat);
    }
    block_atomic_max_key<TCW_GENERIC_THREADS / 32>(key, &maxkey[t], red);
}

// Same arithmetic, ONE WARP per cell: for launches with few cells (the per-walker 1x1 maps of an
// MCMC step, per-segment and cumulative maps, single-cell reads) one thread per cell leaves the
// GPU idle and runs the whole window on a single dependent instruction stream.  Here lane c < 7
// owns channel c -- its float32 sum still runs over the atoms IN ORDER, so the result is
// bit-identical to the thread-per-cell kernel and to the oracle -- and the part of the
// exponential window that does not depend on the running sums (the FP64 division and the table
// lookup of each atom's weight) is done 32 atoms at a time, one per lane, then broadcast by
// shuffles.
#define TCW_GENERIC_WARP_THREADS 128
#define TCW_GENERIC_WARP_MAX_CELLS 16384  // total cells of a launch up to which this kernel is used

template <int WTYPE, bool EXACT_EXP>
__global__ void __launch_bounds__(TCW_GENERIC_WARP_THREADS)
tcw_map_generic_warp_kernel(const float *__restrict__ X, uint32_t xpad, const TplMeta *__restrict__ meta,
                            int t_base, MapWindow w, const MapWindow *__restrict__ wins, int none_window,
                            IndexGeom g, const ExpLut lut,
                            float *__restrict__ Fmn, unsigned long long *__restrict__ maxkey,
                            uint32_t *__restrict__ flags) {
    const int tz = blockIdx.z;
    const int t = t_base + tz;
    if (wins) w = wins[t];
    const uint32_t numAtoms = meta[t].numAtoms;
    const uint32_t t0_data = meta[t].t0_data;
    const size_t cells = (size_t)w.N_t0 * w.N_tau;
    const uint32_t lane = threadIdx.x & 31;
    const size_t flat = (size_t)blockIdx.x * (TCW_GENERIC_WARP_THREADS / 32) + (threadIdx.x >> 5);
    if (flat >= cells) return;  // warp-uniform
    const uint32_t m = (uint32_t)(flat / w.N_tau);
    const uint32_t n = (uint32_t)(flat - (size_t)m * w.N_tau);
    const uint32_t t0_m = (none_window ? t0_data : w.t0) + m * w.dt0;
    const uint32_t tau_n = (none_window ? numAtoms * g.TAtom : w.tau) + n * w.dtau;
    const uint32_t t1 = t0_m + g.ef * tau_n;
    const uint32_t i_t0 = index_t0(t0_m, t0_data, numAtoms, g);
    const uint32_t i_t1 = index_t1(t1, t0_data, numAtoms, g);
    if (lane == 0 && i_t1 == i_t0) atomicOr(&flags[t], TCW_FLAG_DEGENERATE);
    const uint32_t ch = lane < TCW_NCH ? lane : 0u;  // lanes 7..31 shadow channel 0 (results unused)
    const float *xc = X + (size_t)t * TCW_NCH * xpad + (size_t)ch * xpad;
    float S = 0.0f;
    if (WTYPE == TCW_WINDOW_RECT) {
#pragma unroll 4
        for (uint32_t i = i_t0; i <= i_t1; i++) S = __fadd_rn(S, __ldg(xc + i));
    } else if (i_t1 >= i_t0) {
        for (uint32_t base = i_t0; base <= i_t1; base += 32) {
            // weights of atoms base .. base+31, one per lane (Exp.cu:84-90)
            const uint32_t i = base + lane;
            double win = 0.0;
            if (i <= i_t1) {
                const uint32_t t_i = t0_data + i * g.TAtom;
                if (t_i >= t0_m && t_i <= t1) {
                    const double xx = __ddiv_rn((double)(t_i - t0_m), (double)tau_n);
                    win = EXACT_EXP ? exp(-xx) : fast_neg_exp_lut(xx, lut);
                }
            }
            const uint32_t cnt = min(32u, i_t1 - base + 1u);
            // this lane's channel for the 32 atoms: independent loads, all in flight before the
            // (strictly ordered) accumulation starts
            float a[32];
#pragma unroll
            for (int k = 0; k < 32; k++) a[k] = (uint32_t)k < cnt ? __ldg(xc + base + k) : 0.0f;
#pragma unroll
            for (int k = 0; k < 32; k++) {
                if ((uint32_t)k < cnt) {  // warp-uniform
                    const double wk = __shfl_sync(0xffffffffu, win, k);
                    const double wsel = ch < 3 ? __dmul_rn(wk, wk) : wk;
                    // REAL4 accumulator += REAL4 atom * REAL8 window, evaluated in double
                    S = __double2float_rn(__dadd_rn((double)S, __dmul_rn((double)a[k], wsel)));
                }
            }
        }
    }
    float Sc[TCW_NCH];
#pragma unroll
    for (int c = 0; c < TCW_NCH; c++) Sc[c] = __shfl_sync(0xffffffffu, S, c);
    if (lane == 0) {
        const float F = fstat_faithful(Sc[0], Sc[1], Sc[2], Sc[3], Sc[4], Sc[5], Sc[6]);
        if (Fmn) Fmn[(size_t)tz * w.N_t0 * w.pitch + (size_t)m * w.pitch + n] = F;
        if (F > -1.0f) atomicMax(&maxkey[t], pack_key(F, (uint32_t)flat));  // strict > -1 (tcw:135-139)
    }
}
