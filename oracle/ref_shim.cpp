/*
 * ref_shim.cpp -- runs the REFERENCE'S OWN map kernels on the CPU.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/tcw_oracle.c header).
 *
 * The two native sources of the reference's path,
 *   pyfstat/pyCUDAkernels/cudaTransientFstatRectWindow.cu
 *   pyfstat/pyCUDAkernels/cudaTransientFstatExpWindow.cu
 * are self-contained `__global__` functions in plain CUDA C.  This shim #includes them
 * WHERE THEY LIE under /root/reference (path given by -DREF_RECT_CU / -DREF_EXP_CU in
 * oracle/Makefile; nothing is copied into this repo), defines the handful of CUDA builtins
 * they use as host variables, and replays the launch geometry of their host wrappers
 * (tcw_fstat_map_funcs.py:878-898 rect, :959-979 exp) thread by thread.
 *
 * The result, oracle/_ref/libtcw_ref.so, pins oracle/tcw_oracle.c (semantics "pycuda")
 * against real reference code.  It exists only where /root/reference exists (the build
 * container); the committed fixtures under tests/golden/ carry its outputs to the GPU box.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

struct shim_dim3 {
    unsigned int x, y, z;
};
static shim_dim3 blockDim, blockIdx, threadIdx;

#define __global__ static

#include REF_RECT_CU
#include REF_EXP_CU

extern "C" {

/* Launch as pycuda_compute_transient_fstat_map_rect does (tcw:878-898).  `Fmn` must hold
 * rows_alloc*N_tauRange floats with rows_alloc >= ref_rect_rows_needed(): the kernel's guard
 * is `m < N_tauRange` (Rect.cu:48), so surplus threads of the last block write rows beyond
 * N_t0Range. */
unsigned int ref_rect_rows_needed(unsigned int N_t0Range) {
    unsigned int blockRows = N_t0Range < 1024 ? N_t0Range : 1024;
    unsigned int gridRows = (N_t0Range + blockRows - 1) / blockRows;
    return gridRows * blockRows;
}

void ref_rect(float *input, unsigned int numAtoms, unsigned int TAtom, unsigned int t0_data,
              unsigned int win_t0, unsigned int win_dt0, unsigned int win_tau,
              unsigned int win_dtau, unsigned int N_t0Range, unsigned int N_tauRange, float *Fmn) {
    unsigned int blockRows = N_t0Range < 1024 ? N_t0Range : 1024;
    unsigned int gridRows = (N_t0Range + blockRows - 1) / blockRows;
    blockDim.x = blockRows;
    blockDim.y = 1;
    blockDim.z = 1;
    for (unsigned int b = 0; b < gridRows; b++) {
        for (unsigned int t = 0; t < blockRows; t++) {
            blockIdx.x = b;
            blockIdx.y = 0;
            blockIdx.z = 0;
            threadIdx.x = t;
            threadIdx.y = 0;
            threadIdx.z = 0;
            cudaTransientFstatRectWindow(input, numAtoms, TAtom, t0_data, win_t0, win_dt0, win_tau,
                                         win_dtau, N_tauRange, Fmn);
        }
    }
}

/* Launch as pycuda_compute_transient_fstat_map_exp does (tcw:959-979). */
void ref_exp(float *input, unsigned int numAtoms, unsigned int TAtom, unsigned int t0_data,
             unsigned int win_t0, unsigned int win_dt0, unsigned int win_tau, unsigned int win_dtau,
             unsigned int N_t0Range, unsigned int N_tauRange, float *Fmn) {
    unsigned int blockRows = N_t0Range < 32 ? N_t0Range : 32;
    unsigned int blockCols = N_tauRange < 32 ? N_tauRange : 32;
    unsigned int gridRows = (N_t0Range + blockRows - 1) / blockRows;
    unsigned int gridCols = (N_tauRange + blockCols - 1) / blockCols;
    blockDim.x = blockRows;
    blockDim.y = blockCols;
    blockDim.z = 1;
    for (unsigned int bx = 0; bx < gridRows; bx++)
        for (unsigned int by = 0; by < gridCols; by++)
            for (unsigned int tx = 0; tx < blockRows; tx++)
                for (unsigned int ty = 0; ty < blockCols; ty++) {
                    blockIdx.x = bx;
                    blockIdx.y = by;
                    blockIdx.z = 0;
                    threadIdx.x = tx;
                    threadIdx.y = ty;
                    threadIdx.z = 0;
                    cudaTransientFstatExpWindow(input, numAtoms, TAtom, t0_data, win_t0, win_dt0,
                                                win_tau, win_dtau, N_t0Range, N_tauRange, Fmn);
                }
}

} /* extern "C" */
