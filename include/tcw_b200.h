/*
 * tcw_b200.h -- C ABI of the B200-native transient-CW F-statistic map backend.
 *
 * This is the drop-in boundary for ONE hot path of PyFstat: the `tCWFstatMapVersion`
 * backend that turns per-SFT F-stat atoms into the (t0,tau) map F_mn, its max/argmax and
 * the marginalised Bayes factor lnBtSG.  Every entry point cites the reference interface it
 * replaces (paths relative to the PyFstat tree, `tcw` = pyfstat/tcw_fstat_map_funcs.py).
 *
 * Plain C: opaque handle, POD structs, pointers + sizes.  No torch / C++ types.
 * All functions return 0 on success or a negative TCW_E_* code; the message is available
 * from tcw_last_error().  No C++ exception and no exit() ever crosses this boundary.
 *
 * There is NO CPU fallback behind this ABI: every tcw_map_* call runs hand-written sm_100a
 * kernels; if no CUDA device is usable tcw_create() fails with TCW_E_CUDA.
 */
#ifndef TCW_B200_H
#define TCW_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TCW_ABI_VERSION 3

/* lalpulsar transientWindowType_t values used by the reference
 * (tcw:691-697, 742-743, 793-807; pyfstat/core.py:828-840).  lalpulsar is not importable
 * in the build container, so the enum values are carried here as constants. */
#define TCW_WINDOW_NONE 0
#define TCW_WINDOW_RECT 1
#define TCW_WINDOW_EXP  2
#define TCW_WINDOW_LAST 3
/* e-folding truncation of the exponential window: t1 = t0 + 3*tau
 * (pyCUDAkernels/cudaTransientFstatExpWindow.cu:22,53). */
#define TCW_EXP_EFOLDING 3

/* error codes */
#define TCW_OK            0
#define TCW_E_INVALID    -1 /* bad argument (NULL pointer, zero step, unsorted atoms ...)   */
#define TCW_E_WINDOW     -2 /* windowRange.type >= TRANSIENT_LAST (ValueError at tcw:691-697) */
#define TCW_E_CUDA       -3 /* CUDA runtime error / no usable device                        */
#define TCW_E_NOMEM      -4 /* host or device allocation failed                             */
#define TCW_E_DEGENERATE -5 /* a cell has i_t1 == i_t0 (single-atom F-stat): lalpulsar's
                               XLALComputeTransientFstatMap aborts with XLAL_EDOM there
                               (referred to at pyfstat/core.py:2047-2071)                   */
#define TCW_E_STATE      -6 /* call sequence error (e.g. map_resident before upload)        */

/* flags for tcw_map_batch / tcw_map_resident */
#define TCW_WANT_FMN         0x1u /* materialise F_mn (float32, [T][N_t0][N_tau]); without it
                                     F_mn never leaves the chip unless lnBtSG needs a scratch */
#define TCW_WANT_BTSG        0x2u /* also compute lnBtSG, t0_MP, tau_MP (tcw:577-586, 824-828) */
#define TCW_EXP_EXACT        0x4u /* use exact exp() instead of emulating lalpulsar's
                                     XLALFastNegExp lookup table, both for the exponential
                                     window weights and inside lnBtSG (this is what the
                                     reference's pycuda/numpy path does: Exp.cu:88-89,
                                     tcw:210,247,282)                                       */
#define TCW_ALLOW_DEGENERATE 0x8u /* pycuda semantics: single-atom cells silently give the
                                     F=2 fallback / whatever the formula yields, no error    */
#define TCW_FORCE_GENERIC    0x10u /* always use the generic (any window geometry, bit-faithful
                                      sequential float32) kernels instead of the tiled fast
                                      kernels; used by the parity tests                     */
#define TCW_BTSG_TABLE       0x20u /* lnBtSG pass: fetch every term from the XLALFastNegExp table
                                      and sum in FP64 (the very values lalpulsar adds) instead of
                                      recomputing the looked-up entry e^{-i0 dx} on the fly (same
                                      table index, value to ~3e-7 relative).  Implied when a
                                      non-canonical table was installed with tcw_set_exp_lut    */
#define TCW_EXP_DIRECT       0x40u /* exponential window: always the tiled direct sum over the atoms
                                      (Exp.cu:82-102 restructured).  Without it, canonical grids
                                      (rows one atom apart) run the FP64 recurrence down the rows,
                                      plus -- in lookup-table mode -- the tensor-core pass for the
                                      table's deviation from the exact exponential             */

/* Default geometry of the emulated XLALFastNegExp table (lalpulsar/lib/TransientCW_utils.c,
 * not in the reference tree): e^{-x} on [0, XMAX] in LENGTH steps, nearest-point lookup
 * LUT[(UINT4)(x * LENGTH/XMAX + 0.5)], 0 beyond XMAX.  SURVEY A.4-1 records 1/dx = 256
 * (5120 steps); the other recollection on file is 2000 steps (dx = 0.01).  Neither can be
 * verified in the build container, so the geometry is a RUNTIME property of the handle:
 * tcw_set_exp_lut(), or $TCW_EXP_LUT="xmax:length" at tcw_create().  The Python layer measures
 * it from lalpulsar itself whenever lalpulsar is importable (pyfstat_b200/lut_probe.py). */
#define TCW_EXPLUT_DEFAULT_XMAX   20.0
#define TCW_EXPLUT_DEFAULT_LENGTH 5120u

/* One F-stat atom, in the field order of lalpulsar's FstatAtom as read by the reference
 * (tcw:610-617: timestamp u32; a2_alpha, b2_alpha, ab_alpha f32; Fa_alpha, Fb_alpha c8).
 * 32 bytes, no padding. */
typedef struct tcw_atom {
    uint32_t timestamp; /* GPS seconds of the SFT start */
    float a2_alpha;
    float b2_alpha;
    float ab_alpha;
    float Fa_re, Fa_im;
    float Fb_re, Fb_im;
} tcw_atom;

/* lalpulsar transientWindowRange_t as the reference fills it
 * (pyfstat/core.py:828-891, 1447-1449; tests/test_tcw_fstat_map_funcs.py:40-50).
 * All UINT4.  Taken by value semantics: never modified by the library (the reference's
 * pycuda path mutates it for TRANSIENT_NONE, tcw:742-749 -- we do not). */
typedef struct tcw_window_range {
    uint32_t type;
    uint32_t t0, t0Band, dt0;
    uint32_t tau, tauBand, dtau;
} tcw_window_range;

/* Per-template result record = the fields of pyTransientFstatMap (tcw:50-144) that callers
 * read (core.py:1460,1465; grid_based_searches.py:1128-1133), plus the argmax indices that
 * get_maxF_idx() returns (tcw:186-194). */
typedef struct tcw_result {
    double lnBtSG;   /* NaN unless TCW_WANT_BTSG */
    double t0_MP;    /* NaN unless TCW_WANT_BTSG */
    double tau_MP;   /* NaN unless TCW_WANT_BTSG */
    float maxF;      /* max of F (not 2F); -1 if no cell exceeded -1 (tcw:135-139) */
    uint32_t m_ML, n_ML;     /* first-occurrence row-major argmax (np.argmax order, tcw:194) */
    uint32_t t0_ML, tau_ML;  /* t0 + m_ML*dt0, tau + n_ML*dtau (tcw:814-815) */
    uint32_t m_MP, n_MP;     /* argmax of the marginal posteriors (tcw:247-251, 282-286) */
    uint32_t N_t0, N_tau;    /* map shape (tcw:775-780) */
    uint32_t numAtoms;       /* merged atoms on the TAtom grid (tcw:706-709) */
    uint32_t t0_data;        /* first merged timestamp (tcw:725) */
    int32_t status;          /* TCW_OK or TCW_E_DEGENERATE for this template */
    uint32_t path;           /* 0 = generic kernels, 1 = tiled fast kernels, 2 = exp-window recurrence
                                (+ tensor-core correction) (informational) */
    uint32_t reserved;
} tcw_result;

typedef struct tcw_handle tcw_handle;

/* ABI version of the loaded library (== TCW_ABI_VERSION of the header it was built from). */
int tcw_abi_version(void);

/* Device enumeration without opening a context or a handle (replaces drv.Device.count() /
 * drv.Device(n).name() in init_transient_fstat_map_features, tcw:419-432): number of CUDA
 * devices (0 if none / no driver), and the name of device `device`. */
int tcw_device_count(void);
int tcw_device_name_of(int device, char *buf, int buflen);

/* Replaces the pycuda context creation in init_transient_fstat_map_features (tcw:395-484):
 * binds a handle to CUDA device `device`, creates its stream and events.  `device < 0`
 * honours $CUDA_DEVICE like the reference (tcw:434-437, 466-469), default 0. */
int tcw_create(int device, tcw_handle **out);

/* Replaces gpu_context.detach() (pyfstat/core.py:499-516). */
int tcw_destroy(tcw_handle *h);

/* Last error message of this handle (or of the failed tcw_create when h == NULL). */
const char *tcw_last_error(const tcw_handle *h);

/* Device name, as the reference logs it / matches cudaDeviceName against (tcw:419-476). */
int tcw_device_name(const tcw_handle *h, char *buf, int buflen);

/* Geometry (and optionally the contents) of the XLALFastNegExp table this handle emulates for
 * the exponential-window weights (replaces the `lal` path's XLALFastNegExp calls behind
 * lalpulsar.ComputeTransientFstatMap, tcw:571-575) and for the lnBtSG / posterior terms
 * (lalpulsar.ComputeTransientBstat, tcw:578-586).  `table` is NULL (entries exp(-i*xmax/length)
 * computed with the host libm, as XLALCreateExpLUT does) or points to length+1 doubles, e.g.
 * measured from lalpulsar.  Invalidates the cached weight table.  1 <= length <= 2^22. */
int tcw_set_exp_lut(tcw_handle *h, double xmax, uint32_t length, const double *table);
int tcw_get_exp_lut(const tcw_handle *h, double *xmax, uint32_t *length, int *canonical);

/* N_t0Range = floor(t0Band/dt0)+1, N_tauRange = floor(tauBand/dtau)+1 (tcw:775-780).
 * Host-only helper, needs no device. */
int tcw_map_dims(const tcw_window_range *win, uint32_t *N_t0, uint32_t *N_tau);

/* THE hot path.  Replaces fstatmap_versions[...](multiFstatAtoms, windowRange, BtSG)
 * (tcw:320-327, 533) = lalpulsar_compute_transient_fstat_map (tcw:547-587) /
 * pycuda_compute_transient_fstat_map (tcw:656-834) for a BATCH of T templates that share
 * one window range.
 *
 *   atoms    host buffer (pinned or pageable), T*numDet detector vectors, each `atom_stride`
 *            atoms apart: vector (t,X) starts at atoms[(t*numDet + X)*atom_stride].
 *            Timestamps must be non-decreasing within a vector (atoms falling into the same
 *            TAtom bin are summed, as XLALmergeMultiFstatAtomsBinned does).
 *   n_atoms  [T*numDet] number of valid atoms of each vector (0 <= n <= atom_stride; every
 *            template needs at least one atom in some detector)
 *   TAtom    atom duration = multiFstatAtoms.data[0].TAtom (tcw:704), same for all detectors
 *   win      window range (TRANSIENT_NONE is replaced internally by a rect window spanning
 *            the data, tcw:742-749, on a copy)
 *   F_mn_out NULL or host buffer of T*N_t0*N_tau float32 (required with TCW_WANT_FMN);
 *            holds F, not 2F, row-major [t][m over t0][n over tau] (tcw:67-72)
 *   results  [T] records
 *
 * Synchronous: H2D of the atoms, kernels, D2H of the records (and of F_mn if asked).
 * A per-template degenerate window sets results[t].status = TCW_E_DEGENERATE and the call
 * returns TCW_E_DEGENERATE (all other templates are still valid). */
int tcw_map_batch(tcw_handle *h, const tcw_atom *atoms, const uint32_t *n_atoms,
                  uint32_t atom_stride, uint32_t TAtom, int T, int numDet,
                  const tcw_window_range *win, uint32_t flags, float *F_mn_out,
                  tcw_result *results);

/* Same, with ONE WINDOW RANGE PER TEMPLATE (all of the same type and map shape, typically
 * 1x1): the MCMC case, where every walker of a sampler step carries its own transient
 * start time and duration and each likelihood call evaluates a single cell
 * (pyfstat/mcmc_based_searches.py:3448-3466, 3511-3516: windowRange.t0 = int(tstart),
 * windowRange.tau = int(tend - tstart), pyfstat/core.py:1447-1449).  Served by the generic
 * kernels. */
int tcw_map_batch_windows(tcw_handle *h, const tcw_atom *atoms, const uint32_t *n_atoms,
                          uint32_t atom_stride, uint32_t TAtom, int T, int numDet,
                          const tcw_window_range *wins /* [T] */, uint32_t flags, float *F_mn_out,
                          tcw_result *results);

/* Asynchronous split of tcw_map_batch, for search drivers whose host side has work of its own
 * between batches (in PyFstat: lalpulsar.ComputeFstat producing the next batch's atoms,
 * core.py:1359-1365): tcw_submit() enqueues the H2D copies (chunked on a copy stream) and all
 * kernels and returns without waiting; tcw_wait() waits for them and copies the records (and
 * F_mn, when TCW_WANT_FMN was submitted) to the host.  One batch in flight per handle: a second
 * tcw_submit (or any other map / upload call) before tcw_wait returns TCW_E_STATE.  `atoms` and
 * `n_atoms` must stay valid and unchanged until tcw_wait returns (pinned memory for a truly
 * asynchronous copy). */
int tcw_submit(tcw_handle *h, const tcw_atom *atoms, const uint32_t *n_atoms,
               uint32_t atom_stride, uint32_t TAtom, int T, int numDet,
               const tcw_window_range *win, uint32_t flags);
int tcw_wait(tcw_handle *h, float *F_mn_out /* NULL unless TCW_WANT_FMN */, tcw_result *results);

/* Device-resident variant used by batched search drivers and by bench.py's kernel-only
 * timing: upload once, then run any number of windows on the resident atoms. */
int tcw_upload_atoms(tcw_handle *h, const tcw_atom *atoms, const uint32_t *n_atoms,
                     uint32_t atom_stride, uint32_t TAtom, int T, int numDet);
/* Launches the kernels for the resident batch on the handle's stream and returns without
 * synchronising.  F_mn (if TCW_WANT_FMN) stays in device memory. */
int tcw_map_resident(tcw_handle *h, const tcw_window_range *win, uint32_t flags);
/* Waits for the stream, copies the T records to the host. */
int tcw_fetch_results(tcw_handle *h, tcw_result *results);
/* Device address of the T result records of the last map (valid until the next map / upload call on
 * this handle; final once tcw_fetch_results / tcw_map_batch / tcw_wait has returned).  Lets a
 * multi-GPU driver hand the records to a collective (NCCL all_gather) without a host round trip. */
int tcw_results_device(tcw_handle *h, void **ptr, uint64_t *n_records);
/* Copies F_mn of resident template t (after a TCW_WANT_FMN run) to the host. */
int tcw_fetch_fmn(tcw_handle *h, int t, float *F_mn_out);
/* Copies the merged (binned) atoms of resident template t as 7 float32 channel arrays of
 * length numAtoms, order a2,b2,ab,Fa_re,Fa_im,Fb_re,Fb_im = the columns of the reference's
 * atomsInputMatrix (tcw:711-721).  For tests of the on-device detector merge (tcw:706). */
int tcw_fetch_merged(tcw_handle *h, int t, float *channels7xN, uint32_t capacityN);
int tcw_synchronize(tcw_handle *h);

/* CUDA-event timing on the handle's own stream (torch.cuda.Event only sees torch's
 * current stream, so bench.py times through these). */
int tcw_timer_start(tcw_handle *h);
int tcw_timer_stop(tcw_handle *h, float *milliseconds); /* synchronises the stop event */
/* Timings of the last tcw_map_resident call, per stage, from events recorded on the
 * stream: [0] prep (merge+scan), [1] weight table, [2] map kernel, [3] BtSG pass,
 * [4] finalize.  Synchronises. */
int tcw_last_stage_ms(tcw_handle *h, float ms[5]);
/* The map stage of the last call, split for the exponential window's recurrence path (tcw_exp_rec.cuh):
 * [0] operand preparation (scales + chunked atoms), [1] tensor-core pass, [2] the walk (recurrence + F-stat
 * epilogue).  All zero when the last map took another path.  Synchronises. */
int tcw_last_exp_stage_ms(tcw_handle *h, float ms[3]);
/* Kernels launched by this handle since creation (for bench.py's gpu_launches). */
uint64_t tcw_launch_count(const tcw_handle *h);
/* Writes a buffer larger than L2 (256 MiB) on the handle's stream. */
int tcw_flush_l2(tcw_handle *h);
/* FP32-FMA / FP64-add peak microbenchmarks on this device (TFLOP/s, counting FMA = 2 flop,
 * DADD = 1 flop): the roofline denominators MEASURED_PEAKS.json does not carry. */
int tcw_microbench(tcw_handle *h, double *ffma_tflops, double *dadd_tflops);
/* Same for Blackwell's packed FP32 FMA (fma.rn.f32x2 / SASS FFMA2), counting 4 flop per lane-instruction. */
int tcw_microbench_ffma2(tcw_handle *h, double *ffma2_tflops);
/* Pinned host memory helpers (cudaHostAlloc / cudaFreeHost). */
void *tcw_host_alloc(uint64_t bytes);
void tcw_host_free(void *p);

/* Host-only: the uint32 index arithmetic of the reference for one cell
 * (cudaTransientFstatRectWindow.cu:21-31, 54-69; ...ExpWindow.cu:27-65), exactly as the
 * kernels compute it (magic-number division included).  For bit-exactness tests. */
int tcw_cell_index_range(uint32_t window_type, uint32_t t0_m, uint32_t tau_n,
                         uint32_t t0_data, uint32_t TAtom, uint32_t numAtoms,
                         uint32_t *i_t0, uint32_t *i_t1);

#ifdef __cplusplus
}
#endif
#endif /* TCW_B200_H */
