/*
 * tcw_oracle.c -- CPU restatement of the reference's transient F-stat map path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (pyfstat_b200/) may import, link or
 * execute this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, as the checker and as the CPU timing baseline.
 *
 * PARITY STATUS: **parity unpinned against lalpulsar**.  The reference's `lal` backend is a
 * thin wrapper (pyfstat/tcw_fstat_map_funcs.py:571-586) around the third-party C library
 * lalsuite/lalpulsar (setup.py:45 `lalsuite[lalpulsar]>=7.13`; lalpulsar/lib/TransientCW_utils.c),
 * which is NOT in /root/reference and not installable here.  What IS pinned:
 *   - the per-cell arithmetic and index ranges against the reference's own in-tree CUDA ports
 *     of that C code (the two .cu files under pyCUDAkernels/), compiled for the host into oracle/_ref by
 *     oracle/Makefile and compared bit-for-bit (semantics = TCW_SEM_PYCUDA), and
 *   - lnBtSG / t0_MP / tau_MP (exact-exp flavour) and the F_mn text format against the
 *     reference's own Python class pyTransientFstatMap (tests/golden/).
 * The `lal`-only behaviours (TCW_SEM_LAL) are restated from the published lalsuite algorithm
 * as recalled; each is behind a switch and listed in DESIGN.md:
 *   L1 exponential weights and the lnBtSG terms come from the lookup table XLALFastNegExp:
 *      e^{-x} tabulated on [0, xmax] in `length` steps, nearest point
 *      `LUT[(UINT4)(x*length/xmax + 0.5)]`, 0 for x > xmax, libm exp for x < 0; xmax and length
 *      are run-time settings (default 20 / 5120 = SURVEY A.4-1; the other recollection is 20 / 2000);
 *   L2 REAL4 accumulators; in the exponential case each term is `REAL4 * REAL8 window value`
 *      evaluated in double and added to the REAL4 accumulator;
 *   L3 a cell with i_t1 == i_t0 aborts the map (XLAL_EDOM);
 *   L4 detector merge XLALmergeMultiFstatAtomsBinned: bins of width TAtom from the earliest
 *      first timestamp, float32 sums in detector order.
 *
 * Everything else follows in-tree code, cited per function.
 *
 * Build: gcc -O2 -fno-fast-math -ffp-contract=off -fopenmp -shared -fPIC (see Makefile).
 * -ffp-contract=off matters: results are compared bit-for-bit with CUDA code written with
 * explicit _rn intrinsics.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define WIN_NONE 0
#define WIN_RECT 1
#define WIN_EXP 2
#define WIN_LAST 3
#define EXP_EFOLDING 3 /* Exp.cu:22 */

#define SEM_LAL 0    /* LUT weights, REAL4 acc with double products, degenerate -> error */
#define SEM_PYCUDA 1 /* float expf weights, float products, no degenerate check (Exp.cu) */

#define ERR_INVALID -1
#define ERR_WINDOW -2
#define ERR_DEGENERATE -5

typedef struct {
    uint32_t timestamp;
    float a2, b2, ab;
    float Fa_re, Fa_im, Fb_re, Fb_im;
} atom_t;

typedef struct {
    uint32_t type, t0, t0Band, dt0, tau, tauBand, dtau;
} window_range_t;

typedef struct {
    double lnBtSG, t0_MP, tau_MP;
    double maxF; /* value of a REAL4 F held in a REAL8 like transientFstatMap_t.maxF */
    uint32_t m_ML, n_ML, t0_ML, tau_ML, m_MP, n_MP, N_t0, N_tau, numAtoms, t0_data;
    int32_t status;
} oracle_result_t;

/* ---- XLALFastNegExp (L1; lalpulsar/lib/TransientCW_utils.c, restated) ----------------------
 * Table of e^{-x} on [0, xmax] with `length` steps (length + 1 entries, entry i = exp(-(i*dx)),
 * dx = xmax/length), nearest-point lookup LUT[(UINT4)(mx * (length/xmax) + 0.5)], 0 for
 * mx > xmax, libm exp for mx < 0.  The geometry is a RUNTIME setting (oracle_set_exp_lut):
 * SURVEY A.4-1 records xmax = 20, 1/dx = 256 (length 5120), the default here; the other
 * recollection on file is length 2000 (dx = 0.01).  Neither is verifiable without lalsuite. */
static double explut_xmax = 20.0;
static uint32_t explut_length = 5120;
static double *expLUT = NULL;
static uint32_t expLUT_entries = 0;

static void create_exp_lut(void) {
    double dx = explut_xmax / explut_length;
    double *t = (double *)malloc(((size_t)explut_length + 1) * sizeof(double));
    for (uint32_t i = 0; i <= explut_length; i++) t[i] = exp(-(i * dx));
    expLUT = t;
    expLUT_entries = explut_length + 1;
}

/* not thread-safe: call before any map (tests / bench do so from the main thread) */
int oracle_set_exp_lut(double xmax, uint32_t length) {
    if (!(xmax > 0) || length < 1) return ERR_INVALID;
    free(expLUT);
    expLUT = NULL;
    explut_xmax = xmax;
    explut_length = length;
    create_exp_lut();
    return 0;
}

void oracle_get_exp_lut(double *xmax, uint32_t *length) {
    *xmax = explut_xmax;
    *length = explut_length;
}

static inline double fast_neg_exp(double mx) {
    if (mx > explut_xmax) return 0.0;
    if (mx < 0) return exp(-mx);
    uint32_t i0 = (uint32_t)(mx * ((explut_length) / (explut_xmax)) + 0.5);
    return expLUT[i0];
}

double oracle_fast_neg_exp(double mx) {
    if (!expLUT) create_exp_lut();
    return fast_neg_exp(mx);
}

/* copy of the table for the tests (so the CUDA side can be checked entry by entry) */
int oracle_exp_lut(double *out, int capacity) {
    if (!expLUT) create_exp_lut();
    if (capacity < (int)expLUT_entries) return ERR_INVALID;
    memcpy(out, expLUT, (size_t)expLUT_entries * sizeof(double));
    return (int)expLUT_entries;
}

/* ---- detector merge (tcw:702-709; rule recalled, L4) ------------------------------------ */
/* atoms of detector X: atoms[X*stride .. X*stride + n_atoms[X]).  Output: out[numOut]
 * (capacity checked).  Filled bins get timestamp tMin + j*TAtom, empty bins stay all-zero
 * (timestamp 0), as SURVEY A.1/A.4-3 describe. */
int oracle_merge_binned(const atom_t *atoms, const uint32_t *n_atoms, int numDet,
                        uint32_t stride, uint32_t TAtom, atom_t *out, uint32_t capacity,
                        uint32_t *numOut) {
    if (!atoms || !n_atoms || numDet < 1 || TAtom == 0) return ERR_INVALID;
    uint32_t tMin = 0xffffffffu, tMax = 0;
    int any = 0;
    for (int X = 0; X < numDet; X++) {
        if (n_atoms[X] > stride) return ERR_INVALID;
        if (n_atoms[X] == 0) continue; /* a detector without atoms contributes nothing */
        any = 1;
        const atom_t *a = atoms + (size_t)X * stride;
        if (a[0].timestamp < tMin) tMin = a[0].timestamp;
        if (a[n_atoms[X] - 1].timestamp > tMax) tMax = a[n_atoms[X] - 1].timestamp;
    }
    if (!any || tMax < tMin) return ERR_INVALID;
    uint32_t N = (uint32_t)floor(1.0 * (tMax - tMin) / TAtom) + 1;
    *numOut = N;
    if (N > capacity) return ERR_INVALID;
    memset(out, 0, (size_t)N * sizeof(atom_t));
    for (int X = 0; X < numDet; X++) {
        const atom_t *a = atoms + (size_t)X * stride;
        for (uint32_t i = 0; i < n_atoms[X]; i++) {
            uint32_t j = (uint32_t)floor(1.0 * (a[i].timestamp - tMin) / TAtom);
            if (j >= N) return ERR_INVALID; /* unsorted input */
            atom_t *d = &out[j];
            d->timestamp = tMin + j * TAtom;
            d->a2 += a[i].a2;
            d->b2 += a[i].b2;
            d->ab += a[i].ab;
            d->Fa_re += a[i].Fa_re;
            d->Fa_im += a[i].Fa_im;
            d->Fb_re += a[i].Fb_re;
            d->Fb_im += a[i].Fb_im;
        }
    }
    return 0;
}

/* ---- index ranges (Rect.cu:21-31, 54-69; Exp.cu:27-65) ---------------------------------- */
/* all uint32 arithmetic, signed re-interpretation only for the `< 0` clamp */
void oracle_index_range(uint32_t type, uint32_t t0_m, uint32_t tau_n, uint32_t t0_data,
                        uint32_t TAtom, uint32_t numAtoms, uint32_t *i_t0, uint32_t *i_t1,
                        uint32_t *t1_out) {
    uint32_t TAtomHalf = TAtom / 2;
    int32_t i_tmp = (int32_t)((t0_m - t0_data + TAtomHalf) / TAtom);
    if (i_tmp < 0) i_tmp = 0;
    uint32_t a = (uint32_t)i_tmp;
    if (a >= numAtoms) a = numAtoms - 1;
    uint32_t t1 = (type == WIN_EXP) ? t0_m + EXP_EFOLDING * tau_n : t0_m + tau_n;
    i_tmp = (int32_t)((t1 - t0_data + TAtomHalf) / TAtom - 1);
    if (i_tmp < 0) i_tmp = 0;
    uint32_t b = (uint32_t)i_tmp;
    if (b >= numAtoms) b = numAtoms - 1;
    *i_t0 = a;
    *i_t1 = b;
    if (t1_out) *t1_out = t1;
}

/* ---- guarded F-stat epilogue (Rect.cu:98-118 == Exp.cu:109-129) -------------------------
 * The double literals (4.0, 1.0, 2.0) promote parts of the expression exactly as they do in
 * the reference kernels; do not "simplify". */
static inline float fstat_from_sums(float Ad, float Bd, float Cd, float Fa_re, float Fa_im,
                                    float Fb_re, float Fb_im) {
    float sumAB = Ad + Bd;
    float diffAB = Ad - Bd;
    float disc = sqrt(diffAB * diffAB + 4.0 * Cd * Cd);
    float denom = sumAB - disc;
    float cond = (denom > 0) ? ((sumAB + disc) / denom) : INFINITY;
    float DdInv = 0.0f;
    if (cond < 1e4) {
        DdInv = 1.0 / (Ad * Bd - Cd * Cd);
    }
    float F = 2;
    if (DdInv > 0) {
        F = DdInv * (Bd * (Fa_re * Fa_re + Fa_im * Fa_im) + Ad * (Fb_re * Fb_re + Fb_im * Fb_im) -
                     2.0 * Cd * (Fa_re * Fb_re + Fa_im * Fb_im));
    }
    return F;
}

float oracle_fstat_from_sums(float Ad, float Bd, float Cd, float Fa_re, float Fa_im, float Fb_re,
                             float Fb_im) {
    return fstat_from_sums(Ad, Bd, Cd, Fa_re, Fa_im, Fb_re, Fb_im);
}

/* ---- exponential-window sums of one cell ---------------------------------------------------
 * lal flavour (L1, L2): REAL8 window value (lookup table unless exact_exp), each term
 * `REAL4 atom * REAL8 window` evaluated in double and added to the REAL4 accumulator.
 * The kernels use t_i = t0_data + i*TAtom (Exp.cu:84); identical to the stored timestamp of a
 * filled bin, and empty bins are all-zero. */
static inline void exp_cell_lal(const atom_t *merged, uint32_t i_t0, uint32_t i_t1, uint32_t t0_data,
                                uint32_t TAtom, uint32_t t0_m, uint32_t t1, uint32_t tau_n,
                                const int exact_exp, float S[7]) {
    float Ad = 0, Bd = 0, Cd = 0, Fa_re = 0, Fa_im = 0, Fb_re = 0, Fb_im = 0;
    for (uint32_t i = i_t0; i <= i_t1; i++) {
        const atom_t *a = &merged[i];
        uint32_t t_i = t0_data + i * TAtom;
        double win_i = 0.0;
        if (t_i >= t0_m && t_i <= t1) {
            double x = 1.0 * (t_i - t0_m) / tau_n;
            win_i = exact_exp ? exp(-x) : fast_neg_exp(x);
        }
        double win2_i = win_i * win_i;
        Ad += a->a2 * win2_i; /* (float)((double)Ad + (double)a2*win2) */
        Bd += a->b2 * win2_i;
        Cd += a->ab * win2_i;
        Fa_re += a->Fa_re * win_i;
        Fa_im += a->Fa_im * win_i;
        Fb_re += a->Fb_re * win_i;
        Fb_im += a->Fb_im * win_i;
    }
    S[0] = Ad; S[1] = Bd; S[2] = Cd; S[3] = Fa_re; S[4] = Fa_im; S[5] = Fb_re; S[6] = Fb_im;
}

/* Exp.cu:82-102 literally: float window value, float products */
static inline void exp_cell_pycuda(const atom_t *merged, uint32_t i_t0, uint32_t i_t1,
                                   uint32_t t0_data, uint32_t TAtom, uint32_t t0_m, uint32_t t1,
                                   uint32_t tau_n, const int exact_exp, float S[7]) {
    float Ad = 0, Bd = 0, Cd = 0, Fa_re = 0, Fa_im = 0, Fb_re = 0, Fb_im = 0;
    for (uint32_t i = i_t0; i <= i_t1; i++) {
        const atom_t *a = &merged[i];
        uint32_t t_i = t0_data + i * TAtom;
        float win_i = 0.0;
        if (t_i >= t0_m && t_i <= t1) {
            float x = 1.0 * (t_i - t0_m) / tau_n;
            win_i = exact_exp ? expf(-x) : (float)fast_neg_exp(x);
        }
        float win2_i = win_i * win_i;
        Ad += a->a2 * win2_i;
        Bd += a->b2 * win2_i;
        Cd += a->ab * win2_i;
        Fa_re += a->Fa_re * win_i;
        Fa_im += a->Fa_im * win_i;
        Fb_re += a->Fb_re * win_i;
        Fb_im += a->Fb_im * win_i;
    }
    S[0] = Ad; S[1] = Bd; S[2] = Cd; S[3] = Fa_re; S[4] = Fa_im; S[5] = Fb_re; S[6] = Fb_im;
}

/* N_t0Range, N_tauRange (tcw:775-780) */
int oracle_map_dims(const window_range_t *w, uint32_t *N_t0, uint32_t *N_tau) {
    if (w->type >= WIN_LAST) return ERR_WINDOW;
    if (w->type == WIN_NONE) {
        *N_t0 = 1;
        *N_tau = 1;
        return 0;
    }
    if (w->dt0 == 0 || w->dtau == 0) return ERR_INVALID;
    *N_t0 = (uint32_t)floor(1.0 * w->t0Band / w->dt0) + 1;
    *N_tau = (uint32_t)floor(1.0 * w->tauBand / w->dtau) + 1;
    return 0;
}

/* ---- the map (structure of XLALComputeTransientFstatMap as ported in tcw:656-834 and the
 *      two kernels; m outer / n inner; running sums for rect, Rect.cu:33-40, 75-91) -------
 * merged: binned atoms (oracle_merge_binned).  F_mn: [N_t0*N_tau] doubles (lal keeps REAL4 F
 * in a REAL8 gsl_matrix).  rect_vanilla != 0 recomputes each rect cell from
 * scratch (the reference's "#if 0" sanity method; identical floats for sane windows). */
int oracle_map(const atom_t *merged, uint32_t numAtoms, uint32_t TAtom, const window_range_t *win_in,
               int semantics, int exact_exp, int allow_degenerate, int rect_vanilla, double *F_mn,
               oracle_result_t *res) {
    if (!merged || !win_in || !res || numAtoms == 0 || TAtom == 0) return ERR_INVALID;
    if (!expLUT) create_exp_lut();
    window_range_t w = *win_in; /* by value: never mutate the caller's (tcw:742-749 does) */
    if (w.type >= WIN_LAST) return ERR_WINDOW;
    uint32_t t0_data = merged[0].timestamp;
    if (w.type == WIN_NONE) { /* tcw:742-749 */
        w.type = WIN_RECT;
        w.t0 = t0_data;
        w.t0Band = 0;
        w.dt0 = TAtom;
        w.tau = numAtoms * TAtom;
        w.tauBand = 0;
        w.dtau = TAtom;
    }
    uint32_t N_t0, N_tau;
    int rc = oracle_map_dims(&w, &N_t0, &N_tau);
    if (rc) return rc;

    memset(res, 0, sizeof(*res));
    res->N_t0 = N_t0;
    res->N_tau = N_tau;
    res->numAtoms = numAtoms;
    res->t0_data = t0_data;
    res->lnBtSG = res->t0_MP = res->tau_MP = NAN;
    float maxF = -1.0f; /* tcw:135-139 */
    uint32_t m_ML = 0, n_ML = 0, t0_ML = 0, tau_ML = 0;
    int degenerate = 0;

    for (uint32_t m = 0; m < N_t0; m++) {
        uint32_t t0_m = w.t0 + m * w.dt0;
        float Ad = 0, Bd = 0, Cd = 0, Fa_re = 0, Fa_im = 0, Fb_re = 0, Fb_im = 0;
        uint32_t i_t1_last = 0;
        int first = 1;
        for (uint32_t n = 0; n < N_tau; n++) {
            uint32_t tau_n = w.tau + n * w.dtau;
            uint32_t i_t0, i_t1, t1;
            oracle_index_range(w.type, t0_m, tau_n, t0_data, TAtom, numAtoms, &i_t0, &i_t1, &t1);
            if (first) {
                i_t1_last = i_t0;
                first = 0;
            }
            if (i_t1 == i_t0) degenerate = 1; /* L3 */

            if (w.type == WIN_RECT) {
                if (rect_vanilla) {
                    Ad = Bd = Cd = Fa_re = Fa_im = Fb_re = Fb_im = 0;
                    i_t1_last = i_t0;
                }
                for (uint32_t i = i_t1_last; i <= i_t1 && i < numAtoms; i++) {
                    const atom_t *a = &merged[i];
                    Ad += a->a2;
                    Bd += a->b2;
                    Cd += a->ab;
                    Fa_re += a->Fa_re;
                    Fa_im += a->Fa_im;
                    Fb_re += a->Fb_re;
                    Fb_im += a->Fb_im;
                    i_t1_last = i_t1 + 1; /* inside the loop, as Rect.cu:90 */
                }
            } else { /* WIN_EXP, Exp.cu:72-102 */
                float S[7];
                if (semantics == SEM_LAL) {
                    if (exact_exp)
                        exp_cell_lal(merged, i_t0, i_t1, t0_data, TAtom, t0_m, t1, tau_n, 1, S);
                    else
                        exp_cell_lal(merged, i_t0, i_t1, t0_data, TAtom, t0_m, t1, tau_n, 0, S);
                } else {
                    if (exact_exp)
                        exp_cell_pycuda(merged, i_t0, i_t1, t0_data, TAtom, t0_m, t1, tau_n, 1, S);
                    else
                        exp_cell_pycuda(merged, i_t0, i_t1, t0_data, TAtom, t0_m, t1, tau_n, 0, S);
                }
                Ad = S[0]; Bd = S[1]; Cd = S[2];
                Fa_re = S[3]; Fa_im = S[4]; Fb_re = S[5]; Fb_im = S[6];
            }
            float F = fstat_from_sums(Ad, Bd, Cd, Fa_re, Fa_im, Fb_re, Fb_im);
            if (F > maxF) { /* strict >, first occurrence (A.3) */
                maxF = F;
                m_ML = m;
                n_ML = n;
                t0_ML = t0_m;
                tau_ML = tau_n;
            }
            if (F_mn) F_mn[(size_t)m * N_tau + n] = (double)F;
        }
    }
    res->maxF = (double)maxF;
    res->m_ML = m_ML;
    res->n_ML = n_ML;
    res->t0_ML = t0_ML;
    res->tau_ML = tau_ML;
    if (degenerate && !allow_degenerate && semantics == SEM_LAL) {
        res->status = ERR_DEGENERATE;
        return ERR_DEGENERATE;
    }
    return 0;
}

/* ---- lnBtSG and max-posterior estimates -------------------------------------------------
 * tcw:196-287 (Python ports of XLALComputeTransientBstat, XLALComputeTransientPosterior_t0/
 * _tau, XLALFindModeOfPDF1D).  use_lut != 0: lal flavour, each term XLALFastNegExp(maxF-F)
 * (L1); use_lut == 0: exact exp in double.  Sequential double sums, m outer / n inner. */
int oracle_bstat(const double *F_mn, uint32_t N_t0, uint32_t N_tau, double maxF,
                 const window_range_t *win, int use_lut, oracle_result_t *res) {
    if (!F_mn || !res || N_t0 == 0 || N_tau == 0) return ERR_INVALID;
    if (!expLUT) create_exp_lut();
    double sum_eB = 0;
    double *rows = (double *)calloc(N_t0, sizeof(double));
    double *cols = (double *)calloc(N_tau, sizeof(double));
    if (!rows || !cols) {
        free(rows);
        free(cols);
        return ERR_INVALID;
    }
    for (uint32_t m = 0; m < N_t0; m++) {
        for (uint32_t n = 0; n < N_tau; n++) {
            double DeltaF = maxF - F_mn[(size_t)m * N_tau + n];
            double e = use_lut ? fast_neg_exp(DeltaF) : exp(-DeltaF);
            sum_eB += e;
            rows[m] += e;
            cols[n] += e;
        }
    }
    double logBhat = maxF + log(sum_eB);
    double normBh = 70.0 / ((double)N_t0 * (double)N_tau);
    res->lnBtSG = log(normBh) + logBhat;
    uint32_t mb = 0, nb = 0;
    for (uint32_t m = 1; m < N_t0; m++)
        if (rows[m] > rows[mb]) mb = m;
    for (uint32_t n = 1; n < N_tau; n++)
        if (cols[n] > cols[nb]) nb = n;
    res->m_MP = mb;
    res->n_MP = nb;
    /* bin centre, dx = band / N (tcw:248-251, 283-286) */
    res->t0_MP = win->t0 + (mb + 0.5) * ((double)win->t0Band / N_t0);
    res->tau_MP = win->tau + (nb + 0.5) * ((double)win->tauBand / N_tau);
    free(rows);
    free(cols);
    return 0;
}

/* ---- one template end to end: merge + map (+ BtSG) -------------------------------------- */
int oracle_template(const atom_t *atoms, const uint32_t *n_atoms, int numDet, uint32_t stride,
                    uint32_t TAtom, const window_range_t *win, int semantics, int exact_exp,
                    int allow_degenerate, int want_btsg, double *F_mn /* nullable */,
                    oracle_result_t *res) {
    uint32_t cap = 0;
    /* upper bound on bins: span/TAtom + 1 */
    uint32_t tMin = 0xffffffffu, tMax = 0;
    for (int X = 0; X < numDet; X++) {
        if (n_atoms[X] == 0) continue;
        const atom_t *a = atoms + (size_t)X * stride;
        if (a[0].timestamp < tMin) tMin = a[0].timestamp;
        if (a[n_atoms[X] - 1].timestamp > tMax) tMax = a[n_atoms[X] - 1].timestamp;
    }
    if (tMax < tMin) return ERR_INVALID;
    cap = (tMax - tMin) / TAtom + 2;
    atom_t *merged = (atom_t *)malloc((size_t)cap * sizeof(atom_t));
    if (!merged) return ERR_INVALID;
    uint32_t N = 0;
    int rc = oracle_merge_binned(atoms, n_atoms, numDet, stride, TAtom, merged, cap, &N);
    if (rc) {
        free(merged);
        return rc;
    }
    window_range_t w = *win;
    if (w.type == WIN_NONE) {
        w.type = WIN_RECT;
        w.t0 = merged[0].timestamp;
        w.t0Band = 0;
        w.dt0 = TAtom;
        w.tau = N * TAtom;
        w.tauBand = 0;
        w.dtau = TAtom;
    } else if (w.type >= WIN_LAST) {
        free(merged);
        return ERR_WINDOW;
    }
    uint32_t N_t0, N_tau;
    rc = oracle_map_dims(&w, &N_t0, &N_tau);
    if (rc) {
        free(merged);
        return rc;
    }
    double *F = F_mn;
    if (!F && want_btsg) F = (double *)malloc((size_t)N_t0 * N_tau * sizeof(double));
    rc = oracle_map(merged, N, TAtom, &w, semantics, exact_exp, allow_degenerate, 0, F, res);
    if ((rc == 0 || rc == ERR_DEGENERATE) && want_btsg && F) {
        int32_t st = res->status;
        oracle_bstat(F, N_t0, N_tau, res->maxF, &w, !exact_exp, res);
        res->status = st;
    }
    if (F && F != F_mn) free(F);
    free(merged);
    return rc;
}

/* ---- batch: T templates, OpenMP over templates (each template single-threaded, exactly
 *      like one XLALComputeTransientFstatMap call).  Used by bench.py's CPU legs. ---------- */
int oracle_batch(const atom_t *atoms, const uint32_t *n_atoms, uint32_t stride, uint32_t TAtom,
                 int T, int numDet, const window_range_t *win, int semantics, int exact_exp,
                 int allow_degenerate, int want_btsg, int num_threads, oracle_result_t *results) {
    int worst = 0;
    if (!expLUT) create_exp_lut(); /* before the threads start */
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(num_threads > 0 ? num_threads : 1)
#endif
    for (int t = 0; t < T; t++) {
        int rc = oracle_template(atoms + (size_t)t * numDet * stride, n_atoms + (size_t)t * numDet,
                                 numDet, stride, TAtom, win, semantics, exact_exp,
                                 allow_degenerate, want_btsg, NULL, &results[t]);
        if (rc) {
#ifdef _OPENMP
#pragma omp critical
#endif
            worst = rc;
        }
    }
    return worst;
}
