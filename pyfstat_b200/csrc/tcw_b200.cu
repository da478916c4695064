// tcw_b200.cu -- C ABI (include/tcw_b200.h) of the B200-native transient F-stat map backend.
//
// Host orchestration only: buffer management in HBM, the host-side certificate that decides
// between the tiled and the generic kernels, launches on the handle's stream, event timing.
// No CPU compute path exists here: every map is computed by the sm_100a kernels in
// tcw_prep.cuh / tcw_rect.cuh / tcw_exp.cuh / tcw_generic.cuh / tcw_btsg.cuh.
//
// Data layout in HBM (per handle, grown on demand, reused across calls):
//   d_atoms  [T][numDet][stride] tcw_atom (32 B AoS, as uploaded)
//   d_X      [T][7][xpad] float32   merged channels (zero padded)
//   d_X8     [T][xpad][8] float32   the same atoms, channel-interleaved 32-byte records (exp kernel)
//   d_P      [T][7][ppad] float64   exclusive prefix sums (rect window)
//   d_W      [n_tiles][KW][64] float32  exponential-window weight table (cached per window)
//   d_Fmn    [T][N_t0][N_tau] float32   only with TCW_WANT_FMN; else a sub-batch scratch
//                                       sized to stay L2-resident when lnBtSG needs a 2nd pass
//   d_zero   one zero-initialised region per map: max keys [T] u64, lnBtSG marginals [T][N_t0] and
//            [T][N_tau] (64-bit fixed point), flags [T], tile-queue counters; d_results [T]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "tcw_b200.h"
#include "tcw_btsg.cuh"
#include "tcw_common.cuh"
#include "tcw_exp.cuh"
#include "tcw_exp_rec.cuh"
#include "tcw_generic.cuh"
#include "tcw_prep.cuh"
#include "tcw_rect.cuh"
#include "tcw_rect_p.cuh"

static std::string g_create_error;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct tcw_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // H2D of atom chunks, overlapped with the kernels
    std::vector<cudaEvent_t> ev_up;
    cudaEvent_t ev_timer[2] = {nullptr, nullptr};
    cudaEvent_t ev_stage[4] = {nullptr, nullptr, nullptr, nullptr};  // start, prep, table, loops end... finalize
    cudaEvent_t ev_fin = nullptr;
    std::vector<cudaEvent_t> ev_sub;  // 3 per sub-batch: map begin, map end, btsg end
    std::vector<cudaEvent_t> ev_x;    // 2 per sub-batch (exp recurrence path): operands ready, tensor-core pass done
    bool x_valid = false;
    int n_sub_last = 0;
    bool stage_valid = false;
    std::string err;
    cudaDeviceProp prop;
    uint64_t launches = 0;

    // pinned staging of the small per-call arrays (n_atoms, meta, per-template windows): the caller's
    // pageable buffers are consumed before the call returns, with no stream synchronisation
    void *hp_small[2] = {nullptr, nullptr};  // [0] n_atoms + meta, [1] per-template windows
    size_t hp_small_cap[2] = {0, 0};
    cudaEvent_t ev_small[2] = {nullptr, nullptr};
    bool small_pending[2] = {false, false};

    bool in_flight = false, in_flight_fmn = false;  // tcw_submit .. tcw_wait
    // resident batch
    bool uploaded = false;
    int T = 0, numDet = 0;
    uint32_t stride = 0, TAtom = 0;
    uint32_t Nmax = 0, xpad = 0, ppad = 0;
    bool uniform = true;  // all templates share (t0_data, numAtoms)
    std::vector<TplMeta> meta;
    // d_zero: everything a map needs zero-initialised -- max keys, lnBtSG marginals, flags, tile-queue
    // counters -- in ONE region, cleared by one memset per map
    DevBuf d_atoms, d_natoms, d_meta, d_X, d_X8, d_P, d_Fmn, d_scratch, d_zero, d_results, d_W, d_Kn, d_lut, d_flush,
        d_wins, d_tilemax, d_shift, d_G, d_C, d_scale, d_Xd;
    int tc_f16 = 1;  // tensor-core pass of the exp window: FP16 operands (default) or TF32 ($TCW_TC_TF32=1)
    // rect launches through the persistent warp-specialised kernel: $TCW_RECT_PERSIST = 0 never,
    // 1 (default) when the launch has enough tiles to fill the GPU, 2 whenever the plan allows (tests)
    int rect_persist = 1;

    // last map
    bool have_fmn = false;
    uint32_t last_N_t0 = 0, last_N_tau = 0, last_pitch = 0;

    // exp weight-table cache key
    bool w_valid = false;
    tcw_window_range w_key = {};
    uint32_t w_t0_data = 0, w_TAtom = 0, w_KW = 0, w_TN = 0;
    // XLALFastNegExp table geometry (runtime: tcw_set_exp_lut)
    double lut_xmax = 0.0;
    uint32_t lut_len = 0;
    bool lut_canonical = false;
    int exp_variant = 1;  // ExpCfgB: measured fastest (r01: A 14.70 ms, B 13.77 ms, C 13.80 ms per 32 x 30-d templates)
    int w_exact = -1;
    int w_kind = -1;  // what d_W holds: 0 direct-sum weights, 1 tensor-core correction weights, 2 nothing (exact recurrence)
    std::vector<int32_t> w_Kn;
};

static int fail(tcw_handle *h, int code, const std::string &msg) {
    if (h) h->err = msg;
    else g_create_error = msg;
    return code;
}

#define CUDA_TRY(h, expr)                                                                        \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            return fail(h, _e == cudaErrorMemoryAllocation ? TCW_E_NOMEM : TCW_E_CUDA,           \
                        std::string(#expr) + ": " + cudaGetErrorString(_e));                     \
        }                                                                                        \
    } while (0)

static int ensure(tcw_handle *h, DevBuf &b, size_t bytes) {
    if (bytes <= b.cap) return TCW_OK;
    if (b.p) {
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        CUDA_TRY(h, cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = bytes + bytes / 8;  // a little slack against repeated growth
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        e = cudaMalloc(&b.p, want);
    }
    if (e != cudaSuccess) {
        b.p = nullptr;
        return fail(h, TCW_E_NOMEM,
                    "cudaMalloc of " + std::to_string(bytes) + " bytes failed: " + cudaGetErrorString(e));
    }
    b.cap = want;
    return TCW_OK;
}

// Pinned staging area of `bytes` bytes, free to be overwritten (the previous asynchronous copies out
// of it have completed).
static int small_stage(tcw_handle *h, int slot, size_t bytes, unsigned char **out) {
    if (h->small_pending[slot]) {
        CUDA_TRY(h, cudaEventSynchronize(h->ev_small[slot]));
        h->small_pending[slot] = false;
    }
    if (bytes > h->hp_small_cap[slot]) {
        if (h->hp_small[slot]) cudaFreeHost(h->hp_small[slot]);
        h->hp_small[slot] = nullptr;
        h->hp_small_cap[slot] = 0;
        const size_t want = std::max<size_t>(bytes + bytes / 4, 4096);
        if (cudaHostAlloc(&h->hp_small[slot], want, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            return fail(h, TCW_E_NOMEM, "cudaHostAlloc of the staging area failed");
        }
        h->hp_small_cap[slot] = want;
    }
    *out = (unsigned char *)h->hp_small[slot];
    return TCW_OK;
}

static void release(DevBuf &b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

static ExpLut lut_dev(const tcw_handle *h) {
    ExpLut l;
    l.tab = (const double *)h->d_lut.p;
    l.len = h->lut_len;
    l.xmax = h->lut_xmax;
    l.dxinv = (double)h->lut_len / h->lut_xmax;  // EXPLUT_DXINV = LENGTH / XMAX
    const double k = -(h->lut_xmax / (double)h->lut_len) * 1.4426950408889634074;  // -dx log2(e)
    // k_hi keeps 11 significant bits (10 stored) so that i0 * k_hi is exact in FP32 for i0 < 2^13
    float hi = (float)k;
    uint32_t bits;
    memcpy(&bits, &hi, 4);
    bits &= 0xFFFFE000u;
    memcpy(&hi, &bits, 4);
    l.neg_dx_log2e_hi = hi;
    l.neg_dx_log2e_lo = (float)(k - (double)hi);
    l.canonical = h->lut_canonical ? 1u : 0u;
    return l;
}

// Installs the XLALFastNegExp table: `length + 1` entries on [0, xmax], lookup index
// (UINT4)(x * length / xmax + 0.5).  table == NULL: built here as exp(-(i * dx)), dx = xmax / length,
// with the host's libm -- the way lalpulsar's XLALCreateExpLUT fills it.
static int set_exp_lut(tcw_handle *h, double xmax, uint32_t length, const double *table) {
    if (!(xmax > 0.0) || !(xmax < 1e6) || length < 1 || length > (1u << 22))
        return fail(h, TCW_E_INVALID, "tcw_set_exp_lut: need 0 < xmax < 1e6 and 1 <= length <= 2^22");
    std::vector<double> lut((size_t)length + 1);
    const double dx = xmax / (double)length;
    bool canonical = true;
    for (uint32_t i = 0; i <= length; i++) {
        const double ref = exp(-((double)i * dx));
        lut[i] = table ? table[i] : ref;
        if (table && !(fabs(table[i] - ref) <= 1e-13 * ref)) canonical = false;
    }
    int rc = ensure(h, h->d_lut, lut.size() * sizeof(double));
    if (rc) return rc;
    CUDA_TRY(h, cudaMemcpy(h->d_lut.p, lut.data(), lut.size() * sizeof(double), cudaMemcpyHostToDevice));
    h->lut_xmax = xmax;
    h->lut_len = length;
    h->lut_canonical = canonical;
    h->w_valid = false;  // cached exponential-window weights were built from the old table
    const size_t smem = TCW_BTSG_TABLE_SMEM(length);
    if (smem <= (size_t)h->prop.sharedMemPerBlockOptin) {
        CUDA_TRY(h, cudaFuncSetAttribute(tcw_btsg_table_kernel<false, true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(h, cudaFuncSetAttribute(tcw_btsg_table_kernel<false, false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    return TCW_OK;
}

extern "C" int tcw_set_exp_lut(tcw_handle *h, double xmax, uint32_t length, const double *table) {
    if (!h) return TCW_E_INVALID;
    if (h->in_flight) return fail(h, TCW_E_STATE, "a submitted batch is in flight: call tcw_wait first");
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return set_exp_lut(h, xmax, length, table);
}

extern "C" int tcw_get_exp_lut(const tcw_handle *h, double *xmax, uint32_t *length, int *canonical) {
    if (!h) return TCW_E_INVALID;
    if (xmax) *xmax = h->lut_xmax;
    if (length) *length = h->lut_len;
    if (canonical) *canonical = h->lut_canonical ? 1 : 0;
    return TCW_OK;
}

extern "C" int tcw_abi_version(void) { return TCW_ABI_VERSION; }

extern "C" const char *tcw_last_error(const tcw_handle *h) {
    return h ? h->err.c_str() : g_create_error.c_str();
}

extern "C" int tcw_map_dims(const tcw_window_range *win, uint32_t *N_t0, uint32_t *N_tau) {
    if (!win || !N_t0 || !N_tau) return TCW_E_INVALID;
    if (win->type >= TCW_WINDOW_LAST) return TCW_E_WINDOW;
    if (win->type == TCW_WINDOW_NONE) {
        *N_t0 = *N_tau = 1;
        return TCW_OK;
    }
    if (win->dt0 == 0 || win->dtau == 0) return TCW_E_INVALID;
    *N_t0 = win->t0Band / win->dt0 + 1;    // floor(t0Band/dt0)+1 (tcw:775-777)
    *N_tau = win->tauBand / win->dtau + 1;  // tcw:778-780
    return TCW_OK;
}

extern "C" int tcw_cell_index_range(uint32_t window_type, uint32_t t0_m, uint32_t tau_n,
                                    uint32_t t0_data, uint32_t TAtom, uint32_t numAtoms,
                                    uint32_t *i_t0, uint32_t *i_t1) {
    if (!i_t0 || !i_t1 || TAtom == 0 || numAtoms == 0) return TCW_E_INVALID;
    if (window_type != TCW_WINDOW_RECT && window_type != TCW_WINDOW_EXP) return TCW_E_WINDOW;
    IndexGeom g;
    g.TAtom = TAtom;
    g.TAtomHalf = TAtom / 2;
    g.md = make_magic(TAtom);
    g.ef = window_type == TCW_WINDOW_EXP ? TCW_EXP_EFOLDING : 1;
    *i_t0 = index_t0(t0_m, t0_data, numAtoms, g);
    *i_t1 = index_t1(t0_m + g.ef * tau_n, t0_data, numAtoms, g);
    return TCW_OK;
}

extern "C" int tcw_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return count;
}

extern "C" int tcw_device_name_of(int device, char *buf, int buflen) {
    if (!buf || buflen < 1) return TCW_E_INVALID;
    cudaDeviceProp prop;
    if (device < 0 || device >= tcw_device_count() || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        cudaGetLastError();
        return TCW_E_CUDA;
    }
    snprintf(buf, buflen, "%s", prop.name);
    return TCW_OK;
}

extern "C" int tcw_create(int device, tcw_handle **out) {
    if (!out) return fail(nullptr, TCW_E_INVALID, "tcw_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, TCW_E_CUDA,
                    std::string("no usable CUDA device: ") +
                        (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                        " (this backend has no CPU fallback)");
    if (device < 0) {  // $CUDA_DEVICE like the reference (tcw:434-437)
        const char *env = getenv("CUDA_DEVICE");
        device = env ? atoi(env) : 0;
    }
    if (device >= count)
        return fail(nullptr, TCW_E_CUDA,
                    "Requested CUDA device number " + std::to_string(device) +
                        " exceeds number of available devices!");
    tcw_handle *h = new tcw_handle();
    h->device = device;
    CUDA_TRY(nullptr, cudaSetDevice(device));
    CUDA_TRY(nullptr, cudaGetDeviceProperties(&h->prop, device));
    if (h->prop.major < 10) {
        std::string m = std::string("device ") + h->prop.name + " is sm_" + std::to_string(h->prop.major) +
                        std::to_string(h->prop.minor) + "; this library is built for sm_100a only";
        delete h;
        return fail(nullptr, TCW_E_CUDA, m);
    }
    CUDA_TRY(nullptr, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CUDA_TRY(nullptr, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (auto &ev : h->ev_timer) CUDA_TRY(nullptr, cudaEventCreate(&ev));
    for (auto &ev : h->ev_stage) CUDA_TRY(nullptr, cudaEventCreate(&ev));
    for (auto &ev : h->ev_small) CUDA_TRY(nullptr, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CUDA_TRY(nullptr, cudaEventCreate(&h->ev_fin));
    // XLALFastNegExp table.  Default geometry: SURVEY A.4-1 (xmax 20, 1/dx = 256 -> 5120 steps);
    // $TCW_EXP_LUT="xmax:length" or tcw_set_exp_lut() override it (round 1 compiled in 20:2000).
    {
        double xmax = TCW_EXPLUT_DEFAULT_XMAX;
        uint32_t length = TCW_EXPLUT_DEFAULT_LENGTH;
        if (const char *env = getenv("TCW_EXP_LUT")) {
            double x = 0;
            unsigned long n = 0;
            if (sscanf(env, "%lf:%lu", &x, &n) == 2) {
                xmax = x;
                length = (uint32_t)n;
            }
        }
        int rc = set_exp_lut(h, xmax, length, nullptr);
        if (rc) {
            g_create_error = h->err;
            delete h;
            return rc;
        }
    }
    // opt in to large dynamic shared memory once
#define RECT_ATTR(RR, STG)                                                                              \
    CUDA_TRY(nullptr, cudaFuncSetAttribute(tcw_rect_map_kernel<RR, STG>,                                 \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, TCW_RECT_SMEM)); \
    CUDA_TRY(nullptr, cudaFuncSetAttribute(tcw_rect_locate_kernel<RR, STG>,                              \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, TCW_RECT_SMEM))
    RECT_ATTR(4, true);
    RECT_ATTR(4, false);
    RECT_ATTR(1, true);
    RECT_ATTR(1, false);
#undef RECT_ATTR
#define EXP_ATTR(CFG)                                                                                          \
    CUDA_TRY(nullptr, cudaFuncSetAttribute(tcw_exp_map_canon_kernel<CFG>,                                      \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, CFG::kSmemC));         \
    CUDA_TRY(nullptr, cudaFuncSetAttribute(tcw_exp_map_kernel<CFG, true>,                                      \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, CFG::kSmem1));         \
    CUDA_TRY(nullptr, cudaFuncSetAttribute(tcw_exp_map_kernel<CFG, false>,                                     \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, CFG::kSmem))
    EXP_ATTR(ExpCfgA);
    EXP_ATTR(ExpCfgB);
    EXP_ATTR(ExpCfgC);
#undef EXP_ATTR
    if (const char *v = getenv("TCW_EXP_VARIANT")) h->exp_variant = atoi(v);
    CUDA_TRY(nullptr, cudaFuncSetAttribute(tcw_exptc_map_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCX_SMEM));
    CUDA_TRY(nullptr, cudaFuncSetAttribute(tcw_exptc_map_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCX_SMEM));
    if (const char *v = getenv("TCW_TC_TF32")) h->tc_f16 = atoi(v) ? 0 : 1;
#define WALK_ATTR1(HASC, NS, S1)                                                                                     \
    CUDA_TRY(nullptr, cudaFuncSetAttribute(tcw_exp_walk_kernel<HASC, NS, S1>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                           WalkCfg<NS>::smem(HASC)))
#define WALK_ATTR(NS)             \
    WALK_ATTR1(true, NS, true);   \
    WALK_ATTR1(true, NS, false);  \
    WALK_ATTR1(false, NS, true);  \
    WALK_ATTR1(false, NS, false)
    WALK_ATTR(1);
    WALK_ATTR(2);
    WALK_ATTR(4);
    WALK_ATTR(8);
    WALK_ATTR(16);
#undef WALK_ATTR
#undef WALK_ATTR1
    CUDA_TRY(nullptr, cudaFuncSetAttribute(tcw_rect_map_p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           TCW_RECTP_SMEM));
    if (const char *v = getenv("TCW_RECT_PERSIST")) h->rect_persist = atoi(v);
    *out = h;
    return TCW_OK;
}

extern "C" int tcw_destroy(tcw_handle *h) {
    if (!h) return TCW_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (DevBuf *b : {&h->d_atoms, &h->d_natoms, &h->d_meta, &h->d_X, &h->d_X8, &h->d_P, &h->d_Fmn, &h->d_scratch,
                      &h->d_zero, &h->d_results, &h->d_W, &h->d_Kn, &h->d_lut, &h->d_flush, &h->d_wins,
                      &h->d_tilemax, &h->d_shift, &h->d_G, &h->d_C, &h->d_scale, &h->d_Xd})
        release(*b);
    for (auto ev : h->ev_timer)
        if (ev) cudaEventDestroy(ev);
    for (auto ev : h->ev_stage)
        if (ev) cudaEventDestroy(ev);
    if (h->ev_fin) cudaEventDestroy(h->ev_fin);
    for (auto ev : h->ev_small)
        if (ev) cudaEventDestroy(ev);
    for (auto p : h->hp_small)
        if (p) cudaFreeHost(p);
    for (auto ev : h->ev_sub) cudaEventDestroy(ev);
    for (auto ev : h->ev_x) cudaEventDestroy(ev);
    for (auto ev : h->ev_up) cudaEventDestroy(ev);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return TCW_OK;
}

extern "C" int tcw_device_name(const tcw_handle *h, char *buf, int buflen) {
    if (!h || !buf || buflen < 1) return TCW_E_INVALID;
    snprintf(buf, buflen, "%s", h->prop.name);
    return TCW_OK;
}

extern "C" uint64_t tcw_launch_count(const tcw_handle *h) { return h ? h->launches : 0; }

extern "C" int tcw_synchronize(tcw_handle *h) {
    if (!h) return TCW_E_INVALID;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return TCW_OK;
}

extern "C" void *tcw_host_alloc(uint64_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
extern "C" void tcw_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

extern "C" int tcw_timer_start(tcw_handle *h) {
    if (!h) return TCW_E_INVALID;
    CUDA_TRY(h, cudaEventRecord(h->ev_timer[0], h->stream));
    return TCW_OK;
}
extern "C" int tcw_timer_stop(tcw_handle *h, float *ms) {
    if (!h || !ms) return TCW_E_INVALID;
    CUDA_TRY(h, cudaEventRecord(h->ev_timer[1], h->stream));
    CUDA_TRY(h, cudaEventSynchronize(h->ev_timer[1]));
    CUDA_TRY(h, cudaEventElapsedTime(ms, h->ev_timer[0], h->ev_timer[1]));
    return TCW_OK;
}

extern "C" int tcw_flush_l2(tcw_handle *h) {
    if (!h) return TCW_E_INVALID;
    const size_t bytes = 256u << 20;
    int rc = ensure(h, h->d_flush, bytes);
    if (rc) return rc;
    CUDA_TRY(h, cudaMemsetAsync(h->d_flush.p, 0xA5, bytes, h->stream));
    return TCW_OK;
}

// ---------------------------------------------------------------------------------------
// upload
// ---------------------------------------------------------------------------------------
// Geometry of the batch + device buffers + the small per-vector arrays; the atoms themselves are
// copied here (copy_atoms) or, chunk by chunk and overlapped with the kernels, by map_impl.
static int upload_common(tcw_handle *h, const tcw_atom *atoms, const uint32_t *n_atoms,
                         uint32_t atom_stride, uint32_t TAtom, int T, int numDet, bool copy_atoms) {
    if (!h) return TCW_E_INVALID;
    if (!atoms || !n_atoms || T < 1 || numDet < 1 || atom_stride < 1 || TAtom < 1)
        return fail(h, TCW_E_INVALID, "tcw_upload_atoms: bad argument");
    if (h->in_flight) return fail(h, TCW_E_STATE, "a submitted batch is in flight: call tcw_wait first");
    CUDA_TRY(h, cudaSetDevice(h->device));
    h->uploaded = false;
    h->have_fmn = false;
    // merged geometry per template (XLALmergeMultiFstatAtomsBinned sizing):
    // tMin/tMax over detectors, numAtoms = floor((tMax - tMin)/TAtom) + 1
    h->meta.resize(T);
    uint32_t Nmax = 0;
    bool uniform = true;
    for (int t = 0; t < T; t++) {
        uint32_t tMin = 0xFFFFFFFFu, tMax = 0;
        bool any = false;
        for (int X = 0; X < numDet; X++) {
            const uint32_t n = n_atoms[(size_t)t * numDet + X];
            if (n > atom_stride)
                return fail(h, TCW_E_INVALID, "tcw_upload_atoms: n_atoms out of range [0, atom_stride]");
            if (n == 0) continue;  // a detector without atoms contributes nothing to the merge
            any = true;
            const tcw_atom *a = atoms + ((size_t)t * numDet + X) * atom_stride;
            tMin = std::min(tMin, a[0].timestamp);
            tMax = std::max(tMax, a[n - 1].timestamp);
        }
        if (!any) return fail(h, TCW_E_INVALID, "tcw_upload_atoms: a template has no atoms in any detector");
        if (tMax < tMin) return fail(h, TCW_E_INVALID, "tcw_upload_atoms: timestamps not increasing");
        h->meta[t].t0_data = tMin;
        h->meta[t].numAtoms = (tMax - tMin) / TAtom + 1;
        Nmax = std::max(Nmax, h->meta[t].numAtoms);
        if (h->meta[t].t0_data != h->meta[0].t0_data || h->meta[t].numAtoms != h->meta[0].numAtoms)
            uniform = false;
    }
    if ((uint64_t)Nmax * TAtom >= (1ull << 32))
        return fail(h, TCW_E_INVALID, "tcw_upload_atoms: data span does not fit UINT4 seconds");
    h->T = T;
    h->numDet = numDet;
    h->stride = atom_stride;
    h->TAtom = TAtom;
    h->Nmax = Nmax;
    h->uniform = uniform;
    // zero padding: an exp tile reads up to (TM - 1) * A + KC + 8 atoms past the last needed one (A <= 4)
#ifdef TCW_XPAD_OLD  // development: the round-1 padding (valid for A == 1 only)
    h->xpad = ((Nmax + 64 + 2 * TCW_EXP_KC + 8) + 3u) & ~3u;
#else
    h->xpad = ((Nmax + 64 /* max exp tile rows */ * TCW_EXP_AMAX + 2 * TCW_EXP_KC + 8) + 3u) & ~3u;
#endif
    h->ppad = ((Nmax + 1 + 8) + 1u) & ~1u;
    const size_t n_vec = (size_t)T * numDet;
    int rc;
    if ((rc = ensure(h, h->d_atoms, n_vec * atom_stride * sizeof(tcw_atom)))) return rc;
    if ((rc = ensure(h, h->d_natoms, n_vec * sizeof(uint32_t)))) return rc;
    if ((rc = ensure(h, h->d_meta, (size_t)T * sizeof(TplMeta)))) return rc;
    if (copy_atoms)
        CUDA_TRY(h, cudaMemcpyAsync(h->d_atoms.p, atoms, n_vec * atom_stride * sizeof(tcw_atom),
                                    cudaMemcpyHostToDevice, h->stream));
    // n_atoms (caller's, pageable) and meta go through the pinned staging area: consumed here, copied
    // asynchronously, no stream synchronisation on the latency path
    const size_t nb_n = n_vec * sizeof(uint32_t), nb_m = (size_t)T * sizeof(TplMeta);
    unsigned char *stg = nullptr;
    if ((rc = small_stage(h, 0, nb_n + nb_m, &stg))) return rc;
    memcpy(stg, n_atoms, nb_n);
    memcpy(stg + nb_n, h->meta.data(), nb_m);
    CUDA_TRY(h, cudaMemcpyAsync(h->d_natoms.p, stg, nb_n, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_meta.p, stg + nb_n, nb_m, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaEventRecord(h->ev_small[0], h->stream));
    h->small_pending[0] = true;
    // tcw_upload_atoms copies the caller's (possibly pageable) atoms here: consumed before returning
    if (copy_atoms) CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->uploaded = true;
    return TCW_OK;
}

extern "C" int tcw_upload_atoms(tcw_handle *h, const tcw_atom *atoms, const uint32_t *n_atoms,
                                uint32_t atom_stride, uint32_t TAtom, int T, int numDet) {
    return upload_common(h, atoms, n_atoms, atom_stride, TAtom, T, numDet, true);
}

// ---------------------------------------------------------------------------------------
// host-side certificate for the tiled kernels
// ---------------------------------------------------------------------------------------
// True if, for this template, no uint32 wrap-around and no signed re-interpretation can occur
// anywhere in the (skew-extended) map, so i_t0 / i_t1 are monotone in m and n.
static bool no_wrap(const MapWindow &w, uint32_t ef, uint32_t slack_n, const TplMeta &mt, uint32_t TAtom) {
    const int64_t half = TAtom / 2;
    const int64_t x0 = (int64_t)w.t0 - (int64_t)mt.t0_data + half;
    if (x0 < 0) return false;
    const uint64_t t0_last = (uint64_t)w.t0 + (uint64_t)(w.N_t0 - 1) * w.dt0;
    const uint64_t tau_last = (uint64_t)w.tau + (uint64_t)(w.N_tau - 1 + slack_n) * w.dtau;
    const uint64_t t1_max = t0_last + (uint64_t)ef * tau_last;
    if (t1_max >= (1ull << 32)) return false;
    const uint64_t q_max = (t1_max - mt.t0_data + (uint64_t)half) / TAtom;
    return q_max < (1ull << 31);
}

struct ExpPlan {
    bool ok = false;
    bool slide = true;   // A == 1: register sliding-window kernel
    bool canon = false;  // one class, A == 1, no shifts, no start beyond the data end: round 1's kernel
    uint32_t rec_A = 0;  // > 0: the recurrence path applies (one class, rows rec_A atoms apart, no shifts, no start
                         // beyond the data end); `ok` = the tiled direct-sum kernels apply
    uint32_t KW = 0;
    ExpClasses ec = {};
    int32_t delta[TCW_EXP_PMAX] = {0, 0, 0, 0};
    std::vector<int32_t> Kn;     // [P][N_tau]
    std::vector<int32_t> shift;  // [T]: (t0_data of template 0 - t0_data of template t) / TAtom
};

static void exp_tile_dims(int variant, uint32_t *TM, uint32_t *TN) {
    switch (variant) {
        case 0: *TM = ExpCfgA::kTM; *TN = ExpCfgA::kTN; break;
        case 2: *TM = ExpCfgC::kTM; *TN = ExpCfgC::kTN; break;
        default: *TM = ExpCfgB::kTM; *TN = ExpCfgB::kTN; break;
    }
}

static uint32_t gcd_u32(uint32_t a, uint32_t b) {
    while (b) {
        const uint32_t t = a % b;
        a = b;
        b = t;
    }
    return a;
}

// Host certificate + geometry of the tiled exponential-window kernel (tcw_exp.cuh): no uint32
// wrap-around anywhere in the map; rows fall into P <= 4 classes of equal (t0_m - t0_data) mod
// TAtom, consecutive rows of a class A <= 4 atoms apart; the templates' t0_data differ by whole
// atoms (per-template index shift); the `< 0 -> 0` clamp of i_t1 never engages.  Rows whose start
// index lies at or beyond the last atom are fine (zero sums -> the F = 2 fallback, as in the
// reference).  Anything else goes to the generic kernels.
static ExpPlan plan_exp(const tcw_handle *h, const MapWindow &w) {
    ExpPlan p;
    const uint32_t TAtom = h->TAtom;
    const uint32_t rem = w.dt0 % TAtom;
    const uint32_t P = rem ? TAtom / gcd_u32(rem, TAtom) : 1u;
    if (P > TCW_EXP_PMAX) return p;
    const uint64_t A64 = (uint64_t)P * w.dt0 / TAtom;  // exact by construction
    // rows more than AMAX atoms apart: no tiled direct sum, but the recurrence path (one row class) walks any step
    const bool tiled = A64 <= TCW_EXP_AMAX;
    if (A64 < 1 || (!tiled && (P != 1 || A64 > 4096))) return p;
    const uint32_t t0_data = h->meta[0].t0_data;
    p.shift.resize(h->T);
    for (int t = 0; t < h->T; t++) {
        const int64_t d = (int64_t)t0_data - (int64_t)h->meta[t].t0_data;
        if (d % (int64_t)TAtom != 0) return p;  // the weight tables depend on (t0 - t0_data) mod TAtom
        p.shift[t] = (int32_t)(d / (int64_t)TAtom);
        if (!no_wrap(w, TCW_EXP_EFOLDING, 0, h->meta[t], TAtom)) return p;
    }
    const int64_t half = TAtom / 2;
    const int64_t x0 = (int64_t)w.t0 - t0_data + half;  // >= 0 by no_wrap
    uint32_t TM, TN;
    exp_tile_dims(h->exp_variant, &TM, &TN);
    p.ec.P = P;
    p.ec.A = (uint32_t)A64;
    p.slide = A64 == 1;
    p.Kn.resize((size_t)P * w.N_tau);
    int64_t Kmax = -1;
    uint32_t ytiles = 0;
    for (uint32_t r = 0; r < P; r++) {
        const int64_t x0r = x0 + (int64_t)r * w.dt0;
        const int64_t i00 = x0r / TAtom;
        p.ec.i00[r] = (uint32_t)i00;
        p.delta[r] = (int32_t)((int64_t)t0_data + i00 * TAtom - ((int64_t)w.t0 + (int64_t)r * w.dt0));
        p.ec.ybeg[r] = ytiles;
        const uint32_t n_rows = r < w.N_t0 ? (w.N_t0 - r + P - 1) / P : 0u;
        ytiles += (n_rows + TM - 1) / TM;
        bool i00_zero = false;  // some template's first row of this class starts at atom 0
        for (int t = 0; t < h->T; t++) {
            if (i00 + p.shift[t] < 0) return p;
            i00_zero = i00_zero || (i00 + p.shift[t] == 0);
        }
        for (uint32_t n = 0; n < w.N_tau; n++) {
            const int64_t tau_n = (int64_t)w.tau + (int64_t)n * w.dtau;
            const int64_t K = (x0r + (int64_t)TCW_EXP_EFOLDING * tau_n) / TAtom - 1 - i00;  // unclamped i_t1 - i_t0
            if (K < 0 && i00_zero) return p;  // the `< 0 -> 0` clamp of i_t1 would engage
            p.Kn[(size_t)r * w.N_tau + n] = (int32_t)K;
            Kmax = std::max(Kmax, K);
        }
    }
    for (uint32_t r = P; r <= TCW_EXP_PMAX; r++) p.ec.ybeg[r] = ytiles;
    const bool tiled_grid_ok = ytiles <= 65535u;  // grid.y limit of the tiled kernels
    if (Kmax > (int64_t)h->Nmax + 64) Kmax = (int64_t)h->Nmax + 64;  // never need k beyond the data
    for (auto &K : p.Kn) K = (int32_t)std::min<int64_t>(K, Kmax);
    p.KW = (uint32_t)((std::max<int64_t>(Kmax, 0) + 1 + TCW_EXP_KC - 1) / TCW_EXP_KC * TCW_EXP_KC);
    const uint64_t n_tiles = (w.N_tau + TN - 1) / TN;
    const bool table_ok = (uint64_t)P * n_tiles * p.KW * TN * 4ull * TCW_EXP_WP <= (8ull << 30);  // direct-sum weight table
    bool rec = P == 1;
    for (int t = 0; t < h->T && rec; t++)
        rec = p.shift[t] == 0 &&
              (int64_t)p.ec.i00[0] + (int64_t)A64 * ((int64_t)w.N_t0 - 1) <= (int64_t)h->meta[t].numAtoms - 1;
    p.rec_A = rec ? (uint32_t)A64 : 0u;
    p.canon = rec && A64 == 1;
    p.ok = tiled && tiled_grid_ok && table_ok;
    return p;
}

// Tile shape of the rectangular-window kernel for one launch (tcw_rect.cuh): G row groups per
// warp (ROWS = 8 G R rows per tile), the head strip width DD (smallest multiple of 32 such that
// every tile starting at d = DD is off-diagonal; the device re-checks per tile, so DD only affects
// speed) and n_reg regular tiles of DT values of d.  Candidates are ranked by a small cost model:
// work per SM slot (cells + staging, which scales with ROWS + DT) plus half a tile as the
// expected tail -- tall, wide tiles for big batches, small ones when few templates must still
// fill the GPU.  TCW_RECT_G / TCW_RECT_NREG override (development).
struct RectPlan {
    uint32_t G = 1, DD = 32, DT = 32, n_reg = 0;
    bool staged = false;
};

static RectPlan plan_rect(const tcw_handle *h, const MapWindow &w, int S, uint32_t R, uint32_t TAtom,
                          const IndexGeom &g, uint32_t force_G = 0) {
    const TplMeta &mt0 = h->meta[0];
    const uint32_t d_total = w.N_tau + R - 1;
    const double slots = (double)h->prop.multiProcessorCount * TCW_RECT_MINB;
    const uint32_t g_max = R == 4 ? TCW_RECT_GMAX : 2;
    const char *env_g = getenv("TCW_RECT_G"), *env_n = getenv("TCW_RECT_NREG");
    RectPlan best;
    double best_cost = -1.0;
    for (uint32_t G = g_max; G >= 1; G /= 2) {
        if (env_g && !force_G && (uint32_t)atoi(env_g) != G) continue;
        if (force_G && G != force_G) continue;
        const uint32_t rows = TCW_RECT_WARPS * G * R;
        const uint32_t n_gy = (w.N_t0 + rows - 1) / rows;
        // widest regular tile whose end indices still fit the staged slice
        uint32_t dt_cap = 0;
        for (uint32_t dt = TCW_RECT_DT; dt >= 64; dt -= 32) {
            const uint64_t span = ((uint64_t)(rows - 1) * w.dt0 + (uint64_t)(dt - 1) * w.dtau) / TAtom + 6;
            if (span <= TCW_RECT_ECAP) {
                dt_cap = dt;
                break;
            }
        }
        RectPlan p;
        p.G = G;
        p.staged = dt_cap != 0;
        if (!p.staged) dt_cap = 1024;
        p.DD = 32;
        for (uint32_t cand = 32; p.staged && cand <= dt_cap; cand += 32) {
            bool all_off = true;
            for (uint32_t gy = 0; gy < n_gy && all_off; gy++) {
                const uint32_t m0 = gy * rows;
                const uint32_t m_last = std::min(m0 + rows, w.N_t0) - 1;
                const uint32_t e_lo = index_t1(w.t0 + w.tau + m0 * w.dt0 + cand * w.dtau, mt0.t0_data, mt0.numAtoms, g);
                const uint32_t s_hi = index_t0(w.t0 + m_last * w.dt0, mt0.t0_data, mt0.numAtoms, g);
                all_off = ((e_lo + 1) & ~1u) >= s_hi + 2;
            }
            if (all_off) {
                p.DD = cand;
                break;
            }
        }
        const uint32_t rest = d_total > p.DD ? d_total - p.DD : 0;
        const uint32_t n_min = (rest + dt_cap - 1) / dt_cap;
        for (uint32_t n_reg = n_min; n_reg <= n_min + 3; n_reg++) {
            if (env_n && (uint32_t)atoi(env_n) != n_reg && n_min <= (uint32_t)atoi(env_n)) continue;
            p.n_reg = n_reg;
            p.DT = 32;
            if (n_reg) {
                const uint32_t even = (rest + n_reg - 1) / n_reg;
                p.DT = (even + 127u) & ~127u;  // whole guarded blocks (TCW_RECT_JB chunks of 32)
                if (p.DT > dt_cap) p.DT = (even + 31u) & ~31u;
            }
            const double c_setup = 12.0;  // staging cost per staged entry, in cell evaluations (measured ~10 %)
            const double head = 2.0 * rows * p.DD + c_setup * (rows + p.DD);
            const double reg = (double)rows * p.DT + c_setup * (rows + p.DT);
            const double work = (double)S * n_gy * (head + n_reg * reg);
            const double cost = work / slots + 0.5 * std::max(head, n_reg ? reg : 0.0);
            if (best_cost < 0 || cost < best_cost) {
                best_cost = cost;
                best = p;
            }
            if (!n_reg) break;
        }
    }
    return best;
}

// A rect launch may run in the persistent kernel (tcw_rect_p.cuh) if its 128-row tiling (= the G = 4
// plan) is staged, every tile at d >= DD is off-diagonal for EVERY template of the batch (the plan
// checked template 0) and there are enough tiles to fill the GPU a few times over -- small
// launches (single-template calls) keep the one-tile-per-CTA kernel and its finer tiles.  The
// regular tiles are re-cut as wide as the staged slice allows.
static bool plan_rect_p(const tcw_handle *h, const MapWindow &w, int S, uint32_t R, uint32_t TAtom,
                        const IndexGeom &g, RectPlan *out) {
    if (!h->rect_persist || R != 4) return false;
    RectPlan p = plan_rect(h, w, S, R, TAtom, g, TCW_RECT_GMAX);
    if (!p.staged || p.G != TCW_RECT_GMAX) return false;
    const uint32_t rows = TCW_RECTP_ROWS;
    const uint32_t d_total = w.N_tau + R - 1;
    if (p.DD >= d_total) return false;  // no regular tiles at all
    uint32_t dt_cap = 0;
    for (uint32_t dt = TCW_RECT_DT; dt >= 64; dt -= 32)
        if (((uint64_t)(rows - 1) * w.dt0 + (uint64_t)(dt - 1) * w.dtau) / TAtom + 6 <= TCW_RECT_ECAP) {
            dt_cap = dt;
            break;
        }
    if (!dt_cap || ((uint64_t)(rows - 1) * w.dt0 + (uint64_t)(p.DD - 1) * w.dtau) / TAtom + 6 > TCW_RECT_ECAP) return false;
    const uint32_t rest = d_total - p.DD;
    p.n_reg = (rest + dt_cap - 1) / dt_cap;
    const uint32_t even = (rest + p.n_reg - 1) / p.n_reg;
    p.DT = (even + 127u) & ~127u;
    if (p.DT > dt_cap) p.DT = (even + 31u) & ~31u;
    const uint32_t n_gy = (w.N_t0 + rows - 1) / rows;
    if (h->rect_persist < 2 && (uint64_t)S * n_gy * (1 + p.n_reg) < 3ull * (uint64_t)h->prop.multiProcessorCount)
        return false;
    const int n_check = h->uniform ? 1 : h->T;
    for (int t = 0; t < n_check; t++) {
        const TplMeta &mt = h->meta[t];
        for (uint32_t gy = 0; gy < n_gy; gy++) {
            const uint32_t m0 = gy * rows;
            const uint32_t m_last = std::min(m0 + rows, w.N_t0) - 1;
            const uint32_t e_lo = index_t1(w.t0 + w.tau + m0 * w.dt0 + p.DD * w.dtau, mt.t0_data, mt.numAtoms, g);
            const uint32_t s_hi = index_t0(w.t0 + m_last * w.dt0, mt.t0_data, mt.numAtoms, g);
            if (((e_lo + 1) & ~1u) < s_hi + 2) return false;
        }
    }
    *out = p;
    return true;
}

static size_t subbatch_bytes() {
    const char *env = getenv("TCW_SUBBATCH_MB");
    size_t mb = env ? (size_t)atol(env) : 2048;
    if (mb < 1) mb = 1;
    return mb << 20;
}

// ---------------------------------------------------------------------------------------
// the map
// ---------------------------------------------------------------------------------------
// host_atoms == nullptr: the atoms are resident.  Otherwise they are still on the host (pinned
// for real overlap): the batch is cut into chunks whose H2D copies run on a second stream while
// the previous chunk is being computed.
static int map_impl_inner(tcw_handle *h, const tcw_window_range *win, uint32_t flags, const tcw_atom *host_atoms,
                          const tcw_window_range *per_template) {
    if (!h) return TCW_E_INVALID;
    if (!win) return fail(h, TCW_E_INVALID, "tcw_map_resident: window range is NULL");
    if (!h->uploaded) return fail(h, TCW_E_STATE, "tcw_map_resident: no resident atoms (call tcw_upload_atoms)");
    if (h->in_flight && host_atoms == nullptr)
        return fail(h, TCW_E_STATE, "a submitted batch is in flight: call tcw_wait first");
    if (win->type >= TCW_WINDOW_LAST)
        return fail(h, TCW_E_WINDOW, "Unknown window-type (" + std::to_string(win->type) +
                                         ") passed as input. Allowed are [0," +
                                         std::to_string(TCW_WINDOW_LAST - 1) + "].");
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int T = h->T;
    const uint32_t TAtom = h->TAtom;
    const bool want_fmn = flags & TCW_WANT_FMN;
    const bool want_btsg = flags & TCW_WANT_BTSG;
    const bool exact = flags & TCW_EXP_EXACT;
    const bool none_window = win->type == TCW_WINDOW_NONE;

    MapWindow w;
    if (none_window) {  // rect window spanning the data, per template, 1x1 (tcw:742-749)
        w.type = TCW_WINDOW_RECT;
        w.t0 = 0;  // taken from the template's meta inside the kernels
        w.tau = 0;
        w.dt0 = w.dtau = TAtom;
        w.t0Band = w.tauBand = 0;
        w.N_t0 = w.N_tau = 1;
        w.pitch = 4;
    } else {
        if (win->dt0 == 0 || win->dtau == 0)
            return fail(h, TCW_E_INVALID, "windowRange.dt0 and .dtau must be positive");
        w.type = win->type;
        w.t0 = win->t0;
        w.dt0 = win->dt0;
        w.tau = win->tau;
        w.dtau = win->dtau;
        w.t0Band = win->t0Band;
        w.tauBand = win->tauBand;
        w.N_t0 = win->t0Band / win->dt0 + 1;
        w.N_tau = win->tauBand / win->dtau + 1;
        w.pitch = (w.N_tau + 3u) & ~3u;
    }
    const uint64_t cells64 = (uint64_t)w.N_t0 * w.N_tau;
    if (cells64 >= (1ull << 32)) return fail(h, TCW_E_INVALID, "map has more than 2^32-1 cells");
    const size_t cells = (size_t)cells64;
    const size_t pcells = (size_t)w.N_t0 * w.pitch;  // device F_mn elements per template (padded rows)

    IndexGeom g;
    g.TAtom = TAtom;
    g.TAtomHalf = TAtom / 2;
    g.md = make_magic(TAtom);
    g.ef = w.type == TCW_WINDOW_EXP ? TCW_EXP_EFOLDING : 1;

    // ---- per-template window ranges: same type and map shape for all, generic kernels ----
    const MapWindow *d_wins = nullptr;
    if (per_template) {
        if (none_window) return fail(h, TCW_E_INVALID, "per-template windows cannot be TRANSIENT_NONE");
        int rcw = ensure(h, h->d_wins, (size_t)T * sizeof(MapWindow));
        if (rcw) return rcw;
        unsigned char *stg = nullptr;
        if ((rcw = small_stage(h, 1, (size_t)T * sizeof(MapWindow), &stg))) return rcw;
        MapWindow *wins = reinterpret_cast<MapWindow *>(stg);
        for (int t = 0; t < T; t++) {
            const tcw_window_range &pw = per_template[t];
            if (pw.type != win->type || pw.dt0 == 0 || pw.dtau == 0 || pw.t0Band / pw.dt0 + 1 != w.N_t0 ||
                pw.tauBand / pw.dtau + 1 != w.N_tau)
                return fail(h, TCW_E_INVALID, "per-template windows must share type and map shape (template " +
                                                  std::to_string(t) + ")");
            wins[t] = w;
            wins[t].t0 = pw.t0;
            wins[t].dt0 = pw.dt0;
            wins[t].tau = pw.tau;
            wins[t].dtau = pw.dtau;
            wins[t].t0Band = pw.t0Band;
            wins[t].tauBand = pw.tauBand;
        }
        // pinned staging + asynchronous copy: no synchronous pageable cudaMemcpy on the MCMC step path
        CUDA_TRY(h, cudaMemcpyAsync(h->d_wins.p, wins, (size_t)T * sizeof(MapWindow), cudaMemcpyHostToDevice,
                                    h->stream));
        CUDA_TRY(h, cudaEventRecord(h->ev_small[1], h->stream));
        h->small_pending[1] = true;
        d_wins = (const MapWindow *)h->d_wins.p;
    }

    // ---- choose the kernels ----
    enum { PATH_GENERIC = 0, PATH_FAST = 1 };
    int path = PATH_GENERIC;
    int rect_R = 1;
    RectPlan rp;
    ExpPlan ep;
    if (!(flags & TCW_FORCE_GENERIC) && !none_window && !per_template) {
        if (w.type == TCW_WINDOW_RECT) {
            // R = 4 rows per thread share an end index when dt0 == dtau; the diagonal tiles' per-group
            // split point (tcw_rect.cuh) additionally wants windows at least one row step long
            rect_R = (w.dt0 == w.dtau && w.tau >= w.dtau) ? 4 : 1;
            bool ok = true;
            for (int t = 0; t < T && ok; t++) ok = no_wrap(w, 1, rect_R - 1, h->meta[t], TAtom);
            if (ok) path = PATH_FAST;
        } else {
            ep = plan_exp(h, w);
            const bool rec_ok = ep.rec_A > 0 && !(flags & TCW_EXP_DIRECT) && (exact || h->lut_canonical);
            if (ep.ok || rec_ok) path = PATH_FAST;
        }
    }
    // exponential window on a canonical grid: FP64 recurrence down the rows (+ the tensor-core correction
    // in lookup-table mode), tcw_exp_rec.cuh; TCW_EXP_DIRECT keeps the tiled direct sum
    // (an uploaded table that is not e^{-i dx} takes the direct sum: the tensor-core pass relies on the table's
    // deviation from the exact exponential being small)
    // Rows rec_A > 1 atoms apart (dt0 a multiple of TAtom): the same two kernels on the REFINED grid of one row per
    // atom -- the recurrence has to step through every atom anyway, the tensor-core pass computes the correction
    // sums of all refined rows (as much work as the dt0 = TAtom map of the same data, still far below the direct
    // sum's) -- and the walk emits every rec_A-th row.
    const bool exp_rec = path == PATH_FAST && w.type == TCW_WINDOW_EXP && ep.rec_A > 0 && !(flags & TCW_EXP_DIRECT) &&
                         (exact || h->lut_canonical);
    const uint32_t rec_A = exp_rec ? ep.rec_A : 1u;
    const uint32_t tc_rows = exp_rec ? rec_A * (w.N_t0 - 1u) + 1u : w.N_t0;  // refined rows
    const bool exp_tc = exp_rec && !exact;
    const bool tc_f16 = h->tc_f16 != 0;
    const uint32_t tc_rs = tc_f16 ? TcxCfg<true>::kRowStep : TcxCfg<false>::kRowStep, tc_kc = 8 * tc_rs, tc_span = TCX_IROWS;
    const uint32_t tc_n_nt = (w.N_tau + TCX_TAUS - 1) / TCX_TAUS, tc_n_mb = (tc_rows + tc_span - 1) / tc_span;
    const uint32_t tc_cpitch = tc_n_nt * TCX_TAUS, tc_U = (h->Nmax + tc_kc - 1) / tc_kc + 10;
    const uint32_t tc_chunks = path == PATH_FAST && w.type == TCW_WINDOW_EXP ? (ep.KW + tc_kc - 1) / tc_kc : 0;
    const size_t tc_c_per_tpl = (size_t)tc_rows * tc_cpitch * TcxC::kBytesPerCell;

    // ---- buffers ----
    int rc;
    if ((rc = ensure(h, h->d_X, (size_t)T * TCW_NCH * h->xpad * sizeof(float)))) return rc;
    // the atom-interleaved copy only feeds the tiled exp kernel, the prefix sums only the tiled rect kernel
    const bool need_X8 = path == PATH_FAST && w.type == TCW_WINDOW_EXP;
    const bool need_P = path == PATH_FAST && w.type == TCW_WINDOW_RECT;
    if (need_X8 && (rc = ensure(h, h->d_X8, (size_t)T * 8 * h->xpad * sizeof(float)))) return rc;
    if (need_P && (rc = ensure(h, h->d_P, (size_t)T * TCW_NCH * h->ppad * sizeof(double)))) return rc;
    if ((rc = ensure(h, h->d_results, (size_t)T * sizeof(tcw_result)))) return rc;
    // templates per sub-batch: with lnBtSG the F_mn of a sub-batch is written by the map kernel
    // and re-read by the BtSG pass.  Measured (60 d rect, T=64): 64 MB sub-batches (L2-resident
    // scratch, one template per launch) 1.8e11 cells/s, 2-4 GB sub-batches 3.0e11 cells/s --
    // launch gaps and tails cost more than the HBM round trip, so the scratch is sized for big
    // launches (TCW_SUBBATCH_MB, default 2048)
    int S = T;
    if (want_btsg) S = (int)std::max<size_t>(1, std::min<size_t>((size_t)T, subbatch_bytes() / (pcells * 4)));
    S = std::min(S, 32768);
    // exponential window on the fast kernels: compute >> upload, see the small first chunk below
    const bool small_first = host_atoms && T >= 16 && w.type == TCW_WINDOW_EXP && path == PATH_FAST;
    if (host_atoms && T >= 16 && !small_first) {  // sub-batches double as upload chunks: upload/compute overlap
        int chunks = 2;  // measured: MCMC step (23.6 MB, little compute) 1.29 ms with 1-2 chunks, 1.48 with 4, 1.92 with 8
        if (const char *env = getenv("TCW_UPLOAD_CHUNKS")) chunks = std::max(1, atoi(env));
        S = std::min(S, std::max(8, (T + chunks - 1) / chunks));
    }
    if (exp_rec) {  // FP64 copies of the atoms for the walk
        if ((rc = ensure(h, h->d_Xd, (size_t)T * TCW_NCH * h->xpad * sizeof(double)))) return rc;
    }
    // FP16 correction sums: |accumulator| <= max|atom| sum_k |V| <= 2^14 2^12 (e^{dx/2} - 1) 2.01 (terms of the window),
    // stored times 2^-tc_shift so that the bound stays below 2^15
    int tc_shift = 0;
    if (exp_tc) {
        const double dx = h->lut_xmax / (double)h->lut_len;
        const double tau_max = (double)w.tau + (double)(w.N_tau - 1) * (double)w.dtau;
        const double terms = std::min((double)ep.KW, tau_max / (double)TAtom + 1.0) + 1.0;
        const double bound = ldexp(2.01 * expm1(0.5 * dx) * terms, 14 + TCX_VSCALE_LOG2);
        while (tc_shift < 60 && ldexp(bound, -tc_shift) > 32768.0) tc_shift++;
    }
    if (exp_tc) {  // the correction sums of a sub-batch live in an HBM scratch (20 B per cell): cap it at 8 GB
        size_t cap = 8ull << 30;
        if (const char *env = getenv("TCW_EXP_SCRATCH_MB")) cap = std::max<size_t>(64, (size_t)atoll(env)) << 20;
        S = (int)std::max<size_t>(1, std::min<size_t>((size_t)S, cap / tc_c_per_tpl));
        if ((uint64_t)S * tc_n_nt * tc_n_mb >= 0xFFFFFFFFull) return fail(h, TCW_E_INVALID, "too many exp tiles in one launch");
        // a device short of memory gets smaller sub-batches instead of an error
        while ((rc = ensure(h, h->d_C, (size_t)S * tc_c_per_tpl)) == TCW_E_NOMEM && S > 1) S = (S + 1) / 2;
        if (rc) return rc;
        if ((rc = ensure(h, h->d_G, (size_t)S * tc_U * TCX_G_BYTES_PER_U))) return rc;
        if ((rc = ensure(h, h->d_scale, (size_t)S * 4 * sizeof(float)))) return rc;
    }
    float *fmn_full = nullptr, *fmn_scratch = nullptr;
    if (want_fmn) {
        if ((rc = ensure(h, h->d_Fmn, (size_t)T * pcells * sizeof(float)))) return rc;
        fmn_full = (float *)h->d_Fmn.p;
    } else if (want_btsg) {
        if ((rc = ensure(h, h->d_scratch, (size_t)S * pcells * sizeof(float)))) return rc;
        fmn_scratch = (float *)h->d_scratch.p;
    }
    // Sub-batches double as upload chunks.  Where the kernels of a template take much longer than its upload (the
    // exponential window), the FIRST chunk is made small -- nothing can run before it has arrived -- and the rest go
    // in sub-batches of S: sub-batch sb covers templates [sub_lo(sb), sub_lo(sb + 1)).
    int S0 = S;
    if (small_first) S0 = std::min(S, std::max(4, T / 16));
    if (small_first && getenv("TCW_UPLOAD_FIRST")) S0 = std::max(1, std::min(S, atoi(getenv("TCW_UPLOAD_FIRST"))));
    const int n_sub = T <= S0 ? 1 : 1 + (T - S0 + S - 1) / S;
    auto sub_lo = [&](int sb) { return sb == 0 ? 0 : std::min(T, S0 + (sb - 1) * S); };
    // rect tile plan + its group-max table: decided and allocated before anything is enqueued; a map
    // whose row tiles exceed the grid limit takes the generic kernels instead of failing
    uint32_t *groupmax = nullptr;
    bool rect_p = false;
    if (path == PATH_FAST && w.type == TCW_WINDOW_RECT) {
        rect_p = plan_rect_p(h, w, S, (uint32_t)rect_R, TAtom, g, &rp);
        if (!rect_p) rp = plan_rect(h, w, S, (uint32_t)rect_R, TAtom, g);
        const uint32_t rows_per_tile = TCW_RECT_WARPS * rp.G * (uint32_t)rect_R;
        const uint32_t n_gy = (w.N_t0 + rows_per_tile - 1) / rows_per_tile;
        if (n_gy > 65535u) {
            path = PATH_GENERIC;
            rect_p = false;
        } else {
            if (!want_btsg) {
                // The map kernels track max VALUES only.  The argmax is completed by the lnBtSG pass
                // (which re-reads F_mn anyway) or, without it, by the locate kernel: one CTA per
                // template re-evaluates only the row groups whose maximum equals the template maximum
                // (table: one entry per d tile and row group of the map).
                const size_t n_grp = (w.N_t0 + rect_R - 1) / rect_R;
                const size_t n_entries = (size_t)S * (1 + rp.n_reg) * n_grp;
                if ((rc = ensure(h, h->d_tilemax, n_entries * sizeof(uint32_t)))) return rc;
                groupmax = (uint32_t *)h->d_tilemax.p;
            }
        }
    }
    const bool need_P2 = path == PATH_FAST && w.type == TCW_WINDOW_RECT;
    const ExpLut lut = lut_dev(h);
    // fixed-point scale of the lnBtSG marginals: every term is <= 1, so a map's total is <= cells
    int fx_bits = 62;
    for (uint64_t c = cells64; c > 1; c >>= 1) fx_bits--;
    fx_bits -= 1;
    const double fx_scale = ldexp(1.0, fx_bits);
    const bool btsg_table = !exact && ((flags & TCW_BTSG_TABLE) || !h->lut_canonical);
    if (want_btsg && btsg_table && TCW_BTSG_TABLE_SMEM(h->lut_len) > (size_t)h->prop.sharedMemPerBlockOptin)
        return fail(h, TCW_E_INVALID, "the installed exp table does not fit shared memory for the table-fetching lnBtSG pass");
    while ((int)h->ev_sub.size() < 3 * n_sub) {
        cudaEvent_t ev;
        CUDA_TRY(h, cudaEventCreate(&ev));
        h->ev_sub.push_back(ev);
    }
    while (exp_rec && (int)h->ev_x.size() < 2 * n_sub) {
        cudaEvent_t ev;
        CUDA_TRY(h, cudaEventCreate(&ev));
        h->ev_x.push_back(ev);
    }
    h->x_valid = false;

    // the zero-initialised region (one memset): 256-byte aligned sub-arrays
    size_t zero_bytes = 0;
    auto zero_take = [&zero_bytes](size_t bytes) {
        const size_t o = zero_bytes;
        zero_bytes += (bytes + 255) & ~(size_t)255;
        return o;
    };
    const size_t o_maxkey = zero_take((size_t)T * sizeof(unsigned long long));
    const size_t o_rowsum = zero_take(want_btsg ? (size_t)T * w.N_t0 * sizeof(unsigned long long) : 0);
    const size_t o_colsum = zero_take(want_btsg ? (size_t)T * w.N_tau * sizeof(unsigned long long) : 0);
    const size_t o_flags = zero_take((size_t)T * sizeof(uint32_t));
    const size_t o_counter = zero_take((size_t)n_sub * sizeof(uint32_t));
    if ((rc = ensure(h, h->d_zero, zero_bytes))) return rc;
    unsigned char *zbase = (unsigned char *)h->d_zero.p;
    unsigned long long *p_maxkey = (unsigned long long *)(zbase + o_maxkey);
    unsigned long long *p_rowsum = (unsigned long long *)(zbase + o_rowsum);
    unsigned long long *p_colsum = (unsigned long long *)(zbase + o_colsum);
    uint32_t *p_flags = (uint32_t *)(zbase + o_flags);
    uint32_t *p_counter = (uint32_t *)(zbase + o_counter);

    cudaStream_t st = h->stream;
    h->stage_valid = false;
    CUDA_TRY(h, cudaEventRecord(h->ev_stage[0], st));
    CUDA_TRY(h, cudaMemsetAsync(h->d_zero.p, 0, zero_bytes, st));

    // ---- stage 1: exponential-window weight table (cached across calls) ----
    uint32_t exp_TM = 0, exp_TN = 0;
    exp_tile_dims(h->exp_variant, &exp_TM, &exp_TN);
    if (path == PATH_FAST && w.type == TCW_WINDOW_EXP) {
        const int kind = exp_tc ? (tc_f16 ? 3 : 1) : exp_rec ? 2 : 0;
        const bool hit = h->w_valid && memcmp(&h->w_key, win, sizeof(*win)) == 0 &&
                         h->w_t0_data == h->meta[0].t0_data && h->w_TAtom == TAtom && h->w_kind == kind &&
                         h->w_exact == (int)exact && h->w_KW == ep.KW && h->w_Kn == ep.Kn && h->w_TN == exp_TN;
        if (!hit) {
            const uint32_t n_tiles = (w.N_tau + exp_TN - 1) / exp_TN;
            size_t total = (size_t)ep.ec.P * n_tiles * ep.KW * exp_TN;  // table cells (threads of the builder)
            if (kind == 1 || kind == 3) total = (size_t)tc_n_nt * tc_chunks * (TCX_TAUS * tc_kc);
            if (kind == 0 && (rc = ensure(h, h->d_W, total * TCW_EXP_WP * sizeof(float)))) return rc;
            if ((kind == 1 || kind == 3) && (rc = ensure(h, h->d_W, (size_t)tc_n_nt * tc_chunks * 32768))) return rc;
            if ((rc = ensure(h, h->d_Kn, ep.Kn.size() * sizeof(int32_t)))) return rc;
            h->w_valid = false;
            h->w_Kn = ep.Kn;  // keep the host copy alive for the async upload
            CUDA_TRY(h, cudaMemcpyAsync(h->d_Kn.p, h->w_Kn.data(), h->w_Kn.size() * sizeof(int32_t),
                                        cudaMemcpyHostToDevice, st));
            ExpTableGeom eg;
            eg.N_tau = w.N_tau;
            eg.n_tiles = n_tiles;
            eg.KW = ep.KW;
            eg.TN = exp_TN;
            eg.tau = w.tau;
            eg.dtau = w.dtau;
            eg.TAtom = TAtom;
            eg.P = ep.ec.P;
            for (int r = 0; r < TCW_EXP_PMAX; r++) eg.delta[r] = ep.delta[r];
            const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)h->prop.multiProcessorCount * 16);
            if (kind == 0)
                tcw_exp_table_kernel<<<blocks, 256, 0, st>>>((float *)h->d_W.p, (const int32_t *)h->d_Kn.p, eg, lut,
                                                             (int)exact);
            else if (kind == 1)
                tcw_exptc_table_kernel<false><<<blocks, 256, 0, st>>>(h->d_W.p, (const int32_t *)h->d_Kn.p, w.N_tau, tc_n_nt,
                                                                      tc_chunks, w.tau, w.dtau, TAtom, ep.delta[0], lut);
            else if (kind == 3)
                tcw_exptc_table_kernel<true><<<blocks, 256, 0, st>>>(h->d_W.p, (const int32_t *)h->d_Kn.p, w.N_tau, tc_n_nt,
                                                                     tc_chunks, w.tau, w.dtau, TAtom, ep.delta[0], lut);
            if (kind != 2) h->launches++;
            CUDA_TRY(h, cudaGetLastError());
            CUDA_TRY(h, cudaStreamSynchronize(st));  // w_Kn host buffer consumed
            h->w_valid = true;
            h->w_key = *win;
            h->w_t0_data = h->meta[0].t0_data;
            h->w_TAtom = TAtom;
            h->w_exact = (int)exact;
            h->w_kind = kind;
            h->w_KW = ep.KW;
            h->w_TN = exp_TN;
        }
        // per-template index shift (whole atoms between the templates' first timestamps)
        if ((rc = ensure(h, h->d_shift, (size_t)T * sizeof(int32_t)))) return rc;
        unsigned char *stg = nullptr;
        if ((rc = small_stage(h, 1, (size_t)T * sizeof(int32_t), &stg))) return rc;
        memcpy(stg, ep.shift.data(), (size_t)T * sizeof(int32_t));
        CUDA_TRY(h, cudaMemcpyAsync(h->d_shift.p, stg, (size_t)T * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        CUDA_TRY(h, cudaEventRecord(h->ev_small[1], st));
        h->small_pending[1] = true;
    }
    CUDA_TRY(h, cudaEventRecord(h->ev_stage[2], st));

    // ---- stage 0 + 2/3 per sub-batch: [H2D of the chunk on the copy stream] -> merge/scan ->
    //      map kernel (+ lnBtSG pass).  Sub-batches double as upload chunks.
    if (host_atoms) {
        while ((int)h->ev_up.size() < n_sub) {
            cudaEvent_t ev;
            CUDA_TRY(h, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            h->ev_up.push_back(ev);
        }
        const size_t per_tpl = (size_t)h->numDet * h->stride;
        for (int sb = 0; sb < n_sub; sb++) {
            const int lo = sub_lo(sb), cnt = sub_lo(sb + 1) - lo;
            CUDA_TRY(h, cudaMemcpyAsync((tcw_atom *)h->d_atoms.p + (size_t)lo * per_tpl, host_atoms + (size_t)lo * per_tpl,
                                        (size_t)cnt * per_tpl * sizeof(tcw_atom), cudaMemcpyHostToDevice,
                                        h->copy_stream));
            CUDA_TRY(h, cudaEventRecord(h->ev_up[sb], h->copy_stream));
        }
    }
    for (int sb = 0; sb < n_sub; sb++) {
        const int t_base = sub_lo(sb);
        const int cnt = sub_lo(sb + 1) - t_base;
        if (host_atoms) CUDA_TRY(h, cudaStreamWaitEvent(st, h->ev_up[sb], 0));
        // merge detectors, transpose to channels, FP64 prefix scan
        {
            dim3 grid_m((h->xpad + TCW_PREP_THREADS - 1) / TCW_PREP_THREADS, cnt);
            tcw_prep_merge_kernel<<<grid_m, TCW_PREP_THREADS, 0, st>>>(
                (const tcw_atom *)h->d_atoms.p, (const uint32_t *)h->d_natoms.p, (const TplMeta *)h->d_meta.p, t_base,
                h->numDet, h->stride, TAtom, g.md, (float *)h->d_X.p, need_X8 ? (float *)h->d_X8.p : nullptr, h->xpad,
                p_flags);
            h->launches++;
            CUDA_TRY(h, cudaGetLastError());
            if (need_P2) {
                tcw_prep_scan_kernel<<<dim3(TCW_NCH, cnt), TCW_PREP_THREADS, 0, st>>>(
                    (const float *)h->d_X.p, h->xpad, (const TplMeta *)h->d_meta.p, t_base, (double *)h->d_P.p, h->ppad);
                h->launches++;
                CUDA_TRY(h, cudaGetLastError());
            }
        }
        if (sb == 0) CUDA_TRY(h, cudaEventRecord(h->ev_stage[1], st));
        float *fmn = fmn_full ? fmn_full + (size_t)t_base * pcells : fmn_scratch;
        CUDA_TRY(h, cudaEventRecord(h->ev_sub[3 * sb + 0], st));
        if (path == PATH_GENERIC) {
            // few cells in the launch (MCMC steps, per-segment / cumulative / single-cell maps): one
            // warp per cell, channels across lanes; otherwise one thread per cell.  Same results.
            const bool warp_per_cell = (uint64_t)cells * (uint64_t)cnt <= TCW_GENERIC_WARP_MAX_CELLS;
            const unsigned per_block = warp_per_cell ? TCW_GENERIC_WARP_THREADS / 32 : TCW_GENERIC_THREADS;
            dim3 grid((unsigned)((cells + per_block - 1) / per_block), 1, cnt);
#define LAUNCH_GENERIC(WT, EX)                                                                              \
    do {                                                                                                    \
        if (warp_per_cell)                                                                                  \
            tcw_map_generic_warp_kernel<WT, EX><<<grid, TCW_GENERIC_WARP_THREADS, 0, st>>>(                 \
                (const float *)h->d_X.p, h->xpad, (const TplMeta *)h->d_meta.p, t_base, w, d_wins,          \
                (int)none_window, g, lut, fmn, p_maxkey,  \
                p_flags);                                                                  \
        else                                                                                                \
            tcw_map_generic_kernel<WT, EX><<<grid, TCW_GENERIC_THREADS, 0, st>>>(                           \
                (const float *)h->d_X.p, h->xpad, (const TplMeta *)h->d_meta.p, t_base, w, d_wins,          \
                (int)none_window, g, lut, fmn, p_maxkey,  \
                p_flags);                                                                  \
    } while (0)
            if (w.type == TCW_WINDOW_RECT) LAUNCH_GENERIC(TCW_WINDOW_RECT, false);
            else if (exact) LAUNCH_GENERIC(TCW_WINDOW_EXP, true);
            else LAUNCH_GENERIC(TCW_WINDOW_EXP, false);
#undef LAUNCH_GENERIC
        } else if (w.type == TCW_WINDOW_RECT) {
            const uint32_t rect_G = rp.G, DD = rp.DD, DT = rp.DT, n_reg = rp.n_reg;
            const bool rect_staged = rp.staged;
            const uint32_t rows_per_tile = TCW_RECT_WARPS * rect_G * rect_R;
            const uint32_t n_gy = (w.N_t0 + rows_per_tile - 1) / rows_per_tile;
            dim3 grid(1 + n_reg, n_gy, cnt);
            const uint32_t gx_total = 1 + n_reg;
            const size_t smem = TCW_RECT_SMEM;
            if (rect_p) {
                // one persistent CTA per SM, producer / consumer warps, tile queue over all tiles
                const uint64_t n_tiles64 = (uint64_t)cnt * n_gy * gx_total;  // 128-row tiles: same n_gy as G = 4
                if (n_tiles64 >= 0xFFFFFFFFull) return fail(h, TCW_E_INVALID, "too many rect tiles in one launch");
                const uint32_t n_tiles = (uint32_t)n_tiles64;
                const uint32_t ctas = std::min<uint32_t>((uint32_t)h->prop.multiProcessorCount, n_tiles);
                tcw_rect_map_p_kernel<<<ctas, TCW_RECTP_THREADS, TCW_RECTP_SMEM, st>>>(
                    (const double *)h->d_P.p, h->ppad, (const TplMeta *)h->d_meta.p, t_base, w, g, DD, DT, n_reg, n_gy,
                    n_tiles, p_counter + sb, fmn, p_maxkey, groupmax,
                    p_flags);
            }
#define LAUNCH_RECT(RR, STG)                                                                               \
    do {                                                                                                   \
        if (!rect_p)                                                                                       \
            tcw_rect_map_kernel<RR, STG><<<grid, TCW_RECT_THREADS, smem, st>>>(                            \
                (const double *)h->d_P.p, h->ppad, (const TplMeta *)h->d_meta.p, t_base, w, g, DD, DT,     \
                rect_G, gx_total, fmn, p_maxkey, groupmax,                      \
                p_flags);                                                                 \
        if (groupmax) {                                                                                    \
            h->launches++;                                                                                 \
            CUDA_TRY(h, cudaGetLastError());                                                               \
            tcw_rect_locate_kernel<RR, STG><<<cnt, TCW_RECT_THREADS, smem, st>>>(                          \
                (const double *)h->d_P.p, h->ppad, (const TplMeta *)h->d_meta.p, t_base, w, g, DD, DT,     \
                rect_G, gx_total, grid.y, p_maxkey, groupmax,                   \
                p_flags);                                                                 \
        }                                                                                                  \
    } while (0)
            if (rect_R == 4 && rect_staged) LAUNCH_RECT(4, true);
            else if (rect_R == 4) LAUNCH_RECT(4, false);
            else if (rect_staged) LAUNCH_RECT(1, true);
            else LAUNCH_RECT(1, false);
#undef LAUNCH_RECT
        } else if (exp_rec) {
            // correction sums of the sub-batch: group A (a2, b2, ab) then group F (Fa, Fb), TcxC
            unsigned char *c_a = (unsigned char *)h->d_C.p;
            unsigned char *c_f = c_a + (size_t)cnt * 3 * tc_rows * tc_cpitch * TcxC::kElemA;
            MapWindow wt = w;  // the tensor-core pass works on the refined rows
            wt.N_t0 = tc_rows;
            const unsigned char *corr = nullptr;
            tcw_exp_atoms_f64_kernel<<<dim3((h->xpad + 255) / 256, cnt), 256, 0, st>>>(
                (const float *)h->d_X.p, h->xpad, t_base, (double *)h->d_Xd.p);
            h->launches++;
            CUDA_TRY(h, cudaGetLastError());
            if (!exp_tc) {
                CUDA_TRY(h, cudaEventRecord(h->ev_x[2 * sb], st));
                CUDA_TRY(h, cudaEventRecord(h->ev_x[2 * sb + 1], st));
            }
            if (exp_tc) {
                const uint32_t g_elems = tc_U * 56 * (2 * tc_kc);
                const dim3 g_grid(std::min<uint32_t>((g_elems + 255u) / 256u, 1024u), cnt);
                const uint32_t n_tiles = (uint32_t)cnt * tc_n_nt * tc_n_mb;
                const uint32_t ctas = std::min<uint32_t>((uint32_t)h->prop.multiProcessorCount, n_tiles);
                tcw_exptc_scale_kernel<<<cnt, 256, 0, st>>>((const float *)h->d_X.p, h->xpad, (const TplMeta *)h->d_meta.p,
                                                            t_base, (float *)h->d_scale.p);
                h->launches++;
                if (tc_f16) {
                    tcw_exptc_atoms_kernel<true><<<g_grid, 256, 0, st>>>((const float *)h->d_X.p, h->xpad,
                                                                         (const TplMeta *)h->d_meta.p, t_base, ep.ec.i00[0],
                                                                         tc_U, (const float *)h->d_scale.p, h->d_G.p);
                } else {
                    tcw_exptc_atoms_kernel<false><<<g_grid, 256, 0, st>>>((const float *)h->d_X.p, h->xpad,
                                                                          (const TplMeta *)h->d_meta.p, t_base, ep.ec.i00[0],
                                                                          tc_U, (const float *)h->d_scale.p, h->d_G.p);
                }
                h->launches++;
                CUDA_TRY(h, cudaGetLastError());
                CUDA_TRY(h, cudaEventRecord(h->ev_x[2 * sb], st));
                if (tc_f16)
                    tcw_exptc_map_kernel<true><<<ctas, TCX_THREADS, TCX_SMEM, st>>>(
                        h->d_G.p, tc_U, h->d_W.p, tc_chunks, (const int32_t *)h->d_Kn.p, (const TplMeta *)h->d_meta.p, t_base,
                        (uint32_t)cnt, wt, ep.ec.i00[0], tc_n_nt, tc_n_mb, n_tiles, ldexpf(1.0f, -tc_shift), c_a, c_f, tc_cpitch);
                else
                    tcw_exptc_map_kernel<false><<<ctas, TCX_THREADS, TCX_SMEM, st>>>(
                        h->d_G.p, tc_U, h->d_W.p, tc_chunks, (const int32_t *)h->d_Kn.p, (const TplMeta *)h->d_meta.p, t_base,
                        (uint32_t)cnt, wt, ep.ec.i00[0], tc_n_nt, tc_n_mb, n_tiles, ldexpf(1.0f, TCX_VSCALE_LOG2 - tc_shift), c_a, c_f,
                        tc_cpitch);
                h->launches++;
                CUDA_TRY(h, cudaGetLastError());
                CUDA_TRY(h, cudaEventRecord(h->ev_x[2 * sb + 1], st));
                corr = c_a;
            }
            // row segments per column: measured (B200, 120-d maps, final kernels): with 8 templates (46 000 columns)
            // one segment is fastest (1.99 ms; 2: 2.96, 4: 2.55 -- pass 1 of the segmented walk costs more than the
            // extra warps bring); with 4 templates 2-8 segments tie.  So: segments only until the launch has ~256
            // columns per SM.
            // lookup-table mode: the columns whose windows hold only a few atoms go to the generic kernel (TCX_SHORT_K)
            uint32_t n_short = 0;
            if (exp_tc) {
                while (n_short < w.N_tau && ep.Kn[n_short] < TCX_SHORT_K) n_short++;
                if (n_short) {
                    MapWindow ws = w;
                    ws.N_tau = n_short;
                    const size_t cells_s = (size_t)w.N_t0 * n_short;
                    dim3 grid_s((unsigned)((cells_s + TCW_GENERIC_THREADS - 1) / TCW_GENERIC_THREADS), 1, cnt);
                    tcw_map_generic_kernel<TCW_WINDOW_EXP, false><<<grid_s, TCW_GENERIC_THREADS, 0, st>>>(
                        (const float *)h->d_X.p, h->xpad, (const TplMeta *)h->d_meta.p, t_base, ws, nullptr, 0, g, lut, fmn,
                        p_maxkey, p_flags, w.N_tau);
                    h->launches++;
                    CUDA_TRY(h, cudaGetLastError());
                }
            }
            int nseg = 1;
            while (nseg < 16 && (uint64_t)cnt * w.N_tau * nseg < (uint64_t)h->prop.multiProcessorCount * 256) nseg *= 2;
            if (const char *env = getenv("TCW_WALK_NSEG")) nseg = std::max(1, std::min(16, atoi(env)));
#define LAUNCH_WALK(HASC, NS)                                                                                      \
    do {                                                                                                           \
        using WC = WalkCfg<NS>;                                                                                    \
        const size_t smem = WC::smem(HASC);                                                                        \
        dim3 grid((w.N_tau + 32 * WC::kCG - 1) / (32 * WC::kCG), cnt);                                             \
        auto kern = rec_A == 1 ? tcw_exp_walk_kernel<HASC, NS, true> : tcw_exp_walk_kernel<HASC, NS, false>;        \
        kern<<<grid, WC::kThreads, smem, st>>>(                                                                    \
            (const double *)h->d_Xd.p, h->xpad, (const int32_t *)h->d_Kn.p,                                        \
            (const TplMeta *)h->d_meta.p, t_base, w,                                                               \
            ep.ec.i00[0], ep.delta[0], TAtom, (int)rec_A, corr, c_f, tc_rows, tc_cpitch,                         \
            (const float *)h->d_scale.p,                                                                         \
            ldexpf(1.0f, tc_shift), n_short, fmn, p_maxkey, p_flags);                                              \
    } while (0)
            if (exp_tc) {
                if (nseg == 1) LAUNCH_WALK(true, 1);
                else if (nseg == 2) LAUNCH_WALK(true, 2);
                else if (nseg == 4) LAUNCH_WALK(true, 4);
                else if (nseg == 8) LAUNCH_WALK(true, 8);
                else LAUNCH_WALK(true, 16);
            } else {
                if (nseg == 1) LAUNCH_WALK(false, 1);
                else if (nseg == 2) LAUNCH_WALK(false, 2);
                else if (nseg == 4) LAUNCH_WALK(false, 4);
                else if (nseg == 8) LAUNCH_WALK(false, 8);
                else LAUNCH_WALK(false, 16);
            }
#undef LAUNCH_WALK
        } else {
            dim3 grid((w.N_tau + exp_TN - 1) / exp_TN, ep.ec.ybeg[TCW_EXP_PMAX], cnt);
#define LAUNCH_EXP(CFG)                                                                                        \
    do {                                                                                                       \
        if (ep.canon)                                                                                          \
            tcw_exp_map_canon_kernel<CFG><<<grid, CFG::kThreads, CFG::kSmemC, st>>>(                          \
                (const float *)h->d_X8.p, h->xpad, (const float *)h->d_W.p, (const int32_t *)h->d_Kn.p, ep.KW, \
                (const TplMeta *)h->d_meta.p, t_base, w, ep.ec.i00[0], fmn, p_maxkey, p_flags);                \
        else if (ep.slide)                                                                                     \
            tcw_exp_map_kernel<CFG, true><<<grid, CFG::kThreads, CFG::kSmem1, st>>>(                           \
                (const float *)h->d_X8.p, h->xpad, (const float *)h->d_W.p, (const int32_t *)h->d_Kn.p, ep.KW, \
                (const TplMeta *)h->d_meta.p, (const int32_t *)h->d_shift.p, t_base, w, ep.ec, fmn, p_maxkey,  \
                p_flags);                                                                                      \
        else                                                                                                   \
            tcw_exp_map_kernel<CFG, false><<<grid, CFG::kThreads, CFG::kSmem, st>>>(                           \
                (const float *)h->d_X8.p, h->xpad, (const float *)h->d_W.p, (const int32_t *)h->d_Kn.p, ep.KW, \
                (const TplMeta *)h->d_meta.p, (const int32_t *)h->d_shift.p, t_base, w, ep.ec, fmn, p_maxkey,  \
                p_flags);                                                                                      \
    } while (0)
            if (h->exp_variant == 0) LAUNCH_EXP(ExpCfgA);
            else if (h->exp_variant == 2) LAUNCH_EXP(ExpCfgC);
            else LAUNCH_EXP(ExpCfgB);
#undef LAUNCH_EXP
        }
        h->launches++;
        CUDA_TRY(h, cudaGetLastError());
        CUDA_TRY(h, cudaEventRecord(h->ev_sub[3 * sb + 1], st));
        if (want_btsg) {
            // tiles linearised in grid.x (col tile fastest): no 65535 limit on either map dimension
            const uint32_t n_ct = (w.N_tau + TCW_BTSG_COLS - 1) / TCW_BTSG_COLS;
            const uint32_t n_rt = (w.N_t0 + TCW_BTSG_ROWS - 1) / TCW_BTSG_ROWS;
            dim3 grid(n_ct * n_rt, 1, cnt);
            const bool locate = path == PATH_FAST && w.type == TCW_WINDOW_RECT;
#define LAUNCH_BTSG(KERNEL, EX, LOC, SMEM)                                                            \
    KERNEL<EX, LOC><<<grid, TCW_BTSG_THREADS, SMEM, st>>>(                                            \
        fmn, t_base, w.N_t0, w.N_tau, w.pitch, n_ct, p_maxkey, lut, fx_scale, \
        p_rowsum, p_colsum)
            if (btsg_table) {
                const size_t smem = TCW_BTSG_TABLE_SMEM(h->lut_len);
                if (locate) LAUNCH_BTSG(tcw_btsg_table_kernel, false, true, smem);
                else LAUNCH_BTSG(tcw_btsg_table_kernel, false, false, smem);
            } else if (exact && locate) LAUNCH_BTSG(tcw_btsg_kernel, true, true, 0);
            else if (exact) LAUNCH_BTSG(tcw_btsg_kernel, true, false, 0);
            else if (locate) LAUNCH_BTSG(tcw_btsg_kernel, false, true, 0);
            else LAUNCH_BTSG(tcw_btsg_kernel, false, false, 0);
#undef LAUNCH_BTSG
            h->launches++;
            CUDA_TRY(h, cudaGetLastError());
        }
        CUDA_TRY(h, cudaEventRecord(h->ev_sub[3 * sb + 2], st));
    }
    CUDA_TRY(h, cudaEventRecord(h->ev_stage[3], st));

    // ---- stage 4: one result record per template ----
    tcw_finalize_kernel<<<T, TCW_FIN_THREADS, 0, st>>>(
        p_maxkey, p_flags,
        p_rowsum, p_colsum, 1.0 / fx_scale,
        (const TplMeta *)h->d_meta.p, w, d_wins, (int)none_window, TAtom, (int)want_btsg,
        (int)((flags & TCW_ALLOW_DEGENERATE) != 0), (uint32_t)(exp_rec ? 2 : path), (tcw_result *)h->d_results.p);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaEventRecord(h->ev_fin, st));
    h->n_sub_last = n_sub;
    h->x_valid = exp_rec;
    h->stage_valid = true;
    h->have_fmn = want_fmn;
    h->last_N_t0 = w.N_t0;
    h->last_N_tau = w.N_tau;
    h->last_pitch = w.pitch;
    return TCW_OK;
}

// Error exits of map_impl_inner may leave copies / kernels enqueued: drain both streams so that the
// caller may free or reuse its host atoms, and drop the partially uploaded / prepared batch.
static int map_impl(tcw_handle *h, const tcw_window_range *win, uint32_t flags, const tcw_atom *host_atoms,
                    const tcw_window_range *per_template = nullptr) {
    const int rc = map_impl_inner(h, win, flags, host_atoms, per_template);
    if (rc != TCW_OK && h) {
        const std::string keep = h->err;
        if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
        if (h->stream) cudaStreamSynchronize(h->stream);
        cudaGetLastError();
        if (host_atoms) h->uploaded = false;
        h->stage_valid = false;
        h->have_fmn = false;
        h->err = keep;
    }
    return rc;
}

extern "C" int tcw_map_resident(tcw_handle *h, const tcw_window_range *win, uint32_t flags) {
    return map_impl(h, win, flags, nullptr);
}

extern "C" int tcw_last_exp_stage_ms(tcw_handle *h, float ms[3]) {
    if (!h || !ms) return TCW_E_INVALID;
    ms[0] = ms[1] = ms[2] = 0.0f;
    if (!h->stage_valid) return fail(h, TCW_E_STATE, "tcw_last_exp_stage_ms: no map has been run");
    if (!h->x_valid) return TCW_OK;  // the last map did not take the recurrence path
    CUDA_TRY(h, cudaEventSynchronize(h->ev_fin));
    for (int sb = 0; sb < h->n_sub_last; sb++) {
        float a = 0, b = 0, c = 0;
        CUDA_TRY(h, cudaEventElapsedTime(&a, h->ev_sub[3 * sb + 0], h->ev_x[2 * sb]));
        CUDA_TRY(h, cudaEventElapsedTime(&b, h->ev_x[2 * sb], h->ev_x[2 * sb + 1]));
        CUDA_TRY(h, cudaEventElapsedTime(&c, h->ev_x[2 * sb + 1], h->ev_sub[3 * sb + 1]));
        ms[0] += a;
        ms[1] += b;
        ms[2] += c;
    }
    return TCW_OK;
}

extern "C" int tcw_last_stage_ms(tcw_handle *h, float ms[5]) {
    if (!h || !ms) return TCW_E_INVALID;
    if (!h->stage_valid) return fail(h, TCW_E_STATE, "tcw_last_stage_ms: no map has been run");
    CUDA_TRY(h, cudaEventSynchronize(h->ev_fin));
    // order on the stream: start [0] -> weight table [2] -> first sub-batch's merge/scan [1]
    CUDA_TRY(h, cudaEventElapsedTime(&ms[1], h->ev_stage[0], h->ev_stage[2]));
    CUDA_TRY(h, cudaEventElapsedTime(&ms[0], h->ev_stage[2], h->ev_stage[1]));
    ms[0] *= (float)h->n_sub_last;  // merge/scan runs once per sub-batch; the first one is timed
    ms[2] = ms[3] = 0.0f;
    for (int sb = 0; sb < h->n_sub_last; sb++) {
        float a = 0, b = 0;
        CUDA_TRY(h, cudaEventElapsedTime(&a, h->ev_sub[3 * sb + 0], h->ev_sub[3 * sb + 1]));
        CUDA_TRY(h, cudaEventElapsedTime(&b, h->ev_sub[3 * sb + 1], h->ev_sub[3 * sb + 2]));
        ms[2] += a;
        ms[3] += b;
    }
    CUDA_TRY(h, cudaEventElapsedTime(&ms[4], h->ev_stage[3], h->ev_fin));
    return TCW_OK;
}

extern "C" int tcw_fetch_results(tcw_handle *h, tcw_result *results) {
    if (!h || !results) return TCW_E_INVALID;
    if (!h->stage_valid) return fail(h, TCW_E_STATE, "tcw_fetch_results: no map has been run");
    CUDA_TRY(h, cudaMemcpyAsync(results, h->d_results.p, (size_t)h->T * sizeof(tcw_result),
                                cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    int worst = TCW_OK;
    for (int t = 0; t < h->T; t++) {
        if (results[t].status == TCW_E_INVALID) {
            worst = TCW_E_INVALID;
            h->err = "atoms of template " + std::to_string(t) + " are not sorted by timestamp";
        } else if (results[t].status == TCW_E_DEGENERATE && worst == TCW_OK) {
            worst = TCW_E_DEGENERATE;
            h->err = "template " + std::to_string(t) +
                     ": encountered a single-atom Fstat-calculation (i_t1 == i_t0). This is degenerate and "
                     "cannot be computed; t0 must stay away at least 2*TAtom from the end of the data";
        }
    }
    return worst;
}

extern "C" int tcw_results_device(tcw_handle *h, void **ptr, uint64_t *n_records) {
    if (!h || !ptr || !n_records) return TCW_E_INVALID;
    if (!h->stage_valid) return fail(h, TCW_E_STATE, "tcw_results_device: no map has been run");
    *ptr = h->d_results.p;
    *n_records = (uint64_t)h->T;
    return TCW_OK;
}

extern "C" int tcw_fetch_fmn(tcw_handle *h, int t, float *out) {
    if (!h || !out) return TCW_E_INVALID;
    if (!h->have_fmn) return fail(h, TCW_E_STATE, "tcw_fetch_fmn: last map did not materialise F_mn");
    if (t < 0 || t >= h->T) return fail(h, TCW_E_INVALID, "tcw_fetch_fmn: template index out of range");
    // device rows are padded to a multiple of 4 floats; the host array is dense [N_t0][N_tau]
    CUDA_TRY(h, cudaMemcpy2DAsync(out, (size_t)h->last_N_tau * sizeof(float),
                                  (const float *)h->d_Fmn.p + (size_t)t * h->last_N_t0 * h->last_pitch,
                                  (size_t)h->last_pitch * sizeof(float), (size_t)h->last_N_tau * sizeof(float),
                                  h->last_N_t0, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return TCW_OK;
}

extern "C" int tcw_fetch_merged(tcw_handle *h, int t, float *out, uint32_t capacityN) {
    if (!h || !out) return TCW_E_INVALID;
    if (!h->stage_valid) return fail(h, TCW_E_STATE, "tcw_fetch_merged: no map has been run");
    if (t < 0 || t >= h->T) return fail(h, TCW_E_INVALID, "tcw_fetch_merged: template index out of range");
    const uint32_t N = h->meta[t].numAtoms;
    if (capacityN < N) return fail(h, TCW_E_INVALID, "tcw_fetch_merged: capacity too small");
    for (int c = 0; c < TCW_NCH; c++)
        CUDA_TRY(h, cudaMemcpyAsync(out + (size_t)c * capacityN,
                                    (const float *)h->d_X.p + ((size_t)t * TCW_NCH + c) * h->xpad,
                                    (size_t)N * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return TCW_OK;
}

extern "C" int tcw_map_batch(tcw_handle *h, const tcw_atom *atoms, const uint32_t *n_atoms,
                             uint32_t atom_stride, uint32_t TAtom, int T, int numDet,
                             const tcw_window_range *win, uint32_t flags, float *F_mn_out,
                             tcw_result *results) {
    if (!h) return TCW_E_INVALID;
    if (!results) return fail(h, TCW_E_INVALID, "tcw_map_batch: results is NULL");
    if ((flags & TCW_WANT_FMN) && !F_mn_out)
        return fail(h, TCW_E_INVALID, "tcw_map_batch: TCW_WANT_FMN needs F_mn_out");
    int rc = upload_common(h, atoms, n_atoms, atom_stride, TAtom, T, numDet, false);
    if (rc) return rc;
    rc = map_impl(h, win, flags, atoms);
    if (rc) return rc;
    if (flags & TCW_WANT_FMN) {
        CUDA_TRY(h, cudaMemcpy2DAsync(F_mn_out, (size_t)h->last_N_tau * sizeof(float), h->d_Fmn.p,
                                      (size_t)h->last_pitch * sizeof(float), (size_t)h->last_N_tau * sizeof(float),
                                      (size_t)T * h->last_N_t0, cudaMemcpyDeviceToHost, h->stream));
    }
    return tcw_fetch_results(h, results);
}

extern "C" int tcw_submit(tcw_handle *h, const tcw_atom *atoms, const uint32_t *n_atoms, uint32_t atom_stride,
                          uint32_t TAtom, int T, int numDet, const tcw_window_range *win, uint32_t flags) {
    if (!h) return TCW_E_INVALID;
    if (h->in_flight) return fail(h, TCW_E_STATE, "tcw_submit: a batch is already in flight (call tcw_wait)");
    int rc = upload_common(h, atoms, n_atoms, atom_stride, TAtom, T, numDet, false);
    if (rc) return rc;
    rc = map_impl(h, win, flags, atoms);
    if (rc) return rc;
    h->in_flight = true;
    h->in_flight_fmn = (flags & TCW_WANT_FMN) != 0;
    return TCW_OK;
}

extern "C" int tcw_wait(tcw_handle *h, float *F_mn_out, tcw_result *results) {
    if (!h) return TCW_E_INVALID;
    if (!h->in_flight) return fail(h, TCW_E_STATE, "tcw_wait: no batch in flight (call tcw_submit)");
    if (!results) return fail(h, TCW_E_INVALID, "tcw_wait: results is NULL");
    if (h->in_flight_fmn && !F_mn_out) return fail(h, TCW_E_INVALID, "tcw_wait: the batch was submitted with TCW_WANT_FMN");
    h->in_flight = false;
    if (h->in_flight_fmn) {
        CUDA_TRY(h, cudaMemcpy2DAsync(F_mn_out, (size_t)h->last_N_tau * sizeof(float), h->d_Fmn.p,
                                      (size_t)h->last_pitch * sizeof(float), (size_t)h->last_N_tau * sizeof(float),
                                      (size_t)h->T * h->last_N_t0, cudaMemcpyDeviceToHost, h->stream));
    }
    return tcw_fetch_results(h, results);
}

extern "C" int tcw_map_batch_windows(tcw_handle *h, const tcw_atom *atoms, const uint32_t *n_atoms,
                                     uint32_t atom_stride, uint32_t TAtom, int T, int numDet,
                                     const tcw_window_range *wins, uint32_t flags, float *F_mn_out,
                                     tcw_result *results) {
    if (!h) return TCW_E_INVALID;
    if (!results || !wins) return fail(h, TCW_E_INVALID, "tcw_map_batch_windows: results / wins is NULL");
    if ((flags & TCW_WANT_FMN) && !F_mn_out)
        return fail(h, TCW_E_INVALID, "tcw_map_batch_windows: TCW_WANT_FMN needs F_mn_out");
    int rc = upload_common(h, atoms, n_atoms, atom_stride, TAtom, T, numDet, false);
    if (rc) return rc;
    rc = map_impl(h, &wins[0], flags, atoms, wins);
    if (rc) return rc;
    if (flags & TCW_WANT_FMN) {
        CUDA_TRY(h, cudaMemcpy2DAsync(F_mn_out, (size_t)h->last_N_tau * sizeof(float), h->d_Fmn.p,
                                      (size_t)h->last_pitch * sizeof(float), (size_t)h->last_N_tau * sizeof(float),
                                      (size_t)T * h->last_N_t0, cudaMemcpyDeviceToHost, h->stream));
    }
    return tcw_fetch_results(h, results);
}

// ---------------------------------------------------------------------------------------
// SIMT peak microbenchmarks (roofline denominators not in MEASURED_PEAKS.json)
// ---------------------------------------------------------------------------------------
__global__ void tcw_ffma_peak_kernel(float *out, int iters) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3f + i;
    const float b = 1.0000001f, c = 1e-7f;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = fmaf(a[i], b, c);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i];
    if (s == 123.456f) out[0] = s;
}
__global__ void tcw_dadd_peak_kernel(double *out, int iters) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
    const double c = 1e-7;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = __dadd_rn(a[i], c);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i];
    if (s == 123.456) out[0] = s;
}

// packed FP32 (Blackwell FFMA2: two FMAs per lane per instruction)
__global__ void tcw_ffma2_peak_kernel(float *out, int iters) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
    const float2 b = make_float2(1.0000001f, 0.9999999f), c = make_float2(1e-7f, -1e-7f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = __ffma2_rn(a[i], b, c);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i].x + a[i].y;
    if (s == 123.456f) out[0] = s;
}

extern "C" int tcw_microbench_ffma2(tcw_handle *h, double *ffma2_tflops) {
    if (!h || !ffma2_tflops) return TCW_E_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc = ensure(h, h->d_scratch, 1024);
    if (rc) return rc;
    const int blocks = h->prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    float ms = 0;
    for (int rep = 0; rep < 2; rep++) {
        CUDA_TRY(h, cudaEventRecord(h->ev_timer[0], h->stream));
        tcw_ffma2_peak_kernel<<<blocks, threads, 0, h->stream>>>((float *)h->d_scratch.p, iters);
        CUDA_TRY(h, cudaEventRecord(h->ev_timer[1], h->stream));
        CUDA_TRY(h, cudaEventSynchronize(h->ev_timer[1]));
        CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev_timer[0], h->ev_timer[1]));
    }
    *ffma2_tflops = 4.0 * 64.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    h->launches += 2;
    return TCW_OK;
}

extern "C" int tcw_microbench(tcw_handle *h, double *ffma_tflops, double *dadd_tflops) {
    if (!h || !ffma_tflops || !dadd_tflops) return TCW_E_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc = ensure(h, h->d_scratch, 1024);
    if (rc) return rc;
    const int blocks = h->prop.multiProcessorCount * 8, threads = 256;
    const int iters_f = 4096, iters_d = 1024;
    float ms = 0;
    for (int rep = 0; rep < 2; rep++) {  // first rep warms up
        CUDA_TRY(h, cudaEventRecord(h->ev_timer[0], h->stream));
        tcw_ffma_peak_kernel<<<blocks, threads, 0, h->stream>>>((float *)h->d_scratch.p, iters_f);
        CUDA_TRY(h, cudaEventRecord(h->ev_timer[1], h->stream));
        CUDA_TRY(h, cudaEventSynchronize(h->ev_timer[1]));
        CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev_timer[0], h->ev_timer[1]));
    }
    *ffma_tflops = 2.0 * 64.0 * iters_f * (double)blocks * threads / (ms * 1e-3) / 1e12;
    for (int rep = 0; rep < 2; rep++) {
        CUDA_TRY(h, cudaEventRecord(h->ev_timer[0], h->stream));
        tcw_dadd_peak_kernel<<<blocks, threads, 0, h->stream>>>((double *)h->d_scratch.p, iters_d);
        CUDA_TRY(h, cudaEventRecord(h->ev_timer[1], h->stream));
        CUDA_TRY(h, cudaEventSynchronize(h->ev_timer[1]));
        CUDA_TRY(h, cudaEventElapsedTime(&ms, h->ev_timer[0], h->ev_timer[1]));
    }
    *dadd_tflops = 64.0 * iters_d * (double)blocks * threads / (ms * 1e-3) / 1e12;
    h->launches += 4;
    return TCW_OK;
}

#ifdef TCX_TIMING
// development build only (-DTCX_TIMING): role wait cycles of the tensor-core pass, see tcw_exp_rec.cuh
extern "C" int tcw_debug_tcx_timing(unsigned long long *out16, int reset) {
    if (out16 && cudaMemcpyFromSymbol(out16, tcx_timing, sizeof(unsigned long long) * 16) != cudaSuccess) return -1;
    if (reset) {
        unsigned long long z[16] = {0};
        if (cudaMemcpyToSymbol(tcx_timing, z, sizeof(z)) != cudaSuccess) return -1;
    }
    return 0;
}
#endif
