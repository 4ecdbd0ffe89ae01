// C-ABI layer of libysmr_b200.so (include/ysmr_b200.h): context, device buffers, kernel orchestration, streams.
// No torch types, no exceptions across the boundary.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/ysmr_b200.h"
#include "frontend.cuh"
#include "kernels.cuh"
#include "link.cuh"

using namespace ysmr;

static_assert(sizeof(ysmr_row) == 40, "ysmr_row layout");
static_assert(sizeof(RowOut) == sizeof(ysmr_row), "RowOut must mirror ysmr_row");
static_assert(YSMR_MAX_FILTERS == LINK_MAX_FILTERS && YSMR_MAX_HORIZON == LINK_MAX_HORIZON, "filter limits");

namespace {

thread_local std::string g_create_error;

// cv2.getGaussianKernel(11, 0, CV_32F): sigma = 0.3*((11-1)*0.5 - 1) + 0.8 = 2.0.  Bit patterns taken from OpenCV 4.13
// (tests/test_oracle_stages.py checks them against the installed cv2).
const uint32_t kGaussBits[11] = {0x3c10612bu, 0x3cde5c35u, 0x3d855a85u, 0x3df92326u, 0x3e353f0fu, 0x3e4d6105u,
                                 0x3e353f0fu, 0x3df92326u, 0x3d855a85u, 0x3cde5c35u, 0x3c10612bu};

struct Control {             // small device block, zeroed per detect call where noted
    uint32_t work_count;     // per call
    uint32_t big_count;      // per call
    unsigned long long pool_used;  // per call
    int32_t status;          // sticky until ysmr_status
    int32_t first_bad;       // sticky until ysmr_status
};

}  // namespace

constexpr int GATE_FLAGS = 1024;     // chunks per ysmr_track_* call that can use the gated linker launch

struct ysmr_ctx {
    int device = 0, h = 0, w = 0, ww = 0, channels = 1;
    ysmr_params p{};
    int mode = YSMR_MODE_ADAPTIVE_DOUBLE;
    int propagate = 0;            // K2 mode
    int img_is_marker = 0;        // DIRECT on the marker image (dark-on-light quirk)
    int t_mask = 0, t_marker = 0, inverted = 0, signed_offset = 0;
    int window = 0;               // mean/std moving window (frames)
    int32_t *gate_flags = nullptr;   // [GATE_FLAGS] "detections of chunk k are ready" (track_chunks / LinkGate)
    int link_gate = 1;               // YSMR_LINK=...,nogate: sequence linker launches behind detection with events only
    int link_fast = 1;            // YSMR_LINK=general disables the lane fast path of the linker; ,grid / ,cta pick the general kernel (tests)
    int frontend_gen = 4;         // 3: force the three-kernel front-end (ysmr_set_option, A/B measurements in bench.py)
    cudaStream_t s_tail = nullptr; cudaEvent_t ev_tail_fork = nullptr, ev_tail_join = nullptr;   // K1b tail-strip launch
    uint8_t *plane = nullptr; int64_t plane_stride = 0; int pitch = 0;          // blurred planes (K1a -> K1b)
    uint8_t *decisions = nullptr; int64_t dec_stride = 0; int dec_pitch = 0;    // decision bytes (K1b -> K1c)
    std::string err;
    int64_t launches = 0;

    // detection buffers
    uint32_t *mask_bits = nullptr, *marker_bits = nullptr;
    uint8_t *label_scratch = nullptr; size_t label_stride = 0; int label_grid = 0;
    uint32_t *first_xy = nullptr, *counts = nullptr, *work = nullptr;
    uint32_t *big_items = nullptr; int big_cap = 0;
    uint8_t *pool = nullptr; unsigned long long pool_bytes = 0;
    Control *ctl = nullptr;
    unsigned long long *sums = nullptr; double *values = nullptr; int32_t *scalar_thr = nullptr;
    // linker
    LinkConfig lc{};
    LinkState ls{};
    LinkScratch lx{};
    FrameScratch lf{};
    long long *phase_cycles = nullptr;
    double *gain_dev[LINK_MAX_FILTERS] = {nullptr, nullptr, nullptr, nullptr};
    double *exp_tab_dev = nullptr;
    bool gain_set[LINK_MAX_FILTERS] = {false, false, false, false};
    std::vector<std::pair<void *, size_t>> state_parts;   // for export/import
    // pipeline (ysmr_track_*)
    cudaStream_t s_copy = nullptr, s_det = nullptr, s_link = nullptr;
    cudaEvent_t ev_copy[2]{}, ev_det[2]{}, ev_link[2]{}, ev_fork = nullptr, ev_join = nullptr;
    uint8_t *stage[2] = {nullptr, nullptr}; size_t stage_bytes = 0;
    int32_t *pipe_count[2] = {nullptr, nullptr};
    float *pipe_blobs[2] = {nullptr, nullptr};
    ysmr_row *rows_dev = nullptr; int64_t rows_dev_cap = 0;
    // device row archive (ysmr_rows_archive / ysmr_rows_sorted)
    int archive_on = 0;
    ysmr_row *arch = nullptr; int64_t arch_n = 0, arch_cap = 0;
    long long *n_rows_dev = nullptr;
    std::vector<void *> allocs;
    // optional per-kernel timing (ysmr_set_profiling): event pairs recorded around every launch on its own stream
    int profiling = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events[YSMR_PROF_KINDS];
    std::vector<cudaEvent_t> prof_free;
};

namespace {

int fail(ysmr_ctx *c, int code, const std::string &msg)
{
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}

#define CU(c, expr)                                                                                     \
    do {                                                                                                \
        cudaError_t e__ = (expr);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return fail(c, YSMR_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));           \
    } while (0)

template <class T>
cudaError_t dev_alloc(ysmr_ctx *c, T **out, size_t count)
{
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T));
    if (e == cudaSuccess) { c->allocs.push_back(p); *out = (T *)p; }
    return e;
}

cudaEvent_t prof_event(ysmr_ctx *c)
{
    cudaEvent_t e = nullptr;
    if (!c->prof_free.empty()) { e = c->prof_free.back(); c->prof_free.pop_back(); }
    else cudaEventCreate(&e);
    return e;
}

struct ProfScope {          // records start now and stop at scope exit on `st`, filed under `kind`
    ysmr_ctx *c; int kind; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(ysmr_ctx *c_, int kind_, cudaStream_t st_) : c(c_), kind(kind_), st(st_)
    {
        if (c->profiling) { a = prof_event(c); b = prof_event(c); cudaEventRecord(a, st); }
    }
    ~ProfScope()
    {
        if (a) { cudaEventRecord(b, st); c->prof_events[kind].push_back({a, b}); }
    }
};

int derive_thresholds(ysmr_ctx *c)
{
    const ysmr_params &p = c->p;
    c->inverted = p.white_on_dark ? 0 : 1;
    c->signed_offset = p.white_on_dark ? p.offset : -p.offset;       // track_eval.py:132
    if (p.adt > 0) c->mode = YSMR_MODE_ADAPTIVE_DOUBLE;
    else if (p.adt == 0) c->mode = YSMR_MODE_ADAPTIVE_SINGLE;
    else c->mode = YSMR_MODE_MEAN_STD;
    // cv2.adaptiveThreshold: idelta = BINARY ? ceil(C) : floor(C); BINARY: d > -idelta ; BINARY_INV: d <= -idelta
    const double c1 = (double)(c->signed_offset * -1);                           // track_eval.py:196
    const double c2 = ((double)c->signed_offset + p.adt) * -1.0;                 // track_eval.py:206-207
    c->t_mask = p.white_on_dark ? -(int)ceil(c1) : -(int)floor(c1);
    c->t_marker = p.white_on_dark ? -(int)ceil(c2) : -(int)floor(c2);
    c->propagate = 0; c->img_is_marker = 0;
    if (c->mode == YSMR_MODE_ADAPTIVE_DOUBLE) {
        // markers inside mask -> true propagation; otherwise binary_propagation returns the marker image (finding 8)
        const bool marker_in_mask = p.white_on_dark ? (c->t_marker >= c->t_mask) : (c->t_marker <= c->t_mask);
        if (marker_in_mask) c->propagate = 1; else c->img_is_marker = 1;
    }
    c->window = (int)floor(5.0 * p.fps) + 1;                                      // track_eval.py:235-238
    return YSMR_OK;
}

}  // namespace

extern "C" {

int ysmr_abi_version(void) { return YSMR_ABI_VERSION; }

const char *ysmr_last_error(const ysmr_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

void ysmr_default_params(ysmr_params *p)
{
    memset(p, 0, sizeof(*p));
    p->white_on_dark = 1; p->offset = 5; p->adt = 2.0; p->fps = 30.0;       // helper_file.py:160-282
    p->use_gsff = 1; p->n_f = 3; p->n_min = 0; p->n_max = 30;
    p->max_blobs = 4096; p->max_tracks = 8192; p->max_runs = 32768; p->max_batch = 256;
    p->max_distance = 0.0;
}

int ysmr_create(ysmr_ctx **out, int device, int height, int width, int channels, const ysmr_params *params)
{
    if (!out || !params) return fail(nullptr, YSMR_E_INVALID, "null argument");
    *out = nullptr;
    if (height < 16 || width < 16 || height > 32768 || width > 32768)
        return fail(nullptr, YSMR_E_INVALID, "frame size must be within 16..32768 in both dimensions");
    if (channels != 1 && channels != 3) return fail(nullptr, YSMR_E_INVALID, "channels must be 1 (grey) or 3 (BGR)");
    if (params->max_blobs < 1 || params->max_blobs > 32768 || params->max_tracks < 1 || params->max_runs < 16 ||
        params->max_batch < 1 || params->max_batch > 65535)
        return fail(nullptr, YSMR_E_INVALID, "capacities out of range");
    if (!(params->fps > 0)) return fail(nullptr, YSMR_E_INVALID, "fps must be positive");
    if (params->use_gsff && (params->n_f < 1 || params->n_f > YSMR_MAX_FILTERS))
        return fail(nullptr, YSMR_E_INVALID, "number of LSFFs out of range");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, YSMR_E_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, YSMR_E_INVALID, "bad device index");
    ysmr_ctx *c = new ysmr_ctx();
    c->device = device; c->h = height; c->w = width; c->ww = (width + 31) / 32; c->channels = channels; c->p = *params;
    derive_thresholds(c);
    { const char *lk = getenv("YSMR_LINK"); c->link_fast = !(lk && strstr(lk, "general")); c->link_gate = !(lk && strstr(lk, "nogate")); }
#define CC(expr)                                                                                                       \
    do {                                                                                                               \
        cudaError_t e__ = (expr);                                                                                      \
        if (e__ != cudaSuccess) {                                                                                      \
            fail(nullptr, YSMR_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));                           \
            ysmr_destroy(c);                                                                                           \
            return YSMR_E_CUDA;                                                                                        \
        }                                                                                                              \
    } while (0)
    CC(cudaSetDevice(device));
    CC(fused_frontend_init());                     // per-device kernel attributes (shared-memory opt-in)
    CC(link_kernel_init());
    const size_t B = (size_t)params->max_batch, H = (size_t)height, WW = (size_t)c->ww, MB = (size_t)params->max_blobs;
    CC(dev_alloc(c, &c->mask_bits, B * H * WW));
    CC(dev_alloc(c, &c->marker_bits, B * H * WW));
    c->pitch = 16 + ((width + 127) / 128) * 128;
    c->plane_stride = (int64_t)(height + 10) * c->pitch;
    CC(dev_alloc(c, &c->plane, B * (size_t)c->plane_stride));
    CC(cudaMemset(c->plane, 0, B * (size_t)c->plane_stride));
    c->dec_pitch = 32 * ((width + 127) / 128);
    c->dec_stride = (int64_t)height * c->dec_pitch;
    CC(dev_alloc(c, &c->decisions, B * (size_t)c->dec_stride));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    c->label_grid = (int)std::min<size_t>(B, (size_t)sms * 6);
    c->label_stride = label_scratch_bytes(height, params->max_runs);
    CC(dev_alloc(c, &c->label_scratch, c->label_stride * (size_t)c->label_grid));
    CC(dev_alloc(c, &c->first_xy, B * MB));
    CC(dev_alloc(c, &c->counts, B * 4));
    CC(dev_alloc(c, &c->work, B * MB));
    c->big_cap = 4096;
    CC(dev_alloc(c, &c->big_items, (size_t)c->big_cap * 2));
    c->pool_bytes = 64ull << 20;
    CC(dev_alloc(c, &c->pool, (size_t)c->pool_bytes));
    CC(dev_alloc(c, &c->ctl, 1));
    Control init{}; init.first_bad = 0x7fffffff;
    CC(cudaMemcpy(c->ctl, &init, sizeof(init), cudaMemcpyHostToDevice));
    CC(dev_alloc(c, &c->sums, B * 2));
    CC(dev_alloc(c, &c->values, (size_t)c->window + B));
    CC(cudaMemset(c->values, 0, sizeof(double) * ((size_t)c->window + B)));
    CC(dev_alloc(c, &c->scalar_thr, B));

    // linker
    LinkConfig &lc = c->lc;
    lc.max_disappeared = params->fps; lc.max_distance = params->max_distance;
    lc.use_gsff = params->use_gsff; lc.n_f = params->use_gsff ? params->n_f : 0;
    lc.max_tracks = params->max_tracks; lc.max_blobs = params->max_blobs; lc.cross_zero = 1; lc.xy_same = 1;
    if (params->use_gsff) {
        const double step = (double)(params->n_max - params->n_min) / (double)params->n_f;      // gsff.py:103-106
        for (int i = 0; i < params->n_f; ++i) {
            lc.n_i[i] = (int)((double)params->n_min + step * (double)(i + 1));
            // explicit horizons (ysmr_params::reserved): the host evaluated the reference's expression on its operand types
            if (params->reserved[0] > 0 && i < 4) lc.n_i[i] = params->reserved[i];
            if (lc.n_i[i] < 1 || lc.n_i[i] > YSMR_MAX_HORIZON) {
                fail(nullptr, YSMR_E_INVALID, "GSFF horizon out of range (1..64)");
                ysmr_destroy(c);
                return YSMR_E_INVALID;
            }
        }
        lc.hist_len = lc.n_i[params->n_f - 1] + 1;
    } else {
        lc.hist_len = 1;
    }
    const size_t T = (size_t)params->max_tracks;
    LinkState &ls = c->ls;
    auto part = [&](auto **ptr, size_t count) -> cudaError_t {
        cudaError_t e2 = dev_alloc(c, ptr, count);
        if (e2 == cudaSuccess) c->state_parts.push_back({(void *)*ptr, count * sizeof(**ptr)});
        return e2;
    };
    CC(part(&ls.hdr, 8));
    CC(part(&ls.order[0], T)); CC(part(&ls.order[1], T)); CC(part(&ls.free_slots, T));
    CC(part(&ls.id, T)); CC(part(&ls.px, T)); CC(part(&ls.py, T));
    CC(part(&ls.iw, T)); CC(part(&ls.ih, T)); CC(part(&ls.ideg, T));
    CC(part(&ls.gone, T)); CC(part(&ls.mode, T)); CC(part(&ls.hist_n, T));
    CC(part(&ls.hist, T * (size_t)lc.hist_len * 2));
    CC(part(&ls.wgt, T * LINK_MAX_FILTERS)); CC(part(&ls.xh, T * LINK_MAX_FILTERS * 2));
    for (auto &pr : c->state_parts) CC(cudaMemset(pr.first, 0, pr.second));
    LinkScratch &lx = c->lx;
    FrameScratch &lf = c->lf;                  // general path: global-memory home of the per-frame scratch
    CC(dev_alloc(c, &lf.col_best, MB)); CC(dev_alloc(c, &lf.col_row, MB)); CC(dev_alloc(c, &lf.dxy, MB));
    CC(dev_alloc(c, &lf.cell_items, MB)); CC(dev_alloc(c, &lf.cell_start, (size_t)LINK_GRID_CELLS + 2)); CC(dev_alloc(c, &lf.flags, 4));
    CC(dev_alloc(c, &lx.lane_done, 1));
    CC(cudaMemset(lx.lane_done, 0, sizeof(int32_t)));
    CC(dev_alloc(c, &lx.grid_ws, 8));
    CC(cudaMemset(lx.grid_ws, 0, 8 * sizeof(int32_t)));
    {
        // detection grid of the general path: at most 32 x 32 cells of at least 16 pixels covering [0, W] x [0, H]
        const int ext = (width > height ? width : height) + 1;
        int cell = (ext + 31) / 32; if (cell < 16) cell = 16;
        lc.grid_cell = (double)cell; lc.grid_w = width / cell + 1; lc.grid_h = height / cell + 1;
    }
    CC(dev_alloc(c, &lx.row_min, T)); CC(dev_alloc(c, &lx.row_arg, T));
    CC(dev_alloc(c, &lx.flag, std::max(T, MB) + 2)); CC(dev_alloc(c, &lx.list, MB));
    lx.set_table_size = set_table_capacity((int)MB);
    CC(dev_alloc(c, &lx.table, (size_t)lx.set_table_size));
    lx.prep_frames = 2048;           // = FAST_FRAMES of link.cu
    CC(dev_alloc(c, &lx.succ, (size_t)2 * lx.prep_frames * 256));          // two halves: pipelined launches alternate (LinkGate)
    CC(dev_alloc(c, &lx.thr2, (size_t)2 * lx.prep_frames * 256));
    CC(dev_alloc(c, &c->gate_flags, (size_t)GATE_FLAGS));
    CC(cudaMemset(c->gate_flags, 0, GATE_FLAGS * sizeof(int32_t)));
    lx.prep_margin = 2.0e-3f + 1.0e-6f * (float)(height + width);
    CC(dev_alloc(c, &c->phase_cycles, 16));
    CC(cudaMemset(c->phase_cycles, 0, 16 * sizeof(long long)));
    lx.phase_cycles = nullptr;
    for (int i = 0; i < lc.n_f; ++i) CC(dev_alloc(c, &c->gain_dev[i], (size_t)4 * lc.n_i[i]));
    for (int i = 0; i < LINK_MAX_FILTERS; ++i) lc.gain[i] = c->gain_dev[i];
    {
        double tab[NP_EXP_TABLE];
        for (int k = 0; k < NP_EXP_TABLE; ++k) { const unsigned long long u = np_exp_table_bits(k); memcpy(&tab[k], &u, 8); }
        CC(dev_alloc(c, &c->exp_tab_dev, (size_t)NP_EXP_TABLE));
        CC(cudaMemcpy(c->exp_tab_dev, tab, sizeof(tab), cudaMemcpyHostToDevice));
        lc.exp_tab = c->exp_tab_dev;
    }
    CC(launch_link_reset(ls, params->max_tracks, nullptr)); c->launches++;
    CC(cudaDeviceSynchronize());
    // pipeline objects
    CC(cudaStreamCreateWithFlags(&c->s_tail, cudaStreamNonBlocking));
    CC(cudaEventCreateWithFlags(&c->ev_tail_fork, cudaEventDisableTiming));
    CC(cudaEventCreateWithFlags(&c->ev_tail_join, cudaEventDisableTiming));
    CC(cudaStreamCreateWithFlags(&c->s_copy, cudaStreamNonBlocking));
    CC(cudaStreamCreateWithFlags(&c->s_det, cudaStreamNonBlocking));
    {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);          // the serial linker goes first whenever it is ready
        CC(cudaStreamCreateWithPriority(&c->s_link, cudaStreamNonBlocking, hi));
    }
    for (int i = 0; i < 2; ++i) {
        CC(cudaEventCreateWithFlags(&c->ev_copy[i], cudaEventDisableTiming));
        CC(cudaEventCreateWithFlags(&c->ev_det[i], cudaEventDisableTiming));
        CC(cudaEventCreateWithFlags(&c->ev_link[i], cudaEventDisableTiming));
        CC(dev_alloc(c, &c->pipe_count[i], B));
        CC(dev_alloc(c, &c->pipe_blobs[i], B * MB * 5));
    }
    CC(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CC(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    CC(dev_alloc(c, &c->n_rows_dev, 1));
#undef CC
    *out = c;
    return YSMR_OK;
}

int ysmr_destroy(ysmr_ctx *c)
{
    if (!c) return YSMR_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (void *p : c->allocs) cudaFree(p);
    for (int k = 0; k < YSMR_PROF_KINDS; ++k)
        for (auto &pr : c->prof_events[k]) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    for (cudaEvent_t e : c->prof_free) cudaEventDestroy(e);
    if (c->rows_dev) cudaFree(c->rows_dev);
    if (c->arch) cudaFree(c->arch);
    for (int i = 0; i < 2; ++i) {
        if (c->stage[i]) cudaFree(c->stage[i]);
        if (c->ev_copy[i]) cudaEventDestroy(c->ev_copy[i]);
        if (c->ev_det[i]) cudaEventDestroy(c->ev_det[i]);
        if (c->ev_link[i]) cudaEventDestroy(c->ev_link[i]);
    }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->ev_tail_fork) cudaEventDestroy(c->ev_tail_fork);
    if (c->ev_tail_join) cudaEventDestroy(c->ev_tail_join);
    if (c->s_tail) cudaStreamDestroy(c->s_tail);
    if (c->s_copy) cudaStreamDestroy(c->s_copy);
    if (c->s_det) cudaStreamDestroy(c->s_det);
    if (c->s_link) cudaStreamDestroy(c->s_link);
    delete c;
    return YSMR_OK;
}

int ysmr_set_gsff_gain(ysmr_ctx *c, int filter, int horizon, const double *h_gain)
{
    if (!c || !h_gain) return fail(c, YSMR_E_INVALID, "null argument");
    if (!c->p.use_gsff) return fail(c, YSMR_E_STATE, "gsff is disabled in this context");
    if (filter < 0 || filter >= c->lc.n_f) return fail(c, YSMR_E_INVALID, "filter index out of range");
    if (horizon != c->lc.n_i[filter]) return fail(c, YSMR_E_INVALID, "horizon does not match generate_n_i");
    const int n = horizon;
    std::vector<double> g(4 * (size_t)n);
    for (int k = 0; k < n; ++k) {                 // rows 0 (x) and 1 (y) of the 4 x 2n gain (gsff.py:240)
        g[k] = h_gain[2 * k]; g[n + k] = h_gain[2 * k + 1];
        g[2 * n + k] = h_gain[2 * n + 2 * k]; g[3 * n + k] = h_gain[2 * n + 2 * k + 1];
        if (g[n + k] != 0.0 || g[2 * n + k] != 0.0) c->lc.cross_zero = 0;
        if (g[k] != g[3 * n + k]) c->lc.xy_same = 0;      // (the fast linker loads one tap for both axes)
    }
    if (filter < 3 && n <= 30)
        for (int k = 0; k < n; ++k) c->lc.fast_gain[filter][k] = g[k];
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaMemcpy(c->gain_dev[filter], g.data(), sizeof(double) * g.size(), cudaMemcpyHostToDevice));
    c->gain_set[filter] = true;
    return YSMR_OK;
}

int ysmr_detect(ysmr_ctx *c, const uint8_t *d_frames, int n_frames, int64_t frame_stride, int first_frame,
                int32_t *d_blob_count, float *d_blobs, const ysmr_debug_out *dbg, void *stream)
{
    if (!c || !d_frames || !d_blob_count || !d_blobs) return fail(c, YSMR_E_INVALID, "null argument");
    if (n_frames < 0 || n_frames > c->p.max_batch) return fail(c, YSMR_E_INVALID, "n_frames exceeds max_batch");
    if (frame_stride < (int64_t)c->h * c->w * c->channels) return fail(c, YSMR_E_INVALID, "frame_stride too small");
    if (n_frames == 0) return YSMR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    CU(c, cudaSetDevice(c->device));
    FrontParams fp{};
    fp.frames = d_frames; fp.frame_stride = frame_stride; fp.n_frames = n_frames;
    fp.h = c->h; fp.w = c->w; fp.ww = c->ww; fp.channels = c->channels;
    fp.t_mask = c->t_mask; fp.t_marker = c->t_marker; fp.inverted = c->inverted;
    fp.mask_bits = c->mask_bits;
    fp.marker_bits = c->mode == YSMR_MODE_ADAPTIVE_DOUBLE ? c->marker_bits : nullptr;
    memcpy(fp.k, kGaussBits, sizeof(fp.k));
    fp.row_tail_from = c->w - c->w % 4; fp.col_tail_from = c->w - c->w % 8;
    fp.plane = c->plane; fp.plane_stride = c->plane_stride; fp.pitch = c->pitch;
    fp.decisions = c->decisions; fp.dec_stride = c->dec_stride; fp.dec_pitch = c->dec_pitch;
    if (dbg) { fp.dbg_grey = dbg->d_grey; fp.dbg_blurred = dbg->d_blurred; fp.dbg_mean = dbg->d_mean; }
    if (c->mode == YSMR_MODE_MEAN_STD) {
        CU(c, launch_frame_moments(d_frames, frame_stride, n_frames, c->h, c->w, c->channels, c->sums, st));
        CU(c, launch_moving_threshold(c->sums, n_frames, c->h, c->w, c->p.white_on_dark, c->signed_offset, first_frame,
                                      c->window, c->values, c->scalar_thr, st));
        c->launches += 3;
        fp.scalar_thr = c->scalar_thr;
        if (dbg && dbg->d_scalar_thr)
            CU(c, cudaMemcpyAsync(dbg->d_scalar_thr, c->scalar_thr, sizeof(int32_t) * n_frames, cudaMemcpyDeviceToDevice, st));
    }
    const bool fused = c->frontend_gen != 3 && fused_frontend_supported(fp);
    if (dbg && (dbg->d_mean || (!fused && (dbg->d_grey || dbg->d_blurred)))) {
        // debug tile kernel: the only one that can dump the rounded Gaussian mean (and grey / blurred when the production
        // front-end is the three-kernel one).  Its masks are overwritten by the production kernel(s) below, so the mask
        // dumps always come from the production path.
        FrontParams ft = fp;
        if (fused) { ft.dbg_grey = nullptr; ft.dbg_blurred = nullptr; }
        CU(c, launch_frontend_tile(ft, st)); c->launches++;
    }
    if (fused) {
        // generation 4: one fused kernel (fused.cu); grey / blurred dumps come from its own shared-memory tiles
        ProfScope ps(c, YSMR_PROF_FRONTEND, st);
        CU(c, launch_fused_frontend(fp, st)); c->launches += 1;
    } else {
        // generation 3 (three kernels through a padded blurred plane): mean/std mode, widths that are not a multiple of 4,
        // unaligned frames
        ProfScope ps(c, YSMR_PROF_FRONTEND, st);
        { ProfScope p1(c, YSMR_PROF_K1A, st); CU(c, launch_blur_prepass(fp, st)); c->launches += 2; }
        {
            ProfScope p2(c, YSMR_PROF_K1B, st);
            int nl = 0;
            CU(c, cudaEventRecord(c->ev_tail_fork, st));
            CU(c, cudaStreamWaitEvent(c->s_tail, c->ev_tail_fork, 0));
            CU(c, launch_gauss_decide(fp, st, c->s_tail, &nl)); c->launches += nl;
            CU(c, cudaEventRecord(c->ev_tail_join, c->s_tail));
            CU(c, cudaStreamWaitEvent(st, c->ev_tail_join, 0));
        }
        { ProfScope p3(c, YSMR_PROF_K1C, st); CU(c, launch_pack_masks(fp, st)); c->launches += 1; }
    }
    const int64_t rows = (int64_t)n_frames * c->h;
    if (dbg && dbg->d_mask) { CU(c, launch_unpack_bits(c->mask_bits, dbg->d_mask, rows, c->w, c->ww, st)); c->launches++; }
    if (dbg && dbg->d_markers && fp.marker_bits) { CU(c, launch_unpack_bits(c->marker_bits, dbg->d_markers, rows, c->w, c->ww, st)); c->launches++; }
    // per-call counters
    CU(c, cudaMemsetAsync(c->ctl, 0, offsetof(Control, status), st));
    uint32_t *img = c->img_is_marker ? c->marker_bits : c->mask_bits;
    LabelLaunch L{};
    L.n_frames = n_frames; L.first_frame = first_frame; L.h = c->h; L.w = c->w; L.ww = c->ww;
    L.max_runs = c->p.max_runs; L.max_blobs = c->p.max_blobs; L.mode_propagate = c->propagate;
    L.img_bits = img; L.seed_bits = c->propagate ? c->marker_bits : nullptr;
    L.scratch = c->label_scratch; L.scratch_stride = c->label_stride;
    L.blob_count = d_blob_count; L.first_xy = c->first_xy; L.counts = c->counts;
    L.work = c->work; L.work_count = &c->ctl->work_count;
    L.status = &c->ctl->status; L.first_bad = &c->ctl->first_bad;
    {
        ProfScope ps(c, YSMR_PROF_LABEL, st);
        CU(c, launch_label(L, std::min(c->label_grid, n_frames), st)); c->launches++;
    }
    GeoLaunch G{};
    G.h = c->h; G.w = c->w; G.ww = c->ww; G.max_blobs = c->p.max_blobs; G.first_frame = first_frame;
    G.img_bits = img; G.first_xy = c->first_xy; G.work = c->work; G.work_count = &c->ctl->work_count;
    G.blobs = d_blobs; G.first_xy_dbg = dbg ? dbg->d_first_xy : nullptr;
    G.big_items = c->big_items; G.big_count = &c->ctl->big_count; G.big_cap = c->big_cap;
    G.pool = c->pool; G.pool_bytes = c->pool_bytes; G.pool_used = &c->ctl->pool_used;
    G.status = &c->ctl->status; G.first_bad = &c->ctl->first_bad;
    {
        ProfScope ps(c, YSMR_PROF_GEOMETRY, st);
        CU(c, launch_geometry(G, st)); c->launches += 2;
    }
    if (dbg && dbg->d_out) { CU(c, launch_unpack_bits(img, dbg->d_out, rows, c->w, c->ww, st)); c->launches++; }
    return YSMR_OK;
}

static int link_ready(ysmr_ctx *c)
{
    for (int i = 0; i < c->lc.n_f; ++i)
        if (!c->gain_set[i]) return fail(c, YSMR_E_STATE, "ysmr_set_gsff_gain has not been called for every filter");
    return YSMR_OK;
}

static int link_impl(ysmr_ctx *c, const int32_t *d_blob_count, const float *d_blobs, int first_frame, int n_frames,
                     ysmr_row *d_rows, int64_t rows_capacity, int64_t *d_n_rows, int append, cudaStream_t st,
                     const LinkGate *gate = nullptr)
{
    LinkIo io{};
    io.blob_count = d_blob_count; io.blobs = d_blobs; io.rows = (RowOut *)d_rows; io.rows_capacity = rows_capacity;
    io.n_rows = (long long *)d_n_rows; io.append = append;
    io.status = &c->ctl->status; io.first_bad = &c->ctl->first_bad;
    ProfScope ps(c, YSMR_PROF_LINK, st);
    CU(c, launch_link(c->lc, c->ls, c->lx, c->lf, io, first_frame, n_frames, c->link_fast, st, gate));
    c->launches += (c->link_fast) ? 3 * ((n_frames + c->lx.prep_frames - 1) / c->lx.prep_frames) : 1;
    return YSMR_OK;
}

int ysmr_link(ysmr_ctx *c, const int32_t *d_blob_count, const float *d_blobs, int first_frame, int n_frames,
              ysmr_row *d_rows, int64_t rows_capacity, int64_t *d_n_rows, void *stream)
{
    if (!c || !d_blob_count || !d_blobs || !d_rows || !d_n_rows) return fail(c, YSMR_E_INVALID, "null argument");
    if (n_frames < 0) return fail(c, YSMR_E_INVALID, "negative n_frames");
    int r = link_ready(c);
    if (r) return r;
    CU(c, cudaSetDevice(c->device));
    return link_impl(c, d_blob_count, d_blobs, first_frame, n_frames, d_rows, rows_capacity, d_n_rows, 0, (cudaStream_t)stream);
}

int ysmr_link_append(ysmr_ctx *c, const int32_t *d_blob_count, const float *d_blobs, int first_frame, int n_frames,
                     ysmr_row *d_rows, int64_t rows_capacity, int64_t *d_n_rows, void *stream)
{
    if (!c || !d_blob_count || !d_blobs || !d_rows || !d_n_rows) return fail(c, YSMR_E_INVALID, "null argument");
    if (n_frames < 0) return fail(c, YSMR_E_INVALID, "negative n_frames");
    int r = link_ready(c);
    if (r) return r;
    CU(c, cudaSetDevice(c->device));
    return link_impl(c, d_blob_count, d_blobs, first_frame, n_frames, d_rows, rows_capacity, d_n_rows, 1, (cudaStream_t)stream);
}

int ysmr_link_reset(ysmr_ctx *c)
{
    if (!c) return YSMR_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaDeviceSynchronize());
    CU(c, launch_link_reset(c->ls, c->p.max_tracks, nullptr)); c->launches++;
    CU(c, cudaMemset(c->values, 0, sizeof(double) * ((size_t)c->window + (size_t)c->p.max_batch)));
    CU(c, cudaDeviceSynchronize());
    return YSMR_OK;
}

int ysmr_link_state_export(ysmr_ctx *c, void *h_buf, size_t *size)
{
    if (!c || !size) return fail(c, YSMR_E_INVALID, "null argument");
    size_t total = 0;
    for (auto &pr : c->state_parts) total += pr.second;
    if (!h_buf) { *size = total; return YSMR_OK; }
    if (*size < total) { *size = total; return fail(c, YSMR_E_INVALID, "buffer too small"); }
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaDeviceSynchronize());
    uint8_t *q = (uint8_t *)h_buf;
    for (auto &pr : c->state_parts) { CU(c, cudaMemcpy(q, pr.first, pr.second, cudaMemcpyDeviceToHost)); q += pr.second; }
    *size = total;
    return YSMR_OK;
}

int ysmr_link_state_import(ysmr_ctx *c, const void *h_buf, size_t size)
{
    if (!c || !h_buf) return fail(c, YSMR_E_INVALID, "null argument");
    size_t total = 0;
    for (auto &pr : c->state_parts) total += pr.second;
    if (size != total) return fail(c, YSMR_E_INVALID, "state size mismatch (different capacities?)");
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaDeviceSynchronize());
    const uint8_t *q = (const uint8_t *)h_buf;
    for (auto &pr : c->state_parts) { CU(c, cudaMemcpy(pr.first, q, pr.second, cudaMemcpyHostToDevice)); q += pr.second; }
    return YSMR_OK;
}

int ysmr_link_live_tracks(ysmr_ctx *c, int32_t *n_live, int32_t *next_id)
{
    if (!c) return YSMR_E_INVALID;
    int32_t hdr[8];
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaDeviceSynchronize());
    CU(c, cudaMemcpy(hdr, c->ls.hdr, sizeof(hdr), cudaMemcpyDeviceToHost));
    if (n_live) *n_live = hdr[0];
    if (next_id) *next_id = hdr[1];
    return YSMR_OK;
}

int ysmr_status(ysmr_ctx *c, void *stream, int32_t *status_bits, int32_t *first_bad_frame)
{
    if (!c) return YSMR_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize((cudaStream_t)stream));
    CU(c, cudaDeviceSynchronize());
    Control h{};
    CU(c, cudaMemcpy(&h, c->ctl, sizeof(h), cudaMemcpyDeviceToHost));
    if (status_bits) *status_bits = h.status;
    if (first_bad_frame) *first_bad_frame = h.status ? h.first_bad : -1;
    if (h.status) {
        const int32_t reset[2] = {0, 0x7fffffff};
        CU(c, cudaMemcpy(&c->ctl->status, reset, sizeof(reset), cudaMemcpyHostToDevice));
        char buf[160];
        snprintf(buf, sizeof(buf), "device status 0x%x at frame %d (1 runs, 2 blobs, 4 contour points, 8 tracks, 16 rows, 32 linker gate time-out)",
                 h.status, h.first_bad);
        c->err = buf;
        return YSMR_E_OVERFLOW;
    }
    return YSMR_OK;
}

// Chunked pipeline over device-resident frames: detection of chunk i+1 (stream s_det) overlaps the sequential linker of
// chunk i (stream s_link).  `stream` is forked into both and joined again, so the call is asynchronous for the caller.
static int track_chunks(ysmr_ctx *c, const uint8_t *frames, bool host_frames, int n_frames, int64_t frame_stride,
                        int first_frame, ysmr_row *d_rows, int64_t rows_capacity, int64_t *d_n_rows, cudaStream_t user)
{
    const int B = c->p.max_batch;
    const size_t frame_bytes = (size_t)c->h * c->w * c->channels;
    if (host_frames && !c->stage[0]) {
        c->stage_bytes = frame_bytes * (size_t)B;
        for (int i = 0; i < 2; ++i) CU(c, cudaMalloc((void **)&c->stage[i], c->stage_bytes));
    }
    // Gated linker launches (LinkGate, link.cuh): the linker of chunk k is enqueued without waiting for the chunk's detection
    // and spins on gate_flags[k], which a memset behind the chunk's detection kernels sets.  Not while per-kernel timing is on
    // (the event bracket would time the spinning).
    const bool gated = c->link_gate && c->link_fast && !c->profiling && B <= c->lx.prep_frames;
    if (gated) CU(c, cudaMemsetAsync(c->gate_flags, 0, GATE_FLAGS * sizeof(int32_t), user));
    CU(c, cudaEventRecord(c->ev_fork, user));
    CU(c, cudaStreamWaitEvent(c->s_copy, c->ev_fork, 0));
    CU(c, cudaStreamWaitEvent(c->s_det, c->ev_fork, 0));
    CU(c, cudaStreamWaitEvent(c->s_link, c->ev_fork, 0));
    CU(c, cudaMemsetAsync(d_n_rows, 0, sizeof(int64_t), c->s_link));
    // Chunk schedule: a doubling lead-in, full batches while plenty of frames are left, then chunks that shrink geometrically
    // (each 45 % of what is left; measured best of three schedules).  The linker of chunk i runs beside the detection of chunk i+1, so whatever the linker still has to do when
    // the last detection finishes is exposed; with shrinking chunks it has caught up by then (a uniform schedule leaves the
    // link time of about one full batch as a tail).
    int chunk = 0;
    for (int f0 = 0, nf = 0; f0 < n_frames; f0 += nf, ++chunk) {
        const int rem = n_frames - f0;
        nf = std::min(B, std::max(96, (int)(0.45 * rem + 0.5)));
        // lead-in: the first chunks double from an eighth of a batch, so that the sequential linker -- the longer of the two
        // pipelines at ~50 tracks per frame -- starts after a fraction of a batch's detection time instead of a whole one
        if (chunk < 3) nf = std::min(nf, std::max(32, B / 8) << chunk);
        if (nf > rem || rem - nf < 48) nf = std::min(rem, B);
        const int b = chunk & 1;
        const uint8_t *src = frames + (size_t)f0 * frame_stride;
        const uint8_t *dfr = src;
        int64_t dstride = frame_stride;
        if (host_frames) {
            // staging buffer b is free once detection of chunk-2 has consumed it
            if (chunk >= 2) CU(c, cudaStreamWaitEvent(c->s_copy, c->ev_det[b], 0));
            if (frame_stride == (int64_t)frame_bytes)
                CU(c, cudaMemcpyAsync(c->stage[b], src, frame_bytes * (size_t)nf, cudaMemcpyHostToDevice, c->s_copy));
            else
                CU(c, cudaMemcpy2DAsync(c->stage[b], frame_bytes, src, (size_t)frame_stride, frame_bytes, (size_t)nf,
                                        cudaMemcpyHostToDevice, c->s_copy));
            CU(c, cudaEventRecord(c->ev_copy[b], c->s_copy));
            CU(c, cudaStreamWaitEvent(c->s_det, c->ev_copy[b], 0));
            dfr = c->stage[b]; dstride = (int64_t)frame_bytes;
        }
        // detection outputs b are free once the linker of chunk-2 has read them
        if (chunk >= 2) CU(c, cudaStreamWaitEvent(c->s_det, c->ev_link[b], 0));
        int r = ysmr_detect(c, dfr, nf, dstride, first_frame + f0, c->pipe_count[b], c->pipe_blobs[b], nullptr, c->s_det);
        if (r) return r;
        CU(c, cudaEventRecord(c->ev_det[b], c->s_det));
        if (gated && chunk < GATE_FLAGS) {
            // order of submission: tables, flag, and only then the kernel that waits for the flag (LinkGate, link.cuh)
            CU(c, launch_link_prep(c->lc, c->lx, c->pipe_count[b], c->pipe_blobs[b], nf, b, c->s_det)); c->launches++;
            CU(c, cudaMemsetAsync(c->gate_flags + chunk, 1, sizeof(int32_t), c->s_det));
            CU(c, cudaEventRecord(c->ev_det[b], c->s_det));
            const LinkGate gate{c->gate_flags + chunk, b};
            r = link_impl(c, c->pipe_count[b], c->pipe_blobs[b], first_frame + f0, nf, d_rows, rows_capacity, d_n_rows, 1, c->s_link, &gate);
            if (r) return r;
        } else {
            CU(c, cudaStreamWaitEvent(c->s_link, c->ev_det[b], 0));
            r = link_impl(c, c->pipe_count[b], c->pipe_blobs[b], first_frame + f0, nf, d_rows, rows_capacity, d_n_rows, 1, c->s_link);
            if (r) return r;
        }
        CU(c, cudaEventRecord(c->ev_link[b], c->s_link));
    }
    CU(c, cudaEventRecord(c->ev_join, c->s_link));
    CU(c, cudaStreamWaitEvent(user, c->ev_join, 0));
    CU(c, cudaEventRecord(c->ev_join, c->s_det));
    CU(c, cudaStreamWaitEvent(user, c->ev_join, 0));
    CU(c, cudaEventRecord(c->ev_join, c->s_copy));
    CU(c, cudaStreamWaitEvent(user, c->ev_join, 0));
    return YSMR_OK;
}

int ysmr_track_device(ysmr_ctx *c, const uint8_t *d_frames, int n_frames, int64_t frame_stride, int first_frame,
                      ysmr_row *d_rows, int64_t rows_capacity, int64_t *d_n_rows, void *stream)
{
    if (!c || !d_frames || !d_rows || !d_n_rows) return fail(c, YSMR_E_INVALID, "null argument");
    if (n_frames < 0) return fail(c, YSMR_E_INVALID, "negative n_frames");
    int r = link_ready(c);
    if (r) return r;
    CU(c, cudaSetDevice(c->device));
    return track_chunks(c, d_frames, false, n_frames, frame_stride, first_frame, d_rows, rows_capacity, d_n_rows,
                        (cudaStream_t)stream);
}

// ---- row sink on the device: archive + counting sort by (track_id, frame) --------------------------------------------
// Rows leave the linker frame-major, in insertion (= ascending id) order within a frame, and a track has a row in every
// frame from its registration to its deregistration (track_eval.py:313-316 emits every live object).  So the position of a
// row in the (TRACK_ID, POSITION_T) order of helper_file.sort_list is  offset[id] + (frame - first_frame[id])  with
// offset = exclusive scan of the rows per id: one pass of atomics, one scan, one scatter -- no comparison sort.
namespace {

__global__ void rows_count_kernel(const ysmr_row *rows, int64_t n, int32_t *count, int32_t *first, int n_ids, int32_t *bad)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int id = rows[i].track_id;
        if (id < 0 || id >= n_ids) { atomicOr(bad, 1); continue; }
        atomicAdd(&count[id], 1);
        atomicMin(&first[id], rows[i].frame);
    }
}

// exclusive scan of count[0 .. n) into offset (64-bit), one CTA
__global__ void __launch_bounds__(1024) rows_scan_kernel(const int32_t *count, long long *offset, int n)
{
    __shared__ long long warp_tot[32];
    __shared__ long long carry;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + t;
        const long long v = i < n ? count[i] : 0;
        long long inc = v;
        for (int o = 1; o < 32; o <<= 1) { const long long u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
        if (lane == 31) warp_tot[w] = inc;
        __syncthreads();
        if (w == 0) {
            long long x = warp_tot[lane], y = x;
            for (int o = 1; o < 32; o <<= 1) { const long long u = __shfl_up_sync(0xffffffffu, y, o); if (lane >= o) y += u; }
            warp_tot[lane] = y - x;
        }
        __syncthreads();
        const long long c0 = carry;
        if (i < n) offset[i] = c0 + warp_tot[w] + inc - v;
        __syncthreads();
        if (t == 1023) carry = c0 + warp_tot[w] + inc;
        __syncthreads();
    }
    if (t == 0) offset[n] = carry;
}

__global__ void rows_scatter_kernel(const ysmr_row *rows, int64_t n, const long long *offset, const int32_t *first, int n_ids,
                                    ysmr_row *out, int32_t *bad)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const ysmr_row r = rows[i];
        if (r.track_id < 0 || r.track_id >= n_ids) continue;
        const long long dst = offset[r.track_id] + (long long)(r.frame - first[r.track_id]);
        if (dst >= offset[r.track_id + 1]) { atomicOr(bad, 2); continue; }    // a gap in a track's frames: cannot happen
        out[dst] = r;
    }
}

}  // namespace

int ysmr_rows_archive(ysmr_ctx *c, int enabled)
{
    if (!c) return YSMR_E_INVALID;
    c->archive_on = enabled ? 1 : 0;
    c->arch_n = 0;
    if (!enabled && c->arch) { CU(c, cudaSetDevice(c->device)); CU(c, cudaDeviceSynchronize()); cudaFree(c->arch); c->arch = nullptr; c->arch_cap = 0; }
    return YSMR_OK;
}

static int archive_append(ysmr_ctx *c, const ysmr_row *d_rows, int64_t n, cudaStream_t st)
{
    if (n <= 0) return YSMR_OK;
    if (c->arch_n + n > c->arch_cap) {
        int64_t cap = std::max<int64_t>(c->arch_cap * 2, std::max<int64_t>(c->arch_n + n, 1 << 20));
        ysmr_row *bigger = nullptr;
        CU(c, cudaMalloc((void **)&bigger, sizeof(ysmr_row) * (size_t)cap));
        if (c->arch_n > 0) CU(c, cudaMemcpyAsync(bigger, c->arch, sizeof(ysmr_row) * (size_t)c->arch_n, cudaMemcpyDeviceToDevice, st));
        CU(c, cudaStreamSynchronize(st));
        if (c->arch) cudaFree(c->arch);
        c->arch = bigger; c->arch_cap = cap;
    }
    CU(c, cudaMemcpyAsync(c->arch + c->arch_n, d_rows, sizeof(ysmr_row) * (size_t)n, cudaMemcpyDeviceToDevice, st));
    c->arch_n += n;
    return YSMR_OK;
}

int ysmr_rows_sorted(ysmr_ctx *c, ysmr_row *h_rows, int64_t rows_capacity, int64_t *n_rows)
{
    if (!c || !n_rows) return fail(c, YSMR_E_INVALID, "null argument");
    if (!c->archive_on) return fail(c, YSMR_E_STATE, "ysmr_rows_archive has not been enabled");
    *n_rows = c->arch_n;
    if (!h_rows || c->arch_n == 0) return YSMR_OK;
    if (rows_capacity < c->arch_n) return fail(c, YSMR_E_INVALID, "buffer too small");
    CU(c, cudaSetDevice(c->device));
    cudaStream_t st = c->s_link;
    int32_t hdr[8];
    CU(c, cudaStreamSynchronize(st));
    CU(c, cudaMemcpy(hdr, c->ls.hdr, sizeof(hdr), cudaMemcpyDeviceToHost));
    const int n_ids = hdr[1] > 0 ? hdr[1] : 1;
    int32_t *count = nullptr, *first = nullptr, *bad = nullptr; long long *offset = nullptr; ysmr_row *sorted = nullptr;
    cudaError_t e = cudaMalloc((void **)&count, sizeof(int32_t) * (size_t)(2 * n_ids + 1));
    if (e == cudaSuccess) e = cudaMalloc((void **)&offset, sizeof(long long) * (size_t)(n_ids + 1));
    if (e == cudaSuccess) e = cudaMalloc((void **)&sorted, sizeof(ysmr_row) * (size_t)c->arch_n);
    int rc = YSMR_OK;
    if (e != cudaSuccess) rc = fail(c, YSMR_E_CUDA, cudaGetErrorString(e));
    else {
        first = count + n_ids; bad = first + n_ids;
        cudaMemsetAsync(count, 0, sizeof(int32_t) * (size_t)n_ids, st);
        cudaMemsetAsync(first, 0x7f, sizeof(int32_t) * (size_t)n_ids, st);
        cudaMemsetAsync(bad, 0, sizeof(int32_t), st);
        const int grid = (int)std::min<int64_t>((c->arch_n + 255) / 256, 148 * 8);
        rows_count_kernel<<<grid, 256, 0, st>>>(c->arch, c->arch_n, count, first, n_ids, bad);
        rows_scan_kernel<<<1, 1024, 0, st>>>(count, offset, n_ids);
        rows_scatter_kernel<<<grid, 256, 0, st>>>(c->arch, c->arch_n, offset, first, n_ids, sorted, bad);
        c->launches += 3;
        int32_t hbad = 0;
        e = cudaMemcpyAsync(&hbad, bad, sizeof(hbad), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_rows, sorted, sizeof(ysmr_row) * (size_t)c->arch_n, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = fail(c, YSMR_E_CUDA, cudaGetErrorString(e));
        else if (hbad) rc = fail(c, YSMR_E_STATE, "row archive is not frame-contiguous per track (rows from several videos?)");
    }
    if (count) cudaFree(count);
    if (offset) cudaFree(offset);
    if (sorted) cudaFree(sorted);
    return rc;
}

int ysmr_track_host(ysmr_ctx *c, const uint8_t *h_frames, int n_frames, int64_t frame_stride, int first_frame,
                    ysmr_row *h_rows, int64_t rows_capacity, int64_t *n_rows)
{
    if (!c || !h_frames || !n_rows || (!h_rows && !c->archive_on)) return fail(c, YSMR_E_INVALID, "null argument");
    if (n_frames < 0 || rows_capacity < 0) return fail(c, YSMR_E_INVALID, "negative size");
    int r = link_ready(c);
    if (r) return r;
    CU(c, cudaSetDevice(c->device));
    if (rows_capacity > c->rows_dev_cap) {
        if (c->rows_dev) CU(c, cudaFree(c->rows_dev));
        c->rows_dev = nullptr; c->rows_dev_cap = 0;
        CU(c, cudaMalloc((void **)&c->rows_dev, sizeof(ysmr_row) * (size_t)std::max<int64_t>(rows_capacity, 1)));
        c->rows_dev_cap = rows_capacity;
    }
    cudaStream_t user = c->s_link;   // any non-blocking stream of ours will do as the "user" stream here
    r = track_chunks(c, h_frames, true, n_frames, frame_stride, first_frame, c->rows_dev, rows_capacity,
                     (int64_t *)c->n_rows_dev, user);
    if (r) return r;
    long long n = 0;
    CU(c, cudaMemcpyAsync(&n, c->n_rows_dev, sizeof(n), cudaMemcpyDeviceToHost, user));
    CU(c, cudaStreamSynchronize(user));
    if (n > rows_capacity) n = rows_capacity;
    if (n > 0 && h_rows) CU(c, cudaMemcpy(h_rows, c->rows_dev, sizeof(ysmr_row) * (size_t)n, cudaMemcpyDeviceToHost));
    if (c->archive_on) { r = archive_append(c, c->rows_dev, n, user); if (r) return r; CU(c, cudaStreamSynchronize(user)); }
    *n_rows = n;
    int32_t bits = 0, bad = -1;
    return ysmr_status(c, user, &bits, &bad);
}

int ysmr_set_option(ysmr_ctx *c, int option, int value)
{
    if (!c) return YSMR_E_INVALID;
    if (option == YSMR_OPT_FRONTEND_GEN && (value == 3 || value == 4)) { c->frontend_gen = value; return YSMR_OK; }
    return fail(c, YSMR_E_INVALID, "unknown option or value");
}

int64_t ysmr_launch_count(const ysmr_ctx *c) { return c ? c->launches : 0; }

int ysmr_set_profiling(ysmr_ctx *c, int enabled)
{
    if (!c) return YSMR_E_INVALID;
    c->profiling = enabled ? 1 : 0;
    c->lx.phase_cycles = (enabled & 2) ? c->phase_cycles : nullptr;     // bit 1: per-phase clock64 counters in the linker
    return YSMR_OK;
}

int ysmr_link_phase_cycles(ysmr_ctx *c, int64_t *out16)
{
    if (!c || !out16) return YSMR_E_INVALID;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaDeviceSynchronize());
    CU(c, cudaMemcpy(out16, c->phase_cycles, 16 * sizeof(long long), cudaMemcpyDeviceToHost));
    CU(c, cudaMemset(c->phase_cycles, 0, 16 * sizeof(long long)));
    return YSMR_OK;
}

int ysmr_get_profile(ysmr_ctx *c, double *ms, int64_t *launches)
{
    if (!c || !ms || !launches) return fail(c, YSMR_E_INVALID, "null argument");
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaDeviceSynchronize());
    for (int k = 0; k < YSMR_PROF_KINDS; ++k) {
        double total = 0.0;
        for (auto &pr : c->prof_events[k]) {
            float t = 0.f;
            CU(c, cudaEventElapsedTime(&t, pr.first, pr.second));
            total += t;
            c->prof_free.push_back(pr.first); c->prof_free.push_back(pr.second);
        }
        ms[k] = total; launches[k] = (int64_t)c->prof_events[k].size();
        c->prof_events[k].clear();
    }
    return YSMR_OK;
}

}  // extern "C"
