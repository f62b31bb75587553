// vsb_api.cu -- C ABI of libvsb200.so (see include/vsb200.h for the contract and
// the reference interfaces each entry point stands behind).
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include <stdlib.h>

#include <new>

#include "vsb_common.cuh"
#include <vector>

int vsb_kepler_tables(vsb_ctx *c, vsb_kepler_state &k, const double *d_t);
int vsb_launch_vessel_tick(vsb_ctx *c, vsb_kepler_state &k, double warp, double *d_out);

// ---------------------------------------------------------------- errors
static thread_local char g_err[512] = "";

void vsb_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int vsb_buf_reserve(vsb_buf &b, size_t bytes)
{
    if (bytes <= b.cap) return VSB_OK;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        vsb_set_error("cudaMalloc(%zu) -> %s", want, cudaGetErrorString(e));
        b.p = nullptr;
        return VSB_ERR_NOMEM;
    }
    b.cap = want;
    return VSB_OK;
}

void vsb_buf_free(vsb_buf &b)
{
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

#define CTX(c)                                                                      \
    do {                                                                            \
        VSB_CHECK((c) != nullptr, VSB_ERR_ARG, "%s: null context", __func__);       \
        VSB_CUDA(cudaSetDevice((c)->device));                                       \
    } while (0)

static int h2d(vsb_ctx *c, void *dst, const void *src, size_t bytes)
{
    if (bytes == 0) return VSB_OK;
    VSB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
    VSB_CUDA(cudaStreamSynchronize(c->stream));
    return VSB_OK;
}

static int d2h(vsb_ctx *c, void *dst, const void *src, size_t bytes)
{
    if (bytes == 0) return VSB_OK;
    VSB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    VSB_CUDA(cudaStreamSynchronize(c->stream));
    return VSB_OK;
}

// ---------------------------------------------------------------- library / context
extern "C" int vsb_abi_version(void) { return VSB_ABI_VERSION; }

extern "C" const char *vsb_last_error(void) { return g_err; }

extern "C" int vsb_device_count(int *count_out)
{
    VSB_CHECK(count_out, VSB_ERR_ARG, "vsb_device_count: null output");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count_out = 0;
        vsb_set_error("cudaGetDeviceCount -> %s", cudaGetErrorString(e));
        return VSB_ERR_CUDA;
    }
    *count_out = n;
    return VSB_OK;
}

extern "C" int vsb_ctx_create(int device, vsb_ctx **ctx_out)
{
    VSB_CHECK(ctx_out, VSB_ERR_ARG, "vsb_ctx_create: null output");
    *ctx_out = nullptr;
    int ndev = 0;
    VSB_TRY(vsb_device_count(&ndev));
    VSB_CHECK(ndev > 0, VSB_ERR_CUDA, "vsb_ctx_create: no CUDA device (there is no CPU fallback)");
    VSB_CHECK(device >= 0 && device < ndev, VSB_ERR_ARG, "vsb_ctx_create: device %d of %d", device,
              ndev);
    VSB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    VSB_CUDA(cudaGetDeviceProperties(&prop, device));
    VSB_CHECK(prop.major >= 10, VSB_ERR_CUDA,
              "vsb_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only", device,
              prop.major, prop.minor);
    vsb_ctx *c = new (std::nothrow) vsb_ctx();
    VSB_CHECK(c, VSB_ERR_NOMEM, "vsb_ctx_create: out of host memory");
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->kep[0].kind = VSB_KIND_BODY;
    c->kep[1].kind = VSB_KIND_VESSEL;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
    if (e == cudaSuccess) e = cudaMalloc((void **)&c->d_key, 64);
    if (e == cudaSuccess) e = cudaMallocHost((void **)&c->h_key, 64);
    if (e != cudaSuccess) {
        vsb_set_error("vsb_ctx_create: %s", cudaGetErrorString(e));
        delete c;
        return VSB_ERR_CUDA;
    }
    *ctx_out = c;
    return VSB_OK;
}

static void free_bodies(vsb_ctx *c)
{
    if (c->mass) cudaFree(c->mass);  // one slab, mass is its base
    c->mass = nullptr;
    c->cap = 0;
}

static void free_kepler(vsb_kepler_state &k)
{
    if (k.a) cudaFree(k.a);  // one slab
    if (k.ref_pos && !k.ref_pos_borrowed) cudaFree(k.ref_pos);
    vsb_buf_free(k.curves);
    vsb_buf_free(k.quirk_idx);
    vsb_buf_free(k.ev);
    free(k.h_ref);
    int kind = k.kind;
    k = vsb_kepler_state();
    k.kind = kind;
}

extern "C" int vsb_ctx_destroy(vsb_ctx *c)
{
    if (!c) return VSB_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    free_bodies(c);
    free_kepler(c->kep[0]);
    free_kepler(c->kep[1]);
    vsb_buf_free(c->part);
    vsb_buf_free(c->sym_tab);
    vsb_buf_free(c->curve_tabs);
    vsb_buf_free(c->lean_tab);
    vsb_buf_free(c->order);
    vsb_buf_free(c->sorted);
    vsb_buf_free(c->pgrid);
    vsb_buf_free(c->scratch);
    vsb_buf_free(c->host_stage);
    vsb_buf_free(c->march_stats);
    if (c->d_key) cudaFree(c->d_key);
    if (c->h_key) cudaFreeHost(c->h_key);
    if (c->h_frame) cudaFreeHost(c->h_frame);
    free(c->h_curve_tabs);
    free(c->h_small);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return VSB_OK;
}

extern "C" int vsb_ctx_sync(vsb_ctx *c)
{
    CTX(c);
    VSB_CUDA(cudaStreamSynchronize(c->stream));
    return VSB_OK;
}

extern "C" int vsb_ctx_stream(vsb_ctx *c, void **out)
{
    CTX(c);
    VSB_CHECK(out, VSB_ERR_ARG, "vsb_ctx_stream: null output");
    *out = (void *)c->stream;
    return VSB_OK;
}

extern "C" int vsb_ctx_kernel_launches(vsb_ctx *c, int64_t *out)
{
    VSB_CHECK(c && out, VSB_ERR_ARG, "vsb_ctx_kernel_launches: null argument");
    *out = c->launches;
    return VSB_OK;
}

extern "C" int vsb_host_register(void *ptr, int64_t bytes)
{
    VSB_CHECK(ptr && bytes > 0, VSB_ERR_ARG, "vsb_host_register: bad argument");
    const cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault);
    if (e != cudaSuccess) {
        // pinning is optional: report it, but leave no pending error for the next launch check
        cudaGetLastError();
        vsb_set_error("vsb_host_register: %s", cudaGetErrorString(e));
        return VSB_ERR_CUDA;
    }
    return VSB_OK;
}

extern "C" int vsb_host_unregister(void *ptr)
{
    VSB_CHECK(ptr, VSB_ERR_ARG, "vsb_host_unregister: null pointer");
    const cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) {
        cudaGetLastError();
        vsb_set_error("vsb_host_unregister: %s", cudaGetErrorString(e));
        return VSB_ERR_CUDA;
    }
    return VSB_OK;
}

extern "C" int vsb_measure_fp64_peak(vsb_ctx *c, double *flops_out, double *pipe_ops_out)
{
    CTX(c);
    const int iters = 1 << 12, blocks = c->sm_count * 8;
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        float ms = 0;
        VSB_TRY(vsb_launch_dfma_peak(c, iters, blocks, &ms));
        if (ms < best) best = ms;
    }
    const double dfma = (double)blocks * 256.0 * 64.0 * (double)iters;
    if (pipe_ops_out) *pipe_ops_out = dfma / (best * 1e-3);
    if (flops_out) *flops_out = 2.0 * dfma / (best * 1e-3);
    return VSB_OK;
}

// ---------------------------------------------------------------- body-sim state
struct field_info {
    int width;  // 8-byte words per row
};
static const field_info k_fields[VSB_F_COUNT_] = {{1}, {1}, {1}, {1}, {2}, {2}, {2},
                                                  {2}, {1}, {1}, {2}, {1}, {1}, {1}};

static void *field_ptr(vsb_ctx *c, int field)
{
    switch (field) {
        case VSB_F_MASS: return c->mass;
        case VSB_F_DEN: return c->den;
        case VSB_F_RAD: return c->rad;
        case VSB_F_COI: return c->coi;
        case VSB_F_POS: return c->pos;
        case VSB_F_VEL: return c->vel;
        case VSB_F_REL_VEL: return c->rel_vel;
        case VSB_F_ACC: return c->acc;
        case VSB_F_PARENTS: return c->parents;
        case VSB_F_FOCUS: return c->focus;
        case VSB_F_ECC_V: return c->ecc_v;
        case VSB_F_SEMI_MAJOR: return c->semi_major;
        case VSB_F_SEMI_MINOR: return c->semi_minor;
        case VSB_F_PERIAPSIS_ARG: return c->periapsis_arg;
        default: return nullptr;
    }
}

extern "C" int vsb_bodies_resize(vsb_ctx *c, int64_t n)
{
    CTX(c);
    VSB_CHECK(n >= 0 && n < (1ll << 31), VSB_ERR_ARG, "vsb_bodies_resize: n=%lld out of range",
              (long long)n);
    const int64_t npad = vsb_round_up(n > 0 ? n : 1, VSB_PAD);
    if (npad > c->cap) {
        VSB_CUDA(cudaStreamSynchronize(c->stream));
        free_bodies(c);
        // one slab: 20 words per row (9 one-word fields incl. parents + 5 two-word) + parents_old
        const size_t words = 20;
        double *slab = nullptr;
        cudaError_t e = cudaMalloc((void **)&slab, sizeof(double) * words * (size_t)npad);
        VSB_CHECK(e == cudaSuccess, VSB_ERR_NOMEM, "vsb_bodies_resize: cudaMalloc for %lld bodies -> %s",
                  (long long)n, cudaGetErrorString(e));
        size_t o = 0;
        auto take = [&](int w) { double *p = slab + o; o += (size_t)w * npad; return p; };
        c->mass = take(1); c->den = take(1); c->rad = take(1); c->coi = take(1);
        c->focus = take(1); c->semi_major = take(1); c->semi_minor = take(1);
        c->periapsis_arg = take(1);
        c->parents = (int64_t *)take(1);
        c->parents_old = (int64_t *)take(1);
        c->pos = take(2); c->vel = take(2); c->rel_vel = take(2); c->acc = take(2);
        c->ecc_v = take(2);
        c->cap = npad;
        VSB_CUDA(cudaMemsetAsync(slab, 0, sizeof(double) * words * (size_t)npad, c->stream));
    }
    c->n = n;
    c->npad = npad;
    c->row0 = 0;
    c->row1 = n;
    VSB_TRY(vsb_launch_pad_tail(c));
    return VSB_OK;
}

extern "C" int vsb_bodies_count(vsb_ctx *c, int64_t *n_out)
{
    VSB_CHECK(c && n_out, VSB_ERR_ARG, "vsb_bodies_count: null argument");
    *n_out = c->n;
    return VSB_OK;
}

static int field_range(vsb_ctx *c, const char *who, int field, const void *p, int64_t first,
                       int64_t count)
{
    VSB_CHECK(field >= 0 && field < VSB_F_COUNT_, VSB_ERR_ARG, "%s: unknown field %d", who, field);
    VSB_CHECK(c->mass != nullptr, VSB_ERR_STATE, "%s: no bodies (call vsb_bodies_resize)", who);
    VSB_CHECK(first >= 0 && count >= 0 && first + count <= c->n, VSB_ERR_ARG,
              "%s: rows [%lld,%lld) outside [0,%lld)", who, (long long)first,
              (long long)(first + count), (long long)c->n);
    VSB_CHECK(p != nullptr || count == 0, VSB_ERR_ARG, "%s: null buffer", who);
    return VSB_OK;
}

extern "C" int vsb_bodies_set(vsb_ctx *c, int field, const void *src, int64_t first, int64_t count)
{
    CTX(c);
    VSB_TRY(field_range(c, "vsb_bodies_set", field, src, first, count));
    const size_t w = 8 * (size_t)k_fields[field].width;
    return h2d(c, (char *)field_ptr(c, field) + w * first, src, w * count);
}

extern "C" int vsb_bodies_get(vsb_ctx *c, int field, void *dst, int64_t first, int64_t count)
{
    CTX(c);
    VSB_TRY(field_range(c, "vsb_bodies_get", field, dst, first, count));
    const size_t w = 8 * (size_t)k_fields[field].width;
    return d2h(c, dst, (char *)field_ptr(c, field) + w * first, w * count);
}

extern "C" int vsb_bodies_device_ptr(vsb_ctx *c, int field, void **out)
{
    CTX(c);
    VSB_CHECK(out && field >= 0 && field < VSB_F_COUNT_, VSB_ERR_ARG,
              "vsb_bodies_device_ptr: bad argument");
    VSB_CHECK(c->mass != nullptr, VSB_ERR_STATE, "vsb_bodies_device_ptr: no bodies");
    *out = field_ptr(c, field);
    return VSB_OK;
}

extern "C" int vsb_bodies_set_rows(vsb_ctx *c, int64_t row0, int64_t row1)
{
    VSB_CHECK(c, VSB_ERR_ARG, "vsb_bodies_set_rows: null context");
    VSB_CHECK(row0 >= 0 && row0 <= row1 && row1 <= c->n, VSB_ERR_ARG,
              "vsb_bodies_set_rows: [%lld,%lld) outside [0,%lld]", (long long)row0,
              (long long)row1, (long long)c->n);
    c->row0 = row0;
    c->row1 = row1;
    return VSB_OK;
}

static int need_bodies(vsb_ctx *c, const char *who)
{
    VSB_CHECK(c->mass != nullptr && c->n > 0, VSB_ERR_STATE, "%s: no bodies uploaded", who);
    return VSB_OK;
}

// order == NULL reuses the order uploaded by the previous call (valid while n is unchanged)
static int upload_order(vsb_ctx *c, const char *who, const int64_t *order)
{
    if (!order) {
        VSB_CHECK(c->order_n == c->n && c->order.p, VSB_ERR_STATE,
                  "%s: no order resident for %lld bodies", who, (long long)c->n);
        return VSB_OK;
    }
    VSB_TRY(vsb_buf_reserve(c->order, sizeof(int64_t) * (size_t)c->npad));
    c->order_n = -1;
    VSB_TRY(h2d(c, c->order.p, order, sizeof(int64_t) * (size_t)c->n));
    c->order_n = c->n;
    return VSB_OK;
}

// Mapped page-locked result area of the small-system kernels: [frame record: 4 + 14 * 1024 words]
// [sequence word][4 header words][3 spare].
static const size_t k_flag_at = 4 + 14 * (size_t)1024;

// true when results may go through the mapped area: not switched off (VSB_NO_MAPPED_FRAME) and the
// mapped page-locked allocation exists -- a refused allocation (locked-memory limit) is remembered
// and every caller falls back to its copy path: the mapped words are an optimisation, never a
// requirement
static bool use_mapped(vsb_ctx *c)
{
    if (getenv("VSB_NO_MAPPED_FRAME") || c->mapped_refused) return false;
    if (c->h_frame) return true;
    const size_t cap = 8 * (k_flag_at + 8);
    if (cudaHostAlloc((void **)&c->h_frame, cap, cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void **)&c->d_frame, c->h_frame, 0) != cudaSuccess) {
        cudaGetLastError();            // clear the (non-sticky) allocation error
        if (c->h_frame) cudaFreeHost(c->h_frame);
        c->h_frame = nullptr;
        c->mapped_refused = true;
        return false;
    }
    memset(c->h_frame, 0, cap);
    return true;
}

// spins until the kernel launched last has published `seq` (a faulting kernel never does: the
// stream is polled now and then)
static int mapped_wait(vsb_ctx *c, long long seq, const char *who)
{
    volatile long long *flag = (volatile long long *)(c->h_frame + k_flag_at);
    for (unsigned spins = 1; *flag != seq; ++spins) {
        if ((spins & 0x3fff) == 0) {
            cudaError_t q = cudaStreamQuery(c->stream);
            if (q == cudaSuccess && *flag != seq) q = cudaErrorUnknown;
            if (q != cudaErrorNotReady)
                VSB_CHECK(*flag == seq, VSB_ERR_CUDA, "%s: kernel -> %s", who, cudaGetErrorString(q));
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    return VSB_OK;
}

// A few small device results -> host with ONE wait: nf segments of `words[f]` 8-byte words each.
// Up to 14 340 words in all they are packed by one CTA into mapped page-locked memory, followed by a
// sequence word the host spins on (no copy call, no synchronise); above, or with
// VSB_NO_MAPPED_FRAME, asynchronous copies and one synchronise.
static int d2h_small(vsb_ctx *c, int nf, const void *const *src, const int *words, void *const *dst,
                     const char *who)
{
    size_t total_words = 0;
    for (int f = 0; f < nf; ++f) total_words += (size_t)words[f];
    if (total_words == 0) return VSB_OK;
    if (nf <= 16 && total_words <= k_flag_at && use_mapped(c)) {
        int64_t total = 0;
        const long long seq = ++c->frame_seq;
        VSB_TRY(vsb_launch_pack_fields(c, nullptr, src, words, nf, c->d_frame, &total,
                                       (long long *)(c->d_frame + k_flag_at), seq, 1));
        VSB_TRY(mapped_wait(c, seq, who));
        const double *h = c->h_frame;
        for (int f = 0; f < nf; ++f) {
            memcpy(dst[f], h, 8 * (size_t)words[f]);
            h += words[f];
        }
        return VSB_OK;
    }
    for (int f = 0; f < nf; ++f)
        if (words[f])
            VSB_CUDA(cudaMemcpyAsync(dst[f], src[f], 8 * (size_t)words[f], cudaMemcpyDeviceToHost, c->stream));
    VSB_CUDA(cudaStreamSynchronize(c->stream));
    return VSB_OK;
}

static_assert(sizeof(long long) == sizeof(int64_t), "mapped header words");

// The fused small-system frame / tick (n <= 1024, one CTA) and its results on the host: h[0..3] =
// ticks done, del, add, chain-overflow flag; frame_out (optional) receives the frame record.
// The kernel writes the record and the header straight into mapped page-locked host memory and then
// a sequence word the host spins on -- no copy call and no stream synchronise on the 60 Hz path (a
// launch-latency bound frame at 15 bodies: 25 k -> 36 k frames per second).  VSB_NO_MAPPED_FRAME=1
// goes through cudaMemcpyAsync + cudaStreamSynchronize instead (same results).
static int run_frame_small(vsb_ctx *c, double gc, double coi_coef, const int64_t *d_order, int max_ticks,
                           int do_kepler_basic, int do_collision, void *frame_out, int64_t *h)
{
    const size_t rec_words = 4 + 14 * (size_t)c->n;
    if (use_mapped(c)) {
        const long long seq = ++c->frame_seq;
        VSB_TRY(vsb_launch_editor_frame_small(c, gc, coi_coef, d_order, max_ticks, do_kepler_basic,
                                              do_collision, frame_out ? c->d_frame : nullptr,
                                              (long long *)(c->d_frame + k_flag_at), seq));
        VSB_TRY(mapped_wait(c, seq, "fused frame"));
        if (frame_out) memcpy(frame_out, c->h_frame, 8 * rec_words);
        memcpy(h, c->h_frame + k_flag_at + 1, 4 * sizeof(int64_t));
        return VSB_OK;
    }
    if (frame_out) {
        VSB_TRY(vsb_buf_reserve(c->host_stage, 8 * rec_words));
        double *d_packed = (double *)c->host_stage.p;
        VSB_TRY(vsb_launch_editor_frame_small(c, gc, coi_coef, d_order, max_ticks, do_kepler_basic,
                                              do_collision, d_packed));
        VSB_TRY(d2h(c, frame_out, d_packed, 8 * rec_words));
        memcpy(h, frame_out, 4 * sizeof(int64_t));
        return VSB_OK;
    }
    VSB_TRY(vsb_launch_editor_frame_small(c, gc, coi_coef, d_order, max_ticks, do_kepler_basic,
                                          do_collision, nullptr));
    VSB_CUDA(cudaMemcpyAsync(h, c->scratch.p, 4 * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    VSB_CUDA(cudaStreamSynchronize(c->stream));
    return VSB_OK;
}

extern "C" int vsb_find_parents(vsb_ctx *c, const int64_t *order, int64_t *parents_out)
{
    CTX(c);
    VSB_TRY(need_bodies(c, "vsb_find_parents"));
    VSB_TRY(upload_order(c, "vsb_find_parents", order));
    VSB_TRY(vsb_launch_find_parents(c, (const int64_t *)c->order.p));
    if (parents_out) return d2h(c, parents_out, c->parents, sizeof(int64_t) * (size_t)c->n);
    VSB_CUDA(cudaStreamSynchronize(c->stream));
    return VSB_OK;
}

extern "C" int vsb_gravity_ref(vsb_ctx *c, double gc, const int64_t *order)
{
    CTX(c);
    VSB_TRY(need_bodies(c, "vsb_gravity_ref"));
    VSB_TRY(upload_order(c, "vsb_gravity_ref", order));
    if (c->n <= 1024 && !getenv("VSB_NO_FUSED_FRAME")) {
        // small systems: the whole tick in one single-CTA launch (bit-identical results)
        int64_t *h = (int64_t *)c->h_key;
        VSB_TRY(run_frame_small(c, gc, 0.0, (const int64_t *)c->order.p, 1, 0, 0, nullptr, h));
        VSB_CHECK(h[3] == 0, VSB_ERR_STATE, "vsb_gravity_ref: parent chain deeper than 48 levels");
        return VSB_OK;
    }
    return vsb_launch_gravity_ref(c, gc, (const int64_t *)c->order.p);
}

// Multi-GPU form of the MODE_REF tick (SURVEY 8e "find_parents"): rank `part` of `parts` scans its
// row blocks of the mass-sorted bodies; the int32 array at vsb_find_parents_best_ptr (n entries,
// -1 = no enclosing COI) is then reduced with MAX over the ranks and every rank finishes the tick.
extern "C" int vsb_gravity_ref_begin_async(vsb_ctx *c, const int64_t *order, int part, int parts)
{
    CTX(c);
    VSB_TRY(need_bodies(c, "vsb_gravity_ref_begin_async"));
    VSB_CHECK(parts >= 1 && part >= 0 && part < parts, VSB_ERR_ARG,
              "vsb_gravity_ref_begin_async: part %d of %d", part, parts);
    VSB_TRY(upload_order(c, "vsb_gravity_ref_begin_async", order));
    return vsb_launch_gravity_ref_begin(c, (const int64_t *)c->order.p, part, parts);
}

extern "C" int vsb_find_parents_best_ptr(vsb_ctx *c, void **dptr_out)
{
    CTX(c);
    VSB_CHECK(dptr_out != nullptr, VSB_ERR_ARG, "vsb_find_parents_best_ptr: null output");
    VSB_TRY(need_bodies(c, "vsb_find_parents_best_ptr"));
    VSB_TRY(vsb_buf_reserve(c->sorted, (size_t)c->npad * (16 + 8 + 8 + 4)));
    *dptr_out = (char *)c->sorted.p + (size_t)c->npad * 32;
    return VSB_OK;
}

extern "C" int vsb_gravity_ref_finish(vsb_ctx *c, double gc)
{
    CTX(c);
    VSB_TRY(need_bodies(c, "vsb_gravity_ref_finish"));
    VSB_TRY(upload_order(c, "vsb_gravity_ref_finish", nullptr));
    return vsb_launch_gravity_ref_finish(c, gc, (const int64_t *)c->order.p);
}

// Which all-pairs kernel runs: the symmetric one only when this context owns every row
// (the row kernel is the one that shards); VSB_GRAVITY=rows|sym overrides the context mode.
static bool use_sym(vsb_ctx *c)
{
    if (c->row0 != 0 || c->row1 != c->n) return false;
    const char *env = getenv("VSB_GRAVITY");
    if (env && env[0] == 'r') return false;
    if (env && env[0] == 's') return true;
    if (c->gravity_mode >= 0) return c->gravity_mode == 1;
    return c->n >= 8192;   // auto: below that the row kernel's finer grid wins
}

// one acceleration pass (+ integrator) with whichever kernel applies; asynchronous
static int accel_pass(vsb_ctx *c, double gc, int integrate)
{
    if (use_sym(c)) return vsb_launch_allpairs_sym(c, gc, 0, 1, integrate);
    VSB_TRY(vsb_launch_allpairs_accel(c, gc));
    return vsb_launch_allpairs_reduce(c, gc, integrate);
}

extern "C" int vsb_set_gravity_mode(vsb_ctx *c, int mode)
{
    VSB_CHECK(c && mode >= -1 && mode <= 1, VSB_ERR_ARG, "vsb_set_gravity_mode: bad argument");
    c->gravity_mode = mode;
    return VSB_OK;
}

extern "C" int vsb_allpairs_sym_partial_async(vsb_ctx *c, double gc, int rank, int world)
{
    CTX(c);
    VSB_TRY(need_bodies(c, "vsb_allpairs_sym_partial_async"));
    VSB_CHECK(world >= 1 && rank >= 0 && rank < world, VSB_ERR_ARG,
              "vsb_allpairs_sym_partial_async: rank %d of %d", rank, world);
    return vsb_launch_allpairs_sym(c, gc, rank, world, 0);
}

extern "C" int vsb_allpairs_integrate_async(vsb_ctx *c)
{
    CTX(c);
    VSB_TRY(need_bodies(c, "vsb_allpairs_integrate_async"));
    return vsb_launch_allpairs_integrate(c);
}

extern "C" int vsb_allpairs_accel(vsb_ctx *c, double gc)
{
    CTX(c);
    VSB_TRY(need_bodies(c, "vsb_allpairs_accel"));
    VSB_TRY(accel_pass(c, gc, 0));
    VSB_CUDA(cudaStreamSynchronize(c->stream));
    return VSB_OK;
}

extern "C" int vsb_allpairs_integrate(vsb_ctx *c)
{
    CTX(c);
    VSB_TRY(need_bodies(c, "vsb_allpairs_integrate"));
    VSB_TRY(vsb_launch_allpairs_integrate(c));
    VSB_CUDA(cudaStreamSynchronize(c->stream));
    return VSB_OK;
}

// asynchronous variants used by the multi-GPU driver (no host sync inside)
extern "C" int vsb_allpairs_tick_async(vsb_ctx *c, double gc)
{
    CTX(c);
    VSB_TRY(need_bodies(c, "vsb_allpairs_tick_async"));
    return accel_pass(c, gc, 1);
}

extern "C" int vsb_allpairs_accel_async(vsb_ctx *c, double gc)
{
    CTX(c);
    VSB_TRY(need_bodies(c, "vsb_allpairs_accel_async"));
    return accel_pass(c, gc, 0);
}

extern "C" int vsb_allpairs_step(vsb_ctx *c, double gc, int64_t nsteps)
{
    CTX(c);
    VSB_TRY(need_bodies(c, "vsb_allpairs_step"));
    VSB_CHECK(c->row0 == 0 && c->row1 == c->n, VSB_ERR_STATE,
              "vsb_allpairs_step: rows are sharded [%lld,%lld); step per tick and all-gather "
              "positions between ticks instead",
              (long long)c->row0, (long long)c->row1);
    for (int64_t s = 0; s < nsteps; ++s) VSB_TRY(accel_pass(c, gc, 1));
    VSB_CUDA(cudaStreamSynchronize(c->stream));
    return VSB_OK;
}

static const int k_frame_fields[10] = {VSB_F_POS, VSB_F_VEL, VSB_F_REL_VEL, VSB_F_PARENTS,
                                       VSB_F_FOCUS, VSB_F_ECC_V, VSB_F_SEMI_MAJOR,
                                       VSB_F_SEMI_MINOR, VSB_F_PERIAPSIS_ARG, VSB_F_COI};

static int pack_frame(vsb_ctx *c, double *d_dst)
{
    const void *ptrs[10];
    int widths[10];
    for (int f = 0; f < 10; ++f) {
        ptrs[f] = field_ptr(c, k_frame_fields[f]);
        widths[f] = k_fields[k_frame_fields[f]].width;
    }
    int64_t total = 0;
    return vsb_launch_pack_fields(c, k_frame_fields, ptrs, widths, 10, d_dst, &total);
}

extern "C" int vsb_editor_ticks(vsb_ctx *c, double gc, double coi_coef, const int64_t *order,
                                int64_t max_ticks, int do_kepler_basic, int64_t *ticks_done,
                                int64_t *del_out, int64_t *add_out, void *frame_out)
{
    CTX(c);
    VSB_CHECK(ticks_done && del_out && add_out && max_ticks >= 0, VSB_ERR_ARG,
              "vsb_editor_ticks: bad argument");
    VSB_TRY(need_bodies(c, "vsb_editor_ticks"));
    VSB_TRY(upload_order(c, "vsb_editor_ticks", order));
    const int64_t *d_order = (const int64_t *)c->order.p;
    *ticks_done = 0;
    *del_out = -1;
    *add_out = -1;
    const size_t rec_words = 4 + 14 * (size_t)c->n;
    int64_t *h = (int64_t *)c->h_key;
    if (c->n <= 1024 && !getenv("VSB_NO_FUSED_FRAME")) {
        VSB_CHECK(max_ticks < (1ll << 30), VSB_ERR_ARG, "vsb_editor_ticks: max_ticks too large");
        VSB_TRY(run_frame_small(c, gc, coi_coef, d_order, (int)max_ticks, do_kepler_basic, 1, frame_out, h));
        VSB_CHECK(h[3] == 0, VSB_ERR_STATE, "vsb_editor_ticks: parent chain deeper than 48 levels");
        *ticks_done = h[0];
        *del_out = h[1];
        *add_out = h[2];
        return VSB_OK;
    }
    int64_t d = -1, a = -1;
    for (int64_t t = 0; t < max_ticks; ++t) {
        VSB_TRY(vsb_launch_gravity_ref(c, gc, d_order));
        *ticks_done = t + 1;
        VSB_TRY(vsb_check_collision(c, &d, &a));
        if (d >= 0) break;
    }
    *del_out = d;
    *add_out = a;
    if (do_kepler_basic && d < 0) VSB_TRY(vsb_launch_kepler_basic(c, gc, coi_coef));
    if (frame_out) {
        VSB_TRY(vsb_buf_reserve(c->host_stage, 8 * rec_words));
        double *d_packed = (double *)c->host_stage.p;
        VSB_TRY(pack_frame(c, d_packed + 4));
        VSB_TRY(d2h(c, (char *)frame_out + 32, d_packed + 4, 8 * (rec_words - 4)));
        int64_t *fh = (int64_t *)frame_out;
        fh[0] = *ticks_done;
        fh[1] = d;
        fh[2] = a;
        fh[3] = 0;
    }
    VSB_CUDA(cudaStreamSynchronize(c->stream));
    return VSB_OK;
}

extern "C" int vsb_bodies_get_many(vsb_ctx *c, const int *fields, int nfields, void *dst)
{
    CTX(c);
    VSB_CHECK(fields && dst && nfields > 0 && nfields <= 16, VSB_ERR_ARG,
              "vsb_bodies_get_many: bad argument");
    VSB_TRY(need_bodies(c, "vsb_bodies_get_many"));
    const void *ptrs[16];
    int widths[16];
    size_t words = 0;
    for (int f = 0; f < nfields; ++f) {
        VSB_CHECK(fields[f] >= 0 && fields[f] < VSB_F_COUNT_, VSB_ERR_ARG,
                  "vsb_bodies_get_many: unknown field %d", fields[f]);
        ptrs[f] = field_ptr(c, fields[f]);
        widths[f] = k_fields[fields[f]].width;
        words += (size_t)c->n * widths[f];
    }
    int64_t total = 0;
    if (c->n <= 1024 && words <= k_flag_at && use_mapped(c)) {
        // small systems: one CTA packs the fields straight into mapped host memory and publishes a
        // sequence word (no copy call, no stream synchronise: 30 -> 12 us per pull)
        const long long seq = ++c->frame_seq;
        VSB_TRY(vsb_launch_pack_fields(c, fields, ptrs, widths, nfields, c->d_frame, &total,
                                       (long long *)(c->d_frame + k_flag_at), seq));
        VSB_TRY(mapped_wait(c, seq, "vsb_bodies_get_many"));
        memcpy(dst, c->h_frame, 8 * (size_t)total);
        return VSB_OK;
    }
    VSB_TRY(vsb_buf_reserve(c->host_stage, 8 * words));
    VSB_TRY(vsb_launch_pack_fields(c, fields, ptrs, widths, nfields, (double *)c->host_stage.p,
                                   &total));
    return d2h(c, dst, c->host_stage.p, 8 * (size_t)total);
}

extern "C" int vsb_kepler_basic(vsb_ctx *c, double gc, double coi_coef)
{
    CTX(c);
    VSB_TRY(need_bodies(c, "vsb_kepler_basic"));
    VSB_TRY(vsb_launch_kepler_basic(c, gc, coi_coef));
    VSB_CUDA(cudaStreamSynchronize(c->stream));
    return VSB_OK;
}

extern "C" int vsb_simplified_orbit_coi(vsb_ctx *c, double gc, double coi_coef,
                                        const int64_t *order, int64_t *largest_out)
{
    CTX(c);
    VSB_TRY(need_bodies(c, "vsb_simplified_orbit_coi"));
    VSB_TRY(upload_order(c, "vsb_simplified_orbit_coi", order));
    VSB_TRY(vsb_launch_simplified_orbit_coi(c, gc, coi_coef, (const int64_t *)c->order.p));
    int64_t largest = 0;
    VSB_TRY(d2h(c, &largest, c->scratch.p, sizeof(int64_t)));
    if (largest_out) *largest_out = largest;
    return VSB_OK;
}

// VSB_COLLISION=brute|grid forces one path (tests); default: grid from 4096 bodies up
static int collision_use_grid(int64_t n)
{
    const char *env = getenv("VSB_COLLISION");
    if (env && env[0] == 'b') return 0;
    if (env && env[0] == 'g') return n >= 2;
    return n >= 4096;
}

extern "C" int vsb_check_collision(vsb_ctx *c, int64_t *del_out, int64_t *add_out)
{
    CTX(c);
    VSB_CHECK(del_out && add_out, VSB_ERR_ARG, "vsb_check_collision: null output");
    VSB_TRY(need_bodies(c, "vsb_check_collision"));
    int64_t *h = (int64_t *)c->h_key;
    if (c->n >= 1 && c->n <= 1024 && !getenv("VSB_COLLISION") && use_mapped(c)) {
        // small systems: one launch of one CTA, the answer through mapped host memory (the three
        // launches + copy + synchronise of the general path take 56 us at 15 bodies, this 12)
        const long long seq = ++c->frame_seq;
        VSB_TRY(vsb_launch_check_collision_small(c, (long long *)(c->d_frame + k_flag_at), seq));
        VSB_TRY(mapped_wait(c, seq, "vsb_check_collision"));
        memcpy(h, c->h_frame + k_flag_at + 1, 2 * sizeof(int64_t));
        *del_out = h[0];
        *add_out = h[1];
        return VSB_OK;
    }
    if (collision_use_grid(c->n)) {
        VSB_TRY(vsb_launch_broadphase(c));
        VSB_CUDA(cudaMemcpyAsync(h, c->scratch.p, 3 * sizeof(int64_t), cudaMemcpyDeviceToHost,
                                 c->stream));
        VSB_CUDA(cudaStreamSynchronize(c->stream));
        if (h[2] == 0) {
            *del_out = h[0];
            *add_out = h[1];
            return VSB_OK;
        }
        // grid unsuitable for this input: exhaustive scan below
    }
    VSB_TRY(vsb_launch_check_collision(c));
    VSB_CUDA(cudaMemcpyAsync(h, c->scratch.p, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost,
                             c->stream));
    VSB_CUDA(cudaStreamSynchronize(c->stream));
    *del_out = h[0];
    *add_out = h[1];
    return VSB_OK;
}

// Multi-GPU form of the exhaustive scan: rank `part` of `parts` scans its row blocks
// and leaves its local minimum (i << 32 | j, all ones = none) in the device word returned by
// vsb_collision_key_ptr; the caller takes the minimum over ranks (one 64-bit all-reduce) and
// calls vsb_check_collision_finish on every rank.
extern "C" int vsb_check_collision_part_async(vsb_ctx *c, int part, int parts)
{
    CTX(c);
    VSB_TRY(need_bodies(c, "vsb_check_collision_part_async"));
    VSB_CHECK(parts >= 1 && part >= 0 && part < parts, VSB_ERR_ARG,
              "vsb_check_collision_part_async: part %d of %d", part, parts);
    return vsb_launch_check_collision_part(c, part, parts);
}

extern "C" int vsb_collision_key_ptr(vsb_ctx *c, void **dptr_out)
{
    CTX(c);
    VSB_CHECK(dptr_out != nullptr, VSB_ERR_ARG, "vsb_collision_key_ptr: null output");
    *dptr_out = c->d_key;
    return VSB_OK;
}

extern "C" int vsb_check_collision_finish(vsb_ctx *c, int64_t *del_out, int64_t *add_out)
{
    CTX(c);
    VSB_CHECK(del_out && add_out, VSB_ERR_ARG, "vsb_check_collision_finish: null output");
    VSB_TRY(need_bodies(c, "vsb_check_collision_finish"));
    VSB_TRY(vsb_launch_check_collision_finish(c));
    int64_t *h = (int64_t *)c->h_key;
    VSB_CUDA(cudaMemcpyAsync(h, c->scratch.p, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost,
                             c->stream));
    VSB_CUDA(cudaStreamSynchronize(c->stream));
    *del_out = h[0];
    *add_out = h[1];
    return VSB_OK;
}

extern "C" int vsb_delete_row(vsb_ctx *c, int64_t row)
{
    CTX(c);
    VSB_TRY(need_bodies(c, "vsb_delete_row"));
    VSB_CHECK(row >= 0 && row < c->n, VSB_ERR_ARG, "vsb_delete_row: row %lld of %lld",
              (long long)row, (long long)c->n);
    VSB_CHECK(c->n > 1, VSB_ERR_STATE, "vsb_delete_row: the last body cannot be deleted");
    VSB_TRY(vsb_launch_delete_row(c, row));
    VSB_CUDA(cudaStreamSynchronize(c->stream));
    return VSB_OK;
}

extern "C" int vsb_swap_rows(vsb_ctx *c, int64_t a, int64_t b)
{
    CTX(c);
    VSB_TRY(need_bodies(c, "vsb_swap_rows"));
    VSB_CHECK(a >= 0 && a < c->n && b >= 0 && b < c->n, VSB_ERR_ARG,
              "vsb_swap_rows: rows %lld,%lld of %lld", (long long)a, (long long)b, (long long)c->n);
    if (a != b) VSB_TRY(vsb_launch_swap_rows(c, a, b));
    VSB_CUDA(cudaStreamSynchronize(c->stream));
    return VSB_OK;
}

extern "C" int vsb_translate_to(vsb_ctx *c, int64_t row)
{
    CTX(c);
    VSB_TRY(need_bodies(c, "vsb_translate_to"));
    VSB_CHECK(row >= 0 && row < c->n, VSB_ERR_ARG, "vsb_translate_to: row %lld of %lld",
              (long long)row, (long long)c->n);
    VSB_TRY(vsb_launch_translate(c, row));
    VSB_CUDA(cudaStreamSynchronize(c->stream));
    return VSB_OK;
}

extern "C" int vsb_editor_curve(vsb_ctx *c, int64_t P, const double *ell_t, const double *par_t,
                                double *out)
{
    CTX(c);
    VSB_TRY(need_bodies(c, "vsb_editor_curve"));
    VSB_CHECK(P > 0 && ell_t && par_t && out, VSB_ERR_ARG, "vsb_editor_curve: bad argument");
    const size_t out_bytes = sizeof(double) * 2 * (size_t)c->n * (size_t)P;
    const size_t tab_bytes = sizeof(double) * 2 * (size_t)P;
    VSB_TRY(vsb_buf_reserve(c->host_stage, out_bytes));
    double *d_out = (double *)c->host_stage.p;
    // the parameter tables change with the settings only (editor.py calls curve() every frame):
    // they stay on the device and are re-sent when their contents differ
    if (c->curve_tabs_P != P || !c->h_curve_tabs || memcmp(c->h_curve_tabs, ell_t, tab_bytes / 2) != 0 ||
        memcmp(c->h_curve_tabs + P, par_t, tab_bytes / 2) != 0) {
        c->curve_tabs_P = 0;
        VSB_TRY(vsb_buf_reserve(c->curve_tabs, tab_bytes));
        double *h = (double *)realloc(c->h_curve_tabs, tab_bytes);
        VSB_CHECK(h != nullptr, VSB_ERR_NOMEM, "vsb_editor_curve: out of host memory");
        c->h_curve_tabs = h;
        memcpy(h, ell_t, tab_bytes / 2);
        memcpy(h + P, par_t, tab_bytes / 2);
        VSB_TRY(h2d(c, c->curve_tabs.p, h, tab_bytes));
        c->curve_tabs_P = P;
    }
    VSB_TRY(vsb_launch_editor_curve(c, P, (const double *)c->curve_tabs.p, d_out));
    return d2h(c, out, d_out, out_bytes);
}

// ---------------------------------------------------------------- stateless host-buffer calls
// They stage through the context's own body arrays: resident body state is overwritten.
static int stage_bodies(vsb_ctx *c, int64_t n)
{
    if (c->n != n || c->mass == nullptr) VSB_TRY(vsb_bodies_resize(c, n));
    c->row0 = 0;
    c->row1 = n;
    return VSB_OK;
}

extern "C" int vsb_host_allpairs_step(vsb_ctx *c, int64_t n, const double *mass, double *pos,
                                      double *vel, double gc, int64_t nsteps)
{
    CTX(c);
    VSB_CHECK(n > 0 && mass && pos && vel && nsteps >= 0, VSB_ERR_ARG,
              "vsb_host_allpairs_step: bad argument");
    VSB_TRY(stage_bodies(c, n));
    VSB_CUDA(cudaMemcpyAsync(c->mass, mass, 8 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    VSB_CUDA(cudaMemcpyAsync(c->pos, pos, 16 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    VSB_CUDA(cudaMemcpyAsync(c->vel, vel, 16 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    for (int64_t s = 0; s < nsteps; ++s) VSB_TRY(accel_pass(c, gc, 1));
    VSB_CUDA(cudaMemcpyAsync(pos, c->pos, 16 * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    VSB_CUDA(cudaMemcpyAsync(vel, c->vel, 16 * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    VSB_CUDA(cudaStreamSynchronize(c->stream));
    return VSB_OK;
}

extern "C" int vsb_host_allpairs_accel(vsb_ctx *c, int64_t n, const double *mass,
                                       const double *pos, double gc, double *acc_out)
{
    CTX(c);
    VSB_CHECK(n > 0 && mass && pos && acc_out, VSB_ERR_ARG, "vsb_host_allpairs_accel: bad argument");
    VSB_TRY(stage_bodies(c, n));
    VSB_CUDA(cudaMemcpyAsync(c->mass, mass, 8 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    VSB_CUDA(cudaMemcpyAsync(c->pos, pos, 16 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    VSB_TRY(accel_pass(c, gc, 0));
    return d2h(c, acc_out, c->acc, 16 * (size_t)n);
}

extern "C" int vsb_host_find_parents(vsb_ctx *c, int64_t n, const int64_t *order,
                                     const double *pos, const double *coi, int64_t *parents_out)
{
    CTX(c);
    VSB_CHECK(n > 0 && order && pos && coi && parents_out, VSB_ERR_ARG,
              "vsb_host_find_parents: bad argument");
    VSB_TRY(stage_bodies(c, n));
    VSB_CUDA(cudaMemcpyAsync(c->pos, pos, 16 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    VSB_CUDA(cudaMemcpyAsync(c->coi, coi, 8 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    return vsb_find_parents(c, order, parents_out);
}

extern "C" int vsb_host_check_collision(vsb_ctx *c, int64_t n, const double *mass,
                                        const double *pos, const double *rad, int64_t *del_out,
                                        int64_t *add_out)
{
    CTX(c);
    VSB_CHECK(n > 0 && mass && pos && rad, VSB_ERR_ARG, "vsb_host_check_collision: bad argument");
    VSB_TRY(stage_bodies(c, n));
    VSB_CUDA(cudaMemcpyAsync(c->mass, mass, 8 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    VSB_CUDA(cudaMemcpyAsync(c->pos, pos, 16 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    VSB_CUDA(cudaMemcpyAsync(c->rad, rad, 8 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    return vsb_check_collision(c, del_out, add_out);
}

extern "C" int vsb_host_solve_kepler_ell(vsb_ctx *c, int64_t n, const double *ecc,
                                         const double *ma, double tol, double *out)
{
    CTX(c);
    VSB_CHECK(n >= 0 && (n == 0 || (ecc && ma && out)), VSB_ERR_ARG,
              "vsb_host_solve_kepler_ell: bad argument");
    if (n == 0) return VSB_OK;
    VSB_TRY(vsb_buf_reserve(c->host_stage, 24 * (size_t)n));
    double *d = (double *)c->host_stage.p;
    VSB_CUDA(cudaMemcpyAsync(d, ecc, 8 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    VSB_CUDA(cudaMemcpyAsync(d + n, ma, 8 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    VSB_TRY(vsb_launch_solve_ell(c, n, d, d + n, tol, d + 2 * n));
    return d2h(c, out, d + 2 * n, 8 * (size_t)n);
}

extern "C" int vsb_host_newton_root_kepler_hyp(vsb_ctx *c, int64_t n, const double *ecc,
                                               const double *ma, const double *guess, double *out)
{
    CTX(c);
    VSB_CHECK(n >= 0 && (n == 0 || (ecc && ma && guess && out)), VSB_ERR_ARG,
              "vsb_host_newton_root_kepler_hyp: bad argument");
    if (n == 0) return VSB_OK;
    VSB_TRY(vsb_buf_reserve(c->host_stage, 32 * (size_t)n));
    double *d = (double *)c->host_stage.p;
    VSB_CUDA(cudaMemcpyAsync(d, ecc, 8 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    VSB_CUDA(cudaMemcpyAsync(d + n, ma, 8 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    VSB_CUDA(cudaMemcpyAsync(d + 2 * n, guess, 8 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    VSB_TRY(vsb_launch_solve_hyp(c, n, d, d + n, d + 2 * n, d + 3 * n));
    return d2h(c, out, d + 3 * n, 8 * (size_t)n);
}

extern "C" int vsb_host_curve_points(vsb_ctx *c, int64_t n, int64_t P, const double *ecc,
                                     const double *a, const double *b, const double *pea,
                                     const double *t, double *out)
{
    CTX(c);
    VSB_CHECK(n >= 0 && P > 0 && (n == 0 || (ecc && a && b && pea && t && out)), VSB_ERR_ARG,
              "vsb_host_curve_points: bad argument");
    if (n == 0) return VSB_OK;
    const size_t out_bytes = 16 * (size_t)n * (size_t)P;
    const size_t na = ((size_t)n + 31) / 32 * 32;   // keep every staged array 256-byte aligned
    VSB_TRY(vsb_buf_reserve(c->host_stage, out_bytes + 8 * (4 * na + 5 * (size_t)P)));
    double *d_out = (double *)c->host_stage.p;
    double *d = (double *)((char *)c->host_stage.p + out_bytes);
    vsb_kepler_state k;
    k.kind = VSB_KIND_VESSEL;
    k.n = n;
    k.P = P;
    k.ecc = d; k.a = d + na; k.b = d + 2 * na; k.pea = d + 3 * na;
    k.tab = d + 4 * na;         // 4P doubles
    double *d_t = d + 4 * na + 4 * P;
    VSB_CUDA(cudaMemcpyAsync(k.ecc, ecc, 8 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    VSB_CUDA(cudaMemcpyAsync(k.a, a, 8 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    VSB_CUDA(cudaMemcpyAsync(k.b, b, 8 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    VSB_CUDA(cudaMemcpyAsync(k.pea, pea, 8 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    VSB_CUDA(cudaMemcpyAsync(d_t, t, 8 * (size_t)P, cudaMemcpyHostToDevice, c->stream));
    VSB_TRY(vsb_kepler_tables(c, k, d_t));
    VSB_TRY(vsb_launch_kepler_curves(c, k, 0, n, d_out, 0));
    return d2h(c, out, d_out, out_bytes);
}

extern "C" int vsb_host_curve_move_to(vsb_ctx *c, int64_t n, int64_t P, int64_t nref,
                                      const double *curves, const double *body_pos,
                                      const int64_t *ref, const double *f, const double *pea,
                                      double *out)
{
    CTX(c);
    VSB_CHECK(n >= 0 && P > 0 && nref > 0 && (n == 0 || (curves && body_pos && ref && f && pea && out)),
              VSB_ERR_ARG, "vsb_host_curve_move_to: bad argument");
    if (n == 0) return VSB_OK;
    const size_t cb = 16 * (size_t)n * (size_t)P;
    VSB_TRY(vsb_buf_reserve(c->host_stage, 2 * cb + 16 * (size_t)nref + 24 * (size_t)n));
    char *base = (char *)c->host_stage.p;
    double *d_in = (double *)base, *d_out = (double *)(base + cb);
    double *d_bp = (double *)(base + 2 * cb);
    int64_t *d_ref = (int64_t *)(base + 2 * cb + 16 * (size_t)nref);
    double *d_f = (double *)(d_ref + n), *d_pea = d_f + n;
    VSB_CUDA(cudaMemcpyAsync(d_in, curves, cb, cudaMemcpyHostToDevice, c->stream));
    VSB_CUDA(cudaMemcpyAsync(d_bp, body_pos, 16 * (size_t)nref, cudaMemcpyHostToDevice, c->stream));
    VSB_CUDA(cudaMemcpyAsync(d_ref, ref, 8 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    VSB_CUDA(cudaMemcpyAsync(d_f, f, 8 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    VSB_CUDA(cudaMemcpyAsync(d_pea, pea, 8 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    VSB_TRY(vsb_launch_curve_move_to(c, n, P, d_in, d_bp, d_ref, d_f, d_pea, d_out));
    return d2h(c, out, d_out, cb);
}

// ---------------------------------------------------------------- game-mode Kepler state
static int kep_of(vsb_ctx *c, const char *who, int kind, vsb_kepler_state **k)
{
    VSB_CHECK(kind == VSB_KIND_BODY || kind == VSB_KIND_VESSEL, VSB_ERR_ARG, "%s: unknown kind %d",
              who, kind);
    *k = &c->kep[kind];
    return VSB_OK;
}

static void *kep_field_ptr(vsb_kepler_state &k, int field, int *width, int *is_int)
{
    *width = 1;
    *is_int = 0;
    switch (field) {
        case VSB_K_A: return k.a;
        case VSB_K_B: return k.b;
        case VSB_K_F: return k.f;
        case VSB_K_ECC: return k.ecc;
        case VSB_K_PEA: return k.pea;
        case VSB_K_N: return k.nmot;
        case VSB_K_DR: return k.dr;
        case VSB_K_MA: return k.ma;
        case VSB_K_EA: return k.ea;
        case VSB_K_POS: *width = 2; return k.pos;
        case VSB_K_REF: *is_int = 1; return k.ref;
        default: return nullptr;
    }
}

extern "C" int vsb_kepler_load(vsb_ctx *c, int kind, int64_t n, int64_t P, const double *a,
                               const double *b, const double *f, const double *ecc,
                               const double *pea, const double *nmot, const double *dr,
                               const double *ma, const double *ea, const int64_t *ref,
                               const double *t)
{
    CTX(c);
    vsb_kepler_state *kp;
    VSB_TRY(kep_of(c, "vsb_kepler_load", kind, &kp));
    VSB_CHECK(n > 0 && P > 0 && a && b && f && ecc && pea && nmot && dr && ma && ea && ref && t,
              VSB_ERR_ARG, "vsb_kepler_load: bad argument");
    VSB_CUDA(cudaStreamSynchronize(c->stream));
    free_kepler(*kp);
    if (kind == VSB_KIND_BODY && c->kep[VSB_KIND_VESSEL].ref_pos_borrowed) {
        // the vessel state was borrowing the old BODY positions: drop the alias
        c->kep[VSB_KIND_VESSEL].ref_pos = nullptr;
        c->kep[VSB_KIND_VESSEL].ref_pos_borrowed = 0;
        c->kep[VSB_KIND_VESSEL].nref = 0;
    }
    vsb_kepler_state &k = *kp;
    // slab (8-byte words): 9 element arrays + pos(2) + pr(2) + ref + lvl_idx = 15n, tab 4P, t P;
    // every array starts on a 256-byte boundary (double2 / 128-bit accesses)
    auto al = [](size_t w) { return (w + 31) / 32 * 32; };
    const size_t words = 11 * al((size_t)n) + 2 * al(2 * (size_t)n) + al(4 * (size_t)P) +
                         al((size_t)P);
    double *slab = nullptr;
    cudaError_t e = cudaMalloc((void **)&slab, 8 * words);
    VSB_CHECK(e == cudaSuccess, VSB_ERR_NOMEM, "vsb_kepler_load: cudaMalloc -> %s",
              cudaGetErrorString(e));
    size_t o = 0;
    auto take = [&](size_t w) { double *p = slab + o; o += al(w); return p; };
    k.a = take(n); k.b = take(n); k.f = take(n); k.ecc = take(n); k.pea = take(n);
    k.nmot = take(n); k.dr = take(n); k.ma = take(n); k.ea = take(n);
    k.pos = take(2 * n); k.pr = take(2 * n);
    k.ref = (int64_t *)take(n); k.lvl_idx = (int64_t *)take(n);
    k.tab = take(4 * P);
    double *d_t = take(P);
    k.n = n; k.P = P; k.kind = kind; k.nlevels = 0;
    k.h_ref = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    VSB_CHECK(k.h_ref, VSB_ERR_NOMEM, "vsb_kepler_load: out of host memory");
    memcpy(k.h_ref, ref, sizeof(int64_t) * (size_t)n);
    const double *src[9] = {a, b, f, ecc, pea, nmot, dr, ma, ea};
    double *dst[9] = {k.a, k.b, k.f, k.ecc, k.pea, k.nmot, k.dr, k.ma, k.ea};
    for (int i = 0; i < 9; ++i)
        VSB_CUDA(cudaMemcpyAsync(dst[i], src[i], 8 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    VSB_CUDA(cudaMemcpyAsync(k.ref, ref, 8 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    VSB_CUDA(cudaMemsetAsync(k.pos, 0, 16 * (size_t)n, c->stream));
    VSB_CUDA(cudaMemcpyAsync(d_t, t, 8 * (size_t)P, cudaMemcpyHostToDevice, c->stream));
    VSB_TRY(vsb_kepler_tables(c, k, d_t));
    VSB_CUDA(cudaStreamSynchronize(c->stream));
    return VSB_OK;
}

extern "C" int vsb_kepler_set_order(vsb_ctx *c, const int64_t *order)
{
    CTX(c);
    vsb_kepler_state &k = c->kep[VSB_KIND_BODY];
    VSB_CHECK(k.n > 0, VSB_ERR_STATE, "vsb_kepler_set_order: no BODY state loaded");
    VSB_CHECK(order, VSB_ERR_ARG, "vsb_kepler_set_order: null order");
    const int64_t n = k.n;
    int64_t *rank = (int64_t *)malloc(sizeof(int64_t) * 3 * (size_t)n);
    VSB_CHECK(rank, VSB_ERR_NOMEM, "vsb_kepler_set_order: out of host memory");
    int64_t *level = rank + n, *idx = rank + 2 * n;
    for (int64_t r = 0; r < n; ++r) {
        if (order[r] < 0 || order[r] >= n) {
            free(rank);
            VSB_CHECK(false, VSB_ERR_ARG, "vsb_kepler_set_order: order[%lld] out of range",
                      (long long)r);
        }
        rank[order[r]] = r;
    }
    int64_t nlevels = 0;
    int64_t counts[65] = {0};
    for (int64_t r = 0; r < n; ++r) {  // ascending rank: a "new" ref is already levelled
        const int64_t body = order[r], rf = k.h_ref[body];
        const bool dep_new = rf >= 0 && rf < n && rank[rf] < r;
        level[body] = dep_new ? level[rf] + 1 : 0;
        if (level[body] >= 64) {
            free(rank);
            VSB_CHECK(false, VSB_ERR_ARG, "vsb_kepler_set_order: hierarchy deeper than 64 levels");
        }
        counts[level[body]]++;
        if (level[body] + 1 > nlevels) nlevels = level[body] + 1;
    }
    k.lvl_off[0] = 0;
    for (int64_t l = 0; l < nlevels; ++l) k.lvl_off[l + 1] = k.lvl_off[l] + counts[l];
    int64_t fill[65];
    for (int64_t l = 0; l <= nlevels; ++l) fill[l] = k.lvl_off[l];
    for (int64_t r = 0; r < n; ++r) {
        const int64_t body = order[r], rf = k.h_ref[body];
        const bool dep_new = rf >= 0 && rf < n && rank[rf] < r;
        idx[fill[level[body]]++] = body | ((int64_t)(dep_new ? 1 : 0) << 62);
    }
    k.nlevels = nlevels;
    int rc = h2d(c, k.lvl_idx, idx, sizeof(int64_t) * (size_t)n);
    free(rank);
    return rc;
}

static int kep_range(vsb_kepler_state &k, const char *who, int field, const void *p, int64_t first,
                     int64_t count)
{
    VSB_CHECK(k.n > 0, VSB_ERR_STATE, "%s: no state loaded for this kind", who);
    VSB_CHECK(field >= 0 && field < VSB_K_COUNT_, VSB_ERR_ARG, "%s: unknown field %d", who, field);
    VSB_CHECK(first >= 0 && count >= 0 && first + count <= k.n, VSB_ERR_ARG,
              "%s: objects [%lld,%lld) outside [0,%lld)", who, (long long)first,
              (long long)(first + count), (long long)k.n);
    VSB_CHECK(p != nullptr || count == 0, VSB_ERR_ARG, "%s: null buffer", who);
    return VSB_OK;
}

extern "C" int vsb_kepler_set(vsb_ctx *c, int kind, int field, const void *src, int64_t first,
                              int64_t count)
{
    CTX(c);
    vsb_kepler_state *k;
    VSB_TRY(kep_of(c, "vsb_kepler_set", kind, &k));
    VSB_TRY(kep_range(*k, "vsb_kepler_set", field, src, first, count));
    int w, is_int;
    char *p = (char *)kep_field_ptr(*k, field, &w, &is_int);
    if (field == VSB_K_REF) memcpy(k->h_ref + first, src, 8 * (size_t)count);
    return h2d(c, p + 8 * (size_t)w * first, src, 8 * (size_t)w * count);
}

extern "C" int vsb_kepler_get(vsb_ctx *c, int kind, int field, void *dst, int64_t first,
                              int64_t count)
{
    CTX(c);
    vsb_kepler_state *k;
    VSB_TRY(kep_of(c, "vsb_kepler_get", kind, &k));
    VSB_TRY(kep_range(*k, "vsb_kepler_get", field, dst, first, count));
    int w, is_int;
    char *p = (char *)kep_field_ptr(*k, field, &w, &is_int);
    return d2h(c, dst, p + 8 * (size_t)w * first, 8 * (size_t)w * count);
}

extern "C" int vsb_kepler_device_ptr(vsb_ctx *c, int kind, int field, void **out)
{
    CTX(c);
    vsb_kepler_state *k;
    VSB_TRY(kep_of(c, "vsb_kepler_device_ptr", kind, &k));
    VSB_CHECK(out && k->n > 0 && field >= 0 && field < VSB_K_COUNT_, VSB_ERR_ARG,
              "vsb_kepler_device_ptr: bad argument or no state");
    int w, is_int;
    *out = kep_field_ptr(*k, field, &w, &is_int);
    return VSB_OK;
}

static int vessel_ref_pos(vsb_ctx *c, vsb_kepler_state &k, const double *body_pos, int64_t nref)
{
    if (body_pos) {
        VSB_CHECK(nref > 0, VSB_ERR_ARG, "kepler(VESSEL): nref must be positive");
        if (k.ref_pos_borrowed) {
            k.ref_pos = nullptr;
            k.ref_pos_borrowed = 0;
            k.nref = 0;
        }
        if (nref > k.nref || !k.ref_pos) {
            VSB_CUDA(cudaStreamSynchronize(c->stream));
            if (k.ref_pos) cudaFree(k.ref_pos);
            k.ref_pos = nullptr;
            cudaError_t e = cudaMalloc((void **)&k.ref_pos, 16 * (size_t)nref);
            VSB_CHECK(e == cudaSuccess, VSB_ERR_NOMEM, "kepler(VESSEL): cudaMalloc -> %s",
                      cudaGetErrorString(e));
        }
        k.nref = nref;
        VSB_CUDA(cudaMemcpyAsync(k.ref_pos, body_pos, 16 * (size_t)nref, cudaMemcpyHostToDevice,
                                 c->stream));
    } else {
        // NULL: borrow the resident BODY state's device positions (no copy)
        vsb_kepler_state &b = c->kep[VSB_KIND_BODY];
        VSB_CHECK(b.n > 0, VSB_ERR_STATE,
                  "kepler(VESSEL): body_pos is NULL and no BODY state is resident");
        if (k.ref_pos && !k.ref_pos_borrowed) {
            VSB_CUDA(cudaStreamSynchronize(c->stream));
            cudaFree(k.ref_pos);
        }
        k.ref_pos = b.pos;
        k.ref_pos_borrowed = 1;
        k.nref = b.n;
    }
    return VSB_OK;
}

// phys_vessel.move keeps the anomalies of the previous tick (self.prev_ea, phys_vessel.py:486);
// the resident event state needs them for cross_coi and predict_enter_coi_service
static int keep_prev_ea(vsb_ctx *c, vsb_kepler_state &k)
{
    if (k.kind == VSB_KIND_VESSEL && k.ev_n == k.n && k.ev_prev_ea)
        VSB_CUDA(cudaMemcpyAsync(k.ev_prev_ea, k.ea, 8 * (size_t)k.n, cudaMemcpyDeviceToDevice,
                                 c->stream));
    return VSB_OK;
}

extern "C" int vsb_kepler_move(vsb_ctx *c, int kind, double warp, const double *body_pos,
                               int64_t nref)
{
    CTX(c);
    vsb_kepler_state *k;
    VSB_TRY(kep_of(c, "vsb_kepler_move", kind, &k));
    VSB_CHECK(k->n > 0, VSB_ERR_STATE, "vsb_kepler_move: no state loaded for this kind");
    if (kind == VSB_KIND_VESSEL) VSB_TRY(vessel_ref_pos(c, *k, body_pos, nref));
    VSB_TRY(keep_prev_ea(c, *k));
    VSB_TRY(vsb_launch_kepler_move(c, *k, warp));
    VSB_CUDA(cudaStreamSynchronize(c->stream));
    return VSB_OK;
}

// C1 / C4 with the results handed back in the same call: move + (pos, ma, ea) of all objects on the
// host with ONE wait (phys_body.Physics.move -> (pos, ma, ea), phys_body.py:346; phys_vessel's
// :494).  Small states (up to 3 584 objects) come back through mapped page-locked memory: one CTA
// packs the three fields there and publishes a sequence word -- no copy call, no synchronise.
extern "C" int vsb_kepler_move_state(vsb_ctx *c, int kind, double warp, const double *body_pos,
                                     int64_t nref, double *pos_out, double *ma_out, double *ea_out)
{
    CTX(c);
    vsb_kepler_state *k;
    VSB_TRY(kep_of(c, "vsb_kepler_move_state", kind, &k));
    VSB_CHECK(k->n > 0, VSB_ERR_STATE, "vsb_kepler_move_state: no state loaded for this kind");
    VSB_CHECK(pos_out && ma_out && ea_out, VSB_ERR_ARG, "vsb_kepler_move_state: null output");
    if (kind == VSB_KIND_VESSEL) VSB_TRY(vessel_ref_pos(c, *k, body_pos, nref));
    VSB_TRY(keep_prev_ea(c, *k));
    VSB_TRY(vsb_launch_kepler_move(c, *k, warp));
    const size_t n = (size_t)k->n;
    if (4 * n <= k_flag_at && use_mapped(c)) {
        const void *ptrs[3] = {k->pos, k->ma, k->ea};
        const int widths[3] = {2, 1, 1};
        int64_t total = 0;
        const long long seq = ++c->frame_seq;
        VSB_TRY(vsb_launch_pack_fields(c, nullptr, ptrs, widths, 3, c->d_frame, &total,
                                       (long long *)(c->d_frame + k_flag_at), seq, k->n));
        VSB_TRY(mapped_wait(c, seq, "vsb_kepler_move_state"));
        memcpy(pos_out, c->h_frame, 16 * n);
        memcpy(ma_out, c->h_frame + 2 * n, 8 * n);
        memcpy(ea_out, c->h_frame + 3 * n, 8 * n);
        return VSB_OK;
    }
    VSB_CUDA(cudaMemcpyAsync(pos_out, k->pos, 16 * n, cudaMemcpyDeviceToHost, c->stream));
    VSB_CUDA(cudaMemcpyAsync(ma_out, k->ma, 8 * n, cudaMemcpyDeviceToHost, c->stream));
    VSB_CUDA(cudaMemcpyAsync(ea_out, k->ea, 8 * n, cudaMemcpyDeviceToHost, c->stream));
    VSB_CUDA(cudaStreamSynchronize(c->stream));
    return VSB_OK;
}

static int reserve_curves(vsb_kepler_state &k)
{
    return vsb_buf_reserve(k.curves, 16 * (size_t)k.n * (size_t)k.P);
}

extern "C" int vsb_kepler_curves(vsb_ctx *c, int kind)
{
    CTX(c);
    vsb_kepler_state *k;
    VSB_TRY(kep_of(c, "vsb_kepler_curves", kind, &k));
    VSB_CHECK(k->n > 0, VSB_ERR_STATE, "vsb_kepler_curves: no state loaded for this kind");
    if (kind == VSB_KIND_VESSEL)
        VSB_CHECK(k->ref_pos != nullptr, VSB_ERR_STATE,
                  "vsb_kepler_curves(VESSEL): call vsb_kepler_move first (parent positions)");
    VSB_TRY(reserve_curves(*k));
    VSB_TRY(vsb_launch_kepler_curves(c, *k, 0, k->n, (double *)k->curves.p, 1));
    VSB_CUDA(cudaStreamSynchronize(c->stream));
    return VSB_OK;
}

extern "C" int vsb_kepler_tick(vsb_ctx *c, int kind, double warp, const double *body_pos,
                               int64_t nref)
{
    CTX(c);
    vsb_kepler_state *k;
    VSB_TRY(kep_of(c, "vsb_kepler_tick", kind, &k));
    VSB_CHECK(k->n > 0, VSB_ERR_STATE, "vsb_kepler_tick: no state loaded for this kind");
    VSB_TRY(reserve_curves(*k));
    if (kind == VSB_KIND_VESSEL) {
        VSB_TRY(vessel_ref_pos(c, *k, body_pos, nref));
        VSB_TRY(keep_prev_ea(c, *k));
        VSB_TRY(vsb_launch_vessel_tick(c, *k, warp, (double *)k->curves.p));
    } else {
        VSB_TRY(vsb_launch_kepler_move(c, *k, warp));
        VSB_TRY(vsb_launch_kepler_curves(c, *k, 0, k->n, (double *)k->curves.p, 1));
    }
    VSB_CUDA(cudaStreamSynchronize(c->stream));
    return VSB_OK;
}

// asynchronous tick for device-timed benchmarking (no host sync)
extern "C" int vsb_kepler_tick_async(vsb_ctx *c, int kind, double warp)
{
    CTX(c);
    vsb_kepler_state *k;
    VSB_TRY(kep_of(c, "vsb_kepler_tick_async", kind, &k));
    VSB_CHECK(k->n > 0, VSB_ERR_STATE, "vsb_kepler_tick_async: no state loaded for this kind");
    VSB_TRY(reserve_curves(*k));
    if (kind == VSB_KIND_VESSEL) {
        VSB_CHECK(k->ref_pos != nullptr, VSB_ERR_STATE,
                  "vsb_kepler_tick_async(VESSEL): parent positions not resident");
        VSB_TRY(keep_prev_ea(c, *k));
        return vsb_launch_vessel_tick(c, *k, warp, (double *)k->curves.p);
    }
    VSB_TRY(vsb_launch_kepler_move(c, *k, warp));
    return vsb_launch_kepler_curves(c, *k, 0, k->n, (double *)k->curves.p, 1);
}

extern "C" int vsb_kepler_curves_get(vsb_ctx *c, int kind, int64_t first, int64_t count,
                                     double *out)
{
    CTX(c);
    vsb_kepler_state *k;
    VSB_TRY(kep_of(c, "vsb_kepler_curves_get", kind, &k));
    VSB_CHECK(k->n > 0 && k->curves.p, VSB_ERR_STATE,
              "vsb_kepler_curves_get: no curves generated yet");
    VSB_CHECK(first >= 0 && count >= 0 && first + count <= k->n && (out || count == 0), VSB_ERR_ARG,
              "vsb_kepler_curves_get: objects [%lld,%lld) outside [0,%lld)", (long long)first,
              (long long)(first + count), (long long)k->n);
    const size_t stride = 16 * (size_t)k->P;
    return d2h(c, out, (char *)k->curves.p + stride * first, stride * count);
}

// C5 + C6 with the result handed back in the same call: vsb_kepler_curves + vsb_kepler_curves_get
// of all objects, one wait; up to 114 KB of curves come back through mapped page-locked memory.
extern "C" int vsb_kepler_curve_move(vsb_ctx *c, int kind, double *out)
{
    CTX(c);
    vsb_kepler_state *k;
    VSB_TRY(kep_of(c, "vsb_kepler_curve_move", kind, &k));
    VSB_CHECK(k->n > 0 && out, VSB_ERR_STATE, "vsb_kepler_curve_move: no state loaded / null output");
    if (kind == VSB_KIND_VESSEL)
        VSB_CHECK(k->ref_pos != nullptr, VSB_ERR_STATE,
                  "vsb_kepler_curve_move(VESSEL): call vsb_kepler_move first (parent positions)");
    VSB_TRY(reserve_curves(*k));
    VSB_TRY(vsb_launch_kepler_curves(c, *k, 0, k->n, (double *)k->curves.p, 1));
    const size_t words = 2 * (size_t)k->n * (size_t)k->P;
    if (words <= k_flag_at && use_mapped(c)) {
        const void *ptrs[1] = {k->curves.p};
        const int widths[1] = {1};
        int64_t total = 0;
        const long long seq = ++c->frame_seq;
        VSB_TRY(vsb_launch_pack_fields(c, nullptr, ptrs, widths, 1, c->d_frame, &total,
                                       (long long *)(c->d_frame + k_flag_at), seq, (int64_t)words));
        VSB_TRY(mapped_wait(c, seq, "vsb_kepler_curve_move"));
        memcpy(out, c->h_frame, 8 * words);
        return VSB_OK;
    }
    return d2h(c, out, k->curves.p, 8 * words);
}

extern "C" int vsb_kepler_curves_device_ptr(vsb_ctx *c, int kind, void **out)
{
    CTX(c);
    vsb_kepler_state *k;
    VSB_TRY(kep_of(c, "vsb_kepler_curves_device_ptr", kind, &k));
    VSB_CHECK(out && k->curves.p, VSB_ERR_STATE, "vsb_kepler_curves_device_ptr: no curves yet");
    *out = k->curves.p;
    return VSB_OK;
}

// ---------------------------------------------------------------- N1: Newton <-> Kepler
extern "C" int vsb_host_to_kepler(vsb_ctx *c, int64_t n, const int64_t *order, const double *mass,
                                  const double *pos, const double *vel, double gc, double coi_coef,
                                  int64_t *ref_out, double *a_out, double *ecc_out,
                                  double *pe_arg_out, double *ma_out, double *dir_out)
{
    CTX(c);
    VSB_CHECK(n > 0 && order && mass && pos && vel && ref_out && a_out && ecc_out && pe_arg_out &&
                  ma_out && dir_out,
              VSB_ERR_ARG, "vsb_host_to_kepler: bad argument");
    VSB_TRY(stage_bodies(c, n));
    VSB_CUDA(cudaMemcpyAsync(c->mass, mass, 8 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    VSB_CUDA(cudaMemcpyAsync(c->pos, pos, 16 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    VSB_CUDA(cudaMemcpyAsync(c->vel, vel, 16 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    VSB_TRY(upload_order(c, "vsb_host_to_kepler", order));
    VSB_TRY(vsb_buf_reserve(c->host_stage, 48 * (size_t)c->npad));
    int64_t *d_ref = (int64_t *)c->host_stage.p;
    double *d_el = (double *)c->host_stage.p + c->npad;
    VSB_TRY(vsb_launch_to_kepler(c, (const int64_t *)c->order.p, gc, coi_coef, d_ref, d_el));
    double *outs[5] = {a_out, ecc_out, pe_arg_out, ma_out, dir_out};
    for (int k = 0; k < 5; ++k)
        VSB_CUDA(cudaMemcpyAsync(outs[k], d_el + (size_t)k * n, 8 * (size_t)n,
                                 cudaMemcpyDeviceToHost, c->stream));
    return d2h(c, ref_out, d_ref, 8 * (size_t)n);
}

extern "C" int vsb_host_to_newton(vsb_ctx *c, int64_t n, const double *mass, const double *a,
                                  const double *ecc, const double *pe_arg, const double *ma,
                                  const int64_t *ref, const double *dir, double gc,
                                  double *pos_out, double *vel_out)
{
    CTX(c);
    VSB_CHECK(n > 0 && mass && a && ecc && pe_arg && ma && ref && dir && pos_out && vel_out,
              VSB_ERR_ARG, "vsb_host_to_newton: bad argument");
    for (int64_t i = 0; i < n; ++i)
        VSB_CHECK(ref[i] >= 0 && ref[i] < n, VSB_ERR_ARG, "vsb_host_to_newton: ref[%lld] out of range",
                  (long long)i);
    VSB_TRY(stage_bodies(c, n));
    VSB_CUDA(cudaMemcpyAsync(c->mass, mass, 8 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    VSB_TRY(vsb_buf_reserve(c->host_stage, 48 * (size_t)c->npad));
    int64_t *d_ref = (int64_t *)c->host_stage.p;
    double *d_el = (double *)c->host_stage.p + c->npad;
    const double *ins[5] = {a, ecc, pe_arg, ma, dir};
    for (int k = 0; k < 5; ++k)
        VSB_CUDA(cudaMemcpyAsync(d_el + (size_t)k * n, ins[k], 8 * (size_t)n,
                                 cudaMemcpyHostToDevice, c->stream));
    VSB_CUDA(cudaMemcpyAsync(d_ref, ref, 8 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    int err = 0;
    VSB_TRY(vsb_launch_to_newton(c, d_el, d_ref, gc, &err));
    VSB_CHECK(err != 3, VSB_ERR_ARG, "vsb_host_to_newton: ref hierarchy deeper than 64 levels");
    // the reference raises ValueError (math domain error) for hyperbolic bodies, convert.py:170
    VSB_CHECK(err == 0, VSB_ERR_ARG, "vsb_host_to_newton: math domain error");
    VSB_CUDA(cudaMemcpyAsync(pos_out, c->pos, 16 * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    return d2h(c, vel_out, c->vel, 16 * (size_t)n);
}

// ---------------------------------------------------------------- N2: COI-crossing numeric kernels
extern "C" int vsb_host_solve_quartic(vsb_ctx *c, int64_t n, const double *coef5, double *roots8)
{
    CTX(c);
    VSB_CHECK(n >= 0 && (n == 0 || (coef5 && roots8)), VSB_ERR_ARG,
              "vsb_host_solve_quartic: bad argument");
    if (n == 0) return VSB_OK;
    VSB_TRY(vsb_buf_reserve(c->host_stage, 8 * 13 * (size_t)n));
    double *d_in = (double *)c->host_stage.p, *d_out = d_in + 5 * n;
    VSB_CUDA(cudaMemcpyAsync(d_in, coef5, 40 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    VSB_TRY(vsb_launch_quartic(c, n, d_in, d_out));
    return d2h(c, roots8, d_out, 64 * (size_t)n);
}

extern "C" int vsb_host_ell_hyp_intersect_circle(vsb_ctx *c, int64_t n, const double *in6,
                                                 double *ea4_out, int64_t *count_out)
{
    CTX(c);
    VSB_CHECK(n >= 0 && (n == 0 || (in6 && ea4_out && count_out)), VSB_ERR_ARG,
              "vsb_host_ell_hyp_intersect_circle: bad argument");
    if (n == 0) return VSB_OK;
    VSB_TRY(vsb_buf_reserve(c->host_stage, 8 * 11 * (size_t)n));
    double *d_in = (double *)c->host_stage.p, *d_ea = d_in + 6 * n;
    int64_t *d_cnt = (int64_t *)(d_ea + 4 * n);
    VSB_CUDA(cudaMemcpyAsync(d_in, in6, 48 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    VSB_TRY(vsb_launch_intersect(c, n, d_in, d_ea, d_cnt));
    VSB_CUDA(cudaMemcpyAsync(ea4_out, d_ea, 32 * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    return d2h(c, count_out, d_cnt, 8 * (size_t)n);
}

static int march_work_reserve(vsb_ctx *c, int64_t n)
{
    return vsb_buf_reserve(c->march_stats, vsb_march_work_bytes(n));
}

extern "C" int vsb_host_predict_enter_coi(vsb_ctx *c, int64_t n, const double *vessel_data10,
                                          const double *body_data10, double tol,
                                          int64_t steps_limit, double *out5)
{
    CTX(c);
    VSB_CHECK(n >= 0 && steps_limit > 0 && steps_limit < (1ll << 31) &&
                  (n == 0 || (vessel_data10 && body_data10 && out5)),
              VSB_ERR_ARG, "vsb_host_predict_enter_coi: bad argument");
    if (n == 0) return VSB_OK;
    VSB_TRY(march_work_reserve(c, n));
    VSB_TRY(vsb_buf_reserve(c->host_stage, 8 * 25 * (size_t)n));
    double *d_vd = (double *)c->host_stage.p, *d_bd = d_vd + 10 * n, *d_out = d_bd + 10 * n;
    VSB_CUDA(cudaMemcpyAsync(d_vd, vessel_data10, 80 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    VSB_CUDA(cudaMemcpyAsync(d_bd, body_data10, 80 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    VSB_TRY(vsb_launch_predict_enter_coi(c, n, d_vd, d_bd, nullptr, tol, (int)steps_limit, d_out,
                                         c->march_stats.p, c->march_sequential));
    return d2h(c, out5, d_out, 40 * (size_t)n);
}

extern "C" int vsb_march_stats(vsb_ctx *c, uint64_t *out4)
{
    CTX(c);
    VSB_CHECK(out4 != nullptr, VSB_ERR_ARG, "vsb_march_stats: null output");
    out4[0] = out4[1] = out4[2] = out4[3] = 0;
    if (!c->march_stats.p) return VSB_OK;
    uint64_t raw[7];     // mteam: stats[4], ntasks, next_task, refused
    VSB_TRY(d2h(c, raw, c->march_stats.p, sizeof(raw)));
    out4[0] = raw[0];
    out4[1] = raw[1];
    out4[2] = raw[2];
    out4[3] = raw[6];
    return VSB_OK;
}

extern "C" int vsb_host_running_sum(double x0, double c, int64_t kmax, int bounded, double bound,
                                    int64_t nq, const int64_t *ks, double *xs, int64_t *K_out,
                                    int64_t *nseg_out)
{
    VSB_CHECK(kmax >= 0 && nq >= 0 && (nq == 0 || (ks && xs)), VSB_ERR_ARG,
              "vsb_host_running_sum: bad argument");
    const int rc = vsb_march_walk_host(x0, c, kmax, bounded, bound, nq, ks, xs, K_out, nseg_out);
    VSB_CHECK(rc >= 0, VSB_ERR_ARG, "vsb_host_running_sum: step index out of range");
    VSB_CHECK(rc == 0, VSB_ERR_STATE, "vsb_host_running_sum: no closed form (code %d)", rc);
    return VSB_OK;
}

extern "C" int vsb_set_march_mode(vsb_ctx *c, int mode)
{
    CTX(c);
    VSB_CHECK(mode >= 0 && mode <= 3, VSB_ERR_ARG, "vsb_set_march_mode: mode %d", mode);
    c->march_sequential = mode;        // bit 0: serial loop, bit 1: exact trigonometry
    return VSB_OK;
}

// staging helper: consecutive 256-byte aligned device arrays inside host_stage
struct stage_cursor {
    char *base;
    size_t off = 0;
    void *take(size_t bytes)
    {
        void *p = base + off;
        off += (bytes + 255) / 256 * 256;
        return p;
    }
};

static int up_f64(vsb_ctx *c, double *dst, const double *src, size_t count)
{
    if (count == 0) return VSB_OK;
    VSB_CUDA(cudaMemcpyAsync(dst, src, 8 * count, cudaMemcpyHostToDevice, c->stream));
    return VSB_OK;
}

// ---- phys_vessel.Physics.points (phys_vessel.py:372-474): shared between the host-array call
// and the resident vessel state.  The (vessel, candidate body) pairs are listed on the host:
// check_bodies = np.where(body_ref == ref)[0] minus body 0 and left_coi_prev_ref (:423-424).
struct points_plan {
    std::vector<int64_t> sel, offsets, pair_s, pair_body;
};

static int points_plan_build(const char *who, int64_t nv, int64_t nb, const int64_t *v_ref,
                             const int64_t *body_ref, int64_t nsel, const int64_t *sel,
                             int64_t left_coi_prev_ref, points_plan &pl)
{
    if (sel) {
        pl.sel.assign(sel, sel + nsel);
    } else {
        pl.sel.resize((size_t)nv);
        for (int64_t i = 0; i < nv; ++i) pl.sel[(size_t)i] = i;
    }
    for (int64_t v : pl.sel)
        VSB_CHECK(v >= 0 && v < nv, VSB_ERR_ARG, "%s: vessel %lld out of range", who, (long long)v);
    for (int64_t i = 0; i < nv; ++i)
        VSB_CHECK(v_ref[i] >= 0 && v_ref[i] < nb, VSB_ERR_ARG, "%s: ref of vessel %lld out of range",
                  who, (long long)i);
    // children lists by counting sort (ascending body index inside each list)
    std::vector<int64_t> cstart((size_t)nb + 1, 0), child((size_t)nb);
    for (int64_t b = 0; b < nb; ++b)
        if (body_ref[b] >= 0 && body_ref[b] < nb) ++cstart[(size_t)body_ref[b] + 1];
    for (int64_t r = 0; r < nb; ++r) cstart[(size_t)r + 1] += cstart[(size_t)r];
    {
        std::vector<int64_t> fill(cstart.begin(), cstart.end() - 1);
        for (int64_t b = 0; b < nb; ++b)
            if (body_ref[b] >= 0 && body_ref[b] < nb) child[(size_t)fill[(size_t)body_ref[b]]++] = b;
    }
    const size_t ns = pl.sel.size();
    pl.offsets.assign(ns + 1, 0);
    for (size_t s = 0; s < ns; ++s) {
        const int64_t ref = v_ref[pl.sel[s]];
        for (int64_t k = cstart[(size_t)ref]; k < cstart[(size_t)ref + 1]; ++k) {
            const int64_t b = child[(size_t)k];
            if (b == 0 || b == left_coi_prev_ref) continue;
            pl.pair_s.push_back((int64_t)s);
            pl.pair_body.push_back(b);
        }
        pl.offsets[s + 1] = (int64_t)pl.pair_s.size();
    }
    return VSB_OK;
}

static size_t al32(size_t w) { return (w + 31) / 32 * 32; }

// 8-byte words of device scratch the pipeline needs for a plan
static size_t points_scratch_words(const points_plan &pl)
{
    const size_t NS = pl.sel.size(), NP = pl.pair_s.size();
    return al32(NS) + al32(NS + 1) + 2 * al32(NP) + al32(NS) + 2 * al32(10 * NP) + al32(5 * NP) +
           al32((NP + 7) / 8);
}

// scratch: device memory of points_scratch_words(pl) words.  The four result arrays are updated
// in place for the selected vessels.
static int points_run(vsb_ctx *c, const points_plan &pl, double *scratch, const double *d_v12,
                      const int64_t *d_vref, const double *d_b12, int64_t entered_coi,
                      int64_t left_coi, int only_enter_coi, int64_t steps_limit, double *d_imp,
                      double *d_atm, double *d_lv, double *d_en)
{
    const size_t NS = pl.sel.size(), NP = pl.pair_s.size();
    if (NS == 0) return VSB_OK;
    size_t w = 0;
    auto take = [&](size_t words) { double *p = scratch + w; w += al32(words); return p; };
    int64_t *d_sel = (int64_t *)take(NS), *d_off = (int64_t *)take(NS + 1),
            *d_ps = (int64_t *)take(NP), *d_pb = (int64_t *)take(NP);
    double *d_eaeff = take(NS), *d_vd = take(10 * NP), *d_bd = take(10 * NP), *d_out = take(5 * NP);
    unsigned char *d_valid = (unsigned char *)take((NP + 7) / 8);
    auto up = [&](void *dst, const void *src, size_t words) -> int {
        if (words == 0) return VSB_OK;
        VSB_CUDA(cudaMemcpyAsync(dst, src, 8 * words, cudaMemcpyHostToDevice, c->stream));
        return VSB_OK;
    };
    VSB_TRY(up(d_sel, pl.sel.data(), NS));
    VSB_TRY(up(d_off, pl.offsets.data(), NS + 1));
    VSB_TRY(up(d_ps, pl.pair_s.data(), NP));
    VSB_TRY(up(d_pb, pl.pair_body.data(), NP));
    VSB_TRY(vsb_buf_reserve(c->march_stats, vsb_march_work_bytes(NP > 0 ? (int64_t)NP : 1)));
    VSB_CUDA(cudaMemsetAsync(c->march_stats.p, 0, 64, c->stream));
    VSB_TRY(vsb_launch_points_vessel(c, (int64_t)NS, d_sel, d_v12, d_vref, d_b12, entered_coi,
                                     left_coi, only_enter_coi, d_imp, d_atm, d_lv, d_eaeff));
    VSB_TRY(vsb_launch_points_pairs(c, (int64_t)NP, d_ps, d_pb, d_sel, d_v12, d_vref, d_b12, left_coi,
                                    d_lv, d_eaeff, d_vd, d_bd, d_valid));
    VSB_TRY(vsb_launch_predict_enter_coi(c, (int64_t)NP, d_vd, d_bd, d_valid, 1e-5, (int)steps_limit,
                                         d_out, c->march_stats.p, c->march_sequential));
    VSB_TRY(vsb_launch_points_select(c, (int64_t)NS, d_sel, d_off, d_pb, d_valid, d_out, d_eaeff,
                                     d_v12, d_en));
    return VSB_OK;
}

// phys_vessel.Physics.points for the vessels in sel (NULL = all), over host arrays
extern "C" int vsb_host_vessel_points(vsb_ctx *c, int64_t nv, int64_t nb, const double *vessel12,
                                      const int64_t *v_ref, const double *body12,
                                      const int64_t *body_ref, int64_t nsel, const int64_t *sel,
                                      int64_t entered_coi, int64_t left_coi,
                                      int64_t left_coi_prev_ref, int only_enter_coi,
                                      int64_t steps_limit, double *body_impact,
                                      double *body_enter_atm, double *coi_leave, double *coi_enter)
{
    CTX(c);
    VSB_CHECK(nv >= 0 && nb >= 0 && nsel >= 0 && steps_limit > 0 && steps_limit < (1ll << 31),
              VSB_ERR_ARG, "vsb_host_vessel_points: bad argument");
    if (nv == 0) return VSB_OK;
    VSB_CHECK(vessel12 && v_ref && body12 && body_ref && body_impact && body_enter_atm && coi_leave &&
                  coi_enter && nb > 0,
              VSB_ERR_ARG, "vsb_host_vessel_points: null array");
    points_plan pl;
    VSB_TRY(points_plan_build("vsb_host_vessel_points", nv, nb, v_ref, body_ref, nsel, sel,
                              left_coi_prev_ref, pl));
    if (pl.sel.empty()) return VSB_OK;
    const size_t NV = (size_t)nv, NB = (size_t)nb;
    const size_t own = al32(12 * NV) + al32(NV) + al32(12 * NB) + 3 * al32(2 * NV) + al32(6 * NV);
    VSB_TRY(vsb_buf_reserve(c->host_stage, 8 * (own + points_scratch_words(pl))));
    double *base = (double *)c->host_stage.p;
    size_t w = 0;
    auto take = [&](size_t words) { double *p = base + w; w += al32(words); return p; };
    double *d_v12 = take(12 * NV);
    int64_t *d_vref = (int64_t *)take(NV);
    double *d_b12 = take(12 * NB), *d_imp = take(2 * NV), *d_atm = take(2 * NV), *d_lv = take(2 * NV),
           *d_en = take(6 * NV);
    VSB_TRY(up_f64(c, d_v12, vessel12, 12 * NV));
    VSB_CUDA(cudaMemcpyAsync(d_vref, v_ref, 8 * NV, cudaMemcpyHostToDevice, c->stream));
    VSB_TRY(up_f64(c, d_b12, body12, 12 * NB));
    VSB_TRY(up_f64(c, d_imp, body_impact, 2 * NV));
    VSB_TRY(up_f64(c, d_atm, body_enter_atm, 2 * NV));
    VSB_TRY(up_f64(c, d_lv, coi_leave, 2 * NV));
    VSB_TRY(up_f64(c, d_en, coi_enter, 6 * NV));
    VSB_TRY(points_run(c, pl, base + w, d_v12, d_vref, d_b12, entered_coi, left_coi, only_enter_coi,
                       steps_limit, d_imp, d_atm, d_lv, d_en));
    VSB_CUDA(cudaMemcpyAsync(body_impact, d_imp, 16 * NV, cudaMemcpyDeviceToHost, c->stream));
    VSB_CUDA(cudaMemcpyAsync(body_enter_atm, d_atm, 16 * NV, cudaMemcpyDeviceToHost, c->stream));
    VSB_CUDA(cudaMemcpyAsync(coi_leave, d_lv, 16 * NV, cudaMemcpyDeviceToHost, c->stream));
    return d2h(c, coi_enter, d_en, 48 * NV);
}

// ---------------------------------------------------------------- N2: per-tick vessel events
extern "C" int vsb_host_check_warp(vsb_ctx *c, int64_t nv, const double *ecc, const double *dr,
                                   const double *ea, const double *ma, const double *n_motion,
                                   const double *body_impact, const double *body_enter_atm,
                                   const double *coi_leave, const double *coi_enter, double warp,
                                   uint8_t *hold, int64_t *out3)
{
    CTX(c);
    VSB_CHECK(nv >= 0 && out3, VSB_ERR_ARG, "vsb_host_check_warp: bad argument");
    out3[0] = out3[1] = -1;
    out3[2] = 0;
    if (nv == 0) return VSB_OK;
    VSB_CHECK(ecc && dr && ea && ma && n_motion && body_impact && body_enter_atm && coi_leave &&
                  coi_enter && hold,
              VSB_ERR_ARG, "vsb_host_check_warp: null array");
    const size_t N = (size_t)nv;
    VSB_TRY(vsb_buf_reserve(c->host_stage, 8 * 17 * N + N + 20 * 256));
    stage_cursor sc{(char *)c->host_stage.p};
    double *d5[5];
    const double *h5[5] = {ecc, dr, ea, ma, n_motion};
    for (int k = 0; k < 5; ++k) {
        d5[k] = (double *)sc.take(8 * N);
        VSB_TRY(up_f64(c, d5[k], h5[k], N));
    }
    double *d_imp = (double *)sc.take(16 * N), *d_atm = (double *)sc.take(16 * N),
           *d_lv = (double *)sc.take(16 * N), *d_en = (double *)sc.take(48 * N);
    VSB_TRY(up_f64(c, d_imp, body_impact, 2 * N));
    VSB_TRY(up_f64(c, d_atm, body_enter_atm, 2 * N));
    VSB_TRY(up_f64(c, d_lv, coi_leave, 2 * N));
    VSB_TRY(up_f64(c, d_en, coi_enter, 6 * N));
    unsigned char *d_hold = (unsigned char *)sc.take(N);
    unsigned long long *d_key = (unsigned long long *)sc.take(16);
    VSB_CUDA(cudaMemcpyAsync(d_hold, hold, N, cudaMemcpyHostToDevice, c->stream));
    VSB_TRY(vsb_launch_check_warp(c, nv, d5[0], d5[1], d5[2], d5[3], d5[4], d_imp, d_atm, d_lv, d_en,
                                  warp, d_hold, d_key));
    unsigned long long key[2];
    VSB_CUDA(cudaMemcpyAsync(hold, d_hold, N, cudaMemcpyDeviceToHost, c->stream));
    VSB_TRY(d2h(c, key, d_key, 16));
    if (key[0]) {
        out3[0] = (int64_t)(key[0] & 3);
        out3[1] = (int64_t)(key[0] >> 2) - 1;
    }
    out3[2] = (int64_t)key[1];
    return VSB_OK;
}

extern "C" int vsb_host_enter_coi_service(vsb_ctx *c, int64_t nv, const double *dr,
                                          const double *prev_ea, const double *ea,
                                          const double *coi_enter, int crossing_pending,
                                          uint8_t *flags_out)
{
    CTX(c);
    VSB_CHECK(nv >= 0, VSB_ERR_ARG, "vsb_host_enter_coi_service: bad argument");
    if (nv == 0) return VSB_OK;
    VSB_CHECK(dr && prev_ea && ea && coi_enter && flags_out, VSB_ERR_ARG,
              "vsb_host_enter_coi_service: null array");
    const size_t N = (size_t)nv;
    VSB_TRY(vsb_buf_reserve(c->host_stage, 8 * 9 * N + N + 8 * 256));
    stage_cursor sc{(char *)c->host_stage.p};
    double *d_dr = (double *)sc.take(8 * N), *d_pe = (double *)sc.take(8 * N),
           *d_ea = (double *)sc.take(8 * N), *d_en = (double *)sc.take(48 * N);
    unsigned char *d_fl = (unsigned char *)sc.take(N);
    VSB_TRY(up_f64(c, d_dr, dr, N));
    VSB_TRY(up_f64(c, d_pe, prev_ea, N));
    VSB_TRY(up_f64(c, d_ea, ea, N));
    VSB_TRY(up_f64(c, d_en, coi_enter, 6 * N));
    VSB_TRY(vsb_launch_enter_coi_service(c, nv, d_dr, d_pe, d_ea, d_en, crossing_pending, d_fl));
    return d2h(c, flags_out, d_fl, N);
}

extern "C" int vsb_host_cross_scan(vsb_ctx *c, int64_t nv, const double *ecc, const double *dr,
                                   const double *prev_ea, const double *ea,
                                   const double *body_impact, const double *coi_leave,
                                   const double *coi_enter, int64_t *out2)
{
    CTX(c);
    VSB_CHECK(nv >= 0 && out2, VSB_ERR_ARG, "vsb_host_cross_scan: bad argument");
    out2[0] = -1;
    out2[1] = 0;
    if (nv == 0) return VSB_OK;
    VSB_CHECK(ecc && dr && prev_ea && ea && body_impact && coi_leave && coi_enter, VSB_ERR_ARG,
              "vsb_host_cross_scan: null array");
    const size_t N = (size_t)nv;
    VSB_TRY(vsb_buf_reserve(c->host_stage, 8 * 14 * N + 12 * 256));
    stage_cursor sc{(char *)c->host_stage.p};
    double *d4[4];
    const double *h4[4] = {ecc, dr, prev_ea, ea};
    for (int k = 0; k < 4; ++k) {
        d4[k] = (double *)sc.take(8 * N);
        VSB_TRY(up_f64(c, d4[k], h4[k], N));
    }
    double *d_imp = (double *)sc.take(16 * N), *d_lv = (double *)sc.take(16 * N),
           *d_en = (double *)sc.take(48 * N);
    unsigned long long *d_key = (unsigned long long *)sc.take(8);
    VSB_TRY(up_f64(c, d_imp, body_impact, 2 * N));
    VSB_TRY(up_f64(c, d_lv, coi_leave, 2 * N));
    VSB_TRY(up_f64(c, d_en, coi_enter, 6 * N));
    VSB_TRY(vsb_launch_cross_scan(c, nv, d4[0], d4[1], d4[2], d4[3], d_imp, d_lv, d_en, d_key));
    unsigned long long key;
    VSB_TRY(d2h(c, &key, d_key, 8));
    if (key != ~0ull) {
        out2[0] = (int64_t)(key >> 2);
        out2[1] = (int64_t)(key & 3);
    }
    return VSB_OK;
}

extern "C" int vsb_host_cross_apply(vsb_ctx *c, int64_t nev, const int32_t *kind,
                                    const double *vessel8, const int64_t *ref,
                                    const int64_t *new_ref, int64_t nb, const double *body7,
                                    double gc, double *out6, int32_t *status)
{
    CTX(c);
    VSB_CHECK(nev >= 0 && nb >= 0, VSB_ERR_ARG, "vsb_host_cross_apply: bad argument");
    if (nev == 0) return VSB_OK;
    VSB_CHECK(kind && vessel8 && ref && new_ref && body7 && out6 && status && nb > 0, VSB_ERR_ARG,
              "vsb_host_cross_apply: null array");
    for (int64_t i = 0; i < nev; ++i)
        VSB_CHECK((kind[i] == 1 || kind[i] == 2) && ref[i] >= 0 && ref[i] < nb && new_ref[i] >= 0 &&
                      new_ref[i] < nb,
                  VSB_ERR_ARG, "vsb_host_cross_apply: event %lld: bad kind or body index",
                  (long long)i);
    const size_t E = (size_t)nev, B = (size_t)nb;
    VSB_TRY(vsb_buf_reserve(c->host_stage, 8 * (16 * E + 7 * B) + 8 * E + 12 * 256));
    stage_cursor sc{(char *)c->host_stage.p};
    double *d_v8 = (double *)sc.take(64 * E), *d_b7 = (double *)sc.take(56 * B),
           *d_out = (double *)sc.take(48 * E);
    int64_t *d_ref = (int64_t *)sc.take(8 * E), *d_new = (int64_t *)sc.take(8 * E);
    int32_t *d_kind = (int32_t *)sc.take(4 * E), *d_st = (int32_t *)sc.take(4 * E);
    VSB_TRY(up_f64(c, d_v8, vessel8, 8 * E));
    VSB_TRY(up_f64(c, d_b7, body7, 7 * B));
    VSB_CUDA(cudaMemcpyAsync(d_ref, ref, 8 * E, cudaMemcpyHostToDevice, c->stream));
    VSB_CUDA(cudaMemcpyAsync(d_new, new_ref, 8 * E, cudaMemcpyHostToDevice, c->stream));
    VSB_CUDA(cudaMemcpyAsync(d_kind, kind, 4 * E, cudaMemcpyHostToDevice, c->stream));
    VSB_TRY(vsb_launch_cross_apply(c, nev, d_kind, d_v8, d_ref, d_new, d_b7, gc, d_out, d_st));
    VSB_CUDA(cudaMemcpyAsync(status, d_st, 4 * E, cudaMemcpyDeviceToHost, c->stream));
    return d2h(c, out6, d_out, 48 * E);
}

// ---------------------------------------------------------------- N2: resident vessel events
// The orbit points and the per-tick event state of phys_vessel.Physics kept on the device next
// to the vessel elements (KIND_VESSEL state), so that a game tick moves no per-vessel array over
// PCIe: move -> cross_scan -> check_warp -> service all read resident data.
static int need_vessel_events(vsb_ctx *c, const char *who, vsb_kepler_state **out)
{
    vsb_kepler_state &k = c->kep[VSB_KIND_VESSEL];
    VSB_CHECK(k.n > 0, VSB_ERR_STATE, "%s: no VESSEL state loaded", who);
    VSB_CHECK(k.ev_n == k.n && k.ev.p, VSB_ERR_STATE, "%s: call vsb_vessel_events_load first", who);
    *out = &k;
    return VSB_OK;
}

extern "C" int vsb_vessel_events_load(vsb_ctx *c, const double *body_impact,
                                      const double *body_enter_atm, const double *coi_leave,
                                      const double *coi_enter, const uint8_t *hold,
                                      const double *prev_ea, const double *pe_d,
                                      const double *ap_d, const double *period)
{
    CTX(c);
    vsb_kepler_state &k = c->kep[VSB_KIND_VESSEL];
    VSB_CHECK(k.n > 0, VSB_ERR_STATE, "vsb_vessel_events_load: no VESSEL state loaded");
    VSB_CHECK(pe_d && ap_d && period, VSB_ERR_ARG,
              "vsb_vessel_events_load: pe_d, ap_d and period are required");
    const size_t N = (size_t)k.n, A = al32(N), A2 = al32(2 * N), A6 = al32(6 * N);
    VSB_TRY(vsb_buf_reserve(k.ev, 8 * (3 * A2 + A6 + 4 * A) + N + 256));
    double *b = (double *)k.ev.p;
    k.ev_impact = b;
    k.ev_atm = b + A2;
    k.ev_leave = b + 2 * A2;
    k.ev_enter = b + 3 * A2;
    k.ev_prev_ea = b + 3 * A2 + A6;
    k.ev_pe_d = k.ev_prev_ea + A;
    k.ev_ap_d = k.ev_pe_d + A;
    k.ev_period = k.ev_ap_d + A;
    k.ev_hold = (unsigned char *)(k.ev_period + A);
    k.ev_n = k.n;
    // NaN = "no point" (the reference initialises these arrays with NaN, phys_vessel.py:231-234)
    VSB_CUDA(cudaMemsetAsync(b, 0xff, 8 * (3 * A2 + A6), c->stream));
    VSB_CUDA(cudaMemsetAsync(k.ev_hold, 0, N, c->stream));
    if (body_impact) VSB_TRY(up_f64(c, k.ev_impact, body_impact, 2 * N));
    if (body_enter_atm) VSB_TRY(up_f64(c, k.ev_atm, body_enter_atm, 2 * N));
    if (coi_leave) VSB_TRY(up_f64(c, k.ev_leave, coi_leave, 2 * N));
    if (coi_enter) VSB_TRY(up_f64(c, k.ev_enter, coi_enter, 6 * N));
    if (hold) VSB_CUDA(cudaMemcpyAsync(k.ev_hold, hold, N, cudaMemcpyHostToDevice, c->stream));
    if (prev_ea) VSB_TRY(up_f64(c, k.ev_prev_ea, prev_ea, N));
    else VSB_CUDA(cudaMemcpyAsync(k.ev_prev_ea, k.ea, 8 * N, cudaMemcpyDeviceToDevice, c->stream));
    VSB_TRY(up_f64(c, k.ev_pe_d, pe_d, N));
    VSB_TRY(up_f64(c, k.ev_ap_d, ap_d, N));
    VSB_TRY(up_f64(c, k.ev_period, period, N));
    VSB_CUDA(cudaStreamSynchronize(c->stream));
    return VSB_OK;
}

extern "C" int vsb_vessel_events_set_orbit(vsb_ctx *c, int64_t vessel, double pe_d, double ap_d,
                                           double period)
{
    CTX(c);
    vsb_kepler_state *k;
    VSB_TRY(need_vessel_events(c, "vsb_vessel_events_set_orbit", &k));
    VSB_CHECK(vessel >= 0 && vessel < k->n, VSB_ERR_ARG,
              "vsb_vessel_events_set_orbit: vessel %lld of %lld", (long long)vessel, (long long)k->n);
    VSB_TRY(up_f64(c, k->ev_pe_d + vessel, &pe_d, 1));
    VSB_TRY(up_f64(c, k->ev_ap_d + vessel, &ap_d, 1));
    VSB_TRY(up_f64(c, k->ev_period + vessel, &period, 1));
    VSB_CUDA(cudaStreamSynchronize(c->stream));
    return VSB_OK;
}

extern "C" int vsb_vessel_events_get(vsb_ctx *c, int64_t first, int64_t count, double *body_impact,
                                     double *body_enter_atm, double *coi_leave, double *coi_enter,
                                     uint8_t *hold, double *prev_ea)
{
    CTX(c);
    vsb_kepler_state *k;
    VSB_TRY(need_vessel_events(c, "vsb_vessel_events_get", &k));
    VSB_CHECK(first >= 0 && count >= 0 && first + count <= k->n, VSB_ERR_ARG,
              "vsb_vessel_events_get: range [%lld,+%lld) of %lld", (long long)first,
              (long long)count, (long long)k->n);
    const size_t F = (size_t)first, C = (size_t)count;
    auto dn = [&](void *dst, const void *src, size_t bytes) -> int {
        if (dst && bytes)
            VSB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
        return VSB_OK;
    };
    VSB_TRY(dn(body_impact, k->ev_impact + 2 * F, 16 * C));
    VSB_TRY(dn(body_enter_atm, k->ev_atm + 2 * F, 16 * C));
    VSB_TRY(dn(coi_leave, k->ev_leave + 2 * F, 16 * C));
    VSB_TRY(dn(coi_enter, k->ev_enter + 6 * F, 48 * C));
    VSB_TRY(dn(hold, k->ev_hold + F, C));
    VSB_TRY(dn(prev_ea, k->ev_prev_ea + F, 8 * C));
    VSB_CUDA(cudaStreamSynchronize(c->stream));
    return VSB_OK;
}

extern "C" int vsb_vessel_points(vsb_ctx *c, int64_t nb, const double *body12,
                                 const int64_t *body_ref, int64_t nsel, const int64_t *sel,
                                 int64_t entered_coi, int64_t left_coi, int64_t left_coi_prev_ref,
                                 int only_enter_coi, int64_t steps_limit)
{
    CTX(c);
    vsb_kepler_state *k;
    VSB_TRY(need_vessel_events(c, "vsb_vessel_points", &k));
    VSB_CHECK(nb > 0 && body12 && body_ref && nsel >= 0 && steps_limit > 0 &&
                  steps_limit < (1ll << 31),
              VSB_ERR_ARG, "vsb_vessel_points: bad argument");
    points_plan pl;
    VSB_TRY(points_plan_build("vsb_vessel_points", k->n, nb, k->h_ref, body_ref, nsel, sel,
                              left_coi_prev_ref, pl));
    if (pl.sel.empty()) return VSB_OK;
    const size_t NV = (size_t)k->n, NB = (size_t)nb;
    VSB_TRY(vsb_buf_reserve(c->host_stage,
                            8 * (al32(12 * NV) + al32(12 * NB) + points_scratch_words(pl))));
    double *base = (double *)c->host_stage.p;
    double *d_v12 = base, *d_b12 = base + al32(12 * NV);
    VSB_TRY(vsb_launch_pack_vessel12(c, k->n, k->a, k->b, k->f, k->ecc, k->ma, k->ea, k->pea, k->nmot,
                                     k->dr, k->ev_pe_d, k->ev_ap_d, k->ev_period, d_v12));
    VSB_TRY(up_f64(c, d_b12, body12, 12 * NB));
    VSB_TRY(points_run(c, pl, d_b12 + al32(12 * NB), d_v12, k->ref, d_b12, entered_coi, left_coi,
                       only_enter_coi, steps_limit, k->ev_impact, k->ev_atm, k->ev_leave,
                       k->ev_enter));
    VSB_CUDA(cudaStreamSynchronize(c->stream));
    return VSB_OK;
}

extern "C" int vsb_vessel_check_warp(vsb_ctx *c, double warp, int64_t *out3)
{
    CTX(c);
    vsb_kepler_state *k;
    VSB_TRY(need_vessel_events(c, "vsb_vessel_check_warp", &k));
    VSB_CHECK(out3 != nullptr, VSB_ERR_ARG, "vsb_vessel_check_warp: null output");
    VSB_TRY(vsb_buf_reserve(c->scratch, 64));
    unsigned long long *d_key = (unsigned long long *)c->scratch.p;
    VSB_TRY(vsb_launch_check_warp(c, k->n, k->ecc, k->dr, k->ea, k->ma, k->nmot, k->ev_impact,
                                  k->ev_atm, k->ev_leave, k->ev_enter, warp, k->ev_hold, d_key));
    unsigned long long key[2];
    {
        const void *src[1] = {d_key};
        const int words[1] = {2};
        void *dst[1] = {key};
        VSB_TRY(d2h_small(c, 1, src, words, dst, "vsb_vessel_check_warp"));
    }
    out3[0] = out3[1] = -1;
    if (key[0]) {
        out3[0] = (int64_t)(key[0] & 3);
        out3[1] = (int64_t)(key[0] >> 2) - 1;
    }
    out3[2] = (int64_t)key[1];
    return VSB_OK;
}

extern "C" int vsb_vessel_cross_scan(vsb_ctx *c, int64_t *out2)
{
    CTX(c);
    vsb_kepler_state *k;
    VSB_TRY(need_vessel_events(c, "vsb_vessel_cross_scan", &k));
    VSB_CHECK(out2 != nullptr, VSB_ERR_ARG, "vsb_vessel_cross_scan: null output");
    VSB_TRY(vsb_buf_reserve(c->scratch, 64));
    unsigned long long *d_key = (unsigned long long *)c->scratch.p;
    VSB_TRY(vsb_launch_cross_scan(c, k->n, k->ecc, k->dr, k->ev_prev_ea, k->ea, k->ev_impact,
                                  k->ev_leave, k->ev_enter, d_key));
    unsigned long long key;
    {
        const void *src[1] = {d_key};
        const int words[1] = {1};
        void *dst[1] = {&key};
        VSB_TRY(d2h_small(c, 1, src, words, dst, "vsb_vessel_cross_scan"));
    }
    out2[0] = -1;
    out2[1] = 0;
    if (key != ~0ull) {
        out2[0] = (int64_t)(key >> 2);
        out2[1] = (int64_t)(key & 3);
    }
    return VSB_OK;
}

// ---------------------------------------------------------------- stateless element conversions
// generic "upload k input arrays, run, download one output" helper for per-item kernels
struct in_arr {
    const void *host;
    size_t words;    // 8-byte words
};

static int stage_inputs(vsb_ctx *c, const in_arr *ins, int nin, size_t out_words, double **d_ins,
                        double **d_out)
{
    size_t total = al32(out_words);
    for (int k = 0; k < nin; ++k) total += al32(ins[k].words);
    VSB_TRY(vsb_buf_reserve(c->host_stage, 8 * total));
    double *p = (double *)c->host_stage.p;
    for (int k = 0; k < nin; ++k) {
        d_ins[k] = p;
        if (ins[k].words)
            VSB_CUDA(cudaMemcpyAsync(p, ins[k].host, 8 * ins[k].words, cudaMemcpyHostToDevice,
                                     c->stream));
        p += al32(ins[k].words);
    }
    *d_out = p;
    return VSB_OK;
}

extern "C" int vsb_host_kepler_to_velocity(vsb_ctx *c, int64_t n, const double *rel_pos,
                                           const double *a, const double *ecc, const double *pe_arg,
                                           const double *u, const double *dr, double *vel_out)
{
    CTX(c);
    VSB_CHECK(n >= 0 && (n == 0 || (rel_pos && a && ecc && pe_arg && u && dr && vel_out)),
              VSB_ERR_ARG, "vsb_host_kepler_to_velocity: bad argument");
    if (n == 0) return VSB_OK;
    const size_t N = (size_t)n;
    const in_arr ins[6] = {{rel_pos, 2 * N}, {a, N}, {ecc, N}, {pe_arg, N}, {u, N}, {dr, N}};
    double *d[6], *d_out;
    VSB_TRY(stage_inputs(c, ins, 6, 2 * N, d, &d_out));
    VSB_TRY(vsb_launch_k2v(c, n, d[0], d[1], d[2], d[3], d[4], d[5], d_out));
    return d2h(c, vel_out, d_out, 16 * N);
}

extern "C" int vsb_host_velocity_to_kepler(vsb_ctx *c, int64_t n, const double *rel_pos,
                                           const double *rel_vel, const double *u, int failsafe,
                                           double *out5)
{
    CTX(c);
    VSB_CHECK(n >= 0 && failsafe >= 0 && failsafe <= 2 &&
                  (n == 0 || (rel_pos && rel_vel && u && out5)),
              VSB_ERR_ARG, "vsb_host_velocity_to_kepler: bad argument");
    if (n == 0) return VSB_OK;
    const size_t N = (size_t)n;
    const in_arr ins[3] = {{rel_pos, 2 * N}, {rel_vel, 2 * N}, {u, N}};
    double *d[3], *d_out;
    VSB_TRY(stage_inputs(c, ins, 3, 5 * N, d, &d_out));
    VSB_TRY(vsb_launch_v2k(c, n, d[0], d[1], d[2], failsafe, d_out));
    return d2h(c, out5, d_out, 40 * N);
}

extern "C" int vsb_host_orb2xy(vsb_ctx *c, int64_t n, const double *a, const double *b,
                               const double *f, const double *ecc, const double *pea,
                               const double *ref_pos, const double *ea, double *out)
{
    CTX(c);
    VSB_CHECK(n >= 0 && (n == 0 || (a && b && f && ecc && pea && ref_pos && ea && out)),
              VSB_ERR_ARG, "vsb_host_orb2xy: bad argument");
    if (n == 0) return VSB_OK;
    const size_t N = (size_t)n;
    const in_arr ins[7] = {{a, N}, {b, N}, {f, N}, {ecc, N}, {pea, N}, {ref_pos, 2 * N}, {ea, N}};
    double *d[7], *d_out;
    VSB_TRY(stage_inputs(c, ins, 7, 2 * N, d, &d_out));
    VSB_TRY(vsb_launch_orb2xy(c, n, d[0], d[1], d[2], d[3], d[4], d[5], d[6], d_out));
    return d2h(c, out, d_out, 16 * N);
}

extern "C" int vsb_host_kepler_basic(vsb_ctx *c, int64_t n, const int64_t *parents,
                                     const double *mass, const double *pos, const double *vel,
                                     double gc, double coi_coef, double *focus, double *ecc_v,
                                     double *semi_major, double *semi_minor, double *periapsis_arg,
                                     double *coi)
{
    CTX(c);
    VSB_CHECK(n >= 0 && (n == 0 || (parents && mass && pos && vel && focus && ecc_v && semi_major &&
                                    semi_minor && periapsis_arg && coi)),
              VSB_ERR_ARG, "vsb_host_kepler_basic: bad argument");
    if (n == 0) return VSB_OK;
    for (int64_t i = 0; i < n; ++i)
        VSB_CHECK(parents[i] >= 0 && parents[i] < n, VSB_ERR_ARG,
                  "vsb_host_kepler_basic: parent of body %lld out of range", (long long)i);
    const size_t N = (size_t)n;
    const in_arr ins[4] = {{parents, N}, {mass, N}, {pos, 2 * N}, {vel, 2 * N}};
    double *d[4], *o;
    VSB_TRY(stage_inputs(c, ins, 4, 7 * al32(N), d, &o));
    const size_t A = al32(N);     // outputs: focus | semi_major | semi_minor | periapsis_arg | coi | ecc_v(2)
    VSB_TRY(vsb_launch_kepler_basic_arrays(c, n, (const int64_t *)d[0], d[1], d[2], d[3], gc, coi_coef,
                                           o, o + 5 * A, o + A, o + 2 * A, o + 3 * A, o + 4 * A));
    double *outs[5] = {focus, semi_major, semi_minor, periapsis_arg, coi};
    for (int k = 0; k < 5; ++k)
        VSB_CUDA(cudaMemcpyAsync(outs[k], o + k * A, 8 * N, cudaMemcpyDeviceToHost, c->stream));
    return d2h(c, ecc_v, o + 5 * A, 16 * N);
}

// ---------------------------------------------------------------- N4: body thermal passes
extern "C" int vsb_host_body_thermal(vsb_ctx *c, int64_t n, const double *consts12,
                                     const double *mass, const double *den, const double *rad,
                                     const double *atm_pres0, const double *atm_scale_h,
                                     const double *atm_den0, double *atm_h,
                                     const int64_t *base_color, double *surf_grav,
                                     double *luminosity, double *temp, double *rad_sc,
                                     int64_t *color, double *types, int32_t *stellar_class)
{
    CTX(c);
    VSB_CHECK(n >= 0 && consts12, VSB_ERR_ARG, "vsb_host_body_thermal: bad argument");
    if (n == 0) return VSB_OK;
    VSB_CHECK(mass && den && rad && atm_pres0 && atm_scale_h && atm_den0 && atm_h && base_color &&
                  surf_grav && luminosity && temp && rad_sc && color && types && stellar_class,
              VSB_ERR_ARG, "vsb_host_body_thermal: null array");
    // staging (8-byte words): in 6n | atm_h n | base_color 3n | out 4n | color 3n | types n | stellar n/2
    const size_t N = (size_t)n, na = (N + 31) / 32 * 32;
    VSB_TRY(vsb_buf_reserve(c->host_stage, 8 * (20 * na)));
    double *b = (double *)c->host_stage.p;
    double *d_in = b, *d_atm = b + 6 * na, *d_out = b + 10 * na, *d_types = b + 17 * na;
    long long *d_bc = (long long *)(b + 7 * na), *d_col = (long long *)(b + 14 * na);
    int *d_st = (int *)(b + 18 * na);
    const double *ins[6] = {mass, den, rad, atm_pres0, atm_scale_h, atm_den0};
    // the kernel expects the six inputs at stride n (not na): pack them contiguously
    for (int k = 0; k < 6; ++k)
        VSB_CUDA(cudaMemcpyAsync(d_in + k * N, ins[k], 8 * N, cudaMemcpyHostToDevice, c->stream));
    VSB_CUDA(cudaMemcpyAsync(d_atm, atm_h, 8 * N, cudaMemcpyHostToDevice, c->stream));
    VSB_CUDA(cudaMemcpyAsync(d_bc, base_color, 24 * N, cudaMemcpyHostToDevice, c->stream));
    VSB_TRY(vsb_launch_body_thermal(c, n, consts12, d_in, d_atm, d_bc, d_out, d_col, d_types, d_st));
    double *outs[4] = {surf_grav, luminosity, temp, rad_sc};
    for (int k = 0; k < 4; ++k)
        VSB_CUDA(cudaMemcpyAsync(outs[k], d_out + k * N, 8 * N, cudaMemcpyDeviceToHost, c->stream));
    VSB_CUDA(cudaMemcpyAsync(atm_h, d_atm, 8 * N, cudaMemcpyDeviceToHost, c->stream));
    VSB_CUDA(cudaMemcpyAsync(color, d_col, 24 * N, cudaMemcpyDeviceToHost, c->stream));
    VSB_CUDA(cudaMemcpyAsync(types, d_types, 8 * N, cudaMemcpyDeviceToHost, c->stream));
    VSB_CUDA(cudaMemcpyAsync(stellar_class, d_st, 4 * N, cudaMemcpyDeviceToHost, c->stream));
    VSB_CUDA(cudaStreamSynchronize(c->stream));
    return VSB_OK;
}

// ---------------------------------------------------------------- N3: culling / visible subset
// staging layout in host_stage: flags[n] | radius/size[n] | idx[n] | counts[ceil(n/256)] | total
struct vis_stage {
    unsigned char *flags;
    double *vals;
    int64_t *idx;
    int *counts;
    int64_t *total;
    unsigned char *ref_vis;
    double *ref_coi;
};

static int vis_reserve(vsb_ctx *c, int64_t n, int64_t nref, vis_stage *st)
{
    auto al = [](size_t b) { return (b + 255) / 256 * 256; };
    const size_t o_flags = 0, o_vals = al((size_t)n), o_idx = o_vals + al(8 * (size_t)n),
                 o_cnt = o_idx + al(8 * (size_t)n), o_tot = o_cnt + al(4 * (size_t)(n / 256 + 2)),
                 o_rv = o_tot + 256, o_rc = o_rv + al((size_t)nref),
                 total = o_rc + al(8 * (size_t)nref);
    VSB_TRY(vsb_buf_reserve(c->host_stage, total));
    char *b = (char *)c->host_stage.p;
    st->flags = (unsigned char *)(b + o_flags);
    st->vals = (double *)(b + o_vals);
    st->idx = (int64_t *)(b + o_idx);
    st->counts = (int *)(b + o_cnt);
    st->total = (int64_t *)(b + o_tot);
    st->ref_vis = (unsigned char *)(b + o_rv);
    st->ref_coi = (double *)(b + o_rc);
    return VSB_OK;
}

static int vis_finish(vsb_ctx *c, vis_stage &st, int64_t n, int64_t *idx_out, int64_t *count_out)
{
    VSB_TRY(vsb_launch_compact(c, st.flags, n, st.counts, st.total, st.idx));
    int64_t count = 0;
    if (idx_out && n + 1 <= (int64_t)k_flag_at && use_mapped(c)) {
        // small lists: the count and all n index slots in one round trip (the first `count` are valid)
        if (c->h_small_cap < 8 * (size_t)n) {
            void *h = realloc(c->h_small, 8 * (size_t)n);
            VSB_CHECK(h != nullptr, VSB_ERR_NOMEM, "compaction result: out of host memory");
            c->h_small = h;
            c->h_small_cap = 8 * (size_t)n;
        }
        const void *src[2] = {st.total, st.idx};
        const int words[2] = {1, (int)n};
        void *dst[2] = {&count, c->h_small};
        VSB_TRY(d2h_small(c, 2, src, words, dst, "compaction result"));
        *count_out = count;
        if (count > 0) memcpy(idx_out, c->h_small, 8 * (size_t)count);
        return VSB_OK;
    }
    VSB_TRY(d2h(c, &count, st.total, sizeof(int64_t)));
    *count_out = count;
    if (count > 0 && idx_out) VSB_TRY(d2h(c, idx_out, st.idx, 8 * (size_t)count));
    return VSB_OK;
}

extern "C" int vsb_host_culling(vsb_ctx *c, int64_t n, const double *coords, const double *radius,
                                const double *screen_bounds, double zoom, unsigned char *mask_out)
{
    CTX(c);
    VSB_CHECK(n >= 0 && (n == 0 || (coords && radius && screen_bounds && mask_out)) && zoom != 0.0,
              VSB_ERR_ARG, "vsb_host_culling: bad argument");
    if (n == 0) return VSB_OK;
    vis_stage st;
    VSB_TRY(vis_reserve(c, 3 * n, 0, &st));   // idx area doubles as the coords staging (16n bytes)
    double *d_pos = (double *)st.idx;
    VSB_CUDA(cudaMemcpyAsync(d_pos, coords, 16 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    VSB_CUDA(cudaMemcpyAsync(st.vals, radius, 8 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    VSB_TRY(vsb_launch_culling_flags(c, d_pos, st.vals, 0.0, n, screen_bounds, zoom, st.flags));
    return d2h(c, mask_out, st.flags, (size_t)n);
}

extern "C" int vsb_kepler_culling(vsb_ctx *c, int kind, const double *radius,
                                  const double *screen_bounds, double zoom, int64_t *idx_out,
                                  int64_t *count_out)
{
    CTX(c);
    vsb_kepler_state *k;
    VSB_TRY(kep_of(c, "vsb_kepler_culling", kind, &k));
    VSB_CHECK(k->n > 0, VSB_ERR_STATE, "vsb_kepler_culling: no state loaded for this kind");
    VSB_CHECK(screen_bounds && count_out && zoom != 0.0, VSB_ERR_ARG,
              "vsb_kepler_culling: bad argument");
    vis_stage st;
    VSB_TRY(vis_reserve(c, k->n, 0, &st));
    if (radius)
        VSB_CUDA(cudaMemcpyAsync(st.vals, radius, 8 * (size_t)k->n, cudaMemcpyHostToDevice,
                                 c->stream));
    VSB_TRY(vsb_launch_culling_flags(c, k->pos, radius ? st.vals : nullptr, 0.0, k->n,
                                     screen_bounds, zoom, st.flags));
    return vis_finish(c, st, k->n, idx_out, count_out);
}

extern "C" int vsb_kepler_culling_uniform(vsb_ctx *c, int kind, double radius,
                                          const double *screen_bounds, double zoom,
                                          int64_t *idx_out, int64_t *count_out)
{
    CTX(c);
    vsb_kepler_state *k;
    VSB_TRY(kep_of(c, "vsb_kepler_culling_uniform", kind, &k));
    VSB_CHECK(k->n > 0, VSB_ERR_STATE, "vsb_kepler_culling_uniform: no state loaded for this kind");
    VSB_CHECK(screen_bounds && count_out && zoom != 0.0, VSB_ERR_ARG,
              "vsb_kepler_culling_uniform: bad argument");
    vis_stage st;
    VSB_TRY(vis_reserve(c, k->n, 0, &st));
    VSB_TRY(vsb_launch_culling_flags(c, k->pos, nullptr, radius, k->n, screen_bounds, zoom,
                                     st.flags));
    return vis_finish(c, st, k->n, idx_out, count_out);
}

// predict_enter_coi_service on the resident vessel state: ascending indices of the vessels to
// re-predict (feed them to vsb_vessel_points with only_enter_coi)
extern "C" int vsb_vessel_enter_coi_service(vsb_ctx *c, int crossing_pending, int64_t *idx_out,
                                            int64_t *count_out)
{
    CTX(c);
    vsb_kepler_state *k;
    VSB_TRY(need_vessel_events(c, "vsb_vessel_enter_coi_service", &k));
    VSB_CHECK(count_out != nullptr, VSB_ERR_ARG, "vsb_vessel_enter_coi_service: null output");
    vis_stage st;
    VSB_TRY(vis_reserve(c, k->n, 0, &st));
    VSB_TRY(vsb_launch_enter_coi_service(c, k->n, k->dr, k->ev_prev_ea, k->ea, k->ev_enter,
                                         crossing_pending, st.flags));
    return vis_finish(c, st, k->n, idx_out, count_out);
}

extern "C" int vsb_kepler_visible_orbits(vsb_ctx *c, int kind, int64_t nref,
                                         const unsigned char *ref_visible, const double *ref_coi,
                                         const double *size, double zoom, double screen_x,
                                         int64_t *idx_out, int64_t *count_out)
{
    CTX(c);
    vsb_kepler_state *k;
    VSB_TRY(kep_of(c, "vsb_kepler_visible_orbits", kind, &k));
    VSB_CHECK(k->n > 0, VSB_ERR_STATE, "vsb_kepler_visible_orbits: no state loaded for this kind");
    VSB_CHECK(nref > 0 && ref_visible && ref_coi && size && count_out && zoom != 0.0, VSB_ERR_ARG,
              "vsb_kepler_visible_orbits: bad argument");
    vis_stage st;
    VSB_TRY(vis_reserve(c, k->n, nref, &st));
    VSB_CUDA(cudaMemcpyAsync(st.vals, size, 8 * (size_t)k->n, cudaMemcpyHostToDevice, c->stream));
    VSB_CUDA(cudaMemcpyAsync(st.ref_vis, ref_visible, (size_t)nref, cudaMemcpyHostToDevice,
                             c->stream));
    VSB_CUDA(cudaMemcpyAsync(st.ref_coi, ref_coi, 8 * (size_t)nref, cudaMemcpyHostToDevice,
                             c->stream));
    VSB_TRY(vsb_launch_orbit_flags(c, k->ref, st.vals, k->n, st.ref_vis, st.ref_coi, zoom, screen_x,
                                   st.flags));
    return vis_finish(c, st, k->n, idx_out, count_out);
}

extern "C" int vsb_kepler_curves_gather(vsb_ctx *c, int kind, const int64_t *idx, int64_t count,
                                        double *out)
{
    CTX(c);
    vsb_kepler_state *k;
    VSB_TRY(kep_of(c, "vsb_kepler_curves_gather", kind, &k));
    VSB_CHECK(k->n > 0 && k->curves.p, VSB_ERR_STATE,
              "vsb_kepler_curves_gather: no curves generated yet");
    VSB_CHECK(count >= 0 && (count == 0 || (idx && out)), VSB_ERR_ARG,
              "vsb_kepler_curves_gather: bad argument");
    if (count == 0) return VSB_OK;
    for (int64_t i = 0; i < count; ++i)
        VSB_CHECK(idx[i] >= 0 && idx[i] < k->n, VSB_ERR_ARG,
                  "vsb_kepler_curves_gather: index %lld out of range", (long long)idx[i]);
    const size_t bytes = 16 * (size_t)count * (size_t)k->P;
    VSB_TRY(vsb_buf_reserve(c->host_stage, bytes + 8 * (size_t)count + 256));
    double *d_out = (double *)c->host_stage.p;
    int64_t *d_idx = (int64_t *)((char *)c->host_stage.p + (bytes + 255) / 256 * 256);
    VSB_CUDA(cudaMemcpyAsync(d_idx, idx, 8 * (size_t)count, cudaMemcpyHostToDevice, c->stream));
    VSB_TRY(vsb_launch_curves_gather(c, (const double *)k->curves.p, d_idx, count, k->P, d_out));
    return d2h(c, out, d_out, bytes);
}

extern "C" int vsb_vessel_curve_segments(vsb_ctx *c, const int64_t *idx, int64_t count,
                                         double *light, double *dark, double *first_intersect,
                                         int32_t *intersect_type, double *intersect_time,
                                         double *select_range)
{
    CTX(c);
    vsb_kepler_state *k;
    VSB_TRY(need_vessel_events(c, "vsb_vessel_curve_segments", &k));
    VSB_CHECK(k->curves.p && k->ref_pos, VSB_ERR_STATE,
              "vsb_vessel_curve_segments: no moved curves yet (vsb_kepler_curves / vsb_kepler_tick)");
    VSB_CHECK(count >= 0 && (count == 0 || (idx && light && dark && first_intersect &&
                                            intersect_type && intersect_time && select_range)),
              VSB_ERR_ARG, "vsb_vessel_curve_segments: bad argument");
    if (count == 0) return VSB_OK;
    for (int64_t i = 0; i < count; ++i)
        VSB_CHECK(idx[i] >= 0 && idx[i] < k->n, VSB_ERR_ARG,
                  "vsb_vessel_curve_segments: index %lld out of range", (long long)idx[i]);
    // staging: light | dark | first | time | range | type | idx
    const size_t curve_bytes = 16 * (size_t)count * (size_t)k->P;
    auto al = [](size_t b) { return (b + 255) / 256 * 256; };
    const size_t o_dark = al(curve_bytes), o_first = o_dark + al(curve_bytes),
                 o_time = o_first + al(16 * (size_t)count), o_range = o_time + al(8 * (size_t)count),
                 o_type = o_range + al(16 * (size_t)count), o_idx = o_type + al(4 * (size_t)count),
                 total = o_idx + al(8 * (size_t)count);
    VSB_TRY(vsb_buf_reserve(c->host_stage, total));
    char *base = (char *)c->host_stage.p;
    int64_t *d_idx = (int64_t *)(base + o_idx);
    VSB_CUDA(cudaMemcpyAsync(d_idx, idx, 8 * (size_t)count, cudaMemcpyHostToDevice, c->stream));
    VSB_TRY(vsb_launch_curve_segments(c, *k, d_idx, count, (double *)base, (double *)(base + o_dark),
                                      (double *)(base + o_first), (int *)(base + o_type),
                                      (double *)(base + o_time), (double *)(base + o_range)));
    VSB_CUDA(cudaMemcpyAsync(light, base, curve_bytes, cudaMemcpyDeviceToHost, c->stream));
    VSB_CUDA(cudaMemcpyAsync(dark, base + o_dark, curve_bytes, cudaMemcpyDeviceToHost, c->stream));
    VSB_CUDA(cudaMemcpyAsync(first_intersect, base + o_first, 16 * (size_t)count,
                             cudaMemcpyDeviceToHost, c->stream));
    VSB_CUDA(cudaMemcpyAsync(intersect_time, base + o_time, 8 * (size_t)count,
                             cudaMemcpyDeviceToHost, c->stream));
    VSB_CUDA(cudaMemcpyAsync(select_range, base + o_range, 16 * (size_t)count,
                             cudaMemcpyDeviceToHost, c->stream));
    return d2h(c, intersect_type, base + o_type, 4 * (size_t)count);
}

extern "C" int vsb_kepler_free(vsb_ctx *c, int kind)
{
    CTX(c);
    vsb_kepler_state *k;
    VSB_TRY(kep_of(c, "vsb_kepler_free", kind, &k));
    VSB_CUDA(cudaStreamSynchronize(c->stream));
    free_kepler(*k);
    if (kind == VSB_KIND_BODY && c->kep[VSB_KIND_VESSEL].ref_pos_borrowed) {
        c->kep[VSB_KIND_VESSEL].ref_pos = nullptr;
        c->kep[VSB_KIND_VESSEL].ref_pos_borrowed = 0;
        c->kep[VSB_KIND_VESSEL].nref = 0;
    }
    return VSB_OK;
}
