// nagp_api.cu — the C ABI declared in include/nagp.h: context, host/device buffer staging and
// the entry points that replace NowcastAutoGP's calls into AutoGP (file:line per function in the
// header). No torch types, no exceptions across the boundary.
#include "../../include/nagp.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "nagp_kernels.cuh"
#include "nagp_tree.cuh"

using namespace nagp;

// One NVTX range per C-ABI entry point (SURVEY §5 tracing row): header-only NVTX v3, a no-op unless a profiler is attached.
namespace {

// new tile rows from which nagp_factor_append uses the factorisation kernel. Measured at 256 x n = 2048, by rows / streamed:
// one tile row 5.2 / 1.1 ms, two 5.1 / 5.6, three 5.9 / 7.4, four 5.7 / 9.0, eight 6.3 / 17 (1024 x n = 512: two 1.4 / 1.7, four 1.7 / 2.6)
constexpr int kAppendByRowsFrom = 2;
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};
}  // namespace
#define NAGP_RANGE(name) NvtxRange nvtx_range__(name)

namespace {

struct PendingOut { void *host; const void *dev; size_t bytes; };
struct Chunk { char *base; size_t cap; };

std::string g_init_error;

}  // namespace

struct nagp_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t copy_stream = nullptr;     // host->device copies that can run under a kernel of `stream`
    cudaEvent_t copy_done = nullptr, fork = nullptr;
    double jitter = 1e-5;
    int variant = 0;
    int last_kernel = 0;
    int64_t launches = 0;
    int smem_optin = 0, smem_per_sm = 0, num_sms = 0;
    std::string err;
    std::vector<Chunk> chunks;
    size_t chunk_off = 0;          // bump offset in chunks.back()
    std::vector<PendingOut> outs;
    std::vector<nagp::TreeProgram> compiled;   // host-compiled programs of the current call (empty: compile on device)
};

struct nagp_factor {
    int device;
    int64_t P;
    int n, k, h;
    double ya, yb;
    double *proj = nullptr, *Ltail = nullptr, *L33 = nullptr, *logw0 = nullptr, *logml_n = nullptr;
    // appendable store (nagp_factor_store_large): the whole factor stays in HBM
    bool appendable = false;
    int64_t cap = 0;
    int ntab_cap = 0, ncp_cap = 0, nth_cap = 1;
    double step = 0.0, jitter = 0.0;
    size_t L_stride = 0, W_stride = 0;
    double *Lbig = nullptr, *Wbig = nullptr, *d_theta = nullptr, *d_noise = nullptr, *d_t = nullptr, *d_y = nullptr;
    int32_t *d_g = nullptr;
    uint8_t *d_prog = nullptr;
    int64_t *d_prog_off = nullptr, *d_theta_off = nullptr;
    std::vector<int32_t> g_host;      // lag-grid indices of the stored points (empty: pairwise times)
};

namespace {

int32_t fail(nagp_ctx *ctx, int32_t code, const std::string &msg)
{
    if (ctx) ctx->err = msg; else g_init_error = msg;
    return code;
}

#define NAGP_CUDA(ctx, expr)                                                                   \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return fail(ctx, NAGP_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)

bool on_device(const void *p)
{
    if (!p) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// ---- grow-only device arena: bump allocation per call, consolidated at the start of the next ----
int32_t arena_reset(nagp_ctx *ctx)
{
    ctx->outs.clear();
    if (ctx->chunks.size() > 1) {
        NAGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        size_t total = 0;
        for (auto &c : ctx->chunks) { total += c.cap; cudaFree(c.base); }
        ctx->chunks.clear();
        char *base = nullptr;
        NAGP_CUDA(ctx, cudaMalloc(&base, total));
        ctx->chunks.push_back({base, total});
    }
    ctx->chunk_off = 0;
    return NAGP_OK;
}

void *arena_alloc(nagp_ctx *ctx, size_t bytes)
{
    bytes = (bytes + 255) & ~size_t(255);
    if (bytes == 0) bytes = 256;
    if (ctx->chunks.empty() || ctx->chunk_off + bytes > ctx->chunks.back().cap) {
        size_t cap = std::max<size_t>(bytes, ctx->chunks.empty() ? (size_t(8) << 20) : 2 * ctx->chunks.back().cap);
        char *base = nullptr;
        if (cudaMalloc(&base, cap) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        ctx->chunks.push_back({base, cap});
        ctx->chunk_off = 0;
    }
    void *p = ctx->chunks.back().base + ctx->chunk_off;
    ctx->chunk_off += bytes;
    return p;
}

template <class T>
int32_t stage_in(nagp_ctx *ctx, const T *p, size_t count, const T **dev, cudaStream_t on = nullptr)
{
    if (!p || count == 0) { *dev = nullptr; return NAGP_OK; }
    if (on_device(p)) { *dev = p; return NAGP_OK; }
    T *d = static_cast<T *>(arena_alloc(ctx, count * sizeof(T)));
    if (!d) return fail(ctx, NAGP_E_CUDA, "device workspace allocation failed");
    NAGP_CUDA(ctx, cudaMemcpyAsync(d, p, count * sizeof(T), cudaMemcpyHostToDevice, on ? on : ctx->stream));
    *dev = d;
    return NAGP_OK;
}

template <class T>
int32_t stage_out(nagp_ctx *ctx, T *p, size_t count, T **dev)
{
    if (!p || count == 0) { *dev = nullptr; return NAGP_OK; }
    if (on_device(p)) { *dev = p; return NAGP_OK; }
    T *d = static_cast<T *>(arena_alloc(ctx, count * sizeof(T)));
    if (!d) return fail(ctx, NAGP_E_CUDA, "device workspace allocation failed");
    ctx->outs.push_back({p, d, count * sizeof(T)});
    *dev = d;
    return NAGP_OK;
}

template <class T>
int32_t scratch(nagp_ctx *ctx, size_t count, T **dev)
{
    *dev = static_cast<T *>(arena_alloc(ctx, count * sizeof(T)));
    if (!*dev) return fail(ctx, NAGP_E_CUDA, "device workspace allocation failed");
    return NAGP_OK;
}

// Copy host outputs back; blocks iff there are any.
int32_t finish(nagp_ctx *ctx)
{
    for (auto &o : ctx->outs)
        NAGP_CUDA(ctx, cudaMemcpyAsync(o.host, o.dev, o.bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (!ctx->outs.empty()) NAGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->outs.clear();
    return NAGP_OK;
}

#define NAGP_TRY(expr)                 \
    do {                               \
        int32_t rc__ = (expr);         \
        if (rc__ != NAGP_OK) return rc__; \
    } while (0)

// Worst positive info over a host-visible info array (after finish()).
int32_t worst_info(const int32_t *info, int64_t count)
{
    int32_t neg = 0, pos = 0;
    for (int64_t i = 0; i < count; ++i) {
        if (info[i] < 0) neg = info[i];
        else if (info[i] > 0 && pos == 0) pos = info[i];
    }
    return neg ? neg : pos;
}

struct GridInfo { int G; };

// Lag-table extent. g may live on either side; q is small, so a device-resident g costs one tiny
// synchronous copy.
int32_t grid_extent(nagp_ctx *ctx, const int32_t *g, int64_t q, int *G)
{
    *G = 0;
    if (!g) return NAGP_OK;
    std::vector<int32_t> host(q);
    if (on_device(g)) {
        NAGP_CUDA(ctx, cudaMemcpyAsync(host.data(), g, q * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        NAGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    } else {
        std::memcpy(host.data(), g, q * sizeof(int32_t));
    }
    auto mm = std::minmax_element(host.begin(), host.end());
    int64_t ext = (int64_t)*mm.second - (int64_t)*mm.first + 1;
    if (ext > (1 << 20)) return fail(ctx, NAGP_E_ARG, "lag grid extent too large");
    *G = (int)ext;
    return NAGP_OK;
}

// Table capacities: exact maxima when the programs are host-visible, defaults otherwise; shrunk
// until the launch fits the opt-in shared-memory limit. A smaller capacity only means more
// sub-trees are evaluated directly per entry — never a different result.
int32_t plan_tables(nagp_ctx *ctx, int64_t P, const uint8_t *prog, const int64_t *prog_off,
                    const int64_t *theta_off, int q, int G, int *ntab_cap, int *ncp_cap)
{
    int ntab = G > 0 ? 4 : 0, ncp = 4;
    if (prog && !on_device(prog) && !on_device(prog_off) && !on_device(theta_off)) {
        ntab = 0; ncp = 0;
        TreeProgram tp;
        for (int64_t p = 0; p < P; ++p) {
            int64_t len = prog_off[p + 1] - prog_off[p];
            int64_t nth = theta_off[p + 1] - theta_off[p];
            if (len <= 0 || len > MAX_PROG || nth > MAX_THETA)
                return fail(ctx, NAGP_E_PROGRAM, "kernel program length outside 1..64");
            tree_compile(tp, prog + prog_off[p], (int)len, (int)nth, G > 0 ? MAX_TABLES : 0, MAX_CPTAB);
            if (tp.error) return fail(ctx, NAGP_E_PROGRAM, "malformed kernel program");
            ntab = std::max(ntab, tp.ntab);
            ncp = std::max(ncp, tp.ncp);
        }
    }
    *ntab_cap = ntab; *ncp_cap = ncp;
    ctx->compiled.clear();
    if (prog && !on_device(prog) && !on_device(prog_off) && !on_device(theta_off)) {
        // the programs as the device would compile them with the final capacities: uploaded once per call
        ctx->compiled.resize((size_t)P);
        for (int64_t p = 0; p < P; ++p)
            tree_compile(ctx->compiled[p], prog + prog_off[p], (int)(prog_off[p + 1] - prog_off[p]),
                         (int)(theta_off[p + 1] - theta_off[p]), G > 0 ? ntab : 0, ncp);
    }
    return NAGP_OK;
}

// The column kernel keeps its tables in shared memory: shrink the capacities until it fits.
int32_t fit_tables_v1(nagp_ctx *ctx, int q, int G, int *ntab_cap, int *ncp_cap)
{
    int ntab = *ntab_cap, ncp = *ncp_cap;
    while (fused_smem_bytes_v1(q, G, ntab, ncp) > (size_t)ctx->smem_optin && (ntab > 0 || ncp > 0)) {
        if (ntab * (size_t)std::max(G, 1) >= ncp * (size_t)q && ntab > 0) --ntab;
        else if (ncp > 0) --ncp;
        else --ntab;
    }
    if (fused_smem_bytes_v1(q, G, ntab, ncp) > (size_t)ctx->smem_optin)
        return fail(ctx, NAGP_E_SIZE, "problem too large for the shared-memory resident path");
    *ntab_cap = ntab; *ncp_cap = ncp;
    return NAGP_OK;
}

int32_t check_dims(nagp_ctx *ctx, int64_t n, int64_t k, int64_t h)
{
    if (n < 0 || k < 0 || h < 0 || n + k + h <= 0) return fail(ctx, NAGP_E_ARG, "bad n/k/h");
    if (n + k + h > large_max_q()) return fail(ctx, NAGP_E_SIZE, "n+k+h > 4096 not supported");
    return NAGP_OK;
}

// Launch the fused Gram+Cholesky+solve kernel: the tile (DMMA) kernel unless variant 1 is forced
// or the problem does not fit it. `theta_off` (host) gives the per-program theta counts.
int32_t run_fused(nagp_ctx *ctx, FusedArgs a, const int64_t *theta_off_host)
{
    const int q = a.n + a.k + a.h;
    // Slot kernel (three matrices in flight per SM, nagp_fused_v3.cu): selected with variant 3 wherever it applies — it needs
    // host-compiled programs whose stationary leaves and changepoints are all tabulated, and no kept factor. It is
    // parity-green but at 5.6 ms per 32 000 instances still behind the tile kernel (5.1 ms), so auto (variant 0) keeps the
    // tile kernel.
    if (ctx->variant == 3 && q <= fused_v3_max_q() && !a.Lkeep && (int64_t)ctx->compiled.size() == a.P) {
        bool table_only = true;
        int64_t nth = 1;
        for (int64_t p = 0; p < a.P && table_only; ++p) {
            const TreeProgram &tp = ctx->compiled[p];
            nth = std::max(nth, theta_off_host[p + 1] - theta_off_host[p]);
            for (int c = 0; c < tp.clen; ++c) {
                const int op = tp.cop[c];
                if (op == OP_SQEXP || op == OP_GAMMAEXP || op == OP_PERIODIC || op == OP_CHANGEPOINT) table_only = false;
            }
        }
        if (table_only) {
            V3Plan pl = plan_fused_v3(a.n, a.k, a.h, a.proj != nullptr || a.Ltail != nullptr, a.G,
                                      (int)std::min<int64_t>(nth, MAX_THETA), a.ntab_cap, a.ncp_cap, ctx->smem_optin);
            if (pl.ok) {
                NAGP_TRY(stage_in(ctx, ctx->compiled.data(), ctx->compiled.size(), &a.compiled));
                const int grid = fused_v3_grid(a.B, ctx->num_sms);
                if (getenv("NAGP_DEBUG"))
                    fprintf(stderr, "[nagp] slot kernel: q=%d nt=%d pool=%d slots lead=%d smem=%zu B grid=%d\n", q, pl.nt, pl.ns,
                            pl.lead, pl.smem_bytes, grid);
                unsigned long long *counter = nullptr;
                NAGP_TRY(scratch(ctx, 1, &counter));
                NAGP_CUDA(ctx, launch_fused_v3(a, pl, counter, grid, ctx->stream));
                ctx->launches += 1;
                ctx->last_kernel = 3;
                return NAGP_OK;
            }
        }
    }
    if (ctx->variant != 1 && q <= fused_v2_max_q() && ctx->variant != 4) {
        int64_t nth = 1;
        for (int64_t p = 0; p < a.P; ++p) nth = std::max(nth, theta_off_host[p + 1] - theta_off_host[p]);
        V2Plan pl = plan_fused_v2(q, a.G, (int)std::min<int64_t>(nth, MAX_THETA), a.ntab_cap, a.ncp_cap,
                                  ctx->smem_optin, ctx->smem_per_sm);
        if (pl.ok) {
            if ((int64_t)ctx->compiled.size() == a.P)
                NAGP_TRY(stage_in(ctx, ctx->compiled.data(), ctx->compiled.size(), &a.compiled));
            int grid = fused_v2_grid(pl, a.B, ctx->num_sms);
            if (getenv("NAGP_DEBUG"))
                fprintf(stderr, "[nagp] tile kernel: q=%d G=%d caps(tab=%d,cp=%d,theta=%d) smem=%zu B aux_in_smem(th,gg,tt,sig,tab)=%d%d%d%d%d scratch/CTA=%d grid=%d\n",
                        q, a.G, a.ntab_cap, a.ncp_cap, (int)nth, pl.smem_bytes, pl.aux_smem[0], pl.aux_smem[1],
                        pl.aux_smem[2], pl.aux_smem[3], pl.aux_smem[4], pl.scratch_stride, grid);
            char *scr = nullptr;
            if (pl.scratch_stride) NAGP_TRY(scratch(ctx, (size_t)grid * pl.scratch_stride, &scr));
            unsigned long long *counter = nullptr;
            NAGP_TRY(scratch(ctx, 1, &counter));
            NAGP_CUDA(ctx, launch_fused_v2(a, pl, scr, counter, grid, ctx->stream));
            ctx->launches += 1;
            ctx->last_kernel = 2;
            return NAGP_OK;
        }
    }
    if (ctx->variant != 1 || q > 234) {
        // beyond shared memory: blocked factorisation with the factor in HBM (per-CTA workspace)
        int64_t nth = 1;
        for (int64_t p = 0; p < a.P; ++p) nth = std::max(nth, theta_off_host[p + 1] - theta_off_host[p]);
        LargePlan pl = plan_large(q, q, a.G, (int)std::min<int64_t>(nth, MAX_THETA), a.ntab_cap, a.ncp_cap,
                                  ctx->smem_per_sm, false);
        if (!pl.ok) return fail(ctx, NAGP_E_SIZE, "problem too large");
        const int grid = large_grid(pl, a.B, ctx->num_sms, false);
        if (getenv("NAGP_DEBUG"))
            fprintf(stderr, "[nagp] large kernel: q=%d ntp=%d G=%d smem=%zu B scratch/CTA=%d L/CTA=%zu B grid=%d\n",
                    q, pl.ntp, a.G, pl.smem_bytes, pl.scratch_stride, pl.L_stride * 8, grid);
        char *scr = nullptr;
        if (pl.scratch_stride) NAGP_TRY(scratch(ctx, (size_t)grid * pl.scratch_stride, &scr));
        double *Lws = nullptr;
        NAGP_TRY(scratch(ctx, (size_t)grid * pl.L_stride, &Lws));
        unsigned long long *counter = nullptr;
        NAGP_TRY(scratch(ctx, 1, &counter));
        NAGP_CUDA(ctx, launch_chol_large(a, pl, scr, Lws, 0, nullptr, counter, grid, ctx->stream));
        ctx->launches += 1;
        ctx->last_kernel = 4;
        return NAGP_OK;
    }
    NAGP_TRY(fit_tables_v1(ctx, q, a.G, &a.ntab_cap, &a.ncp_cap));
    NAGP_CUDA(ctx, launch_fused_v1(a, ctx->stream));
    ctx->launches += 1;
    ctx->last_kernel = 1;
    return NAGP_OK;
}

}  // namespace

extern "C" {

int32_t nagp_version(void) { return 100; }

int32_t nagp_init(int32_t device, nagp_ctx **out)
{
    if (!out) return fail(nullptr, NAGP_E_ARG, "out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, NAGP_E_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(nullptr, NAGP_E_ARG, "device index out of range");
    nagp_ctx *ctx = new (std::nothrow) nagp_ctx();
    if (!ctx) return fail(nullptr, NAGP_E_ARG, "out of host memory");
    ctx->device = device;
    if ((e = cudaSetDevice(device)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ctx->copy_done, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&ctx->fork, cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaDeviceGetAttribute(&ctx->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device)) != cudaSuccess ||
        (e = cudaDeviceGetAttribute(&ctx->smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device)) != cudaSuccess ||
        (e = cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, NAGP_E_CUDA, std::string("context setup: ") + cudaGetErrorString(e));
    }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return NAGP_OK;
}

void nagp_destroy(nagp_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &c : ctx->chunks) cudaFree(c.base);
    if (ctx->copy_done) cudaEventDestroy(ctx->copy_done);
    if (ctx->fork) cudaEventDestroy(ctx->fork);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char *nagp_last_error(const nagp_ctx *ctx) { return ctx ? ctx->err.c_str() : g_init_error.c_str(); }

int32_t nagp_set_stream(nagp_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return NAGP_E_ARG;
    NAGP_CUDA(ctx, cudaSetDevice(ctx->device));
    NAGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return NAGP_OK;
}

int32_t nagp_set_jitter(nagp_ctx *ctx, double jitter)
{
    if (!ctx || !(jitter >= 0.0)) return NAGP_E_ARG;
    ctx->jitter = jitter;
    return NAGP_OK;
}

int64_t nagp_launch_count(const nagp_ctx *ctx) { return ctx ? ctx->launches : 0; }

int32_t nagp_last_kernel(const nagp_ctx *ctx) { return ctx ? ctx->last_kernel : 0; }

int32_t nagp_set_variant(nagp_ctx *ctx, int32_t variant)
{
    if (!ctx || variant < 0 || variant > 4) return NAGP_E_ARG;
    ctx->variant = variant;
    return NAGP_OK;
}

int32_t nagp_logml_batch(nagp_ctx *ctx, int64_t B, const uint8_t *prog, const int64_t *prog_off,
                         const double *theta, const int64_t *theta_off, const double *noise,
                         int64_t n, const double *t, const int32_t *g, double step,
                         const double *y, int64_t y_stride, double *logml, int32_t *info)
{
    NAGP_RANGE("nagp_logml_batch");
    if (!ctx) return NAGP_E_ARG;
    if (B <= 0 || !prog || !prog_off || !theta || !theta_off || !noise || !t || !y || !logml || !info)
        return fail(ctx, NAGP_E_ARG, "nagp_logml_batch: null or empty argument");
    NAGP_TRY(check_dims(ctx, n, 0, 0));
    NAGP_CUDA(ctx, cudaSetDevice(ctx->device));
    NAGP_TRY(arena_reset(ctx));
    FusedArgs a{};
    NAGP_TRY(grid_extent(ctx, g, n, &a.G));
    NAGP_TRY(plan_tables(ctx, B, prog, prog_off, theta_off, (int)n, a.G, &a.ntab_cap, &a.ncp_cap));
    int64_t nprog, ntheta;
    if (on_device(prog_off) || on_device(theta_off))
        return fail(ctx, NAGP_E_ARG, "prog_off/theta_off must be host arrays");
    nprog = prog_off[B]; ntheta = theta_off[B];
    a.B = B; a.P = B;
    NAGP_TRY(stage_in(ctx, prog, (size_t)nprog, &a.prog));
    NAGP_TRY(stage_in(ctx, prog_off, (size_t)B + 1, &a.prog_off));
    NAGP_TRY(stage_in(ctx, theta, (size_t)ntheta, &a.theta));
    NAGP_TRY(stage_in(ctx, theta_off, (size_t)B + 1, &a.theta_off));
    NAGP_TRY(stage_in(ctx, noise, (size_t)B, &a.noise));
    a.jitter = ctx->jitter; a.noise_pred = -1.0;
    a.n = (int)n; a.k = 0; a.h = 0;
    NAGP_TRY(stage_in(ctx, t, (size_t)n, &a.t));
    NAGP_TRY(stage_in(ctx, g, (size_t)n, &a.g));
    a.step = step;
    NAGP_TRY(stage_in(ctx, y, (size_t)(y_stride ? (B - 1) * y_stride + n : n), &a.y1));
    a.y1_stride = y_stride;
    a.ya = 1.0; a.yb = 0.0;
    NAGP_TRY(stage_out(ctx, logml, (size_t)B, &a.logml_n));
    NAGP_TRY(stage_out(ctx, info, (size_t)B, &a.info));
    const bool host_info = !ctx->outs.empty() && !on_device(info);
    NAGP_TRY(run_fused(ctx, a, theta_off));
    NAGP_TRY(finish(ctx));
    return host_info ? worst_info(info, B) : NAGP_OK;
}

int32_t nagp_forecast_instances(nagp_ctx *ctx, int64_t K, int64_t P, const uint8_t *prog,
                                const int64_t *prog_off, const double *theta, const int64_t *theta_off,
                                int64_t theta_stride_k, const double *noise, int64_t noise_stride_k,
                                double noise_pred, int64_t n, int64_t k, int64_t h, const double *t,
                                const int32_t *g, double step, const double *y1, const double *y2,
                                double ya, double yb, const double *logw0, double *logw, double *mu,
                                double *L, int32_t *info, double *logml_n, double *logml_m)
{
    NAGP_RANGE("nagp_forecast_instances");
    if (!ctx) return NAGP_E_ARG;
    if (K <= 0 || P <= 0 || !prog || !prog_off || !theta || !theta_off || !noise || !t || !y1 || !info ||
        (k > 0 && !y2) || ya == 0.0)
        return fail(ctx, NAGP_E_ARG, "nagp_forecast_instances: null or empty argument");
    NAGP_TRY(check_dims(ctx, n, k, h));
    if (on_device(prog_off) || on_device(theta_off))
        return fail(ctx, NAGP_E_ARG, "prog_off/theta_off must be host arrays");
    NAGP_CUDA(ctx, cudaSetDevice(ctx->device));
    NAGP_TRY(arena_reset(ctx));
    const int64_t q = n + k + h, B = K * P;
    FusedArgs a{};
    NAGP_TRY(grid_extent(ctx, g, q, &a.G));
    NAGP_TRY(plan_tables(ctx, P, prog, prog_off, theta_off, (int)q, a.G, &a.ntab_cap, &a.ncp_cap));
    const int64_t nprog = prog_off[P], ntheta = theta_off[P];
    a.B = B; a.P = P;
    NAGP_TRY(stage_in(ctx, prog, (size_t)nprog, &a.prog));
    NAGP_TRY(stage_in(ctx, prog_off, (size_t)P + 1, &a.prog_off));
    NAGP_TRY(stage_in(ctx, theta, (size_t)(theta_stride_k ? (K - 1) * theta_stride_k + ntheta : ntheta), &a.theta));
    NAGP_TRY(stage_in(ctx, theta_off, (size_t)P + 1, &a.theta_off));
    a.theta_stride_k = theta_stride_k;
    NAGP_TRY(stage_in(ctx, noise, (size_t)(noise_stride_k ? (K - 1) * noise_stride_k + P : P), &a.noise));
    a.noise_stride_k = noise_stride_k;
    a.jitter = ctx->jitter; a.noise_pred = noise_pred;
    a.n = (int)n; a.k = (int)k; a.h = (int)h;
    NAGP_TRY(stage_in(ctx, t, (size_t)q, &a.t));
    NAGP_TRY(stage_in(ctx, g, (size_t)q, &a.g));
    a.step = step;
    NAGP_TRY(stage_in(ctx, y1, (size_t)n, &a.y1));
    NAGP_TRY(stage_in(ctx, y2, (size_t)(K * k), &a.y2));
    a.ya = ya; a.yb = yb;
    NAGP_TRY(stage_in(ctx, logw0, (size_t)P, &a.logw0));
    NAGP_TRY(stage_out(ctx, logw, (size_t)B, &a.logw));
    NAGP_TRY(stage_out(ctx, mu, (size_t)(B * h), &a.mu));
    NAGP_TRY(stage_out(ctx, L, (size_t)(B * h * h), &a.L33));
    NAGP_TRY(stage_out(ctx, info, (size_t)B, &a.info));
    NAGP_TRY(stage_out(ctx, logml_n, (size_t)B, &a.logml_n));
    NAGP_TRY(stage_out(ctx, logml_m, (size_t)B, &a.logml_m));
    const bool host_info = !on_device(info);
    NAGP_TRY(run_fused(ctx, a, theta_off));
    NAGP_TRY(finish(ctx));
    return host_info ? worst_info(info, B) : NAGP_OK;
}

int32_t nagp_factor_store(nagp_ctx *ctx, int64_t P, const uint8_t *prog, const int64_t *prog_off,
                          const double *theta, const int64_t *theta_off, const double *noise,
                          double noise_pred, int64_t n, int64_t k, int64_t h, const double *t,
                          const int32_t *g, double step, const double *y1, double ya, double yb,
                          const double *logw0, nagp_factor **out, double *logml_n, int32_t *info)
{
    NAGP_RANGE("nagp_factor_store");
    if (!ctx) return NAGP_E_ARG;
    if (!out || P <= 0 || !prog || !prog_off || !theta || !theta_off || !noise || !t || !y1 || !info || ya == 0.0)
        return fail(ctx, NAGP_E_ARG, "nagp_factor_store: null or empty argument");
    *out = nullptr;
    NAGP_TRY(check_dims(ctx, n, k, h));
    if (on_device(prog_off) || on_device(theta_off))
        return fail(ctx, NAGP_E_ARG, "prog_off/theta_off must be host arrays");
    NAGP_CUDA(ctx, cudaSetDevice(ctx->device));
    NAGP_TRY(arena_reset(ctx));
    const int64_t q = n + k + h, kh = k + h;
    nagp_factor *f = new (std::nothrow) nagp_factor();
    if (!f) return fail(ctx, NAGP_E_ARG, "out of host memory");
    f->device = ctx->device; f->P = P; f->n = (int)n; f->k = (int)k; f->h = (int)h; f->ya = ya; f->yb = yb;
    auto dalloc = [&](double **p, size_t cnt) { return cudaMalloc(p, std::max<size_t>(cnt, 1) * sizeof(double)); };
    if (dalloc(&f->proj, P * kh) != cudaSuccess || dalloc(&f->Ltail, P * kh * kh) != cudaSuccess ||
        dalloc(&f->L33, P * h * h) != cudaSuccess || dalloc(&f->logw0, P) != cudaSuccess ||
        dalloc(&f->logml_n, P) != cudaSuccess) {
        nagp_factor_free(f);
        return fail(ctx, NAGP_E_CUDA, "factor allocation failed");
    }
    FusedArgs a{};
    int32_t rc;
    auto bail = [&](int32_t code) { nagp_factor_free(f); return code; };
    if ((rc = grid_extent(ctx, g, q, &a.G)) != NAGP_OK) return bail(rc);
    if ((rc = plan_tables(ctx, P, prog, prog_off, theta_off, (int)q, a.G, &a.ntab_cap, &a.ncp_cap)) != NAGP_OK) return bail(rc);
    a.B = P; a.P = P;
    if ((rc = stage_in(ctx, prog, (size_t)prog_off[P], &a.prog)) != NAGP_OK) return bail(rc);
    if ((rc = stage_in(ctx, prog_off, (size_t)P + 1, &a.prog_off)) != NAGP_OK) return bail(rc);
    if ((rc = stage_in(ctx, theta, (size_t)theta_off[P], &a.theta)) != NAGP_OK) return bail(rc);
    if ((rc = stage_in(ctx, theta_off, (size_t)P + 1, &a.theta_off)) != NAGP_OK) return bail(rc);
    if ((rc = stage_in(ctx, noise, (size_t)P, &a.noise)) != NAGP_OK) return bail(rc);
    a.jitter = ctx->jitter; a.noise_pred = noise_pred;
    a.n = (int)n; a.k = (int)k; a.h = (int)h;
    if ((rc = stage_in(ctx, t, (size_t)q, &a.t)) != NAGP_OK) return bail(rc);
    if ((rc = stage_in(ctx, g, (size_t)q, &a.g)) != NAGP_OK) return bail(rc);
    a.step = step;
    if ((rc = stage_in(ctx, y1, (size_t)n, &a.y1)) != NAGP_OK) return bail(rc);
    a.y2 = nullptr; a.ya = ya; a.yb = yb;
    a.logml_n = f->logml_n; a.proj = f->proj; a.Ltail = f->Ltail; a.L33 = f->L33;
    if ((rc = stage_out(ctx, info, (size_t)P, &a.info)) != NAGP_OK) return bail(rc);
    if (logw0) {
        cudaError_t e = cudaMemcpyAsync(f->logw0, logw0, P * sizeof(double), cudaMemcpyDefault, ctx->stream);
        if (e != cudaSuccess) return bail(fail(ctx, NAGP_E_CUDA, cudaGetErrorString(e)));
    } else {
        cudaMemsetAsync(f->logw0, 0, P * sizeof(double), ctx->stream);
    }
    if ((rc = run_fused(ctx, a, theta_off)) != NAGP_OK) return bail(rc);
    if (logml_n) {
        cudaError_t e = cudaMemcpyAsync(logml_n, f->logml_n, P * sizeof(double), cudaMemcpyDefault, ctx->stream);
        if (e != cudaSuccess) return bail(fail(ctx, NAGP_E_CUDA, cudaGetErrorString(e)));
    }
    if ((rc = finish(ctx)) != NAGP_OK) return bail(rc);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return bail(fail(ctx, NAGP_E_CUDA, cudaGetErrorString(e)));
    *out = f;
    return on_device(info) ? NAGP_OK : worst_info(info, P);
}

void nagp_factor_free(nagp_factor *f)
{
    if (!f) return;
    cudaSetDevice(f->device);
    cudaFree(f->proj); cudaFree(f->Ltail); cudaFree(f->L33); cudaFree(f->logw0); cudaFree(f->logml_n);
    cudaFree(f->Lbig); cudaFree(f->Wbig); cudaFree(f->d_theta); cudaFree(f->d_noise); cudaFree(f->d_t);
    cudaFree(f->d_y); cudaFree(f->d_g); cudaFree(f->d_prog); cudaFree(f->d_prog_off); cudaFree(f->d_theta_off);
    delete f;
}


// ---- (f1) gradient of the log marginal likelihood: the HMC primitive --------------------------------------
}  // extern "C"

namespace {

// Everything one "logML + gradient" evaluation launches, on device buffers that stay valid for the call:
// the tile kernel (keeping the factor) followed by the gradient kernel. Built once, enqueued as often as needed
// (once by nagp_logml_grad, once per leapfrog stage by nagp_hmc — also inside a CUDA-graph capture).
struct GradJob {
    FusedArgs a{};
    V2Plan vpl{};
    int vgrid = 0;
    char *vscr = nullptr;
    unsigned long long *vcounter = nullptr;
    GradArgs ga{};
    GradTilePlan gpl{};
    int ggrid = 0;
    char *gscr = nullptr;
    unsigned long long *gcounter = nullptr;
    size_t col_smem = 0;       // column-kernel fallback
};

// theta/noise/logml/grad_theta/grad_noise/info may be host or device; theta_dev etc. come back as device pointers.
int32_t build_grad_job(nagp_ctx *ctx, int64_t K, int64_t P, const uint8_t *prog, const int64_t *prog_off,
                       const double *theta, const int64_t *theta_off, int64_t theta_stride_k, const double *noise,
                       int64_t noise_stride_k, int64_t n, int64_t k, const double *t, const int32_t *g, double step,
                       const double *y1, int64_t y1_stride, const double *y2, double *logml, double *grad_theta,
                       double *grad_noise, int32_t *info, GradJob *job)
{
    const int64_t m = n + k, B = K * P, ntheta = theta_off[P];
    const int nt = (int)((m + 7) / 8);
    FusedArgs &a = job->a;
    NAGP_TRY(grid_extent(ctx, g, m, &a.G));
    NAGP_TRY(plan_tables(ctx, P, prog, prog_off, theta_off, (int)m, a.G, &a.ntab_cap, &a.ncp_cap));
    a.B = B; a.P = P;
    NAGP_TRY(stage_in(ctx, prog, (size_t)prog_off[P], &a.prog));
    NAGP_TRY(stage_in(ctx, prog_off, (size_t)P + 1, &a.prog_off));
    NAGP_TRY(stage_in(ctx, theta, (size_t)(theta_stride_k ? (K - 1) * theta_stride_k + ntheta : ntheta), &a.theta));
    NAGP_TRY(stage_in(ctx, theta_off, (size_t)P + 1, &a.theta_off));
    a.theta_stride_k = theta_stride_k;
    NAGP_TRY(stage_in(ctx, noise, (size_t)(noise_stride_k ? (K - 1) * noise_stride_k + P : P), &a.noise));
    a.noise_stride_k = noise_stride_k;
    a.jitter = ctx->jitter; a.noise_pred = -1.0;
    a.n = (int)n; a.k = (int)k; a.h = 0;
    NAGP_TRY(stage_in(ctx, t, (size_t)m, &a.t));
    NAGP_TRY(stage_in(ctx, g, (size_t)m, &a.g));
    a.step = step;
    NAGP_TRY(stage_in(ctx, y1, (size_t)(y1_stride ? (B - 1) * y1_stride + n : n), &a.y1));
    a.y1_stride = y1_stride;
    NAGP_TRY(stage_in(ctx, y2, (size_t)(K * k), &a.y2));
    a.ya = 1.0; a.yb = 0.0;
    NAGP_TRY(stage_out(ctx, logml, (size_t)B, &a.logml_m));
    int32_t *d_info = nullptr;
    NAGP_TRY(stage_out(ctx, info, (size_t)B, &d_info));
    a.info = d_info;
    NAGP_TRY(scratch(ctx, (size_t)B * (size_t)(nt * (nt + 1) / 2) * 64, &a.Lkeep));
    NAGP_TRY(scratch(ctx, (size_t)B * nt * 8, &a.zkeep));
    // tile version of the gradient kernel: needs a strictly increasing lag grid (one point per grid index) or
    // pairwise times, and the factor resident in shared memory
    bool increasing = true;
    if (g) {
        std::vector<int32_t> gh((size_t)m);
        if (on_device(g)) {
            NAGP_CUDA(ctx, cudaMemcpyAsync(gh.data(), g, m * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
            NAGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        } else {
            std::memcpy(gh.data(), g, m * sizeof(int32_t));
        }
        for (int64_t i = 1; i < m; ++i) increasing = increasing && gh[i] > gh[i - 1];
    }
    if (increasing && ctx->variant != 1)
        job->gpl = plan_grad_tile((int)m, a.G, a.ntab_cap, a.ncp_cap, ctx->smem_optin, ctx->smem_per_sm);
    if (job->gpl.ok) NAGP_TRY(scratch(ctx, (size_t)B * nt * 64, &a.Wkeep));
    // the tile kernel is the one that keeps the factor
    int64_t nth = 1;
    for (int64_t p = 0; p < P; ++p) nth = std::max(nth, theta_off[p + 1] - theta_off[p]);
    job->vpl = plan_fused_v2((int)m, a.G, (int)std::min<int64_t>(nth, MAX_THETA), a.ntab_cap, a.ncp_cap, ctx->smem_optin,
                             ctx->smem_per_sm);
    if (!job->vpl.ok) return fail(ctx, NAGP_E_SIZE, "nagp_logml_grad: problem too large for the tile kernel");
    if ((int64_t)ctx->compiled.size() == P)
        NAGP_TRY(stage_in(ctx, ctx->compiled.data(), ctx->compiled.size(), &a.compiled));
    job->vgrid = fused_v2_grid(job->vpl, B, ctx->num_sms);
    if (job->vpl.scratch_stride) NAGP_TRY(scratch(ctx, (size_t)job->vgrid * job->vpl.scratch_stride, &job->vscr));
    NAGP_TRY(scratch(ctx, 1, &job->vcounter));

    GradArgs &ga = job->ga;
    ga.B = B; ga.P = P; ga.prog = a.prog; ga.prog_off = a.prog_off; ga.theta = a.theta; ga.theta_off = a.theta_off;
    ga.theta_stride_k = theta_stride_k; ga.n = (int)m; ga.t = a.t; ga.g = a.g; ga.step = step;
    ga.L = a.Lkeep; ga.z = a.zkeep; ga.info = d_info;
    NAGP_TRY(stage_out(ctx, grad_theta, (size_t)(K * ntheta), &ga.grad_theta));
    NAGP_TRY(stage_out(ctx, grad_noise, (size_t)B, &ga.grad_noise));
    if (job->gpl.ok) {
        ga.Winv = a.Wkeep; ga.G = a.G; ga.ntab_cap = a.ntab_cap; ga.ncp_cap = a.ncp_cap;
        ga.compiled = a.compiled;
        job->ggrid = grad_tile_grid(job->gpl, B, ctx->num_sms);
        if (getenv("NAGP_DEBUG"))
            fprintf(stderr, "[nagp] grad tile kernel: n=%d G=%d Gd=%d smem=%zu B region=%d scratch/CTA=%d grid=%d\n",
                    (int)m, a.G, job->gpl.Gd, job->gpl.smem_bytes, job->gpl.region_bytes, job->gpl.scratch_stride, job->ggrid);
        NAGP_TRY(scratch(ctx, (size_t)job->ggrid * job->gpl.scratch_stride, &job->gscr));
        NAGP_TRY(scratch(ctx, 1, &job->gcounter));
    } else {
        bool s_in_smem = true;
        job->col_smem = grad_smem_bytes((int)m, ctx->smem_optin, &s_in_smem);
        job->ggrid = (int)std::min<int64_t>(B, ctx->num_sms);
        if (!s_in_smem) NAGP_TRY(scratch(ctx, (size_t)job->ggrid * m * m, &ga.S));
    }
    return NAGP_OK;
}

int32_t enqueue_grad_job(nagp_ctx *ctx, const GradJob &job)
{
    NAGP_CUDA(ctx, launch_fused_v2(job.a, job.vpl, job.vscr, job.vcounter, job.vgrid, ctx->stream));
    if (job.gpl.ok) NAGP_CUDA(ctx, launch_grad_tile(job.ga, job.gpl, job.gscr, job.gcounter, job.ggrid, ctx->stream));
    else NAGP_CUDA(ctx, launch_grad(job.ga, job.ggrid, job.col_smem, ctx->stream));
    ctx->launches += 2;
    return NAGP_OK;
}

}  // namespace

extern "C" {

int32_t nagp_logml_grad(nagp_ctx *ctx, int64_t K, int64_t P, const uint8_t *prog, const int64_t *prog_off,
                        const double *theta, const int64_t *theta_off, int64_t theta_stride_k,
                        const double *noise, int64_t noise_stride_k, int64_t n, int64_t k, const double *t,
                        const int32_t *g, double step, const double *y1, int64_t y1_stride, const double *y2,
                        double *logml, double *grad_theta, double *grad_noise, int32_t *info)
{
    NAGP_RANGE("nagp_logml_grad");
    if (!ctx) return NAGP_E_ARG;
    if (K <= 0 || P <= 0 || !prog || !prog_off || !theta || !theta_off || !noise || !t || !y1 || !logml ||
        !grad_theta || !grad_noise || !info || (k > 0 && !y2) || n <= 0 || k < 0 || (y1_stride != 0 && y1_stride < n))
        return fail(ctx, NAGP_E_ARG, "nagp_logml_grad: null or empty argument");
    if (n + k > fused_v2_max_q()) return fail(ctx, NAGP_E_SIZE, "nagp_logml_grad: n + k > 232 not supported yet");
    if (on_device(prog_off) || on_device(theta_off))
        return fail(ctx, NAGP_E_ARG, "prog_off/theta_off must be host arrays");
    NAGP_CUDA(ctx, cudaSetDevice(ctx->device));
    NAGP_TRY(arena_reset(ctx));
    GradJob job;
    NAGP_TRY(build_grad_job(ctx, K, P, prog, prog_off, theta, theta_off, theta_stride_k, noise, noise_stride_k, n, k, t, g,
                            step, y1, y1_stride, y2, logml, grad_theta, grad_noise, info, &job));
    NAGP_TRY(enqueue_grad_job(ctx, job));
    NAGP_TRY(finish(ctx));
    return on_device(info) ? NAGP_OK : worst_info(info, K * P);
}

// ---- (f1) HMC on the hyperparameters, leapfrog on the device ---------------------------------------------------
int32_t nagp_hmc(nagp_ctx *ctx, int64_t K, int64_t P, const uint8_t *prog, const int64_t *prog_off,
                 const int64_t *theta_off, const int32_t *slot_kind, const double *slot_a, const double *slot_b,
                 int32_t noise_kind, double noise_a, double noise_b, double *z, double *noise_z,
                 int64_t n, int64_t k, const double *t, const int32_t *g, double step,
                 const double *y1, int64_t y1_stride, const double *y2,
                 int64_t n_steps, int64_t n_leapfrog, double eps,
                 const double *momenta, const double *noise_momenta, const double *log_u,
                 double *logml, int32_t *n_accept, int32_t *info)
{
    NAGP_RANGE("nagp_hmc");
    if (!ctx) return NAGP_E_ARG;
    const bool learn_noise = noise_kind != 5;
    if (K <= 0 || P <= 0 || !prog || !prog_off || !theta_off || !slot_kind || !slot_a || !slot_b || !z || !noise_z ||
        !t || !y1 || (k > 0 && !y2) || n <= 0 || k < 0 || n_steps < 0 || n_leapfrog <= 0 || !(eps > 0.0) ||
        (n_steps > 0 && (!momenta || !log_u || (learn_noise && !noise_momenta))) || !logml || !n_accept || !info ||
        (y1_stride != 0 && y1_stride < n))
        return fail(ctx, NAGP_E_ARG, "nagp_hmc: null or empty argument");
    if (n + k > fused_v2_max_q()) return fail(ctx, NAGP_E_SIZE, "nagp_hmc: n + k > 232 not supported yet");
    if (on_device(prog_off) || on_device(theta_off) || on_device(z) || on_device(noise_z))
        return fail(ctx, NAGP_E_ARG, "nagp_hmc: prog_off/theta_off/z/noise_z must be host arrays");
    NAGP_CUDA(ctx, cudaSetDevice(ctx->device));
    NAGP_TRY(arena_reset(ctx));
    const int64_t total = theta_off[P], B = K * P;
    HmcArgs h{};
    h.K = K; h.P = P; h.total = total; h.L = (int)n_leapfrog; h.eps = eps;
    h.noise_kind = noise_kind; h.noise_a = noise_a; h.noise_b = noise_b;
    NAGP_TRY(stage_in(ctx, slot_kind, (size_t)total, &h.slot_kind));
    NAGP_TRY(stage_in(ctx, slot_a, (size_t)total, &h.slot_a));
    NAGP_TRY(stage_in(ctx, slot_b, (size_t)total, &h.slot_b));
    NAGP_TRY(stage_in(ctx, momenta, (size_t)(n_steps * K * total), &h.momenta));
    NAGP_TRY(stage_in(ctx, noise_momenta, (size_t)(learn_noise ? n_steps * B : 0), &h.noise_momenta));
    NAGP_TRY(stage_in(ctx, log_u, (size_t)(n_steps * B), &h.log_u));
    // chain state on the device; z / noise_z are copied in here and copied back at the end
    const double *z_in = nullptr, *nz_in = nullptr;
    NAGP_TRY(stage_in(ctx, (const double *)z, (size_t)(K * total), &z_in));
    NAGP_TRY(stage_in(ctx, (const double *)noise_z, (size_t)B, &nz_in));
    h.Z = const_cast<double *>(z_in); h.NZ = const_cast<double *>(nz_in);
    ctx->outs.push_back({z, h.Z, (size_t)(K * total) * sizeof(double)});
    ctx->outs.push_back({noise_z, h.NZ, (size_t)B * sizeof(double)});
    NAGP_TRY(scratch(ctx, (size_t)(K * total), &h.Zq));
    NAGP_TRY(scratch(ctx, (size_t)B, &h.NZq));
    NAGP_TRY(scratch(ctx, (size_t)(K * total), &h.mom));
    NAGP_TRY(scratch(ctx, (size_t)B, &h.mnz));
    NAGP_TRY(scratch(ctx, (size_t)(K * total), &h.gZ));
    NAGP_TRY(scratch(ctx, (size_t)B, &h.gNZ));
    NAGP_TRY(scratch(ctx, (size_t)(K * total), &h.gq));
    NAGP_TRY(scratch(ctx, (size_t)B, &h.gnq));
    NAGP_TRY(scratch(ctx, (size_t)B, &h.lp));
    NAGP_TRY(scratch(ctx, (size_t)(K * total), &h.theta));
    NAGP_TRY(scratch(ctx, (size_t)B, &h.noise));
    NAGP_TRY(scratch(ctx, (size_t)B, &h.logml_q));
    NAGP_TRY(scratch(ctx, (size_t)(K * total), &h.grad_theta));
    NAGP_TRY(scratch(ctx, (size_t)B, &h.grad_noise));
    NAGP_TRY(scratch(ctx, (size_t)B, &h.info_q));
    NAGP_TRY(scratch(ctx, 1, &h.iter));
    NAGP_TRY(stage_out(ctx, logml, (size_t)B, &h.logml_cur));
    NAGP_TRY(stage_out(ctx, n_accept, (size_t)B, &h.n_accept));
    NAGP_TRY(stage_out(ctx, info, (size_t)B, &h.info_cur));

    GradJob job;
    NAGP_TRY(build_grad_job(ctx, K, P, prog, prog_off, h.theta, theta_off, total, h.noise, P, n, k, t, g, step, y1,
                            y1_stride, y2, h.logml_q, h.grad_theta, h.grad_noise, h.info_q, &job));
    h.theta_off = job.a.theta_off;

    // initial state: log posterior and gradient at (Z, NZ)
    NAGP_CUDA(ctx, launch_hmc_stage(h, HMC_INIT, ctx->num_sms, ctx->stream));
    NAGP_TRY(enqueue_grad_job(ctx, job));
    NAGP_CUDA(ctx, launch_hmc_stage(h, HMC_INIT_DONE, ctx->num_sms, ctx->stream));
    ctx->launches += 2;

    // one HMC iteration = begin, L x (half kick + drift + transform, logML + gradient, half kick), accept.
    // Every kernel reads the iteration counter from device memory, so the sequence is identical each time:
    // captured once into a CUDA graph and replayed n_steps times (no host round trip inside the chain).
    auto enqueue_iteration = [&]() -> int32_t {
        NAGP_CUDA(ctx, launch_hmc_stage(h, HMC_BEGIN, ctx->num_sms, ctx->stream));
        for (int l = 0; l < h.L; ++l) {
            NAGP_CUDA(ctx, launch_hmc_stage(h, HMC_LEAP_PRE, ctx->num_sms, ctx->stream));
            NAGP_TRY(enqueue_grad_job(ctx, job));
            NAGP_CUDA(ctx, launch_hmc_stage(h, HMC_LEAP_POST, ctx->num_sms, ctx->stream));
        }
        NAGP_CUDA(ctx, launch_hmc_stage(h, HMC_ACCEPT, ctx->num_sms, ctx->stream));
        ctx->launches += 2 + 2 * h.L;
        return NAGP_OK;
    };
    bool graphed = false;
    if (n_steps > 1 && !getenv("NAGP_NO_GRAPH")) {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        const int64_t launches_before = ctx->launches;
        if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
            const int32_t rc = enqueue_iteration();
            const cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
            if (rc == NAGP_OK && e == cudaSuccess && graph &&
                cudaGraphInstantiate(&exec, graph, nullptr, nullptr, 0) == cudaSuccess) {
                graphed = true;
                for (int64_t it = 0; it < n_steps && graphed; ++it)
                    if (cudaGraphLaunch(exec, ctx->stream) != cudaSuccess) graphed = false;
                ctx->launches = launches_before + (int64_t)(2 + 4 * h.L) * n_steps;
            }
            if (exec) {
                cudaStreamSynchronize(ctx->stream);
                cudaGraphExecDestroy(exec);
            }
            if (graph) cudaGraphDestroy(graph);
        }
        if (!graphed) {
            cudaGetLastError();
            ctx->launches = launches_before;
            return fail(ctx, NAGP_E_CUDA, "nagp_hmc: CUDA graph capture of the HMC iteration failed (set NAGP_NO_GRAPH=1 to run it as plain stream launches)");
        }
    }
    if (!graphed)
        for (int64_t it = 0; it < n_steps; ++it) NAGP_TRY(enqueue_iteration());
    NAGP_TRY(finish(ctx));
    return on_device(info) ? NAGP_OK : worst_info(info, B);
}

// ---- (f4) inverse transformation + row quantiles of the forecast matrix -------------------------------------
int32_t nagp_forecast_summary(nagp_ctx *ctx, int32_t kind, double lambda, double offset, double max_value,
                              int64_t h, int64_t N, const double *x, double *x_out,
                              int64_t nq, const double *probs, double *q)
{
    NAGP_RANGE("nagp_forecast_summary");
    if (!ctx) return NAGP_E_ARG;
    if (kind < 0 || kind > 3 || h <= 0 || N <= 0 || !x || nq < 0 || (nq > 0 && (!probs || !q)) || (!x_out && nq == 0))
        return fail(ctx, NAGP_E_ARG, "nagp_forecast_summary: bad argument");
    if (nq > 0 && !on_device(probs))
        for (int64_t j = 0; j < nq; ++j)
            if (!(probs[j] >= 0.0 && probs[j] <= 1.0)) return fail(ctx, NAGP_E_ARG, "nagp_forecast_summary: probability outside [0, 1]");
    NAGP_CUDA(ctx, cudaSetDevice(ctx->device));
    NAGP_TRY(arena_reset(ctx));
    const double *d_x = nullptr, *d_probs = nullptr;
    double *d_out = nullptr, *d_rows = nullptr, *d_q = nullptr;
    NAGP_TRY(stage_in(ctx, x, (size_t)(h * N), &d_x));
    if (x_out) {
        if (x_out == x && on_device(x)) d_out = const_cast<double *>(d_x);       // in place on the device
        else NAGP_TRY(stage_out(ctx, x_out, (size_t)(h * N), &d_out));
    }
    if (nq > 0) {
        NAGP_TRY(scratch(ctx, (size_t)(h * N), &d_rows));
        NAGP_TRY(stage_in(ctx, probs, (size_t)nq, &d_probs));
        NAGP_TRY(stage_out(ctx, q, (size_t)(h * nq), &d_q));
    }
    NAGP_CUDA(ctx, launch_inverse_transform(kind, lambda, offset, max_value, h, N, d_x, d_out, d_rows, ctx->num_sms,
                                            ctx->stream));
    ctx->launches += 1;
    if (nq > 0) {
        NAGP_CUDA(ctx, launch_row_quantiles(h, N, d_rows, nq, d_probs, d_q, ctx->stream));
        ctx->launches += 1;
    }
    return finish(ctx);
}

// ---- appendable factor store for long series (SMC data annealing, BASELINE config 5) -----------------
int32_t nagp_factor_store_large(nagp_ctx *ctx, int64_t P, const uint8_t *prog, const int64_t *prog_off,
                                const double *theta, const int64_t *theta_off, const double *noise,
                                int64_t n, int64_t capacity, const double *t, const int32_t *g, double step,
                                const double *y, nagp_factor **out, double *logml, int32_t *info)
{
    NAGP_RANGE("nagp_factor_store_large");
    if (!ctx) return NAGP_E_ARG;
    if (!out || P <= 0 || !prog || !prog_off || !theta || !theta_off || !noise || !t || !y || !info || n <= 0)
        return fail(ctx, NAGP_E_ARG, "nagp_factor_store_large: null or empty argument");
    *out = nullptr;
    if (capacity < n) capacity = n;
    NAGP_TRY(check_dims(ctx, capacity, 0, 0));
    if (on_device(prog) || on_device(prog_off) || on_device(theta_off) || on_device(t) || on_device(g) || on_device(y))
        return fail(ctx, NAGP_E_ARG, "nagp_factor_store_large: prog/offsets/t/g/y must be host arrays");
    NAGP_CUDA(ctx, cudaSetDevice(ctx->device));
    NAGP_TRY(arena_reset(ctx));
    nagp_factor *f = new (std::nothrow) nagp_factor();
    if (!f) return fail(ctx, NAGP_E_ARG, "out of host memory");
    f->device = ctx->device; f->P = P; f->n = (int)n; f->k = 0; f->h = 0; f->ya = 1.0; f->yb = 0.0;
    f->appendable = true; f->cap = capacity; f->step = step; f->jitter = ctx->jitter;
    auto bail = [&](int32_t code) { nagp_factor_free(f); return code; };
    FusedArgs a{};
    int32_t rc;
    if ((rc = grid_extent(ctx, g, n, &a.G)) != NAGP_OK) return bail(rc);
    // table capacities are fixed for the life of the factor: plan them as for the full capacity
    if ((rc = plan_tables(ctx, P, prog, prog_off, theta_off, (int)capacity, a.G, &a.ntab_cap, &a.ncp_cap)) != NAGP_OK) return bail(rc);
    f->ntab_cap = a.ntab_cap; f->ncp_cap = a.ncp_cap;
    int64_t nth = 1;
    for (int64_t p = 0; p < P; ++p) nth = std::max(nth, theta_off[p + 1] - theta_off[p]);
    f->nth_cap = (int)std::min<int64_t>(nth, MAX_THETA);
    LargePlan pl = plan_large((int)n, (int)capacity, a.G, f->nth_cap, a.ntab_cap, a.ncp_cap, ctx->smem_per_sm, false);
    f->L_stride = pl.L_stride; f->W_stride = (size_t)pl.ntp_cap * 64;
    const int64_t nprog = prog_off[P], ntheta = theta_off[P];
    auto dal = [&](void **p, size_t bytes) { return cudaMalloc(p, std::max<size_t>(bytes, 16)); };
    if (dal((void **)&f->Lbig, (size_t)P * f->L_stride * 8) != cudaSuccess || dal((void **)&f->Wbig, (size_t)P * f->W_stride * 8) != cudaSuccess ||
        dal((void **)&f->d_theta, ntheta * 8) != cudaSuccess || dal((void **)&f->d_noise, P * 8) != cudaSuccess ||
        dal((void **)&f->d_t, capacity * 8) != cudaSuccess || dal((void **)&f->d_y, capacity * 8) != cudaSuccess ||
        dal((void **)&f->d_g, capacity * 4) != cudaSuccess || dal((void **)&f->d_prog, nprog) != cudaSuccess ||
        dal((void **)&f->d_prog_off, (P + 1) * 8) != cudaSuccess || dal((void **)&f->d_theta_off, (P + 1) * 8) != cudaSuccess ||
        dal((void **)&f->logml_n, P * 8) != cudaSuccess) {
        cudaGetLastError();
        return bail(fail(ctx, NAGP_E_CUDA, "factor allocation failed (capacity too large for device memory?)"));
    }
    cudaStream_t st = ctx->stream;
    cudaError_t e = cudaSuccess;
    auto cp = [&](void *d, const void *s_, size_t bytes) { if (e == cudaSuccess && bytes) e = cudaMemcpyAsync(d, s_, bytes, cudaMemcpyDefault, st); };
    cp(f->d_theta, theta, ntheta * 8); cp(f->d_noise, noise, P * 8); cp(f->d_t, t, n * 8); cp(f->d_y, y, n * 8);
    if (g) { cp(f->d_g, g, n * 4); f->g_host.assign(g, g + n); }
    cp(f->d_prog, prog, nprog); cp(f->d_prog_off, prog_off, (P + 1) * 8); cp(f->d_theta_off, theta_off, (P + 1) * 8);
    if (e != cudaSuccess) return bail(fail(ctx, NAGP_E_CUDA, cudaGetErrorString(e)));
    a.B = P; a.P = P; a.prog = f->d_prog; a.prog_off = f->d_prog_off; a.theta = f->d_theta; a.theta_off = f->d_theta_off;
    a.noise = f->d_noise; a.jitter = ctx->jitter; a.noise_pred = -1.0;
    a.n = (int)n; a.k = 0; a.h = 0; a.t = f->d_t; a.g = g ? f->d_g : nullptr; a.step = step;
    a.y1 = f->d_y; a.ya = 1.0; a.yb = 0.0; a.logml_n = f->logml_n;
    if ((rc = stage_out(ctx, info, (size_t)P, &a.info)) != NAGP_OK) return bail(rc);
    const int grid = large_grid(pl, P, ctx->num_sms, false);
    char *scr = nullptr;
    if (pl.scratch_stride && (rc = scratch(ctx, (size_t)grid * pl.scratch_stride, &scr)) != NAGP_OK) return bail(rc);
    unsigned long long *counter = nullptr;
    if ((rc = scratch(ctx, 1, &counter)) != NAGP_OK) return bail(rc);
    e = launch_chol_large(a, pl, scr, f->Lbig, 1, f->Wbig, counter, grid, st);
    if (e != cudaSuccess) return bail(fail(ctx, NAGP_E_CUDA, cudaGetErrorString(e)));
    ctx->launches += 1;
    if (logml) {
        e = cudaMemcpyAsync(logml, f->logml_n, P * sizeof(double), cudaMemcpyDefault, st);
        if (e != cudaSuccess) return bail(fail(ctx, NAGP_E_CUDA, cudaGetErrorString(e)));
    }
    if ((rc = finish(ctx)) != NAGP_OK) return bail(rc);
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return bail(fail(ctx, NAGP_E_CUDA, cudaGetErrorString(e)));
    *out = f;
    return on_device(info) ? NAGP_OK : worst_info(info, P);
}

int64_t nagp_factor_size(const nagp_factor *f) { return f ? f->n : -1; }

int32_t nagp_factor_append(nagp_ctx *ctx, nagp_factor *f, int64_t k_new, const double *t_new, const int32_t *g_new,
                           const double *y_new, double *dlogml, double *logml, int32_t *info)
{
    NAGP_RANGE("nagp_factor_append");
    if (!ctx) return NAGP_E_ARG;
    if (!f || !f->appendable || k_new <= 0 || !t_new || !y_new || !info)
        return fail(ctx, NAGP_E_ARG, "nagp_factor_append: null argument or factor not appendable");
    if (f->device != ctx->device) return fail(ctx, NAGP_E_ARG, "factor lives on another device");
    if ((int64_t)f->n + k_new > f->cap) return fail(ctx, NAGP_E_SIZE, "nagp_factor_append: capacity exceeded");
    if (f->g_host.empty() != (g_new == nullptr)) return fail(ctx, NAGP_E_ARG, "nagp_factor_append: lag-grid mode must match the stored factor");
    if (on_device(t_new) || on_device(g_new) || on_device(y_new)) return fail(ctx, NAGP_E_ARG, "nagp_factor_append: t/g/y must be host arrays");
    NAGP_CUDA(ctx, cudaSetDevice(ctx->device));
    NAGP_TRY(arena_reset(ctx));
    const int n_old = f->n, n_new = f->n + (int)k_new;
    cudaStream_t st = ctx->stream;
    NAGP_CUDA(ctx, cudaMemcpyAsync(f->d_t + n_old, t_new, k_new * 8, cudaMemcpyHostToDevice, st));
    NAGP_CUDA(ctx, cudaMemcpyAsync(f->d_y + n_old, y_new, k_new * 8, cudaMemcpyHostToDevice, st));
    FusedArgs a{};
    if (g_new) {
        NAGP_CUDA(ctx, cudaMemcpyAsync(f->d_g + n_old, g_new, k_new * 4, cudaMemcpyHostToDevice, st));
        f->g_host.insert(f->g_host.end(), g_new, g_new + k_new);
        NAGP_TRY(grid_extent(ctx, f->g_host.data(), n_new, &a.G));
    }
    a.B = f->P; a.P = f->P; a.prog = f->d_prog; a.prog_off = f->d_prog_off; a.theta = f->d_theta; a.theta_off = f->d_theta_off;
    a.noise = f->d_noise; a.jitter = f->jitter; a.noise_pred = -1.0;
    a.n = n_new; a.k = 0; a.h = 0; a.t = f->d_t; a.g = g_new ? f->d_g : nullptr; a.step = f->step;
    a.y1 = f->d_y; a.ya = 1.0; a.yb = 0.0;
    a.ntab_cap = f->ntab_cap; a.ncp_cap = f->ncp_cap;
    NAGP_TRY(stage_out(ctx, info, (size_t)f->P, &a.info));
    double *d_dl = nullptr;
    NAGP_TRY(stage_out(ctx, dlogml, (size_t)f->P, &d_dl));
    const int new_tile_rows = (n_new + 7) / 8 - n_old / 8;
    const bool one_tile_row = new_tile_rows <= 1;
    // Appends of many rows at once (the steps of a fit_smc schedule) go through the factorisation kernel restricted to the
    // new tile rows: it reads the stored factor once however many rows arrive and sums two rows per warp against each
    // panel tile, where the row-streaming kernel takes one pass over the factor per eight new tile rows.
    const char *from_env = getenv("NAGP_APPEND_BY_ROWS_FROM");        // measurement hook
    const bool by_rows = new_tile_rows >= (from_env ? atoi(from_env) : kAppendByRowsFrom) && !getenv("NAGP_APPEND_STREAM");
    LargePlan pl = plan_large(n_new, (int)f->cap, a.G, f->nth_cap, a.ntab_cap, a.ncp_cap, ctx->smem_per_sm, !by_rows, one_tile_row);
    if (pl.L_stride != f->L_stride) return fail(ctx, NAGP_E_ARG, "nagp_factor_append: internal layout mismatch");
    const int grid = large_grid(pl, f->P, ctx->num_sms, !by_rows);
    char *scr = nullptr;
    if (pl.scratch_stride) NAGP_TRY(scratch(ctx, (size_t)grid * pl.scratch_stride, &scr));
    if (by_rows) {
        unsigned long long *counter = nullptr;
        NAGP_TRY(scratch(ctx, 1, &counter));
        NAGP_CUDA(ctx, launch_chol_large(a, pl, scr, f->Lbig, 1, f->Wbig, counter, grid, st, n_old, d_dl, f->logml_n));
    } else {
        NAGP_CUDA(ctx, launch_rank_append(a, pl, scr, f->Lbig, f->Wbig, n_old, f->logml_n, d_dl, grid, st));
    }
    ctx->launches += 1;
    f->n = n_new;
    if (logml) NAGP_CUDA(ctx, cudaMemcpyAsync(logml, f->logml_n, f->P * sizeof(double), cudaMemcpyDefault, st));
    NAGP_TRY(finish(ctx));
    NAGP_CUDA(ctx, cudaStreamSynchronize(st));
    return on_device(info) ? NAGP_OK : worst_info(info, f->P);
}

int32_t nagp_append(nagp_ctx *ctx, const nagp_factor *f, int64_t K, const double *y2, double *logw, double *mu)
{
    NAGP_RANGE("nagp_append");
    if (!ctx) return NAGP_E_ARG;
    if (!f || K <= 0 || !logw || (f->k > 0 && !y2)) return fail(ctx, NAGP_E_ARG, "nagp_append: null or empty argument");
    if (f->appendable) return fail(ctx, NAGP_E_ARG, "nagp_append: factor was created by nagp_factor_store_large (use nagp_factor_append)");
    if (f->device != ctx->device) return fail(ctx, NAGP_E_ARG, "factor lives on another device");
    if (f->k > 16) return fail(ctx, NAGP_E_SIZE, "nagp_append: more than 16 nowcast points (use nagp_forecast_instances)");
    NAGP_CUDA(ctx, cudaSetDevice(ctx->device));
    NAGP_TRY(arena_reset(ctx));
    AppendArgs a{};
    a.K = K; a.P = f->P; a.k = f->k; a.h = f->h;
    NAGP_TRY(stage_in(ctx, y2, (size_t)(K * f->k), &a.y2));
    a.proj = f->proj; a.Ltail = f->Ltail; a.logw0 = f->logw0; a.ya = f->ya; a.yb = f->yb;
    NAGP_TRY(stage_out(ctx, logw, (size_t)(K * f->P), &a.logw));
    NAGP_TRY(stage_out(ctx, mu, (size_t)(K * f->P * f->h), &a.mu));
    NAGP_CUDA(ctx, launch_append(a, ctx->stream));
    ctx->launches += 1;
    return finish(ctx);
}

int32_t nagp_predict(nagp_ctx *ctx, const nagp_factor *f, double *mu, double *L)
{
    NAGP_RANGE("nagp_predict");
    if (!ctx) return NAGP_E_ARG;
    if (!f || f->appendable) return fail(ctx, NAGP_E_ARG, "nagp_predict: null or appendable-only factor");
    if (f->device != ctx->device) return fail(ctx, NAGP_E_ARG, "factor lives on another device");
    if (mu && f->k != 0) return fail(ctx, NAGP_E_ARG, "nagp_predict: means need k == 0 (use nagp_append)");
    NAGP_CUDA(ctx, cudaSetDevice(ctx->device));
    NAGP_TRY(arena_reset(ctx));
    if (mu) {
        // k == 0: the append kernel with zero nowcast points is exactly the un-scaling of proj
        AppendArgs a{};
        a.K = 1; a.P = f->P; a.k = 0; a.h = f->h;
        a.y2 = nullptr; a.proj = f->proj; a.Ltail = f->Ltail; a.logw0 = nullptr; a.ya = f->ya; a.yb = f->yb;
        double *lw;
        NAGP_TRY(scratch(ctx, (size_t)f->P, &lw));
        a.logw = lw;
        NAGP_TRY(stage_out(ctx, mu, (size_t)(f->P * f->h), &a.mu));
        NAGP_CUDA(ctx, launch_append(a, ctx->stream));
        ctx->launches += 1;
    }
    if (L)
        NAGP_CUDA(ctx, cudaMemcpyAsync(L, f->L33, (size_t)(f->P * f->h * f->h) * sizeof(double),
                                       cudaMemcpyDefault, ctx->stream));
    NAGP_TRY(finish(ctx));
    if (L && !on_device(L)) NAGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NAGP_OK;
}

int32_t nagp_ess(nagp_ctx *ctx, int64_t K, int64_t P, const double *logw, double *ess, double *w)
{
    NAGP_RANGE("nagp_ess");
    if (!ctx) return NAGP_E_ARG;
    if (K <= 0 || P <= 0 || !logw || !ess) return fail(ctx, NAGP_E_ARG, "nagp_ess: null or empty argument");
    NAGP_CUDA(ctx, cudaSetDevice(ctx->device));
    NAGP_TRY(arena_reset(ctx));
    DrawArgs a{};
    a.K = K; a.P = P; a.h = 0; a.D = 0;
    NAGP_TRY(stage_in(ctx, logw, (size_t)(K * P), &a.logw));
    NAGP_TRY(stage_out(ctx, ess, (size_t)K, &a.ess_out));
    NAGP_TRY(stage_out(ctx, w, (size_t)(K * P), &a.w_out));
    NAGP_CUDA(ctx, launch_draw(a, ctx->stream));
    ctx->launches += 1;
    return finish(ctx);
}

int32_t nagp_draw(nagp_ctx *ctx, int64_t K, int64_t P, int64_t h, int64_t D, const double *logw,
                  const double *mu, int64_t mu_stride_k, const double *L, int64_t l_stride_k,
                  const int32_t *comp, const double *u, const double *u_res, double ess_thr,
                  const double *zeta, double *x, double *ess_out, int32_t *comp_out)
{
    NAGP_RANGE("nagp_draw");
    if (!ctx) return NAGP_E_ARG;
    if (K <= 0 || P <= 0 || h <= 0 || D <= 0 || !logw || !mu || !L || !zeta || !x || (!comp && !u))
        return fail(ctx, NAGP_E_ARG, "nagp_draw: null or empty argument");
    if (u_res && !u) return fail(ctx, NAGP_E_ARG, "nagp_draw: resampling needs u");
    NAGP_CUDA(ctx, cudaSetDevice(ctx->device));
    NAGP_TRY(arena_reset(ctx));
    DrawArgs a{};
    a.K = K; a.P = P; a.h = (int)h; a.D = D;
    NAGP_TRY(stage_in(ctx, logw, (size_t)(K * P), &a.logw));
    NAGP_TRY(stage_in(ctx, mu, (size_t)((K - 1) * mu_stride_k + P * h), &a.mu));
    a.mu_stride_k = mu_stride_k;
    NAGP_TRY(stage_in(ctx, L, (size_t)((K - 1) * l_stride_k + P * h * h), &a.L));
    a.l_stride_k = l_stride_k;
    NAGP_TRY(stage_in(ctx, comp, (size_t)(K * D), &a.comp));
    NAGP_TRY(stage_in(ctx, u, (size_t)(K * D), &a.u));
    NAGP_TRY(stage_in(ctx, u_res, (size_t)(K * P), &a.u_res));
    a.ess_thr = ess_thr;
    NAGP_TRY(stage_in(ctx, zeta, (size_t)(K * D * h), &a.zeta));
    NAGP_TRY(stage_out(ctx, x, (size_t)(K * D * h), &a.x));
    NAGP_TRY(stage_out(ctx, ess_out, (size_t)K, &a.ess_out));
    NAGP_TRY(stage_out(ctx, comp_out, (size_t)(K * D), &a.comp_out));
    NAGP_CUDA(ctx, launch_draw(a, ctx->stream));
    ctx->launches += 1;
    return finish(ctx);
}

int32_t nagp_forecast_with_nowcasts_theta(nagp_ctx *ctx, int64_t K, int64_t P, int64_t D, const uint8_t *prog,
                                          const int64_t *prog_off, const double *theta, const int64_t *theta_off,
                                          int64_t theta_stride_k, const double *noise, int64_t noise_stride_k,
                                          double noise_pred, int64_t n, int64_t k, int64_t h, const double *t,
                                          const int32_t *g, double step, const double *y1, const double *y2, double ya,
                                          double yb, const double *logw0, const int32_t *comp, const double *u,
                                          const double *u_res, double ess_thr, const double *zeta, double *x,
                                          double *logw_out, double *ess_out, int32_t *info)
{
    NAGP_RANGE("nagp_forecast_with_nowcasts_theta");
    if (!ctx) return NAGP_E_ARG;
    if (K <= 0 || P <= 0 || D <= 0 || h <= 0 || !prog || !prog_off || !theta || !theta_off || !noise || !t || !y1 ||
        (k > 0 && !y2) || !zeta || !x || !info || (!comp && !u) || ya == 0.0)
        return fail(ctx, NAGP_E_ARG, "nagp_forecast_with_nowcasts_theta: null or empty argument");
    if (u_res && !u) return fail(ctx, NAGP_E_ARG, "nagp_forecast_with_nowcasts_theta: resampling needs u");
    NAGP_TRY(check_dims(ctx, n, k, h));
    if (on_device(prog_off) || on_device(theta_off))
        return fail(ctx, NAGP_E_ARG, "prog_off/theta_off must be host arrays");
    NAGP_CUDA(ctx, cudaSetDevice(ctx->device));
    NAGP_TRY(arena_reset(ctx));
    const int64_t q = n + k + h, B = K * P;

    // (1) one fused Gram + factorisation per (scenario, particle): log-weights, mu*, L33 stay on the device
    FusedArgs a{};
    NAGP_TRY(grid_extent(ctx, g, q, &a.G));
    NAGP_TRY(plan_tables(ctx, P, prog, prog_off, theta_off, (int)q, a.G, &a.ntab_cap, &a.ncp_cap));
    const int64_t nprog = prog_off[P], ntheta = theta_off[P];
    a.B = B; a.P = P;
    NAGP_TRY(stage_in(ctx, prog, (size_t)nprog, &a.prog));
    NAGP_TRY(stage_in(ctx, prog_off, (size_t)P + 1, &a.prog_off));
    NAGP_TRY(stage_in(ctx, theta, (size_t)(theta_stride_k ? (K - 1) * theta_stride_k + ntheta : ntheta), &a.theta));
    NAGP_TRY(stage_in(ctx, theta_off, (size_t)P + 1, &a.theta_off));
    a.theta_stride_k = theta_stride_k;
    NAGP_TRY(stage_in(ctx, noise, (size_t)(noise_stride_k ? (K - 1) * noise_stride_k + P : P), &a.noise));
    a.noise_stride_k = noise_stride_k;
    a.jitter = ctx->jitter; a.noise_pred = noise_pred;
    a.n = (int)n; a.k = (int)k; a.h = (int)h;
    NAGP_TRY(stage_in(ctx, t, (size_t)q, &a.t));
    NAGP_TRY(stage_in(ctx, g, (size_t)q, &a.g));
    a.step = step;
    NAGP_TRY(stage_in(ctx, y1, (size_t)n, &a.y1));
    NAGP_TRY(stage_in(ctx, y2, (size_t)(K * k), &a.y2));
    a.ya = ya; a.yb = yb;
    NAGP_TRY(stage_in(ctx, logw0, (size_t)P, &a.logw0));
    if (logw_out) NAGP_TRY(stage_out(ctx, logw_out, (size_t)B, &a.logw));
    else NAGP_TRY(scratch(ctx, (size_t)B, &a.logw));
    NAGP_TRY(scratch(ctx, (size_t)(B * h), &a.mu));
    NAGP_TRY(scratch(ctx, (size_t)(B * h * h), &a.L33));
    NAGP_TRY(stage_out(ctx, info, (size_t)B, &a.info));
    // the inputs of the draw kernel travel on the copy stream, under the fused kernel (ordered after whatever the
    // caller's stream has already queued; the arena memory they land in is not touched by the fused kernel)
    DrawArgs d{};
    NAGP_CUDA(ctx, cudaEventRecord(ctx->fork, ctx->stream));
    NAGP_TRY(run_fused(ctx, a, theta_off));          // its own inputs go first through the copy engine
    NAGP_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->fork, 0));
    NAGP_TRY(stage_in(ctx, comp, (size_t)(K * D), &d.comp, ctx->copy_stream));
    NAGP_TRY(stage_in(ctx, u, (size_t)(K * D), &d.u, ctx->copy_stream));
    NAGP_TRY(stage_in(ctx, u_res, (size_t)(K * P), &d.u_res, ctx->copy_stream));
    NAGP_TRY(stage_in(ctx, zeta, (size_t)(K * D * h), &d.zeta, ctx->copy_stream));
    NAGP_CUDA(ctx, cudaEventRecord(ctx->copy_done, ctx->copy_stream));

    // (2) maybe_resample! + rand(MixtureModel, D) per scenario
    d.K = K; d.P = P; d.h = (int)h; d.D = D;
    d.logw = a.logw; d.mu = a.mu; d.mu_stride_k = P * h; d.L = a.L33; d.l_stride_k = P * h * h;
    d.ess_thr = ess_thr;
    NAGP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->copy_done, 0));
    NAGP_TRY(stage_out(ctx, x, (size_t)(K * D * h), &d.x));
    NAGP_TRY(stage_out(ctx, ess_out, (size_t)K, &d.ess_out));
    NAGP_CUDA(ctx, launch_draw(d, ctx->stream));
    ctx->launches += 1;
    NAGP_TRY(finish(ctx));
    return on_device(info) ? NAGP_OK : worst_info(info, B);
}

int32_t nagp_forecast_with_nowcasts(nagp_ctx *ctx, int64_t K, int64_t P, int64_t D, const uint8_t *prog,
                                    const int64_t *prog_off, const double *theta, const int64_t *theta_off,
                                    const double *noise, double noise_pred, int64_t n, int64_t k, int64_t h,
                                    const double *t, const int32_t *g, double step, const double *y1,
                                    const double *y2, double ya, double yb, const double *logw0,
                                    const int32_t *comp, const double *u, const double *u_res,
                                    double ess_thr, const double *zeta, double *x, double *logw_out,
                                    double *ess_out, int32_t *info)
{
    NAGP_RANGE("nagp_forecast_with_nowcasts");
    if (!ctx) return NAGP_E_ARG;
    if (K <= 0 || P <= 0 || D <= 0 || h <= 0 || !prog || !prog_off || !theta || !theta_off || !noise || !t ||
        !y1 || (k > 0 && !y2) || !zeta || !x || !info || (!comp && !u) || ya == 0.0)
        return fail(ctx, NAGP_E_ARG, "nagp_forecast_with_nowcasts: null or empty argument");
    if (u_res && !u) return fail(ctx, NAGP_E_ARG, "nagp_forecast_with_nowcasts: resampling needs u");
    if (k > 16) return fail(ctx, NAGP_E_SIZE, "nagp_forecast_with_nowcasts: more than 16 nowcast points (use nagp_forecast_instances + nagp_draw)");
    NAGP_TRY(check_dims(ctx, n, k, h));
    if (on_device(prog_off) || on_device(theta_off))
        return fail(ctx, NAGP_E_ARG, "prog_off/theta_off must be host arrays");
    NAGP_CUDA(ctx, cudaSetDevice(ctx->device));
    NAGP_TRY(arena_reset(ctx));
    const int64_t q = n + k + h, kh = k + h;

    // (1) one factorisation per particle over [train | nowcast | forecast]
    FusedArgs a{};
    NAGP_TRY(grid_extent(ctx, g, q, &a.G));
    NAGP_TRY(plan_tables(ctx, P, prog, prog_off, theta_off, (int)q, a.G, &a.ntab_cap, &a.ncp_cap));
    a.B = P; a.P = P;
    NAGP_TRY(stage_in(ctx, prog, (size_t)prog_off[P], &a.prog));
    NAGP_TRY(stage_in(ctx, prog_off, (size_t)P + 1, &a.prog_off));
    NAGP_TRY(stage_in(ctx, theta, (size_t)theta_off[P], &a.theta));
    NAGP_TRY(stage_in(ctx, theta_off, (size_t)P + 1, &a.theta_off));
    NAGP_TRY(stage_in(ctx, noise, (size_t)P, &a.noise));
    a.jitter = ctx->jitter; a.noise_pred = noise_pred;
    a.n = (int)n; a.k = (int)k; a.h = (int)h;
    NAGP_TRY(stage_in(ctx, t, (size_t)q, &a.t));
    NAGP_TRY(stage_in(ctx, g, (size_t)q, &a.g));
    a.step = step;
    NAGP_TRY(stage_in(ctx, y1, (size_t)n, &a.y1));
    a.y2 = nullptr; a.ya = ya; a.yb = yb;
    NAGP_TRY(scratch(ctx, (size_t)P, &a.logml_n));
    NAGP_TRY(scratch(ctx, (size_t)(P * kh), &a.proj));
    NAGP_TRY(scratch(ctx, (size_t)(P * kh * kh), &a.Ltail));
    NAGP_TRY(scratch(ctx, (size_t)(P * h * h), &a.L33));
    NAGP_TRY(stage_out(ctx, info, (size_t)P, &a.info));
    NAGP_TRY(run_fused(ctx, a, theta_off));

    // (2) add_data! for every scenario: O(k^2 + hk) per (scenario, particle)
    AppendArgs ap{};
    ap.K = K; ap.P = P; ap.k = (int)k; ap.h = (int)h;
    NAGP_TRY(stage_in(ctx, y2, (size_t)(K * k), &ap.y2));
    ap.proj = a.proj; ap.Ltail = a.Ltail;
    NAGP_TRY(stage_in(ctx, logw0, (size_t)P, &ap.logw0));
    ap.ya = ya; ap.yb = yb;
    if (logw_out) NAGP_TRY(stage_out(ctx, logw_out, (size_t)(K * P), &ap.logw));
    else NAGP_TRY(scratch(ctx, (size_t)(K * P), &ap.logw));
    NAGP_TRY(scratch(ctx, (size_t)(K * P * h), &ap.mu));
    NAGP_CUDA(ctx, launch_append(ap, ctx->stream));
    ctx->launches += 1;

    // (3) maybe_resample! + rand(MixtureModel, D)
    DrawArgs d{};
    d.K = K; d.P = P; d.h = (int)h; d.D = D;
    d.logw = ap.logw; d.mu = ap.mu; d.mu_stride_k = P * h; d.L = a.L33; d.l_stride_k = 0;
    NAGP_TRY(stage_in(ctx, comp, (size_t)(K * D), &d.comp));
    NAGP_TRY(stage_in(ctx, u, (size_t)(K * D), &d.u));
    NAGP_TRY(stage_in(ctx, u_res, (size_t)(K * P), &d.u_res));
    d.ess_thr = ess_thr;
    NAGP_TRY(stage_in(ctx, zeta, (size_t)(K * D * h), &d.zeta));
    NAGP_TRY(stage_out(ctx, x, (size_t)(K * D * h), &d.x));
    NAGP_TRY(stage_out(ctx, ess_out, (size_t)K, &d.ess_out));
    NAGP_CUDA(ctx, launch_draw(d, ctx->stream));
    ctx->launches += 1;
    NAGP_TRY(finish(ctx));
    return on_device(info) ? NAGP_OK : worst_info(info, P);
}

}  // extern "C"
