// extern "C" surface of libvpm_b200.so (declared in include/vpm_b200.h): contexts, particle storage,
// spaces, operator-level entry points and the whole-step device-resident steppers.
#include <dlfcn.h>
#include <sched.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <map>
#include <mutex>
#include <new>

#include "vpm_internal.h"

namespace vpm {

static thread_local std::string g_err;

void set_error(const std::string& msg) { g_err = msg; }
int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}

static int grow(double** buf, size_t* cap, size_t doubles, cudaStream_t stream)
{
    if (*cap >= doubles) return VPM_OK;
    if (*buf) {
        VPM_CUDA(cudaStreamSynchronize(stream));
        VPM_CUDA(cudaFree(*buf));
        *buf = nullptr;
        *cap = 0;
    }
    size_t want = doubles + doubles / 4 + 64;
    cudaError_t e = cudaMalloc((void**)buf, want * sizeof(double));
    if (e != cudaSuccess) return fail(VPM_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    *cap = want;
    return VPM_OK;
}

int kernel_occupancy(vpm_ctx* ctx, const void* kern, int block, size_t smem, int* occ)
{
    struct Key {
        int dev;
        const void* k;
        size_t smem;
        bool operator<(const Key& o) const { return dev != o.dev ? dev < o.dev : (k != o.k ? k < o.k : smem < o.smem); }
    };
    static std::map<Key, int> cache;                           // occupancy per (device, kernel, smem)
    static std::map<std::pair<int, const void*>, size_t> attr;  // largest dynamic-smem attribute set so far
    static std::mutex mu;                                      // contexts of different GPUs may live on different host threads
    std::lock_guard<std::mutex> guard(mu);
    size_t& cur = attr[{ctx->device, kern}];
    if (smem > cur) {  // the attribute is per kernel: only ever raise it
        VPM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cur = smem;
    }
    const Key key{ctx->device, kern, smem};
    auto it = cache.find(key);
    if (it != cache.end()) {
        *occ = it->second;
        return VPM_OK;
    }
    VPM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, block, smem));
    cache[key] = *occ;
    return VPM_OK;
}

bool pdl_enabled()
{
    static const bool on = []() {
        const char* e = getenv("VPM_TUNE_PDL");
        return !(e && atoi(e) == 0);
    }();
    return on;
}

int ensure_partials(vpm_ctx* ctx, size_t doubles) { return grow(&ctx->partials, &ctx->partials_cap, doubles, ctx->stream); }
int ensure_red(vpm_ctx* ctx, size_t doubles) { return grow(&ctx->red, &ctx->red_cap, doubles, ctx->stream); }
int ensure_staging(vpm_ctx* ctx, size_t doubles) { return grow(&ctx->staging, &ctx->staging_cap, doubles, ctx->stream); }

// ---- NCCL through dlopen: no link-time dependency, works with torch's bundled libnccl.so.2 ----
struct NcclUniqueId {
    char internal[128];
};
struct Nccl {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclUniqueId, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

static Nccl* nccl_api()
{
    static Nccl api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (api.handle) {
            api.GetUniqueId = (int (*)(NcclUniqueId*))dlsym(api.handle, "ncclGetUniqueId");
            api.CommInitRank = (int (*)(void**, int, NcclUniqueId, int))dlsym(api.handle, "ncclCommInitRank");
            api.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(api.handle, "ncclAllReduce");
            api.CommDestroy = (int (*)(void*))dlsym(api.handle, "ncclCommDestroy");
            api.GetErrorString = (const char* (*)(int))dlsym(api.handle, "ncclGetErrorString");
        }
    }
    if (!api.handle || !api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) return nullptr;
    return &api;
}

int comm_allreduce(vpm_ctx* ctx, double* buf, size_t count)
{
    if (!ctx->comm.comm) return VPM_OK;
    const int kFloat64 = 8, kSum = 0;
    int r = ctx->comm.api->AllReduce(buf, buf, count, kFloat64, kSum, ctx->comm.comm, ctx->stream);
    if (r != 0)
        return fail(VPM_ERR_COMM, std::string("ncclAllReduce: ") + (ctx->comm.api->GetErrorString ? ctx->comm.api->GetErrorString(r) : "error"));
    return VPM_OK;
}

void prof_begin(vpm_ctx* ctx, int kind, int sub)
{
    if (!ctx->profile) return;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    ctx->prof_events.push_back(e0);
    ctx->prof_events.push_back(e1);
    ctx->prof_kinds.push_back(kind);
    ctx->prof_subs.push_back(sub);
    cudaEventRecord(e0, ctx->stream);
}

void prof_end(vpm_ctx* ctx)
{
    if (!ctx->profile) return;
    cudaEventRecord(ctx->prof_events.back(), ctx->stream);
}

template <typename T>
static int upload(vpm_ctx* ctx, double** dst, const std::vector<T>& src)
{
    VPM_CUDA(cudaMalloc((void**)dst, sizeof(double) * (src.empty() ? 1 : src.size())));
    if (!src.empty()) VPM_CUDA(cudaMemcpy(*dst, src.data(), sizeof(double) * src.size(), cudaMemcpyHostToDevice));
    return VPM_OK;
}

static int grow_diag(vpm_ctx* ctx, double** buf, size_t* cap, size_t doubles)
{
    int rc = grow(buf, cap, doubles, ctx->stream);
    if (rc) return rc;
    VPM_CUDA(cudaMemsetAsync(*buf, 0, sizeof(double) * doubles, ctx->stream));
    return VPM_OK;
}

}  // namespace vpm

using namespace vpm;

#define VPM_CHECK(expr)             \
    do {                            \
        int _rc = (expr);           \
        if (_rc != VPM_OK) return _rc; \
    } while (0)
#define VPM_REQUIRE(cond, msg) \
    do {                       \
        if (!(cond)) return fail(VPM_ERR_INVALID, msg); \
    } while (0)

// A peer that never showed up in a fused all-reduce leaves its sequence number in the local mailbox's error word
// (p2p.cuh).  Called by the synchronous steppers after their stream synchronisation.
static int p2p_status(vpm_ctx* ctx)
{
    if (!ctx->p2p_local || ctx->p2p.nranks <= 1) return VPM_OK;
    unsigned long long e = 0;
    VPM_CUDA(cudaMemcpy(&e, &ctx->p2p_local->error, sizeof(e), cudaMemcpyDeviceToHost));
    if (e) return fail(VPM_ERR_COMM, "peer-memory all-reduce timed out waiting for a rank (sequence " + std::to_string(e) + "); results are invalid");
    return VPM_OK;
}

extern "C" {

const char* vpm_last_error(void) { return g_err.c_str(); }
int vpm_version(void) { return 100; }

int vpm_ctx_create(int device, void* stream, vpm_ctx** out)
{
    VPM_REQUIRE(out, "vpm_ctx_create: out is NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(VPM_ERR_CUDA, std::string("no CUDA device available (libvpm_b200 has no CPU fallback): ") + cudaGetErrorString(e));
    VPM_REQUIRE(device >= 0 && device < ndev, "vpm_ctx_create: bad device index");
    VPM_CUDA(cudaSetDevice(device));
    vpm_ctx* c = new (std::nothrow) vpm_ctx();
    if (!c) return fail(VPM_ERR_NOMEM, "out of host memory");
    c->device = device;
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        cudaError_t es = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (es != cudaSuccess) {
            delete c;
            return fail(VPM_ERR_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(es));
        }
        c->own_stream = true;
    }
    cudaDeviceProp prop;
    VPM_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    c->smem_sm = prop.sharedMemPerMultiprocessor;
    c->smem_reserved = prop.reservedSharedMemPerBlock;
    *out = c;
    return VPM_OK;
}

int vpm_ctx_destroy(vpm_ctx* ctx)
{
    if (!ctx) return VPM_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    vpm_comm_destroy(ctx);
    vpm_p2p_detach(ctx);
    cudaFree(ctx->partials);
    cudaFree(ctx->red);
    cudaFree(ctx->staging);
    for (cudaEvent_t e : ctx->copy_events) cudaEventDestroy(e);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return VPM_OK;
}

int vpm_sync(vpm_ctx* ctx)
{
    VPM_REQUIRE(ctx, "vpm_sync: ctx is NULL");
    VPM_CUDA(cudaStreamSynchronize(ctx->stream));
    return VPM_OK;
}

int vpm_device_info(vpm_ctx* ctx, int* sm_count, int64_t* smem_optin_bytes, int64_t* total_mem_bytes)
{
    VPM_REQUIRE(ctx, "vpm_device_info: ctx is NULL");
    if (sm_count) *sm_count = ctx->sm_count;
    if (smem_optin_bytes) *smem_optin_bytes = (int64_t)ctx->smem_optin;
    if (total_mem_bytes) {
        size_t fr = 0, tot = 0;
        VPM_CUDA(cudaMemGetInfo(&fr, &tot));
        *total_mem_bytes = (int64_t)tot;
    }
    return VPM_OK;
}

int64_t vpm_launch_count(vpm_ctx* ctx) { return ctx ? (int64_t)ctx->launches : 0; }

int vpm_profile(vpm_ctx* ctx, int enable)
{
    VPM_REQUIRE(ctx, "vpm_profile: ctx is NULL");
    VPM_CUDA(cudaStreamSynchronize(ctx->stream));
    for (cudaEvent_t e : ctx->prof_events) cudaEventDestroy(e);
    ctx->prof_events.clear();
    ctx->prof_kinds.clear();
    ctx->prof_subs.clear();
    ctx->profile = enable != 0;
    return VPM_OK;
}

int vpm_profile_get(vpm_ctx* ctx, double* ms_by_kind, int64_t* count_by_kind)
{
    VPM_REQUIRE(ctx && ms_by_kind && count_by_kind, "vpm_profile_get: NULL argument");
    VPM_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < PROF_NKIND; k++) { ms_by_kind[k] = 0.0; count_by_kind[k] = 0; }
    for (size_t i = 0; i < ctx->prof_kinds.size(); i++) {
        float ms = 0.f;
        VPM_CUDA(cudaEventElapsedTime(&ms, ctx->prof_events[2 * i], ctx->prof_events[2 * i + 1]));
        ms_by_kind[ctx->prof_kinds[i]] += ms;
        count_by_kind[ctx->prof_kinds[i]] += 1;
    }
    return VPM_OK;
}

int vpm_profile_get_lb(vpm_ctx* ctx, double* ms_by_mode, int64_t* count_by_mode)
{
    VPM_REQUIRE(ctx && ms_by_mode && count_by_mode, "vpm_profile_get_lb: NULL argument");
    VPM_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < 8; k++) { ms_by_mode[k] = 0.0; count_by_mode[k] = 0; }
    for (size_t i = 0; i < ctx->prof_kinds.size(); i++) {
        if (ctx->prof_kinds[i] != PROF_LB_PASS) continue;
        const int m = ctx->prof_subs[i] & 7;
        float ms = 0.f;
        VPM_CUDA(cudaEventElapsedTime(&ms, ctx->prof_events[2 * i], ctx->prof_events[2 * i + 1]));
        ms_by_mode[m] += ms;
        count_by_mode[m] += 1;
    }
    return VPM_OK;
}

int vpm_ctx_bind_numa(vpm_ctx* ctx, char* cpulist_out, int cap)
{
    VPM_REQUIRE(ctx, "vpm_ctx_bind_numa: ctx is NULL");
    if (cpulist_out && cap > 0) cpulist_out[0] = 0;
    char bus[32] = {0};
    VPM_CUDA(cudaDeviceGetPCIBusId(bus, (int)sizeof(bus), ctx->device));
    for (char* c = bus; *c; c++) *c = (char)tolower(*c);
    const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist";
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return fail(VPM_ERR_UNSUPPORTED, "vpm_ctx_bind_numa: cannot read " + path);
    char buf[1024] = {0};
    const bool got = fgets(buf, sizeof(buf), f) != nullptr;
    fclose(f);
    if (!got) return fail(VPM_ERR_UNSUPPORTED, "vpm_ctx_bind_numa: empty " + path);
    cpu_set_t set;
    CPU_ZERO(&set);
    int ncpu = 0;
    for (const char* c = buf; *c && *c != '\n';) {   // "a-b,c,d-e"
        char* end = nullptr;
        const long a = strtol(c, &end, 10);
        if (end == c) break;
        long b = a;
        c = end;
        if (*c == '-') {
            b = strtol(c + 1, &end, 10);
            c = end;
        }
        for (long k = a; k <= b && k < CPU_SETSIZE; k++) {
            CPU_SET((int)k, &set);
            ncpu++;
        }
        if (*c == ',') c++;
    }
    if (ncpu == 0) return fail(VPM_ERR_UNSUPPORTED, "vpm_ctx_bind_numa: no CPUs listed in " + path);
    if (sched_setaffinity(0, sizeof(set), &set) != 0) return fail(VPM_ERR_UNSUPPORTED, "vpm_ctx_bind_numa: sched_setaffinity failed");
    if (cpulist_out && cap > 0) {
        size_t len = strcspn(buf, "\n");
        if (len >= (size_t)cap) len = (size_t)cap - 1;
        memcpy(cpulist_out, buf, len);
        cpulist_out[len] = 0;
    }
    return VPM_OK;
}

int vpm_host_alloc(int64_t bytes, void** out)
{
    VPM_REQUIRE(out && bytes >= 0, "vpm_host_alloc: bad arguments");
    VPM_CUDA(cudaMallocHost(out, (size_t)(bytes > 0 ? bytes : 1)));
    return VPM_OK;
}

int vpm_host_free(void* p)
{
    if (p) VPM_CUDA(cudaFreeHost(p));
    return VPM_OK;
}

int vpm_dev_alloc(vpm_ctx* ctx, int64_t n_doubles, double** out_dev)
{
    VPM_REQUIRE(ctx && out_dev && n_doubles >= 0, "vpm_dev_alloc: bad arguments");
    VPM_CUDA(cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc((void**)out_dev, sizeof(double) * (size_t)(n_doubles + 2));
    if (e != cudaSuccess) return fail(VPM_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    return VPM_OK;
}

int vpm_dev_free(vpm_ctx* ctx, double* dev)
{
    VPM_REQUIRE(ctx, "vpm_dev_free: ctx is NULL");
    if (!dev) return VPM_OK;
    VPM_CUDA(cudaSetDevice(ctx->device));
    VPM_CUDA(cudaStreamSynchronize(ctx->stream));
    VPM_CUDA(cudaFree(dev));
    return VPM_OK;
}

int vpm_memcpy_h2d(vpm_ctx* ctx, double* dst_dev, const double* src_host, int64_t n_doubles)
{
    VPM_REQUIRE(ctx && (n_doubles == 0 || (dst_dev && src_host)) && n_doubles >= 0, "vpm_memcpy_h2d: bad arguments");
    VPM_CUDA(cudaSetDevice(ctx->device));
    VPM_CUDA(cudaMemcpyAsync(dst_dev, src_host, sizeof(double) * (size_t)n_doubles, cudaMemcpyHostToDevice, ctx->stream));
    VPM_CUDA(cudaStreamSynchronize(ctx->stream));
    return VPM_OK;
}

int vpm_memcpy_d2h(vpm_ctx* ctx, double* dst_host, const double* src_dev, int64_t n_doubles)
{
    VPM_REQUIRE(ctx && (n_doubles == 0 || (dst_host && src_dev)) && n_doubles >= 0, "vpm_memcpy_d2h: bad arguments");
    VPM_CUDA(cudaSetDevice(ctx->device));
    VPM_CUDA(cudaMemcpyAsync(dst_host, src_dev, sizeof(double) * (size_t)n_doubles, cudaMemcpyDeviceToHost, ctx->stream));
    VPM_CUDA(cudaStreamSynchronize(ctx->stream));
    return VPM_OK;
}

/* ---------------------------------------------------------------- particles */

static uint64_t next_generation_seed()
{
    // generation counters of the spaces start at distinct values, so that a space re-created at the address of a destroyed
    // one can never look like the space a carried stagger / projection was recorded against
    static std::atomic<uint64_t> seed{1};
    return seed.fetch_add(1) << 32;
}

// The collision steppers advance a velocity-sorted MIRROR of (v, w) (kernels_lbs.cu) and leave p->v behind (v_stale) until
// somebody needs it in the caller's order: every reader of p->v calls particles_sync_v first, every writer of v or w
// outside those steppers calls mirror_invalidate (which syncs, then drops the mirror).
static int particles_sync_v(vpm_particles* p)
{
    if (!p->v_stale) return VPM_OK;
    VPM_CHECK(launch_lbs_writeback(p->ctx, p->sv, p->sinv, p->v, p->n));
    p->v_stale = false;
    return VPM_OK;
}

static int mirror_invalidate(vpm_particles* p)
{
    VPM_CHECK(particles_sync_v(p));
    p->mirror_valid = false;
    p->stag_valid = false;   // (and the carried half drift of the Strang stepper, which was taken with the old v)
    return VPM_OK;
}



int vpm_particles_create(vpm_ctx* ctx, int64_t n, vpm_particles** out)
{
    VPM_REQUIRE(ctx && out && n >= 0, "vpm_particles_create: bad arguments");
    VPM_CUDA(cudaSetDevice(ctx->device));
    vpm_particles* p = new (std::nothrow) vpm_particles();
    if (!p) return fail(VPM_ERR_NOMEM, "out of host memory");
    p->ctx = ctx;
    p->n = n;
    const size_t bytes = sizeof(double) * (size_t)(n > 0 ? n + (n & 1) : 2);
    cudaError_t e1 = cudaMalloc((void**)&p->x, bytes), e2 = cudaMalloc((void**)&p->v, bytes), e3 = cudaMalloc((void**)&p->w, bytes);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
        cudaFree(p->x); cudaFree(p->v); cudaFree(p->w);
        delete p;
        return fail(VPM_ERR_NOMEM, "cudaMalloc of particle arrays failed");
    }
    cudaMemsetAsync(p->x, 0, bytes, ctx->stream);
    cudaMemsetAsync(p->v, 0, bytes, ctx->stream);
    cudaMemsetAsync(p->w, 0, bytes, ctx->stream);
    *out = p;
    return VPM_OK;
}

int vpm_particles_destroy(vpm_particles* p)
{
    if (!p) return VPM_OK;
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    cudaFree(p->x); cudaFree(p->v); cudaFree(p->w);
    cudaFree(p->q); cudaFree(p->ka); cudaFree(p->kb);
    cudaFree(p->sv); cudaFree(p->sw); cudaFree(p->sinv); cudaFree(p->sort_counts);
    delete p;
    return VPM_OK;
}

int64_t vpm_particles_size(const vpm_particles* p) { return p ? p->n : 0; }

int vpm_particles_ptrs(vpm_particles* p, double** x, double** v, double** w)
{
    VPM_REQUIRE(p, "vpm_particles_ptrs: p is NULL");
    VPM_CUDA(cudaSetDevice(p->ctx->device));
    // the caller may write x, v or w through these pointers at any later time: from now on the velocity-sorted mirror of the
    // collision steppers is rebuilt (and v written back) at every call, and the Strang stepper's stagger is not carried
    if (x || v || w) {
        VPM_CHECK(mirror_invalidate(p));
        p->exposed = true;
    }
    if (x) *x = p->x;
    if (v) *v = p->v;
    if (w) {
        *w = p->w;
        p->uw = false;   // the caller may rewrite the weights through this pointer: the uniform-weight declaration ends here
    }
    return VPM_OK;
}

int vpm_particles_ptrs_const(const vpm_particles* p, const double** x, const double** v, const double** w)
{
    VPM_REQUIRE(p, "vpm_particles_ptrs_const: p is NULL");
    if (v) {   // the array behind the pointer must hold the current velocities (logically const: a cached layout is refreshed)
        VPM_CUDA(cudaSetDevice(p->ctx->device));
        VPM_CHECK(particles_sync_v(const_cast<vpm_particles*>(p)));
    }
    if (x) *x = p->x;
    if (v) *v = p->v;
    if (w) *w = p->w;
    return VPM_OK;
}

int vpm_particles_upload_aos(vpm_particles* p, const double* z, int ld)
{
    VPM_REQUIRE(p && z && (ld == 2 || ld == 3), "vpm_particles_upload_aos: bad arguments (ld must be 2 or 3)");
    vpm_ctx* ctx = p->ctx;
    VPM_CUDA(cudaSetDevice(ctx->device));
    const int64_t chunk = 1 << 24;
    VPM_CHECK(ensure_staging(ctx, (size_t)std::min<int64_t>(chunk, std::max<int64_t>(p->n, 1)) * ld));
    p->v_stale = false;   // v is overwritten as a whole
    p->mirror_valid = false;
    p->stag_valid = false;
    for (int64_t o = 0; o < p->n; o += chunk) {
        const int64_t m = std::min(chunk, p->n - o);
        VPM_CUDA(cudaMemcpyAsync(ctx->staging, z + o * ld, sizeof(double) * m * ld, cudaMemcpyHostToDevice, ctx->stream));
        VPM_CHECK(launch_aos_to_soa(ctx, ctx->staging, ld, m, p->x + o, p->v + o, ld == 3 ? p->w + o : nullptr));
    }
    if (ld == 3) p->uw = false;
    VPM_CUDA(cudaStreamSynchronize(ctx->stream));
    return VPM_OK;
}

int vpm_particles_download_aos(vpm_particles* p, double* z, int ld)
{
    VPM_REQUIRE(p && z && (ld == 2 || ld == 3), "vpm_particles_download_aos: bad arguments (ld must be 2 or 3)");
    vpm_ctx* ctx = p->ctx;
    VPM_CUDA(cudaSetDevice(ctx->device));
    const int64_t chunk = 1 << 24;
    VPM_CHECK(ensure_staging(ctx, (size_t)std::min<int64_t>(chunk, std::max<int64_t>(p->n, 1)) * ld));
    VPM_CHECK(particles_sync_v(p));
    for (int64_t o = 0; o < p->n; o += chunk) {
        const int64_t m = std::min(chunk, p->n - o);
        VPM_CHECK(launch_soa_to_aos(ctx, p->x + o, p->v + o, ld == 3 ? p->w + o : nullptr, ld, m, ctx->staging));
        VPM_CUDA(cudaMemcpyAsync(z + o * ld, ctx->staging, sizeof(double) * m * ld, cudaMemcpyDeviceToHost, ctx->stream));
        VPM_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return VPM_OK;
}

int vpm_particles_upload_soa(vpm_particles* p, const double* x, const double* v, const double* w)
{
    VPM_REQUIRE(p, "vpm_particles_upload_soa: p is NULL");
    vpm_ctx* ctx = p->ctx;
    VPM_CUDA(cudaSetDevice(ctx->device));
    const size_t bytes = sizeof(double) * (size_t)p->n;
    if (v) p->v_stale = false;   // overwritten as a whole
    if (v || w) VPM_CHECK(mirror_invalidate(p));
    if (x) p->stag_valid = false;
    if (x) VPM_CUDA(cudaMemcpyAsync(p->x, x, bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (v) VPM_CUDA(cudaMemcpyAsync(p->v, v, bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (w) {
        VPM_CUDA(cudaMemcpyAsync(p->w, w, bytes, cudaMemcpyHostToDevice, ctx->stream));
        p->uw = false;
    }
    VPM_CUDA(cudaStreamSynchronize(ctx->stream));
    return VPM_OK;
}

int vpm_particles_download_soa(vpm_particles* p, double* x, double* v, double* w)
{
    VPM_REQUIRE(p, "vpm_particles_download_soa: p is NULL");
    vpm_ctx* ctx = p->ctx;
    VPM_CUDA(cudaSetDevice(ctx->device));
    const size_t bytes = sizeof(double) * (size_t)p->n;
    if (v) VPM_CHECK(particles_sync_v(p));
    if (x) VPM_CUDA(cudaMemcpyAsync(x, p->x, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (v) VPM_CUDA(cudaMemcpyAsync(v, p->v, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (w) VPM_CUDA(cudaMemcpyAsync(w, p->w, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    VPM_CUDA(cudaStreamSynchronize(ctx->stream));
    return VPM_OK;
}

int vpm_particles_set_uniform_weight(vpm_particles* p, double w)
{
    VPM_REQUIRE(p && std::isfinite(w), "vpm_particles_set_uniform_weight: bad arguments");
    VPM_CUDA(cudaSetDevice(p->ctx->device));
    VPM_CHECK(mirror_invalidate(p));
    VPM_CHECK(launch_fill(p->ctx, p->w, p->n, w));
    p->uw = true;
    p->wu = w;
    return VPM_OK;
}

int vpm_sample_bump_on_tail(vpm_particles* p, int64_t offset, int64_t ntotal, uint64_t seed, double eps, double kappa,
                            double alpha, double sigma, double v0)
{
    VPM_REQUIRE(p && ntotal > 0 && kappa > 0, "vpm_sample_bump_on_tail: bad arguments");
    VPM_CUDA(cudaSetDevice(p->ctx->device));
    p->uw = false;
    p->v_stale = false;   // x, v and w are overwritten as a whole
    p->mirror_valid = false;
    p->stag_valid = false;
    return launch_sample_bump_on_tail(p->ctx, p, offset, ntotal, seed, eps, kappa, alpha, sigma, v0);
}

int vpm_sample_normal(vpm_particles* p, int64_t offset, int64_t ntotal, uint64_t seed, double xlo, double xhi, double xmax,
                      double* xmax_used)
{
    VPM_REQUIRE(p && ntotal > 0 && xhi > xlo, "vpm_sample_normal: bad arguments");
    VPM_CUDA(cudaSetDevice(p->ctx->device));
    p->uw = false;
    p->v_stale = false;   // x, v and w are overwritten as a whole
    p->mirror_valid = false;
    p->stag_valid = false;
    return launch_sample_normal(p->ctx, p, offset, ntotal, seed, xlo, xhi, xmax, xmax_used);
}

int vpm_sample_maxwellian(vpm_particles* p, int64_t offset, int64_t ntotal, uint64_t seed, double xlo, double xhi,
                          double shift, int doubled, double wnum)
{
    VPM_REQUIRE(p && ntotal > 0, "vpm_sample_maxwellian: bad arguments");
    VPM_CUDA(cudaSetDevice(p->ctx->device));
    p->uw = false;
    p->v_stale = false;   // x, v and w are overwritten as a whole
    p->mirror_valid = false;
    p->stag_valid = false;
    return launch_sample_maxwellian(p->ctx, p, offset, ntotal, seed, xlo, xhi, shift, doubled, wnum);
}

int vpm_sample_uniform(vpm_particles* p, int64_t offset, int64_t ntotal, uint64_t seed, double xlo, double xhi, double vlo,
                       double vhi, double shift, double wnum)
{
    VPM_REQUIRE(p && ntotal > 0, "vpm_sample_uniform: bad arguments");
    VPM_CUDA(cudaSetDevice(p->ctx->device));
    p->uw = false;
    p->v_stale = false;   // x, v and w are overwritten as a whole
    p->mirror_valid = false;
    p->stag_valid = false;
    return launch_sample_uniform(p->ctx, p, offset, ntotal, seed, xlo, xhi, vlo, vhi, shift, wnum);
}

/* ---------------------------------------------------------------- x-space */

int vpm_xspace_create(vpm_ctx* ctx, double lo, double hi, int order, int n_basis, vpm_xspace** out)
{
    VPM_REQUIRE(ctx && out, "vpm_xspace_create: NULL argument");
    VPM_REQUIRE(order >= 2 && order <= kMaxOrder, "vpm_xspace_create: spline order must be 2..6");
    VPM_REQUIRE(n_basis >= 1 && n_basis <= (1 << 20), "vpm_xspace_create: n_basis out of range");
    VPM_REQUIRE(hi > lo && std::isfinite(lo) && std::isfinite(hi), "vpm_xspace_create: need finite lo < hi");
    VPM_CUDA(cudaSetDevice(ctx->device));
    vpm_xspace* xs = new (std::nothrow) vpm_xspace();
    if (!xs) return fail(VPM_ERR_NOMEM, "out of host memory");
    xs->ctx = ctx;
    xs->field_gen = next_generation_seed();
    xs->lo = lo; xs->hi = hi; xs->K = order; xs->nh = n_basis;
    xs->h = (hi - lo) / n_basis;
    xs->invh = 1.0 / xs->h;
    xs->fm = make_fastmod(n_basis);
    periodic_stencils(order, xs->h, xs->mass_stencil, xs->stiff_stencil);
    std::vector<double> mrow, srow, dpiece;
    circulant_first_row(xs->mass_stencil, order, n_basis, mrow);
    circulant_first_row(xs->stiff_stencil, order, n_basis, srow);
    circulant_pinv(srow, true, xs->ginv_host);
    circulant_pinv(mrow, false, xs->minv_host);
    uniform_piece_table(order - 1, dpiece);
    const int ES = (order - 1) | 1;
    int rc = VPM_OK;
    if ((rc = upload(ctx, &xs->ginv, xs->ginv_host)) || (rc = upload(ctx, &xs->minv, xs->minv_host)) ||
        (rc = upload(ctx, &xs->stiff, xs->stiff_stencil)) || (rc = upload(ctx, &xs->dpiece, dpiece))) {
        vpm_xspace_destroy(xs);
        return rc;
    }
    std::vector<double> z1((size_t)n_basis + 2, 0.0), z2((size_t)n_basis, 0.0), z3((size_t)n_basis * ES, 0.0);
    if ((rc = upload(ctx, &xs->rhs, z1)) || (rc = upload(ctx, &xs->phi, z2)) || (rc = upload(ctx, &xs->etab, z3))) {
        vpm_xspace_destroy(xs);
        return rc;
    }
    *out = xs;
    return VPM_OK;
}

int vpm_xspace_destroy(vpm_xspace* xs)
{
    if (!xs) return VPM_OK;
    cudaSetDevice(xs->ctx->device);
    cudaStreamSynchronize(xs->ctx->stream);
    cudaFree(xs->ginv); cudaFree(xs->minv); cudaFree(xs->stiff); cudaFree(xs->dpiece);
    cudaFree(xs->rhs); cudaFree(xs->phi); cudaFree(xs->etab); cudaFree(xs->diag);
    delete xs;
    return VPM_OK;
}

int vpm_xspace_stencils(const vpm_xspace* xs, double* mass, double* stiffness)
{
    VPM_REQUIRE(xs, "vpm_xspace_stencils: xs is NULL");
    const size_t n = 2 * xs->K - 1;
    if (mass) std::memcpy(mass, xs->mass_stencil.data(), n * sizeof(double));
    if (stiffness) std::memcpy(stiffness, xs->stiff_stencil.data(), n * sizeof(double));
    return VPM_OK;
}

static int d2h(vpm_ctx* ctx, double* dst, const double* src, size_t n)
{
    VPM_CUDA(cudaMemcpyAsync(dst, src, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
    VPM_CUDA(cudaStreamSynchronize(ctx->stream));
    return VPM_OK;
}

static int h2d(vpm_ctx* ctx, double* dst, const double* src, size_t n)
{
    VPM_CUDA(cudaMemcpyAsync(dst, src, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    VPM_CUDA(cudaStreamSynchronize(ctx->stream));  // src may be pageable / reused by the caller
    return VPM_OK;
}

int vpm_deposit_x(vpm_xspace* xs, const double* x_dev, const double* w_dev, int64_t n, double* rhs_host)
{
    VPM_REQUIRE(xs && x_dev && w_dev && n >= 0, "vpm_deposit_x: bad arguments");
    vpm_ctx* ctx = xs->ctx;
    VPM_CUDA(cudaSetDevice(ctx->device));
    VpPass p{};
    p.x_in = x_dev; p.w = w_dev; p.n = n; p.flags = VP_DEPOSIT;
    int grid = 0;
    VPM_CHECK(launch_vp_pass(ctx, xs, p, &grid));
    VPM_CHECK(launch_vp_field(ctx, xs, FIELD_REDUCE, grid, 1, 0, -1.0, 1.0, -1, -1));
    if (rhs_host) VPM_CHECK(d2h(ctx, rhs_host, xs->rhs, xs->nh));
    return VPM_OK;
}

int vpm_poisson_solve(vpm_xspace* xs, const double* rhs_host, double* phi_host)
{
    VPM_REQUIRE(xs, "vpm_poisson_solve: xs is NULL");
    vpm_ctx* ctx = xs->ctx;
    VPM_CUDA(cudaSetDevice(ctx->device));
    if (rhs_host) VPM_CHECK(h2d(ctx, xs->rhs, rhs_host, xs->nh));
    // solve only: the rhs is taken as given (already global); without the REDUCE phase nothing is communicated
    VPM_CHECK(launch_vp_field(ctx, xs, FIELD_SOLVE | FIELD_TABLE, 0, 0, 0, -1.0, 1.0, -1, -1));
    if (phi_host) VPM_CHECK(d2h(ctx, phi_host, xs->phi, xs->nh));
    return VPM_OK;
}

int vpm_mass_solve_x(vpm_xspace* xs, const double* rhs_host, double* rho_host)
{
    VPM_REQUIRE(xs && rhs_host && rho_host, "vpm_mass_solve_x: NULL argument");
    vpm_ctx* ctx = xs->ctx;
    VPM_CUDA(cudaSetDevice(ctx->device));
    VPM_CHECK(ensure_red(ctx, 2 * (size_t)xs->nh));
    VPM_CHECK(h2d(ctx, ctx->red, rhs_host, xs->nh));
    VPM_CHECK(launch_circulant_apply(ctx, xs->minv, ctx->red, ctx->red + xs->nh, xs->nh));
    return d2h(ctx, rho_host, ctx->red + xs->nh, xs->nh);
}

int vpm_gather_x(vpm_xspace* xs, const double* coef_host, const double* x_dev, int64_t n, int deriv, double* out_dev)
{
    VPM_REQUIRE(xs && x_dev && out_dev && n >= 0 && (deriv == 0 || deriv == 1), "vpm_gather_x: bad arguments");
    vpm_ctx* ctx = xs->ctx;
    VPM_CUDA(cudaSetDevice(ctx->device));
    const int nh = xs->nh, K = xs->K;
    // scratch in staging: coef (nh) | table (nh * (K|1))
    VPM_CHECK(ensure_staging(ctx, (size_t)nh * (2 + (K | 1))));
    double* coef = ctx->staging;
    double* tab = ctx->staging + nh;
    if (coef_host) VPM_CHECK(h2d(ctx, coef, coef_host, nh));
    else VPM_CUDA(cudaMemcpyAsync(coef, xs->phi, sizeof(double) * nh, cudaMemcpyDeviceToDevice, ctx->stream));
    int nc = 0;
    VPM_CHECK(launch_x_table(ctx, xs, coef, deriv, tab, &nc));
    return launch_x_gather(ctx, xs, tab, nc, x_dev, n, out_dev);
}

int vpm_field_energy(vpm_xspace* xs, const double* phi_host, double* energy)
{
    VPM_REQUIRE(xs && phi_host && energy, "vpm_field_energy: NULL argument");
    vpm_ctx* ctx = xs->ctx;
    VPM_CUDA(cudaSetDevice(ctx->device));
    VPM_CHECK(ensure_red(ctx, (size_t)xs->nh + 2));
    VPM_CHECK(h2d(ctx, ctx->red, phi_host, xs->nh));
    VPM_CHECK(launch_x_energy(ctx, xs, ctx->red, ctx->red + xs->nh));
    return d2h(ctx, energy, ctx->red + xs->nh, 1);
}

int vpm_push_drift(vpm_xspace* xs, vpm_particles* p, double tau)
{
    VPM_REQUIRE(xs && p, "vpm_push_drift: NULL argument");
    VPM_CUDA(cudaSetDevice(xs->ctx->device));
    VPM_CHECK(particles_sync_v(p));   // reads v
    p->stag_valid = false;            // writes x
    VpPass ps{};
    ps.x_in = p->x; ps.v_in = p->v; ps.w = p->w; ps.x_out = p->x; ps.n = p->n;
    ps.flags = VP_POST1 | VP_WRITE_X;
    ps.tau_post1 = tau;
    return launch_vp_pass(xs->ctx, xs, ps, nullptr);
}

int vpm_push_kick(vpm_xspace* xs, vpm_particles* p, const double* phi_host, double tau, double scale)
{
    VPM_REQUIRE(xs && p, "vpm_push_kick: NULL argument");
    vpm_ctx* ctx = xs->ctx;
    VPM_CUDA(cudaSetDevice(ctx->device));
    if (phi_host) VPM_CHECK(h2d(ctx, xs->phi, phi_host, xs->nh));
    VPM_CHECK(launch_vp_field(ctx, xs, FIELD_TABLE, 0, 0, 0, -scale, 1.0, -1, -1));
    VpPass ps{};
    ps.x_in = p->x; ps.v_in = p->v; ps.w = p->w; ps.v_out = p->v; ps.n = p->n;
    VPM_CHECK(mirror_invalidate(p));
    ps.flags = VP_KICK1 | VP_WRITE_V;
    ps.tau_kick = tau;
    return launch_vp_pass(ctx, xs, ps, nullptr);
}

int vpm_update_potential(vpm_xspace* xs, vpm_particles* p, double* rhs_host, double* phi_host)
{
    VPM_REQUIRE(xs && p, "vpm_update_potential: NULL argument");
    vpm_ctx* ctx = xs->ctx;
    VPM_CUDA(cudaSetDevice(ctx->device));
    VpPass ps{};
    ps.x_in = p->x; ps.w = p->w; ps.n = p->n; ps.flags = VP_DEPOSIT;
    int grid = 0;
    VPM_CHECK(launch_vp_pass(ctx, xs, ps, &grid));
    VPM_CHECK(launch_vp_field(ctx, xs, FIELD_REDUCE | FIELD_SOLVE | FIELD_TABLE, grid, 1, 0, -1.0, 1.0, -1, -1));
    if (rhs_host) VPM_CHECK(d2h(ctx, rhs_host, xs->rhs, xs->nh));
    if (phi_host) VPM_CHECK(d2h(ctx, phi_host, xs->phi, xs->nh));
    return VPM_OK;
}

int vpm_xspace_get(vpm_xspace* xs, double* rhs_host, double* phi_host)
{
    VPM_REQUIRE(xs, "vpm_xspace_get: xs is NULL");
    VPM_CUDA(cudaSetDevice(xs->ctx->device));
    if (rhs_host) VPM_CHECK(d2h(xs->ctx, rhs_host, xs->rhs, xs->nh));
    if (phi_host) VPM_CHECK(d2h(xs->ctx, phi_host, xs->phi, xs->nh));
    return VPM_OK;
}

// Strang stepping on SoA arrays (x, v evolve; w and the frozen-field deposit positions are inputs)
static int vp_steps(vpm_xspace* xs, double* x, double* v, const double* w, int64_t n, const double* xdep,
                    const double* wdep, int64_t ndep, double dt, double chi, int nsteps, int mode, int diag_mode,
                    bool uw = false, double wu = 0.0, bool keep_field = false, bool later_leg = false)
{
    vpm_ctx* ctx = xs->ctx;
    const double Dt = dt * chi, escale = -1.0 / (chi * chi), wscale = 1.0 / (chi * chi);
    const int ALL = FIELD_REDUCE | FIELD_SOLVE | FIELD_TABLE;
    int grid = 0;
    if (diag_mode) VPM_CHECK(grow_diag(ctx, &xs->diag, &xs->diag_cap, 3 * ((size_t)nsteps + 2)));

    VpPass ps{};
    ps.x_in = x; ps.v_in = v; ps.w = w; ps.x_out = x; ps.v_out = v; ps.n = n;
    ps.use_uw = uw; ps.w_uniform = wu;

    if (mode == VPM_VP_FROZEN) {
        // field of model.distribution, fixed for the whole run
        // (keep_field: a later leg of one run -- the kick table of the first leg stays in place)
        if (!keep_field) {
            VpPass pd{};
            pd.x_in = xdep; pd.w = wdep; pd.n = ndep; pd.flags = VP_DEPOSIT;
            pd.use_uw = uw; pd.w_uniform = wu;
            VPM_CHECK(launch_vp_pass(ctx, xs, pd, &grid));
            VPM_CHECK(launch_vp_field(ctx, xs, ALL, grid, 1, 0, escale, wscale, diag_mode ? 0 : -1, -1));
        }
        if (diag_mode && !keep_field) {
            VpPass pk = ps;
            pk.flags = VP_DIAG;
            VPM_CHECK(launch_vp_pass(ctx, xs, pk, &grid));
            VPM_CHECK(launch_vp_field(ctx, xs, FIELD_REDUCE, grid, 0, 1, escale, wscale, -1, 0));
        }
        // (K, M are accumulated only when diagnostics were asked for)
        ps.flags = VP_PRE | VP_KICK1 | VP_KICK2 | VP_POST1 | (diag_mode ? VP_DIAG : 0) | VP_WRITE_X | VP_WRITE_V;
        ps.tau_pre = 0.5 * Dt; ps.tau_kick = 0.5 * Dt; ps.tau_post1 = 0.5 * Dt;
        for (int it = 1; it <= nsteps; it++) {
            VPM_CHECK(launch_vp_pass(ctx, xs, ps, &grid));
            if (diag_mode) VPM_CHECK(launch_vp_field(ctx, xs, FIELD_REDUCE, grid, 0, 1, escale, wscale, -1, it));
        }
        return VPM_OK;
    }

    // self-consistent (legacy integrate_vp!): passes are staggered by half a drift so that one pass per
    // step does kick(n) + drift/2 + drift/2 + deposit(n+1)
    if (diag_mode && !later_leg) {   // row 0 of a later leg of one run repeats the previous leg's last row: not recomputed
        VpPass p0 = ps;
        p0.flags = VP_DEPOSIT | VP_DIAG;
        VPM_CHECK(launch_vp_pass(ctx, xs, p0, &grid));
        VPM_CHECK(launch_vp_field(ctx, xs, FIELD_REDUCE | FIELD_SOLVE, grid, 1, 1, escale, wscale, 0, 0));
    }
    if (nsteps <= 0) return VPM_OK;
    {
        VpPass p1 = ps;
        p1.flags = VP_POST2 | VP_DEPOSIT | VP_WRITE_X;
        p1.tau_post2 = 0.5 * Dt;
        VPM_CHECK(launch_vp_pass(ctx, xs, p1, &grid));
        VPM_CHECK(launch_vp_field(ctx, xs, ALL, grid, 1, 0, escale, wscale, diag_mode == 1 ? 1 : -1, -1));
    }
    ps.tau_kick = Dt; ps.tau_post1 = 0.5 * Dt; ps.tau_post2 = 0.5 * Dt;
    for (int it = 1; it <= nsteps; it++) {
        const bool last = it == nsteps;
        if (diag_mode == 2) {
            ps.flags = VP_KICK1 | VP_POST1 | VP_DIAG | VP_DEPOSIT | VP_WRITE_X | VP_WRITE_V;
            VPM_CHECK(launch_vp_pass(ctx, xs, ps, &grid));
            VPM_CHECK(launch_vp_field(ctx, xs, FIELD_REDUCE | FIELD_SOLVE, grid, 1, 1, escale, wscale, it, it));
            if (!last) {
                VpPass p2 = ps;
                p2.flags = VP_POST2 | VP_DEPOSIT | VP_WRITE_X;
                VPM_CHECK(launch_vp_pass(ctx, xs, p2, &grid));
                VPM_CHECK(launch_vp_field(ctx, xs, ALL, grid, 1, 0, escale, wscale, -1, -1));
            }
        } else if (!last) {
            ps.flags = VP_KICK1 | VP_POST1 | (diag_mode ? VP_DIAG : 0) | VP_POST2 | VP_DEPOSIT | VP_WRITE_X | VP_WRITE_V;
            VPM_CHECK(launch_vp_pass(ctx, xs, ps, &grid));
            VPM_CHECK(launch_vp_field(ctx, xs, ALL, grid, 1, diag_mode ? 1 : 0, escale, wscale, diag_mode ? it + 1 : -1, diag_mode ? it : -1));
        } else {
            ps.flags = VP_KICK1 | VP_POST1 | (diag_mode ? VP_DIAG : 0) | VP_WRITE_X | VP_WRITE_V;
            VPM_CHECK(launch_vp_pass(ctx, xs, ps, &grid));
            if (diag_mode) VPM_CHECK(launch_vp_field(ctx, xs, FIELD_REDUCE, grid, 0, 1, escale, wscale, -1, it));
        }
    }
    return VPM_OK;
}

// Self-consistent Strang steps with a CARRIED stagger (no diagnostics).  The fused pass keeps the particles half a drift
// ahead (kick(n) + drift/2 | drift/2 + deposit(n+1) in one launch), so a call used to cost nsteps + 1 passes: a prologue
// (drift/2 + deposit) and an epilogue (kick + drift/2).  Here the LAST pass of a call is a full fused pass as well: it
// stores the caller-visible position (after the trailing half drift) but still deposits at the staggered one, leaving the
// next step's field solved in xs; the FIRST pass of the next call -- same (xs, dt, chi, weights), nobody having touched the
// particles or the field in between -- redoes that half drift in registers (the same fma on the same operands: the same
// bits) and carries on.  Every pass moves 40 B per particle, a call of nsteps steps is nsteps passes, and callers that
// step one step per call pay one pass per step instead of two.  p->x, p->v hold the official state at every call boundary.
static int vp_steps_carry(vpm_xspace* xs, vpm_particles* p, double dt, double chi, int nsteps, bool carried)
{
    vpm_ctx* ctx = xs->ctx;
    const double Dt = dt * chi, escale = -1.0 / (chi * chi), wscale = 1.0 / (chi * chi);
    const int ALL = FIELD_REDUCE | FIELD_SOLVE | FIELD_TABLE;
    const int kFused = VP_KICK1 | VP_POST1 | VP_POST2 | VP_DEPOSIT | VP_WRITE_X | VP_WRITE_V;
    int grid = 0;
    VpPass ps{};
    ps.x_in = p->x; ps.v_in = p->v; ps.w = p->w; ps.x_out = p->x; ps.v_out = p->v; ps.n = p->n;
    ps.use_uw = p->uw; ps.w_uniform = p->wu;
    ps.tau_pre = 0.5 * Dt; ps.tau_kick = Dt; ps.tau_post1 = 0.5 * Dt; ps.tau_post2 = 0.5 * Dt;
    if (!carried) {   // deposit at x + dt/2 v without moving anything (24 B per particle)
        VpPass p1 = ps;
        p1.flags = VP_POST2 | VP_DEPOSIT;
        VPM_CHECK(launch_vp_pass(ctx, xs, p1, &grid));
        VPM_CHECK(launch_vp_field(ctx, xs, ALL, grid, 1, 0, escale, wscale, -1, -1));
    }
    for (int it = 1; it <= nsteps; it++) {
        const bool first = it == 1, last = it == nsteps;
        ps.flags = (first || last) ? (kFused | VP_PRE | VP_WRITE_XU) : kFused;
        ps.rt_pre = first; ps.rt_store_mid = last;
        VPM_CHECK(launch_vp_pass(ctx, xs, ps, &grid));
        VPM_CHECK(launch_vp_field(ctx, xs, ALL, grid, 1, 0, escale, wscale, -1, -1));
    }
    p->stag_valid = true;
    p->stag_xs = xs; p->stag_gen = xs->field_gen;
    p->stag_Dt = Dt; p->stag_chi = chi; p->stag_uw = p->uw; p->stag_wu = p->wu;
    return VPM_OK;
}

static bool vp_carry_enabled()
{
    // VPM_TUNE_VPCARRY=0: every call ends unstaggered (prologue and epilogue passes, as in round 1); read per call: the tests switch it
    if (const char* e = getenv("VPM_TUNE_VPCARRY")) return atoi(e) != 0;
    return true;
}

static bool vp_stag_matches(const vpm_xspace* xs, const vpm_particles* p, double dt, double chi)
{
    return p->stag_valid && p->stag_xs == xs && p->stag_gen == xs->field_gen && p->stag_Dt == dt * chi && p->stag_chi == chi &&
           p->stag_uw == p->uw && (!p->uw || p->stag_wu == p->wu);
}

int vpm_vp_strang_steps_async(vpm_xspace* xs, vpm_particles* p, double dt, double chi, int nsteps, int mode, int diag_mode)
{
    VPM_REQUIRE(xs && p && xs->ctx == p->ctx, "vpm_vp_strang_steps: bad handles");
    VPM_REQUIRE(nsteps >= 0 && chi > 0 && (mode == VPM_VP_SELFCONSISTENT || mode == VPM_VP_FROZEN) && diag_mode >= 0 && diag_mode <= 2,
                "vpm_vp_strang_steps: bad arguments");
    VPM_CUDA(cudaSetDevice(xs->ctx->device));
    const double* xdep = p->x;
    if (mode == VPM_VP_FROZEN) {
        // the deposit positions are the particles' positions at call time; the pass that computes the field
        // runs before any push on the same stream, so no copy is needed
    }
    const bool carried = vp_stag_matches(xs, p, dt, chi);
    VPM_CHECK(mirror_invalidate(p));   // the kick changes v (this also drops the stagger: `carried` was read first)
    if (vp_carry_enabled() && mode == VPM_VP_SELFCONSISTENT && diag_mode == 0 && nsteps >= 1 && !p->exposed)
        return vp_steps_carry(xs, p, dt, chi, nsteps, carried);
    return vp_steps(xs, p->x, p->v, p->w, p->n, xdep, p->w, p->n, dt, chi, nsteps, mode, diag_mode, p->uw, p->wu);
}

int vpm_vp_strang_steps(vpm_xspace* xs, vpm_particles* p, double dt, double chi, int nsteps, int mode, int diag_mode,
                        double* diag_host)
{
    if (!diag_host) diag_mode = 0;
    VPM_CHECK(vpm_vp_strang_steps_async(xs, p, dt, chi, nsteps, mode, diag_mode));
    vpm_ctx* ctx = xs->ctx;
    if (diag_host && diag_mode) {
        VPM_CHECK(d2h(ctx, diag_host, xs->diag, 3 * ((size_t)nsteps + 1)));
        if (mode == VPM_VP_FROZEN)
            for (int it = 1; it <= nsteps; it++) diag_host[3 * it] = diag_host[0];  // the field never changes
        if (mode == VPM_VP_SELFCONSISTENT && diag_mode == 1 && nsteps == 0) { /* row 0 only */ }
    } else {
        VPM_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return p2p_status(ctx);
}

int vpm_vp_strang_step_host(vpm_xspace* xs, vpm_particles* p, const double* z_in, double* z_out, double dt, double chi, int mode)
{
    VPM_REQUIRE(xs && p && z_in && z_out && xs->ctx == p->ctx, "vpm_vp_strang_step_host: bad arguments");
    VPM_REQUIRE(chi > 0 && (mode == VPM_VP_SELFCONSISTENT || mode == VPM_VP_FROZEN), "vpm_vp_strang_step_host: bad arguments");
    vpm_ctx* ctx = xs->ctx;
    VPM_CUDA(cudaSetDevice(ctx->device));
    const int64_t n = p->n, npad = n + (n & 1);
    // staging: z (2n) | x (npad) | v (npad)
    VPM_CHECK(ensure_staging(ctx, (size_t)(2 * n + 2 * npad + 4)));
    double* z = ctx->staging;
    double* x = z + 2 * npad;
    double* v = x + npad;
    // The two PCIe copies dominate (3.2 GB per step at 1e8 particles) and cannot overlap each other: the kick of any
    // particle needs the deposit of all.  What can overlap is the layout work: the state travels in kChunks pieces on a
    // second stream, and the AoS -> SoA pass of piece c runs while piece c + 1 is on the wire (likewise on the way out).
    constexpr int kMaxChunks = 8;
    int kChunks = kMaxChunks;
    if (const char* e = getenv("VPM_TUNE_E2E_CHUNKS")) kChunks = std::min(kMaxChunks, std::max(1, atoi(e)));   // 1: no overlap (A/B)
    if (!ctx->copy_stream) VPM_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    while (ctx->copy_events.size() < 2 * kMaxChunks + 1) {
        cudaEvent_t e;
        VPM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->copy_events.push_back(e);
    }
    const int64_t piece = ((n + kChunks - 1) / kChunks + 1) & ~(int64_t)1;   // even: keeps the 16-byte alignment of every piece
    cudaStream_t cs = ctx->copy_stream;
    // the staging buffer may still be read by work enqueued earlier on the compute stream
    VPM_CUDA(cudaEventRecord(ctx->copy_events[2 * kMaxChunks], ctx->stream));
    VPM_CUDA(cudaStreamWaitEvent(cs, ctx->copy_events[2 * kMaxChunks], 0));
    int c = 0;
    for (int64_t o = 0; o < n; o += piece, c++) {
        const int64_t m = std::min(piece, n - o);
        VPM_CUDA(cudaMemcpyAsync(z + 2 * o, z_in + 2 * o, sizeof(double) * 2 * m, cudaMemcpyHostToDevice, cs));
        VPM_CUDA(cudaEventRecord(ctx->copy_events[c], cs));
        VPM_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->copy_events[c], 0));
        VPM_CHECK(launch_aos_to_soa(ctx, z + 2 * o, 2, m, x + o, v + o, nullptr));
    }
    VPM_CHECK(vp_steps(xs, x, v, p->w, n, p->x, p->w, n, dt, chi, 1, mode, 0, p->uw, p->wu));
    c = 0;
    for (int64_t o = 0; o < n; o += piece, c++) {
        const int64_t m = std::min(piece, n - o);
        VPM_CHECK(launch_soa_to_aos(ctx, x + o, v + o, nullptr, 2, m, z + 2 * o));
        VPM_CUDA(cudaEventRecord(ctx->copy_events[kMaxChunks + c], ctx->stream));
        VPM_CUDA(cudaStreamWaitEvent(cs, ctx->copy_events[kMaxChunks + c], 0));
        VPM_CUDA(cudaMemcpyAsync(z_out + 2 * o, z + 2 * o, sizeof(double) * 2 * m, cudaMemcpyDeviceToHost, cs));
    }
    VPM_CUDA(cudaStreamSynchronize(cs));
    VPM_CUDA(cudaStreamSynchronize(ctx->stream));
    return p2p_status(ctx);
}

/* ---------------------------------------------------------------- v-space */

int vpm_vspace_create(vpm_ctx* ctx, double lo, double hi, int nknots, int order, int dirichlet, vpm_vspace** out)
{
    VPM_REQUIRE(ctx && out, "vpm_vspace_create: NULL argument");
    VPM_REQUIRE(order >= 2 && order <= kMaxOrder, "vpm_vspace_create: spline order must be 2..6");
    VPM_REQUIRE(nknots >= 2 && nknots <= (1 << 16), "vpm_vspace_create: nknots out of range");
    VPM_REQUIRE(hi > lo && std::isfinite(lo) && std::isfinite(hi), "vpm_vspace_create: need finite lo < hi");
    VPM_REQUIRE(!dirichlet || nknots + order - 4 >= 1, "vpm_vspace_create: Dirichlet basis would be empty");
    VPM_CUDA(cudaSetDevice(ctx->device));
    vpm_vspace* vs = new (std::nothrow) vpm_vspace();
    if (!vs) return fail(VPM_ERR_NOMEM, "out of host memory");
    vs->ctx = ctx;
    vs->field_gen = next_generation_seed();
    vs->lo = lo; vs->hi = hi; vs->K = order; vs->nknots = nknots; vs->ncell = nknots - 1;
    vs->h = (hi - lo) / (nknots - 1);
    vs->invh = 1.0 / vs->h;
    vs->nbfull = nknots + order - 2;
    vs->dirichlet = dirichlet ? 1 : 0;
    vs->nv = vs->nbfull - 2 * vs->dirichlet;
    std::vector<double> pieces;
    clamped_piece_table(lo, hi, nknots, order, pieces);
    clamped_mass(pieces, vs->ncell, order, vs->h, vs->dirichlet, vs->mass_host);
    if (banded_cholesky(vs->mass_host, vs->nv, order, vs->chol_host) != 0) {
        delete vs;
        return fail(VPM_ERR_INVALID, "vpm_vspace_create: mass matrix is not positive definite");
    }
    const int TS = 2 * order - 1;
    std::vector<double> z1((size_t)vs->nv + 8, 0.0), z2((size_t)vs->nv, 0.0), z3((size_t)vs->ncell * TS, 0.0), z4(12, 0.0);   // scal: A1, A2 | five moments | - | last diagnostics row of a stepper call (8, 9)
    std::vector<double> z5((size_t)vs->ncell * (2 * order + 2) + 8, 0.0);
    int rc = VPM_OK;
    if (vs->nv <= 64) {
        // dense inverse from the banded factor, column by column (cond(M) ~ 20 on these grids: nothing is lost): the field
        // kernel then solves with one row product per thread instead of a serial chain of 2 nv substitution rows
        const int nv = vs->nv, K = order;
        std::vector<double> inv((size_t)nv * nv, 0.0), y(nv);
        const std::vector<double>& L = vs->chol_host;   // L[i*K + k] = L(i, i-k)
        for (int col = 0; col < nv; col++) {
            for (int i = 0; i < nv; i++) {
                double t = i == col ? 1.0 : 0.0;
                for (int k = 1; k < K && k <= i; k++) t -= L[(size_t)i * K + k] * y[i - k];
                y[i] = t / L[(size_t)i * K];
            }
            for (int i = nv - 1; i >= 0; i--) {
                double t = y[i];
                for (int k = 1; k < K && i + k < nv; k++) t -= L[(size_t)(i + k) * K + k] * y[i + k];
                y[i] = t / L[(size_t)i * K];
            }
            for (int i = 0; i < nv; i++) inv[(size_t)i * nv + col] = y[i];
        }
        if ((rc = upload(ctx, &vs->minv, inv))) {
            vpm_vspace_destroy(vs);
            return rc;
        }
    }
    if ((rc = upload(ctx, &vs->psum, z5)) || (rc = upload(ctx, &vs->pieces, pieces)) || (rc = upload(ctx, &vs->chol, vs->chol_host)) || (rc = upload(ctx, &vs->rhs, z1)) ||
        (rc = upload(ctx, &vs->coef, z2)) || (rc = upload(ctx, &vs->ftab, z3)) || (rc = upload(ctx, &vs->scal, z4))) {
        vpm_vspace_destroy(vs);
        return rc;
    }
    *out = vs;
    return VPM_OK;
}

int vpm_vspace_destroy(vpm_vspace* vs)
{
    if (!vs) return VPM_OK;
    cudaSetDevice(vs->ctx->device);
    cudaStreamSynchronize(vs->ctx->stream);
    cudaFree(vs->pieces); cudaFree(vs->chol); cudaFree(vs->rhs); cudaFree(vs->coef); cudaFree(vs->ftab);
    cudaFree(vs->scal); cudaFree(vs->diag); cudaFree(vs->ent); cudaFree(vs->psum); cudaFree(vs->minv);
    delete vs;
    return VPM_OK;
}

int vpm_vspace_size(const vpm_vspace* vs) { return vs ? vs->nv : 0; }

int vpm_vspace_mass(const vpm_vspace* vs, double* M)
{
    VPM_REQUIRE(vs && M, "vpm_vspace_mass: NULL argument");
    std::memcpy(M, vs->mass_host.data(), sizeof(double) * vs->mass_host.size());
    return VPM_OK;
}

int vpm_deposit_v(vpm_vspace* vs, const double* v_dev, const double* w_dev, int64_t n, double* rhs_host)
{
    VPM_REQUIRE(vs && v_dev && w_dev && n >= 0, "vpm_deposit_v: bad arguments");
    vpm_ctx* ctx = vs->ctx;
    VPM_CUDA(cudaSetDevice(ctx->device));
    LbPass p{};
    p.mode = LB_DEPOSIT_ONLY; p.q = v_dev; p.w = w_dev; p.n = n;
    int grid = 0;
    VPM_CHECK(launch_lb_pass(ctx, vs, p, &grid));
    VPM_CHECK(launch_lb_field(ctx, vs, LBF_REDUCE, grid, 0, -1));
    if (rhs_host) VPM_CHECK(d2h(ctx, rhs_host, vs->rhs, vs->nv));
    return VPM_OK;
}

static int lb_field_local(vpm_ctx* ctx, vpm_vspace* vs, int phases)
{
    // phases without REDUCE / SCALRED never communicate
    return launch_lb_field(ctx, vs, phases & ~(LBF_REDUCE | LBF_SCALRED), 0, 0, -1);
}

int vpm_mass_solve_v(vpm_vspace* vs, const double* rhs_host, double* coef_host)
{
    VPM_REQUIRE(vs, "vpm_mass_solve_v: vs is NULL");
    vpm_ctx* ctx = vs->ctx;
    VPM_CUDA(cudaSetDevice(ctx->device));
    if (rhs_host) VPM_CHECK(h2d(ctx, vs->rhs, rhs_host, vs->nv));
    VPM_CHECK(lb_field_local(ctx, vs, LBF_SOLVE | LBF_TABLE));
    if (coef_host) VPM_CHECK(d2h(ctx, coef_host, vs->coef, vs->nv));
    return VPM_OK;
}

int vpm_project_v(vpm_vspace* vs, const double* v_dev, const double* w_dev, int64_t n, double* coef_host)
{
    VPM_REQUIRE(vs && v_dev && w_dev && n >= 0, "vpm_project_v: bad arguments");
    vpm_ctx* ctx = vs->ctx;
    VPM_CUDA(cudaSetDevice(ctx->device));
    LbPass p{};
    p.mode = LB_DEPOSIT_ONLY; p.q = v_dev; p.w = w_dev; p.n = n;
    int grid = 0;
    VPM_CHECK(launch_lb_pass(ctx, vs, p, &grid));
    VPM_CHECK(launch_lb_field(ctx, vs, LBF_REDUCE | LBF_SOLVE | LBF_TABLE, grid, 0, -1));
    if (coef_host) VPM_CHECK(d2h(ctx, coef_host, vs->coef, vs->nv));
    return VPM_OK;
}

static int set_coef(vpm_vspace* vs, const double* coef_host)
{
    if (!coef_host) return VPM_OK;
    VPM_CHECK(h2d(vs->ctx, vs->coef, coef_host, vs->nv));
    return lb_field_local(vs->ctx, vs, LBF_TABLE);
}

int vpm_resample_v(vpm_vspace* vs, const double* coef_host, vpm_particles* p, int64_t offset, int64_t ntotal, uint64_t seed, int jitter,
                   double* mass_out)
{
    VPM_REQUIRE(vs && p && vs->ctx == p->ctx && ntotal > 0 && offset >= 0, "vpm_resample_v: bad arguments");
    vpm_ctx* ctx = vs->ctx;
    VPM_CUDA(cudaSetDevice(ctx->device));
    VPM_CHECK(set_coef(vs, coef_host));
    p->uw = false;
    p->v_stale = false;   // x, v and w are overwritten as a whole
    p->mirror_valid = false;
    p->stag_valid = false;
    return launch_resample_v(ctx, vs, p, offset, ntotal, seed, jitter, mass_out);
}

int vpm_gather_v(vpm_vspace* vs, const double* coef_host, const double* v_dev, int64_t n, double* f_dev, double* df_dev)
{
    VPM_REQUIRE(vs && v_dev && n >= 0, "vpm_gather_v: bad arguments");
    vpm_ctx* ctx = vs->ctx;
    VPM_CUDA(cudaSetDevice(ctx->device));
    VPM_CHECK(set_coef(vs, coef_host));
    LbPass p{};
    p.mode = LB_EVAL; p.q = v_dev; p.n = n; p.out = f_dev; p.out2 = df_dev;
    return launch_lb_pass(ctx, vs, p, nullptr);
}

int vpm_moments(vpm_vspace* vs, const double* coef_host, const double* v_dev, int64_t n, double* out5_host)
{
    VPM_REQUIRE(vs && v_dev && n >= 0 && out5_host, "vpm_moments: bad arguments");
    vpm_ctx* ctx = vs->ctx;
    VPM_CUDA(cudaSetDevice(ctx->device));
    VPM_CHECK(set_coef(vs, coef_host));
    LbPass p{};
    p.mode = LB_MOMENTS; p.q = v_dev; p.n = n;
    int grid = 0;
    VPM_CHECK(launch_lb_pass(ctx, vs, p, &grid));
    VPM_CHECK(launch_lb_field(ctx, vs, LBF_SCALRED, grid, 5, -1));
    return d2h(ctx, out5_host, vs->rhs + vs->nv, 5);
}

// projection + (CLB: moments + coefficients) for the stage input q; leaves tables and A on the device
static int lb_prepare_clb(vpm_ctx* ctx, vpm_vspace* vs, const double* q, int64_t n)
{
    LbPass pm{};
    pm.mode = LB_MOMENTS; pm.q = q; pm.n = n;
    int grid = 0;
    VPM_CHECK(launch_lb_pass(ctx, vs, pm, &grid));
    return launch_lb_field(ctx, vs, LBF_SCALRED | LBF_COEFF, grid, 5, -1);
}

int vpm_lb_rhs(vpm_vspace* vs, const double* v_dev, const double* w_dev, int64_t n, double nu, int conservative, double* vdot_dev,
               double* coef_host, double* A_host)
{
    VPM_REQUIRE(vs && v_dev && w_dev && vdot_dev && n >= 0, "vpm_lb_rhs: bad arguments");
    vpm_ctx* ctx = vs->ctx;
    VPM_CUDA(cudaSetDevice(ctx->device));
    VPM_CHECK(vpm_project_v(vs, v_dev, w_dev, n, coef_host));
    if (conservative) VPM_CHECK(lb_prepare_clb(ctx, vs, v_dev, n));
    LbPass p{};
    p.mode = LB_RHS_OUT; p.q = v_dev; p.n = n; p.out = vdot_dev; p.nu = nu; p.conservative = conservative;
    VPM_CHECK(launch_lb_pass(ctx, vs, p, nullptr));
    if (A_host) {
        if (conservative) VPM_CHECK(d2h(ctx, A_host, vs->scal, 2));
        else { A_host[0] = 0.0; A_host[1] = 1.0; }
    }
    return VPM_OK;
}

// one LB_ENTROPY pass over (q, w) with the device-side spline of vs; the two sums (S, floored count) are reduced
// (all-reduced across ranks) into row `slot` of the entropy history
static int lb_entropy_row(vpm_ctx* ctx, vpm_vspace* vs, const double* q, const double* w, int64_t n, bool uw, double wu, int slot)
{
    LbPass pe{};
    pe.mode = LB_ENTROPY; pe.q = q; pe.w = w; pe.n = n; pe.f_floor = vs->f_floor;
    pe.use_uw = uw; pe.w_uniform = wu;
    int grid = 0;
    VPM_CHECK(launch_lb_pass(ctx, vs, pe, &grid));
    return launch_lb_field(ctx, vs, LBF_SCALRED | LBF_ENT, grid, 2, slot);
}

int vpm_entropy_v(vpm_vspace* vs, const double* coef_host, const double* v_dev, const double* w_dev, int64_t n, double f_floor,
                  double* S_host, double* nfloored_host)
{
    VPM_REQUIRE(vs && v_dev && w_dev && n >= 0 && S_host && f_floor > 0.0, "vpm_entropy_v: bad arguments");
    vpm_ctx* ctx = vs->ctx;
    VPM_CUDA(cudaSetDevice(ctx->device));
    VPM_CHECK(set_coef(vs, coef_host));
    VPM_CHECK(grow_diag(ctx, &vs->ent, &vs->ent_cap, 2));
    const double keep = vs->f_floor;
    vs->f_floor = f_floor;
    const int rc = lb_entropy_row(ctx, vs, v_dev, w_dev, n, false, 0.0, 0);
    vs->f_floor = keep;
    VPM_CHECK(rc);
    double r[2];
    VPM_CHECK(d2h(ctx, r, vs->ent, 2));
    *S_host = r[0];
    if (nfloored_host) *nfloored_host = r[1];
    return VPM_OK;
}

int vpm_vspace_entropy_history(vpm_vspace* vs, int enable, double f_floor)
{
    VPM_REQUIRE(vs && (!enable || f_floor > 0.0), "vpm_vspace_entropy_history: bad arguments");
    vs->want_entropy = enable != 0;
    if (enable) vs->f_floor = f_floor;
    vs->ent_rows = 0;
    return VPM_OK;
}

int vpm_vspace_entropy_get(vpm_vspace* vs, double* S_host, double* nfloored_host, int rows)
{
    VPM_REQUIRE(vs && S_host && rows >= 0, "vpm_vspace_entropy_get: bad arguments");
    VPM_REQUIRE(rows <= vs->ent_rows, "vpm_vspace_entropy_get: more rows requested than the last stepper call recorded");
    if (rows == 0) return VPM_OK;
    VPM_CUDA(cudaSetDevice(vs->ctx->device));
    std::vector<double> tmp(2 * (size_t)rows);
    VPM_CHECK(d2h(vs->ctx, tmp.data(), vs->ent, tmp.size()));
    for (int i = 0; i < rows; i++) {
        S_host[i] = tmp[2 * i];
        if (nfloored_host) nfloored_host[i] = tmp[2 * i + 1];
    }
    return VPM_OK;
}

static int alloc_scratch(vpm_particles* p)
{
    if (p->q) return VPM_OK;
    const size_t bytes = sizeof(double) * (size_t)(p->n > 0 ? p->n + (p->n & 1) : 2);
    cudaError_t e1 = cudaMalloc((void**)&p->q, bytes), e2 = cudaMalloc((void**)&p->ka, bytes), e3 = cudaMalloc((void**)&p->kb, bytes);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
        cudaFree(p->q); cudaFree(p->ka); cudaFree(p->kb);
        p->q = p->ka = p->kb = nullptr;
        return fail(VPM_ERR_NOMEM, "cudaMalloc of RK438 stage arrays failed");
    }
    return VPM_OK;
}

// Velocity-sorted mirror of (v, w) for the collision steppers (kernels_lbs.cu).  The flow of the collision models cannot
// reorder particles in v, so the mirror is sorted ONCE and stays sorted while only those steppers advance it; it is rebuilt
// when anything else wrote v or w (mirror_valid), when the spline domain that defined the sort keys changed, or -- once
// the caller holds writable device pointers -- at every call.
static constexpr int kSortGridPerSm = 4;

static int lb_sort_mode(const vpm_vspace* vs, const vpm_particles* p)
{
    // VPM_TUNE_LBSORT: 0 = never (private-histogram passes), 1 = default (large ensembles), 2 = always (tests),
    // 3 = always, but the mirror keeps the caller's order (tests: every trip of the sorted passes takes the mixed-cell path)
    // Returns < 0 when no consistent choice exists.
    int mode = 1;
    if (const char* e = getenv("VPM_TUNE_LBSORT")) mode = atoi(e);
    if (mode <= 0 || !lbs_supported(vs->ctx, vs)) return 0;
    // Multi-GPU: every rank must take the same passes (the all-reduce carries power sums on one path, spline right-hand
    // sides on the other), so the choice may only depend on what all ranks share -- the space and the environment -- and not
    // on the size of the local slab: sorted passes at any slab size, including an empty one.
    const bool multi = vs->ctx->p2p.nranks > 1 || vs->ctx->comm.comm != nullptr;
    if (p->n >= ((int64_t)1 << 32)) return multi ? -1 : 0;   // 32-bit sort indices
    if (multi) return mode;
    if (p->n < 1) return 0;
    if (mode == 1 && p->n < ((int64_t)1 << 18)) return 0;   // launch-latency-bound sizes: the sort buys nothing
    return mode;
}

static int ensure_mirror(vpm_vspace* vs, vpm_particles* p, int sort_mode, bool* rebuilt)
{
    *rebuilt = false;
    vpm_ctx* ctx = vs->ctx;
    const bool need_w = !p->uw;
    if (p->mirror_valid && !p->exposed && p->mirror_lo == vs->lo && p->mirror_hi == vs->hi && (p->mirror_has_w || !need_w)) return VPM_OK;
    VPM_CHECK(particles_sync_v(p));   // a rebuild starts from v in the caller's order: bring it up to date first
    *rebuilt = true;
    const size_t bytes = sizeof(double) * (size_t)(p->n > 0 ? p->n + (p->n & 1) : 2);
    const int sort_grid = ctx->sm_count * kSortGridPerSm;
    if (!p->sv) {
        cudaError_t e1 = cudaMalloc((void**)&p->sv, bytes), e2 = cudaMalloc((void**)&p->sinv, sizeof(unsigned) * (size_t)(p->n + 1));
        cudaError_t e3 = cudaMalloc((void**)&p->sort_counts, sizeof(unsigned) * (256 * (size_t)sort_grid + 1));
        if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
            cudaFree(p->sv); cudaFree(p->sinv); cudaFree(p->sort_counts);
            p->sv = nullptr; p->sinv = nullptr; p->sort_counts = nullptr;
            cudaGetLastError();
            return fail(VPM_ERR_NOMEM, "cudaMalloc of the velocity-sorted mirror failed");
        }
    }
    if (need_w && !p->sw) {
        if (cudaMalloc((void**)&p->sw, bytes) != cudaSuccess) {
            p->sw = nullptr;
            cudaGetLastError();
            return fail(VPM_ERR_NOMEM, "cudaMalloc of the velocity-sorted mirror failed");
        }
    }
    // the RK438 stage arrays ka, kb serve as the two key buffers of the sort
    VPM_CHECK(launch_lbs_sort(ctx, p->v, need_w ? p->w : nullptr, p->n, vs->lo, vs->hi, p->ka, p->kb, p->sort_counts, sort_grid, p->sv,
                              need_w ? p->sw : nullptr, p->sinv, sort_mode != 3));
    p->mirror_valid = true;
    p->mirror_has_w = need_w;
    p->mirror_lo = vs->lo;
    p->mirror_hi = vs->hi;
    return VPM_OK;
}

int vpm_lb_rk438_steps_async(vpm_vspace* vs, vpm_particles* p, double nu, double dt, int nsteps, int conservative)
{
    VPM_REQUIRE(vs && p && vs->ctx == p->ctx && nsteps >= 0, "vpm_lb_rk438_steps: bad arguments");
    vpm_ctx* ctx = vs->ctx;
    VPM_CUDA(cudaSetDevice(ctx->device));
    VPM_CHECK(alloc_scratch(p));
    VPM_CHECK(grow_diag(ctx, &vs->diag, &vs->diag_cap, 2 * ((size_t)nsteps + 2)));
    if (nsteps > 0) p->stag_valid = false;   // v changes: a carried half drift of the Strang stepper is void
    const int PROJ = LBF_REDUCE | LBF_SOLVE | LBF_TABLE;
    int grid = 0;
    LbPass ps{};
    ps.n = p->n; ps.w = p->w; ps.nu = nu; ps.dt = dt; ps.conservative = conservative;
    ps.use_uw = p->uw; ps.w_uniform = p->wu;
    ps.ka = p->ka; ps.kb = p->kb;
    // entropy history (opt-in, vpm_vspace_entropy_history): one extra gather pass per row, right after the projection of
    // that row's state; a leg of a longer run (vpm_lb_run) continues at row vs->ent_row0 and skips its row 0
    const bool ent = vs->want_entropy != 0;
    const int erow0 = ent ? vs->ent_row0 : 0;
    if (ent) {
        if (erow0 == 0) VPM_CHECK(grow_diag(ctx, &vs->ent, &vs->ent_cap, 2 * ((size_t)nsteps + 2)));
        else VPM_REQUIRE(vs->ent_cap >= 2 * ((size_t)erow0 + nsteps + 1), "entropy history buffer too small for this leg");
        vs->ent_rows = erow0 + nsteps + 1;
    }
    int sort_mode = lb_sort_mode(vs, p);
    if (sort_mode < 0) return fail(VPM_ERR_UNSUPPORTED, "multi-GPU collision steppers need slabs below 2^32 particles per rank");
    bool rebuilt = false;
    if (sort_mode) {
        const int rc = ensure_mirror(vs, p, sort_mode, &rebuilt);
        // no room for the mirror (+20 B per particle): the histogram passes need none -- but a single rank must not change
        // passes on its own (see lb_sort_mode), so with a communicator attached the error stands
        if (rc == VPM_ERR_NOMEM && ctx->p2p.nranks <= 1 && !ctx->comm.comm) sort_mode = 0;
        else if (rc) return rc;
    }
    if (sort_mode) {
        // ---- velocity-sorted path (kernels_lbs.cu): four passes per step for both models, no histograms, no moments passes.
        // The stage passes deposit per-cell power sums; the field kernel turns them into the right-hand side and, for the
        // conservative model, into the five moments of the freshly solved spline (A1, A2 ready for the next pass).
        const int PS = LBF_PS_REDUCE | LBF_PS_CONVERT | LBF_SOLVE | LBF_TABLE | (conservative ? (LBF_PS_COEFF | LBF_COEFF) : 0);
        ps.w = p->sw;
        // The previous call on this mirror ended with the projection of its final state solved in vs (spline table, A1, A2):
        // if nothing touched the particles (the mirror was not rebuilt) or the space (generation counter) since, the
        // deposit-only pass that re-projects the initial state is skipped and row 0 of the history is that call's last row
        const bool carried = !rebuilt && p->lb_carry_vs == vs && p->lb_carry_gen == vs->field_gen && p->lb_carry_cons == (conservative != 0) &&
                             p->lb_carry_uw == p->uw && (!p->uw || p->lb_carry_wu == p->wu);
        if (carried) {
            VPM_CUDA(cudaMemcpyAsync(vs->diag, vs->scal + 8, 2 * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        } else {   // projection of the initial state + step-0 diagnostics
            LbPass p0 = ps;
            p0.mode = LB_DEPOSIT_ONLY; p0.q = p->sv; p0.diag = 1;
            VPM_CHECK(launch_lbs_pass(ctx, vs, p0, &grid));
            VPM_CHECK(launch_lb_field(ctx, vs, PS, grid, 2, 0, p->uw, p->wu));
        }
        if (ent && erow0 == 0) VPM_CHECK(lb_entropy_row(ctx, vs, p->sv, p->sw, p->n, p->uw, p->wu, 0));
        for (int it = 1; it <= nsteps; it++) {
            for (int s = 1; s <= 4; s++) {
                LbPass st = ps;
                st.mode = LB_STAGE1 + (s - 1);
                st.q = s == 1 ? p->sv : p->q;   // stage input in memory: v (s = 1), q4 (s = 4); stages 2, 3 recompute theirs
                st.v0 = p->sv;
                st.qout = s == 4 ? p->sv : p->q;
                st.diag = s == 4;
                VPM_CHECK(launch_lbs_pass(ctx, vs, st, &grid));
                VPM_CHECK(launch_lb_field(ctx, vs, PS, grid, s == 4 ? 2 : 0, s == 4 ? it : -1, p->uw, p->wu));
            }
            if (ent) VPM_CHECK(lb_entropy_row(ctx, vs, p->sv, p->sw, p->n, p->uw, p->wu, erow0 + it));
        }
        // p->v gets the new velocities back in the caller's order when somebody asks for them (particles_sync_v): a random
        // gather of 1e8 doubles costs as much as a whole RK438 step
        // (a caller holding writable device pointers may look at v at any time: written back at once)
        if (nsteps > 0) {
            p->v_stale = true;
            if (p->exposed) VPM_CHECK(particles_sync_v(p));
        }
        VPM_CUDA(cudaMemcpyAsync(vs->scal + 8, vs->diag + 2 * (size_t)nsteps, 2 * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        p->lb_carry_vs = vs; p->lb_carry_gen = vs->field_gen; p->lb_carry_cons = conservative != 0;
        p->lb_carry_uw = p->uw; p->lb_carry_wu = p->wu;
        return VPM_OK;
    }
    VPM_CHECK(mirror_invalidate(p));   // this path advances v itself
    {   // projection of the initial state + step-0 diagnostics
        LbPass p0 = ps;
        p0.mode = LB_DEPOSIT_ONLY; p0.q = p->v; p0.diag = 1;
        VPM_CHECK(launch_lb_pass(ctx, vs, p0, &grid));
        VPM_CHECK(launch_lb_field(ctx, vs, PROJ | LBF_SCALRED | LBF_DIAG, grid, 2, 0));
        if (ent && erow0 == 0) VPM_CHECK(lb_entropy_row(ctx, vs, p->v, p->w, p->n, p->uw, p->wu, 0));
    }
    for (int it = 1; it <= nsteps; it++) {
        for (int s = 1; s <= 4; s++) {
            // stage input in memory: v (s = 1), q4 (s = 4); q2, q3 only for the conservative model's moments
            // pass -- the stage passes themselves recompute them from v and the stored derivatives
            const double* q = s == 1 ? p->v : p->q;
            if (conservative) VPM_CHECK(lb_prepare_clb(ctx, vs, q, p->n));
            LbPass st = ps;
            st.mode = LB_STAGE1 + (s - 1);
            st.q = q; st.v0 = p->v;
            st.qout = s == 4 ? p->v : ((s == 3 || conservative) ? p->q : nullptr);
            st.diag = s == 4;
            VPM_CHECK(launch_lb_pass(ctx, vs, st, &grid));
            if (s == 4) VPM_CHECK(launch_lb_field(ctx, vs, PROJ | LBF_SCALRED | LBF_DIAG, grid, 2, it));
            else VPM_CHECK(launch_lb_field(ctx, vs, PROJ, grid, 0, -1));
        }
        if (ent) VPM_CHECK(lb_entropy_row(ctx, vs, p->v, p->w, p->n, p->uw, p->wu, erow0 + it));
    }
    return VPM_OK;
}

int vpm_lb_rk438_steps(vpm_vspace* vs, vpm_particles* p, double nu, double dt, int nsteps, int conservative, double* diag_host)
{
    VPM_CHECK(vpm_lb_rk438_steps_async(vs, p, nu, dt, nsteps, conservative));
    if (diag_host) VPM_CHECK(d2h(vs->ctx, diag_host, vs->diag, 2 * ((size_t)nsteps + 1)));
    else VPM_CUDA(cudaStreamSynchronize(vs->ctx->stream));
    return p2p_status(vs->ctx);
}

int vpm_vspace_get(vpm_vspace* vs, double* rhs_host, double* coef_host)
{
    VPM_REQUIRE(vs, "vpm_vspace_get: vs is NULL");
    VPM_CUDA(cudaSetDevice(vs->ctx->device));
    if (rhs_host) VPM_CHECK(d2h(vs->ctx, rhs_host, vs->rhs, vs->nv));
    if (coef_host) VPM_CHECK(d2h(vs->ctx, coef_host, vs->coef, vs->nv));
    return VPM_OK;
}

/* ---------------------------------------------------------------- run! drivers with trajectory output */

}  // extern "C"

namespace {

// Saved frames leave the device without stalling the stepper: the compute stream snapshots the state into one
// of two device buffers (a D2D pass, ~0.5 ms per 1e8 particles), records an event, and carries on with the next
// leg of steps; a second stream copies the snapshot to the host through a small ring of pinned buffers and the
// host thread writes each piece into its place in the HDF5 file (h5min.cpp) while the GPU computes.  The file
// system sets the pace (measured 3-4 GB/s of buffered writes into the page cache or tmpfs against ~50 GB/s of
// PCIe; a pool of 8 writer threads measured no faster, buffered writes to one file serialise on its inode lock),
// so a frame is hidden completely once a leg of steps computes for longer than frame bytes / 3 GB/s.
struct FrameWriter {
    static constexpr int kSlots = 4;
    vpm_ctx* ctx = nullptr;
    vpm_h5* h5 = nullptr;
    int ds_z = -1, ds_t = -1;
    cudaStream_t copy = nullptr;
    cudaEvent_t ready[2] = {nullptr, nullptr};
    double* snap[2] = {nullptr, nullptr};
    size_t frame_doubles = 0, piece = 0;
    double* slot[kSlots] = {};
    cudaEvent_t slot_ev[kSlots] = {};

    int open(vpm_ctx* c, const char* path, int64_t nframes, int64_t n, int nd)
    {
        ctx = c;
        if (n < 1) return fail(VPM_ERR_INVALID, "trajectory output needs at least one particle");
        frame_doubles = (size_t)n * (size_t)nd;
        piece = std::min<size_t>(frame_doubles, (size_t)4 << 20);  // 32 MiB pieces
        VPM_CHECK(vpm_h5_create(path, &h5));
        const int64_t dz3[3] = {nframes, n, nd}, dz2[2] = {nframes, n}, dt1[1] = {nframes};
        VPM_CHECK(nd > 1 ? vpm_h5_add_dataset(h5, "z", 3, dz3, &ds_z) : vpm_h5_add_dataset(h5, "z", 2, dz2, &ds_z));
        VPM_CHECK(vpm_h5_add_dataset(h5, "t", 1, dt1, &ds_t));
        VPM_CHECK(vpm_h5_commit(h5));
        VPM_CUDA(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking));
        for (int b = 0; b < 2; b++) {
            VPM_CUDA(cudaEventCreateWithFlags(&ready[b], cudaEventDisableTiming));
            if (cudaMalloc((void**)&snap[b], frame_doubles * sizeof(double)) != cudaSuccess) {
                cudaGetLastError();
                return fail(VPM_ERR_NOMEM, "cudaMalloc of the trajectory snapshot buffers failed");
            }
        }
        for (int i = 0; i < kSlots; i++) {
            VPM_CUDA(cudaEventCreateWithFlags(&slot_ev[i], cudaEventDisableTiming));
            VPM_CUDA(cudaMallocHost((void**)&slot[i], piece * sizeof(double)));
        }
        return VPM_OK;
    }

    // the snapshot of frame f has been enqueued into snap[f & 1] on the compute stream
    int mark_ready(int64_t f)
    {
        VPM_CUDA(cudaEventRecord(ready[f & 1], ctx->stream));
        return VPM_OK;
    }

    // device -> pinned ring -> file, for frame f; returns when the frame is in the file
    int drain(int64_t f, double t)
    {
        const double* src = snap[f & 1];
        VPM_CUDA(cudaStreamWaitEvent(copy, ready[f & 1], 0));
        const size_t npieces = (frame_doubles + piece - 1) / piece;
        for (size_t i = 0; i < npieces + kSlots; i++) {
            if (i >= (size_t)kSlots) {
                const size_t j = i - kSlots;
                if (j < npieces) {
                    const size_t cnt = std::min(piece, frame_doubles - j * piece);
                    VPM_CUDA(cudaEventSynchronize(slot_ev[j % kSlots]));
                    VPM_CHECK(vpm_h5_write(h5, ds_z, f, (int64_t)(j * piece), (int64_t)cnt, slot[j % kSlots]));
                }
            }
            if (i < npieces) {
                const size_t cnt = std::min(piece, frame_doubles - i * piece);
                VPM_CUDA(cudaMemcpyAsync(slot[i % kSlots], src + i * piece, cnt * sizeof(double), cudaMemcpyDeviceToHost, copy));
                VPM_CUDA(cudaEventRecord(slot_ev[i % kSlots], copy));
            }
        }
        return vpm_h5_write(h5, ds_t, f, 0, 1, &t);
    }

    ~FrameWriter()
    {
        if (copy) cudaStreamSynchronize(copy);
        for (int i = 0; i < kSlots; i++) {
            if (slot[i]) cudaFreeHost(slot[i]);
            if (slot_ev[i]) cudaEventDestroy(slot_ev[i]);
        }
        for (int b = 0; b < 2; b++) {
            if (snap[b]) cudaFree(snap[b]);
            if (ready[b]) cudaEventDestroy(ready[b]);
        }
        if (copy) cudaStreamDestroy(copy);
        if (h5) vpm_h5_close(h5);
    }
};

struct DeviceScratch {
    double* p = nullptr;
    ~DeviceScratch() { if (p) cudaFree(p); }
};

}  // namespace

extern "C" {

int vpm_vp_run(vpm_xspace* xs, vpm_particles* p, double dt, double chi, int nsteps, int mode, int diag_mode, int save_stride,
               const char* h5path, double* diag_host, int* frames_out)
{
    if (frames_out) *frames_out = 0;
    if (!h5path || save_stride <= 0) return vpm_vp_strang_steps(xs, p, dt, chi, nsteps, mode, diag_mode, diag_host);
    VPM_REQUIRE(xs && p && xs->ctx == p->ctx, "vpm_vp_run: bad handles");
    VPM_REQUIRE(nsteps >= 0 && chi > 0 && (mode == VPM_VP_SELFCONSISTENT || mode == VPM_VP_FROZEN) && diag_mode >= 0 && diag_mode <= 2,
                "vpm_vp_run: bad arguments");
    if (!diag_host) diag_mode = 0;
    vpm_ctx* ctx = xs->ctx;
    VPM_CUDA(cudaSetDevice(ctx->device));
    const int64_t nframes = 1 + ((int64_t)nsteps + save_stride - 1) / save_stride;
    FrameWriter fw;
    VPM_CHECK(fw.open(ctx, h5path, nframes, p->n, 2));
    DeviceScratch hist;  // W, K, M rows of the whole run (every leg restarts its own history at row 0)
    if (diag_mode) {
        if (cudaMalloc((void**)&hist.p, sizeof(double) * 3 * ((size_t)nsteps + 1)) != cudaSuccess) {
            cudaGetLastError();
            return fail(VPM_ERR_NOMEM, "cudaMalloc of the diagnostics history failed");
        }
    }
    VPM_CHECK(mirror_invalidate(p));   // the kicks change v
    VPM_CHECK(launch_soa_to_aos(ctx, p->x, p->v, nullptr, 2, p->n, fw.snap[0]));
    VPM_CHECK(fw.mark_ready(0));
    int done = 0;
    if (nsteps == 0 && diag_mode) {
        VPM_CHECK(vp_steps(xs, p->x, p->v, p->w, p->n, p->x, p->w, p->n, dt, chi, 0, mode, diag_mode, p->uw, p->wu, false));
        VPM_CUDA(cudaMemcpyAsync(hist.p, xs->diag, sizeof(double) * 3, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    for (int64_t f = 0; f < nframes; f++) {
        if (done < nsteps) {
            // enqueue the next leg and its snapshot before draining frame f, so the GPU computes while the host writes
            const int leg = std::min(save_stride, nsteps - done);
            if (vp_carry_enabled() && mode == VPM_VP_SELFCONSISTENT && diag_mode == 0 && !p->exposed) {
                // without diagnostics the legs carry the stagger: a saved frame costs its snapshot pass and nothing else
                const bool carried = done > 0 && vp_stag_matches(xs, p, dt, chi);
                p->stag_valid = false;
                VPM_CHECK(vp_steps_carry(xs, p, dt, chi, leg, carried));
            } else
            VPM_CHECK(vp_steps(xs, p->x, p->v, p->w, p->n, p->x, p->w, p->n, dt, chi, leg, mode, diag_mode, p->uw, p->wu, done > 0, done > 0));
            if (diag_mode) {
                const int skip = done > 0 ? 1 : 0;   // row 0 of a later leg repeats the previous leg's last row
                VPM_CUDA(cudaMemcpyAsync(hist.p + 3 * (size_t)(done + skip), xs->diag + 3 * skip,
                                         sizeof(double) * 3 * (size_t)(leg + 1 - skip), cudaMemcpyDeviceToDevice, ctx->stream));
            }
            done += leg;
            VPM_CHECK(launch_soa_to_aos(ctx, p->x, p->v, nullptr, 2, p->n, fw.snap[(f + 1) & 1]));
            VPM_CHECK(fw.mark_ready(f + 1));
        }
        const int64_t step = std::min<int64_t>(f * save_stride, nsteps);
        VPM_CHECK(fw.drain(f, (double)step * dt));
    }
    if (diag_mode) {
        VPM_CHECK(d2h(ctx, diag_host, hist.p, 3 * ((size_t)nsteps + 1)));
        if (mode == VPM_VP_FROZEN)
            for (int it = 1; it <= nsteps; it++) diag_host[3 * it] = diag_host[0];  // the field never changes
    } else {
        VPM_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    if (frames_out) *frames_out = (int)nframes;
    return p2p_status(ctx);
}

int vpm_lb_run(vpm_vspace* vs, vpm_particles* p, double nu, double dt, double t0, int nsteps, int conservative, int save_stride,
               const char* h5path, double* diag_host, int* frames_out)
{
    if (frames_out) *frames_out = 0;
    if (!h5path || save_stride <= 0) return vpm_lb_rk438_steps(vs, p, nu, dt, nsteps, conservative, diag_host);
    VPM_REQUIRE(vs && p && vs->ctx == p->ctx && nsteps >= 0, "vpm_lb_run: bad arguments");
    vpm_ctx* ctx = vs->ctx;
    VPM_CUDA(cudaSetDevice(ctx->device));
    const int64_t nframes = 1 + ((int64_t)nsteps + save_stride - 1) / save_stride;
    FrameWriter fw;
    VPM_CHECK(fw.open(ctx, h5path, nframes, p->n, 1));
    DeviceScratch hist;  // (sum v, sum v^2) rows of the whole run
    if (diag_host) {
        if (cudaMalloc((void**)&hist.p, sizeof(double) * 2 * ((size_t)nsteps + 1)) != cudaSuccess) {
            cudaGetLastError();
            return fail(VPM_ERR_NOMEM, "cudaMalloc of the diagnostics history failed");
        }
    }
    const size_t vbytes = sizeof(double) * (size_t)p->n;
    VPM_CHECK(particles_sync_v(p));
    VPM_CUDA(cudaMemcpyAsync(fw.snap[0], p->v, vbytes, cudaMemcpyDeviceToDevice, ctx->stream));
    VPM_CHECK(fw.mark_ready(0));
    int done = 0;
    if (vs->want_entropy) {   // rows of the whole run; the legs below continue at row `done`
        VPM_CHECK(grow_diag(ctx, &vs->ent, &vs->ent_cap, 2 * ((size_t)nsteps + 2)));
        vs->ent_row0 = 0;
    }
    struct EntRowReset {
        vpm_vspace* v;
        ~EntRowReset() { v->ent_row0 = 0; }
    } ent_reset{vs};
    if (nsteps == 0 && diag_host) {
        VPM_CHECK(vpm_lb_rk438_steps_async(vs, p, nu, dt, 0, conservative));
        VPM_CUDA(cudaMemcpyAsync(hist.p, vs->diag, sizeof(double) * 2, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    for (int64_t f = 0; f < nframes; f++) {
        if (done < nsteps) {
            const int leg = std::min(save_stride, nsteps - done);
            vs->ent_row0 = done;
            VPM_CHECK(vpm_lb_rk438_steps_async(vs, p, nu, dt, leg, conservative));
            if (diag_host) {
                const int skip = done > 0 ? 1 : 0;
                VPM_CUDA(cudaMemcpyAsync(hist.p + 2 * (size_t)(done + skip), vs->diag + 2 * skip,
                                         sizeof(double) * 2 * (size_t)(leg + 1 - skip), cudaMemcpyDeviceToDevice, ctx->stream));
            }
            done += leg;
            // frames are in the caller's particle order: straight from the sorted mirror when that is where the state lives
            if (p->v_stale) VPM_CHECK(launch_lbs_writeback(ctx, p->sv, p->sinv, fw.snap[(f + 1) & 1], p->n));
            else VPM_CUDA(cudaMemcpyAsync(fw.snap[(f + 1) & 1], p->v, vbytes, cudaMemcpyDeviceToDevice, ctx->stream));
            VPM_CHECK(fw.mark_ready(f + 1));
        }
        const int64_t step = std::min<int64_t>(f * save_stride, nsteps);
        VPM_CHECK(fw.drain(f, t0 + (double)step * dt));
    }
    if (diag_host) VPM_CHECK(d2h(ctx, diag_host, hist.p, 2 * ((size_t)nsteps + 1)));
    else VPM_CUDA(cudaStreamSynchronize(ctx->stream));
    if (frames_out) *frames_out = (int)nframes;
    return p2p_status(ctx);
}

/* ---------------------------------------------------------------- host-side operators */

int vpm_galerkin_periodic(double lo, double hi, int order, int n_basis, double* mass_row, double* stiff_row, double* pinv_row)
{
    VPM_REQUIRE(order >= 2 && order <= kMaxOrder && n_basis >= 1 && hi > lo, "vpm_galerkin_periodic: bad arguments");
    std::vector<double> ms, ss, mrow, srow, pinv;
    periodic_stencils(order, (hi - lo) / n_basis, ms, ss);
    circulant_first_row(ms, order, n_basis, mrow);
    circulant_first_row(ss, order, n_basis, srow);
    if (mass_row) std::memcpy(mass_row, mrow.data(), sizeof(double) * n_basis);
    if (stiff_row) std::memcpy(stiff_row, srow.data(), sizeof(double) * n_basis);
    if (pinv_row) {
        circulant_pinv(srow, true, pinv);
        std::memcpy(pinv_row, pinv.data(), sizeof(double) * n_basis);
    }
    return VPM_OK;
}

int vpm_galerkin_clamped(double lo, double hi, int nknots, int order, int dirichlet, int* size, double* M, double* chol_band)
{
    VPM_REQUIRE(order >= 2 && order <= kMaxOrder && nknots >= 2 && hi > lo && size, "vpm_galerkin_clamped: bad arguments");
    const int nv = nknots + order - 2 - (dirichlet ? 2 : 0);
    VPM_REQUIRE(nv >= 1, "vpm_galerkin_clamped: empty basis");
    *size = nv;
    if (!M && !chol_band) return VPM_OK;
    std::vector<double> pieces, mass, chol;
    clamped_piece_table(lo, hi, nknots, order, pieces);
    clamped_mass(pieces, nknots - 1, order, (hi - lo) / (nknots - 1), dirichlet, mass);
    if (M) std::memcpy(M, mass.data(), sizeof(double) * mass.size());
    if (chol_band) {
        if (banded_cholesky(mass, nv, order, chol) != 0) return fail(VPM_ERR_INVALID, "mass matrix is not positive definite");
        std::memcpy(chol_band, chol.data(), sizeof(double) * chol.size());
    }
    return VPM_OK;
}

int vpm_selftest_wrap(int d)
{
    VPM_REQUIRE(d >= 1 && d <= (1 << 20), "vpm_selftest_wrap: divisor out of range");
    const FastMod fm = make_fastmod(d);
    auto wrap = [&](int ci) -> int {   // the device code of splines.cuh, instruction for instruction
        if ((d & (d - 1)) == 0 && (ci & (d - 1)) != (int)(((long long)ci % d + d) % d)) return -1;   // POW2 path
        const uint32_t n = (uint32_t)(ci + fm.bias);
        const uint32_t q = (uint32_t)(((uint64_t)n * fm.magic) >> 32) >> fm.shift;
        const uint32_t r = n - q * fm.d;
        return (int)std::min(r, fm.d - 1u);
    };
    auto ref = [&](long long ci) -> int { long long r = ci % d; return (int)(r < 0 ? r + d : r); };
    const int lim = 1 << 30;
    for (long long ci = -3LL * d - 5; ci <= 3LL * d + 5; ci++)
        if (wrap((int)ci) != ref(ci)) return fail(VPM_ERR_INVALID, "wrap mismatch near zero");
    uint64_t s = 0x9E3779B97F4A7C15ull;
    for (int it = 0; it < 2000000; it++) {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        const int span = lim - d;   // valid index range of the kernels: |ci| <= 2^30 - d
        const int ci = (int)((s >> 33) % (2ull * span + 1)) - span;
        if (wrap(ci) != ref(ci)) return fail(VPM_ERR_INVALID, "wrap mismatch at " + std::to_string(ci));
    }
    for (int ci : {lim - d, -(lim - d), lim - 1 - d, -(lim - 1) + d})
        if (wrap(ci) != ref(ci)) return fail(VPM_ERR_INVALID, "wrap mismatch at the range end");
    return VPM_OK;
}

/* ---------------------------------------------------------------- multi-GPU */

int vpm_comm_unique_id(void* unique_id_128)
{
    VPM_REQUIRE(unique_id_128, "vpm_comm_unique_id: NULL argument");
    Nccl* api = nccl_api();
    if (!api) return fail(VPM_ERR_COMM, "libnccl.so.2 could not be loaded");
    NcclUniqueId id;
    int r = api->GetUniqueId(&id);
    if (r != 0) return fail(VPM_ERR_COMM, "ncclGetUniqueId failed");
    std::memcpy(unique_id_128, &id, sizeof(id));
    return VPM_OK;
}

int vpm_comm_init(vpm_ctx* ctx, int nranks, int rank, const void* unique_id_128)
{
    VPM_REQUIRE(ctx && unique_id_128 && nranks >= 1 && rank >= 0 && rank < nranks, "vpm_comm_init: bad arguments");
    Nccl* api = nccl_api();
    if (!api) return fail(VPM_ERR_COMM, "libnccl.so.2 could not be loaded");
    VPM_CUDA(cudaSetDevice(ctx->device));
    vpm_comm_destroy(ctx);
    NcclUniqueId id;
    std::memcpy(&id, unique_id_128, sizeof(id));
    void* comm = nullptr;
    int r = api->CommInitRank(&comm, nranks, id, rank);
    if (r != 0) return fail(VPM_ERR_COMM, std::string("ncclCommInitRank: ") + (api->GetErrorString ? api->GetErrorString(r) : "error"));
    ctx->comm.api = api;
    ctx->comm.comm = comm;
    ctx->comm.nranks = nranks;
    ctx->comm.rank = rank;
    return VPM_OK;
}

int vpm_comm_destroy(vpm_ctx* ctx)
{
    if (!ctx || !ctx->comm.comm) return VPM_OK;
    cudaStreamSynchronize(ctx->stream);
    ctx->comm.api->CommDestroy(ctx->comm.comm);
    ctx->comm = vpm::Comm{};
    return VPM_OK;
}

/* ---- fused peer-memory all-reduce (p2p.cuh) ---- */

int vpm_p2p_prepare(vpm_ctx* ctx, void* ipc_handle_64)
{
    VPM_REQUIRE(ctx && ipc_handle_64, "vpm_p2p_prepare: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    VPM_CUDA(cudaSetDevice(ctx->device));
    vpm_p2p_detach(ctx);
    cudaError_t e = cudaMalloc((void**)&ctx->p2p_local, sizeof(P2PMailbox));
    if (e != cudaSuccess) return fail(VPM_ERR_NOMEM, std::string("cudaMalloc(mailbox): ") + cudaGetErrorString(e));
    VPM_CUDA(cudaMemset(ctx->p2p_local, 0, sizeof(P2PMailbox)));
    VPM_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    VPM_CUDA(cudaIpcGetMemHandle(&h, ctx->p2p_local));
    std::memcpy(ipc_handle_64, &h, sizeof(h));
    return VPM_OK;
}

int vpm_p2p_attach(vpm_ctx* ctx, int nranks, int rank, const void* ipc_handles)
{
    VPM_REQUIRE(ctx && ipc_handles && ctx->p2p_local, "vpm_p2p_attach: call vpm_p2p_prepare first");
    VPM_REQUIRE(nranks >= 1 && nranks <= kP2PMaxRanks && rank >= 0 && rank < nranks, "vpm_p2p_attach: 1..8 ranks supported");
    VPM_CUDA(cudaSetDevice(ctx->device));
    P2PDev d{};
    d.nranks = nranks;
    d.rank = rank;
    for (int r = 0; r < nranks; r++) {
        if (r == rank) {
            d.mbox[r] = ctx->p2p_local;
            continue;
        }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, (const char*)ipc_handles + 64 * (size_t)r, sizeof(h));
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            for (int q = 0; q < r; q++)
                if (ctx->p2p_opened[q]) { cudaIpcCloseMemHandle(ctx->p2p_opened[q]); ctx->p2p_opened[q] = nullptr; }
            return fail(VPM_ERR_COMM, std::string("cudaIpcOpenMemHandle (peer access over NVLink unavailable?): ") + cudaGetErrorString(e));
        }
        ctx->p2p_opened[r] = ptr;
        d.mbox[r] = (P2PMailbox*)ptr;
    }
    {
        double ms = 20000.0;   // ~20 s at 2 GHz
        if (const char* e = getenv("VPM_P2P_TIMEOUT_MS")) ms = std::max(1.0, atof(e));
        d.timeout_cycles = (long long)(ms * 2.0e6);
    }
    ctx->p2p = d;
    ctx->p2p_seq = 0;
    return VPM_OK;
}

int vpm_p2p_detach(vpm_ctx* ctx)
{
    if (!ctx) return VPM_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int r = 0; r < kP2PMaxRanks; r++)
        if (ctx->p2p_opened[r]) { cudaIpcCloseMemHandle(ctx->p2p_opened[r]); ctx->p2p_opened[r] = nullptr; }
    if (ctx->p2p_local) cudaFree(ctx->p2p_local);
    ctx->p2p_local = nullptr;
    ctx->p2p = P2PDev{};
    ctx->p2p_seq = 0;
    return VPM_OK;
}

int vpm_p2p_error(vpm_ctx* ctx, uint64_t* failed_seq)
{
    VPM_REQUIRE(ctx && failed_seq, "vpm_p2p_error: NULL argument");
    *failed_seq = 0;
    if (!ctx->p2p_local) return VPM_OK;
    VPM_CUDA(cudaSetDevice(ctx->device));
    VPM_CUDA(cudaStreamSynchronize(ctx->stream));
    unsigned long long e = 0;
    VPM_CUDA(cudaMemcpy(&e, &ctx->p2p_local->error, sizeof(e), cudaMemcpyDeviceToHost));
    *failed_seq = e;
    if (e) return fail(VPM_ERR_COMM, "peer-memory all-reduce timed out waiting for a rank (sequence " + std::to_string(e) + ")");
    return VPM_OK;
}

int vpm_comm_allreduce(vpm_ctx* ctx, double* buf_dev, int64_t count)
{
    VPM_REQUIRE(ctx && buf_dev && count >= 0, "vpm_comm_allreduce: bad arguments");
    VPM_CUDA(cudaSetDevice(ctx->device));
    return comm_allreduce(ctx, buf_dev, (size_t)count);
}

}  // extern "C"
