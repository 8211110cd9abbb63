// abi.cu -- error reporting, version and device facts of the C ABI (include/keynet_b200.h).
#include "common.cuh"
#include <string.h>

static thread_local char g_err[512] = "";

void kn_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

KN_API int kn_abi_version(void) { return KN_ABI_VERSION; }

KN_API const char *kn_last_error(void) { return g_err; }

int kn_sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

KN_API int kn_device_info(int *sm_count, int *cc_major, int *cc_minor, int64_t *total_mem_bytes) {
    int dev = 0;
    KN_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    KN_CUDA(cudaGetDeviceProperties(&p, dev));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (total_mem_bytes) *total_mem_bytes = (int64_t)p.totalGlobalMem;
    return KN_OK;
}

static thread_local KnPeers g_peers = {0, {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}, nullptr};

KnPeers kn_current_peers() { return g_peers; }

KnPeersScope::KnPeersScope(const kn_peers *p) : saved(g_peers), ok(true) {
    KnPeers v = {0, {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}, nullptr};
    if (p != nullptr && p->n != 0) {
        if (p->n < 0 || p->n > 8) { kn_set_error("peers: between 0 and 8 destinations (got %d)", (int)p->n); ok = false; return; }
        v.n = p->n;
        for (int i = 0; i < p->n; i++) {
            if (p->y[i] == 0) { kn_set_error("peers: destination %d is null", i); ok = false; return; }
            v.y[i] = reinterpret_cast<float *>(p->y[i]);
        }
        v.row_mask = p->row_mask;
    }
    g_peers = v;
}

KnPeersScope::~KnPeersScope() { g_peers = saved; }
