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

// ---- neighbourhood synchronisation of the fused row-sharded forward (keynet_b200/dist.py) -------------------------------
// After a layer's SpMM has stored its rows into the peers' activation buffers, a rank SIGNALS the ranks that depend on it and
// WAITS only for the ranks it depends on (its spatial neighbours for conv / pool layers), instead of a barrier over all ranks:
// flags[r][p] (int32, one array per rank in NVLink-mapped memory) holds the last epoch rank p signalled to rank r.
namespace {
struct KnSyncArgs { int32_t *flags[8]; int world, me; unsigned signal_mask, wait_mask; int32_t *epoch_counter; int *timeout_flag; };

__global__ void peer_sync_kernel(const __grid_constant__ KnSyncArgs a) {
    const int t = threadIdx.x;
    // the epoch lives in device memory and advances by one per call: every rank makes the same sequence of calls, so the
    // counters agree, and a CUDA graph that captured this launch can be replayed (an epoch passed by value would repeat)
    __shared__ int s_epoch;
    if (t == 0) { s_epoch = *a.epoch_counter + 1; *a.epoch_counter = s_epoch; }
    __syncthreads();
    const int epoch = s_epoch;
    __threadfence_system();                       // the preceding kernels' peer stores are performed before the flags are
    if (t < a.world && ((a.signal_mask >> t) & 1u)) {
        int32_t *dst = a.flags[t] + a.me;
        asm volatile("st.release.sys.global.s32 [%0], %1;" :: "l"(dst), "r"(epoch) : "memory");
    }
    if (t < a.world && ((a.wait_mask >> t) & 1u)) {
        const int32_t *src = a.flags[a.me] + t;
        int v;
        long long spins = 0;
        do {
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
            if (v >= epoch) break;
            __nanosleep(64);
        } while (++spins < (1ll << 24));          // ~seconds: a peer that never arrives must not hang the GPU
        if (v < epoch && a.timeout_flag) atomicExch(a.timeout_flag, 1);
    }
    __syncthreads();
    __threadfence_system();
}
}  // namespace

KN_API int kn_peer_sync(const uint64_t *flags_host, int32_t world, int32_t my_rank, uint32_t signal_mask, uint32_t wait_mask, int32_t *epoch_counter,
                        int32_t *timeout_flag, void *stream) {
    KN_REQUIRE(epoch_counter != nullptr, "peer_sync: null epoch counter");
    KN_REQUIRE(flags_host != nullptr && world >= 1 && world <= 8 && my_rank >= 0 && my_rank < world, "peer_sync: bad arguments (world=%d rank=%d)", world, my_rank);
    KnSyncArgs a;
    for (int i = 0; i < 8; i++) a.flags[i] = (i < world) ? reinterpret_cast<int32_t *>(flags_host[i]) : nullptr;
    a.world = world; a.me = my_rank; a.signal_mask = signal_mask & ~(1u << my_rank); a.wait_mask = wait_mask & ~(1u << my_rank); a.epoch_counter = epoch_counter;
    a.timeout_flag = timeout_flag;
    peer_sync_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(a);
    KN_CHECK_LAUNCH();
    return KN_OK;
}
