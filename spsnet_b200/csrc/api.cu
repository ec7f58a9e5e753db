// api.cu -- error plumbing and library identity for libspsk.so.
#include "common.cuh"
#include <atomic>

namespace spsk {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return SPSK_ERR_CUDA;
}

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// fp16 range guard: one word per device, OR-ed by the tensor-core kernels with the caller's tag bit whenever a value they
// store as fp16 exceeds the fp16 range (spsk_fp16_overflow_poll reads it)
__device__ unsigned int g_fp16_overflow = 0u;

unsigned int *fp16_overflow_word() {
    static unsigned int *ptr[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (!ptr[dev]) {
        void *p = nullptr;
        if (cudaGetSymbolAddress(&p, g_fp16_overflow) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
        ptr[dev] = static_cast<unsigned int *>(p);
    }
    return ptr[dev];
}

}  // namespace spsk

extern "C" unsigned long long spsk_launch_count(void) { return spsk::g_launches.load(std::memory_order_relaxed); }
extern "C" const char *spsk_last_error(void) { return spsk::g_err; }
extern "C" int spsk_abi_version(void) { return 5; }  // 4: batch-statistics pass of spsk_sa_mma_forward (stats, stats_parts) + spsk_sa_mma_stats_parts; 5: spsk_sa_pack_layer, spsk_bn_stats_reduce / _finalize, spsk_sa_mma_schedule

extern "C" int spsk_fp16_overflow_poll(unsigned int *mask, int clear) {
    using namespace spsk;
    SPSK_REQUIRE(mask, SPSK_ERR_INVALID_ARG, "fp16_overflow_poll: null output");
    unsigned int *w = fp16_overflow_word();
    SPSK_REQUIRE(w, SPSK_ERR_CUDA, "fp16_overflow_poll: cannot resolve the flag word");
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(mask, w, sizeof(unsigned int), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && clear && *mask) e = cudaMemset(w, 0, sizeof(unsigned int));
    if (e != cudaSuccess) return cuda_fail(e, "fp16_overflow_poll");
    return SPSK_OK;
}
extern "C" int spsk_built_for_sm(void) { return 100; }
