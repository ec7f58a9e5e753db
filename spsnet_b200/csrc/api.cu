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

}  // namespace spsk

extern "C" unsigned long long spsk_launch_count(void) { return spsk::g_launches.load(std::memory_order_relaxed); }
extern "C" const char *spsk_last_error(void) { return spsk::g_err; }
extern "C" int spsk_abi_version(void) { return 2; }  // 2: sections 3 (iou3d / NMS / detect), 4 (edge conv) and spsk_scatter_grad added
extern "C" int spsk_built_for_sm(void) { return 100; }
