// Host-side plumbing shared by runtime.cu and capi.cu (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <functional>
#include <mutex>
#include <vector>

#include "runtime.h"

namespace ldpc {

// Mirror of the LABRADOR_LDPC_ERR_* values in include/labrador_ldpc.h.
constexpr int LDPC_OK = 0;
constexpr int LDPC_ERR_BAD_CODE = -1;
constexpr int LDPC_ERR_NULL_POINTER = -2;
constexpr int LDPC_ERR_CUDA = -3;
constexpr int LDPC_ERR_MIXED_POINTERS = -4;
constexpr int LDPC_ERR_BAD_ARGUMENT = -5;

// One frame-major array of a batched call.  host_in != nullptr: copied to the
// device before the kernel; host_out != nullptr: copied back after it.
struct HostArray {
    const void *host_in;
    void *host_out;
    size_t bytes_per_frame;
};

// Launches the kernel(s) of one chunk: device pointers in the order of the
// HostArray list, `frames` frames, on `stream`.
typedef std::function<cudaError_t(DeviceCtx &, const std::vector<void *> &, size_t frames, cudaStream_t)> BatchLaunch;
// Same, for kernels whose result depends on the absolute frame index (counter-based generators):
// `first` is the index, within the whole batch, of the chunk's first frame.
typedef std::function<cudaError_t(DeviceCtx &, const std::vector<void *> &, size_t frames, cudaStream_t, size_t first)>
    BatchLaunchAt;

int runtime_init(const int *devices, int n_devices);
void runtime_shutdown();
int runtime_device_count();
int get_ctx(int device, DeviceCtx **out, std::mutex **mu_out);
int classify_pointer(const void *p, int *device);
int run_host_batch_at(const std::vector<HostArray> &arrays, size_t batch, const BatchLaunchAt &launch);
inline int run_host_batch(const std::vector<HostArray> &arrays, size_t batch, const BatchLaunch &launch) {
    return run_host_batch_at(arrays, batch, [&launch](DeviceCtx &ctx, const std::vector<void *> &d, size_t nf,
                                                      cudaStream_t st, size_t) { return launch(ctx, d, nf, st); });
}
int run_device_batch(int device, cudaStream_t stream, bool synchronize,
                     const std::function<cudaError_t(DeviceCtx &, cudaStream_t)> &launch);

unsigned long long launch_count();
const char *last_error();
int fail(int code, const char *msg);

}  // namespace ldpc
