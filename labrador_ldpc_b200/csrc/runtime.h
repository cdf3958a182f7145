// Internal interfaces of the B200 LDPC library: per-device context, device
// tables and the kernel launchers.  Nothing here crosses the C ABI.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "code_tables.h"

namespace ldpc {

// Device ordinals the per-kernel "already configured" tables are sized for; contexts for larger ordinals are refused.
constexpr int kMaxDevices = 64;

enum LlrType : int { kI8 = 0, kI16 = 1, kI32 = 2, kF32 = 3, kF64 = 4, kNumLlrTypes = 5 };

inline size_t llr_size(int t) {
    switch (t) {
        case kI8: return 1;
        case kI16: return 2;
        case kI32: return 4;
        case kF32: return 4;
        case kF64: return 8;
        default: return 0;
    }
}

// Input front end of decode_ms (see front.cuh): what the `llrs` pointer of a launch holds.
enum FrontKind : int { kFrontNone = 0, kFrontSoftF32 = 1, kFrontHard = 2 };
struct Front {
    int kind = kFrontNone;
    float scale = 1.0f;   // kFrontSoftF32: llr = clamp(rint(soft * scale), -limit, limit)
    float limit = 0.0f;
};
inline size_t front_frame_bytes(const Front &f, int n, int llr_type) {
    if (f.kind == kFrontSoftF32) return (size_t)n * 4;
    if (f.kind == kFrontHard) return (size_t)n / 8;
    return (size_t)n * llr_size(llr_type);
}
bool front_supported(int kind, int llr_type);   // fused kernels exist for (soft, i8|i16) and (hard, i8)

// Device-resident tables of one code (built once per device by DeviceCtx).
struct DeviceCode {
    int n, k, p, m, b;
    int edges, checks, vars;
    int max_var_degree, max_check_degree;
    int n_blocks;
    const uint64_t *var_tab;   // [max_var_degree][vars]   idx | check << 32
    const uint64_t *chk_tab;   // [max_check_degree][checks] idx | var << 32
    const uint64_t *gen;       // compact generator rows (nullptr for the k = 16384 codes, which the reference does not encode)
    const uint32_t *gen32;     // same rows as big-endian-ordered 32-bit words
    const uint32_t *enc_ainv;  // TM codes: first columns of the circulants of A^-1 (code_tables.h: tm_encoder_table), else null
    const uint32_t *enc_tc_lut; // TC codes: byte (TC128, TC256) / nibble (TC512) table of parity contributions, else null
    const uint32_t *enc_lut;   // TM codes: the same as a nibble lookup table (code_tables.h: tm_encoder_lut), else null
};

// Everything one host thread needs to push a host-pointer batch through a device: its own streams, device chunk
// buffers and pinned staging block.  Lanes live in a pool of the device context (DeviceCtx::lanes), so concurrent
// callers -- the reference's usage is one decoder per host thread, perftest/src/main.rs:39-45 -- never share a
// stream or a buffer and only meet at the context mutex for the few microseconds of a kernel launch.
struct HostLane {
    static constexpr int kPipe = 3;
    static constexpr size_t kSmallBytes = 256 << 10;
    cudaStream_t stream[kPipe] = {nullptr, nullptr, nullptr};
    void *buf[kPipe] = {nullptr, nullptr, nullptr};
    size_t bytes[kPipe] = {0, 0, 0};
    // small calls (the single-codeword reference API): one pinned, device-mapped staging block
    void *small_host = nullptr;   // host address
    void *small_dev = nullptr;    // the same block as the device sees it
    // one work counter per stream of the lane: launches on a lane's stream are ordered, so its counter needs neither the
    // context's ring nor events (class WorkCounter picks it up through set_lane_counter)
    unsigned long long *counters = nullptr;     // [kPipe][16]
};

// The calling thread's next launch runs on a lane stream that owns `counter` (nullptr: use the context's ring).
void set_lane_counter(unsigned long long *counter);

struct DeviceCtx {
    int device = -1;
    int sm_count = 0;
    int max_smem_optin = 0;       // bytes of dynamic shared memory a CTA may opt in to
    DeviceCode codes[kNumCodes];
    void *table_blob = nullptr;   // one allocation holding every table
    // scratch for decode paths whose message array does not fit in shared memory
    void *vscratch = nullptr;
    size_t vscratch_bytes = 0;
    // event that orders the reuse of vscratch across streams (decode_ms_generic.cu)
    cudaEvent_t vscratch_done = nullptr;
    // work counters of the persistent kernels (class WorkCounter below): one slot per launch, reused round-robin;
    // `counter_done[s]` is recorded behind the kernel that used slot s and awaited by the next user of the slot
    static constexpr int kCounterSlots = 64;
    static constexpr int kCounterStride = 16;   // 128 bytes apart: concurrent kernels never share a line
    unsigned long long *counters = nullptr;
    cudaEvent_t counter_done[kCounterSlots] = {};
    int counter_next = 0;
    // list of undecided frames of the two-pass TC bit-flipping decoder (decode_bf_tc.cu); `retry_done` orders
    // its reuse across streams
    unsigned *retry_list = nullptr;
    size_t retry_list_bytes = 0;
    cudaEvent_t retry_done = nullptr;
    // host-pointer pipeline: pool of per-caller lanes (guarded by the context mutex)
    std::vector<std::unique_ptr<HostLane>> lanes;
    std::vector<HostLane *> lanes_free;
};

// A zeroed 8-byte frame-claim counter for ONE launch of a persistent kernel.  Construct it (under the context mutex,
// like every launcher) before the launch and let it go out of scope after the launch: the destructor records the
// slot's event on the stream, and the constructor makes the stream wait for the slot's previous user, so a slot is
// never zeroed or claimed from while an earlier kernel -- on any stream -- still uses it.
class WorkCounter {
public:
    WorkCounter(DeviceCtx &ctx, cudaStream_t stream);
    ~WorkCounter();
    WorkCounter(const WorkCounter &) = delete;
    WorkCounter &operator=(const WorkCounter &) = delete;
    cudaError_t error() const { return err_; }
    unsigned long long *ptr() const { return ptr_; }

private:
    DeviceCtx &ctx_;
    cudaStream_t stream_;
    int slot_ = -1;
    unsigned long long *ptr_ = nullptr;
    cudaError_t err_ = cudaSuccess;
};

// Global launch counter (every kernel launch of this library increments it).
void count_launch(int n = 1);

// ---- kernel launchers (device pointers, stream-ordered, no synchronisation) ----
cudaError_t launch_decode_ms(DeviceCtx &ctx, int code, int llr_type, const void *llrs, uint8_t *output,
                             size_t batch, size_t max_iters, uint8_t *success, uint32_t *iters,
                             cudaStream_t stream, const Front &front = Front());
const char *decode_ms_kernel_name(int code, int llr_type);

cudaError_t launch_decode_bf(DeviceCtx &ctx, int code, const uint8_t *input, uint8_t *output, size_t batch,
                             size_t max_iters, uint8_t *success, uint32_t *iters, cudaStream_t stream);

// codewords: [batch][n/8]; if data != nullptr it is [batch][k/8] and is copied in first.
cudaError_t launch_encode(DeviceCtx &ctx, int code, const uint8_t *data, uint8_t *codewords, size_t batch,
                          cudaStream_t stream);

cudaError_t launch_hard_to_llrs(DeviceCtx &ctx, int code, int llr_type, const uint8_t *input, void *llrs,
                                size_t batch, cudaStream_t stream);
// llrs[i] = clamp(rint(soft[i] * scale), -limit, limit) as i8 / i16 (the un-fused form of kFrontSoftF32).
cudaError_t launch_quantise(DeviceCtx &ctx, int code, int llr_type, const float *soft, void *llrs, size_t batch,
                            float scale, float limit, cudaStream_t stream);
cudaError_t launch_llrs_to_hard(DeviceCtx &ctx, int code, int llr_type, const void *llrs, uint8_t *output,
                                size_t batch, cudaStream_t stream);

// ---- harness kernels (channel.cu): counter-based frame generator and error counter ----
cudaError_t launch_random_data(DeviceCtx &ctx, int code, unsigned long long seed, unsigned long long first_frame,
                               uint8_t *data, size_t batch, cudaStream_t stream);
// out_type kF32: out = y * scale;  kI8 / kI16: out = clamp(rint(y * scale), -limit, limit);  y = (1 - 2 bit) + sigma * z
cudaError_t launch_awgn(DeviceCtx &ctx, int code, int out_type, const uint8_t *codewords, float sigma, float scale,
                        float limit, unsigned long long seed, unsigned long long first_frame, void *out, size_t batch,
                        cudaStream_t stream);
cudaError_t launch_count_errors(DeviceCtx &ctx, int code, const uint8_t *decoded, const uint8_t *data, uint32_t *errors,
                                size_t batch, cudaStream_t stream);

}  // namespace ldpc
