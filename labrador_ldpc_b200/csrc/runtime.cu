// Device contexts, device-resident tables, kernel dispatch and the
// host-buffer streaming pipeline behind the `_batch` C entry points.
#include "runtime.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "host_api.h"

namespace ldpc {

// implemented in decode_ms_generic.cu / decode_ms_tm.cu
cudaError_t launch_decode_ms_generic(DeviceCtx &ctx, int code, int llr_type, const void *llrs, uint8_t *output,
                                     size_t batch, size_t max_iters, uint8_t *success, uint32_t *iters,
                                     cudaStream_t stream, const Front &front);

bool launch_decode_ms_tm_i8(DeviceCtx &ctx, int code, const void *llrs, uint8_t *output, size_t batch,
                            size_t max_iters, uint8_t *success, uint32_t *iters, cudaStream_t stream,
                            cudaError_t *err, const Front &front);
bool has_decode_ms_tm_i8(int code);
// decode_ms_tm_i16.cu: i16 on the TM codes, 32-bit variable side + packed 16-bit check side
bool launch_decode_ms_tm_i16(DeviceCtx &ctx, int code, const void *llrs, uint8_t *output, size_t batch, size_t max_iters,
                             uint8_t *success, uint32_t *iters, cudaStream_t stream, cudaError_t *err, const Front &front);
bool has_decode_ms_tm_i16(int code);
// decode_ms_tm_cluster.cu: i8 on the k = 16384 codes, one codeword per cluster of four CTAs
bool launch_decode_ms_tm_cluster(DeviceCtx &ctx, int code, const void *llrs, uint8_t *output, size_t batch, size_t max_iters,
                                 uint8_t *success, uint32_t *iters, cudaStream_t stream, cudaError_t *err, const Front &front);
bool has_decode_ms_tm_cluster(int code);
bool launch_decode_ms_tm_wide(DeviceCtx &ctx, int code, int llr_type, const void *llrs, uint8_t *output, size_t batch,
                              size_t max_iters, uint8_t *success, uint32_t *iters, cudaStream_t stream,
                              cudaError_t *err, const Front &front);
bool has_decode_ms_tm_wide(int code, int llr_type);
bool launch_decode_ms_tc(DeviceCtx &ctx, int code, int llr_type, const void *llrs, uint8_t *output, size_t batch,
                         size_t max_iters, uint8_t *success, uint32_t *iters, cudaStream_t stream, cudaError_t *err,
                         const Front &front);
bool has_decode_ms_tc(int code);
// decode_ms_tc_x2.cu: i8 on the TC codes, two codewords per register
bool launch_decode_ms_tc_x2(DeviceCtx &ctx, int code, const void *llrs, uint8_t *output, size_t batch, size_t max_iters,
                            uint8_t *success, uint32_t *iters, cudaStream_t stream, cudaError_t *err, const Front &front);
bool tc_x2_enabled();

namespace {

// LABRADOR_LDPC_FORCE_GENERIC=1 routes everything through the generic kernel (A/B testing only).
bool force_generic() {
    static const bool v = [] { const char *e = getenv("LABRADOR_LDPC_FORCE_GENERIC"); return e && e[0] == '1'; }();
    return v;
}

std::atomic<unsigned long long> g_launches{0};
thread_local std::string t_last_error;

struct Runtime {
    std::mutex mu;
    bool inited = false;
    std::vector<int> devices;                          // devices used for host-pointer batches
    std::vector<std::unique_ptr<DeviceCtx>> ctxs;      // every context created so far
    std::vector<std::unique_ptr<std::mutex>> ctx_mu;   // serialises pipeline use per context
};

Runtime &rt() {
    static Runtime r;
    return r;
}

int set_error(int code, const std::string &msg) {
    t_last_error = msg;
    return code;
}

int cuda_error(cudaError_t e, const char *what) {
    return set_error(LDPC_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define CUDA_TRY(expr)                                    \
    do {                                                  \
        cudaError_t e__ = (expr);                         \
        if (e__ != cudaSuccess) return cuda_error(e__, #expr); \
    } while (0)

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Builds every device table of every code in one blob (one cudaMalloc + one copy).
int build_ctx(int device, std::unique_ptr<DeviceCtx> &out) {
    if (device < 0 || device >= kMaxDevices) return set_error(LDPC_ERR_BAD_ARGUMENT, "device ordinal out of range");
    CUDA_TRY(cudaSetDevice(device));
    std::unique_ptr<DeviceCtx> ctx(new DeviceCtx());
    ctx->device = device;
    CUDA_TRY(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device));
    CUDA_TRY(cudaDeviceGetAttribute(&ctx->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));

    std::vector<unsigned char> blob;
    struct Off { size_t var_tab, chk_tab, gen, gen32, ainv, lut, tc_lut; bool has_ainv, has_tc_lut; };
    Off offs[kNumCodes];
    bool has_lut_of[kNumCodes] = {};
    for (int ci = 0; ci < kNumCodes; ci++) {
        const CodeInfo &c = *code_info(ci);
        std::vector<uint64_t> vt, ct;
        build_ell_tables(c, vt, ct);
        auto append = [&blob](const void *p, size_t bytes) {
            const size_t off = align_up(blob.size(), 256);
            blob.resize(off + bytes);
            memcpy(blob.data() + off, p, bytes);
            return off;
        };
        offs[ci].var_tab = append(vt.data(), vt.size() * 8);
        offs[ci].chk_tab = append(ct.data(), ct.size() * 8);
        const size_t gwords = c.gen ? (size_t)(c.k / c.b) * ((c.n - c.k) / 64) : 0;
        offs[ci].gen = append(c.gen, gwords * 8);
        std::vector<uint32_t> g32(gwords * 2);
        for (size_t i = 0; i < gwords; i++) {
            g32[2 * i] = (uint32_t)(c.gen[i] >> 32);
            g32[2 * i + 1] = (uint32_t)(c.gen[i] & 0xFFFFFFFFu);
        }
        offs[ci].gen32 = append(g32.data(), g32.size() * 4);
        std::vector<uint32_t> ainv;
        offs[ci].has_ainv = tm_encoder_table(ci, ainv);
        offs[ci].ainv = offs[ci].has_ainv ? append(ainv.data(), ainv.size() * 4) : 0;
        std::vector<uint32_t> tc_lut;
        offs[ci].has_tc_lut = c.gen && ((ci == 1 || ci == 2) ? tc_rot_encoder_lut(ci, tc_lut) : tc_encoder_lut(ci, tc_encoder_group_bits(ci), tc_lut));
        offs[ci].tc_lut = offs[ci].has_tc_lut ? append(tc_lut.data(), tc_lut.size() * 4) : 0;
        std::vector<uint32_t> lut;
        // the nibble lookup table is 512 rows of M bits: it fits the shared memory of a CTA up to M = 2048
        const bool has_lut = offs[ci].has_ainv && c.m <= 2048 && tm_encoder_lut(ci, lut);
        has_lut_of[ci] = has_lut;
        offs[ci].lut = has_lut ? append(lut.data(), lut.size() * 4) : 0;
    }
    CUDA_TRY(cudaMalloc(&ctx->table_blob, blob.size()));
    CUDA_TRY(cudaMemcpy(ctx->table_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    const unsigned char *base = static_cast<const unsigned char *>(ctx->table_blob);
    for (int ci = 0; ci < kNumCodes; ci++) {
        const CodeInfo &c = *code_info(ci);
        DeviceCode &d = ctx->codes[ci];
        d.n = c.n; d.k = c.k; d.p = c.p; d.m = c.m; d.b = c.b;
        d.edges = c.edges; d.checks = c.checks; d.vars = c.vars;
        d.max_var_degree = c.max_var_degree; d.max_check_degree = c.max_check_degree;
        d.n_blocks = c.n_blocks;
        d.var_tab = reinterpret_cast<const uint64_t *>(base + offs[ci].var_tab);
        d.chk_tab = reinterpret_cast<const uint64_t *>(base + offs[ci].chk_tab);
        d.gen = c.gen ? reinterpret_cast<const uint64_t *>(base + offs[ci].gen) : nullptr;
        d.gen32 = c.gen ? reinterpret_cast<const uint32_t *>(base + offs[ci].gen32) : nullptr;
        d.enc_ainv = offs[ci].has_ainv ? reinterpret_cast<const uint32_t *>(base + offs[ci].ainv) : nullptr;
        d.enc_tc_lut = offs[ci].has_tc_lut ? reinterpret_cast<const uint32_t *>(base + offs[ci].tc_lut) : nullptr;
        d.enc_lut = has_lut_of[ci] ? reinterpret_cast<const uint32_t *>(base + offs[ci].lut) : nullptr;
    }
    out = std::move(ctx);
    return LDPC_OK;
}

// Caller holds rt().mu.
int ctx_for_device_locked(int device, DeviceCtx **out, std::mutex **mu_out) {
    Runtime &r = rt();
    for (size_t i = 0; i < r.ctxs.size(); i++)
        if (r.ctxs[i]->device == device) {
            *out = r.ctxs[i].get();
            if (mu_out) *mu_out = r.ctx_mu[i].get();
            return LDPC_OK;
        }
    std::unique_ptr<DeviceCtx> ctx;
    int prev = -1;
    cudaGetDevice(&prev);
    const int rc = build_ctx(device, ctx);
    if (prev >= 0) cudaSetDevice(prev);
    if (rc != LDPC_OK) return rc;
    r.ctxs.push_back(std::move(ctx));
    r.ctx_mu.emplace_back(new std::mutex());
    *out = r.ctxs.back().get();
    if (mu_out) *mu_out = r.ctx_mu.back().get();
    return LDPC_OK;
}

}  // namespace

namespace {
thread_local unsigned long long *t_lane_counter = nullptr;
}
void set_lane_counter(unsigned long long *counter) { t_lane_counter = counter; }

WorkCounter::WorkCounter(DeviceCtx &ctx, cudaStream_t stream) : ctx_(ctx), stream_(stream) {
    if (t_lane_counter) {      // host-pointer path: the lane's stream owns this counter, stream order is all it needs
        ptr_ = t_lane_counter;
        err_ = cudaMemsetAsync(ptr_, 0, sizeof(unsigned long long), stream);
        return;
    }
    if (!ctx.counters) {
        err_ = cudaMalloc(&ctx.counters, sizeof(unsigned long long) * DeviceCtx::kCounterStride * DeviceCtx::kCounterSlots);
        if (err_ != cudaSuccess) return;
    }
    const int s = ctx.counter_next;
    ctx.counter_next = (s + 1) % DeviceCtx::kCounterSlots;
    if (ctx.counter_done[s]) {
        err_ = cudaStreamWaitEvent(stream, ctx.counter_done[s], 0);      // the slot's previous user, whatever its stream
    } else {
        err_ = cudaEventCreateWithFlags(&ctx.counter_done[s], cudaEventDisableTiming);
    }
    if (err_ != cudaSuccess) return;
    ptr_ = ctx.counters + (size_t)s * DeviceCtx::kCounterStride;
    err_ = cudaMemsetAsync(ptr_, 0, sizeof(unsigned long long), stream);
    slot_ = s;
}

WorkCounter::~WorkCounter() {
    if (slot_ >= 0) cudaEventRecord(ctx_.counter_done[slot_], stream_);
}

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
unsigned long long launch_count() { return g_launches.load(std::memory_order_relaxed); }
const char *last_error() { return t_last_error.c_str(); }
int fail(int code, const char *msg) { return set_error(code, msg); }

int runtime_init(const int *devices, int n_devices) {
    Runtime &r = rt();
    std::lock_guard<std::mutex> lock(r.mu);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess) return cuda_error(e, "cudaGetDeviceCount");
    if (count == 0) return set_error(LDPC_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    std::vector<int> want;
    if (devices == nullptr || n_devices <= 0) {
        if (r.inited) return LDPC_OK;
        int cur = 0;
        CUDA_TRY(cudaGetDevice(&cur));
        want.push_back(cur);
    } else {
        for (int i = 0; i < n_devices; i++) {
            if (devices[i] < 0 || devices[i] >= count)
                return set_error(LDPC_ERR_BAD_ARGUMENT, "device index out of range");
            want.push_back(devices[i]);
        }
    }
    for (int d : want) {
        DeviceCtx *ctx = nullptr;
        const int rc = ctx_for_device_locked(d, &ctx, nullptr);
        if (rc != LDPC_OK) return rc;
    }
    r.devices = want;
    r.inited = true;
    return LDPC_OK;
}

void runtime_shutdown() {
    Runtime &r = rt();
    std::lock_guard<std::mutex> lock(r.mu);
    int prev = -1;
    cudaGetDevice(&prev);
    for (auto &c : r.ctxs) {
        cudaSetDevice(c->device);
        for (auto &lane : c->lanes) {
            for (int i = 0; i < HostLane::kPipe; i++) {
                if (lane->stream[i]) { cudaStreamSynchronize(lane->stream[i]); cudaStreamDestroy(lane->stream[i]); }
                if (lane->buf[i]) cudaFree(lane->buf[i]);
            }
            if (lane->small_host) cudaFreeHost(lane->small_host);
            if (lane->counters) cudaFree(lane->counters);
        }
        c->lanes.clear();
        c->lanes_free.clear();
        if (c->vscratch) cudaFree(c->vscratch);
        if (c->vscratch_done) { cudaEventSynchronize(c->vscratch_done); cudaEventDestroy(c->vscratch_done); }
        for (int i = 0; i < DeviceCtx::kCounterSlots; i++)
            if (c->counter_done[i]) { cudaEventSynchronize(c->counter_done[i]); cudaEventDestroy(c->counter_done[i]); }
        if (c->counters) cudaFree(c->counters);
        if (c->retry_done) { cudaEventSynchronize(c->retry_done); cudaEventDestroy(c->retry_done); }
        if (c->retry_list) cudaFree(c->retry_list);
        if (c->table_blob) cudaFree(c->table_blob);
    }
    r.ctxs.clear();
    r.ctx_mu.clear();
    r.devices.clear();
    r.inited = false;
    if (prev >= 0) cudaSetDevice(prev);
}

int runtime_device_count() {
    Runtime &r = rt();
    std::lock_guard<std::mutex> lock(r.mu);
    return (int)r.devices.size();
}

int get_ctx(int device, DeviceCtx **out, std::mutex **mu_out) {
    int rc = runtime_init(nullptr, 0);
    if (rc != LDPC_OK) return rc;
    Runtime &r = rt();
    std::lock_guard<std::mutex> lock(r.mu);
    return ctx_for_device_locked(device, out, mu_out);
}

// 0 = host (pageable or pinned), 1 = device/managed; *device receives the owning device.
int classify_pointer(const void *p, int *device) {
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    if (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) {
        if (device) *device = attr.device;
        return 1;
    }
    return 0;
}

// ---------------------------------------------------------------------------
// kernel dispatch
// ---------------------------------------------------------------------------
cudaError_t launch_decode_ms(DeviceCtx &ctx, int code, int llr_type, const void *llrs, uint8_t *output,
                             size_t batch, size_t max_iters, uint8_t *success, uint32_t *iters,
                             cudaStream_t stream, const Front &front) {
    if (batch == 0) return cudaSuccess;
    if (front.kind != kFrontNone && !front_supported(front.kind, llr_type)) return cudaErrorInvalidValue;
    if (llr_type == kI8 && !force_generic()) {
        cudaError_t err = cudaSuccess;
        if (launch_decode_ms_tm_i8(ctx, code, llrs, output, batch, max_iters, success, iters, stream, &err, front))
            return err;
        if (launch_decode_ms_tm_cluster(ctx, code, llrs, output, batch, max_iters, success, iters, stream, &err, front))
            return err;
    }
    if (!force_generic()) {
        cudaError_t err = cudaSuccess;
        if (llr_type == kI16 &&
            launch_decode_ms_tm_i16(ctx, code, llrs, output, batch, max_iters, success, iters, stream, &err, front))
            return err;
        if (launch_decode_ms_tm_wide(ctx, code, llr_type, llrs, output, batch, max_iters, success, iters, stream, &err,
                                     front))
            return err;
        if (llr_type == kI8 &&
            launch_decode_ms_tc_x2(ctx, code, llrs, output, batch, max_iters, success, iters, stream, &err, front))
            return err;
        if (launch_decode_ms_tc(ctx, code, llr_type, llrs, output, batch, max_iters, success, iters, stream, &err, front))
            return err;
    }
    return launch_decode_ms_generic(ctx, code, llr_type, llrs, output, batch, max_iters, success, iters, stream, front);
}

bool front_supported(int kind, int llr_type) {
    if (kind == kFrontNone) return llr_type >= 0 && llr_type < kNumLlrTypes;
    if (kind == kFrontSoftF32) return llr_type == kI8 || llr_type == kI16;
    if (kind == kFrontHard) return llr_type == kI8;
    return false;
}

const char *decode_ms_kernel_name(int code, int llr_type) {
    if (llr_type == kI8 && has_decode_ms_tm_i8(code) && !force_generic()) return "ms_tm_s16x2<i8>";
    if (llr_type == kI8 && has_decode_ms_tm_cluster(code) && !force_generic()) return "ms_tm_cluster<i8>";
    if (llr_type == kI16 && has_decode_ms_tm_i16(code) && !force_generic()) return "ms_tm_s16x2<i16>";
    if (has_decode_ms_tm_wide(code, llr_type) && !force_generic()) {
        static const char *wide[kNumLlrTypes] = {"ms_tm_wide<i8>", "ms_tm_wide<i16>", "ms_tm_wide<i32>", "ms_tm_wide<f32>", "ms_tm_wide<f64>"};
        return wide[llr_type];
    }
    if (has_decode_ms_tc(code) && !force_generic() && llr_type == kI8 && tc_x2_enabled()) return "ms_tc_x2<i8>";
    if (has_decode_ms_tc(code) && !force_generic() && llr_type >= 0 && llr_type < kNumLlrTypes) {
        static const char *tc[kNumLlrTypes] = {"ms_tc_warp<i8>", "ms_tc_warp<i16>", "ms_tc_warp<i32>", "ms_tc_warp<f32>", "ms_tc_warp<f64>"};
        return tc[llr_type];
    }
    static const char *names[kNumLlrTypes] = {"ms_generic<i8>", "ms_generic<i16>", "ms_generic<i32>",
                                              "ms_generic<f32>", "ms_generic<f64>"};
    if (llr_type < 0 || llr_type >= kNumLlrTypes) return "invalid";
    return names[llr_type];
}

// ---------------------------------------------------------------------------
// Host-buffer pipeline: frames are streamed through the GPU in chunks; chunk c
// uses buffer/stream c % 3, so the H2D copy of one chunk, the kernel of the
// previous one and the D2H copy of the one before overlap.
// ---------------------------------------------------------------------------
namespace {

size_t chunk_bytes_target() {
    static size_t v = [] {
        const char *e = getenv("LABRADOR_LDPC_CHUNK_MB");
        size_t mb = e ? (size_t)atol(e) : 16;      // 16 MB: shortest fill and drain of the three-slot pipeline that still keeps PCIe busy (profiles/raw/r02d_chunk_sizes.txt)
        if (mb < 1) mb = 1;
        return mb << 20;
    }();
    return v;
}

// A lane of the context's pool for the duration of one call (RAII; the pool is guarded by the context mutex).
class LaneLease {
public:
    LaneLease(DeviceCtx &ctx, std::mutex &mu) : ctx_(ctx), mu_(mu) {
        std::lock_guard<std::mutex> lock(mu_);
        if (!ctx_.lanes_free.empty()) {
            lane_ = ctx_.lanes_free.back();
            ctx_.lanes_free.pop_back();
            return;
        }
        std::unique_ptr<HostLane> lane(new HostLane());
        for (int i = 0; i < HostLane::kPipe; i++) {
            err_ = cudaStreamCreateWithFlags(&lane->stream[i], cudaStreamNonBlocking);
            if (err_ != cudaSuccess) return;
        }
        err_ = cudaMalloc(&lane->counters, sizeof(unsigned long long) * 16 * HostLane::kPipe);
        if (err_ != cudaSuccess) return;
        lane_ = lane.get();
        ctx_.lanes.push_back(std::move(lane));
    }
    ~LaneLease() {
        if (!lane_) return;
        std::lock_guard<std::mutex> lock(mu_);
        ctx_.lanes_free.push_back(lane_);
    }
    HostLane *lane() const { return lane_; }
    cudaError_t error() const { return err_; }

private:
    DeviceCtx &ctx_;
    std::mutex &mu_;
    HostLane *lane_ = nullptr;
    cudaError_t err_ = cudaSuccess;
};

// Launchers touch per-context state (work counters, scratch): they run under the context mutex.  Copies and
// synchronisation do not, so concurrent callers overlap everything but the launch call itself.
cudaError_t locked_launch(DeviceCtx &ctx, std::mutex &mu, const BatchLaunchAt &launch, const std::vector<void *> &dptr,
                          size_t frames, cudaStream_t st, size_t first, unsigned long long *lane_counter) {
    std::lock_guard<std::mutex> lock(mu);
    set_lane_counter(lane_counter);
    const cudaError_t e = launch(ctx, dptr, frames, st, first);
    set_lane_counter(nullptr);
    return e;
}

int run_on_device_host_ptrs(DeviceCtx &ctx, std::mutex &mu, const std::vector<HostArray> &arrays, size_t first,
                            size_t count, const BatchLaunchAt &launch) {
    CUDA_TRY(cudaSetDevice(ctx.device));
    size_t per_frame = 0;
    for (const HostArray &a : arrays) per_frame += align_up(a.bytes_per_frame, 16);
    if (per_frame == 0 || count == 0) return LDPC_OK;
    LaneLease lease(ctx, mu);
    if (lease.error() != cudaSuccess || !lease.lane()) return cuda_error(lease.error(), "stream create");
    HostLane &lane = *lease.lane();
    // Small calls (the reference's single-codeword API is a batch of one): the arrays are packed into one pinned,
    // device-mapped block; the kernel reads its input and writes its results through that mapping, so the call is two
    // host memcpys, one launch and one stream synchronisation instead of a staged copy per array (profiles/r01_latency.md).
    static const bool small_path = [] { const char *e = getenv("LABRADOR_LDPC_SMALL_CALLS"); return !e || atoi(e) != 0; }();
    if (small_path && per_frame * count + 256 * arrays.size() <= HostLane::kSmallBytes) {
        if (!lane.small_host) {
            CUDA_TRY(cudaHostAlloc(&lane.small_host, HostLane::kSmallBytes, cudaHostAllocMapped | cudaHostAllocPortable));
            CUDA_TRY(cudaHostGetDevicePointer(&lane.small_dev, lane.small_host, 0));
        }
        std::vector<void *> dptr(arrays.size());
        std::vector<size_t> off(arrays.size());
        size_t total = 0;
        for (size_t i = 0; i < arrays.size(); i++) {
            off[i] = total;
            total += align_up(arrays[i].bytes_per_frame * count, 256);
            dptr[i] = static_cast<unsigned char *>(lane.small_dev) + off[i];
            if (arrays[i].host_in)
                memcpy(static_cast<unsigned char *>(lane.small_host) + off[i],
                       static_cast<const unsigned char *>(arrays[i].host_in) + first * arrays[i].bytes_per_frame,
                       arrays[i].bytes_per_frame * count);
        }
        cudaStream_t st = lane.stream[0];
        cudaError_t e = locked_launch(ctx, mu, launch, dptr, count, st, first, lane.counters);
        if (e != cudaSuccess) return cuda_error(e, "kernel launch");
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return cuda_error(e, "stream synchronize");
        for (size_t i = 0; i < arrays.size(); i++)
            if (arrays[i].host_out)
                memcpy(static_cast<unsigned char *>(arrays[i].host_out) + first * arrays[i].bytes_per_frame,
                       static_cast<unsigned char *>(lane.small_host) + off[i], arrays[i].bytes_per_frame * count);
        return LDPC_OK;
    }
    size_t chunk = chunk_bytes_target() / per_frame;
    const size_t min_chunk = (size_t)ctx.sm_count * 8;
    if (chunk < min_chunk) chunk = min_chunk;
    if (chunk > count) chunk = count;
    // carve one device buffer per pipeline slot
    std::vector<size_t> offs(arrays.size());
    size_t total = 0;
    for (size_t i = 0; i < arrays.size(); i++) {
        offs[i] = total;
        total += align_up(arrays[i].bytes_per_frame * chunk, 256);
    }
    const size_t n_chunks = (count + chunk - 1) / chunk;
    const int slots = (int)(n_chunks < (size_t)HostLane::kPipe ? n_chunks : (size_t)HostLane::kPipe);
    for (int s = 0; s < slots; s++) {
        if (lane.bytes[s] < total) {
            if (lane.buf[s]) { CUDA_TRY(cudaStreamSynchronize(lane.stream[s])); CUDA_TRY(cudaFree(lane.buf[s])); }
            lane.buf[s] = nullptr; lane.bytes[s] = 0;
            CUDA_TRY(cudaMalloc(&lane.buf[s], total));
            lane.bytes[s] = total;
        }
    }
    int rc = LDPC_OK;
    for (size_t c = 0; c < n_chunks && rc == LDPC_OK; c++) {
        const int s = (int)(c % HostLane::kPipe);
        cudaStream_t st = lane.stream[s];
        unsigned char *base = static_cast<unsigned char *>(lane.buf[s]);
        const size_t f0 = first + c * chunk;
        const size_t nf = (c + 1 == n_chunks) ? (count - c * chunk) : chunk;
        std::vector<void *> dptr(arrays.size());
        for (size_t i = 0; i < arrays.size(); i++) {
            const HostArray &a = arrays[i];
            dptr[i] = base + offs[i];
            if (a.host_in) {
                const unsigned char *src = static_cast<const unsigned char *>(a.host_in) + f0 * a.bytes_per_frame;
                cudaError_t e = cudaMemcpyAsync(dptr[i], src, nf * a.bytes_per_frame, cudaMemcpyHostToDevice, st);
                if (e != cudaSuccess) { rc = cuda_error(e, "H2D copy"); break; }
            }
        }
        if (rc != LDPC_OK) break;
        cudaError_t e = locked_launch(ctx, mu, launch, dptr, nf, st, f0, lane.counters + 16 * s);
        if (e != cudaSuccess) { rc = cuda_error(e, "kernel launch"); break; }
        for (size_t i = 0; i < arrays.size(); i++) {
            const HostArray &a = arrays[i];
            if (a.host_out) {
                unsigned char *dst = static_cast<unsigned char *>(a.host_out) + f0 * a.bytes_per_frame;
                e = cudaMemcpyAsync(dst, dptr[i], nf * a.bytes_per_frame, cudaMemcpyDeviceToHost, st);
                if (e != cudaSuccess) { rc = cuda_error(e, "D2H copy"); break; }
            }
        }
    }
    for (int s = 0; s < slots; s++) {
        cudaError_t e = cudaStreamSynchronize(lane.stream[s]);
        if (e != cudaSuccess && rc == LDPC_OK) rc = cuda_error(e, "stream synchronize");
    }
    return rc;
}

}  // namespace

int run_host_batch_at(const std::vector<HostArray> &arrays, size_t batch, const BatchLaunchAt &launch) {
    if (batch == 0) return LDPC_OK;
    int rc = runtime_init(nullptr, 0);
    if (rc != LDPC_OK) return rc;
    std::vector<int> devices;
    {
        Runtime &r = rt();
        std::lock_guard<std::mutex> lock(r.mu);
        devices = r.devices;
    }
    int prev = -1;
    cudaGetDevice(&prev);
    const size_t nd = devices.size();
    if (nd <= 1 || batch < nd) {
        DeviceCtx *ctx = nullptr;
        std::mutex *mu = nullptr;
        rc = get_ctx(devices.empty() ? 0 : devices[0], &ctx, &mu);
        if (rc == LDPC_OK) rc = run_on_device_host_ptrs(*ctx, *mu, arrays, 0, batch, launch);
    } else {
        // contiguous shards of independent codewords, one host thread per device, no collective
        std::vector<int> rcs(nd, LDPC_OK);
        std::vector<std::string> errs(nd);
        std::vector<std::thread> threads;
        for (size_t g = 0; g < nd; g++) {
            const size_t f0 = batch * g / nd, f1 = batch * (g + 1) / nd;
            threads.emplace_back([&, g, f0, f1]() {
                DeviceCtx *ctx = nullptr;
                std::mutex *mu = nullptr;
                int r2 = get_ctx(devices[g], &ctx, &mu);
                if (r2 == LDPC_OK) r2 = run_on_device_host_ptrs(*ctx, *mu, arrays, f0, f1 - f0, launch);
                rcs[g] = r2;
                if (r2 != LDPC_OK) errs[g] = last_error();
            });
        }
        for (auto &t : threads) t.join();
        for (size_t g = 0; g < nd; g++)
            if (rcs[g] != LDPC_OK) { rc = set_error(rcs[g], errs[g]); break; }
    }
    if (prev >= 0) cudaSetDevice(prev);
    return rc;
}

int run_device_batch(int device, cudaStream_t stream, bool synchronize,
                     const std::function<cudaError_t(DeviceCtx &, cudaStream_t)> &launch) {
    DeviceCtx *ctx = nullptr;
    std::mutex *mu = nullptr;
    int rc = get_ctx(device, &ctx, &mu);
    if (rc != LDPC_OK) return rc;
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != device) CUDA_TRY(cudaSetDevice(device));
    cudaError_t e;
    {
        std::lock_guard<std::mutex> lock(*mu);   // launchers may grow per-context scratch
        e = launch(*ctx, stream);
    }
    if (e == cudaSuccess && synchronize) e = cudaStreamSynchronize(stream);
    if (prev != device && prev >= 0) cudaSetDevice(prev);
    if (e != cudaSuccess) return cuda_error(e, "device batch");
    return LDPC_OK;
}

}  // namespace ldpc
