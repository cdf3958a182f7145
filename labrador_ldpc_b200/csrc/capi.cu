// C ABI of the library: the reference's 21 entry points (single codeword =
// batch of one on the GPU) plus the batched / stream-ordered additions.
// Declarations and reference citations: include/labrador_ldpc.h.
#include <cstdio>
#include <cstdlib>

#include "../../include/labrador_ldpc.h"
#include "host_api.h"

using namespace ldpc;

namespace {

const CodeInfo *info(int code) { return code_info(code); }

// 0 host, 1 device, -1 mixed.  Null pointers are skipped.
int common_kind(std::initializer_list<const void *> ptrs, int *device) {
    int kind = -2, dev = -1;
    for (const void *p : ptrs) {
        if (!p) continue;
        int d = -1;
        const int k = classify_pointer(p, &d);
        if (kind == -2) { kind = k; dev = d; }
        else if (kind != k || (k == 1 && d != dev)) return -1;
    }
    if (device) *device = dev;
    return kind == -2 ? 0 : kind;
}

// front: what `llrs` holds (runtime.h) -- LLRs of type ty, float soft values quantised to ty on load, or hard bits.
int decode_ms_impl(int code, int ty, const void *llrs, uint8_t *output, size_t batch, size_t max_iters,
                   uint8_t *success, uint32_t *iters, bool async, cudaStream_t stream, const Front &front = Front()) {
    const CodeInfo *c = info(code);
    if (!c) return fail(LDPC_ERR_BAD_CODE, "code out of range");
    if (ty < 0 || ty >= kNumLlrTypes) return fail(LDPC_ERR_BAD_ARGUMENT, "bad llr_type");
    if (!front_supported(front.kind, ty))
        return fail(LDPC_ERR_BAD_ARGUMENT, "front end not available for this llr_type (soft: i8/i16, hard: i8)");
    if (front.kind == kFrontSoftF32) {
        const float max_limit = ty == kI8 ? 127.0f : 32767.0f;
        if (!(front.limit >= 1.0f && front.limit <= max_limit) || front.limit != (float)(int)front.limit)
            return fail(LDPC_ERR_BAD_ARGUMENT, "limit must be an integer in 1..127 (i8) or 1..32767 (i16)");
        if (!(front.scale == front.scale) || front.scale - front.scale != 0.0f)
            return fail(LDPC_ERR_BAD_ARGUMENT, "scale must be finite");
    }
    if (batch == 0) return LDPC_OK;
    if (!llrs || !output) return fail(LDPC_ERR_NULL_POINTER, "llrs/output must not be NULL");
    int dev = -1;
    const int kind = common_kind({llrs, output, success, iters}, &dev);
    if (async && kind != 1) return fail(LDPC_ERR_MIXED_POINTERS, "_async entry points take device pointers only");
    if (kind < 0) return fail(LDPC_ERR_MIXED_POINTERS, "pointers must be all host or all on one device");
    if (kind == 1) {
        return run_device_batch(dev, stream, !async, [&](DeviceCtx &ctx, cudaStream_t st) {
            return launch_decode_ms(ctx, code, ty, llrs, output, batch, max_iters, success, iters, st, front);
        });
    }
    std::vector<HostArray> arrays;
    arrays.push_back({llrs, nullptr, front_frame_bytes(front, c->n, ty)});
    arrays.push_back({nullptr, output, c->output_len()});
    const int si = success ? (int)arrays.size() : -1;
    if (success) arrays.push_back({nullptr, success, 1});
    const int ii = iters ? (int)arrays.size() : -1;
    if (iters) arrays.push_back({nullptr, iters, 4});
    return run_host_batch(arrays, batch, [=](DeviceCtx &ctx, const std::vector<void *> &d, size_t nf, cudaStream_t st) {
        return launch_decode_ms(ctx, code, ty, d[0], static_cast<uint8_t *>(d[1]), nf, max_iters,
                                si >= 0 ? static_cast<uint8_t *>(d[si]) : nullptr,
                                ii >= 0 ? static_cast<uint32_t *>(d[ii]) : nullptr, st, front);
    });
}

int quantise_impl(int code, int ty, const float *soft, float scale, int limit, void *llrs, size_t batch, bool async,
                  cudaStream_t stream) {
    const CodeInfo *c = info(code);
    if (!c) return fail(LDPC_ERR_BAD_CODE, "code out of range");
    if (ty != kI8 && ty != kI16) return fail(LDPC_ERR_BAD_ARGUMENT, "quantise produces i8 or i16 LLRs");
    if (limit < 1 || limit > (ty == kI8 ? 127 : 32767))
        return fail(LDPC_ERR_BAD_ARGUMENT, "limit must be in 1..127 (i8) or 1..32767 (i16)");
    if (!(scale == scale) || scale - scale != 0.0f) return fail(LDPC_ERR_BAD_ARGUMENT, "scale must be finite");
    if (batch == 0) return LDPC_OK;
    if (!soft || !llrs) return fail(LDPC_ERR_NULL_POINTER, "soft/llrs must not be NULL");
    int dev = -1;
    const int kind = common_kind({soft, llrs}, &dev);
    if (async && kind != 1) return fail(LDPC_ERR_MIXED_POINTERS, "_async entry points take device pointers only");
    if (kind < 0) return fail(LDPC_ERR_MIXED_POINTERS, "pointers must be all host or all on one device");
    const float flimit = (float)limit;
    if (kind == 1) {
        return run_device_batch(dev, stream, !async, [&](DeviceCtx &ctx, cudaStream_t st) {
            return launch_quantise(ctx, code, ty, soft, llrs, batch, scale, flimit, st);
        });
    }
    std::vector<HostArray> arrays;
    arrays.push_back({soft, nullptr, (size_t)c->n * 4});
    arrays.push_back({nullptr, llrs, (size_t)c->n * llr_size(ty)});
    return run_host_batch(arrays, batch, [=](DeviceCtx &ctx, const std::vector<void *> &d, size_t nf, cudaStream_t st) {
        return launch_quantise(ctx, code, ty, static_cast<const float *>(d[0]), d[1], nf, scale, flimit, st);
    });
}

int random_data_impl(int code, unsigned long long seed, unsigned long long first_frame, uint8_t *data, size_t batch,
                     bool async, cudaStream_t stream) {
    const CodeInfo *c = info(code);
    if (!c) return fail(LDPC_ERR_BAD_CODE, "code out of range");
    if (batch == 0) return LDPC_OK;
    if (!data) return fail(LDPC_ERR_NULL_POINTER, "data must not be NULL");
    int dev = -1;
    const int kind = common_kind({data}, &dev);
    if (async && kind != 1) return fail(LDPC_ERR_MIXED_POINTERS, "_async entry points take device pointers only");
    if (kind == 1) {
        return run_device_batch(dev, stream, !async, [&](DeviceCtx &ctx, cudaStream_t st) {
            return launch_random_data(ctx, code, seed, first_frame, data, batch, st);
        });
    }
    std::vector<HostArray> arrays;
    arrays.push_back({nullptr, data, (size_t)c->k / 8});
    return run_host_batch_at(arrays, batch, [=](DeviceCtx &ctx, const std::vector<void *> &d, size_t nf, cudaStream_t st,
                                                size_t first) {
        return launch_random_data(ctx, code, seed, first_frame + first, static_cast<uint8_t *>(d[0]), nf, st);
    });
}

int awgn_impl(int code, int ty, const uint8_t *codewords, float sigma, float scale, int limit, unsigned long long seed,
              unsigned long long first_frame, void *out, size_t batch, bool async, cudaStream_t stream) {
    const CodeInfo *c = info(code);
    if (!c) return fail(LDPC_ERR_BAD_CODE, "code out of range");
    if (ty != kI8 && ty != kI16 && ty != kF32) return fail(LDPC_ERR_BAD_ARGUMENT, "channel output is i8, i16 or f32");
    if (ty != kF32 && (limit < 1 || limit > (ty == kI8 ? 127 : 32767)))
        return fail(LDPC_ERR_BAD_ARGUMENT, "limit must be in 1..127 (i8) or 1..32767 (i16)");
    if (!(sigma >= 0.0f) || sigma - sigma != 0.0f || !(scale == scale) || scale - scale != 0.0f)
        return fail(LDPC_ERR_BAD_ARGUMENT, "sigma must be finite and >= 0, scale finite");
    if (batch == 0) return LDPC_OK;
    if (!codewords || !out) return fail(LDPC_ERR_NULL_POINTER, "codewords/out must not be NULL");
    int dev = -1;
    const int kind = common_kind({codewords, out}, &dev);
    if (async && kind != 1) return fail(LDPC_ERR_MIXED_POINTERS, "_async entry points take device pointers only");
    if (kind < 0) return fail(LDPC_ERR_MIXED_POINTERS, "pointers must be all host or all on one device");
    const float flimit = (float)limit;
    if (kind == 1) {
        return run_device_batch(dev, stream, !async, [&](DeviceCtx &ctx, cudaStream_t st) {
            return launch_awgn(ctx, code, ty, codewords, sigma, scale, flimit, seed, first_frame, out, batch, st);
        });
    }
    std::vector<HostArray> arrays;
    arrays.push_back({codewords, nullptr, (size_t)c->n / 8});
    arrays.push_back({nullptr, out, (size_t)c->n * llr_size(ty)});
    return run_host_batch_at(arrays, batch, [=](DeviceCtx &ctx, const std::vector<void *> &d, size_t nf, cudaStream_t st,
                                                size_t first) {
        return launch_awgn(ctx, code, ty, static_cast<const uint8_t *>(d[0]), sigma, scale, flimit, seed,
                           first_frame + first, d[1], nf, st);
    });
}

int count_errors_impl(int code, const uint8_t *decoded, const uint8_t *data, uint32_t *errors, size_t batch, bool async,
                      cudaStream_t stream) {
    const CodeInfo *c = info(code);
    if (!c) return fail(LDPC_ERR_BAD_CODE, "code out of range");
    if (batch == 0) return LDPC_OK;
    if (!decoded || !data || !errors) return fail(LDPC_ERR_NULL_POINTER, "decoded/data/errors must not be NULL");
    int dev = -1;
    const int kind = common_kind({decoded, data, errors}, &dev);
    if (async && kind != 1) return fail(LDPC_ERR_MIXED_POINTERS, "_async entry points take device pointers only");
    if (kind < 0) return fail(LDPC_ERR_MIXED_POINTERS, "pointers must be all host or all on one device");
    if (kind == 1) {
        return run_device_batch(dev, stream, !async, [&](DeviceCtx &ctx, cudaStream_t st) {
            return launch_count_errors(ctx, code, decoded, data, errors, batch, st);
        });
    }
    std::vector<HostArray> arrays;
    arrays.push_back({decoded, nullptr, c->output_len()});
    arrays.push_back({data, nullptr, (size_t)c->k / 8});
    arrays.push_back({nullptr, errors, 4});
    return run_host_batch(arrays, batch, [=](DeviceCtx &ctx, const std::vector<void *> &d, size_t nf, cudaStream_t st) {
        return launch_count_errors(ctx, code, static_cast<const uint8_t *>(d[0]), static_cast<const uint8_t *>(d[1]),
                                   static_cast<uint32_t *>(d[2]), nf, st);
    });
}

Front soft_front(float scale, int limit) {
    Front f;
    f.kind = kFrontSoftF32;
    f.scale = scale;
    f.limit = (float)limit;
    return f;
}

int decode_bf_impl(int code, const uint8_t *input, uint8_t *output, size_t batch, size_t max_iters,
                   uint8_t *success, uint32_t *iters, bool async, cudaStream_t stream) {
    const CodeInfo *c = info(code);
    if (!c) return fail(LDPC_ERR_BAD_CODE, "code out of range");
    if (batch == 0) return LDPC_OK;
    if (!input || !output) return fail(LDPC_ERR_NULL_POINTER, "input/output must not be NULL");
    int dev = -1;
    const int kind = common_kind({input, output, success, iters}, &dev);
    if (async && kind != 1) return fail(LDPC_ERR_MIXED_POINTERS, "_async entry points take device pointers only");
    if (kind < 0) return fail(LDPC_ERR_MIXED_POINTERS, "pointers must be all host or all on one device");
    if (kind == 1) {
        return run_device_batch(dev, stream, !async, [&](DeviceCtx &ctx, cudaStream_t st) {
            return launch_decode_bf(ctx, code, input, output, batch, max_iters, success, iters, st);
        });
    }
    std::vector<HostArray> arrays;
    arrays.push_back({input, nullptr, (size_t)c->n / 8});
    arrays.push_back({nullptr, output, c->output_len()});
    const int si = success ? (int)arrays.size() : -1;
    if (success) arrays.push_back({nullptr, success, 1});
    const int ii = iters ? (int)arrays.size() : -1;
    if (iters) arrays.push_back({nullptr, iters, 4});
    return run_host_batch(arrays, batch, [=](DeviceCtx &ctx, const std::vector<void *> &d, size_t nf, cudaStream_t st) {
        return launch_decode_bf(ctx, code, static_cast<const uint8_t *>(d[0]), static_cast<uint8_t *>(d[1]), nf,
                                max_iters, si >= 0 ? static_cast<uint8_t *>(d[si]) : nullptr,
                                ii >= 0 ? static_cast<uint32_t *>(d[ii]) : nullptr, st);
    });
}

// data == nullptr: encode in place (first k/8 bytes of every codeword already set).
int encode_impl(int code, const uint8_t *data, uint8_t *codewords, size_t batch, bool async, cudaStream_t stream) {
    const CodeInfo *c = info(code);
    if (!c) return fail(LDPC_ERR_BAD_CODE, "code out of range");
    if (batch == 0) return LDPC_OK;
    if (!codewords) return fail(LDPC_ERR_NULL_POINTER, "codewords must not be NULL");
    int dev = -1;
    const int kind = common_kind({data, codewords}, &dev);
    if (async && kind != 1) return fail(LDPC_ERR_MIXED_POINTERS, "_async entry points take device pointers only");
    if (kind < 0) return fail(LDPC_ERR_MIXED_POINTERS, "pointers must be all host or all on one device");
    if (kind == 1) {
        return run_device_batch(dev, stream, !async, [&](DeviceCtx &ctx, cudaStream_t st) {
            return launch_encode(ctx, code, data, codewords, batch, st);
        });
    }
    std::vector<HostArray> arrays;
    if (data && data != codewords) {
        arrays.push_back({data, nullptr, (size_t)c->k / 8});
        arrays.push_back({nullptr, codewords, (size_t)c->n / 8});
        return run_host_batch(arrays, batch, [=](DeviceCtx &ctx, const std::vector<void *> &d, size_t nf, cudaStream_t st) {
            return launch_encode(ctx, code, static_cast<const uint8_t *>(d[0]), static_cast<uint8_t *>(d[1]), nf, st);
        });
    }
    arrays.push_back({codewords, codewords, (size_t)c->n / 8});
    return run_host_batch(arrays, batch, [=](DeviceCtx &ctx, const std::vector<void *> &d, size_t nf, cudaStream_t st) {
        return launch_encode(ctx, code, nullptr, static_cast<uint8_t *>(d[0]), nf, st);
    });
}

int h2l_impl(int code, int ty, const uint8_t *input, void *llrs, size_t batch, bool async, cudaStream_t stream) {
    const CodeInfo *c = info(code);
    if (!c) return fail(LDPC_ERR_BAD_CODE, "code out of range");
    if (ty < 0 || ty >= kNumLlrTypes) return fail(LDPC_ERR_BAD_ARGUMENT, "bad llr_type");
    if (batch == 0) return LDPC_OK;
    if (!input || !llrs) return fail(LDPC_ERR_NULL_POINTER, "input/llrs must not be NULL");
    int dev = -1;
    const int kind = common_kind({input, llrs}, &dev);
    if (async && kind != 1) return fail(LDPC_ERR_MIXED_POINTERS, "_async entry points take device pointers only");
    if (kind < 0) return fail(LDPC_ERR_MIXED_POINTERS, "pointers must be all host or all on one device");
    if (kind == 1) {
        return run_device_batch(dev, stream, !async, [&](DeviceCtx &ctx, cudaStream_t st) {
            return launch_hard_to_llrs(ctx, code, ty, input, llrs, batch, st);
        });
    }
    std::vector<HostArray> arrays;
    arrays.push_back({input, nullptr, (size_t)c->n / 8});
    arrays.push_back({nullptr, llrs, (size_t)c->n * llr_size(ty)});
    return run_host_batch(arrays, batch, [=](DeviceCtx &ctx, const std::vector<void *> &d, size_t nf, cudaStream_t st) {
        return launch_hard_to_llrs(ctx, code, ty, static_cast<const uint8_t *>(d[0]), d[1], nf, st);
    });
}

int l2h_impl(int code, int ty, const void *llrs, uint8_t *output, size_t batch, bool async, cudaStream_t stream) {
    const CodeInfo *c = info(code);
    if (!c) return fail(LDPC_ERR_BAD_CODE, "code out of range");
    if (ty < 0 || ty >= kNumLlrTypes) return fail(LDPC_ERR_BAD_ARGUMENT, "bad llr_type");
    if (batch == 0) return LDPC_OK;
    if (!llrs || !output) return fail(LDPC_ERR_NULL_POINTER, "llrs/output must not be NULL");
    int dev = -1;
    const int kind = common_kind({llrs, output}, &dev);
    if (async && kind != 1) return fail(LDPC_ERR_MIXED_POINTERS, "_async entry points take device pointers only");
    if (kind < 0) return fail(LDPC_ERR_MIXED_POINTERS, "pointers must be all host or all on one device");
    if (kind == 1) {
        return run_device_batch(dev, stream, !async, [&](DeviceCtx &ctx, cudaStream_t st) {
            return launch_llrs_to_hard(ctx, code, ty, llrs, output, batch, st);
        });
    }
    std::vector<HostArray> arrays;
    arrays.push_back({llrs, nullptr, (size_t)c->n * llr_size(ty)});
    arrays.push_back({nullptr, output, (size_t)c->n / 8});
    return run_host_batch(arrays, batch, [=](DeviceCtx &ctx, const std::vector<void *> &d, size_t nf, cudaStream_t st) {
        return launch_llrs_to_hard(ctx, code, ty, d[0], static_cast<uint8_t *>(d[1]), nf, st);
    });
}

// The reference-signature functions have no error channel: report and abort.
void die_if(int rc, const char *fn) {
    if (rc == LDPC_OK) return;
    fprintf(stderr, "labrador_ldpc (B200): %s failed: %s (code %d); this library has no CPU fallback\n", fn,
            last_error(), rc);
    abort();
}

bool single_ms(int code, int ty, const void *llrs, uint8_t *output, size_t max_iters, size_t *iters_run,
               const char *fn) {
    uint8_t ok = 0;
    uint32_t it = 0;
    die_if(decode_ms_impl(code, ty, llrs, output, 1, max_iters, &ok, &it, false, nullptr), fn);
    if (iters_run) *iters_run = it;
    return ok != 0;
}

}  // namespace

extern "C" {

size_t labrador_ldpc_code_n(enum labrador_ldpc_code code) { const CodeInfo *c = info(code); return c ? c->n : 0; }
size_t labrador_ldpc_code_k(enum labrador_ldpc_code code) { const CodeInfo *c = info(code); return c ? c->k : 0; }
size_t labrador_ldpc_bf_working_len(enum labrador_ldpc_code code) { const CodeInfo *c = info(code); return c ? c->bf_working_len() : 0; }
size_t labrador_ldpc_ms_working_u8_len(enum labrador_ldpc_code code) { const CodeInfo *c = info(code); return c ? c->ms_working_u8_len() : 0; }
size_t labrador_ldpc_ms_working_len(enum labrador_ldpc_code code) { const CodeInfo *c = info(code); return c ? c->ms_working_len() : 0; }
size_t labrador_ldpc_output_len(enum labrador_ldpc_code code) { const CodeInfo *c = info(code); return c ? c->output_len() : 0; }

void labrador_ldpc_encode(enum labrador_ldpc_code code, uint8_t *codeword) {
    die_if(encode_impl(code, nullptr, codeword, 1, false, nullptr), "labrador_ldpc_encode");
}

void labrador_ldpc_copy_encode(enum labrador_ldpc_code code, const uint8_t *data, uint8_t *codeword) {
    die_if(encode_impl(code, data, codeword, 1, false, nullptr), "labrador_ldpc_copy_encode");
}

bool labrador_ldpc_decode_bf(enum labrador_ldpc_code code, const uint8_t *input, uint8_t *output,
                             uint8_t *working, size_t max_iters, size_t *iters_run) {
    (void)working;
    uint8_t ok = 0;
    uint32_t it = 0;
    die_if(decode_bf_impl(code, input, output, 1, max_iters, &ok, &it, false, nullptr), "labrador_ldpc_decode_bf");
    if (iters_run) *iters_run = it;
    return ok != 0;
}

#define LDPC_SINGLE_MS(SUFFIX, T, TY)                                                                       \
    bool labrador_ldpc_decode_ms_##SUFFIX(enum labrador_ldpc_code code, const T *llrs, uint8_t *output,      \
                                          T *working, uint8_t *working_u8, size_t max_iters,                 \
                                          size_t *iters_run) {                                               \
        (void)working; (void)working_u8;                                                                    \
        return single_ms(code, TY, llrs, output, max_iters, iters_run, "labrador_ldpc_decode_ms_" #SUFFIX); \
    }                                                                                                       \
    int labrador_ldpc_decode_ms_##SUFFIX##_batch(enum labrador_ldpc_code code, const T *llrs, uint8_t *output, \
                                                 size_t batch, size_t max_iters, uint8_t *success,           \
                                                 uint32_t *iters_run) {                                      \
        return decode_ms_impl(code, TY, llrs, output, batch, max_iters, success, iters_run, false, nullptr); \
    }                                                                                                       \
    void labrador_ldpc_hard_to_llrs_##SUFFIX(enum labrador_ldpc_code code, const uint8_t *input, T *llrs) {  \
        die_if(h2l_impl(code, TY, input, llrs, 1, false, nullptr), "labrador_ldpc_hard_to_llrs_" #SUFFIX);   \
    }                                                                                                       \
    void labrador_ldpc_llrs_to_hard_##SUFFIX(enum labrador_ldpc_code code, const T *llrs, uint8_t *output) { \
        die_if(l2h_impl(code, TY, llrs, output, 1, false, nullptr), "labrador_ldpc_llrs_to_hard_" #SUFFIX);  \
    }                                                                                                       \
    int labrador_ldpc_hard_to_llrs_##SUFFIX##_batch(enum labrador_ldpc_code code, const uint8_t *input,      \
                                                    T *llrs, size_t batch) {                                 \
        return h2l_impl(code, TY, input, llrs, batch, false, nullptr);                                       \
    }                                                                                                       \
    int labrador_ldpc_llrs_to_hard_##SUFFIX##_batch(enum labrador_ldpc_code code, const T *llrs,             \
                                                    uint8_t *output, size_t batch) {                         \
        return l2h_impl(code, TY, llrs, output, batch, false, nullptr);                                      \
    }

LDPC_SINGLE_MS(i8, int8_t, kI8)
LDPC_SINGLE_MS(i16, int16_t, kI16)
LDPC_SINGLE_MS(i32, int32_t, kI32)
LDPC_SINGLE_MS(f32, float, kF32)
LDPC_SINGLE_MS(f64, double, kF64)

int labrador_ldpc_cuda_init(const int *devices, int n_devices) { return runtime_init(devices, n_devices); }
void labrador_ldpc_cuda_shutdown(void) { runtime_shutdown(); }
int labrador_ldpc_cuda_device_count(void) { return runtime_device_count(); }
const char *labrador_ldpc_last_error(void) { return last_error(); }
const char *labrador_ldpc_version(void) { return "labrador-ldpc-b200 0.1.0 (drop-in for labrador-ldpc 1.2.1 C API, sm_100a)"; }

void *labrador_ldpc_alloc_pinned(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void labrador_ldpc_free_pinned(void *ptr) { if (ptr) cudaFreeHost(ptr); }

int labrador_ldpc_decode_bf_batch(enum labrador_ldpc_code code, const uint8_t *input, uint8_t *output,
                                  size_t batch, size_t max_iters, uint8_t *success, uint32_t *iters_run) {
    return decode_bf_impl(code, input, output, batch, max_iters, success, iters_run, false, nullptr);
}

int labrador_ldpc_encode_batch(enum labrador_ldpc_code code, uint8_t *codewords, size_t batch) {
    return encode_impl(code, nullptr, codewords, batch, false, nullptr);
}

int labrador_ldpc_copy_encode_batch(enum labrador_ldpc_code code, const uint8_t *data, uint8_t *codewords,
                                    size_t batch) {
    if (!data && batch) return fail(LDPC_ERR_NULL_POINTER, "data must not be NULL");
    return encode_impl(code, data, codewords, batch, false, nullptr);
}

int labrador_ldpc_decode_ms_batch_async(enum labrador_ldpc_code code, int llr_type, const void *llrs,
                                        uint8_t *output, size_t batch, size_t max_iters, uint8_t *success,
                                        uint32_t *iters_run, void *cuda_stream) {
    return decode_ms_impl(code, llr_type, llrs, output, batch, max_iters, success, iters_run, true,
                          static_cast<cudaStream_t>(cuda_stream));
}

int labrador_ldpc_decode_bf_batch_async(enum labrador_ldpc_code code, const uint8_t *input, uint8_t *output,
                                        size_t batch, size_t max_iters, uint8_t *success, uint32_t *iters_run,
                                        void *cuda_stream) {
    return decode_bf_impl(code, input, output, batch, max_iters, success, iters_run, true,
                          static_cast<cudaStream_t>(cuda_stream));
}

int labrador_ldpc_copy_encode_batch_async(enum labrador_ldpc_code code, const uint8_t *data, uint8_t *codewords,
                                          size_t batch, void *cuda_stream) {
    return encode_impl(code, data, codewords, batch, true, static_cast<cudaStream_t>(cuda_stream));
}

int labrador_ldpc_hard_to_llrs_batch_async(enum labrador_ldpc_code code, int llr_type, const uint8_t *input,
                                           void *llrs, size_t batch, void *cuda_stream) {
    return h2l_impl(code, llr_type, input, llrs, batch, true, static_cast<cudaStream_t>(cuda_stream));
}

int labrador_ldpc_llrs_to_hard_batch_async(enum labrador_ldpc_code code, int llr_type, const void *llrs,
                                           uint8_t *output, size_t batch, void *cuda_stream) {
    return l2h_impl(code, llr_type, llrs, output, batch, true, static_cast<cudaStream_t>(cuda_stream));
}

// Copy-only control of the host-pointer decode path: the same arrays go through the same chunked pipeline
// (runtime.cu), but no kernel is launched.  What it measures is the host<->device transport alone.
int labrador_ldpc_copy_control_batch(enum labrador_ldpc_code code, int llr_type, const void *llrs, uint8_t *output,
                                     size_t batch, uint8_t *success, uint32_t *iters_run) {
    const CodeInfo *c = info(code);
    if (!c) return fail(LDPC_ERR_BAD_CODE, "code out of range");
    if (llr_type < 0 || llr_type >= kNumLlrTypes) return fail(LDPC_ERR_BAD_ARGUMENT, "bad llr_type");
    if (batch == 0) return LDPC_OK;
    if (!llrs || !output) return fail(LDPC_ERR_NULL_POINTER, "llrs/output must not be NULL");
    if (common_kind({llrs, output, success, iters_run}, nullptr) != 0)
        return fail(LDPC_ERR_MIXED_POINTERS, "the copy control takes host pointers only");
    std::vector<HostArray> arrays;
    arrays.push_back({llrs, nullptr, (size_t)c->n * llr_size(llr_type)});
    arrays.push_back({nullptr, output, c->output_len()});
    if (success) arrays.push_back({nullptr, success, 1});
    if (iters_run) arrays.push_back({nullptr, iters_run, 4});
    return run_host_batch(arrays, batch, [](DeviceCtx &, const std::vector<void *> &, size_t, cudaStream_t) {
        return cudaSuccess;
    });
}

// ---- fused front ends (csrc/front.cuh) ----
int labrador_ldpc_decode_ms_i8_soft_batch(enum labrador_ldpc_code code, const float *soft, float scale, int limit,
                                          uint8_t *output, size_t batch, size_t max_iters, uint8_t *success,
                                          uint32_t *iters_run) {
    return decode_ms_impl(code, kI8, soft, output, batch, max_iters, success, iters_run, false, nullptr,
                          soft_front(scale, limit));
}

int labrador_ldpc_decode_ms_i16_soft_batch(enum labrador_ldpc_code code, const float *soft, float scale, int limit,
                                           uint8_t *output, size_t batch, size_t max_iters, uint8_t *success,
                                           uint32_t *iters_run) {
    return decode_ms_impl(code, kI16, soft, output, batch, max_iters, success, iters_run, false, nullptr,
                          soft_front(scale, limit));
}

int labrador_ldpc_decode_ms_i8_hard_batch(enum labrador_ldpc_code code, const uint8_t *input, uint8_t *output,
                                          size_t batch, size_t max_iters, uint8_t *success, uint32_t *iters_run) {
    Front f;
    f.kind = kFrontHard;
    return decode_ms_impl(code, kI8, input, output, batch, max_iters, success, iters_run, false, nullptr, f);
}

int labrador_ldpc_decode_ms_front_batch_async(enum labrador_ldpc_code code, int llr_type, int front, const void *input,
                                              float scale, int limit, uint8_t *output, size_t batch, size_t max_iters,
                                              uint8_t *success, uint32_t *iters_run, void *cuda_stream) {
    Front f;
    if (front == LABRADOR_LDPC_FRONT_SOFT_F32) f = soft_front(scale, limit);
    else if (front == LABRADOR_LDPC_FRONT_HARD) f.kind = kFrontHard;
    else if (front != LABRADOR_LDPC_FRONT_NONE) return fail(LDPC_ERR_BAD_ARGUMENT, "bad front");
    return decode_ms_impl(code, llr_type, input, output, batch, max_iters, success, iters_run, true,
                          static_cast<cudaStream_t>(cuda_stream), f);
}

int labrador_ldpc_quantise_i8_batch(enum labrador_ldpc_code code, const float *soft, float scale, int limit,
                                    int8_t *llrs, size_t batch) {
    return quantise_impl(code, kI8, soft, scale, limit, llrs, batch, false, nullptr);
}

int labrador_ldpc_quantise_i16_batch(enum labrador_ldpc_code code, const float *soft, float scale, int limit,
                                     int16_t *llrs, size_t batch) {
    return quantise_impl(code, kI16, soft, scale, limit, llrs, batch, false, nullptr);
}

int labrador_ldpc_quantise_batch_async(enum labrador_ldpc_code code, int llr_type, const float *soft, float scale,
                                       int limit, void *llrs, size_t batch, void *cuda_stream) {
    return quantise_impl(code, llr_type, soft, scale, limit, llrs, batch, true, static_cast<cudaStream_t>(cuda_stream));
}

// ---- harness kernels: counter-based frame generator and error counter (csrc/channel.cu) ----
int labrador_ldpc_random_data_batch(enum labrador_ldpc_code code, uint64_t seed, uint64_t first_frame, uint8_t *data,
                                    size_t batch) {
    return random_data_impl(code, seed, first_frame, data, batch, false, nullptr);
}

int labrador_ldpc_awgn_batch(enum labrador_ldpc_code code, int out_type, const uint8_t *codewords, float sigma,
                             float scale, int limit, uint64_t seed, uint64_t first_frame, void *out, size_t batch) {
    return awgn_impl(code, out_type, codewords, sigma, scale, limit, seed, first_frame, out, batch, false, nullptr);
}

int labrador_ldpc_count_errors_batch(enum labrador_ldpc_code code, const uint8_t *decoded, const uint8_t *data,
                                     uint32_t *bit_errors, size_t batch) {
    return count_errors_impl(code, decoded, data, bit_errors, batch, false, nullptr);
}

int labrador_ldpc_random_data_batch_async(enum labrador_ldpc_code code, uint64_t seed, uint64_t first_frame,
                                          uint8_t *data, size_t batch, void *cuda_stream) {
    return random_data_impl(code, seed, first_frame, data, batch, true, static_cast<cudaStream_t>(cuda_stream));
}

int labrador_ldpc_awgn_batch_async(enum labrador_ldpc_code code, int out_type, const uint8_t *codewords, float sigma,
                                   float scale, int limit, uint64_t seed, uint64_t first_frame, void *out, size_t batch,
                                   void *cuda_stream) {
    return awgn_impl(code, out_type, codewords, sigma, scale, limit, seed, first_frame, out, batch, true,
                     static_cast<cudaStream_t>(cuda_stream));
}

int labrador_ldpc_count_errors_batch_async(enum labrador_ldpc_code code, const uint8_t *decoded, const uint8_t *data,
                                           uint32_t *bit_errors, size_t batch, void *cuda_stream) {
    return count_errors_impl(code, decoded, data, bit_errors, batch, true, static_cast<cudaStream_t>(cuda_stream));
}

unsigned long long labrador_ldpc_kernel_launch_count(void) { return launch_count(); }

int labrador_ldpc_host_encode_model(enum labrador_ldpc_code code, const uint8_t *data, uint8_t *parity_tables,
                                    uint8_t *parity_generator) {
    if (!ldpc::code_info((int)code)) return LABRADOR_LDPC_ERR_BAD_CODE;
    if (!data || !parity_tables) return LABRADOR_LDPC_ERR_NULL_POINTER;
    const bool has_gen = ldpc::code_info((int)code)->gen != nullptr;
    if (has_gen) {
        if (!parity_generator) return LABRADOR_LDPC_ERR_NULL_POINTER;
        ldpc::host_encode_generator((int)code, data, parity_generator);
    }
    if (!ldpc::host_encode_tables((int)code, data, parity_tables)) return LABRADOR_LDPC_ERR_BAD_ARGUMENT;
    return has_gen ? 0 : 1;      // 1: k = 16384 code -- no generator exists, parity_generator was not written
}

const char *labrador_ldpc_decode_ms_kernel_name(enum labrador_ldpc_code code, int llr_type) {
    if (!info(code)) return "invalid";
    return decode_ms_kernel_name(code, llr_type);
}

uint32_t labrador_ldpc_edge_table_crc(enum labrador_ldpc_code code) {
    const CodeInfo *c = info(code);
    return c ? edge_crc(*c) : 0;
}

}  // extern "C"
