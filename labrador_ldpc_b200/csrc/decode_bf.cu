// Batched bit-flipping decoder with the erasure pre-pass for punctured codes.
//
// Replaces LDPCCode::decode_bf (reference src/decoder.rs:243-301) and the
// private decode_erasures (src/decoder.rs:144-223).  One CTA per codeword.
// Both algorithms only use XOR parities, counts and a max, so every pass is
// order-independent and is run one-thread-per-check / one-thread-per-variable
// over the device edge tables.
//
// decode_erasures, as written in the reference, increments `bits_fixed` for
// every still-erased variable whether or not it got a majority
// (src/decoder.rs:205-213), so it always returns (true, 0) after exactly one
// pass when max_iters >= 1 and contributes 0 to the returned iteration count;
// that behaviour is reproduced here (one pass, erasure_iters = 0).
#include <cuda_runtime.h>

#include <cstdlib>

#include "runtime.h"

namespace ldpc {

bool launch_decode_bf_tm(DeviceCtx &ctx, int code, const uint8_t *input, uint8_t *output, size_t batch,
                         size_t max_iters, uint8_t *success, uint32_t *iters, cudaStream_t stream, cudaError_t *err);

bool launch_decode_bf_tc(DeviceCtx &ctx, int code, const uint8_t *input, uint8_t *output, size_t batch,
                         size_t max_iters, uint8_t *success, uint32_t *iters, cudaStream_t stream, cudaError_t *err);

namespace {

__global__ void decode_bf_kernel(const DeviceCode code, const uint8_t *__restrict__ in_all,
                                 uint8_t *__restrict__ out_all, unsigned long long batch, unsigned max_iters,
                                 uint8_t *__restrict__ success, uint32_t *__restrict__ iters_out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int n = code.n, nv = code.vars, nc = code.checks;
    const int dv = code.max_var_degree, dc = code.max_check_degree;
    uint8_t *bits = smem;            // [nv] hard decision per variable
    uint8_t *cpar = bits + nv;       // [nc] per-check parity (bit0) / single-erasure flag (bit1)
    uint8_t *cnt = cpar + nc;        // [nv] violated-check count per variable
    const int tid = threadIdx.x, nt = blockDim.x;
    const int in_len = n / 8, out_len = nv / 8;
    __shared__ unsigned s_seen;      // OR over the CTA of (1 << violation count)

    for (unsigned long long frame = blockIdx.x; frame < batch; frame += gridDim.x) {
        const uint8_t *in = in_all + frame * (unsigned long long)in_len;
        // output[..n/8] = input (:251); punctured bits start as zero (:167)
        for (int a = tid; a < nv; a += nt)
            bits[a] = a < n ? ((in[a >> 3] >> (7 - (a & 7))) & 1) : 0;
        __syncthreads();

        if (code.p > 0 && max_iters > 0) {
            // erasure pass A: parity over known bits + number of erased neighbours (:177-189)
            for (int c = tid; c < nc; c += nt) {
                int par = 0, er = 0;
                for (int j = 0; j < dc; j++) {
                    const uint64_t ent = __ldg(code.chk_tab + (size_t)j * nc + c);
                    if (ent == kNoEdge64) break;
                    const int var = (int)(ent >> 32);
                    if (var >= n) er++;
                    else par ^= bits[var];
                }
                cpar[c] = (uint8_t)(par | (er == 1 ? 2 : 0));
            }
            __syncthreads();
            // pass B + C: votes from single-erasure checks, majority sets the bit (:192-213)
            for (int a = n + tid; a < nv; a += nt) {
                int votes = 0;
                for (int j = 0; j < dv; j++) {
                    const uint64_t ent = __ldg(code.var_tab + (size_t)j * nv + a);
                    if (ent == kNoEdge64) break;
                    const int cp = cpar[(int)(ent >> 32)];
                    if (cp & 2) votes += (cp & 1) ? 1 : -1;
                }
                if (votes > 0) bits[a] = 1;
            }
            __syncthreads();
        }

        unsigned iters_run = max_iters;
        bool ok = false;
        for (unsigned iter = 0; iter < max_iters; iter++) {
            for (int c = tid; c < nc; c += nt) {                         // :269-273
                int par = 0;
                for (int j = 0; j < dc; j++) {
                    const uint64_t ent = __ldg(code.chk_tab + (size_t)j * nc + c);
                    if (ent == kNoEdge64) break;
                    par ^= bits[(int)(ent >> 32)];
                }
                cpar[c] = (uint8_t)par;
            }
            if (tid == 0) s_seen = 0;
            __syncthreads();
            unsigned seen = 0;                                           // :276-286
            for (int a = tid; a < nv; a += nt) {
                int viol = 0;
                for (int j = 0; j < dv; j++) {
                    const uint64_t ent = __ldg(code.var_tab + (size_t)j * nv + a);
                    if (ent == kNoEdge64) break;
                    viol += cpar[(int)(ent >> 32)];
                }
                cnt[a] = (uint8_t)viol;
                seen |= 1u << viol;
            }
            seen = __reduce_or_sync(0xFFFFFFFFu, seen);
            if ((tid & 31) == 0) atomicOr(&s_seen, seen);
            __syncthreads();
            const int max_viol = 31 - __clz((int)s_seen);
            if (max_viol == 0) { ok = true; iters_run = iter; break; }   // :288-289
            for (int a = tid; a < nv; a += nt)                           // :292-296
                if (cnt[a] == max_viol) bits[a] ^= 1;
            __syncthreads();
        }

        uint8_t *out = out_all + frame * (unsigned long long)out_len;
        for (int o = tid; o < out_len; o += nt) {
            unsigned byte = 0;
#pragma unroll
            for (int b = 0; b < 8; b++) byte |= (unsigned)bits[o * 8 + b] << (7 - b);
            out[o] = (uint8_t)byte;
        }
        if (tid == 0) {
            if (success) success[frame] = ok ? 1 : 0;
            if (iters_out) iters_out[frame] = iters_run;
        }
        __syncthreads();
    }
}

}  // namespace

cudaError_t launch_decode_bf(DeviceCtx &ctx, int code, const uint8_t *input, uint8_t *output, size_t batch,
                             size_t max_iters, uint8_t *success, uint32_t *iters, cudaStream_t stream) {
    if (batch == 0) return cudaSuccess;
    {   // bit-packed kernels: lane groups per codeword for the TM codes, one thread per codeword for the TC codes
        // (LABRADOR_LDPC_FORCE_GENERIC=1 keeps the table-driven kernel below)
        static const bool generic = [] { const char *e = getenv("LABRADOR_LDPC_FORCE_GENERIC"); return e && e[0] == '1'; }();
        cudaError_t err = cudaSuccess;
        if (!generic && launch_decode_bf_tm(ctx, code, input, output, batch, max_iters, success, iters, stream, &err))
            return err;
        if (!generic && launch_decode_bf_tc(ctx, code, input, output, batch, max_iters, success, iters, stream, &err))
            return err;
    }
    const DeviceCode &dc = ctx.codes[code];
    const size_t smem = (size_t)dc.vars * 2 + dc.checks;
    int threads = dc.vars < 512 ? ((dc.vars + 31) / 32) * 32 : 512;
    cudaError_t err = cudaFuncSetAttribute(decode_bf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)(smem > 64 * 1024 ? smem : 64 * 1024));
    if (err != cudaSuccess) return err;
    unsigned long long grid = batch > 0x7FFFFFFFull ? 0x7FFFFFFFull : batch;
    const unsigned mi = max_iters > 0xFFFFFFFFull ? 0xFFFFFFFFu : (unsigned)max_iters;
    decode_bf_kernel<<<(unsigned)grid, threads, smem, stream>>>(dc, input, output, (unsigned long long)batch, mi,
                                                               success, iters);
    count_launch();
    return cudaGetLastError();
}

}  // namespace ldpc
