// Generic batched min-sum decoder: any code, any LLR type.
//
// Replaces LDPCCode::decode_ms<T> (reference src/decoder.rs:347-475) for every
// (code, T) pair that has no specialised kernel.  One CTA decodes one codeword
// at a time.  The reference's two edge loops per iteration are re-expressed as
//   phase A (one thread per variable):  recompute u for the variable's edges
//            from the previous v and the previous per-check (min1, min2, sign)
//            (src/decoder.rs:391-405), accumulate the marginal with saturating
//            adds in ascending edge order (:408 -- the only order-sensitive
//            operation), then produce the self-corrected new v for each edge
//            (:421-426);
//   phase B (one thread per check):  min1/min2 of |v|, sign product and the
//            parity of the marginals' hard bits (:430-447), followed by the
//            all-parities-zero exit test (:453).
// u is never stored: it is a pure function of the old v and the old per-check
// state, so the state is v[E] + min1[C] + min2[C] + sign[C] (+ hard bits).
// v lives in shared memory when it fits, else in an L2-resident global scratch
// slot owned by the CTA.
#include <cuda_runtime.h>

#include "front.cuh"
#include "llr_arith.cuh"
#include "runtime.h"

namespace ldpc {
namespace {

constexpr int kThreads = 512;
constexpr int kMaxVarDeg = 6;

struct MsLayout {
    // byte offsets into dynamic shared memory
    unsigned v_off, min1_off, min2_off, llr_off, sgn_off, hb_off, total;
    bool v_in_smem, mm_in_smem, llr_in_smem;
    unsigned slot_elems;      // elements of T per persistent CTA in the global scratch: v (edges) and / or min1, min2 (2 * checks)
};

template <class T, int FRONT = kFrontNone>
__global__ void __launch_bounds__(kThreads)
decode_ms_generic_kernel(const DeviceCode code, const MsLayout lay,
                         const typename FrontSrc<FRONT, T>::type *__restrict__ llrs_all,
                         uint8_t *__restrict__ out_all, unsigned long long batch, unsigned max_iters,
                         uint8_t *__restrict__ success, uint32_t *__restrict__ iters_out,
                         T *__restrict__ vscratch, const float fscale, const float flimit) {
    typedef Arith<T> A;
    extern __shared__ __align__(16) unsigned char smem[];
    T *slot = vscratch + (size_t)blockIdx.x * lay.slot_elems;
    T *v = lay.v_in_smem ? reinterpret_cast<T *>(smem + lay.v_off) : slot;
    T *mm = slot + (lay.v_in_smem ? 0 : code.edges);
    T *min1 = lay.mm_in_smem ? reinterpret_cast<T *>(smem + lay.min1_off) : mm;
    T *min2 = lay.mm_in_smem ? reinterpret_cast<T *>(smem + lay.min2_off) : mm + code.checks;
    T *llr_s = reinterpret_cast<T *>(smem + lay.llr_off);
    uint8_t *sgn = smem + lay.sgn_off;
    uint8_t *hb = smem + lay.hb_off;

    const int tid = threadIdx.x;
    const int n = code.n, nv = code.vars, nc = code.checks, ne = code.edges;
    const int dv = code.max_var_degree, dc = code.max_check_degree;
    const int out_len = nv / 8;

    for (unsigned long long frame = blockIdx.x; frame < batch; frame += gridDim.x) {
        const typename FrontSrc<FRONT, T>::type *llr_g =
            llrs_all + frame * (unsigned long long)(FRONT == kFrontHard ? n / 8 : n);   // front.cuh
        // Zero-initialised state, every call (reference :368, :374).
        for (int i = tid; i < ne; i += kThreads) v[i] = A::zero();
        for (int i = tid; i < nc; i += kThreads) { min1[i] = A::zero(); min2[i] = A::zero(); sgn[i] = 0; }
        for (int i = tid; i < nv; i += kThreads) hb[i] = 0;
        if (lay.llr_in_smem)
            for (int i = tid; i < n; i += kThreads) llr_s[i] = (T)front_load<FRONT, T>(llr_g, i, fscale, flimit);
        __syncthreads();

        unsigned iters_run = max_iters;
        bool ok = false;
        for (unsigned iter = 0; iter < max_iters; iter++) {
            // ---- phase A: variables ----
            for (int a = tid; a < nv; a += kThreads) {
                T va = A::zero();                                           // :382-383
                if (a < n) va = lay.llr_in_smem ? llr_s[a] : (T)front_load<FRONT, T>(llr_g, a, fscale, flimit);
                T ul[kMaxVarDeg];
                T vold[kMaxVarDeg];
#pragma unroll
                for (int j = 0; j < kMaxVarDeg; j++) {
                    ul[j] = A::zero(); vold[j] = A::zero();
                    if (j < dv) {
                        const uint64_t ent = __ldg(code.var_tab + (size_t)j * nv + a);
                        if (ent != kNoEdge64) {
                            const int idx = (int)(uint32_t)ent, c = (int)(ent >> 32);
                            const T vv = v[idx];
                            T u = (A::abs(vv) == min1[c]) ? min2[c] : min1[c];   // :391-395
                            if (sgn[c]) u = A::neg(u);                           // :398-400
                            if (A::hard_bit(vv)) u = A::neg(u);                  // :403-405
                            va = A::sat_add(va, u);                              // :408
                            ul[j] = u; vold[j] = vv;
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < kMaxVarDeg; j++) {
                    if (j < dv) {
                        const uint64_t ent = __ldg(code.var_tab + (size_t)j * nv + a);
                        if (ent != kNoEdge64) {
                            const int idx = (int)(uint32_t)ent;
                            const T nvv = A::sat_sub(va, ul[j]);                 // :421
                            const bool keep = (A::hard_bit(nvv) == A::hard_bit(vold[j])) || (vold[j] == A::zero());
                            v[idx] = keep ? nvv : A::zero();                     // :422-426
                        }
                    }
                }
                hb[a] = A::hard_bit(va) ? 1 : 0;
            }
            __syncthreads();
            // ---- phase B: checks ----
            int par_any = 0;
            for (int c = tid; c < nc; c += kThreads) {
                T m1 = A::maxval(), m2 = A::maxval();                       // :414-415
                int s = 0, par = 0;
                for (int j = 0; j < dc; j++) {
                    const uint64_t ent = __ldg(code.chk_tab + (size_t)j * nc + c);
                    if (ent == kNoEdge64) break;
                    const int idx = (int)(uint32_t)ent, var = (int)(ent >> 32);
                    const T vv = v[idx];
                    const T av = A::abs(vv);
                    if (av < m1) { m2 = m1; m1 = av; }                      // :430-435
                    else if (av < m2) { m2 = av; }
                    s ^= A::hard_bit(vv) ? 1 : 0;                           // :439-441
                    par ^= hb[var];                                         // :445-447
                }
                min1[c] = m1; min2[c] = m2; sgn[c] = (uint8_t)s;
                par_any |= par;
            }
            if (__syncthreads_or(par_any) == 0) {                           // :453
                ok = true; iters_run = iter;                                // :462
                break;
            }
        }
        // Hard decisions of all n+p marginals, MSB first (:455-461, :466-473).
        uint8_t *out = out_all + frame * (unsigned long long)out_len;
        for (int o = tid; o < out_len; o += kThreads) {
            unsigned byte = 0;
#pragma unroll
            for (int b = 0; b < 8; b++) byte |= (unsigned)hb[o * 8 + b] << (7 - b);
            out[o] = (uint8_t)byte;
        }
        if (tid == 0) {
            if (success) success[frame] = ok ? 1 : 0;
            if (iters_out) iters_out[frame] = iters_run;
        }
        __syncthreads();
    }
}

inline unsigned align16(unsigned x) { return (x + 15u) & ~15u; }

template <class T>
MsLayout make_layout(const DeviceCode &c, int max_smem) {
    MsLayout l{};
    const unsigned ts = sizeof(T);
    // per-check sign bytes and per-variable hard-decision bytes always sit in shared memory; the per-check minima, the
    // messages and the LLRs join them in this order as far as they fit (the k = 16384 codes keep f32 / f64 minima, and
    // every code beyond TM2048 its messages, in an L2-resident slot per persistent CTA)
    const unsigned base = align16(c.checks) + align16(c.vars);
    const unsigned mbytes = align16(c.checks * ts) * 2;
    const unsigned vbytes = align16(c.edges * ts);
    const unsigned lbytes = align16(c.n * ts);
    l.mm_in_smem = base + mbytes <= (unsigned)max_smem;
    const unsigned fixed = base + (l.mm_in_smem ? mbytes : 0);
    l.v_in_smem = l.mm_in_smem && fixed + vbytes <= (unsigned)max_smem;
    l.llr_in_smem = fixed + (l.v_in_smem ? vbytes : 0) + lbytes <= (unsigned)max_smem;
    l.slot_elems = (l.v_in_smem ? 0u : (unsigned)c.edges) + (l.mm_in_smem ? 0u : 2u * (unsigned)c.checks);
    unsigned off = 0;
    l.v_off = off; if (l.v_in_smem) off += vbytes;
    l.min1_off = off; if (l.mm_in_smem) off += align16(c.checks * ts);
    l.min2_off = off; if (l.mm_in_smem) off += align16(c.checks * ts);
    l.llr_off = off; if (l.llr_in_smem) off += lbytes;
    l.sgn_off = off; off += align16(c.checks);
    l.hb_off = off; off += align16(c.vars);
    l.total = off;
    return l;
}

template <class T, int FRONT = kFrontNone>
cudaError_t launch_generic(DeviceCtx &ctx, int code, const void *llrs, uint8_t *output, size_t batch,
                           size_t max_iters, uint8_t *success, uint32_t *iters, cudaStream_t stream,
                           const Front &front = Front()) {
    const DeviceCode &dc = ctx.codes[code];
    const MsLayout lay = make_layout<T>(dc, ctx.max_smem_optin);
    auto kern = decode_ms_generic_kernel<T, FRONT>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.total);
    if (err != cudaSuccess) return err;
    unsigned long long grid = batch;
    T *scratch = nullptr;
    if (!lay.v_in_smem) {
        // persistent CTAs, one global message slot each
        grid = (unsigned long long)ctx.sm_count;
        if (grid > batch) grid = batch;
        const size_t need = (size_t)grid * lay.slot_elems * sizeof(T);
        if (ctx.vscratch_bytes < need) {
            if (ctx.vscratch) cudaFree(ctx.vscratch);
            ctx.vscratch = nullptr; ctx.vscratch_bytes = 0;
            err = cudaMalloc(&ctx.vscratch, need);
            if (err != cudaSuccess) return err;
            ctx.vscratch_bytes = need;
        }
        scratch = static_cast<T *>(ctx.vscratch);
        // the slots are shared by every launch that needs them: order this one behind the previous user, whatever
        // stream that ran on (as decode_bf_tc.cu does for its retry list)
        if (ctx.vscratch_done) err = cudaStreamWaitEvent(stream, ctx.vscratch_done, 0);
        else err = cudaEventCreateWithFlags(&ctx.vscratch_done, cudaEventDisableTiming);
        if (err != cudaSuccess) return err;
    }
    if (grid > 0x7FFFFFFFull) grid = 0x7FFFFFFFull;
    const unsigned mi = max_iters > 0xFFFFFFFFull ? 0xFFFFFFFFu : (unsigned)max_iters;
    kern<<<(unsigned)grid, kThreads, lay.total, stream>>>(dc, lay, static_cast<const typename FrontSrc<FRONT, T>::type *>(llrs),
                                                          output, (unsigned long long)batch, mi, success, iters, scratch,
                                                          front.scale, front.limit);
    count_launch();
    if (scratch) cudaEventRecord(ctx.vscratch_done, stream);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_decode_ms_generic(DeviceCtx &ctx, int code, int llr_type, const void *llrs, uint8_t *output,
                                     size_t batch, size_t max_iters, uint8_t *success, uint32_t *iters,
                                     cudaStream_t stream, const Front &front) {
    if (front.kind == kFrontSoftF32 && llr_type == kI8)
        return launch_generic<int8_t, kFrontSoftF32>(ctx, code, llrs, output, batch, max_iters, success, iters, stream, front);
    if (front.kind == kFrontSoftF32 && llr_type == kI16)
        return launch_generic<int16_t, kFrontSoftF32>(ctx, code, llrs, output, batch, max_iters, success, iters, stream, front);
    if (front.kind == kFrontHard && llr_type == kI8)
        return launch_generic<int8_t, kFrontHard>(ctx, code, llrs, output, batch, max_iters, success, iters, stream, front);
    if (front.kind != kFrontNone) return cudaErrorInvalidValue;
    switch (llr_type) {
        case kI8: return launch_generic<int8_t>(ctx, code, llrs, output, batch, max_iters, success, iters, stream);
        case kI16: return launch_generic<int16_t>(ctx, code, llrs, output, batch, max_iters, success, iters, stream);
        case kI32: return launch_generic<int32_t>(ctx, code, llrs, output, batch, max_iters, success, iters, stream);
        case kF32: return launch_generic<float>(ctx, code, llrs, output, batch, max_iters, success, iters, stream);
        case kF64: return launch_generic<double>(ctx, code, llrs, output, batch, max_iters, success, iters, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace ldpc
