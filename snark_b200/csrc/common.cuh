// Library-internal plumbing shared by the .cu translation units: the context object behind
// `b2s_ctx` (include/b200snark.h), error handling that never unwinds across the C ABI, stream-ordered
// device buffers and the kernel-launch counter that bench.py reports as `gpu_launches`.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b200snark.h"
#include "curves.cuh"

namespace b2s {

struct NttPlan;  // ntt.cu
struct MsmDedupCache;  // msm.cu: classification of a scalar vector, shared by the MSMs of one proof that use the same scalars

struct Ctx {
    int curve = 0;
    int device = 0;
    cudaStream_t stream = nullptr;
    // second stream for latency-bound MSM tails (Horner) so they overlap the next MSM's bucket work
    cudaStream_t aux = nullptr;
    cudaEvent_t ev_tail = nullptr, ev_done = nullptr;
    // third stream: the witness map of a proof runs here, side by side with the MSMs that only need z (their latency-bound
    // phases -- sampling synchronisations, inversion levels, small reductions -- leave the SMs to the transforms)
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool aux_pending = false;
    std::mutex mu;
    std::string err;
    uint64_t launches = 0;
    int sm_count = 148;
    uint64_t total_mem = 0;     // device memory, queried once (cudaMemGetInfo costs milliseconds with a multi-GiB pool: not per MSM)
    std::map<uint32_t, NttPlan*> ntt_plans;   // keyed by log_n
    // optional per-kernel timing (b2s_profile_*): CUDA events around every launch on `stream`
    bool profiling = false;
    struct ProfRec { const char* name; cudaEvent_t e0, e1; };
    std::vector<ProfRec> prof;
    void* fixed_base_tables[2] = {nullptr, nullptr};  // G1 / G2 window tables (setup.cu)
    MsmDedupCache* dedup_cache = nullptr;             // non-null between msm_dedup_scope_begin / _end (one proof)
    // small buffers read by aux-stream kernels (heavy-list sums, tiny-rest products): a ring of persistent slots instead of
    // stream-ordered allocations freed on the other stream -- cross-stream frees make the pool insert dependencies between
    // the streams (observed as tens of milliseconds of main-stream stalls in some timed regions)
    static constexpr size_t AUX_SLOT_BYTES = 16384, AUX_SLOTS = 16;
    void* aux_ring = nullptr;
    uint32_t aux_ring_next = 0;
    void* aux_slot() {
        void* p = static_cast<char*>(aux_ring) + (size_t)(aux_ring_next % AUX_SLOTS) * AUX_SLOT_BYTES;
        aux_ring_next++;
        return p;
    }
};

inline int32_t fail(Ctx* c, int32_t code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
}

#define B2S_CUDA(ctx, expr)                                                                         \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return ::b2s::fail(ctx, e__ == cudaErrorMemoryAllocation ? B2S_ERR_OOM : B2S_ERR_CUDA, \
                               "%s:%d %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
    } while (0)

#define B2S_TRY(expr)                 \
    do {                              \
        int32_t s__ = (expr);         \
        if (s__ != B2S_OK) return s__; \
    } while (0)

// Launch + count + check.  Usage: B2S_LAUNCH(ctx, kernel<T>, grid, block, smem, args...)
#define B2S_LAUNCH(ctx, kern, grid, block, smem, ...) B2S_LAUNCH_N(ctx, #kern, kern, grid, block, smem, __VA_ARGS__)
#define B2S_LAUNCH_N(ctx, label, kern, grid, block, smem, ...) \
    B2S_LAUNCH_SN(ctx, (ctx)->stream, label, kern, grid, block, smem, __VA_ARGS__)
#define B2S_LAUNCH_SN(ctx, strm, label, kern, grid, block, smem, ...)                   \
    do {                                                                                \
        ::b2s::Ctx::ProfRec pr__{label, nullptr, nullptr};                              \
        if ((ctx)->profiling) {                                                         \
            cudaEventCreate(&pr__.e0);                                                  \
            cudaEventCreate(&pr__.e1);                                                  \
            cudaEventRecord(pr__.e0, (strm));                                    \
        }                                                                               \
        kern<<<(grid), (block), (smem), (strm)>>>(__VA_ARGS__);                  \
        (ctx)->launches++;                                                              \
        if ((ctx)->profiling) {                                                         \
            cudaEventRecord(pr__.e1, (strm));                                    \
            (ctx)->prof.push_back(pr__);                                                \
        }                                                                               \
        B2S_CUDA(ctx, cudaGetLastError());                                              \
    } while (0)

// Stream-ordered device allocation (cudaMallocAsync pool; freed on the same stream).
struct DevBuf {
    Ctx* ctx = nullptr;
    void* p = nullptr;
    size_t bytes = 0;
    bool secret = false;   // holds trapdoor-derived scalars: cleared before the block returns to the pool
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    int32_t alloc(Ctx* c, size_t n) {
        release();
        ctx = c;
        bytes = n;
        if (n == 0) return B2S_OK;
        cudaError_t e = cudaMallocAsync(&p, n, c->stream);
        if (e != cudaSuccess) {
            p = nullptr;
            return fail(c, e == cudaErrorMemoryAllocation ? B2S_ERR_OOM : B2S_ERR_CUDA, "cudaMallocAsync(%zu): %s", n,
                        cudaGetErrorString(e));
        }
        return B2S_OK;
    }
    void release() {
        if (p && secret) cudaMemsetAsync(p, 0, bytes, ctx->stream);
        if (p) cudaFreeAsync(p, ctx->stream);
        p = nullptr;
        bytes = 0;
    }
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

// Bring a caller buffer onto the device (copy if it lives on the host, alias if already there).
struct InBuf {
    DevBuf own;
    const void* dptr = nullptr;
    int32_t bind(Ctx* c, const void* src, size_t bytes, int32_t mem) {
        if (mem == B2S_MEM_DEVICE) { dptr = src; return B2S_OK; }
        B2S_TRY(own.alloc(c, bytes));
        if (bytes) B2S_CUDA(c, cudaMemcpyAsync(own.p, src, bytes, cudaMemcpyHostToDevice, c->stream));
        dptr = own.p;
        return B2S_OK;
    }
    template <class T>
    const T* as() const { return reinterpret_cast<const T*>(dptr); }
};

// Kernels that need more than 48 KiB of dynamic shared memory: the attribute is per device, so it is (re)applied
// on every launch path rather than cached in a process-wide flag (a second ctx on another GPU needs it too).
#define B2S_SMEM_ATTR(ctx, kern, bytes) \
    B2S_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)))

// clears a host object holding secrets when the scope ends, whichever way it ends
struct HostWipe {
    void* p;
    size_t n;
    ~HostWipe() {
        volatile unsigned char* q = reinterpret_cast<volatile unsigned char*>(p);
        for (size_t i = 0; i < n; i++) q[i] = 0;
    }
};

inline unsigned cdiv(uint64_t a, uint64_t b) { return (unsigned)((a + b - 1) / b); }

// Per-curve dispatch helper: F is a generic lambda taking a curve tag.
template <class F>
inline int32_t dispatch_curve(Ctx* c, F&& f) {
    switch (c->curve) {
        case B2S_CURVE_BLS12_381: return f(Bls12_381{});
        case B2S_CURVE_BN254: return f(Bn254{});
    }
    return fail(c, B2S_ERR_INVALID_ARG, "unknown curve id %d", c->curve);
}

// ---- entry points implemented per translation unit (all take the ctx lock in api.cu) -----------
int32_t ntt_run(Ctx* c, void* data_dev, uint32_t log_n, bool inverse, bool coset);
void ntt_free_plans(Ctx* c);
// wins_ext == nullptr: the whole MSM runs on c->stream.  Otherwise wins_ext is caller-owned scratch for the
// window sums (>= 64 XYZZ points, alive until msm_join_tails): the Horner tail is queued on c->aux and the
// caller must call msm_join_tails(c) before reading out_xyzz_dev on c->stream.
// pre (optional): bases_dev is a precomputed table [pre->nwin][pre->stride] with row w = 2^(pre->c w) * (row 0), see msm_precompute
struct MsmPre { uint32_t c, nwin; uint32_t stride; };
int32_t msm_run(Ctx* c, int group, const void* bases_dev, const void* scalars_dev, uint64_t n, bool scalars_mont,
                void* out_xyzz_dev, void* wins_ext = nullptr, const MsmPre* pre = nullptr);
// table[w * n + i] = 2^(c w) bases[i] (affine), w < nwin; picks c / nwin for n points itself and reports them in *pre
int32_t msm_precompute(Ctx* c, int group, const void* bases_dev, uint64_t n, void* table_dev, MsmPre* pre);
uint32_t msm_precompute_windows(Ctx* c, uint64_t n, uint32_t* c_out);
int32_t msm_join_tails(Ctx* c);
// a, b_g1 and b_g2 of a proof are MSMs over the SAME scalar vector: inside a scope the multiplicity-aware front end
// classifies a (pointer, length) pair once and the following MSMs reuse the lists (no second sample / count round trip)
void msm_dedup_scope_begin(Ctx* c);
void msm_dedup_scope_end(Ctx* c);
int32_t scan_counts(Ctx* c, const uint32_t* counts, uint32_t n, uint32_t L, uint32_t* offsets, uint32_t* task_off);
int32_t fixed_base_run(Ctx* c, int group, const void* scalars_dev, uint64_t n, bool mont, void* out_dev);
int32_t group_sum_to_affine(Ctx* c, int group, const void* xyzz_dev, uint32_t count, void* out_affine_dev);

}  // namespace b2s
