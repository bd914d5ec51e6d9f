// b2s_r1cs_upload_lcmap: build the device CSR of A, B, C from the constraint system's flat LcMap (lcmap.cuh has the
// reference citations and the per-row logic).  Three launches per matrix set: count, scan (msm.cu's scan kernels),
// fill; everything is HBM-bound index work and runs once per circuit.
#include <cstring>
#include <vector>

#include "common.cuh"
#include "lcmap.cuh"
#include "r1cs.cuh"

namespace b2s {

using lcmap::View;

struct LcArgs { const uint64_t* a[3]; };
struct LcOut { uint64_t* row_ptr[3]; uint32_t* col[3]; uint32_t* coeff_id[3]; };

// one thread per (matrix, row): counts[k * n_rows + row] = nonzeros make_row keeps
// total64: the same sum in 64 bits (one atomic per warp) -- the scan below is 32-bit and must not wrap unnoticed
__global__ void lcmap_count_kernel(View v, LcArgs args, uint64_t n_rows, uint32_t* __restrict__ counts, uint32_t* __restrict__ err,
                                   unsigned long long* __restrict__ total64) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * n_rows) return;
    const uint32_t k = (uint32_t)(t / n_rows);
    const uint64_t row = t - (uint64_t)k * n_rows;
    uint32_t e = 0;
    const uint32_t n = lcmap::count_row(v, args.a[k][row], &e);
    counts[t] = n;
    if (e) atomicOr(err, e);
    atomicAdd(total64, (unsigned long long)n);
}

// offsets: exclusive scan of counts over all 3 * n_rows entries (one scan for the three matrices); matrix k's entries
// start at offsets[k * n_rows]
__global__ void lcmap_fill_kernel(View v, LcArgs args, uint64_t n_rows, const uint32_t* __restrict__ offsets, LcOut out) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > 3 * n_rows) return;
    if (t == 3 * n_rows) {      // closing entries of the three row_ptr arrays
        for (uint32_t k = 0; k < 3; k++) out.row_ptr[k][n_rows] = (uint64_t)(offsets[(uint64_t)(k + 1) * n_rows] - offsets[(uint64_t)k * n_rows]);
        return;
    }
    const uint32_t k = (uint32_t)(t / n_rows);
    const uint64_t row = t - (uint64_t)k * n_rows;
    const uint32_t base = offsets[(uint64_t)k * n_rows];
    const uint32_t at = offsets[t] - base;
    out.row_ptr[k][row] = at;
    lcmap::fill_row(v, args.a[k][row], out.col[k] + at, out.coeff_id[k] + at);
}

int32_t r1cs_upload_lcmap(Ctx* c, uint64_t n_rows, uint64_t n_instance, uint64_t n_witness, const uint64_t* const args[3],
                          uint64_t n_lcs, const uint64_t* lc_offsets, const uint64_t* lc_vars, const uint32_t* lc_coeffs,
                          const void* pool, uint32_t pool_len, b2s_r1cs** out) {
    return dispatch_curve(c, [&](auto curve) -> int32_t {
        using FrP = typename decltype(curve)::FrP;
        constexpr size_t FR_BYTES = 32;
        const uint64_t n_vars = n_instance + n_witness;
        if (n_instance == 0) return fail(c, B2S_ERR_INVALID_ARG, "r1cs: n_instance counts the constant One and must be >= 1");
        if (n_lcs == 0 || !lc_offsets || lc_offsets[0] != 0) return fail(c, B2S_ERR_INVALID_ARG, "lcmap: offsets must start with 0 (LC 0 is the empty LC)");
        if (pool_len < 2 || !pool) return fail(c, B2S_ERR_INVALID_ARG, "lcmap: the interner pool holds at least ONE and -ONE");
        if (3 * n_rows >= (1ull << 32)) return fail(c, B2S_ERR_POLYNOMIAL_DEGREE_TOO_LARGE, "lcmap: too many rows");
        for (uint64_t j = 0; j < n_lcs; j++)
            if (lc_offsets[j + 1] < lc_offsets[j]) return fail(c, B2S_ERR_INVALID_ARG, "lcmap: offsets not monotone at %llu", (unsigned long long)j);
        const uint64_t total = lc_offsets[n_lcs];
        {   // pool[0] must be ONE: the SpMV kernel skips the multiplication for id 0 (sr1cs/mod.rs:42-46)
            uint32_t one[8];
            for (int i = 0; i < 8; i++) one[i] = FrP::r1(i);
            if (memcmp(pool, one, FR_BYTES) != 0) return fail(c, B2S_ERR_INVALID_ARG, "lcmap: pool[0] is not ONE (Montgomery form)");
        }
        uint32_t logd = 0;
        while ((1ull << logd) < n_rows + n_instance) logd++;
        if (logd > (uint32_t)FrP::TWO_ADICITY || logd > 27) return fail(c, B2S_ERR_POLYNOMIAL_DEGREE_TOO_LARGE, "r1cs: domain 2^%u unsupported", logd);

        // zero flags of the pool (byte comparison on the host; the pool is one entry per DISTINCT coefficient)
        std::vector<uint8_t> is_zero(pool_len);
        {
            const uint8_t* p = reinterpret_cast<const uint8_t*>(pool);
            static const uint8_t zeros[FR_BYTES] = {0};
            for (uint32_t i = 0; i < pool_len; i++) is_zero[i] = memcmp(p + (size_t)i * FR_BYTES, zeros, FR_BYTES) == 0;
        }

        b2s_r1cs* m = new b2s_r1cs();
        m->n_rows = n_rows; m->n_instance = n_instance; m->n_witness = n_witness; m->log_domain = logd;
        m->pool_size = pool_len;
        int32_t st = [&]() -> int32_t {
            DevBuf d_off, d_vars, d_coeffs, d_zero, d_args[3], d_counts, d_offsets, d_task, d_err;
            B2S_TRY(m->pool.alloc(c, (size_t)pool_len * FR_BYTES));
            B2S_CUDA(c, cudaMemcpyAsync(m->pool.p, pool, (size_t)pool_len * FR_BYTES, cudaMemcpyHostToDevice, c->stream));
            B2S_TRY(d_off.alloc(c, (n_lcs + 1) * 8));
            B2S_CUDA(c, cudaMemcpyAsync(d_off.p, lc_offsets, (n_lcs + 1) * 8, cudaMemcpyHostToDevice, c->stream));
            B2S_TRY(d_vars.alloc(c, total * 8));
            B2S_TRY(d_coeffs.alloc(c, total * 4));
            if (total) {
                B2S_CUDA(c, cudaMemcpyAsync(d_vars.p, lc_vars, total * 8, cudaMemcpyHostToDevice, c->stream));
                B2S_CUDA(c, cudaMemcpyAsync(d_coeffs.p, lc_coeffs, total * 4, cudaMemcpyHostToDevice, c->stream));
            }
            B2S_TRY(d_zero.alloc(c, pool_len));
            B2S_CUDA(c, cudaMemcpyAsync(d_zero.p, is_zero.data(), pool_len, cudaMemcpyHostToDevice, c->stream));
            LcArgs la{};
            for (int k = 0; k < 3; k++) {
                B2S_TRY(d_args[k].alloc(c, n_rows * 8));
                if (n_rows) B2S_CUDA(c, cudaMemcpyAsync(d_args[k].p, args[k], n_rows * 8, cudaMemcpyHostToDevice, c->stream));
                la.a[k] = d_args[k].as<uint64_t>();
                B2S_TRY(m->row_ptr[k].alloc(c, (n_rows + 1) * 8));
            }
            View v{d_off.as<uint64_t>(), d_vars.as<uint64_t>(), d_coeffs.as<uint32_t>(), d_zero.as<uint8_t>(), n_lcs, pool_len, n_instance, n_vars};
            const uint64_t n3 = 3 * n_rows;
            B2S_TRY(d_err.alloc(c, 16));         // [0..3] error bits, [8..15] 64-bit nonzero total
            B2S_CUDA(c, cudaMemsetAsync(d_err.p, 0, 16, c->stream));
            unsigned long long* d_total64 = reinterpret_cast<unsigned long long*>(d_err.as<uint8_t>() + 8);
            uint32_t h_tot[4] = {0, 0, 0, 0};   // offsets[0], [n_rows], [2 n_rows], [3 n_rows]
            uint32_t h_err = 0;
            unsigned long long h_total64 = 0;
            if (n_rows) {
                B2S_TRY(d_counts.alloc(c, n3 * 4));
                B2S_TRY(d_offsets.alloc(c, (n3 + 1) * 4));
                B2S_TRY(d_task.alloc(c, (n3 + 1) * 4));
                B2S_LAUNCH(c, lcmap_count_kernel, cdiv(n3, 256), 256, 0, v, la, n_rows, d_counts.as<uint32_t>(), d_err.as<uint32_t>(), d_total64);
                B2S_TRY(scan_counts(c, d_counts.as<uint32_t>(), (uint32_t)n3, 1u, d_offsets.as<uint32_t>(), d_task.as<uint32_t>()));
                for (int k = 1; k <= 3; k++)
                    B2S_CUDA(c, cudaMemcpyAsync(&h_tot[k], d_offsets.as<uint32_t>() + (uint64_t)k * n_rows, 4, cudaMemcpyDeviceToHost, c->stream));
            }
            B2S_CUDA(c, cudaMemcpyAsync(&h_err, d_err.p, 4, cudaMemcpyDeviceToHost, c->stream));
            B2S_CUDA(c, cudaMemcpyAsync(&h_total64, d_total64, 8, cudaMemcpyDeviceToHost, c->stream));
            B2S_CUDA(c, cudaStreamSynchronize(c->stream));
            if (h_err & (lcmap::ERR_NESTED_LC))
                return fail(c, B2S_ERR_INVALID_ARG, "lcmap: a linear combination refers to another one -- call finalize() (inline_all_lcs) first");
            if (h_err & lcmap::ERR_COLUMN) return fail(c, B2S_ERR_ASSIGNMENT_MISSING, "lcmap: a variable index is outside the %llu variables", (unsigned long long)n_vars);
            if (h_err) return fail(c, B2S_ERR_INVALID_ARG, "lcmap: malformed input (error bits 0x%x: 1 tag, 2 lc index, 16 coefficient id)", h_err);
            if (h_total64 >> 32) return fail(c, B2S_ERR_POLYNOMIAL_DEGREE_TOO_LARGE, "lcmap: %llu nonzeros in A, B, C together; the limit is 2^32 - 1", h_total64);
            LcOut lo{};
            for (int k = 0; k < 3; k++) {
                m->nnz[k] = (uint64_t)(h_tot[k + 1] - h_tot[k]);
                B2S_TRY(m->col[k].alloc(c, m->nnz[k] * 4));
                B2S_TRY(m->coeff_id[k].alloc(c, m->nnz[k] * 4));
                lo.row_ptr[k] = m->row_ptr[k].as<uint64_t>(); lo.col[k] = m->col[k].as<uint32_t>(); lo.coeff_id[k] = m->coeff_id[k].as<uint32_t>();
            }
            if (n_rows) {
                B2S_LAUNCH(c, lcmap_fill_kernel, cdiv(n3 + 1, 256), 256, 0, v, la, n_rows, (const uint32_t*)d_offsets.as<uint32_t>(), lo);
            } else {
                for (int k = 0; k < 3; k++) B2S_CUDA(c, cudaMemsetAsync(m->row_ptr[k].p, 0, 8, c->stream));
            }
            B2S_CUDA(c, cudaStreamSynchronize(c->stream));   // host inputs may be released by the caller after return
            return B2S_OK;
        }();
        if (st != B2S_OK) { delete m; return st; }
        *out = m;
        return B2S_OK;
    });
}

}  // namespace b2s
