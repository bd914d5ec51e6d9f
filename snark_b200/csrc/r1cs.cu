// K1: R1CS matrices x assignment (CSR SpMV), K3: the pointwise QAP quotient, and their composition
// into `witness_map`.
//
// Replaces, on the GPU:
//   mat_vec_mul                         /root/reference/relations/src/utils/matrix.rs:26-36
//   Sr1csAdapter::evaluate_constraint   /root/reference/relations/src/sr1cs/mod.rs:24-56
//   (out of tree) ark-groth16 LibsnarkReduction::witness_map_from_matrices, SURVEY.md Appendix A.2
// The matrices are those exported by ConstraintSystem::to_matrices()
// (/root/reference/relations/src/gr1cs/constraint_system.rs:768-804): rows may hold duplicate or unsorted
// columns; the product simply sums.  Column c reads z[c] with z = instance || witness.
//
// SpMV is the one HBM-bound kernel of the path: per nonzero 4 B column + 4 B coefficient id + a 32 B
// gather from z (40 B/nnz), plus 8 B row_ptr and a 32 B result per row.  Coefficients are
// interned like the reference's FieldInterner (relations/src/gr1cs/field_interner.rs:13-35, id 0 = ONE)
// so real circuits -- whose coefficients are almost all 1 -- skip the multiplication entirely.
#define B2S_INLINE_MUL 1   // Fr only in this unit
#include <unordered_map>

#include "ntt.cuh"
#include "r1cs.cuh"

namespace b2s {

template <class Fr>
__device__ __forceinline__ Fr fr_ld(const Fr* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    Fr r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
template <class Fr>
__device__ __forceinline__ void fr_st(Fr* p, const Fr& r) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

struct SpmvMat {
    const uint64_t* row_ptr;
    const uint32_t* col;
    const uint32_t* cid;
    void* out;
};

// One thread per (matrix, row).
template <class Fr>
__global__ void __launch_bounds__(256)
spmv_kernel(SpmvMat m0, SpmvMat m1, SpmvMat m2, const Fr* __restrict__ pool, const Fr* __restrict__ z, uint64_t n_rows) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * n_rows) return;
    const uint32_t k = (uint32_t)(t / n_rows);
    const uint64_t row = t - (uint64_t)k * n_rows;
    const SpmvMat m = k == 0 ? m0 : (k == 1 ? m1 : m2);
    const uint64_t beg = m.row_ptr[row], end = m.row_ptr[row + 1];
    Fr acc = Fr::zero();
    for (uint64_t e = beg; e < end; e++) {
        const uint32_t cid = m.cid[e];
        Fr v = fr_ld(z + m.col[e]);
        if (cid != 0) v = v * fr_ld(pool + cid);
        acc = acc + v;
    }
    fr_st(reinterpret_cast<Fr*>(m.out) + row, acc);
}

// a[n_rows + i] = z[i], i < n_instance   (input-consistency rows of the LibsnarkReduction)
template <class Fr>
__global__ void copy_instance_kernel(Fr* a, const Fr* z, uint64_t n_rows, uint64_t n_instance) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_instance) fr_st(a + n_rows + i, fr_ld(z + i));
}

// K3, first half: a[i] *= b[i]   (evaluations of A B on the coset g H)
template <class Fr>
__global__ void __launch_bounds__(256) qap_mul_kernel(Fr* a, const Fr* b, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_st(a + i, fr_ld(a + i) * fr_ld(b + i));
}
// K3, second half, on COEFFICIENTS: h[j] = q[j] * alpha - c[j] * beta   (in place over q; alpha, beta one element each)
template <class Fr>
__global__ void __launch_bounds__(256) qap_quotient_kernel(Fr* q, const Fr* c, const Fr* alpha, const Fr* beta, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_st(q + i, fr_ld(q + i) * fr_ld(alpha) - fr_ld(c + i) * fr_ld(beta));
}

// -------------------------------------------------------------------------------------------
struct Key32 {
    uint64_t w[4];
    bool operator==(const Key32& o) const { return w[0] == o.w[0] && w[1] == o.w[1] && w[2] == o.w[2] && w[3] == o.w[3]; }
};
struct Key32Hash {
    size_t operator()(const Key32& k) const {
        uint64_t h = k.w[0] * 0x9E3779B97F4A7C15ull;
        h ^= (k.w[1] + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full;
        h ^= (k.w[2] + 0x165667B1ull) * 0x9E3779B97F4A7C15ull;
        h ^= (k.w[3] + 0x27D4EB2Full) * 0xC2B2AE3D27D4EB4Full;
        return (size_t)(h ^ (h >> 29));
    }
};

template <class Curve>
static int32_t r1cs_upload_t(Ctx* c, uint64_t n_rows, uint64_t n_instance, uint64_t n_witness,
                             const uint64_t* const row_ptr[3], const uint32_t* const col[3], const void* const coeff[3],
                             b2s_r1cs** out) {
    using Fr = typename Curve::Fr;
    using FrP = typename Curve::FrP;
    const uint64_t n_vars = n_instance + n_witness;
    if (n_instance == 0) return fail(c, B2S_ERR_INVALID_ARG, "r1cs: n_instance counts the constant One and must be >= 1");
    uint64_t need = n_rows + n_instance;
    uint32_t logd = 0;
    while ((1ull << logd) < need) logd++;
    if (logd > (uint32_t)FrP::TWO_ADICITY || logd > 27)
        return fail(c, B2S_ERR_POLYNOMIAL_DEGREE_TOO_LARGE, "r1cs: domain 2^%u unsupported", logd);
    b2s_r1cs* m = new b2s_r1cs();
    m->n_rows = n_rows; m->n_instance = n_instance; m->n_witness = n_witness; m->log_domain = logd;
    // intern coefficients (host, once per circuit; byte comparisons only -- no field arithmetic)
    Key32 one;
    {
        uint32_t o[8];
        for (int i = 0; i < 8; i++) o[i] = FrP::r1(i);
        memcpy(one.w, o, 32);
    }
    std::unordered_map<Key32, uint32_t, Key32Hash> ids;
    std::vector<Key32> pool;
    pool.push_back(one);
    ids.emplace(one, 0u);
    int32_t st = B2S_OK;
    for (int k = 0; k < 3 && st == B2S_OK; k++) {
        if (row_ptr[k][0] != 0) { st = fail(c, B2S_ERR_INVALID_ARG, "r1cs: row_ptr[%d][0] != 0", k); break; }
        const uint64_t nnz = row_ptr[k][n_rows];
        m->nnz[k] = nnz;
        std::vector<uint32_t> cid(nnz);
        const Key32* vals = reinterpret_cast<const Key32*>(coeff[k]);
        for (uint64_t e = 0; e < nnz; e++) {
            if (col[k][e] >= n_vars) { st = fail(c, B2S_ERR_ASSIGNMENT_MISSING, "r1cs: column %u >= %llu variables", col[k][e], (unsigned long long)n_vars); break; }
            Key32 v;
            memcpy(&v, vals + e, 32);
            if (v == one) { cid[e] = 0; continue; }
            auto it = ids.find(v);
            if (it == ids.end()) {
                it = ids.emplace(v, (uint32_t)pool.size()).first;
                pool.push_back(v);
            }
            cid[e] = it->second;
        }
        if (st != B2S_OK) break;
        for (uint64_t r = 0; r < n_rows; r++)
            if (row_ptr[k][r + 1] < row_ptr[k][r]) { st = fail(c, B2S_ERR_INVALID_ARG, "r1cs: row_ptr[%d] not monotone", k); break; }
        if (st != B2S_OK) break;
        if ((st = m->row_ptr[k].alloc(c, (n_rows + 1) * 8)) != B2S_OK) break;
        if ((st = m->col[k].alloc(c, nnz * 4)) != B2S_OK) break;
        if ((st = m->coeff_id[k].alloc(c, nnz * 4)) != B2S_OK) break;
        cudaError_t ce = cudaMemcpyAsync(m->row_ptr[k].p, row_ptr[k], (n_rows + 1) * 8, cudaMemcpyHostToDevice, c->stream);
        if (ce == cudaSuccess && nnz) ce = cudaMemcpyAsync(m->col[k].p, col[k], nnz * 4, cudaMemcpyHostToDevice, c->stream);
        if (ce == cudaSuccess && nnz) ce = cudaMemcpyAsync(m->coeff_id[k].p, cid.data(), nnz * 4, cudaMemcpyHostToDevice, c->stream);
        const cudaError_t se = cudaStreamSynchronize(c->stream);  // cid goes out of scope
        if (ce == cudaSuccess) ce = se;
        if (ce != cudaSuccess) { st = fail(c, B2S_ERR_CUDA, "r1cs upload of matrix %d failed: %s", k, cudaGetErrorString(ce)); break; }
    }
    if (st == B2S_OK) {
        m->pool_size = (uint32_t)pool.size();
        st = m->pool.alloc(c, pool.size() * 32);
        if (st == B2S_OK) {
            cudaMemcpyAsync(m->pool.p, pool.data(), pool.size() * 32, cudaMemcpyHostToDevice, c->stream);
            if (cudaStreamSynchronize(c->stream) != cudaSuccess) st = fail(c, B2S_ERR_CUDA, "r1cs upload failed");
        }
    }
    if (st != B2S_OK) { delete m; return st; }
    *out = m;
    return B2S_OK;
}

int32_t r1cs_upload(Ctx* c, uint64_t n_rows, uint64_t n_instance, uint64_t n_witness, const uint64_t* const row_ptr[3],
                    const uint32_t* const col[3], const void* const coeff[3], b2s_r1cs** out) {
    return dispatch_curve(c, [&](auto curve) {
        return r1cs_upload_t<decltype(curve)>(c, n_rows, n_instance, n_witness, row_ptr, col, coeff, out);
    });
}

template <class Curve>
static int32_t spmv_t(Ctx* c, const b2s_r1cs* m, const void* z_dev, void* oa, void* ob, void* oc) {
    using Fr = typename Curve::Fr;
    if (m->n_rows == 0) return B2S_OK;
    SpmvMat mm[3];
    void* outs[3] = {oa, ob, oc};
    for (int k = 0; k < 3; k++)
        mm[k] = SpmvMat{m->row_ptr[k].as<uint64_t>(), m->col[k].as<uint32_t>(), m->coeff_id[k].as<uint32_t>(), outs[k]};
    B2S_LAUNCH(c, spmv_kernel<Fr>, cdiv(3 * m->n_rows, 256), 256, 0, mm[0], mm[1], mm[2], m->pool.as<Fr>(),
               reinterpret_cast<const Fr*>(z_dev), m->n_rows);
    return B2S_OK;
}

int32_t spmv_run(Ctx* c, const b2s_r1cs* m, const void* z_dev, void* oa, void* ob, void* oc) {
    return dispatch_curve(c, [&](auto curve) { return spmv_t<decltype(curve)>(c, m, z_dev, oa, ob, oc); });
}

// h = coefficients of (A B - C) / Z.  ark-groth16 evaluates a, b AND c on the coset g H, forms
// (a b - c) / Z there and interpolates: 7 transforms.  Z is the constant g^N - 1 on that coset and C has degree < N, so the
// interpolation of the c term gives back C's own coefficients: h = (cosetiNTT(a_coset * b_coset) - iNTT(c)) * Zinv, exactly,
// for every assignment (satisfying or not) -- 6 transforms, the same field elements.  With the plan's full-size tables the
// scalings are merged as well: the three inverse transforms run unscaled, the 1/N goes into the coset input scaling of a and
// b (g^j / N), Zinv into the output scaling of the closing transform, and c's 1/N * Zinv into the (HBM-bound) last kernel.
template <class Curve>
static int32_t witness_map_t(Ctx* c, const b2s_r1cs* m, const void* z_dev, void* h_dev) {
    using Fr = typename Curve::Fr;
    const uint64_t N = 1ull << m->log_domain;
    Fr* a = reinterpret_cast<Fr*>(h_dev);
    DevBuf bb, cb;
    B2S_TRY(bb.alloc(c, N * sizeof(Fr)));
    B2S_TRY(cb.alloc(c, N * sizeof(Fr)));
    Fr* b = bb.as<Fr>();
    Fr* cc = cb.as<Fr>();
    // zero the padding [n_rows, N)
    B2S_CUDA(c, cudaMemsetAsync(a + m->n_rows, 0, (N - m->n_rows) * sizeof(Fr), c->stream));
    B2S_CUDA(c, cudaMemsetAsync(b + m->n_rows, 0, (N - m->n_rows) * sizeof(Fr), c->stream));
    B2S_CUDA(c, cudaMemsetAsync(cc + m->n_rows, 0, (N - m->n_rows) * sizeof(Fr), c->stream));
    B2S_TRY(spmv_t<Curve>(c, m, z_dev, a, b, cc));
    B2S_LAUNCH(c, copy_instance_kernel<Fr>, cdiv(m->n_instance, 256), 256, 0, a, reinterpret_cast<const Fr*>(z_dev),
               m->n_rows, m->n_instance);
    NttPlan* pl = nullptr;
    B2S_TRY(ntt_get_full(c, m->log_domain, &pl));
    const uint32_t wm = pl ? NTT_M_WM : 0u;
    if (!pl) B2S_TRY(ntt_get_plan(c, m->log_domain, &pl));
    Fr* bufs[3] = {a, b, cc};
    for (Fr* v : bufs) B2S_TRY(ntt_run_mode(c, v, m->log_domain, NTT_M_INVERSE | wm));
    B2S_TRY(ntt_run_mode(c, a, m->log_domain, NTT_M_COSET | wm));
    B2S_TRY(ntt_run_mode(c, b, m->log_domain, NTT_M_COSET | wm));
    B2S_LAUNCH(c, qap_mul_kernel<Fr>, cdiv(N, 256), 256, 0, a, (const Fr*)b, N);
    B2S_TRY(ntt_run_mode(c, a, m->log_domain, NTT_M_INVERSE | NTT_M_COSET | wm));
    // composed scalings: q and c both still lack Zinv;  merged: q is finished, c lacks Zinv / N
    const Fr* alpha = reinterpret_cast<const Fr*>(wm ? pl->one : pl->zinv);
    const Fr* beta = reinterpret_cast<const Fr*>(wm ? pl->full->wm_beta : pl->zinv);
    B2S_LAUNCH(c, qap_quotient_kernel<Fr>, cdiv(N, 256), 256, 0, a, (const Fr*)cc, alpha, beta, N);
    return B2S_OK;
}

int32_t witness_map_run(Ctx* c, const b2s_r1cs* m, const void* z_dev, void* h_dev) {
    return dispatch_curve(c, [&](auto curve) { return witness_map_t<decltype(curve)>(c, m, z_dev, h_dev); });
}

}  // namespace b2s
