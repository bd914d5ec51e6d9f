// Groth16 prover composition on the GPU: witness_map -> five MSMs -> r/s epilogue.
//
// GPU counterpart of ark-groth16 `create_proof_with_reduction` / `create_proof_with_assignment`
// (upstream crate, not in /root/reference), behind the trait method
// `SNARK::prove` (/root/reference/snark/src/lib.rs:50-54).
//
//   h      = witness_map(A, B, C, z)                                        (r1cs.cu)
//   h_acc  = MSM(h_query, h[0 .. N-1))          l_acc = MSM(l_query, z_witness)
//   a_acc  = MSM(a_query, z)   b1_acc = MSM(b_g1_query, z)   b2_acc = MSM(b_g2_query, z)
//   A  = alpha_1 + a_acc  + r delta_1          B2 = beta_2 + b2_acc + s delta_2
//   B1 = beta_1  + b1_acc + s delta_1          C  = s A + r B1 - (r s) delta_1 + l_acc + h_acc
// (ark adds query[0] separately because z[0] = 1; including index 0 in the MSM is the same group
// element.)  For multi-GPU the five MSMs are cut by base range: each rank holds a shard of the key,
// returns five XYZZ partial sums, and the join adds them before the epilogue (EC addition is not an
// NCCL reduction, so the exchange is an all-gather of 5 small points).
#include "r1cs.cuh"

namespace b2s {

template <class F>
__device__ __forceinline__ void st_pt(Affine<F>* p, const Affine<F>& v) { *p = v; }

// Threads 0, 32, 64 (three warps) run the independent scalar multiplications side by side.
template <class Curve>
__global__ void groth16_epilogue_g1_kernel(const Affine<typename Curve::Fq>* consts /*alpha,beta,delta*/,
                                           const XYZZ<typename Curve::Fq>* sums /*h,l,a,b1*/, const typename Curve::Fr* rs,
                                           Affine<typename Curve::Fq>* out_a, Affine<typename Curve::Fq>* out_c) {
    using Fq = typename Curve::Fq;
    using Fr = typename Curve::Fr;
    using P = XYZZ<Fq>;
    __shared__ P sh[3];
    const int role = threadIdx.x >> 5;
    const bool lead = (threadIdx.x & 31) == 0;
    Fr r = rs[0].from_mont(), s = rs[1].from_mont();
    Fr rsp = (rs[0] * rs[1]).from_mont();
    const P delta = P::from_affine(consts[2]);
    if (lead) {
        if (role == 0) sh[0] = scalar_mul_words(delta, r.v, Fr::N);
        if (role == 1) sh[1] = scalar_mul_words(delta, s.v, Fr::N);
        if (role == 2) sh[2] = scalar_mul_words(delta, rsp.v, Fr::N);
    }
    __syncthreads();
    P acc = P::identity();
    if (lead && role == 0) {   // A
        acc = sh[0]; acc.add(sums[2]); acc.add_affine(consts[0]);
        sh[0] = acc;
    }
    if (lead && role == 1) {   // B1
        acc = sh[1]; acc.add(sums[3]); acc.add_affine(consts[1]);
        sh[1] = acc;
    }
    __syncthreads();
    if (lead && role == 0) { P t = scalar_mul_words(sh[0], s.v, Fr::N); acc = sh[0]; sh[0] = t; *out_a = acc.to_affine(); }
    if (lead && role == 1) { sh[1] = scalar_mul_words(sh[1], r.v, Fr::N); }
    __syncthreads();
    if (lead && role == 0) {
        P cacc = sh[0];
        cacc.add(sh[1]);
        cacc.add(sh[2].neg());
        cacc.add(sums[1]);
        cacc.add(sums[0]);
        *out_c = cacc.to_affine();
    }
}

template <class Curve>
__global__ void groth16_epilogue_g2_kernel(const Affine<typename Curve::Fq2>* consts /*beta,delta*/,
                                           const XYZZ<typename Curve::Fq2>* b2_sum, const typename Curve::Fr* rs,
                                           Affine<typename Curve::Fq2>* out_b) {
    using Fq2 = typename Curve::Fq2;
    using Fr = typename Curve::Fr;
    if (threadIdx.x != 0) return;
    Fr s = rs[1].from_mont();
    XYZZ<Fq2> acc = scalar_mul_words(XYZZ<Fq2>::from_affine(consts[1]), s.v, Fr::N);
    acc.add(b2_sum[0]);
    acc.add_affine(consts[0]);
    *out_b = acc.to_affine();
}

// sums[j] = sum over shards of partials[shard * stride + j]
template <class F>
__global__ void sum_shards_kernel(const XYZZ<F>* partials, uint32_t n_shards, uint32_t stride, XYZZ<F>* sums) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= stride) return;
    XYZZ<F> acc = XYZZ<F>::identity();
    for (uint32_t sidx = 0; sidx < n_shards; sidx++) acc.add(partials[sidx * stride + j]);
    sums[j] = acc;
}

static int32_t copy_in(Ctx* c, DevBuf& dst, const void* src, size_t bytes, int32_t mem) {
    B2S_TRY(dst.alloc(c, bytes));
    if (bytes)
        B2S_CUDA(c, cudaMemcpyAsync(dst.p, src, bytes, mem == B2S_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                                    c->stream));
    return B2S_OK;
}

int32_t pk_upload(Ctx* c, const b2s_pk_desc* d, int32_t mem, b2s_pk** out) {
    const size_t fq = c->curve == B2S_CURVE_BLS12_381 ? 48 : 32;
    const size_t g1 = 2 * fq, g2 = 4 * fq;
    const uint64_t n_vars = d->n_instance + d->n_witness;
    if (d->a_off + d->a_len > n_vars || d->b1_off + d->b1_len > n_vars || d->b2_off + d->b2_len > n_vars ||
        d->l_off + d->l_len > d->n_witness || d->h_off + d->h_len > d->domain_size)
        return fail(c, B2S_ERR_MALFORMED_VK, "pk: a query range exceeds the key dimensions");
    if (!d->alpha_g1 || !d->beta_g1 || !d->delta_g1 || !d->beta_g2 || !d->delta_g2)
        return fail(c, B2S_ERR_MALFORMED_VK, "pk: missing group constants");
    b2s_pk* pk = new b2s_pk();
    pk->n_instance = d->n_instance; pk->n_witness = d->n_witness; pk->domain_size = d->domain_size;
    pk->a_off = d->a_off; pk->a_len = d->a_len; pk->b1_off = d->b1_off; pk->b1_len = d->b1_len;
    pk->b2_off = d->b2_off; pk->b2_len = d->b2_len; pk->h_off = d->h_off; pk->h_len = d->h_len;
    pk->l_off = d->l_off; pk->l_len = d->l_len;
    int32_t st = pk->consts_g1.alloc(c, 3 * g1);
    if (st == B2S_OK) st = pk->consts_g2.alloc(c, 2 * g2);
    const cudaMemcpyKind kind = mem == B2S_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (st == B2S_OK) {
        char* p1 = pk->consts_g1.as<char>();
        char* p2 = pk->consts_g2.as<char>();
        cudaMemcpyAsync(p1, d->alpha_g1, g1, kind, c->stream);
        cudaMemcpyAsync(p1 + g1, d->beta_g1, g1, kind, c->stream);
        cudaMemcpyAsync(p1 + 2 * g1, d->delta_g1, g1, kind, c->stream);
        cudaMemcpyAsync(p2, d->beta_g2, g2, kind, c->stream);
        cudaMemcpyAsync(p2 + g2, d->delta_g2, g2, kind, c->stream);
    }
    if (st == B2S_OK) st = copy_in(c, pk->a_query, d->a_query, d->a_len * g1, mem);
    if (st == B2S_OK) st = copy_in(c, pk->b_g1_query, d->b_g1_query, d->b1_len * g1, mem);
    if (st == B2S_OK) st = copy_in(c, pk->b_g2_query, d->b_g2_query, d->b2_len * g2, mem);
    if (st == B2S_OK) st = copy_in(c, pk->h_query, d->h_query, d->h_len * g1, mem);
    if (st == B2S_OK) st = copy_in(c, pk->l_query, d->l_query, d->l_len * g1, mem);
    if (st == B2S_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) st = fail(c, B2S_ERR_CUDA, "pk upload failed");
    if (st != B2S_OK) { delete pk; return st; }
    *out = pk;
    return B2S_OK;
}

template <class Curve>
static int32_t shard_t(Ctx* c, const b2s_pk* pk, const b2s_r1cs* m, const void* z_inst, const void* z_wit, const void* z_dev,
                       void* g1_out, void* g2_out) {
    using Fr = typename Curve::Fr;
    using P1 = XYZZ<typename Curve::Fq>;
    const uint64_t N = 1ull << m->log_domain;
    if (pk->n_instance != m->n_instance || pk->n_witness != m->n_witness || pk->domain_size != N)
        return fail(c, B2S_ERR_ASSIGNMENT_MISSING, "prove: key (%llu,%llu,%llu) does not match matrices (%llu,%llu,%llu)",
                    (unsigned long long)pk->n_instance, (unsigned long long)pk->n_witness, (unsigned long long)pk->domain_size,
                    (unsigned long long)m->n_instance, (unsigned long long)m->n_witness, (unsigned long long)N);
    const uint64_t n_vars = m->n_instance + m->n_witness;
    DevBuf z, h;
    B2S_TRY(h.alloc(c, N * sizeof(Fr)));
    const Fr* zd = reinterpret_cast<const Fr*>(z_dev);
    if (!zd) {
        B2S_TRY(z.alloc(c, n_vars * sizeof(Fr)));
        Fr* zw = z.as<Fr>();
        B2S_CUDA(c, cudaMemcpyAsync(zw, z_inst, m->n_instance * sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
        if (m->n_witness)
            B2S_CUDA(c, cudaMemcpyAsync(zw + m->n_instance, z_wit, m->n_witness * sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
        zd = zw;
    }
    B2S_TRY(witness_map_run(c, m, zd, h.p));
    P1* g1 = reinterpret_cast<P1*>(g1_out);
    B2S_TRY(msm_run(c, 1, pk->h_query.p, h.as<Fr>() + pk->h_off, pk->h_len, true, g1 + 0));
    B2S_TRY(msm_run(c, 1, pk->l_query.p, zd + m->n_instance + pk->l_off, pk->l_len, true, g1 + 1));
    B2S_TRY(msm_run(c, 1, pk->a_query.p, zd + pk->a_off, pk->a_len, true, g1 + 2));
    B2S_TRY(msm_run(c, 1, pk->b_g1_query.p, zd + pk->b1_off, pk->b1_len, true, g1 + 3));
    B2S_TRY(msm_run(c, 2, pk->b_g2_query.p, zd + pk->b2_off, pk->b2_len, true, g2_out));
    return B2S_OK;
}

int32_t groth16_shard(Ctx* c, const b2s_pk* pk, const b2s_r1cs* m, const void* z_inst, const void* z_wit, const void* z_dev,
                      void* g1_out, void* g2_out) {
    return dispatch_curve(c, [&](auto curve) { return shard_t<decltype(curve)>(c, pk, m, z_inst, z_wit, z_dev, g1_out, g2_out); });
}

template <class Curve>
static int32_t finish_t(Ctx* c, const b2s_pk* pk, const void* g1_partials, const void* g2_partials, uint32_t n_shards,
                        const void* r_host, const void* s_host, void* out_a, void* out_b, void* out_c) {
    using Fr = typename Curve::Fr;
    using Fq = typename Curve::Fq;
    using Fq2 = typename Curve::Fq2;
    DevBuf rs, sums1, sums2, outs;
    B2S_TRY(rs.alloc(c, 2 * sizeof(Fr)));
    B2S_CUDA(c, cudaMemcpyAsync(rs.p, r_host, sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
    B2S_CUDA(c, cudaMemcpyAsync(rs.as<Fr>() + 1, s_host, sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
    B2S_TRY(sums1.alloc(c, 4 * sizeof(XYZZ<Fq>)));
    B2S_TRY(sums2.alloc(c, sizeof(XYZZ<Fq2>)));
    B2S_LAUNCH(c, sum_shards_kernel<Fq>, 1, 32, 0, reinterpret_cast<const XYZZ<Fq>*>(g1_partials), n_shards, 4u, sums1.as<XYZZ<Fq>>());
    B2S_LAUNCH(c, sum_shards_kernel<Fq2>, 1, 32, 0, reinterpret_cast<const XYZZ<Fq2>*>(g2_partials), n_shards, 1u, sums2.as<XYZZ<Fq2>>());
    const size_t g1 = sizeof(Affine<Fq>), g2 = sizeof(Affine<Fq2>);
    B2S_TRY(outs.alloc(c, 2 * g1 + g2));
    char* o = outs.as<char>();
    B2S_LAUNCH(c, groth16_epilogue_g1_kernel<Curve>, 1, 96, 0, pk->consts_g1.as<Affine<Fq>>(), sums1.as<XYZZ<Fq>>(), rs.as<Fr>(),
               reinterpret_cast<Affine<Fq>*>(o), reinterpret_cast<Affine<Fq>*>(o + g1));
    B2S_LAUNCH(c, groth16_epilogue_g2_kernel<Curve>, 1, 32, 0, pk->consts_g2.as<Affine<Fq2>>(), sums2.as<XYZZ<Fq2>>(), rs.as<Fr>(),
               reinterpret_cast<Affine<Fq2>*>(o + 2 * g1));
    B2S_CUDA(c, cudaMemcpyAsync(out_a, o, g1, cudaMemcpyDeviceToHost, c->stream));
    B2S_CUDA(c, cudaMemcpyAsync(out_c, o + g1, g1, cudaMemcpyDeviceToHost, c->stream));
    B2S_CUDA(c, cudaMemcpyAsync(out_b, o + 2 * g1, g2, cudaMemcpyDeviceToHost, c->stream));
    B2S_CUDA(c, cudaStreamSynchronize(c->stream));
    return B2S_OK;
}

int32_t groth16_finish(Ctx* c, const b2s_pk* pk, const void* g1_partials_dev, const void* g2_partials_dev, uint32_t n_shards,
                       const void* r_host, const void* s_host, void* out_a, void* out_b, void* out_c) {
    return dispatch_curve(c, [&](auto curve) {
        return finish_t<decltype(curve)>(c, pk, g1_partials_dev, g2_partials_dev, n_shards, r_host, s_host, out_a, out_b, out_c);
    });
}

}  // namespace b2s
