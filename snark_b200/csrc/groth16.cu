// Groth16 prover composition on the GPU: witness_map -> five MSMs -> r/s epilogue.
//
// GPU counterpart of ark-groth16 `create_proof_with_reduction` / `create_proof_with_assignment`
// (upstream crate, not in /root/reference), behind the trait method
// `SNARK::prove` (/root/reference/snark/src/lib.rs:50-54).
//
//   h      = witness_map(A, B, C, z)                                        (r1cs.cu)
//   h_acc  = MSM(h_query, h[0 .. N-1))          l_acc = MSM(l_query, z_witness)
//   a_acc  = MSM(a_query ++ [delta_1, O], z ++ [r, s])      = sum z_j a_j + r delta_1
//   b1_acc = MSM(b_g1_query ++ [O, delta_1], z ++ [r, s])   = sum z_j b_j + s delta_1
//   b2_acc = MSM(b_g2_query ++ [O, delta_2], z ++ [r, s])   = sum z_j b_j + s delta_2      (G2)
//   A = alpha_1 + a_acc      B1 = beta_1 + b1_acc      B2 = beta_2 + b2_acc
//   C = s A + r B1 - (r s) delta_1 + l_acc + h_acc
// ark adds query[0] and r*delta / s*delta separately (z[0] = 1, fresh r, s); putting them into the MSMs as
// two extra (base, scalar) pairs is the same group element and removes three 255-bit scalar multiplications
// (one of them in G2) from the latency-bound tail.  What remains serial is s*A, r*B1, (rs)*delta_1, run
// side by side in three warps.
//
// Multi-GPU: the five MSMs are cut by base range; each rank holds a shard of the key and returns five XYZZ
// partial sums; the rank that owns the END of a query range also owns its two extra pairs.  The join adds
// the partials (EC addition is not an NCCL reduction, so the exchange is an all-gather of 5 small points)
// and applies the epilogue.
#include "ec_team.cuh"
#include "r1cs.cuh"

namespace b2s {

// Warps 0, 1, 2 run the three remaining scalar multiplications side by side (4-bit windows, four-lane teams of
// ec_team.cuh: a doubling costs 3 multiplication latencies instead of 9); warp 3 normalises A meanwhile.
template <class Curve>
__global__ void __launch_bounds__(128)
groth16_epilogue_g1_kernel(const Affine<typename Curve::Fq>* consts /*alpha,beta,delta*/, const XYZZ<typename Curve::Fq>* sums /*h,l,a,b1*/,
                           const typename Curve::Fr* rs, Affine<typename Curve::Fq>* out_a, Affine<typename Curve::Fq>* out_c) {
    using Fq = typename Curve::Fq;
    using Fr = typename Curve::Fr;
    using P = XYZZ<Fq>;
    __shared__ P sh[3];
    __shared__ P table[3][16];
    const int role = threadIdx.x >> 5;
    const bool lead = (threadIdx.x & 31) == 0;
    if (role == 0) {   // s * A,  A = alpha + a_acc
        P a = sums[2];
        a.add_affine(consts[0]);
        Fr s = rs[1].from_mont();
        P v = team_scalar_mul(a, s.v, Fr::N, table[0]);
        if (lead) sh[0] = v;
    } else if (role == 1) {   // r * B1,  B1 = beta + b1_acc
        P b = sums[3];
        b.add_affine(consts[1]);
        Fr r = rs[0].from_mont();
        P v = team_scalar_mul(b, r.v, Fr::N, table[1]);
        if (lead) sh[1] = v;
    } else if (role == 2) {   // (r s) * delta
        Fr rsp = (rs[0] * rs[1]).from_mont();
        P v = team_scalar_mul(P::from_affine(consts[2]), rsp.v, Fr::N, table[2]);
        if (lead) sh[2] = v;
    } else {   // A itself, normalised (one inversion) while the others multiply
        P a = sums[2];
        a.add_affine(consts[0]);
        if (lead) *out_a = a.to_affine();
    }
    __syncthreads();
    if (role == 0) {
        P c = sh[0];
        team_add(c, sh[1]);
        team_add(c, sh[2].neg());
        team_add(c, sums[1]);
        team_add(c, sums[0]);
        if (lead) *out_c = c.to_affine();
    }
}

template <class Curve>
__global__ void groth16_epilogue_g2_kernel(const Affine<typename Curve::Fq2>* consts /*beta,delta*/,
                                           const XYZZ<typename Curve::Fq2>* b2_sum, Affine<typename Curve::Fq2>* out_b) {
    if (threadIdx.x != 0) return;
    XYZZ<typename Curve::Fq2> acc = b2_sum[0];
    acc.add_affine(consts[0]);
    *out_b = acc.to_affine();
}

// sums[j] = sum over shards of partials[shard * stride + j], j < count
template <class F>
__global__ void sum_shards_kernel(const XYZZ<F>* partials, uint32_t n_shards, uint32_t stride, uint32_t count, XYZZ<F>* sums) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    XYZZ<F> acc = XYZZ<F>::identity();
    for (uint32_t sidx = 0; sidx < n_shards; sidx++) acc.add(partials[sidx * stride + j]);
    sums[j] = acc;
}

// query ++ extras: `len` points from src, then (if this shard owns the end of the range) two extra points
static int32_t copy_query(Ctx* c, DevBuf& dst, const void* src, uint64_t len, size_t pt, int32_t mem, const void* extra0,
                          const void* extra1, cudaMemcpyKind kind) {
    B2S_TRY(dst.alloc(c, (len + 2) * pt));
    char* d = dst.as<char>();
    if (len) B2S_CUDA(c, cudaMemcpyAsync(d, src, len * pt, kind, c->stream));
    B2S_CUDA(c, cudaMemsetAsync(d + len * pt, 0, 2 * pt, c->stream));   // O = all-zero bytes
    if (extra0) B2S_CUDA(c, cudaMemcpyAsync(d + len * pt, extra0, pt, kind, c->stream));
    if (extra1) B2S_CUDA(c, cudaMemcpyAsync(d + (len + 1) * pt, extra1, pt, kind, c->stream));
    (void)mem;
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
    pk->a_ext = (d->a_off + d->a_len == n_vars) ? 2 : 0;
    pk->b1_ext = (d->b1_off + d->b1_len == n_vars) ? 2 : 0;
    pk->b2_ext = (d->b2_off + d->b2_len == n_vars) ? 2 : 0;
    const cudaMemcpyKind kind = mem == B2S_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    auto body = [&]() -> int32_t {
        B2S_TRY(pk->consts_g1.alloc(c, 3 * g1));
        B2S_TRY(pk->consts_g2.alloc(c, 2 * g2));
        char* p1 = pk->consts_g1.as<char>();
        char* p2 = pk->consts_g2.as<char>();
        B2S_CUDA(c, cudaMemcpyAsync(p1, d->alpha_g1, g1, kind, c->stream));
        B2S_CUDA(c, cudaMemcpyAsync(p1 + g1, d->beta_g1, g1, kind, c->stream));
        B2S_CUDA(c, cudaMemcpyAsync(p1 + 2 * g1, d->delta_g1, g1, kind, c->stream));
        B2S_CUDA(c, cudaMemcpyAsync(p2, d->beta_g2, g2, kind, c->stream));
        B2S_CUDA(c, cudaMemcpyAsync(p2 + g2, d->delta_g2, g2, kind, c->stream));
        // extras: a: [delta_1, O] (scalars r, s)   b1: [O, delta_1]   b2: [O, delta_2]
        B2S_TRY(copy_query(c, pk->a_query, d->a_query, d->a_len, g1, mem, pk->a_ext ? d->delta_g1 : nullptr, nullptr, kind));
        B2S_TRY(copy_query(c, pk->b_g1_query, d->b_g1_query, d->b1_len, g1, mem, nullptr, pk->b1_ext ? d->delta_g1 : nullptr, kind));
        B2S_TRY(copy_query(c, pk->b_g2_query, d->b_g2_query, d->b2_len, g2, mem, nullptr, pk->b2_ext ? d->delta_g2 : nullptr, kind));
        B2S_TRY(copy_query(c, pk->h_query, d->h_query, d->h_len, g1, mem, nullptr, nullptr, kind));
        B2S_TRY(copy_query(c, pk->l_query, d->l_query, d->l_len, g1, mem, nullptr, nullptr, kind));
        // Fixed-base window table for the h query: its scalars (the quotient polynomial) are never repeated values, so this is
        // the MSM that always pays the full Pippenger price; the other queries run over the witness, where the multiplicity-aware
        // front end usually leaves little.  13 x the query (18 GiB at 2^24): only when it fits comfortably.
        {
            const char* env = getenv("B2S_PK_PRECOMP");
            const uint64_t min_n = getenv("B2S_PK_PRECOMP_MIN") ? strtoull(getenv("B2S_PK_PRECOMP_MIN"), nullptr, 10) : (1ull << 18);
            uint32_t cc = 0;
            const uint32_t nw = msm_precompute_windows(c, d->h_len, &cc);
            size_t free_b = 0, total_b = 0;
            cudaMemGetInfo(&free_b, &total_b);
            const uint64_t need = (uint64_t)nw * d->h_len * g1;
            if (!(env && env[0] == '0') && d->h_len >= min_n && (uint64_t)nw * d->h_len < (1ull << 31) && need * 4 < (uint64_t)free_b) {
                B2S_TRY(pk->h_table.alloc(c, need));
                B2S_TRY(msm_precompute(c, 1, pk->h_query.p, d->h_len, pk->h_table.p, &pk->h_pre));
            }
        }
        B2S_CUDA(c, cudaStreamSynchronize(c->stream));
        return B2S_OK;
    };
    const int32_t st = body();
    if (st != B2S_OK) { delete pk; return st; }
    *out = pk;
    return B2S_OK;
}

// the whole h on this GPU
struct ReplicatedH : HSource {
    DevBuf h;
    int32_t get(Ctx* c, const b2s_pk* pk, const b2s_r1cs* m, const void* z_dev, const void** h_for_shard) override {
        B2S_TRY(h.alloc(c, (size_t)32 << m->log_domain));
        B2S_TRY(witness_map_run(c, m, z_dev, h.p));
        *h_for_shard = h.as<char>() + pk->h_off * 32;
        return B2S_OK;
    }
};

template <class Curve>
static int32_t shard_t(Ctx* c, const b2s_pk* pk, const b2s_r1cs* m, const void* z_inst, const void* z_wit, const void* z_dev,
                       const void* r_host, const void* s_host, void* g1_out, void* g2_out, HSource* hs) {
    using Fr = typename Curve::Fr;
    using P1 = XYZZ<typename Curve::Fq>;
    using P2 = XYZZ<typename Curve::Fq2>;
    const uint64_t N = 1ull << m->log_domain;
    if (pk->n_instance != m->n_instance || pk->n_witness != m->n_witness || pk->domain_size != N)
        return fail(c, B2S_ERR_ASSIGNMENT_MISSING, "prove: key (%llu,%llu,%llu) does not match matrices (%llu,%llu,%llu)",
                    (unsigned long long)pk->n_instance, (unsigned long long)pk->n_witness, (unsigned long long)pk->domain_size,
                    (unsigned long long)m->n_instance, (unsigned long long)m->n_witness, (unsigned long long)N);
    const uint64_t n_vars = m->n_instance + m->n_witness;
    // z_ext = z ++ [r, s]
    DevBuf z, tails;
    ReplicatedH replicated;
    if (!hs) hs = &replicated;
    B2S_TRY(z.alloc(c, (n_vars + 2) * sizeof(Fr)));
    B2S_TRY(tails.alloc(c, 4 * 64 * sizeof(P1) + 64 * sizeof(P2)));
    // Horner tails run on c->aux and read `tails` / write the outputs: whatever way this function is left, the aux
    // stream must be done before the buffers above (and the caller's outputs) go back to the pool
    struct AuxGuard {
        Ctx* c;
        ~AuxGuard() {
            if (c->aux_pending) {
                cudaStreamSynchronize(c->aux);
                c->aux_pending = false;
            }
        }
    } aux_guard{c};
    Fr* zd = z.as<Fr>();
    if (z_dev) {
        B2S_CUDA(c, cudaMemcpyAsync(zd, z_dev, n_vars * sizeof(Fr), cudaMemcpyDeviceToDevice, c->stream));
    } else {
        B2S_CUDA(c, cudaMemcpyAsync(zd, z_inst, m->n_instance * sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
        if (m->n_witness)
            B2S_CUDA(c, cudaMemcpyAsync(zd + m->n_instance, z_wit, m->n_witness * sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
    }
    B2S_CUDA(c, cudaMemcpyAsync(zd + n_vars, r_host, sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
    B2S_CUDA(c, cudaMemcpyAsync(zd + n_vars + 1, s_host, sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
    P1* g1 = reinterpret_cast<P1*>(g1_out);
    P1* w1 = tails.as<P1>();
    void* w2 = w1 + 4 * 64;
    // Fork: the witness map (SpMV, six transforms, quotient -- or its distributed form) is enqueued on the side stream and
    // runs side by side with the four MSMs that only need z; the h-query MSM joins them.  The ctx's launch stream is swapped
    // for the duration of the enqueue (everything below the C ABI launches and allocates on c->stream).
    const void* h_shard = nullptr;
    // OFF by default: measured no gain at domain 2^24 (both sides are fmaheavy-bound, the transforms just run slower next to
    // the MSM kernels: 208 vs 204 ms per proof) and the host-buffer path got slower; B2S_SIDE_STREAM=1 enables it
    const bool fork = getenv("B2S_SIDE_STREAM") != nullptr;
    struct SideGuard {   // an error return must not leave work in flight on the side stream over buffers being released
        Ctx* c; cudaStream_t main; bool active;
        ~SideGuard() { if (active) { c->stream = main; cudaStreamSynchronize(c->side); } }
    } side_guard{c, c->stream, false};
    if (fork) {
        B2S_CUDA(c, cudaEventRecord(c->ev_fork, c->stream));
        B2S_CUDA(c, cudaStreamWaitEvent(c->side, c->ev_fork, 0));
        side_guard.active = true;
        c->stream = c->side;
        const int32_t st = hs->get(c, pk, m, zd, &h_shard);
        c->stream = side_guard.main;
        if (st != B2S_OK) return st;
        B2S_CUDA(c, cudaEventRecord(c->ev_join, c->side));
    }
    // the MSMs that only need z (their Horner tails run on the aux stream under the following work); G2 first because its
    // tail is the longest
    // a, b_g1, b_g2 (and l, when its range coincides) run over the same scalars: classify them once (msm.cu)
    struct DedupScope {
        Ctx* c;
        explicit DedupScope(Ctx* ctx) : c(ctx) { msm_dedup_scope_begin(c); }
        ~DedupScope() { msm_dedup_scope_end(c); }
    } dedup_scope(c);
    B2S_TRY(msm_run(c, 2, pk->b_g2_query.p, zd + pk->b2_off, pk->b2_len + pk->b2_ext, true, g2_out, w2));
    B2S_TRY(msm_run(c, 1, pk->a_query.p, zd + pk->a_off, pk->a_len + pk->a_ext, true, g1 + 2, w1 + 2 * 64));
    B2S_TRY(msm_run(c, 1, pk->b_g1_query.p, zd + pk->b1_off, pk->b1_len + pk->b1_ext, true, g1 + 3, w1 + 3 * 64));
    B2S_TRY(msm_run(c, 1, pk->l_query.p, zd + m->n_instance + pk->l_off, pk->l_len, true, g1 + 1, w1 + 1 * 64));
    if (fork) {
        B2S_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_join, 0));
        side_guard.active = false;
    } else {
        B2S_TRY(hs->get(c, pk, m, zd, &h_shard));
    }
    if (pk->h_table.p) B2S_TRY(msm_run(c, 1, pk->h_table.p, h_shard, pk->h_len, true, g1 + 0, w1 + 0 * 64, &pk->h_pre));
    else B2S_TRY(msm_run(c, 1, pk->h_query.p, h_shard, pk->h_len, true, g1 + 0, w1 + 0 * 64));
    B2S_TRY(msm_join_tails(c));
    return B2S_OK;
}

int32_t groth16_shard(Ctx* c, const b2s_pk* pk, const b2s_r1cs* m, const void* z_inst, const void* z_wit, const void* z_dev,
                      const void* r_host, const void* s_host, void* g1_out, void* g2_out, HSource* hs) {
    return dispatch_curve(c, [&](auto curve) {
        return shard_t<decltype(curve)>(c, pk, m, z_inst, z_wit, z_dev, r_host, s_host, g1_out, g2_out, hs);
    });
}

// g1_partials: shard i's four G1 sums start at element i * g1_stride (XYZZ<Fq> units); g2_partials likewise (XYZZ<Fq2>)
template <class Curve>
static int32_t finish_t(Ctx* c, const b2s_pk* pk, const void* g1_partials, uint32_t g1_stride, const void* g2_partials, uint32_t g2_stride,
                        uint32_t n_shards, const void* r_host, const void* s_host, void* out_a, void* out_b, void* out_c) {
    using Fr = typename Curve::Fr;
    using Fq = typename Curve::Fq;
    using Fq2 = typename Curve::Fq2;
    DevBuf rs, sums1, sums2, outs;
    B2S_TRY(rs.alloc(c, 2 * sizeof(Fr)));
    B2S_CUDA(c, cudaMemcpyAsync(rs.p, r_host, sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
    B2S_CUDA(c, cudaMemcpyAsync(rs.as<Fr>() + 1, s_host, sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
    B2S_TRY(sums1.alloc(c, 4 * sizeof(XYZZ<Fq>)));
    B2S_TRY(sums2.alloc(c, sizeof(XYZZ<Fq2>)));
    B2S_LAUNCH(c, sum_shards_kernel<Fq>, 1, 32, 0, reinterpret_cast<const XYZZ<Fq>*>(g1_partials), n_shards, g1_stride, 4u, sums1.as<XYZZ<Fq>>());
    B2S_LAUNCH(c, sum_shards_kernel<Fq2>, 1, 32, 0, reinterpret_cast<const XYZZ<Fq2>*>(g2_partials), n_shards, g2_stride, 1u, sums2.as<XYZZ<Fq2>>());
    const size_t g1 = sizeof(Affine<Fq>), g2 = sizeof(Affine<Fq2>);
    B2S_TRY(outs.alloc(c, 2 * g1 + g2));
    char* o = outs.as<char>();
    B2S_LAUNCH(c, groth16_epilogue_g1_kernel<Curve>, 1, 128, 0, pk->consts_g1.as<Affine<Fq>>(), sums1.as<XYZZ<Fq>>(), rs.as<Fr>(),
               reinterpret_cast<Affine<Fq>*>(o), reinterpret_cast<Affine<Fq>*>(o + g1));
    B2S_LAUNCH(c, groth16_epilogue_g2_kernel<Curve>, 1, 32, 0, pk->consts_g2.as<Affine<Fq2>>(), sums2.as<XYZZ<Fq2>>(),
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
        return finish_t<decltype(curve)>(c, pk, g1_partials_dev, 4u, g2_partials_dev, 1u, n_shards, r_host, s_host, out_a, out_b, out_c);
    });
}

// The all-gathered layout of group.cu: rank i's packet = [4 G1 XYZZ | 1 G2 XYZZ] at byte i * (4 |P1| + |P2|); |P2| = 2 |P1|,
// so the G1 sums sit at stride 6 (P1 units) and the G2 sum at element 2 + 3 i (P2 units).
int32_t groth16_finish_strided(Ctx* c, const b2s_pk* pk, const void* packed_dev, uint32_t n_shards, const void* r_host, const void* s_host,
                               void* out_a, void* out_b, void* out_c) {
    return dispatch_curve(c, [&](auto curve) {
        using C = decltype(curve);
        static_assert(sizeof(XYZZ<typename C::Fq2>) == 2 * sizeof(XYZZ<typename C::Fq>), "packet layout");
        const char* base = reinterpret_cast<const char*>(packed_dev);
        return finish_t<C>(c, pk, base, 6u, base + 4 * sizeof(XYZZ<typename C::Fq>), 3u, n_shards, r_host, s_host, out_a, out_b, out_c);
    });
}

}  // namespace b2s
