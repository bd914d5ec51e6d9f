// Distributed witness_map: the transforms of the LibsnarkReduction (six of them, see r1cs.cu) cut over G = 2^lg ranks by the
// four-step (Bailey) schedule, SpMV and the pointwise quotient on the same distribution.
//
// Distribution.  N = 2^n (n even), N1 = N2 = 2^(n/2).  A vector x[i], i = i1 N2 + i2, is held by the rank that owns the
// COLUMN i2:  rank g owns i2 in [g W, (g+1) W), W = N2 / G, as the local array L[i1][c] = x[i1 N2 + g W + c].
//   X[k1 + N1 k2] = sum_{i2} w_N^(i2 k1) w_N2^(i2 k2)  sum_{i1} x[i1 N2 + i2] w_N1^(i1 k1)
//   1  local:    N1-point transforms down the columns                       (passes A, B)
//   2  local:    twiddle w_N^(i2 k1), fused into the stores of pass B
//   3  EXCHANGE: rows k1 go to the rank that owns them (all-to-all, N/G^2 elements per pair; pass B writes its results
//                straight into per-destination blocks, so the exchange moves contiguous buffers)
//   4  local:    N2-point transforms along the rows                         (passes C, D)
// and the rank that owns rows k1 in [g W, (g+1) W) ends with X[k1 + N1 k2] for all k2 -- which, because N1 = N2, is again
// "the columns i2 = k1 of the next transform's input": pass D stores transposed (L'[k2][k1 - g W]) and the chain
// iNTT -> coset NTT -> pointwise -> coset iNTT keeps ONE distribution with no re-gathering.  Only h is moved once more,
// into contiguous coefficient slabs [g N/G, (g+1) N/G) for the base-range shard of the h-query MSM.
// Per rank: 1/G of the butterflies, 4 exchanges per proof (a,b,c travel together), 3 N/G^2 elements per pair each.
//
// Every pass is the same kernel: a radix-2^r decimation-in-frequency transform over one index of a tile in shared memory,
// with the source / destination addresses, twiddle exponents and coset-scaling indices given as BIT-FIELD MAPS of the
// logical coordinates (p, r, m) -- all dimensions are powers of two, so every layout above is a permutation of index bits.
//
// The same schedule runs with G virtual ranks on ONE GPU (b2s_witness_map_sim: exchange = device-to-device copies), which
// is how the index algebra is tested bit-exactly against the single-GPU witness_map without a multi-GPU box.
#define B2S_INLINE_MUL 1
#include "dntt.cuh"

namespace b2s {

// sum over segments: bits[i] bits of x (consumed from the least significant end) placed at shift[i]
struct BitMap {
    uint32_t n = 0;
    uint8_t bits[4] = {0, 0, 0, 0}, shift[4] = {0, 0, 0, 0};
    __host__ __device__ uint64_t map(uint64_t x) const {
        uint64_t r = 0;
#pragma unroll
        for (uint32_t i = 0; i < 4; i++)
            if (i < n) {
                r |= (x & ((1ull << bits[i]) - 1ull)) << shift[i];
                x >>= bits[i];
            }
        return r;
    }
    BitMap& seg(uint32_t b, uint32_t s) {
        if (b) { bits[n] = (uint8_t)b; shift[n] = (uint8_t)s; n++; }
        return *this;
    }
};

struct DPass {
    uint32_t log_r = 0, log_c = 0, log_m = 0, log_p = 0;
    uint32_t load_r_fast = 0, store_r_fast = 0;   // which coordinate runs over consecutive threads (coalescing)
    BitMap in_p, in_r, in_m, out_p, out_r, out_m;
    PowTab tw_r; uint32_t tw_r_shift = 0;          // butterflies: w_R^e = tw_r^(e << tw_r_shift)
    PowTab tw_in; uint32_t in_mshift = 0;          // output k *= tw_in^((m >> in_mshift) k)
    PowTab tw_gl; BitMap gl_i2_m, gl_k1_p, gl_k1_k; uint32_t gl_i2_base = 0, log_n = 0;   // *= tw_gl^(i2 k1 mod N)
    PowTab pre; BitMap pre_p, pre_r, pre_m; uint64_t pre_base = 0;      // input  *= pre^(global input index)
    PowTab post; BitMap post_p, post_r, post_m; uint64_t post_base = 0; // output *= post^(global output index)
    const void* post_const = nullptr;
};

template <class Fr>
__global__ void __launch_bounds__(NTT_THREADS) dntt_pass_kernel(const Fr* __restrict__ src, Fr* __restrict__ dst, DPass a) {
    extern __shared__ uint4 smem[];
    const uint32_t R = 1u << a.log_r, C = 1u << a.log_c;
    const uint32_t pitch = 2 * C + 1;
    uint4* tile = smem;
    Fr* wtab = reinterpret_cast<Fr*>(smem + (size_t)R * pitch + 1);
    const uint32_t tiles_per_p = 1u << (a.log_m - a.log_c);
    const uint64_t p = blockIdx.x / tiles_per_p;
    const uint64_t m0 = (uint64_t)(blockIdx.x % tiles_per_p) << a.log_c;
    const uint64_t src_p = a.in_p.map(p), dst_p = a.out_p.map(p);

    for (uint32_t e = threadIdx.x; e < R / 2; e += blockDim.x) gst<Fr>(wtab + e, pow_lookup<Fr>(a.tw_r, (uint64_t)e << a.tw_r_shift));
    for (uint32_t idx = threadIdx.x; idx < R * C; idx += blockDim.x) {
        const uint32_t j = a.load_r_fast ? (idx & (R - 1)) : (idx >> a.log_c);
        const uint32_t c = a.load_r_fast ? (idx >> a.log_r) : (idx & (C - 1));
        const uint64_t m = m0 + c;
        Fr v = gld<Fr>(src + (src_p + a.in_r.map(j) + a.in_m.map(m)));
        if (a.pre.lo) v = v * pow_lookup<Fr>(a.pre, a.pre_base + a.pre_p.map(p) + a.pre_r.map(j) + a.pre_m.map(m));
        tile_st<Fr>(tile, j, c, pitch, v);
    }
    __syncthreads();
    tile_dif<Fr>(tile, wtab, a.log_r, a.log_c, pitch);
    const uint64_t nmask = (1ull << a.log_n) - 1ull;
    for (uint32_t idx = threadIdx.x; idx < R * C; idx += blockDim.x) {
        const uint32_t rho = a.store_r_fast ? (idx & (R - 1)) : (idx >> a.log_c);
        const uint32_t c = a.store_r_fast ? (idx >> a.log_r) : (idx & (C - 1));
        const uint64_t k = bitrev(rho, a.log_r);
        const uint64_t m = m0 + c;
        Fr v = tile_ld<Fr>(tile, rho, c, pitch);
        if (a.tw_in.lo) {
            const uint64_t e = (m >> a.in_mshift) * k;
            if (e) v = v * pow_lookup<Fr>(a.tw_in, e);
        }
        if (a.tw_gl.lo) {
            const uint64_t e = ((a.gl_i2_base + a.gl_i2_m.map(m)) * (a.gl_k1_p.map(p) + a.gl_k1_k.map(k))) & nmask;
            if (e) v = v * pow_lookup<Fr>(a.tw_gl, e);
        }
        if (a.post.lo) v = v * pow_lookup<Fr>(a.post, a.post_base + a.post_p.map(p) + a.post_r.map(k) + a.post_m.map(m));
        else if (a.post_const) v = v * gld<Fr>(reinterpret_cast<const Fr*>(a.post_const));
        gst<Fr>(dst + (dst_p + a.out_r.map(k) + a.out_m.map(m)), v);
    }
}

// SpMV on the distributed layout: local element t = i1 W + c is row i = i1 N2 + g W + c of the padded evaluation vectors
// (rows >= n_rows: the input-consistency rows a[n_rows + j] = z[j], zero elsewhere) -- r1cs.cu's kernel, other row order.
struct DSpmvMat { const uint64_t* row_ptr; const uint32_t* col; const uint32_t* cid; void* out; };
template <class Fr>
__global__ void __launch_bounds__(256)
dspmv_kernel(DSpmvMat m0, DSpmvMat m1, DSpmvMat m2, const Fr* __restrict__ pool, const Fr* __restrict__ z, uint64_t n_rows, uint64_t n_inst,
             uint32_t log_local, uint32_t lw, uint32_t log_n2, uint64_t col_base) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if ((t >> log_local) >= 3) return;
    const uint32_t k = (uint32_t)(t >> log_local);
    const uint64_t loc = t & ((1ull << log_local) - 1ull);
    const uint64_t row = ((loc >> lw) << log_n2) + col_base + (loc & ((1ull << lw) - 1ull));
    const DSpmvMat m = k == 0 ? m0 : (k == 1 ? m1 : m2);
    Fr acc = Fr::zero();
    if (row < n_rows) {
        const uint64_t beg = m.row_ptr[row], end = m.row_ptr[row + 1];
        for (uint64_t e = beg; e < end; e++) {
            const uint32_t cid = m.cid[e];
            Fr v = gld<Fr>(z + m.col[e]);
            if (cid != 0) v = v * gld<Fr>(pool + cid);
            acc = acc + v;
        }
    } else if (k == 0 && row < n_rows + n_inst) {
        acc = gld<Fr>(z + (row - n_rows));
    }
    gst<Fr>(reinterpret_cast<Fr*>(m.out) + loc, acc);
}

// a *= b on the coset; later, on coefficients, h = (q - c) * zinv: the c term of (a b - c) / Z never needs its own coset
// transform, because Z is constant on the coset and deg C < N (r1cs.cu, witness_map_t)
template <class Fr>
__global__ void __launch_bounds__(256) dqap_mul_kernel(Fr* a, const Fr* b, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    gst<Fr>(a + i, gld<Fr>(a + i) * gld<Fr>(b + i));
}
template <class Fr>
__global__ void __launch_bounds__(256) dqap_quotient_kernel(Fr* q, const Fr* c, const Fr* zinv, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    gst<Fr>(q + i, (gld<Fr>(q + i) - gld<Fr>(c + i)) * gld<Fr>(zinv));
}
template <class Fr, class FrP>
__global__ void dvanishing_inv_kernel(Fr* out, uint64_t domain) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Fr g;
    for (int i = 0; i < Fr::N; i++) g.v[i] = FrP::gen(i);
    out[0] = (g.pow_u64(domain) - Fr::one()).inverse();
}
// received slab blocks [src][i1 local][c] -> slab[i1 local * N2 + src W + c]
template <class Fr>
__global__ void __launch_bounds__(256) dunpack_kernel(const Fr* __restrict__ recv, Fr* __restrict__ slab, uint32_t log_blk, uint32_t lw, uint32_t log_n2, uint64_t n) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint64_t s = t >> log_blk, loc = t & ((1ull << log_blk) - 1ull);
    gst<Fr>(slab + (((loc >> lw) << log_n2) + (s << lw) + (loc & ((1ull << lw) - 1ull))), gld<Fr>(recv + t));
}

// ---------------------------------------------------------------------------------------------------------------------
bool dist_supported(uint32_t log_n, uint32_t lg) {
    if (log_n & 1u) return false;
    const uint32_t h = log_n / 2, ha = (h + 1) / 2, hb = h - ha;
    return lg >= 1 && lg <= hb && ha <= (uint32_t)NTT_MAX_RADIX_LOG;
}

struct DGeom {
    uint32_t n, h, ha, hb, lg, lw;
    DGeom(uint32_t log_n, uint32_t lg_) : n(log_n), h(log_n / 2), ha((log_n / 2 + 1) / 2), hb(log_n / 2 - (log_n / 2 + 1) / 2), lg(lg_), lw(log_n / 2 - lg_) {}
};

template <class Fr>
static int32_t launch_pass(Ctx* c, const char* label, const Fr* src, Fr* dst, DPass& a) {
    a.log_c = std::min<uint32_t>((uint32_t)NTT_TILE_LOG - a.log_r, a.log_m);
    const unsigned grid = 1u << (a.log_p + a.log_m - a.log_c);
    B2S_SMEM_ATTR(c, dntt_pass_kernel<Fr>, ntt_pass_smem_bytes());
    B2S_LAUNCH_N(c, label, dntt_pass_kernel<Fr>, grid, NTT_THREADS, ntt_pass_smem_bytes(), src, dst, a);
    return B2S_OK;
}

// First half of one distributed transform on virtual / real rank g: passes A and B (+ twiddle), results in `send` as
// per-destination blocks of N / G^2 elements.
template <class Curve>
static int32_t dntt_half1(Ctx* c, const DGeom& q, uint32_t g, typename Curve::Fr* data, typename Curve::Fr* send, bool inverse, bool coset) {
    using Fr = typename Curve::Fr;
    NttPlan *pn = nullptr, *ph = nullptr;
    B2S_TRY(ntt_get_plan(c, q.n, &pn));
    B2S_TRY(ntt_get_plan(c, q.h, &ph));
    const PowTab twh = inverse ? ph->inv : ph->fwd, twn = inverse ? pn->inv : pn->fwd;
    {   // A: over a (i1 = a Rb + b), m = (b, c), in place
        DPass a;
        a.log_r = q.ha; a.log_m = q.hb + q.lw; a.log_p = 0;
        a.in_r.seg(q.ha, q.hb + q.lw); a.in_m.seg(q.hb + q.lw, 0);
        a.out_r = a.in_r; a.out_m = a.in_m;
        a.tw_r = twh; a.tw_r_shift = q.h - q.ha;
        a.tw_in = twh; a.in_mshift = q.lw;
        if (coset && !inverse) {   // x_i *= g^i, i = (a Rb + b) N2 + g W + c
            a.pre = pn->coset_in;
            a.pre_r.seg(q.ha, q.hb + q.h);
            a.pre_m.seg(q.lw, 0).seg(q.hb, q.h);
            a.pre_base = (uint64_t)g << q.lw;
        }
        B2S_TRY(launch_pass<Fr>(c, "dntt_pass_a", data, data, a));
    }
    {   // B: over b, p = ka, m = c; twiddle w_N^(i2 k1); stores into [dest][ka][kb low][c]
        DPass a;
        a.log_r = q.hb; a.log_m = q.lw; a.log_p = q.ha;
        a.in_p.seg(q.ha, q.hb + q.lw); a.in_r.seg(q.hb, q.lw); a.in_m.seg(q.lw, 0);
        a.out_p.seg(q.ha, q.hb - q.lg + q.lw);
        a.out_r.seg(q.hb - q.lg, q.lw).seg(q.lg, q.n - 2 * q.lg);
        a.out_m.seg(q.lw, 0);
        a.tw_r = twh; a.tw_r_shift = q.h - q.hb;
        a.tw_gl = twn; a.log_n = q.n;
        a.gl_i2_base = g << q.lw; a.gl_i2_m.seg(q.lw, 0);
        a.gl_k1_p.seg(q.ha, 0); a.gl_k1_k.seg(q.hb, q.ha);
        B2S_TRY(launch_pass<Fr>(c, "dntt_pass_b", data, send, a));
    }
    return B2S_OK;
}

// Second half on rank d: passes C and D over the received rows [src][ka][kb low][c]; results in `data`, natural local layout.
template <class Curve>
static int32_t dntt_half2(Ctx* c, const DGeom& q, uint32_t d, typename Curve::Fr* recv, typename Curve::Fr* data, bool inverse, bool coset) {
    using Fr = typename Curve::Fr;
    NttPlan *pn = nullptr, *ph = nullptr;
    B2S_TRY(ntt_get_plan(c, q.n, &pn));
    B2S_TRY(ntt_get_plan(c, q.h, &ph));
    const PowTab twh = inverse ? ph->inv : ph->fwd;
    {   // C: over a' (top ha bits of i2 = [src | c]), p = row x, m = b' (low hb bits of c), in place
        DPass a;
        a.log_r = q.ha; a.log_m = q.hb; a.log_p = q.lw;
        a.in_p.seg(q.lw, q.lw);
        a.in_r.seg(q.ha - q.lg, q.hb).seg(q.lg, q.n - 2 * q.lg);
        a.in_m.seg(q.hb, 0);
        a.out_p = a.in_p; a.out_r = a.in_r; a.out_m = a.in_m;
        a.tw_r = twh; a.tw_r_shift = q.h - q.ha;
        a.tw_in = twh; a.in_mshift = 0;
        B2S_TRY(launch_pass<Fr>(c, "dntt_pass_c", recv, recv, a));
    }
    {   // D: over b' (contiguous), p = row x, m = ka'; X[k1 + N1 k2] -> L'[k2][k1 local], k2 = ka' + Ra kb', k1 local = ka + Ra kbl
        DPass a;
        a.log_r = q.hb; a.log_m = q.ha; a.log_p = q.lw;
        a.load_r_fast = 1;
        a.in_p.seg(q.lw, q.lw);
        a.in_m.seg(q.ha - q.lg, q.hb).seg(q.lg, q.n - 2 * q.lg);
        a.in_r.seg(q.hb, 0);
        a.out_p.seg(q.hb - q.lg, q.ha).seg(q.ha, 0);        // x = ka 2^(hb-lg) + kbl  ->  k1 local = ka + Ra kbl
        a.out_m.seg(q.ha, q.lw);
        a.out_r.seg(q.hb, q.lw + q.ha);
        a.tw_r = twh; a.tw_r_shift = q.h - q.hb;
        if (inverse && coset) {   // X_k *= g^-k / N, k = d W + k1 local + N1 k2
            a.post = pn->coset_out_scaled;
            a.post_p.seg(q.hb - q.lg, q.ha).seg(q.ha, 0);
            a.post_m.seg(q.ha, q.h);
            a.post_r.seg(q.hb, q.h + q.ha);
            a.post_base = (uint64_t)d << q.lw;
        } else if (inverse) {
            a.post_const = pn->n_inv;
        }
        B2S_TRY(launch_pass<Fr>(c, "dntt_pass_d", recv, data, a));
    }
    return B2S_OK;
}

template <class Curve>
static int32_t dspmv(Ctx* c, const DGeom& q, uint32_t g, const b2s_r1cs* m, const typename Curve::Fr* z, typename Curve::Fr* a, typename Curve::Fr* b,
                     typename Curve::Fr* cc) {
    using Fr = typename Curve::Fr;
    const uint32_t log_local = q.n - q.lg;
    DSpmvMat mm[3];
    void* outs[3] = {a, b, cc};
    for (int k = 0; k < 3; k++) mm[k] = DSpmvMat{m->row_ptr[k].as<uint64_t>(), m->col[k].as<uint32_t>(), m->coeff_id[k].as<uint32_t>(), outs[k]};
    B2S_LAUNCH_N(c, "dspmv_kernel", dspmv_kernel<Fr>, cdiv(3ull << log_local, 256), 256, 0, mm[0], mm[1], mm[2], m->pool.as<Fr>(), z, m->n_rows,
                 m->n_instance, log_local, q.lw, q.h, (uint64_t)g << q.lw);
    return B2S_OK;
}

// ---- driver: `ranks` = the ranks this process plays (one real rank, or all G virtual ranks of the simulation) -----------
template <class Curve>
static int32_t witness_map_dist_t(Ctx* c, const b2s_r1cs* m, const void* z_dev, uint32_t lg, const std::vector<uint32_t>& ranks,
                                  DistExchange* xch, void* const* h_slab_out) {
    using Fr = typename Curve::Fr;
    using FrP = typename Curve::FrP;
    const uint32_t n = m->log_domain;
    if (!dist_supported(n, lg)) return fail(c, B2S_ERR_INVALID_ARG, "distributed witness_map: domain 2^%u over 2^%u ranks unsupported", n, lg);
    const DGeom q(n, lg);
    const uint64_t local = 1ull << (n - lg), blk = 1ull << (n - 2 * lg);
    const Fr* z = reinterpret_cast<const Fr*>(z_dev);
    const size_t R = ranks.size();
    // per played rank: a, b, c (data), 3 send, 3 recv buffers of N/G elements
    std::vector<DevBuf> bufs(R);
    DevBuf zi;
    B2S_TRY(zi.alloc(c, sizeof(Fr)));
    B2S_LAUNCH_N(c, "dvanishing_inv_kernel", (dvanishing_inv_kernel<Fr, FrP>), 1, 32, 0, zi.as<Fr>(), 1ull << n);
    auto data = [&](size_t r, int v) { return bufs[r].as<Fr>() + (size_t)v * local; };
    auto send = [&](size_t r, int v) { return bufs[r].as<Fr>() + (size_t)(3 + v) * local; };
    auto recv = [&](size_t r, int v) { return bufs[r].as<Fr>() + (size_t)(6 + v) * local; };
    for (size_t r = 0; r < R; r++) {
        B2S_TRY(bufs[r].alloc(c, 9 * local * sizeof(Fr)));
        B2S_TRY((dspmv<Curve>(c, q, ranks[r], m, z, data(r, 0), data(r, 1), data(r, 2))));
    }
    // one distributed transform of `nv` vectors
    auto transform = [&](int nv, bool inverse, bool coset) -> int32_t {
        for (size_t r = 0; r < R; r++)
            for (int v = 0; v < nv; v++) B2S_TRY((dntt_half1<Curve>(c, q, ranks[r], data(r, v), send(r, v), inverse, coset)));
        std::vector<const void*> s(R);
        std::vector<void*> d(R);
        for (int v = 0; v < nv; v++) {
            for (size_t r = 0; r < R; r++) { s[r] = send(r, v); d[r] = recv(r, v); }
            B2S_TRY(xch->all_to_all(c, s.data(), d.data(), blk * sizeof(Fr), v == nv - 1));
        }
        for (size_t r = 0; r < R; r++)
            for (int v = 0; v < nv; v++) B2S_TRY((dntt_half2<Curve>(c, q, ranks[r], recv(r, v), data(r, v), inverse, coset)));
        return B2S_OK;
    };
    B2S_TRY(transform(3, true, false));      // a, b, c -> coefficients (every transform maps the local layout onto itself)
    B2S_TRY(transform(2, false, true));      // a, b -> coset evaluations; c stays in coefficient form
    for (size_t r = 0; r < R; r++)
        B2S_LAUNCH_N(c, "dqap_mul_kernel", dqap_mul_kernel<Fr>, cdiv(local, 256), 256, 0, data(r, 0), (const Fr*)data(r, 1), local);
    B2S_TRY(transform(1, true, true));
    for (size_t r = 0; r < R; r++)
        B2S_LAUNCH_N(c, "dqap_quotient_kernel", dqap_quotient_kernel<Fr>, cdiv(local, 256), 256, 0, data(r, 0), (const Fr*)data(r, 2), (const Fr*)zi.as<Fr>(), local);
    // h into contiguous coefficient slabs: rows i1 in [d N1/G, (d+1) N1/G) of the local layout are one contiguous block
    {
        std::vector<const void*> s(R);
        std::vector<void*> d(R);
        for (size_t r = 0; r < R; r++) { s[r] = data(r, 0); d[r] = recv(r, 0); }
        B2S_TRY(xch->all_to_all(c, s.data(), d.data(), blk * sizeof(Fr), true));
        for (size_t r = 0; r < R; r++)
            B2S_LAUNCH_N(c, "dunpack_kernel", dunpack_kernel<Fr>, cdiv(local, 256), 256, 0, recv(r, 0), reinterpret_cast<Fr*>(h_slab_out[r]), n - 2 * lg, q.lw,
                         q.h, local);
    }
    return B2S_OK;
}

int32_t witness_map_dist(Ctx* c, const b2s_r1cs* m, const void* z_dev, uint32_t lg, const std::vector<uint32_t>& ranks, DistExchange* xch,
                         void* const* h_slab_out) {
    return dispatch_curve(c, [&](auto curve) { return witness_map_dist_t<decltype(curve)>(c, m, z_dev, lg, ranks, xch, h_slab_out); });
}

// ---- simulation: all G ranks on this GPU, the exchange is a set of device-to-device copies --------------------------------
struct SimExchange : DistExchange {
    uint32_t G;
    explicit SimExchange(uint32_t g) : G(g) {}
    int32_t all_to_all(Ctx* c, const void* const* send, void* const* recv, size_t block_bytes, bool) override {
        for (uint32_t s = 0; s < G; s++)
            for (uint32_t d = 0; d < G; d++)
                B2S_CUDA(c, cudaMemcpyAsync(reinterpret_cast<char*>(recv[d]) + (size_t)s * block_bytes,
                                            reinterpret_cast<const char*>(send[s]) + (size_t)d * block_bytes, block_bytes, cudaMemcpyDeviceToDevice, c->stream));
        return B2S_OK;
    }
};

int32_t witness_map_sim(Ctx* c, const b2s_r1cs* m, const void* z_dev, uint32_t lg, void* h_dev) {
    const uint32_t G = 1u << lg;
    std::vector<uint32_t> ranks(G);
    std::vector<void*> slabs(G);
    const size_t slab_bytes = (size_t)32 << (m->log_domain - lg);
    for (uint32_t g = 0; g < G; g++) { ranks[g] = g; slabs[g] = reinterpret_cast<char*>(h_dev) + g * slab_bytes; }
    SimExchange x(G);
    return witness_map_dist(c, m, z_dev, lg, ranks, &x, slabs.data());
}

}  // namespace b2s
