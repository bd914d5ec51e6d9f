// Batched-affine pairwise reduction rounds in front of the XYZZ bucket accumulation.
//
// A round halves every bucket: the points of a bucket (contiguous in the bucket-sorted order) are added
// in adjacent pairs IN AFFINE coordinates; all pairs of a round are independent, so each thread takes K
// consecutive outputs and shares ONE field inversion among them (Montgomery's trick):
//   pass 1  d_i = x2 - x1 (or 2 y1 when doubling, 1 when nothing is to be inverted), prefix products to scratch
//   invert  the product of the K denominators
//   pass 2  backwards: 1/d_i from the running inverse and the stored prefix, lambda = num_i / d_i,
//           x3 = lambda^2 - x1 - x2,  y3 = lambda (x1 - x3) - y1
// i.e. 6 multiplications per addition plus inverse/K, against 10 for the XYZZ mixed addition (Fq2: 17 vs 28
// base multiplications).  The price is HBM traffic (points are read twice, prefixes written and read) -- the
// resource this ALU-bound path leaves idle.  After R rounds every bucket holds ceil(n / 2^R) points and the
// XYZZ kernel (msm_acc.cuh) finishes.  All special cases keep the result an exact group element:
// missing partner / infinity -> copy, P + P -> tangent, P + (-P) -> infinity.
#pragma once
#include "msm_acc.cuh"

namespace b2s {

static constexpr int BA_THREADS = 128;

enum : uint32_t { BA_COPY1 = 0, BA_COPY2 = 1, BA_ADD = 2, BA_DBL = 3, BA_INF = 4 };

template <class F>
struct BaPair { Affine<F> p1, p2; uint32_t kind; };

// inputs of output element `o` of bucket g (round input layout in_off, this round's counts via in_off)
template <class F, bool FIRST>
__device__ __forceinline__ BaPair<F> ba_load(const Affine<F>* __restrict__ bases, const uint32_t* __restrict__ sorted,
                                             const Affine<F>* __restrict__ prev, uint32_t in0, bool has2) {
    BaPair<F> r;
    if (FIRST) {
        const uint32_t e1 = sorted[in0];
        r.p1 = ld_struct(bases + (e1 & 0x7fffffffu));
        if (e1 >> 31) r.p1.y = r.p1.y.neg();
        if (has2) {
            const uint32_t e2 = sorted[in0 + 1];
            r.p2 = ld_struct(bases + (e2 & 0x7fffffffu));
            if (e2 >> 31) r.p2.y = r.p2.y.neg();
        }
    } else {
        r.p1 = ld_struct(prev + in0);
        if (has2) r.p2 = ld_struct(prev + in0 + 1);
    }
    if (!has2 || r.p2.is_inf()) r.kind = BA_COPY1;
    else if (r.p1.is_inf()) r.kind = BA_COPY2;
    else if (r.p1.x == r.p2.x) r.kind = (r.p1.y == r.p2.y && !r.p1.y.is_zero()) ? BA_DBL : BA_INF;
    else r.kind = BA_ADD;
    return r;
}

template <class F>
__device__ __forceinline__ F ba_denominator(const BaPair<F>& q) {
    if (q.kind == BA_ADD) return q.p2.x - q.p1.x;
    if (q.kind == BA_DBL) return q.p1.y.dbl();
    return F::one();
}

// Bucket of output `o`: the largest g with off[g] <= o (empty buckets have off[g] == off[g+1] and are skipped).
// `hint` is a bucket at or before it.  The next bucket is tried first (dense case); otherwise a binary search --
// never a linear walk: with skewed scalars two non-empty buckets can be 2^19 empty ones apart.
__device__ __forceinline__ uint32_t ba_bucket_fwd(const uint32_t* __restrict__ off, uint32_t G, uint32_t hint, uint32_t o) {
    if (off[hint + 1] > o) return hint;
    if (hint + 2 <= G && off[hint + 2] > o) return hint + 1;
    uint32_t lo = hint + 1, hi = G;          // off[lo] <= o < off[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (off[mid] <= o) lo = mid; else hi = mid;
    }
    return lo;
}
// same, searching downwards from a bucket `hint` whose range starts after o
__device__ __forceinline__ uint32_t ba_bucket_bwd(const uint32_t* __restrict__ off, uint32_t hint, uint32_t o) {
    if (hint > 0 && off[hint - 1] <= o) return hint - 1;
    uint32_t lo = 0, hi = hint;              // off[lo] <= o < off[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (off[mid] <= o) lo = mid; else hi = mid;
    }
    return lo;
}

// One thread: outputs [t*K, (t+1)*K) of this round.
template <class F, bool FIRST>
__global__ void __launch_bounds__(BA_THREADS)
msm_ba_round_kernel(const Affine<F>* __restrict__ bases, const uint32_t* __restrict__ sorted, const Affine<F>* __restrict__ prev,
                    const uint32_t* __restrict__ in_off, const uint32_t* __restrict__ out_off, uint32_t G, uint32_t K,
                    F* __restrict__ prefix, Affine<F>* __restrict__ out) {
    const uint32_t total = out_off[G];
    const uint64_t o_beg64 = (uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) * K;
    if (o_beg64 >= total) return;
    const uint32_t o_beg = (uint32_t)o_beg64;
    const uint32_t o_end = (uint32_t)min((uint64_t)total, o_beg64 + K);
    // bucket of the first output: largest g with out_off[g] <= o_beg (skipping empty buckets)
    uint32_t lo = 0, hi = G;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (out_off[mid] <= o_beg) lo = mid; else hi = mid;
    }
    const uint32_t g0 = lo;

    // ---- pass 1: prefix products of the denominators.  Only the x coordinates are needed unless a pair is special
    // (missing partner, x = 0 which may be the (0,0) encoding of infinity, or equal x): then the y's are fetched and
    // the pair is classified exactly as pass 2 will.  The next pair's x's are requested before the current
    // multiplication so that the gather latency hides behind it.
    F prod = F::one();
    {
        uint32_t g = g0, g_out_end = out_off[g + 1], g_in = in_off[g], g_in_end = in_off[g + 1], g_out = out_off[g];
        struct Xs { F x1, x2; uint32_t in0; bool has2; };
        auto fetch = [&](uint32_t o) {
            if (o >= g_out_end) { g = ba_bucket_fwd(out_off, G, g, o); g_out = out_off[g]; g_out_end = out_off[g + 1]; g_in = in_off[g]; g_in_end = in_off[g + 1]; }
            Xs r;
            r.in0 = g_in + 2 * (o - g_out);
            r.has2 = r.in0 + 1 < g_in_end;
            if (FIRST) {
                r.x1 = ld_struct(&bases[sorted[r.in0] & 0x7fffffffu].x);
                r.x2 = r.has2 ? ld_struct(&bases[sorted[r.in0 + 1] & 0x7fffffffu].x) : F::zero();
            } else {
                r.x1 = ld_struct(&prev[r.in0].x);
                r.x2 = r.has2 ? ld_struct(&prev[r.in0 + 1].x) : F::zero();
            }
            return r;
        };
        Xs nxt = fetch(o_beg);
        for (uint32_t o = o_beg; o < o_end; o++) {
            const Xs cur = nxt;
            if (o + 1 < o_end) nxt = fetch(o + 1);
            F d;
            if (cur.has2 && !cur.x1.is_zero() && !cur.x2.is_zero() && cur.x1 != cur.x2) {
                d = cur.x2 - cur.x1;
            } else {
                BaPair<F> q = ba_load<F, FIRST>(bases, sorted, prev, cur.in0, cur.has2);   // rare: full classification
                d = ba_denominator(q);
            }
            st_struct(prefix + o, prod);
            prod = prod * d;
        }
    }
    F inv = prod.inverse();
    // ---- pass 2: backwards
    {
        // bucket of the last output
        uint32_t g = ba_bucket_fwd(out_off, G, g0, o_end - 1);
        uint32_t g_out = out_off[g], g_in = in_off[g], g_in_end = in_off[g + 1];
        for (uint32_t o = o_end; o-- > o_beg;) {
            if (o < g_out) { g = ba_bucket_bwd(out_off, g, o); g_out = out_off[g]; g_in = in_off[g]; g_in_end = in_off[g + 1]; }
            const uint32_t in0 = g_in + 2 * (o - g_out);
            BaPair<F> q = ba_load<F, FIRST>(bases, sorted, prev, in0, in0 + 1 < g_in_end);
            const F d = ba_denominator(q);
            const F dinv = inv * ld_struct(prefix + o);
            inv = inv * d;
            Affine<F> res;
            if (q.kind == BA_ADD || q.kind == BA_DBL) {
                F num;
                if (q.kind == BA_ADD) num = q.p2.y - q.p1.y;
                else { F xx = q.p1.x.sqr(); num = xx.dbl() + xx; }
                const F lam = num * dinv;
                const F x3 = lam.sqr() - q.p1.x - q.p2.x;   // DBL: p2 == p1, so this is lambda^2 - 2 x1
                res.x = x3;
                res.y = lam * (q.p1.x - x3) - q.p1.y;
            } else if (q.kind == BA_COPY1) res = q.p1;
            else if (q.kind == BA_COPY2) res = q.p2;
            else res = Affine<F>::inf();
            st_struct(out + o, res);
        }
    }
}

// counts_next[g] = ceil(counts[g] / 2)
static __global__ void msm_ba_halve_kernel(const uint32_t* __restrict__ counts, uint32_t G, uint32_t* __restrict__ next) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < G) next[g] = (counts[g] + 1) >> 1;
}

// implemented in msm_acc_g1.cu / msm_acc_g2.cu (multiplication inlined)
int32_t msm_ba_round_g1(Ctx* c, bool first, const void* bases, const uint32_t* sorted, const void* prev, const uint32_t* in_off,
                        const uint32_t* out_off, uint32_t G, uint32_t K, uint64_t out_bound, void* prefix, void* out);
int32_t msm_ba_round_g2(Ctx* c, bool first, const void* bases, const uint32_t* sorted, const void* prev, const uint32_t* in_off,
                        const uint32_t* out_off, uint32_t G, uint32_t K, uint64_t out_bound, void* prefix, void* out);

template <class F>
static int32_t msm_ba_round_launch(Ctx* c, const char* label, bool first, const void* bases, const uint32_t* sorted, const void* prev,
                                   const uint32_t* in_off, const uint32_t* out_off, uint32_t G, uint32_t K, uint64_t out_bound,
                                   void* prefix, void* out) {
    const unsigned grid = cdiv(cdiv(out_bound, K), BA_THREADS);
    if (grid == 0) return B2S_OK;
    if (first)
        B2S_LAUNCH_N(c, label, (msm_ba_round_kernel<F, true>), grid, BA_THREADS, 0, reinterpret_cast<const Affine<F>*>(bases), sorted,
                     reinterpret_cast<const Affine<F>*>(prev), in_off, out_off, G, K, reinterpret_cast<F*>(prefix),
                     reinterpret_cast<Affine<F>*>(out));
    else
        B2S_LAUNCH_N(c, label, (msm_ba_round_kernel<F, false>), grid, BA_THREADS, 0, reinterpret_cast<const Affine<F>*>(bases), sorted,
                     reinterpret_cast<const Affine<F>*>(prev), in_off, out_off, G, K, reinterpret_cast<F*>(prefix),
                     reinterpret_cast<Affine<F>*>(out));
    return B2S_OK;
}

}  // namespace b2s
