// K4: variable-base multi-scalar multiplication (Pippenger, signed radix-2^c windows) for G1 and G2.
//
// GPU counterpart of ark-ec `VariableBaseMSM::msm` / `msm_bigint` (upstream crate, not in
// /root/reference; callers: ark-groth16 `create_proof_with_assignment`,
// Appendix A.1).  The output is a group element, so after normalisation it is bit-identical to any
// other correct MSM regardless of window size or summation order.
//
// Pipeline (no host synchronisation inside; everything on the ctx stream except the Horner tail, which can run on
// the ctx's aux stream so that it overlaps the next MSM of a proof):
//   1 count     one thread per scalar: Montgomery -> canonical, signed digits, histogram of
//               (window, |digit|) bucket sizes (warp-aggregated global atomics, 4 B each)
//   2 scan      exclusive prefix sums (tile sums, spine, apply): bucket offsets in the sorted index array, and task offsets
//               (a bucket of s points is cut into ceil(s / L) tasks so that no thread ever owns more
//               than L points -- this is what keeps degenerate scalar distributions, e.g. the all-equal
//               witness of the reference's DummyCircuit (relations/src/sr1cs/mod.rs:306-309), balanced)
//   3 scatter   counting-sort the point indices (sign in bit 31) by bucket
//   3a rank     buckets are ranked by decreasing size (second counting sort) and task numbers follow the ranks, so
//               the 32 tasks a warp runs in lockstep have equal length (uniform scalars give Poisson bucket sizes)
//   3b for big problems (windows * n >= 2^27; B2S_MSM_AFFINE_ROUNDS overrides) three batched-affine halving rounds,
//               msm_affine.cuh, instead of 3a
//   4 accumulate one thread per task: XYZZ accumulator += affine base, 8M+2S per point; bases are
//               gathered from HBM (96 B / 192 B per point), everything else stays in registers
//   5 reduce    buckets that were split: CTAs sum the task partials of a bucket (two stages, shared-memory tree)
//   6 bucket sum per window sum_b b*B_b by segments: running sums over 16-32 buckets per thread, then
//               seg_start * (segment total) by double-and-add; one CTA per window adds the segments
//   7 horner    sum_w 2^(c w) S_w, one thread (255 doublings; multiplication inlined for ILP, msm_acc_g*.cu)
//
// Roofline: per (point, scalar) the algorithmic HBM traffic is 96+32 B (G1 BLS12-381), but each point
// costs ceil(255/c) mixed additions of ~10 Fq multiplications = ~3000 wide IMADs; the kernel is
// bound by the fma pipe by two orders of magnitude over HBM (DESIGN.md has the numbers).
#include <algorithm>

#include "msm_affine.cuh"

namespace b2s {

// ---- signed-digit recoding -------------------------------------------------------------------------
// Digits of a canonical 256-bit scalar, least significant window first.  Windows other than the
// last are recoded into [-2^(c-1), 2^(c-1)); the last keeps the carry (shape guarantees it fits B).
struct DigitIter {
    uint32_t k[8];
    uint32_t carry;
    __device__ __forceinline__ int32_t next(uint32_t w, uint32_t c, uint32_t nwin) {
        const uint32_t bit = w * c;
        const uint32_t word = bit >> 5, off = bit & 31;
        uint64_t v = 0;
        if (word < 8) v = k[word];
        if (word + 1 < 8) v |= (uint64_t)k[word + 1] << 32;
        uint32_t d = (uint32_t)((v >> off) & ((1u << c) - 1u)) + carry;
        carry = 0;
        if (w != nwin - 1 && d >= (1u << (c - 1))) {
            carry = 1;
            return (int32_t)d - (int32_t)(1u << c);
        }
        return (int32_t)d;
    }
};

template <class Fr>
__device__ __forceinline__ void load_scalar(DigitIter& it, const Fr* scalars, uint64_t i, bool mont) {
    Fr s = ld_struct(scalars + i);
    // canonical input may be any 256-bit value (the ABI does not promise < r): to_mont / from_mont leaves s mod r, so a
    // top-window digit can never index past the bucket arrays
    if (!mont) s = s.to_mont();
    s = s.from_mont();
#pragma unroll
    for (int j = 0; j < 8; j++) it.k[j] = s.v[j];
    it.carry = 0;
}

// Histogram of (window, |digit|).  Lanes of a warp that hit the same bucket are combined with
// match.any so that skewed scalar distributions (all-equal witnesses, many 0/1 values) issue one
// atomic per distinct bucket per warp instead of 32 to the same address.
template <class Fr>
__global__ void msm_count_kernel(const Fr* __restrict__ scalars, uint64_t n, bool mont, MsmShape sh,
                                 uint32_t* __restrict__ counts) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned active = __activemask();
    const unsigned lane = threadIdx.x & 31;
    DigitIter it;
    load_scalar(it, scalars, i, mont);
    for (uint32_t w = 0; w < sh.nwin; w++) {
        const int32_t d = it.next(w, sh.c, sh.nwin);
        const uint32_t key = d != 0 ? w * sh.B + (uint32_t)(d < 0 ? -d : d) - 1 : 0xffffffffu;
        const unsigned peers = __match_any_sync(active, key);
        if (key != 0xffffffffu && lane == (unsigned)(__ffs(peers) - 1)) atomicAdd(&counts[key], (uint32_t)__popc(peers));
    }
}

// Buckets ordered by decreasing size (counting sort on min(count, SIZE_BINS - 1)): task ranks follow this order,
// so the 32 tasks a warp executes in lockstep have (nearly) the same length.  With uniformly random scalars the
// bucket sizes are Poisson; in natural order a warp waits for its longest bucket (+37 % at a mean of 32 points).
static constexpr uint32_t SIZE_BINS = 4096;

__global__ void msm_size_hist_kernel(const uint32_t* __restrict__ counts, uint32_t G, uint32_t* __restrict__ hist) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    const unsigned active = __activemask();
    const uint32_t key = SIZE_BINS - 1 - min(counts[g], SIZE_BINS - 1);
    const unsigned peers = __match_any_sync(active, key);
    if ((threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&hist[key], (uint32_t)__popc(peers));
}
// hist -> exclusive offsets in place in `binoff`; clears hist for reuse as the scatter cursor
__global__ void msm_size_scan_kernel(uint32_t* __restrict__ hist, uint32_t* __restrict__ binoff) {
    __shared__ uint32_t part[1024];
    const uint32_t per = SIZE_BINS / 1024;
    uint32_t local[per], sum = 0;
    for (uint32_t k = 0; k < per; k++) { local[k] = hist[threadIdx.x * per + k]; sum += local[k]; }
    part[threadIdx.x] = sum;
    __syncthreads();
    for (uint32_t d = 1; d < 1024; d <<= 1) {
        uint32_t a = threadIdx.x >= d ? part[threadIdx.x - d] : 0;
        __syncthreads();
        part[threadIdx.x] += a;
        __syncthreads();
    }
    uint32_t run = part[threadIdx.x] - sum;
    for (uint32_t k = 0; k < per; k++) { binoff[threadIdx.x * per + k] = run; run += local[k]; hist[threadIdx.x * per + k] = 0; }
}
__global__ void msm_size_scatter_kernel(const uint32_t* __restrict__ counts, uint32_t G, const uint32_t* __restrict__ binoff,
                                        uint32_t* __restrict__ cursor, uint32_t* __restrict__ perm) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    const unsigned active = __activemask();
    const unsigned lane = threadIdx.x & 31;
    const uint32_t key = SIZE_BINS - 1 - min(counts[g], SIZE_BINS - 1);
    const unsigned peers = __match_any_sync(active, key);
    const unsigned leader = (unsigned)(__ffs(peers) - 1);
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(&cursor[key], (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    perm[binoff[key] + base + __popc(peers & ((1u << lane) - 1u))] = g;
}

// Exclusive scans over the G bucket counts, in three kernels (tile sums -> scan of tile sums -> apply):
//   offsets[g]  = number of points in buckets < g        (position in the sorted index array)
//   task_off[g] = number of tasks in buckets < g         (a bucket of s points has ceil(s / L) tasks)
//   heavy[1..]  = buckets that were split (more than one task), heavy[0] = how many
static constexpr int SCAN_THREADS = 256;
static constexpr int SCAN_PER_THREAD = 8;
static constexpr int SCAN_TILE = SCAN_THREADS * SCAN_PER_THREAD;

struct Scan3 { uint32_t pts, tsk, hvy; };
__device__ __forceinline__ Scan3 operator+(const Scan3& a, const Scan3& b) { return {a.pts + b.pts, a.tsk + b.tsk, a.hvy + b.hvy}; }

// inclusive block scan of one Scan3 per thread; returns the inclusive prefix, total in *total
__device__ __forceinline__ Scan3 block_scan(Scan3 v, Scan3* total) {
    __shared__ Scan3 warp_tot[32];
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Scan3 o = {__shfl_up_sync(0xffffffffu, v.pts, d), __shfl_up_sync(0xffffffffu, v.tsk, d), __shfl_up_sync(0xffffffffu, v.hvy, d)};
        if (lane >= (unsigned)d) v = v + o;
    }
    if (lane == 31) warp_tot[wid] = v;
    __syncthreads();
    const unsigned nw = (blockDim.x + 31) >> 5;
    if (wid == 0) {
        Scan3 w = lane < nw ? warp_tot[lane] : Scan3{0, 0, 0};
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            Scan3 o = {__shfl_up_sync(0xffffffffu, w.pts, d), __shfl_up_sync(0xffffffffu, w.tsk, d), __shfl_up_sync(0xffffffffu, w.hvy, d)};
            if (lane >= (unsigned)d) w = w + o;
        }
        warp_tot[lane] = w;
    }
    __syncthreads();
    if (wid > 0) v = v + warp_tot[wid - 1];
    *total = warp_tot[nw - 1];
    __syncthreads();
    return v;
}

__global__ void __launch_bounds__(SCAN_THREADS)
msm_scan_tiles_kernel(const uint32_t* __restrict__ counts, const uint32_t* __restrict__ perm, MsmShape sh,
                      Scan3* __restrict__ tile_sums) {
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_PER_THREAD;
    Scan3 v{0, 0, 0};
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) {
        const uint32_t g = base + k;
        if (g < sh.G) {
            const uint32_t s = counts[perm ? perm[g] : g], t = (s + sh.L - 1) / sh.L;
            v = v + Scan3{s, t, t > 1 ? 1u : 0u};
        }
    }
    Scan3 total;
    block_scan(v, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// one CTA: exclusive scan of the tile sums in place; totals to offsets[G], task_off[G], heavy[0]
__global__ void __launch_bounds__(1024)
msm_scan_spine_kernel(Scan3* __restrict__ tile_sums, uint32_t ntiles, MsmShape sh, uint32_t* __restrict__ offsets,
                      uint32_t* __restrict__ task_off, uint32_t* __restrict__ heavy) {
    const uint32_t per = (ntiles + blockDim.x - 1) / blockDim.x;
    const uint32_t lo = min(threadIdx.x * per, ntiles), hi = min(lo + per, ntiles);
    Scan3 v{0, 0, 0};
    for (uint32_t i = lo; i < hi; i++) v = v + tile_sums[i];
    Scan3 total;
    Scan3 incl = block_scan(v, &total);
    Scan3 run = {incl.pts - v.pts, incl.tsk - v.tsk, incl.hvy - v.hvy};
    for (uint32_t i = lo; i < hi; i++) {
        Scan3 t = tile_sums[i];
        tile_sums[i] = run;
        run = run + t;
    }
    if (threadIdx.x == 0) {
        if (offsets) offsets[sh.G] = total.pts;
        task_off[sh.G] = total.tsk;
        heavy[0] = total.hvy;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS)
msm_scan_apply_kernel(const uint32_t* __restrict__ counts, const uint32_t* __restrict__ perm, MsmShape sh,
                      const Scan3* __restrict__ tile_sums,
                      uint32_t* __restrict__ offsets, uint32_t* __restrict__ task_off, uint32_t* __restrict__ heavy) {
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_PER_THREAD;
    uint32_t s[SCAN_PER_THREAD], t[SCAN_PER_THREAD];
    Scan3 v{0, 0, 0};
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) {
        const uint32_t g = base + k;
        s[k] = g < sh.G ? counts[perm ? perm[g] : g] : 0u;
        t[k] = (s[k] + sh.L - 1) / sh.L;
        v = v + Scan3{s[k], t[k], t[k] > 1 ? 1u : 0u};
    }
    Scan3 total;
    Scan3 incl = block_scan(v, &total);
    const Scan3 tb = tile_sums[blockIdx.x];
    Scan3 run = {tb.pts + incl.pts - v.pts, tb.tsk + incl.tsk - v.tsk, tb.hvy + incl.hvy - v.hvy};
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) {
        const uint32_t g = base + k;
        if (g < sh.G) {
            if (offsets) offsets[g] = run.pts;
            task_off[g] = run.tsk;
            if (t[k] > 1) heavy[1 + run.hvy] = g;
            run = run + Scan3{s[k], t[k], t[k] > 1 ? 1u : 0u};
        }
    }
}

template <class Fr>
__global__ void msm_scatter_kernel(const Fr* __restrict__ scalars, uint64_t n, bool mont, MsmShape sh,
                                   const uint32_t* __restrict__ offsets, uint32_t* __restrict__ cursor,
                                   uint32_t* __restrict__ sorted) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned active = __activemask();
    const unsigned lane = threadIdx.x & 31;
    DigitIter it;
    load_scalar(it, scalars, i, mont);
    for (uint32_t w = 0; w < sh.nwin; w++) {
        const int32_t d = it.next(w, sh.c, sh.nwin);
        const uint32_t key = d != 0 ? w * sh.B + (uint32_t)(d < 0 ? -d : d) - 1 : 0xffffffffu;
        const unsigned peers = __match_any_sync(active, key);
        const unsigned leader = (unsigned)(__ffs(peers) - 1);
        uint32_t base = 0;
        if (key != 0xffffffffu && lane == leader) base = atomicAdd(&cursor[key], (uint32_t)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        if (key != 0xffffffffu) {
            const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
            sorted[offsets[key] + base + rank] = (uint32_t)i | (d < 0 ? 0x80000000u : 0u);
        }
    }
}

// Buckets that were split into tasks: sum their task partials.  Big ones (>= HEAVY_BIG partials, e.g. the
// one-bucket-per-window case of an all-equal witness) first get HEAVY_SPLIT CTAs each, which leave their
// slice sums in tmp[t0 / HEAVY_SPLIT + j] (t0 = first task of the bucket; slots of different big buckets
// cannot overlap because each owns >= HEAVY_BIG consecutive tasks); then one CTA per bucket finishes.
static constexpr uint32_t HEAVY_SPLIT = 16;
static constexpr uint32_t HEAVY_BIG = 16 * HEAVY_SPLIT;

template <class F>
__global__ void __launch_bounds__(MSM_RED_THREADS)
msm_reduce_heavy_stage1_kernel(const uint32_t* __restrict__ heavy, const uint32_t* __restrict__ task_off,
                               const XYZZ<F>* __restrict__ partials, XYZZ<F>* __restrict__ tmp) {   // heavy[] holds ranks
    extern __shared__ uint4 smem_raw[];
    XYZZ<F>* smem = reinterpret_cast<XYZZ<F>*>(smem_raw);
    const uint32_t nitems = heavy[0] * HEAVY_SPLIT;
    for (uint32_t it = blockIdx.x; it < nitems; it += gridDim.x) {
        const uint32_t g = heavy[1 + it / HEAVY_SPLIT], j = it % HEAVY_SPLIT;
        const uint32_t t0 = task_off[g], cnt = task_off[g + 1] - t0;
        if (cnt < HEAVY_BIG) continue;
        const uint32_t lo = (uint32_t)((uint64_t)cnt * j / HEAVY_SPLIT), hi = (uint32_t)((uint64_t)cnt * (j + 1) / HEAVY_SPLIT);
        XYZZ<F> s = cta_sum(partials + t0 + lo, hi - lo, smem);
        if (threadIdx.x == 0) st_struct(tmp + t0 / HEAVY_SPLIT + j, s);
        __syncthreads();
    }
}

template <class F>
__global__ void __launch_bounds__(MSM_RED_THREADS)
msm_reduce_heavy_kernel(const uint32_t* __restrict__ heavy, const uint32_t* __restrict__ task_off, const uint32_t* __restrict__ perm,
                        const XYZZ<F>* __restrict__ partials, const XYZZ<F>* __restrict__ tmp, XYZZ<F>* __restrict__ bucket_acc) {
    extern __shared__ uint4 smem_raw[];
    XYZZ<F>* smem = reinterpret_cast<XYZZ<F>*>(smem_raw);
    const uint32_t nheavy = heavy[0];
    for (uint32_t h = blockIdx.x; h < nheavy; h += gridDim.x) {
        const uint32_t rk = heavy[1 + h];
        const uint32_t g = perm ? perm[rk] : rk;
        const uint32_t t0 = task_off[rk], cnt = task_off[rk + 1] - t0;
        XYZZ<F> s = cnt < HEAVY_BIG ? cta_sum(partials + t0, cnt, smem) : cta_sum(tmp + t0 / HEAVY_SPLIT, HEAVY_SPLIT, smem);
        if (threadIdx.x == 0) st_struct(bucket_acc + g, s);
        __syncthreads();
    }
}

// sum of `count` XYZZ points -> affine (join of multi-GPU shard partials; final normalisation)
template <class F>
__global__ void __launch_bounds__(MSM_RED_THREADS)
group_sum_affine_kernel(const XYZZ<F>* __restrict__ pts, uint32_t count, Affine<F>* __restrict__ out) {
    extern __shared__ uint4 smem_raw[];
    XYZZ<F>* smem = reinterpret_cast<XYZZ<F>*>(smem_raw);
    XYZZ<F> s = cta_sum(pts, count, smem);
    if (threadIdx.x == 0) {
        Affine<F> a = s.to_affine();
        st_struct(out, a);
    }
}

// G2 bucket reduction stays in this unit (out-of-line multiplication): inlining 42 base multiplications per general
// addition into its kernels costs a quarter of an hour of ptxas for a ~5 % kernel
int32_t msm_bucket_reduce_g2(Ctx* c, const void* bucket_acc, MsmShape sh, uint32_t seg, void* segs, uint32_t segs_per_win, void* wins) {
    return dispatch_curve(c, [&](auto curve) {
        using F = typename decltype(curve)::Fq2;
        return msm_bucket_reduce_launch<F>(c, "msm_bucket_segments_g2", "msm_window_sum_g2", bucket_acc, sh, seg, segs, segs_per_win, wins);
    });
}

// ---------------------------------------------------------------------------------------------------
static uint32_t env_u32(const char* name, uint32_t dflt) {
    const char* v = getenv(name);
    return v ? (uint32_t)strtoul(v, nullptr, 10) : dflt;
}

// Window size: minimise  n * nwin  (bucket accumulation, mixed additions)  +  nwin * 2^(c-1) * 4.7
// (bucket reduction: two general additions per bucket at ~1.4x the cost of a mixed one, plus the
// per-segment double-and-add), subject to the bucket array staying under 4 GiB.
static MsmShape msm_shape(uint64_t n, uint32_t scalar_bits, size_t point_bytes) {
    MsmShape sh{};
    auto nwin_of = [&](uint32_t c) {
        uint32_t nw = (scalar_bits + c - 1) / c;
        // the last window keeps the recoding carry: it must fit in 2^(c-1) buckets
        if (scalar_bits - (nw - 1) * c >= c) nw += 1;
        return nw;
    };
    uint32_t best_c = 5;
    double best = 1e300;
    for (uint32_t c = 5; c <= 20; c++) {
        const uint32_t nw = nwin_of(c);
        const double buckets = (double)nw * (double)(1u << (c - 1));
        if (buckets * (double)point_bytes > 4.0 * 1024 * 1024 * 1024) break;
        const double cost = (double)n * nw + buckets * 4.7;
        if (cost < best) { best = cost; best_c = c; }
    }
    uint32_t c = env_u32("B2S_MSM_C", best_c);
    if (c < 2) c = 2;
    if (c > 24) c = 24;
    sh.c = c;
    sh.nwin = nwin_of(c);
    sh.B = 1u << (c - 1);
    sh.G = sh.nwin * sh.B;
    const uint64_t t_upper = (uint64_t)sh.nwin * n;
    uint64_t L = t_upper >> 18;
    if (L < 64) L = 64;
    L = env_u32("B2S_MSM_L", (uint32_t)L);
    sh.L = (uint32_t)L;
    sh.max_tasks = t_upper / sh.L + sh.G + 1;
    return sh;
}

template <class Curve, class F>
static int32_t msm_run_t(Ctx* c, const void* bases_dev, const void* scalars_dev, uint64_t n, bool mont, void* out_dev, void* wins_ext) {
    using Fr = typename Curve::Fr;
    using Pt = XYZZ<F>;
    if (n >= (1ull << 31)) return fail(c, B2S_ERR_INVALID_ARG, "msm: n = %llu exceeds 2^31 - 1", (unsigned long long)n);
    Pt* out = reinterpret_cast<Pt*>(out_dev);
    if (n == 0) {
        B2S_CUDA(c, cudaMemsetAsync(out, 0, sizeof(Pt), c->stream));
        return B2S_OK;
    }
    constexpr bool is_g1_t = sizeof(F) == sizeof(typename Curve::Fq);
    MsmShape sh = msm_shape(n, Curve::FrP::BITS, sizeof(Pt));
    if ((uint64_t)sh.nwin * n >= (1ull << 32)) return fail(c, B2S_ERR_INVALID_ARG, "msm: n * windows exceeds 2^32");
    // batched-affine halving rounds (msm_affine.cuh); afterwards the points are already in bucket order.
    // Rounds are worth it when buckets hold several points: R = ceil(log2(points per bucket)) rounds leave <= 2 points
    // per average bucket for the XYZZ kernel; heavy buckets (skewed scalars) shrink by 2^R as well.
    uint32_t ba_auto = 0;
    {
        const uint64_t per_bucket = ((uint64_t)sh.nwin * n) / sh.G;
        while ((1ull << ba_auto) < per_bucket) ba_auto++;
        // a round has a fixed price -- one latency-bound inversion level (~0.7 ms) plus a dozen small launches -- and
        // saves 4 (G1) / 11 (G2) base multiplications on each of its T / 2^(r+1) additions: keep the rounds that pay
        const double min_adds = is_g1_t ? 6.0e6 : 2.5e6;
        uint32_t pays = 0;
        while (pays < 16 && (double)((uint64_t)sh.nwin * n >> (pays + 1)) > min_adds) pays++;
        ba_auto = std::min(ba_auto, pays);
    }
    uint32_t ba_rounds = env_u32("B2S_MSM_AFFINE_ROUNDS", ba_auto);
    // scratch of the rounds: two output buffers, the prefix products and the lane totals (bounded by the first round)
    {
        const uint64_t t0 = std::min<uint64_t>((uint64_t)sh.nwin * n, ((uint64_t)sh.nwin * n + sh.G) / 2 + 1);
        const uint64_t need = t0 * (sizeof(Affine<F>) * 3 / 2 + sizeof(F)) + t0 / 4;
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        uint64_t pool_held = 0;
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, c->device) == cudaSuccess) {
            uint64_t reserved = 0, used = 0;
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used);
            pool_held = reserved > used ? reserved - used : 0;
        }
        if (ba_rounds && need + ((uint64_t)2 << 30) > (uint64_t)free_b + pool_held && !getenv("B2S_MSM_AFFINE_ROUNDS")) ba_rounds = 0;
    }
    if (ba_rounds && !getenv("B2S_MSM_L")) {
        // what the XYZZ kernel sees after the rounds is 2^-R of the input: cut its tasks accordingly, otherwise the heavy
        // buckets of skewed scalars (a few thousand tasks of ~1000 points) leave most of the machine idle
        uint64_t t_after = (uint64_t)sh.nwin * n;
        for (uint32_t r = 0; r < ba_rounds; r++) t_after = std::min<uint64_t>(t_after, (t_after + sh.G) / 2 + 1);
        sh.L = (uint32_t)std::max<uint64_t>(16, t_after >> 18);
        sh.max_tasks = t_after / sh.L + sh.G + 1;
    }
    const Fr* scalars = reinterpret_cast<const Fr*>(scalars_dev);
    const Affine<F>* bases = reinterpret_cast<const Affine<F>*>(bases_dev);

    const uint32_t MSM_SEG = env_u32("B2S_MSM_SEG", sh.B >= (1u << 16) ? 32u : 16u);
    const uint32_t ntiles = (sh.G + SCAN_TILE - 1) / SCAN_TILE;
    DevBuf ibuf, sorted, bucket_acc, partials, segs, wins, tiles, heavy_tmp;
    B2S_TRY(tiles.alloc(c, (size_t)ntiles * sizeof(Scan3)));
    B2S_TRY(heavy_tmp.alloc(c, ((size_t)sh.max_tasks / HEAVY_SPLIT + HEAVY_SPLIT + 1) * sizeof(Pt)));
    // u32 arrays: counts[G] cursor[G] offsets[G+1] task_off[G+1] heavy[G+1]
    const size_t ints = (size_t)5 * sh.G + 3;
    B2S_TRY(ibuf.alloc(c, ints * sizeof(uint32_t)));
    uint32_t* counts = ibuf.as<uint32_t>();
    uint32_t* cursor = counts + sh.G;
    uint32_t* offsets = cursor + sh.G;
    uint32_t* task_off = offsets + sh.G + 1;
    uint32_t* heavy = task_off + sh.G + 1;
    B2S_CUDA(c, cudaMemsetAsync(counts, 0, (size_t)2 * sh.G * sizeof(uint32_t), c->stream));
    B2S_TRY(sorted.alloc(c, (size_t)sh.nwin * n * sizeof(uint32_t)));
    B2S_TRY(bucket_acc.alloc(c, (size_t)sh.G * sizeof(Pt)));
    B2S_CUDA(c, cudaMemsetAsync(bucket_acc.p, 0, (size_t)sh.G * sizeof(Pt), c->stream));  // identity = zeros
    B2S_TRY(partials.alloc(c, (size_t)sh.max_tasks * sizeof(Pt)));
    const uint32_t segs_per_win = (sh.B + MSM_SEG - 1) / MSM_SEG;
    B2S_TRY(segs.alloc(c, (size_t)segs_per_win * sh.nwin * sizeof(Pt)));
    if (sh.nwin > 64) wins_ext = nullptr;   // caller scratch holds 64 window sums; tiny windows take the in-stream path
    if (getenv("B2S_NO_AUX")) wins_ext = nullptr;   // debugging knob: keep the Horner tail on the main stream
    if (!wins_ext) B2S_TRY(wins.alloc(c, (size_t)sh.nwin * sizeof(Pt)));
    Pt* wins_p = wins_ext ? reinterpret_cast<Pt*>(wins_ext) : wins.as<Pt>();

    B2S_LAUNCH(c, msm_count_kernel<Fr>, cdiv(n, 256), 256, 0, scalars, n, mont, sh, counts);
    const uint32_t* no_perm = nullptr;
    B2S_LAUNCH(c, msm_scan_tiles_kernel, ntiles, SCAN_THREADS, 0, counts, no_perm, sh, tiles.as<Scan3>());
    B2S_LAUNCH(c, msm_scan_spine_kernel, 1, 1024, 0, tiles.as<Scan3>(), ntiles, sh, offsets, task_off, heavy);
    B2S_LAUNCH(c, msm_scan_apply_kernel, ntiles, SCAN_THREADS, 0, counts, no_perm, sh, tiles.as<Scan3>(), offsets, task_off, heavy);
    B2S_LAUNCH(c, msm_scatter_kernel<Fr>, cdiv(n, 256), 256, 0, scalars, n, mont, sh, offsets, cursor, sorted.as<uint32_t>());
    constexpr bool is_g1 = sizeof(F) == sizeof(typename Curve::Fq);
    const void* acc_bases = bases;
    const uint32_t* acc_sorted = sorted.as<uint32_t>();
    const uint32_t* acc_offsets = offsets;
    DevBuf ba_ints, ba_out[2], ba_prefix, ba_tot, ba_bits;
    if (ba_rounds) {
        B2S_TRY(ba_ints.alloc(c, ((size_t)2 * sh.G + 1) * sizeof(uint32_t)));
        uint32_t* cnt[2] = {counts, ba_ints.as<uint32_t>()};
        uint32_t* off[2] = {offsets, ba_ints.as<uint32_t>() + sh.G};
        uint64_t t_in = (uint64_t)sh.nwin * n;
        // outputs of a round: every bucket keeps ceil(count / 2) points -- at most (t_in + G) / 2 and never more than t_in
        auto round_bound = [&](uint64_t tin) { return std::min<uint64_t>(tin, (tin + sh.G) / 2 + 1); };
        const uint64_t out_bound0 = round_bound(t_in);
        const uint64_t words_bound0 = out_bound0 / 32 + 2;
        const uint64_t tot_bound = (words_bound0 / BA_KMIN + 2) * 32;
        const uint32_t rank_tiles0 = cdiv(words_bound0, BA_SCAN_TILE);
        // bitmap | wrank | rank tile sums
        B2S_TRY(ba_bits.alloc(c, (2 * words_bound0 + rank_tiles0 + 1) * sizeof(uint32_t)));
        uint32_t* bitmap = ba_bits.as<uint32_t>();
        uint32_t* wrank = bitmap + words_bound0;
        uint32_t* rtiles = wrank + words_bound0;
        B2S_TRY(ba_prefix.alloc(c, out_bound0 * sizeof(F)));
        B2S_TRY(ba_tot.alloc(c, 2 * tot_bound * sizeof(F)));
        const uint32_t target_units = 4u * 16u * (uint32_t)c->sm_count;   // ~4 units per resident warp, handed out dynamically
        DevBuf ba_ctr;
        B2S_TRY(ba_ctr.alloc(c, (size_t)2 * ba_rounds * sizeof(uint32_t)));
        B2S_CUDA(c, cudaMemsetAsync(ba_ctr.p, 0, (size_t)2 * ba_rounds * sizeof(uint32_t), c->stream));
        // the two ping-pong output buffers, sized for the rounds that use them (even rounds write [1], odd rounds [0])
        B2S_TRY(ba_out[1].alloc(c, out_bound0 * sizeof(Affine<F>)));
        if (ba_rounds > 1) B2S_TRY(ba_out[0].alloc(c, round_bound(out_bound0) * sizeof(Affine<F>)));
        int cur = 0;
        const void* prev = nullptr;
        for (uint32_t r = 0; r < ba_rounds; r++) {
            const int nxt = cur ^ 1;
            const uint64_t out_bound = round_bound(t_in);
            const uint32_t n_words = (uint32_t)(out_bound / 32 + 2);
            const uint32_t rank_tiles = cdiv(n_words, BA_SCAN_TILE);
            B2S_LAUNCH(c, msm_ba_halve_kernel, cdiv(sh.G, 256), 256, 0, cnt[cur], sh.G, cnt[nxt]);
            B2S_LAUNCH(c, msm_scan_tiles_kernel, ntiles, SCAN_THREADS, 0, cnt[nxt], no_perm, sh, tiles.as<Scan3>());
            B2S_LAUNCH(c, msm_scan_spine_kernel, 1, 1024, 0, tiles.as<Scan3>(), ntiles, sh, off[nxt], task_off, heavy);
            B2S_LAUNCH(c, msm_scan_apply_kernel, ntiles, SCAN_THREADS, 0, cnt[nxt], no_perm, sh, tiles.as<Scan3>(), off[nxt], task_off, heavy);
            B2S_CUDA(c, cudaMemsetAsync(bitmap, 0, (size_t)n_words * sizeof(uint32_t), c->stream));
            B2S_LAUNCH(c, msm_ba_singles_kernel, cdiv(sh.G, 256), 256, 0, cnt[cur], off[nxt], sh.G, bitmap);
            B2S_LAUNCH(c, msm_ba_rank_tiles_kernel, rank_tiles, BA_SCAN_THREADS, 0, bitmap, n_words, rtiles);
            B2S_LAUNCH(c, msm_ba_rank_spine_kernel, 1, 1024, 0, rtiles, rank_tiles);
            B2S_LAUNCH(c, msm_ba_rank_apply_kernel, rank_tiles, BA_SCAN_THREADS, 0, bitmap, n_words, rtiles, wrank);
            BaRoundArgs ra{};
            ra.first = r == 0;
            ra.bases = bases; ra.sorted = sorted.as<uint32_t>(); ra.prev = prev;
            ra.bitmap = bitmap; ra.wrank = wrank; ra.t_out = off[nxt] + sh.G;
            ra.target_units = target_units;
            ra.unit_ctr = ba_ctr.as<uint32_t>() + 2 * r;
            ra.prefix = ba_prefix.p; ra.tot = ba_tot.p; ra.inv_scratch = ba_tot.as<F>() + tot_bound; ra.out = ba_out[nxt].p;
            if (is_g1) B2S_TRY(msm_ba_round_g1(c, ra));
            else B2S_TRY(msm_ba_round_g2(c, ra));
            prev = ba_out[nxt].p;
            t_in = out_bound;
            cur = nxt;
        }
        acc_bases = prev;
        acc_sorted = nullptr;
        acc_offsets = off[cur];
    }
    // task ranks by decreasing bucket size (skipped after affine rounds, which leave the natural order)
    const uint32_t* perm = nullptr;
    DevBuf perm_buf;
    if (!ba_rounds && env_u32("B2S_MSM_SIZE_SORT", 1)) {
        B2S_TRY(perm_buf.alloc(c, ((size_t)sh.G + 2 * SIZE_BINS) * sizeof(uint32_t)));
        uint32_t* pm = perm_buf.as<uint32_t>();
        uint32_t* hist = pm + sh.G;
        uint32_t* binoff = hist + SIZE_BINS;
        uint32_t* none = nullptr;
        B2S_CUDA(c, cudaMemsetAsync(hist, 0, SIZE_BINS * sizeof(uint32_t), c->stream));
        B2S_LAUNCH(c, msm_size_hist_kernel, cdiv(sh.G, 256), 256, 0, counts, sh.G, hist);
        B2S_LAUNCH(c, msm_size_scan_kernel, 1, 1024, 0, hist, binoff);
        B2S_LAUNCH(c, msm_size_scatter_kernel, cdiv(sh.G, 256), 256, 0, counts, sh.G, binoff, hist, pm);
        B2S_LAUNCH(c, msm_scan_tiles_kernel, ntiles, SCAN_THREADS, 0, counts, (const uint32_t*)pm, sh, tiles.as<Scan3>());
        B2S_LAUNCH(c, msm_scan_spine_kernel, 1, 1024, 0, tiles.as<Scan3>(), ntiles, sh, none, task_off, heavy);
        B2S_LAUNCH(c, msm_scan_apply_kernel, ntiles, SCAN_THREADS, 0, counts, (const uint32_t*)pm, sh, tiles.as<Scan3>(), none, task_off, heavy);
        perm = pm;
    }
    if (is_g1) B2S_TRY(msm_accumulate_g1(c, acc_bases, acc_sorted, acc_offsets, task_off, perm, sh, bucket_acc.p, partials.p));
    else B2S_TRY(msm_accumulate_g2(c, acc_bases, acc_sorted, acc_offsets, task_off, perm, sh, bucket_acc.p, partials.p));
    const size_t red_smem = (size_t)MSM_RED_THREADS * sizeof(Pt);
    B2S_SMEM_ATTR(c, msm_reduce_heavy_kernel<F>, red_smem);
    B2S_SMEM_ATTR(c, msm_reduce_heavy_stage1_kernel<F>, red_smem);
    B2S_LAUNCH(c, msm_reduce_heavy_stage1_kernel<F>, 4 * c->sm_count, MSM_RED_THREADS, red_smem, heavy, task_off,
               partials.as<Pt>(), heavy_tmp.as<Pt>());
    B2S_LAUNCH(c, msm_reduce_heavy_kernel<F>, 2 * c->sm_count, MSM_RED_THREADS, red_smem, heavy, task_off, perm,
               partials.as<Pt>(), heavy_tmp.as<Pt>(), bucket_acc.as<Pt>());
    // bucket reduction: compiled with the multiplication inlined (msm_acc_g1.cu / msm_acc_g2.cu), 2 general additions per bucket
    if (is_g1) B2S_TRY(msm_bucket_reduce_g1(c, bucket_acc.p, sh, MSM_SEG, segs.p, segs_per_win, wins_p));
    else B2S_TRY(msm_bucket_reduce_g2(c, bucket_acc.p, sh, MSM_SEG, segs.p, segs_per_win, wins_p));
    if (!wins_ext) {
        B2S_TRY(msm_horner(c, c->stream, is_g1 ? 1 : 2, wins_p, sh, out));
    } else {
        // tail on the aux stream: it only needs the window sums, the main stream goes on with the next MSM
        B2S_CUDA(c, cudaEventRecord(c->ev_tail, c->stream));
        B2S_CUDA(c, cudaStreamWaitEvent(c->aux, c->ev_tail, 0));
        B2S_TRY(msm_horner(c, c->aux, is_g1 ? 1 : 2, wins_p, sh, out));
        c->aux_pending = true;
    }
    return B2S_OK;
}

// Generic use of the scan kernels (setup_groth16.cu: column-sorted matrices): offsets[i] = sum_{j<i} counts[j],
// task_off[i] = sum_{j<i} ceil(counts[j] / L); both arrays have n + 1 entries.
int32_t scan_counts(Ctx* c, const uint32_t* counts, uint32_t n, uint32_t L, uint32_t* offsets, uint32_t* task_off) {
    MsmShape sh{};
    sh.G = n;
    sh.L = L;
    const uint32_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    DevBuf tiles, heavy;
    B2S_TRY(tiles.alloc(c, (size_t)ntiles * sizeof(Scan3)));
    B2S_TRY(heavy.alloc(c, ((size_t)n + 1) * sizeof(uint32_t)));
    const uint32_t* no_perm = nullptr;
    B2S_LAUNCH(c, msm_scan_tiles_kernel, ntiles, SCAN_THREADS, 0, counts, no_perm, sh, tiles.as<Scan3>());
    B2S_LAUNCH(c, msm_scan_spine_kernel, 1, 1024, 0, tiles.as<Scan3>(), ntiles, sh, offsets, task_off, heavy.as<uint32_t>());
    B2S_LAUNCH(c, msm_scan_apply_kernel, ntiles, SCAN_THREADS, 0, counts, no_perm, sh, tiles.as<Scan3>(), offsets, task_off, heavy.as<uint32_t>());
    return B2S_OK;
}

int32_t msm_join_tails(Ctx* c) {
    if (!c->aux_pending) return B2S_OK;
    B2S_CUDA(c, cudaEventRecord(c->ev_done, c->aux));
    B2S_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_done, 0));
    c->aux_pending = false;
    return B2S_OK;
}

int32_t msm_run(Ctx* c, int group, const void* bases_dev, const void* scalars_dev, uint64_t n, bool scalars_mont,
                void* out_xyzz_dev, void* wins_ext) {
    return dispatch_curve(c, [&](auto curve) {
        using C = decltype(curve);
        if (group == 1) return msm_run_t<C, typename C::Fq>(c, bases_dev, scalars_dev, n, scalars_mont, out_xyzz_dev, wins_ext);
        return msm_run_t<C, typename C::Fq2>(c, bases_dev, scalars_dev, n, scalars_mont, out_xyzz_dev, wins_ext);
    });
}

template <class Curve, class F>
static int32_t group_sum_t(Ctx* c, const void* xyzz_dev, uint32_t count, void* out_affine_dev) {
    const size_t red_smem = (size_t)MSM_RED_THREADS * sizeof(XYZZ<F>);
    B2S_CUDA(c, cudaFuncSetAttribute(group_sum_affine_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)red_smem));
    B2S_LAUNCH(c, group_sum_affine_kernel<F>, 1, MSM_RED_THREADS, red_smem, reinterpret_cast<const XYZZ<F>*>(xyzz_dev),
               count, reinterpret_cast<Affine<F>*>(out_affine_dev));
    return B2S_OK;
}

int32_t group_sum_to_affine(Ctx* c, int group, const void* xyzz_dev, uint32_t count, void* out_affine_dev) {
    return dispatch_curve(c, [&](auto curve) {
        using C = decltype(curve);
        if (group == 1) return group_sum_t<C, typename C::Fq>(c, xyzz_dev, count, out_affine_dev);
        return group_sum_t<C, typename C::Fq2>(c, xyzz_dev, count, out_affine_dev);
    });
}

}  // namespace b2s
