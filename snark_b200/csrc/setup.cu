// Fixed-base batch multiplication out[i] = k_i * G: the group part of
// `CircuitSpecificSetupSNARK::setup` (/root/reference/snark/src/lib.rs:84-93; ark-groth16
// generator, Appendix A.5: a_query[j] = A_j(tau) G1, ...), also used to make synthetic bases with known
// discrete logs for the full-size MSM parity checks.
//
// 8-bit windows: T[j][d] = d * 2^(8j) * G (32 x 256 affine entries, < 1.6 MiB, L2 resident), so a
// scalar costs at most 32 mixed additions plus one normalisation.
#include "common.cuh"

namespace b2s {

static constexpr int FB_WIN = 8;
static constexpr int FB_NWIN = 32;

template <class Curve, class F>
struct GenOf;
template <class Curve>
struct GenOf<Curve, typename Curve::Fq> {
    __device__ static Affine<typename Curve::Fq> get() { return Curve::g1_generator(); }
};
template <class Curve>
struct GenOf<Curve, typename Curve::Fq2> {
    __device__ static Affine<typename Curve::Fq2> get() { return Curve::g2_generator(); }
};

// thread (j, d): T[j][d] = d * 2^(8 j) * G
template <class Curve, class F>
__global__ void fixed_base_table_kernel(Affine<F>* table) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= FB_NWIN * (1u << FB_WIN)) return;
    const uint32_t j = t >> FB_WIN, d = t & ((1u << FB_WIN) - 1);
    if (d == 0) { table[t] = Affine<F>::inf(); return; }
    XYZZ<F> base = XYZZ<F>::from_affine(GenOf<Curve, F>::get());
    for (uint32_t i = 0; i < j * FB_WIN; i++) base = base.dbl();
    XYZZ<F> acc = XYZZ<F>::identity();
    for (int b = FB_WIN - 1; b >= 0; b--) {
        acc = acc.dbl();
        if ((d >> b) & 1) acc.add(base);
    }
    table[t] = acc.to_affine();
}

template <class F, class Fr>
__global__ void __launch_bounds__(128)
fixed_base_kernel(const Affine<F>* __restrict__ table, const Fr* __restrict__ scalars, uint64_t n, bool mont,
                  Affine<F>* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr s = scalars[i];
    if (mont) s = s.from_mont();
    XYZZ<F> acc = XYZZ<F>::identity();
#pragma unroll 1
    for (int j = 0; j < FB_NWIN; j++) {
        const uint32_t d = (s.v[j >> 2] >> ((j & 3) * 8)) & 0xffu;
        if (d) acc.add_affine(table[j * (1 << FB_WIN) + d]);
    }
    out[i] = acc.to_affine();
}

template <class Curve, class F>
static int32_t fixed_base_t(Ctx* c, int gi, const void* scalars_dev, uint64_t n, bool mont, void* out_dev) {
    using Fr = typename Curve::Fr;
    if (!c->fixed_base_tables[gi]) {
        void* p = nullptr;
        B2S_CUDA(c, cudaMalloc(&p, sizeof(Affine<F>) * FB_NWIN * (1 << FB_WIN)));
        c->fixed_base_tables[gi] = p;
        B2S_LAUNCH(c, (fixed_base_table_kernel<Curve, F>), cdiv(FB_NWIN * (1 << FB_WIN), 64), 64, 0,
                   reinterpret_cast<Affine<F>*>(p));
    }
    if (n == 0) return B2S_OK;
    B2S_LAUNCH(c, (fixed_base_kernel<F, Fr>), cdiv(n, 128), 128, 0,
               reinterpret_cast<const Affine<F>*>(c->fixed_base_tables[gi]), reinterpret_cast<const Fr*>(scalars_dev), n,
               mont, reinterpret_cast<Affine<F>*>(out_dev));
    return B2S_OK;
}

int32_t fixed_base_run(Ctx* c, int group, const void* scalars_dev, uint64_t n, bool mont, void* out_dev) {
    return dispatch_curve(c, [&](auto curve) {
        using C = decltype(curve);
        if (group == 1) return fixed_base_t<C, typename C::Fq>(c, 0, scalars_dev, n, mont, out_dev);
        return fixed_base_t<C, typename C::Fq2>(c, 1, scalars_dev, n, mont, out_dev);
    });
}

void fixed_base_free(Ctx* c) {
    for (int i = 0; i < 2; i++) {
        if (c->fixed_base_tables[i]) cudaFree(c->fixed_base_tables[i]);
        c->fixed_base_tables[i] = nullptr;
    }
}

}  // namespace b2s
