// Element-wise field / group kernels behind b2s_field_op / b2s_group_op: they run the exact device
// templates the hot-path kernels use (ff.cuh / ec.cuh) on arrays, so tests/test_gpu_field.py can pin
// the device arithmetic (PTX carry chains, special cases of the addition law) against the oracle.
#include "common.cuh"

namespace b2s {

template <class F>
__global__ void field_op_kernel(int op, const F* a, const F* b, F* out, uint64_t count) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    F x = a[i], y = b[i], r;
    switch (op) {
        case 0: r = x * y; break;
        case 1: r = x + y; break;
        case 2: r = x - y; break;
        case 3: r = x.inverse(); break;
        case 4: r = x.neg(); break;
        case 5: r = x.to_mont(); break;
        case 6: r = x.from_mont(); break;
        case 7: r = x.sqr(); break;
        default: r = F::zero();
    }
    out[i] = r;
}

template <class F>
static int32_t field_op_t(Ctx* c, int op, const void* a, const void* b, void* out, uint64_t count) {
    InBuf A, B;
    B2S_TRY(A.bind(c, a, count * sizeof(F), B2S_MEM_HOST));
    B2S_TRY(B.bind(c, b, count * sizeof(F), B2S_MEM_HOST));
    DevBuf O;
    B2S_TRY(O.alloc(c, count * sizeof(F)));
    if (count) B2S_LAUNCH(c, field_op_kernel<F>, cdiv(count, 128), 128, 0, op, A.as<F>(), B.as<F>(), O.as<F>(), count);
    if (count) B2S_CUDA(c, cudaMemcpyAsync(out, O.p, count * sizeof(F), cudaMemcpyDeviceToHost, c->stream));
    B2S_CUDA(c, cudaStreamSynchronize(c->stream));
    return B2S_OK;
}

int32_t field_op_run(Ctx* c, int field, int op, const void* a, const void* b, void* out, uint64_t count) {
    return dispatch_curve(c, [&](auto curve) {
        using C = decltype(curve);
        if (field == 0) return field_op_t<typename C::Fq>(c, op, a, b, out, count);
        if (field == 1) return field_op_t<typename C::Fr>(c, op, a, b, out, count);
        return fail(c, B2S_ERR_INVALID_ARG, "field_op: field must be 0 (Fq) or 1 (Fr)");
    });
}

template <class F, class Fr>
__global__ void group_op_kernel(int op, const Affine<F>* a, const Affine<F>* b, const Fr* k, Affine<F>* out, uint64_t count) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    Affine<F> pa = a[i], pb = b[i];
    XYZZ<F> r = XYZZ<F>::from_affine(pa);
    switch (op) {
        case 0: r.add_affine(pb); break;
        case 1: {
            // general addition with non-trivial denominators on both sides: (2a - a) + (2b - b)
            XYZZ<F> x = XYZZ<F>::from_affine(pa).dbl(); x.add_affine(pa.neg());
            XYZZ<F> y = XYZZ<F>::from_affine(pb).dbl(); y.add_affine(pb.neg());
            x.add(y); r = x; break;
        }
        case 2: r = r.dbl(); break;
        case 3: { Fr s = k[i]; r = scalar_mul_words(r, s.v, Fr::N); break; }
    }
    out[i] = r.to_affine();
}

template <class Curve, class F>
static int32_t group_op_t(Ctx* c, int op, const void* a, const void* b, const void* k, void* out, uint64_t count) {
    using Fr = typename Curve::Fr;
    InBuf A, B, K;
    B2S_TRY(A.bind(c, a, count * sizeof(Affine<F>), B2S_MEM_HOST));
    B2S_TRY(B.bind(c, b, count * sizeof(Affine<F>), B2S_MEM_HOST));
    B2S_TRY(K.bind(c, k, count * sizeof(Fr), B2S_MEM_HOST));
    DevBuf O;
    B2S_TRY(O.alloc(c, count * sizeof(Affine<F>)));
    if (count) {
        B2S_LAUNCH(c, (group_op_kernel<F, Fr>), cdiv(count, 64), 64, 0, op, A.as<Affine<F>>(), B.as<Affine<F>>(), K.as<Fr>(),
                   O.as<Affine<F>>(), count);
        B2S_CUDA(c, cudaMemcpyAsync(out, O.p, count * sizeof(Affine<F>), cudaMemcpyDeviceToHost, c->stream));
    }
    B2S_CUDA(c, cudaStreamSynchronize(c->stream));
    return B2S_OK;
}

int32_t group_op_run(Ctx* c, int group, int op, const void* a, const void* b, const void* k, void* out, uint64_t count) {
    return dispatch_curve(c, [&](auto curve) {
        using C = decltype(curve);
        if (group == 1) return group_op_t<C, typename C::Fq>(c, op, a, b, k, out, count);
        return group_op_t<C, typename C::Fq2>(c, op, a, b, k, out, count);
    });
}

}  // namespace b2s
