// Compressed wire format of group elements and proofs: what
// `CanonicalSerialize::serialize_compressed` gives for the associated types of `SNARK`
// (/root/reference/snark/src/lib.rs:25-36 bounds; encodings of ark-serialize / ark-bls12-381 / ark-bn254, not in
// /root/reference):
//   BLS12-381 (zcash / IETF form): x big-endian; top bits of byte 0: 0x80 compressed, 0x40 infinity, 0x20 y is the
//                                  lexicographically larger root; G2 writes x.c1 || x.c0
//   BN254 (ark-ec SWFlags):        x little-endian; top bits of the LAST byte: 0x80 y > -y, 0x40 infinity; G2 writes
//                                  x.c0 || x.c1
//   "larger" compares canonical integers; for Fq2, c1 first then c0.   Proof = A || B || C.
// The GPU turns Montgomery limbs into canonical ones and decides the sign bit (field arithmetic stays on the
// device); the host only orders bytes.  Known answers: the standard compressed BLS12-381 generators
// (tests/test_gpu_serialize.py).
#include "common.cuh"

namespace b2s {

struct CanonPoint { uint32_t x[24]; uint32_t flags; uint32_t pad[3]; };   // x: up to 2 x 12 limbs; flags: 1 = inf, 2 = y larger

template <class B>
__device__ __forceinline__ int cmp_canon(const B& a, const B& b) {   // canonical (non-Montgomery) values
    for (int i = B::N - 1; i >= 0; i--) {
        if (a.v[i] != b.v[i]) return a.v[i] > b.v[i] ? 1 : -1;
    }
    return 0;
}
template <class P>
__device__ __forceinline__ bool y_is_larger(const Fp<P>& y) {
    const Fp<P> a = y.from_mont(), b = y.neg().from_mont();
    return cmp_canon(a, b) > 0;
}
template <class P>
__device__ __forceinline__ bool y_is_larger(const Fp2<P>& y) {
    const Fp2<P> n = y.neg();
    const int c1 = cmp_canon(y.c1.from_mont(), n.c1.from_mont());
    if (c1 != 0) return c1 > 0;
    return cmp_canon(y.c0.from_mont(), n.c0.from_mont()) > 0;
}
template <class P>
__device__ __forceinline__ void put_x(CanonPoint& o, const Fp<P>& x) {
    const Fp<P> c = x.from_mont();
    for (int i = 0; i < Fp<P>::N; i++) o.x[i] = c.v[i];
}
template <class P>
__device__ __forceinline__ void put_x(CanonPoint& o, const Fp2<P>& x) {
    const Fp<P> c0 = x.c0.from_mont(), c1 = x.c1.from_mont();
    for (int i = 0; i < Fp<P>::N; i++) { o.x[i] = c0.v[i]; o.x[Fp<P>::N + i] = c1.v[i]; }
}

template <class F>
__global__ void canon_points_kernel(const Affine<F>* pts, uint32_t count, CanonPoint* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const Affine<F> p = pts[i];
    CanonPoint o;
    for (int k = 0; k < 24; k++) o.x[k] = 0;
    o.pad[0] = o.pad[1] = o.pad[2] = 0;
    if (p.is_inf()) o.flags = 1;
    else { put_x(o, p.x); o.flags = y_is_larger(p.y) ? 2u : 0u; }
    out[i] = o;
}

// host: bytes of `count` points of `group` (HOST affine Montgomery in) -> compressed bytes
int32_t serialize_points(Ctx* c, int group, const void* affine_host, uint32_t count, uint8_t* out, uint64_t cap) {
    const bool bls = c->curve == B2S_CURVE_BLS12_381;
    const size_t fq = bls ? 48 : 32, words = fq / 4;
    const size_t in_bytes = (group == 1 ? 2 : 4) * fq, out_bytes = (group == 1 ? 1 : 2) * fq;
    if ((uint64_t)count * out_bytes > cap) return fail(c, B2S_ERR_INVALID_ARG, "serialize: output buffer too small");
    if (count == 0) return B2S_OK;
    InBuf in;
    B2S_TRY(in.bind(c, affine_host, (size_t)count * in_bytes, B2S_MEM_HOST));
    DevBuf d;
    B2S_TRY(d.alloc(c, (size_t)count * sizeof(CanonPoint)));
    int32_t st = dispatch_curve(c, [&](auto curve) {
        using C = decltype(curve);
        if (group == 1) B2S_LAUNCH(c, canon_points_kernel<typename C::Fq>, cdiv(count, 64), 64, 0, in.as<Affine<typename C::Fq>>(), count, d.as<CanonPoint>());
        else B2S_LAUNCH(c, canon_points_kernel<typename C::Fq2>, cdiv(count, 64), 64, 0, in.as<Affine<typename C::Fq2>>(), count, d.as<CanonPoint>());
        return (int32_t)B2S_OK;
    });
    B2S_TRY(st);
    std::vector<CanonPoint> h(count);
    B2S_CUDA(c, cudaMemcpyAsync(h.data(), d.p, (size_t)count * sizeof(CanonPoint), cudaMemcpyDeviceToHost, c->stream));
    B2S_CUDA(c, cudaStreamSynchronize(c->stream));
    for (uint32_t i = 0; i < count; i++) {
        uint8_t* o = out + (size_t)i * out_bytes;
        const CanonPoint& p = h[i];
        const uint8_t* le0 = reinterpret_cast<const uint8_t*>(p.x);             // c0 (or x) little-endian bytes
        const uint8_t* le1 = reinterpret_cast<const uint8_t*>(p.x + words);     // c1
        if (bls) {
            // big-endian; G2: c1 then c0
            if (group == 1) for (size_t b = 0; b < fq; b++) o[b] = le0[fq - 1 - b];
            else for (size_t b = 0; b < fq; b++) { o[b] = le1[fq - 1 - b]; o[fq + b] = le0[fq - 1 - b]; }
            o[0] |= 0x80;
            if (p.flags & 1) o[0] |= 0x40;
            if (p.flags & 2) o[0] |= 0x20;
        } else {
            // little-endian; G2: c0 then c1; flags on the last byte
            memcpy(o, le0, fq);
            if (group == 2) memcpy(o + fq, le1, fq);
            if (p.flags & 1) o[out_bytes - 1] |= 0x40;
            if (p.flags & 2) o[out_bytes - 1] |= 0x80;
        }
    }
    return B2S_OK;
}

}  // namespace b2s
