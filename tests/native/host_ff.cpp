// Host-side harness: compiles the SAME field / curve templates the kernels use (ff.cuh, ec.cuh) with
// the PTX carry flag emulated, and exposes them to ctypes so the limb schedules can be checked
// against the big-int oracle on a machine without a GPU.  Test infrastructure only.
#include <cstdint>
#include <cstring>
#include "../../snark_b200/csrc/field_params.h"

using namespace b2s;

template <class P>
static void binop(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
    Fp<P> x, y, r;
    memcpy(x.v, a, sizeof(x.v));
    memcpy(y.v, b, sizeof(y.v));
    switch (op) {
        case 0: r = x * y; break;
        case 1: r = x + y; break;
        case 2: r = x - y; break;
        case 3: r = x.inverse(); break;
        case 4: r = x.neg(); break;
        case 5: r = x.to_mont(); break;
        case 6: r = x.from_mont(); break;
        case 7: r = x.sqr(); break;
        default: r = Fp<P>::zero();
    }
    memcpy(out, r.v, sizeof(r.v));
}

extern "C" int ht_field_limbs(int field) { return field == 0 ? 12 : 8; }

// field: 0 BlsFq, 1 BlsFr, 2 BnFq, 3 BnFr
extern "C" void ht_field_op(int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* out, int count) {
    int n = ht_field_limbs(field);
    for (int i = 0; i < count; i++) {
        const uint32_t *x = a + (size_t)i * n, *y = b + (size_t)i * n;
        uint32_t* o = out + (size_t)i * n;
        switch (field) {
            case 0: binop<BlsFqP>(op, x, y, o); break;
            case 1: binop<BlsFrP>(op, x, y, o); break;
            case 2: binop<BnFqP>(op, x, y, o); break;
            case 3: binop<BnFrP>(op, x, y, o); break;
        }
    }
}

extern "C" void ht_field_const(int field, int which, uint32_t* out) {
    int n = ht_field_limbs(field);
    for (int i = 0; i < n; i++) {
        uint32_t v = 0;
        switch (field) {
            case 0: v = which == 0 ? BlsFqP::mod(i) : which == 1 ? BlsFqP::r1(i) : BlsFqP::r2(i); break;
            case 1: v = which == 0 ? BlsFrP::mod(i) : which == 1 ? BlsFrP::r1(i) : which == 2 ? BlsFrP::r2(i) : which == 3 ? BlsFrP::gen(i) : BlsFrP::root(i); break;
            case 2: v = which == 0 ? BnFqP::mod(i) : which == 1 ? BnFqP::r1(i) : BnFqP::r2(i); break;
            case 3: v = which == 0 ? BnFrP::mod(i) : which == 1 ? BnFrP::r1(i) : which == 2 ? BnFrP::r2(i) : which == 3 ? BnFrP::gen(i) : BnFrP::root(i); break;
        }
        out[i] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// group law (ec.cuh) on the host
// ---------------------------------------------------------------------------------------------
#include "../../snark_b200/csrc/curves.cuh"

template <class F>
static void ec_op(int op, const uint32_t* a, const uint32_t* b, const uint32_t* k, int kwords, uint32_t* out) {
    Affine<F> pa, pb;
    memcpy(&pa, a, sizeof(pa));
    memcpy(&pb, b, sizeof(pb));
    XYZZ<F> r = XYZZ<F>::from_affine(pa);
    switch (op) {
        case 0: r.add_affine(pb); break;                       // mixed add
        case 1: r.add(XYZZ<F>::from_affine(pb)); break;        // general add (both Z = 1)
        case 2: r = r.dbl(); break;
        case 3: r = scalar_mul_words(r, k, kwords); break;
        case 4: {                                              // general add with non-trivial Z on both sides
            XYZZ<F> x = XYZZ<F>::from_affine(pa).dbl(); x.add_affine(pa);   // 3a
            XYZZ<F> y = XYZZ<F>::from_affine(pb).dbl();                     // 2b
            x.add(y); r = x; break;                                          // 3a + 2b
        }
        case 5: {                                              // (a + b) + (-(a + b)) + a via general adds
            XYZZ<F> x = XYZZ<F>::from_affine(pa); x.add_affine(pb);
            XYZZ<F> y = x.neg(); XYZZ<F> z = x; z.add(y);                   // identity
            z.add_affine(pa); r = z; break;
        }
        case 6: {                                              // doubling through the general add path
            XYZZ<F> x = XYZZ<F>::from_affine(pa).dbl(); XYZZ<F> y = x; x.add(y); r = x; break;  // 4a
        }
    }
    Affine<F> o = r.to_affine();
    memcpy(out, &o, sizeof(o));
}

// curve: 0 bls12-381, 1 bn254; group: 1 or 2
extern "C" void ht_ec_op(int curve, int group, int op, const uint32_t* a, const uint32_t* b, const uint32_t* k,
                          int kwords, uint32_t* out) {
    if (curve == 0 && group == 1) ec_op<Bls12_381::Fq>(op, a, b, k, kwords, out);
    if (curve == 0 && group == 2) ec_op<Bls12_381::Fq2>(op, a, b, k, kwords, out);
    if (curve == 1 && group == 1) ec_op<Bn254::Fq>(op, a, b, k, kwords, out);
    if (curve == 1 && group == 2) ec_op<Bn254::Fq2>(op, a, b, k, kwords, out);
}

extern "C" void ht_generator(int curve, int group, uint32_t* out) {
    if (curve == 0 && group == 1) { auto g = Bls12_381::g1_generator(); memcpy(out, &g, sizeof(g)); }
    if (curve == 0 && group == 2) { auto g = Bls12_381::g2_generator(); memcpy(out, &g, sizeof(g)); }
    if (curve == 1 && group == 1) { auto g = Bn254::g1_generator(); memcpy(out, &g, sizeof(g)); }
    if (curve == 1 && group == 2) { auto g = Bn254::g2_generator(); memcpy(out, &g, sizeof(g)); }
}

// ---------------------------------------------------------------------------------------------
// LcMap -> CSR (snark_b200/csrc/lcmap.cuh): the kernels' per-row functions run over all (matrix, row) pairs in
// the order the GPU grid covers them, with the scan done serially.  Returns the OR of the error bits.
// Outputs are caller-allocated: row_ptr[k] n_rows + 1 entries; col / coeff_id with room for `cap` entries each.
#include "../../snark_b200/csrc/lcmap.cuh"

extern "C" uint32_t ht_lcmap_csr(uint64_t n_rows, uint64_t n_instance, uint64_t n_vars, const uint64_t* a0, const uint64_t* a1,
                                 const uint64_t* a2, uint64_t n_lcs, const uint64_t* lc_offsets, const uint64_t* lc_vars,
                                 const uint32_t* lc_coeffs, const uint8_t* pool_is_zero, uint32_t pool_len, uint64_t cap,
                                 uint64_t* rp0, uint64_t* rp1, uint64_t* rp2, uint32_t* col0, uint32_t* col1, uint32_t* col2,
                                 uint32_t* id0, uint32_t* id1, uint32_t* id2) {
    b2s::lcmap::View v{lc_offsets, lc_vars, lc_coeffs, pool_is_zero, n_lcs, pool_len, n_instance, n_vars};
    const uint64_t* args[3] = {a0, a1, a2};
    uint64_t* rp[3] = {rp0, rp1, rp2};
    uint32_t* col[3] = {col0, col1, col2};
    uint32_t* id[3] = {id0, id1, id2};
    uint32_t err = 0;
    for (int k = 0; k < 3; k++) {
        uint64_t at = 0;
        for (uint64_t r = 0; r < n_rows; r++) {      // count + exclusive scan
            rp[k][r] = at;
            at += b2s::lcmap::count_row(v, args[k][r], &err);
        }
        rp[k][n_rows] = at;
        if (at > cap) return err | 0x80000000u;
        for (uint64_t r = 0; r < n_rows; r++) b2s::lcmap::fill_row(v, args[k][r], col[k] + rp[k][r], id[k] + rp[k][r]);
    }
    return err;
}
