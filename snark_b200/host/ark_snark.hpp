// C++ mirror of the ark-snark trait family with a Groth16 instantiation over libb200snark.so.
//
//   SNARK::{circuit_specific_setup, prove}            /root/reference/snark/src/lib.rs:22-54
//   CircuitSpecificSetupSNARK::setup                  /root/reference/snark/src/lib.rs:84-93
// The Groth16 algebra follows ark-groth16 (generator / prover; not in /root/reference, SURVEY.md App. A.1,
// A.5).  The host does what the Rust host does -- synthesis and witness assignment -- and calls the C ABI
// (include/b200snark.h) for everything else: b2s_groth16_setup builds the key, b2s_groth16_prove the proof.
// No setup or prover arithmetic runs on the CPU.
#pragma once
#include <cstdlib>
#include <istream>
#include <ostream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/b200snark.h"
#include "../csrc/curves.cuh"
#include "ark_relations.hpp"
#include "ark_std_rng.hpp"

namespace ark_snark {

using ark_relations::gr1cs::ConstraintSynthesizer;
using ark_relations::gr1cs::ConstraintSystemRef;
using ark_relations::gr1cs::Matrix;
using ark_relations::gr1cs::OptimizationGoal;
using ark_relations::gr1cs::SynthesisMode;

struct BackendError : std::runtime_error {
    int32_t status;
    BackendError(int32_t st, const std::string& msg) : std::runtime_error(msg), status(st) {}
};

template <class Curve>
struct Trapdoor { typename Curve::Fr tau, alpha, beta, gamma, delta; };   // the setup's toxic waste (sampled from rng upstream)

// Keys and proofs are kept in the C-ABI layout (Montgomery limbs, affine x||y), ready for upload / comparison.
template <class Curve>
struct ProvingKey {
    uint64_t n_instance = 0, n_witness = 0, domain_size = 0;
    std::vector<uint32_t> alpha_g1, beta_g1, delta_g1, beta_g2, delta_g2, gamma_g2, gamma_abc_g1;
    std::vector<uint32_t> a_query, b_g1_query, b_g2_query, h_query, l_query;
};
template <class Curve>
struct Proof { std::vector<uint32_t> a, b, c; };

// A proving key on disk in the layout the backend uploads (Montgomery limbs, affine x||y, zeros = infinity): written once
// after setup, read back by every prover process -- the reference's "checkpoint" is the serialized key
// (snark/src/lib.rs:25-36 bounds).  This is the backend's own container, not the ark-serialize encoding (that one is
// b2s_serialize_*_compressed for points and proofs).  Header: magic, curve id, the three counts, then thirteen
// length-prefixed arrays in declaration order.  read_proving_key rejects a wrong magic / curve / inconsistent lengths.
namespace detail {
constexpr uint64_t PK_MAGIC = 0x31304b5053323042ull;   // "B02SPK01" little-endian
inline void put_u64(std::ostream& o, uint64_t v) { o.write(reinterpret_cast<const char*>(&v), 8); }
inline uint64_t get_u64(std::istream& i) {
    uint64_t v = 0;
    i.read(reinterpret_cast<char*>(&v), 8);
    if (!i) throw std::runtime_error("proving key file: truncated");
    return v;
}
inline void put_vec(std::ostream& o, const std::vector<uint32_t>& v) {
    put_u64(o, v.size());
    o.write(reinterpret_cast<const char*>(v.data()), std::streamsize(v.size() * 4));
}
inline std::vector<uint32_t> get_vec(std::istream& i, uint64_t expect_words, const char* what) {
    const uint64_t n = get_u64(i);
    if (n != expect_words) throw std::runtime_error(std::string("proving key file: unexpected length of ") + what);
    std::vector<uint32_t> v(n);
    i.read(reinterpret_cast<char*>(v.data()), std::streamsize(n * 4));
    if (!i) throw std::runtime_error("proving key file: truncated");
    return v;
}
}  // namespace detail

template <class Curve>
void write_proving_key(std::ostream& o, const ProvingKey<Curve>& pk) {
    detail::put_u64(o, detail::PK_MAGIC);
    detail::put_u64(o, uint64_t(Curve::id));
    detail::put_u64(o, pk.n_instance); detail::put_u64(o, pk.n_witness); detail::put_u64(o, pk.domain_size);
    for (const auto* v : {&pk.alpha_g1, &pk.beta_g1, &pk.delta_g1, &pk.beta_g2, &pk.delta_g2, &pk.gamma_g2, &pk.gamma_abc_g1, &pk.a_query,
                          &pk.b_g1_query, &pk.b_g2_query, &pk.h_query, &pk.l_query})
        detail::put_vec(o, *v);
    if (!o) throw std::runtime_error("proving key file: write failed");
}

template <class Curve>
ProvingKey<Curve> read_proving_key(std::istream& i) {
    if (detail::get_u64(i) != detail::PK_MAGIC) throw std::runtime_error("proving key file: bad magic");
    if (detail::get_u64(i) != uint64_t(Curve::id)) throw std::runtime_error("proving key file: written for another curve");
    ProvingKey<Curve> pk;
    pk.n_instance = detail::get_u64(i); pk.n_witness = detail::get_u64(i); pk.domain_size = detail::get_u64(i);
    if (pk.n_instance == 0 || pk.domain_size == 0 || (pk.domain_size & (pk.domain_size - 1)))
        throw std::runtime_error("proving key file: implausible header");
    const uint64_t g1 = 2 * Curve::Fq::N, g2 = 4 * Curve::Fq::N, n_vars = pk.n_instance + pk.n_witness;
    pk.alpha_g1 = detail::get_vec(i, g1, "alpha_g1"); pk.beta_g1 = detail::get_vec(i, g1, "beta_g1"); pk.delta_g1 = detail::get_vec(i, g1, "delta_g1");
    pk.beta_g2 = detail::get_vec(i, g2, "beta_g2"); pk.delta_g2 = detail::get_vec(i, g2, "delta_g2"); pk.gamma_g2 = detail::get_vec(i, g2, "gamma_g2");
    pk.gamma_abc_g1 = detail::get_vec(i, pk.n_instance * g1, "gamma_abc_g1");
    pk.a_query = detail::get_vec(i, n_vars * g1, "a_query"); pk.b_g1_query = detail::get_vec(i, n_vars * g1, "b_g1_query");
    pk.b_g2_query = detail::get_vec(i, n_vars * g2, "b_g2_query"); pk.h_query = detail::get_vec(i, (pk.domain_size - 1) * g1, "h_query");
    pk.l_query = detail::get_vec(i, pk.n_witness * g1, "l_query");
    return pk;
}

template <class Curve>
class Groth16 {
public:
    using F = typename Curve::Fr;
    using FrP = typename Curve::FrP;
    static constexpr size_t FQ_WORDS = Curve::Fq::N;

    explicit Groth16(int device = 0) {
        int32_t st = b2s_ctx_create(Curve::id, device, &ctx_);
        if (st != B2S_OK) throw BackendError(st, "b2s_ctx_create failed (no sm_100 GPU?)");
    }
    ~Groth16() { if (ctx_) b2s_ctx_destroy(ctx_); }
    Groth16(const Groth16&) = delete;
    Groth16& operator=(const Groth16&) = delete;

    // SNARK::circuit_specific_setup (snark/src/lib.rs:43-46), trapdoor supplied by the caller (upstream draws it
    // from rng).  Synthesis runs in Setup mode on the host; the key itself is built on the GPU (b2s_groth16_setup).
    ProvingKey<Curve> circuit_specific_setup(ConstraintSynthesizer<F>& circuit, const Trapdoor<Curve>& td) {
        auto cs = ConstraintSystemRef<F>::new_ref();
        cs.set_optimization_goal(OptimizationGoal::Constraints);
        cs.set_mode(SynthesisMode::Setup());
        circuit.generate_constraints(cs);
        cs.finalize();
        b2s_r1cs* mat = upload_matrices(cs);
        const size_t ell = cs.num_instance_variables(), m = cs.num_witness_variables(), n_vars = ell + m;
        const size_t N = b2s_r1cs_domain_size(mat);
        ProvingKey<Curve> pk;
        pk.n_instance = ell; pk.n_witness = m; pk.domain_size = N;
        const size_t g1 = 2 * FQ_WORDS, g2 = 4 * FQ_WORDS;
        pk.alpha_g1.resize(g1); pk.beta_g2.resize(g2); pk.gamma_g2.resize(g2); pk.delta_g2.resize(g2); pk.gamma_abc_g1.resize(ell * g1);
        const F trap[5] = {td.tau, td.alpha, td.beta, td.gamma, td.delta};
        b2s_pk* pkh = nullptr;
        int32_t st = b2s_groth16_setup(ctx_, mat, trap, &pkh, pk.alpha_g1.data(), pk.beta_g2.data(), pk.gamma_g2.data(), pk.delta_g2.data(),
                                       pk.gamma_abc_g1.data());
        b2s_r1cs_free(ctx_, mat);
        check(st);
        auto fetch = [&](int which, size_t count, size_t words) {
            std::vector<uint32_t> v(count * words);
            int32_t s2 = b2s_pk_query(ctx_, pkh, which, v.data(), v.size() * 4);
            if (s2 != B2S_OK) { b2s_pk_free(ctx_, pkh); check(s2); }
            return v;
        };
        pk.a_query = fetch(0, n_vars, g1); pk.b_g1_query = fetch(1, n_vars, g1); pk.b_g2_query = fetch(2, n_vars, g2);
        pk.h_query = fetch(3, N - 1, g1); pk.l_query = fetch(4, m, g1);
        auto k1 = fetch(5, 3, g1); auto k2 = fetch(6, 2, g2);
        pk.beta_g1.assign(k1.begin() + g1, k1.begin() + 2 * g1);
        pk.delta_g1.assign(k1.begin() + 2 * g1, k1.end());
        b2s_pk_free(ctx_, pkh);
        return pk;
    }

    // The reference's signatures, `circuit_specific_setup(circuit, &mut rng)` / `prove(&pk, circuit, &mut rng)`
    // (snark/src/lib.rs:43-54), for any generator with next_u64() (ark_std::test_rng() in tests).  prove draws r then s, as
    // ark-groth16 does.  setup draws alpha, beta, gamma, delta, tau; upstream additionally draws random G1/G2 generators
    // between delta and tau, where this backend uses the curve's standard generators -- a valid key, not upstream's bytes.
    template <class R, class = decltype(std::declval<R&>().next_u64())>
    ProvingKey<Curve> circuit_specific_setup(ConstraintSynthesizer<F>& circuit, R& rng) {
        Trapdoor<Curve> td;
        td.alpha = ark_std::rand<F>(rng); td.beta = ark_std::rand<F>(rng); td.gamma = ark_std::rand<F>(rng); td.delta = ark_std::rand<F>(rng);
        td.tau = ark_std::rand<F>(rng);
        return circuit_specific_setup(circuit, static_cast<const Trapdoor<Curve>&>(td));
    }
    template <class R, class = decltype(std::declval<R&>().next_u64())>
    Proof<Curve> prove(const ProvingKey<Curve>& pk, ConstraintSynthesizer<F>& circuit, R& rng) {
        const F r = ark_std::rand<F>(rng);
        const F s = ark_std::rand<F>(rng);
        return prove(pk, circuit, r, s);
    }

    // SNARK::prove (snark/src/lib.rs:50-54); r, s are the two field elements upstream draws from rng.
    Proof<Curve> prove(const ProvingKey<Curve>& pk, ConstraintSynthesizer<F>& circuit, const F& r, const F& s) {
        auto cs = ConstraintSystemRef<F>::new_ref();
        cs.set_optimization_goal(OptimizationGoal::Constraints);
        circuit.generate_constraints(cs);
        cs.finalize();
        b2s_r1cs* mat = upload_matrices(cs);
        b2s_pk_desc d{};
        d.n_instance = pk.n_instance; d.n_witness = pk.n_witness; d.domain_size = pk.domain_size;
        d.alpha_g1 = pk.alpha_g1.data(); d.beta_g1 = pk.beta_g1.data(); d.delta_g1 = pk.delta_g1.data();
        d.beta_g2 = pk.beta_g2.data(); d.delta_g2 = pk.delta_g2.data();
        const uint64_t n_vars = pk.n_instance + pk.n_witness;
        d.a_query = pk.a_query.data(); d.a_len = n_vars;
        d.b_g1_query = pk.b_g1_query.data(); d.b1_len = n_vars;
        d.b_g2_query = pk.b_g2_query.data(); d.b2_len = n_vars;
        d.h_query = pk.h_query.data(); d.h_len = pk.domain_size - 1;
        d.l_query = pk.l_query.data(); d.l_len = pk.n_witness;
        b2s_pk* pkh = nullptr;
        int32_t st = b2s_pk_upload(ctx_, &d, B2S_MEM_HOST, &pkh);
        if (st != B2S_OK) { b2s_r1cs_free(ctx_, mat); check(st); }
        Proof<Curve> proof;
        proof.a.resize(2 * FQ_WORDS); proof.b.resize(4 * FQ_WORDS); proof.c.resize(2 * FQ_WORDS);
        const auto& zi = cs->instance_assignment();
        const auto& zw = cs->witness_assignment();
        st = b2s_groth16_prove(ctx_, pkh, mat, zi.data(), zw.data(), &r, &s, proof.a.data(), proof.b.data(), proof.c.data());
        b2s_pk_free(ctx_, pkh);
        b2s_r1cs_free(ctx_, mat);
        check(st);
        return proof;
    }

    static F from_u64(uint64_t x) { F r = F::zero(); r.v[0] = uint32_t(x); r.v[1] = uint32_t(x >> 32); return r.to_mont(); }

private:
    // Matrix<F> (to_matrices(), constraint_system.rs:768-774) -> CSR -> device (once per circuit in a long-lived
    // prover; per call in this thin mirror)
    b2s_r1cs* upload_matrices(const ConstraintSystemRef<F>& cs) {
        // B2S_HOST_LCMAP=1: hand the flat LcMap to the device and let kernels build the CSR (b2s_r1cs_upload_lcmap),
        // skipping to_matrices().  Opt-in until the device path has been run on a B200 (DESIGN.md section 0, row f).
        const char* lc = std::getenv("B2S_HOST_LCMAP");
        if (lc && lc[0] == '1') return upload_lcmap(cs);
        const auto all = cs.to_matrices();   // BTreeMap<Label, Vec<Matrix>> upstream; Groth16 takes the R1CS entry
        const auto it = all.find(ark_relations::gr1cs::R1CS_PREDICATE_LABEL);
        if (it == all.end() || it->second.size() != 3) throw BackendError(B2S_ERR_INVALID_ARG, "constraint system has no R1CS predicate");
        const auto& mats = it->second;
        std::vector<uint64_t> rp[3];
        std::vector<uint32_t> col[3];
        std::vector<F> co[3];
        for (int k = 0; k < 3; k++) {
            rp[k].push_back(0);
            for (const auto& row : mats[k]) {
                for (const auto& [c, j] : row) { col[k].push_back(uint32_t(j)); co[k].push_back(c); }
                rp[k].push_back(col[k].size());
            }
        }
        const uint64_t* rpp[3] = {rp[0].data(), rp[1].data(), rp[2].data()};
        const uint32_t* colp[3] = {col[0].data(), col[1].data(), col[2].data()};
        const void* cop[3] = {co[0].data(), co[1].data(), co[2].data()};
        b2s_r1cs* mat = nullptr;
        check(b2s_r1cs_upload(ctx_, mats[0].size(), cs.num_instance_variables(), cs.num_witness_variables(), rpp, colp, cop, &mat));
        return mat;
    }
    b2s_r1cs* upload_lcmap(const ConstraintSystemRef<F>& cs) {
        const auto& lm = cs->lc_map();
        const auto& pool = cs->field_interner().vec();
        const auto it = cs->predicates().find(ark_relations::gr1cs::R1CS_PREDICATE_LABEL);
        if (it == cs->predicates().end()) throw BackendError(B2S_ERR_INVALID_ARG, "constraint system has no R1CS predicate");
        const auto& args = it->second.get_constraints();   // argument_lcs[k]: Variable is the raw u64 the ABI takes
        static_assert(sizeof(ark_relations::gr1cs::Variable) == 8 && sizeof(ark_relations::gr1cs::InternedField) == 4, "ABI layout");
        const uint64_t* a[3] = {reinterpret_cast<const uint64_t*>(args[0].data()), reinterpret_cast<const uint64_t*>(args[1].data()),
                                reinterpret_cast<const uint64_t*>(args[2].data())};
        b2s_r1cs* mat = nullptr;
        check(b2s_r1cs_upload_lcmap(ctx_, it->second.num_constraints(), cs.num_instance_variables(), cs.num_witness_variables(), a,
                                    lm.num_lcs(), lm.offsets().data(), reinterpret_cast<const uint64_t*>(lm.vars().data()),
                                    reinterpret_cast<const uint32_t*>(lm.coeffs().data()), pool.data(), uint32_t(pool.size()), &mat));
        return mat;
    }
    void check(int32_t st) { if (st != B2S_OK) throw BackendError(st, b2s_last_error(ctx_)); }
    b2s_ctx* ctx_ = nullptr;
};

}  // namespace ark_snark
