// C++ mirror of the ark-snark trait family with a Groth16 instantiation over libb200snark.so.
//
//   SNARK::{circuit_specific_setup, prove}            /root/reference/snark/src/lib.rs:22-54
//   CircuitSpecificSetupSNARK::setup                  /root/reference/snark/src/lib.rs:84-93
// The Groth16 algebra follows ark-groth16 (generator / prover; not in /root/reference, SURVEY.md App. A.1,
// A.5).  The host does what the Rust host does -- synthesis, witness assignment, and at setup time the
// O(nnz) evaluation of the QAP at tau -- and calls the C ABI (include/b200snark.h) for every group operation:
// b2s_fixed_base_g{1,2} for the key, b2s_groth16_prove for the proof.  No prover arithmetic runs on the CPU.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/b200snark.h"
#include "../csrc/curves.cuh"
#include "ark_relations.hpp"

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

    // SNARK::circuit_specific_setup (snark/src/lib.rs:43-46), trapdoor supplied by the caller.
    ProvingKey<Curve> circuit_specific_setup(ConstraintSynthesizer<F>& circuit, const Trapdoor<Curve>& td) {
        auto cs = ConstraintSystemRef<F>::new_ref();
        cs.set_optimization_goal(OptimizationGoal::Constraints);
        cs.set_mode(SynthesisMode::Setup());
        circuit.generate_constraints(cs);
        cs.finalize();
        const auto mats = cs.to_matrices();
        const size_t n = cs.num_constraints(), ell = cs.num_instance_variables(), m = cs.num_witness_variables();
        const size_t n_vars = ell + m;
        size_t log_n = 0;
        while ((size_t(1) << log_n) < n + ell) log_n++;
        const size_t N = size_t(1) << log_n;
        // Lagrange coefficients at tau over the radix-2 domain: u_i = Z(tau) w^i / (N (tau - w^i))
        F w = fr_const(&FrP::root);
        for (size_t i = log_n; i < size_t(FrP::TWO_ADICITY); i++) w = w.sqr();
        F zt = pow_u64(td.tau, N) - F::one();
        if (zt.is_zero()) throw std::domain_error("tau lies in the evaluation domain");
        F n_inv = F::one(), half = fr_const(&FrP::half);
        for (size_t i = 0; i < log_n; i++) n_inv = n_inv * half;
        std::vector<F> wi(N), den(N), u(N);
        F cur = F::one();
        for (size_t i = 0; i < N; i++) { wi[i] = cur; den[i] = td.tau - cur; cur = cur * w; }
        batch_invert(den);
        for (size_t i = 0; i < N; i++) u[i] = zt * wi[i] * n_inv * den[i];
        std::vector<F> a(n_vars, F::zero()), b(n_vars, F::zero()), c(n_vars, F::zero());
        for (size_t i = 0; i < n; i++) {
            for (const auto& [co, col] : mats[0][i]) a[col] = a[col] + u[i] * co;
            for (const auto& [co, col] : mats[1][i]) b[col] = b[col] + u[i] * co;
            for (const auto& [co, col] : mats[2][i]) c[col] = c[col] + u[i] * co;
        }
        for (size_t i = 0; i < ell; i++) a[i] = a[i] + u[n + i];   // input-consistency rows (LibsnarkReduction)
        const F dinv = td.delta.inverse(), ginv = td.gamma.inverse();
        std::vector<F> hq(N - 1), lq(m), abc(ell);
        F t = zt * dinv;
        for (size_t i = 0; i + 1 < N; i++) { hq[i] = t; t = t * td.tau; }
        for (size_t j = 0; j < n_vars; j++) {
            F v = td.beta * a[j] + td.alpha * b[j] + c[j];
            if (j < ell) abc[j] = v * ginv; else lq[j - ell] = v * dinv;
        }
        ProvingKey<Curve> pk;
        pk.n_instance = ell; pk.n_witness = m; pk.domain_size = N;
        pk.a_query = fixed_base(1, a); pk.b_g1_query = fixed_base(1, b); pk.b_g2_query = fixed_base(2, b);
        pk.h_query = fixed_base(1, hq); pk.l_query = fixed_base(1, lq); pk.gamma_abc_g1 = fixed_base(1, abc);
        pk.alpha_g1 = fixed_base(1, {td.alpha}); pk.beta_g1 = fixed_base(1, {td.beta}); pk.delta_g1 = fixed_base(1, {td.delta});
        pk.beta_g2 = fixed_base(2, {td.beta}); pk.delta_g2 = fixed_base(2, {td.delta}); pk.gamma_g2 = fixed_base(2, {td.gamma});
        return pk;
    }

    // SNARK::prove (snark/src/lib.rs:50-54); r, s are the two field elements upstream draws from rng.
    Proof<Curve> prove(const ProvingKey<Curve>& pk, ConstraintSynthesizer<F>& circuit, const F& r, const F& s) {
        auto cs = ConstraintSystemRef<F>::new_ref();
        cs.set_optimization_goal(OptimizationGoal::Constraints);
        circuit.generate_constraints(cs);
        cs.finalize();
        const auto mats = cs.to_matrices();
        // Matrix<F> -> CSR (once per circuit in a long-lived prover; per call in this thin mirror)
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
        check(b2s_r1cs_upload(ctx_, cs.num_constraints(), cs.num_instance_variables(), cs.num_witness_variables(), rpp, colp, cop, &mat));
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
    static F fr_const(uint32_t (*f)(int)) { F r; for (int i = 0; i < F::N; i++) r.v[i] = f(i); return r; }
    static F pow_u64(const F& b, uint64_t e) { return b.pow_u64(e); }
    static void batch_invert(std::vector<F>& v) {   // Montgomery's trick; zeros are not expected here
        std::vector<F> pre(v.size());
        F acc = F::one();
        for (size_t i = 0; i < v.size(); i++) { pre[i] = acc; acc = acc * v[i]; }
        F inv = acc.inverse();
        for (size_t i = v.size(); i-- > 0;) { F t = inv * pre[i]; inv = inv * v[i]; v[i] = t; }
    }
    std::vector<uint32_t> fixed_base(int group, const std::vector<F>& scalars) {
        std::vector<uint32_t> out(scalars.size() * (group == 1 ? 2 : 4) * FQ_WORDS);
        if (scalars.empty()) return out;
        check(group == 1 ? b2s_fixed_base_g1(ctx_, scalars.data(), scalars.size(), 1, B2S_MEM_HOST, out.data())
                         : b2s_fixed_base_g2(ctx_, scalars.data(), scalars.size(), 1, B2S_MEM_HOST, out.data()));
        return out;
    }
    void check(int32_t st) { if (st != B2S_OK) throw BackendError(st, b2s_last_error(ctx_)); }
    b2s_ctx* ctx_ = nullptr;
};

}  // namespace ark_snark
