// Host-side synthesis benchmark of the C++ mirror, after the reference's only benchmark example
// (/root/reference/relations/examples/bench.rs:14-108): BenchCircuit with 1..10-term unit-coefficient linear
// combinations over the last 10 variables, an extra symbolic LC on every second constraint, Prove mode with
// construct_matrices = true and generate_lc_assignments = false, OptimizationGoal::Constraints, timed over
// generate_constraints + finalize.  Also times to_matrices() (the export the LcMap upload avoids).
// Not on the GPU path: synthesis stays on the host by design; this is the number that sits NEXT to a proof time.
//   g++ -O2 -std=c++17 -o /tmp/bench_synthesis tools/bench_synthesis.cpp && /tmp/bench_synthesis 20
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "../snark_b200/csrc/curves.cuh"
#include "../snark_b200/host/ark_relations.hpp"
#include "../snark_b200/host/ark_std_rng.hpp"

using namespace ark_relations::gr1cs;
using F = b2s::Bls12_381::Fr;
constexpr size_t NUM_COEFFS_IN_LC = 10;

// rand_core's default seed_from_u64 (PCG32 expansion of the state into the 32-byte seed), as recalled
static ark_std::StdRng seed_from_u64(uint64_t state) {
    ark_std::StdRng::Seed seed{};
    for (int c = 0; c < 8; c++) {
        state = state * 6364136223846793005ull + 11634580027462260723ull;
        const uint32_t xorshifted = uint32_t(((state >> 18) ^ state) >> 27), rot = uint32_t(state >> 59);
        const uint32_t x = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
        for (int b = 0; b < 4; b++) seed[4 * c + b] = uint8_t(x >> (8 * b));
    }
    return ark_std::StdRng::from_seed(seed);
}
// uniform in [lo, hi) -- the reference uses rand's gen_range; the exact draw does not matter for a timing
static size_t range(ark_std::StdRng& r, size_t lo, size_t hi) { return lo + size_t(r.next_u64() % uint64_t(hi - lo)); }

struct BenchCircuit : ConstraintSynthesizer<F> {
    F a, b, c;
    size_t num_constraints;
    void generate_constraints(ConstraintSystemRef<F> cs) override {
        std::vector<Variable> variables = {cs.new_witness_variable([&] { return a; }), cs.new_witness_variable([&] { return b; }),
                                           cs.new_witness_variable([&] { return c; })};
        variables.reserve(3 * num_constraints + 3);
        auto rng_a = seed_from_u64(0);
        auto rng_b = seed_from_u64(rng_a.next_u64());
        auto rng_c = seed_from_u64(rng_a.next_u64());
        const F one = F::one();
        for (size_t i = 0; i < num_constraints; i++) {
            const size_t cur = std::min<size_t>(variables.size(), 10), lower = variables.size() - cur, upper = variables.size();
            const size_t a_size = range(rng_a, 1, NUM_COEFFS_IN_LC + 1);
            auto a_i = [&] {
                LinearCombination<F> l;
                for (size_t k = 0; k < a_size; k++) l.terms.emplace_back(one, variables[range(rng_a, lower, upper)]);
                return l;
            };
            const size_t b_size = range(rng_b, 1, NUM_COEFFS_IN_LC + 1);
            auto b_i = [&] {
                LinearCombination<F> l;
                for (size_t k = 0; k < b_size; k++) l.terms.emplace_back(one, variables[range(rng_b, lower, upper)]);
                return l;
            };
            const Variable c_i = variables[range(rng_c, lower, upper)];
            auto c_lc = [&] { return LinearCombination<F>::from(c_i); };
            if (i % 2 == 0) {
                const Variable extra = cs.new_lc([&] {
                    LinearCombination<F> l;
                    for (size_t k = 0; k < a_size; k++) l.terms.emplace_back(one, variables[range(rng_c, lower, upper)]);
                    return l;
                });
                cs.enforce_r1cs_constraint([&] { return a_i() + extra; }, b_i, c_lc);
            } else {
                cs.enforce_r1cs_constraint(a_i, b_i, c_lc);
            }
            for (int k = 0; k < 3; k++) variables.push_back(cs.new_witness_variable([&] { return a; }));
        }
    }
};

int main(int argc, char** argv) {
    const int log_n = argc > 1 ? atoi(argv[1]) : 20;
    auto rng = ark_std::test_rng();
    BenchCircuit circuit;
    circuit.a = ark_std::rand<F>(rng); circuit.b = ark_std::rand<F>(rng); circuit.c = ark_std::rand<F>(rng);
    circuit.num_constraints = size_t(1) << log_n;
    auto cs = ConstraintSystemRef<F>::new_ref();
    cs.set_optimization_goal(OptimizationGoal::Constraints);
    cs.set_mode(SynthesisMode::Prove(true, false));
    auto t0 = std::chrono::steady_clock::now();
    circuit.generate_constraints(cs);
    auto tg = std::chrono::steady_clock::now();
    cs.finalize();
    auto t1 = std::chrono::steady_clock::now();
    printf("  generate_constraints %.3f s, finalize %.3f s\n", std::chrono::duration<double>(tg - t0).count(), std::chrono::duration<double>(t1 - tg).count());
    printf("Synthesizing 2^%d constraints took %.3f s (%zu LCs, %zu terms, %zu pooled coefficients)\n", log_n,
           std::chrono::duration<double>(t1 - t0).count(), cs->lc_map().num_lcs(), cs->lc_map().total_lc_size(), cs->field_interner().vec().size());
    auto m = cs.to_matrices();
    auto t2 = std::chrono::steady_clock::now();
    size_t nnz = 0;
    for (const auto& mat : m.at(R1CS_PREDICATE_LABEL)) for (const auto& row : mat) nnz += row.size();
    printf("to_matrices(): %.3f s, %zu nonzeros (40 B each on the host); the LcMap arrays b2s_r1cs_upload_lcmap takes are %zu MiB\n",
           std::chrono::duration<double>(t2 - t1).count(), nnz,
           (cs->lc_map().total_lc_size() * 12 + cs->lc_map().num_lcs() * 8 + 3 * circuit.num_constraints * 8) >> 20);
    return 0;
}
