// Tests of the C++ host mirror (snark_b200/host/*.hpp), written after the reference's own unit tests
// (/root/reference/relations/src/gr1cs/tests/mod.rs, circuit2.rs, sr1cs/mod.rs:276-330, variable.rs:206-266).
//   ./host_relations_test cpu            -> builder tests, no GPU
//   ./host_relations_test gpu <curve> <circuit> tau alpha beta gamma delta r s   -> setup + prove through the C ABI, prints the proof
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../snark_b200/host/ark_snark.hpp"

using namespace ark_relations::gr1cs;

template <class F>
struct Circuit2 : ConstraintSynthesizer<F> {   // gr1cs/tests/circuit2.rs:47-60
    F a, b, c;
    Circuit2(F a_, F b_, F c_) : a(a_), b(b_), c(c_) {}
    void generate_constraints(ConstraintSystemRef<F> cs) override {
        const F two = F::one() + F::one();
        Variable va = cs.new_input_variable([&] { return a; });
        Variable vb = cs.new_witness_variable([&] { return b; });
        Variable vc = cs.new_witness_variable([&] { return c; });
        cs.enforce_r1cs_constraint([&] { return lc<F>() + va; }, [&] { return lc<F>() + std::make_pair(two, vb); }, [&] { return lc<F>() + vc; });
        Variable d = cs.new_lc([&] { return lc<F>() + va + vb; });
        cs.enforce_r1cs_constraint([&] { return lc<F>() + va; }, [&] { return lc<F>() + d; }, [&] { return lc<F>() + d; });
        Variable e = cs.new_lc([&] { return lc<F>() + d + d; });
        cs.enforce_r1cs_constraint([&] { return lc<F>() + Variable::One(); }, [&] { return lc<F>() + e; }, [&] { return lc<F>() + e; });
    }
};

template <class F>
struct DummyCircuit : ConstraintSynthesizer<F> {   // sr1cs/mod.rs:276-319
    F a, b;
    size_t num_variables, num_constraints;
    DummyCircuit(F a_, F b_, size_t nv, size_t nc) : a(a_), b(b_), num_variables(nv), num_constraints(nc) {}
    void generate_constraints(ConstraintSystemRef<F> cs) override {
        Variable va = cs.new_witness_variable([&] { return a; });
        Variable vb = cs.new_witness_variable([&] { return b; });
        Variable vc = cs.new_input_variable([&] { return a * b; });
        for (size_t i = 0; i < num_variables - 3; i++) cs.new_witness_variable([&] { return a; });
        for (size_t i = 0; i + 1 < num_constraints; i++)
            cs.enforce_r1cs_constraint([&] { return lc<F>({va}); }, [&] { return lc<F>({vb}); }, [&] { return lc<F>({vc}); });
        cs.enforce_r1cs_constraint([&] { return lc<F>(); }, [&] { return lc<F>(); }, [&] { return lc<F>(); });
    }
};

static int failures = 0;
#define CHECK(cond)                                                      \
    do {                                                                 \
        if (!(cond)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); failures++; } \
    } while (0)

template <class Curve>
static void cpu_tests(const char* name) {
    using F = typename Curve::Fr;
    auto U = [](uint64_t x) { return ark_snark::Groth16<Curve>::from_u64(x); };
    const F one = F::one(), two = U(2);
    {   // test_circuit2_matrices (tests/mod.rs:136-147): golden A, B, C of circuit2.rs:21-43 after finalize
        Circuit2<F> c(one, one, two);
        auto cs = ConstraintSystemRef<F>::new_ref();
        c.generate_constraints(cs);
        cs.finalize();
        using Row = std::vector<std::pair<F, size_t>>;
        std::vector<Matrix<F>> golden = {
            {Row{{one, 1}}, Row{{one, 1}}, Row{{one, 0}}},
            {Row{{two, 2}}, Row{{one, 1}, {one, 2}}, Row{{two, 1}, {two, 2}}},
            {Row{{one, 3}}, Row{{one, 1}, {one, 2}}, Row{{two, 1}, {two, 2}}},
        };
        CHECK(cs.to_matrices() == golden);
        CHECK(cs.is_satisfied());
        CHECK(cs.num_constraints() == 3 && cs.num_instance_variables() == 2 && cs.num_witness_variables() == 2);
        std::vector<F> z = cs->instance_assignment();
        z.insert(z.end(), cs->witness_assignment().begin(), cs->witness_assignment().end());
        auto az = mat_vec_mul(golden[0], z), bz = mat_vec_mul(golden[1], z), cz = mat_vec_mul(golden[2], z);
        for (int i = 0; i < 3; i++) CHECK(az[i] * bz[i] == cz[i]);
        auto t = transpose(golden[1], 4);
        CHECK(t[1].size() == 2 && t[2].size() == 3 && t[0].empty());
    }
    {   // unsatisfied witness is reported at the first failing constraint
        Circuit2<F> c(one, one, U(3));
        auto cs = ConstraintSystemRef<F>::new_ref();
        c.generate_constraints(cs);
        CHECK(!cs.is_satisfied());
        CHECK(cs->which_is_unsatisfied().value() == 0);
    }
    {   // r1cs_to_sr1cs test's DummyCircuit{128,128} synthesizes (sr1cs/mod.rs:320-330); shape checks
        DummyCircuit<F> c(U(3), U(5), 128, 128);
        auto cs = ConstraintSystemRef<F>::new_ref();
        c.generate_constraints(cs);
        CHECK(cs.num_constraints() == 128 && cs.num_instance_variables() == 2 && cs.num_witness_variables() == 127);
        CHECK(cs.is_satisfied());
        auto m = cs.to_matrices();
        CHECK(m[0][0].size() == 1 && m[0][0][0].second == 2 && m[1][0][0].second == 3 && m[2][0][0].second == 1);
        CHECK(m[0][127].empty() && m[1][127].empty() && m[2][127].empty());
    }
    {   // setup mode: closures are not evaluated, assignments are missing (constraint_system.rs:193-206, 598, 613)
        auto cs = ConstraintSystemRef<F>::new_ref();
        cs.set_mode(SynthesisMode::Setup());
        bool called = false;
        cs.new_witness_variable([&] { called = true; return one; });
        CHECK(!called);
        bool threw = false;
        try { cs->witness_assignment(); } catch (const SynthesisFailure& e) { threw = e.kind == SynthesisError::AssignmentMissing; }
        CHECK(threw);
        threw = false;
        try { ConstraintSystemRef<F>().num_constraints(); } catch (const SynthesisFailure& e) { threw = e.kind == SynthesisError::MissingCS; }
        CHECK(threw);
    }
    {   // test_variable_ordering (variable.rs:206-266) and column mapping (variable.rs:105-113)
        CHECK(Variable::Zero() < Variable::One());
        CHECK(Variable::One() < Variable::instance(0));
        CHECK(Variable::instance(7) < Variable::witness(0));
        CHECK(Variable::witness(9) < Variable::symbolic_lc(0));
        CHECK(Variable::instance(1) < Variable::instance(2));
        CHECK(*Variable::One().get_variable_index(5) == 0 && *Variable::instance(3).get_variable_index(5) == 3);
        CHECK(*Variable::witness(2).get_variable_index(5) == 7 && !Variable::symbolic_lc(1).get_variable_index(5));
    }
    {   // `lc + var` on a short LC inserts a duplicate (linear_combination.rs:174-190); compactify merges
        Variable v = Variable::witness(0);
        auto l = lc<F>() + v + v;
        CHECK(l.len() == 2);
        l.compactify();
        CHECK(l.len() == 1 && l.terms[0].first == two);
    }
    printf("%s cpu tests: %s\n", name, failures ? "FAILED" : "ok");
}

static void print_words(const char* tag, const std::vector<uint32_t>& w) {
    printf("%s", tag);
    for (uint32_t x : w) printf(" %08x", x);
    printf("\n");
}

template <class Curve>
static int gpu_prove(const char* circuit, char** a) {
    using F = typename Curve::Fr;
    using G = ark_snark::Groth16<Curve>;
    auto U = [&](const char* s) { return G::from_u64(strtoull(s, nullptr, 10)); };
    ark_snark::Trapdoor<Curve> td{U(a[0]), U(a[1]), U(a[2]), U(a[3]), U(a[4])};
    F r = U(a[5]), s = U(a[6]);
    G g(0);
    Circuit2<F> c2(F::one(), F::one(), G::from_u64(2));
    DummyCircuit<F> dc(G::from_u64(3), G::from_u64(5), 16, 16);
    ConstraintSynthesizer<F>& circ = strcmp(circuit, "circuit2") == 0 ? static_cast<ConstraintSynthesizer<F>&>(c2)
                                                                        : static_cast<ConstraintSynthesizer<F>&>(dc);
    auto pk = g.circuit_specific_setup(circ, td);
    auto proof = g.prove(pk, circ, r, s);
    print_words("A", proof.a);
    print_words("B", proof.b);
    print_words("C", proof.c);
    print_words("alpha_g1", pk.alpha_g1);
    print_words("h_query0", std::vector<uint32_t>(pk.h_query.begin(), pk.h_query.begin() + 2 * Curve::Fq::N));
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 2 && strcmp(argv[1], "cpu") == 0) {
        cpu_tests<b2s::Bls12_381>("bls12_381");
        cpu_tests<b2s::Bn254>("bn254");
        return failures ? 1 : 0;
    }
    if (argc == 11 && strcmp(argv[1], "gpu") == 0) {
        try {
            return atoi(argv[2]) == 0 ? gpu_prove<b2s::Bls12_381>(argv[3], argv + 4) : gpu_prove<b2s::Bn254>(argv[3], argv + 4);
        } catch (const std::exception& e) {
            printf("ERROR %s\n", e.what());
            return 2;
        }
    }
    printf("usage: %s cpu | gpu <curve 0|1> <circuit2|dummy> tau alpha beta gamma delta r s\n", argv[0]);
    return 64;
}
