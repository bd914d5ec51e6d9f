// Tests of the C++ host mirror (snark_b200/host/*.hpp), written after the reference's own unit tests
// (/root/reference/relations/src/gr1cs/tests/mod.rs, circuit1.rs, circuit2.rs, sr1cs/mod.rs:276-330, variable.rs:206-266).
//   ./host_relations_test cpu            -> builder tests, no GPU
//   ./host_relations_test rng            -> ark_std::test_rng() words and Fr::rand draws (compared with oracle/rng.py)
//   ./host_relations_test gpu <curve> <circuit> tau alpha beta gamma delta r s   -> setup + prove through the C ABI, prints the proof
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>

#include "../../snark_b200/host/ark_snark.hpp"

using namespace ark_relations::gr1cs;
using ark_relations::sr1cs::Sr1csAdapter;

template <class F>
struct Circuit1 : ConstraintSynthesizer<F> {   // gr1cs/tests/circuit1.rs:10-24, 63-164
    F x[5], w[8];
    Circuit1(const F (&x_)[5], const F (&w_)[8]) { for (int i = 0; i < 5; i++) x[i] = x_[i]; for (int i = 0; i < 8; i++) w[i] = w_[i]; }
    void generate_constraints(ConstraintSystemRef<F> cs) override {
        auto input_ns = ns(cs, "Input variables");
        Variable xv[5], wv[8];
        for (int i = 0; i < 5; i++) xv[i] = input_ns.cs().new_input_variable([&, i] { return x[i]; });
        ns(cs, "Witness variables");
        for (int i = 0; i < 8; i++) wv[i] = cs.new_witness_variable([&, i] { return w[i]; });
        const F one = F::one(), minus_one = F::zero() - one, three = one + one + one, seven = three + three + one;
        cs.register_predicate("poly-predicate-A", PredicateConstraintSystem<F>::new_polynomial_predicate_cs(
                                                      4, {{one, {{0, 1}, {1, 1}}}, {three, {{2, 2}}}, {minus_one, {{3, 1}}}}));
        cs.register_predicate("poly-predicate-B", PredicateConstraintSystem<F>::new_polynomial_predicate_cs(
                                                      3, {{seven, {{1, 1}}}, {one, {{0, 3}}}, {minus_one, {{2, 1}}}}));
        cs.register_predicate("poly-predicate-C", PredicateConstraintSystem<F>::new_polynomial_predicate_cs(
                                                      3, {{one, {{0, 1}, {1, 1}}}, {minus_one, {{2, 1}}}}));
        auto V = [](Variable v) { return [v] { return lc<F>() + v; }; };
        const Variable x1 = xv[0], x2 = xv[1], x3 = xv[2], x4 = xv[3], x5 = xv[4];
        const Variable w1 = wv[0], w2 = wv[1], w3 = wv[2], w4 = wv[3], w5 = wv[4], w6 = wv[5], w8 = wv[7];
        ns(cs, "Predicate A constraints");
        cs.enforce_constraint_arity_4("poly-predicate-A", V(x1), V(x2), V(x3), V(w4));
        ns(cs, "Predicate B constraints");
        cs.enforce_constraint_arity_3("poly-predicate-B", V(x4), V(w1), V(w5));
        cs.enforce_constraint_arity_3("poly-predicate-B", V(w5), V(w6), V(w8));
        ns(cs, "Predicate C constraints");
        cs.enforce_constraint_arity_3("poly-predicate-C", V(w2), V(w3), V(w6));
        cs.enforce_constraint_arity_3("poly-predicate-C", [=] { return lc<F>() + w5 + w4; }, V(w8), V(x5));
    }
    static std::map<Label, std::vector<Matrix<F>>> get_matrices() {   // circuit1.rs:28-61
        using Row = std::vector<std::pair<F, size_t>>;
        const F one = F::one();
        auto R = [&](size_t col) { return Row{{one, col}}; };
        std::map<Label, std::vector<Matrix<F>>> m;
        m[R1CS_PREDICATE_LABEL] = {{}, {}, {}};
        m["poly-predicate-A"] = {{R(1)}, {R(2)}, {R(3)}, {R(9)}};
        m["poly-predicate-B"] = {{R(4), R(10)}, {R(6), R(11)}, {R(10), R(13)}};
        m["poly-predicate-C"] = {{R(7), Row{{one, 9}, {one, 10}}}, {R(8), R(13)}, {R(11), R(5)}};
        return m;
    }
};

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
        CHECK(cs.to_matrices().at(R1CS_PREDICATE_LABEL) == golden);
        CHECK(cs.to_matrices().size() == 1 && cs.num_predicates() == 1);
        CHECK(cs.is_satisfied());
        CHECK(cs.num_constraints() == 3 && cs.num_instance_variables() == 2 && cs.num_witness_variables() == 2);
        std::vector<F> z = cs->instance_assignment();
        z.insert(z.end(), cs->witness_assignment().begin(), cs->witness_assignment().end());
        auto az = mat_vec_mul(golden[0], z), bz = mat_vec_mul(golden[1], z), cz = mat_vec_mul(golden[2], z);
        for (int i = 0; i < 3; i++) CHECK(az[i] * bz[i] == cz[i]);
        auto t = transpose(golden[1], 4);
        CHECK(t[1].size() == 2 && t[2].size() == 3 && t[0].empty());
    }
    {   // Setup mode builds the same matrices without evaluating a single assignment closure (what circuit_specific_setup uploads)
        Circuit2<F> c(one, one, two);
        auto prove_cs = ConstraintSystemRef<F>::new_ref();
        prove_cs.set_optimization_goal(OptimizationGoal::Constraints);
        c.generate_constraints(prove_cs);
        prove_cs.finalize();
        auto setup_cs = ConstraintSystemRef<F>::new_ref();
        setup_cs.set_optimization_goal(OptimizationGoal::Constraints);
        setup_cs.set_mode(SynthesisMode::Setup());
        c.generate_constraints(setup_cs);
        setup_cs.finalize();
        CHECK(setup_cs.to_matrices() == prove_cs.to_matrices());
        CHECK(setup_cs.num_witness_variables() == 2 && setup_cs.num_instance_variables() == 2 && setup_cs.num_constraints() == 3);
        CHECK(setup_cs->lc_map().offsets() == prove_cs->lc_map().offsets() && setup_cs->lc_map().vars() == prove_cs->lc_map().vars());
        DummyCircuit<F> d(U(3), U(5), 16, 16);
        auto ds = ConstraintSystemRef<F>::new_ref();
        ds.set_mode(SynthesisMode::Setup());
        d.generate_constraints(ds);
        ds.finalize();
        auto dp = ConstraintSystemRef<F>::new_ref();
        d.generate_constraints(dp);
        dp.finalize();
        CHECK(ds.to_matrices() == dp.to_matrices());
    }
    {   // unsatisfied witness is reported at the first failing constraint
        Circuit2<F> c(one, one, U(3));
        auto cs = ConstraintSystemRef<F>::new_ref();
        c.generate_constraints(cs);
        CHECK(!cs.is_satisfied());
        CHECK(cs.which_is_unsatisfied().value() == "R1CS - 0");
    }
    {   // r1cs_to_sr1cs test's DummyCircuit{128,128} synthesizes (sr1cs/mod.rs:320-330); shape checks
        DummyCircuit<F> c(U(3), U(5), 128, 128);
        auto cs = ConstraintSystemRef<F>::new_ref();
        c.generate_constraints(cs);
        CHECK(cs.num_constraints() == 128 && cs.num_instance_variables() == 2 && cs.num_witness_variables() == 127);
        CHECK(cs.is_satisfied());
        auto m = cs.to_matrices().at(R1CS_PREDICATE_LABEL);
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
    auto U32 = [&](uint64_t v) { return U(v); };
    const F sat_x[5] = {U32(1), U32(2), U32(3), U32(0), U32(1255254)};
    const F sat_w[8] = {U32(4), U32(2), U32(5), U32(29), U32(28), U32(10), U32(57), U32(22022)};
    const F zeros5[5] = {F::zero(), F::zero(), F::zero(), F::zero(), F::zero()};
    const F zeros8[8] = {F::zero(), F::zero(), F::zero(), F::zero(), F::zero(), F::zero(), F::zero(), F::zero()};
    {   // test_circuit1_sat (tests/mod.rs:17-45): satisfied with and without the Constraints optimisation goal
        Circuit1<F> c(sat_x, sat_w);
        auto cs = ConstraintSystemRef<F>::new_ref();
        c.generate_constraints(cs);
        cs.finalize();
        CHECK(cs.is_satisfied());
        auto cs2 = ConstraintSystemRef<F>::new_ref();
        cs2.set_optimization_goal(OptimizationGoal::Constraints);
        c.generate_constraints(cs2);
        cs2.finalize();
        CHECK(cs2.is_satisfied());
        CHECK(cs.num_constraints() == 5 && cs.num_predicates() == 4 && cs.num_instance_variables() == 6 && cs.num_witness_variables() == 8);
        CHECK(cs->get_predicate_arity("poly-predicate-A").value() == 4 && cs->get_predicate_num_constraints("poly-predicate-B").value() == 2);
        CHECK(!cs->get_predicate_arity("nope").has_value());
        CHECK(cs->predicates().at("poly-predicate-B").get_predicate().degree() == 3);
    }
    {   // test_circuit1_non_sat (tests/mod.rs:47-76): x1 = 4 breaks predicate A; reported without finalize
        F bad_x[5] = {U32(4), sat_x[1], sat_x[2], sat_x[3], sat_x[4]};
        Circuit1<F> c(bad_x, sat_w);
        auto cs = ConstraintSystemRef<F>::new_ref();
        c.generate_constraints(cs);
        CHECK(!cs.is_satisfied());
        CHECK(cs.which_is_unsatisfied().value() == "poly-predicate-A - 0");
    }
    {   // test_circuit1_matrices (tests/mod.rs:78-103): golden map of circuit1.rs:28-61, BEFORE finalize
        Circuit1<F> c(zeros5, zeros8);
        auto cs = ConstraintSystemRef<F>::new_ref();
        c.generate_constraints(cs);
        CHECK(Circuit1<F>::get_matrices() == cs.to_matrices());
        cs.set_instance_outliner({R1CS_PREDICATE_LABEL, outline_r1cs<F>});
        cs.finalize();
    }
    {   // test_circuit1_instance_outlined (tests/mod.rs:105-131)
        Circuit1<F> c(zeros5, zeros8);
        auto cs = ConstraintSystemRef<F>::new_ref();
        c.generate_constraints(cs);
        const size_t num_instance = cs.num_instance_variables(), prev_num_witness = cs.num_witness_variables();
        cs.set_instance_outliner({R1CS_PREDICATE_LABEL, outline_r1cs<F>});
        CHECK(cs.should_outline_instances());
        cs.finalize();
        CHECK(num_instance == cs.num_witness_variables() - prev_num_witness);
        CHECK(!cs.should_outline_instances());
        CHECK(cs->get_predicate_num_constraints(R1CS_PREDICATE_LABEL).value() == num_instance);   // one tie per instance column
    }
    {   // what outlining does to circuit2: stored LCs lose their instance columns, two tie rows, still satisfied
        Circuit2<F> c(one, one, two);
        auto cs = ConstraintSystemRef<F>::new_ref();
        c.generate_constraints(cs);
        cs.set_instance_outliner({R1CS_PREDICATE_LABEL, outline_r1cs<F>});
        cs.finalize();
        using Row = std::vector<std::pair<F, size_t>>;
        auto m = cs.to_matrices().at(R1CS_PREDICATE_LABEL);
        CHECK(cs.num_witness_variables() == 4 && cs.num_constraints() == 5 && cs.is_satisfied());
        CHECK((m[1][1] == Row{{one, 5}, {one, 2}}) && (m[2][2] == Row{{two, 5}, {two, 2}}));   // rewritten in place: unsorted
        CHECK((m[0][3] == Row{{one, 4}}) && (m[1][3] == Row{{one, 4}}) && (m[2][3] == Row{{one, 0}}));
        CHECK((m[0][4] == Row{{one, 4}}) && (m[1][4] == Row{{one, 5}}) && (m[2][4] == Row{{one, 1}}));
        // an outliner naming an unregistered predicate is dropped (constraint_system.rs:701)
        auto cs2 = ConstraintSystemRef<F>::new_ref();
        c.generate_constraints(cs2);
        cs2.set_instance_outliner({"no-such-predicate", outline_r1cs<F>});
        cs2.finalize();
        CHECK(cs2.num_witness_variables() == 2 && !cs2.should_outline_instances());
    }
    {   // predicate errors: unknown label, arity quirk of predicate/mod.rs:156-174
        auto cs = ConstraintSystemRef<F>::new_ref();
        Variable v = cs.new_witness_variable([&] { return one; });
        auto L = [v] { return lc<F>() + v; };
        bool threw = false;
        try { cs.enforce_constraint_arity_2("missing", L, L); } catch (const SynthesisFailure& e) { threw = e.kind == SynthesisError::PredicateNotFound; }
        CHECK(threw);
        threw = false;
        try { cs.enforce_constraint_arity_2(R1CS_PREDICATE_LABEL, L, L); } catch (const SynthesisFailure& e) { threw = e.kind == SynthesisError::ArityMismatch; }
        CHECK(threw && cs.num_constraints() == 0);
        CHECK(cs->predicates().at(R1CS_PREDICATE_LABEL).get_constraints()[0].size() == 1);   // the partial push stays, as upstream
        auto cs2 = ConstraintSystemRef<F>::new_ref();
        Variable u = cs2.new_witness_variable([&] { return one; });
        auto M = [u] { return lc<F>() + u; };
        cs2.enforce_constraint_arity_4(R1CS_PREDICATE_LABEL, M, M, M, M);   // surplus argument dropped silently
        CHECK(cs2.num_constraints() == 1 && cs2.is_satisfied());
        cs2.remove_predicate(R1CS_PREDICATE_LABEL);
        CHECK(!cs2.has_predicate(R1CS_PREDICATE_LABEL) && cs2.num_predicates() == 0 && cs2.num_constraints() == 0);
        CHECK(!ConstraintSystemRef<F>().has_predicate(R1CS_PREDICATE_LABEL));
        ConstraintSystemRef<F>().finalize();   // None: no-op
    }
    {   // LC (+|-) LC merges, negation, lc_diff (linear_combination.rs:32-38, 300-470)
        Variable a = Variable::instance(1), b = Variable::witness(0), c = Variable::witness(3);
        auto l1 = lc_pairs<F>({{two, a}, {one, c}}), l2 = lc_pairs<F>({{one, a}, {two, b}});
        using T = std::vector<std::pair<F, Variable>>;
        const F three = two + one, minus_one = F::zero() - one, minus_two = F::zero() - two;
        CHECK(((l1 + l2).terms == T{{three, a}, {two, b}, {one, c}}));
        CHECK(((l1 - l2).terms == T{{one, a}, {minus_two, b}, {one, c}}));
        CHECK(((lc<F>() - l2).terms == T{{minus_one, a}, {minus_two, b}}) && (l1 + lc<F>()).terms == l1.terms);
        CHECK((lc_diff<F>(a, b).terms == T{{one, a}, {minus_one, b}}) && lc_diff<F>(a, a).terms.empty());
        CHECK(LinearCombination<F>::from(Variable::Zero()).terms.empty() && LinearCombination<F>::from(F::zero(), a).terms.empty());
    }
    {   // test_lc_map_iter_mut (lc_map.rs:524-568) and the interner's conventions (field_interner.rs:27-68)
        FieldInterner<F> interner;
        LcMap<F> lcmap;
        lcmap.push({{U(1), Variable::One()}, {U(2), Variable::instance(2)}}, interner);
        lcmap.push({{U(3), Variable::witness(4)}, {U(4), Variable::instance(4)}}, interner);
        lcmap.lc_vars_iter_mut([](Variable* v, Variable* end) {
            for (; v != end; ++v) if (v->is_instance()) *v = Variable::instance(*v->index() + 1);
        });
        std::vector<std::pair<F, Variable>> flattened;
        for (size_t i = 0; i < lcmap.num_lcs(); i++) for (const auto& t : lcmap.get_lc(i, interner)) flattened.push_back(t);
        std::vector<std::pair<F, Variable>> expected = {{U(1), Variable::One()}, {U(2), Variable::instance(3)},
                                                        {U(3), Variable::witness(4)}, {U(4), Variable::instance(5)}};
        CHECK(flattened == expected);
        CHECK(lcmap.num_lcs() == 2 && lcmap.total_lc_size() == 4 && (lcmap.offsets() == std::vector<uint64_t>{0, 2, 4}));
        CHECK(!lcmap.get(2).has_value() && lcmap.get(1)->size() == 2);
        const F minus_one = F::zero() - one;
        CHECK(interner.vec()[0] == one && interner.vec()[1] == minus_one);
        CHECK(interner.get_or_intern(one).id == 0 && interner.get_or_intern(minus_one).id == 1);
        const uint32_t id7 = interner.get_or_intern(U(7)).id;
        CHECK(id7 == interner.vec().size() - 1 && interner.get_or_intern(U(7)).id == id7 && *interner.value({id7}) == U(7));
        CHECK(lcmap.coeffs()[0].id == 0 && !interner.value({1000}).has_value());
    }
    {   // proving key container: round trip, and rejection of a wrong curve / truncated / inconsistent file
        ark_snark::ProvingKey<Curve> pk;
        pk.n_instance = 2; pk.n_witness = 3; pk.domain_size = 8;
        const size_t g1 = 2 * Curve::Fq::N, g2 = 4 * Curve::Fq::N;
        uint32_t ctr = 1;
        auto fill = [&](std::vector<uint32_t>& v, size_t words) { v.resize(words); for (auto& w : v) w = ctr++ * 2654435761u; };
        fill(pk.alpha_g1, g1); fill(pk.beta_g1, g1); fill(pk.delta_g1, g1); fill(pk.beta_g2, g2); fill(pk.delta_g2, g2); fill(pk.gamma_g2, g2);
        fill(pk.gamma_abc_g1, 2 * g1); fill(pk.a_query, 5 * g1); fill(pk.b_g1_query, 5 * g1); fill(pk.b_g2_query, 5 * g2);
        fill(pk.h_query, 7 * g1); fill(pk.l_query, 3 * g1);
        std::stringstream file;
        ark_snark::write_proving_key(file, pk);
        const std::string bytes = file.str();
        auto back = ark_snark::read_proving_key<Curve>(file);
        CHECK(back.n_instance == 2 && back.n_witness == 3 && back.domain_size == 8 && back.h_query == pk.h_query && back.b_g2_query == pk.b_g2_query);
        CHECK(back.alpha_g1 == pk.alpha_g1 && back.gamma_abc_g1 == pk.gamma_abc_g1 && back.l_query == pk.l_query && back.gamma_g2 == pk.gamma_g2);
        auto rejects = [&](std::string b) { std::stringstream f(b); try { ark_snark::read_proving_key<Curve>(f); } catch (const std::runtime_error&) { return true; } return false; };
        CHECK(rejects(bytes.substr(0, bytes.size() - 5)));                    // truncated
        std::string other = bytes; other[8] ^= 1;                             // curve id flipped
        CHECK(rejects(other));
        std::string magic = bytes; magic[0] ^= 0x40;
        CHECK(rejects(magic));
        std::string counts = bytes; counts[24] ^= 1;                          // n_witness 3 -> 2: array lengths no longer match
        CHECK(rejects(counts));
        CHECK(!rejects(bytes));
    }
    {   // Sr1csAdapter (sr1cs/mod.rs:122-264): a*b = c  ->  (a+b)^2 = 4c + s, (a-b)^2 = s; instances re-exposed
        DummyCircuit<F> c(U(3), U(5), 8, 8);
        auto cs = ConstraintSystemRef<F>::new_ref();
        c.generate_constraints(cs);
        cs.finalize();
        auto shape = Sr1csAdapter<F>::r1cs_to_sr1cs(cs);
        // 8 rows -> 16 square constraints + 1 tie for the one public column in use; witnesses: a, b, c copies + 8 squares
        CHECK(shape.is_in_setup_mode() && shape.num_constraints() == 17 && shape.num_instance_variables() == 2 && shape.num_witness_variables() == 11);
        CHECK(!shape.has_predicate(R1CS_PREDICATE_LABEL) && shape.has_predicate(SR1CS_PREDICATE_LABEL));
        auto full = Sr1csAdapter<F>::r1cs_to_sr1cs_with_assignment(cs.inner());
        CHECK(full.num_constraints() == 17 && full.is_satisfied());
        CHECK(full->instance_assignment()[1] == U(15));                    // the new instance carries c = a*b
        auto rows = full.to_matrices().at(SR1CS_PREDICATE_LABEL);
        std::vector<F> z = full->instance_assignment();
        z.insert(z.end(), full->witness_assignment().begin(), full->witness_assignment().end());
        for (size_t i = 0; i < rows[0].size(); i++) {
            F l = Sr1csAdapter<F>::evaluate_constraint(rows[0][i], z), r = Sr1csAdapter<F>::evaluate_constraint(rows[1][i], z);
            CHECK(l * l == r);
        }
        CHECK(mat_vec_mul(rows[0], z)[0] == U(8));                         // first row: a + b = 3 + 5
        // a violated R1CS row gives a violated SR1CS system
        Circuit2<F> bad(one, one, U(3));
        auto bcs = ConstraintSystemRef<F>::new_ref();
        bad.generate_constraints(bcs);
        bcs.finalize();
        CHECK(!Sr1csAdapter<F>::r1cs_to_sr1cs_with_assignment(bcs.inner()).is_satisfied());
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

template <class Curve>
static void rng_draws(const char* tag) {
    auto rng = ark_std::test_rng();
    for (int k = 0; k < 20; k++) {
        auto x = ark_std::rand<typename Curve::Fr>(rng);
        print_words(tag, std::vector<uint32_t>(x.v, x.v + Curve::Fr::N));
    }
}

// SNARK::circuit_specific_setup(circuit, rng) then SNARK::prove(pk, circuit, rng) on one test_rng() stream
template <class Curve>
static int gpu_prove_rng(const char* circuit) {
    using F = typename Curve::Fr;
    using G = ark_snark::Groth16<Curve>;
    G g(0);
    Circuit2<F> c2(F::one(), F::one(), G::from_u64(2));
    DummyCircuit<F> dc(G::from_u64(3), G::from_u64(5), 16, 16);
    ConstraintSynthesizer<F>& circ = strcmp(circuit, "circuit2") == 0 ? static_cast<ConstraintSynthesizer<F>&>(c2)
                                                                        : static_cast<ConstraintSynthesizer<F>&>(dc);
    auto rng = ark_std::test_rng();
    auto pk = g.circuit_specific_setup(circ, rng);
    auto proof = g.prove(pk, circ, rng);
    print_words("A", proof.a);
    print_words("B", proof.b);
    print_words("C", proof.c);
    print_words("alpha_g1", pk.alpha_g1);
    return 0;
}

// the flat storage of a finalized system, one array per line (compared with oracle/r1cs.py: to_lcmap)
template <class Curve>
static void dump_lcmap(const char* tag, bool outlined) {
    using F = typename Curve::Fr;
    Circuit2<F> c(F::one(), F::one(), ark_snark::Groth16<Curve>::from_u64(2));
    auto cs = ConstraintSystemRef<F>::new_ref();
    c.generate_constraints(cs);
    if (outlined) cs.set_instance_outliner({R1CS_PREDICATE_LABEL, outline_r1cs<F>});
    cs.finalize();
    const auto& lm = cs->lc_map();
    printf("%s offsets", tag);
    for (uint64_t o : lm.offsets()) printf(" %llu", (unsigned long long)o);
    printf("\n%s vars", tag);
    for (const Variable& v : lm.vars()) printf(" %llu", (unsigned long long)v.raw);
    printf("\n%s coeffs", tag);
    for (const InternedField& i : lm.coeffs()) printf(" %u", i.id);
    printf("\n");
    for (const F& v : cs->field_interner().vec()) print_words((std::string(tag) + " pool").c_str(), std::vector<uint32_t>(v.v, v.v + F::N));
    const auto& args = cs->predicates().at(R1CS_PREDICATE_LABEL).get_constraints();
    for (int k = 0; k < 3; k++) {
        printf("%s args%d", tag, k);
        for (const Variable& v : args[k]) printf(" %llu", (unsigned long long)v.raw);
        printf("\n");
    }
}

int main(int argc, char** argv) {
    if (argc >= 2 && strcmp(argv[1], "lcmap") == 0) {
        dump_lcmap<b2s::Bls12_381>("bls12_381", false);
        dump_lcmap<b2s::Bls12_381>("bls12_381-outlined", true);
        dump_lcmap<b2s::Bn254>("bn254", false);
        return 0;
    }
    if (argc >= 2 && strcmp(argv[1], "rng") == 0) {
        auto rng = ark_std::test_rng();
        std::vector<uint32_t> w;
        for (int i = 0; i < 63; i++) w.push_back(rng.next_u32());
        const uint64_t straddle = rng.next_u64();          // one word left: low = word 63, high = word 64
        w.push_back(uint32_t(straddle)); w.push_back(uint32_t(straddle >> 32));
        for (int i = 0; i < 3; i++) { const uint64_t v = rng.next_u64(); w.push_back(uint32_t(v)); w.push_back(uint32_t(v >> 32)); }
        print_words("words", w);
        uint8_t bytes[10];
        ark_std::test_rng().fill_bytes(bytes, sizeof(bytes));
        printf("bytes");
        for (uint8_t b : bytes) printf(" %02x", b);
        printf("\n");
        rng_draws<b2s::Bls12_381>("bls12_381");
        rng_draws<b2s::Bn254>("bn254");
        return 0;
    }
    if (argc >= 2 && strcmp(argv[1], "cpu") == 0) {
        cpu_tests<b2s::Bls12_381>("bls12_381");
        cpu_tests<b2s::Bn254>("bn254");
        return failures ? 1 : 0;
    }
    if (argc == 4 && strcmp(argv[1], "gpu-rng") == 0) {   // setup + prove with the reference's (circuit, rng) signatures
        try {
            return atoi(argv[2]) == 0 ? gpu_prove_rng<b2s::Bls12_381>(argv[3]) : gpu_prove_rng<b2s::Bn254>(argv[3]);
        } catch (const std::exception& e) {
            printf("ERROR %s\n", e.what());
            return 2;
        }
    }
    if (argc == 11 && strcmp(argv[1], "gpu") == 0) {
        try {
            return atoi(argv[2]) == 0 ? gpu_prove<b2s::Bls12_381>(argv[3], argv + 4) : gpu_prove<b2s::Bn254>(argv[3], argv + 4);
        } catch (const std::exception& e) {
            printf("ERROR %s\n", e.what());
            return 2;
        }
    }
    printf("usage: %s cpu | rng | lcmap | gpu-rng <curve 0|1> <circuit2|dummy> | gpu <curve 0|1> <circuit2|dummy> tau alpha beta gamma delta r s\n", argv[0]);
    return 64;
}
