// C++ mirror of the ark-relations GR1CS builder API -- the host side that sits ABOVE the C ABI.
//
// The reference is Rust and keeps constraint synthesis on the host ("Rust host code owns constraint
// synthesis and witness assignment exactly as the reference does"); this image has no Rust toolchain, so
// the same interface is provided in C++ with the reference's names, argument meaning and error behaviour,
// for hosts that are C++ and so that tests/native/host_relations_test.cpp can read like the reference's own
// tests (relations/src/gr1cs/tests/mod.rs).  R1CS predicate only (what Groth16 consumes).
//
//   Variable                          relations/src/utils/variable.rs:4-14,105-113,177-183
//   LinearCombination, lc()           relations/src/utils/linear_combination.rs:15-38,53-82,174-211
//   SynthesisError                    relations/src/utils/error.rs:5-21
//   SynthesisMode / OptimizationGoal  relations/src/gr1cs/mod.rs:75-106
//   ConstraintSystem                  relations/src/gr1cs/constraint_system.rs:44-139,323-353,431-438,472-532,
//                                     591-617,652-707,717-804
//   ConstraintSystemRef               relations/src/gr1cs/constraint_system_ref.rs:26-34,235-250,345-383
//   ConstraintSynthesizer             relations/src/gr1cs/mod.rs:54-61
//   Matrix, mat_vec_mul, transpose    relations/src/utils/matrix.rs:4-36
//
// F is any field type with zero()/one()/+/-/*/==/is_zero() -- in this repository b2s::Fp<...> compiled for the
// host.  Coefficient arithmetic here is the builder's own (a handful of additions/multiplications per
// constraint), exactly as in the reference; the prover hot path never runs on the CPU.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <functional>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace ark_relations {
namespace gr1cs {

enum class SynthesisError {   // utils/error.rs:5-21
    MissingCS = 1, AssignmentMissing = 2, DivisionByZero = 3, Unsatisfiable = 4, PolynomialDegreeTooLarge = 5,
    UnexpectedIdentity = 6, MalformedVerifyingKey = 7, ArityMismatch = 8, PredicateNotFound = 9,
};
struct SynthesisFailure : std::runtime_error {
    SynthesisError kind;
    explicit SynthesisFailure(SynthesisError k) : std::runtime_error("SynthesisError"), kind(k) {}
};

enum class VarKind : uint8_t { Zero = 0, One = 1, Instance = 2, Witness = 3, SymbolicLc = 4 };

// 8-byte tagged id: top 3 bits = kind, low 61 bits = index (variable.rs:4-14).
struct Variable {
    uint64_t raw;
    static constexpr int TAG_SHIFT = 61;
    static constexpr uint64_t PAYLOAD_MASK = (uint64_t(1) << TAG_SHIFT) - 1;
    static Variable pack(VarKind k, uint64_t payload) { return {(uint64_t(k) << TAG_SHIFT) | (payload & PAYLOAD_MASK)}; }
    static Variable Zero() { return pack(VarKind::Zero, 0); }
    static Variable One() { return pack(VarKind::One, 0); }
    static Variable instance(size_t i) { return pack(VarKind::Instance, i); }
    static Variable witness(size_t i) { return pack(VarKind::Witness, i); }
    static Variable symbolic_lc(size_t i) { return pack(VarKind::SymbolicLc, i); }
    VarKind kind() const { return VarKind(raw >> TAG_SHIFT); }
    uint64_t payload() const { return raw & PAYLOAD_MASK; }
    bool is_zero() const { return kind() == VarKind::Zero; }
    bool is_one() const { return kind() == VarKind::One; }
    bool is_instance() const { return kind() == VarKind::Instance; }
    bool is_witness() const { return kind() == VarKind::Witness; }
    bool is_lc() const { return kind() == VarKind::SymbolicLc; }
    std::optional<size_t> index() const {   // variable.rs:120-127
        if (is_zero() || is_one()) return std::nullopt;
        return size_t(payload());
    }
    std::optional<size_t> get_lc_index() const { return is_lc() ? std::optional<size_t>(payload()) : std::nullopt; }
    // column of this variable in the constraint matrices (variable.rs:105-113)
    std::optional<size_t> get_variable_index(size_t witness_offset) const {
        switch (kind()) {
            case VarKind::One: return size_t(0);
            case VarKind::Instance: return size_t(payload());
            case VarKind::Witness: return size_t(payload()) + witness_offset;
            default: return std::nullopt;
        }
    }
    // Ord: Zero < One < Instance < Witness < SymbolicLc, then by index (variable.rs:206-266)
    bool operator<(const Variable& o) const { return raw < o.raw; }
    bool operator>=(const Variable& o) const { return raw >= o.raw; }
    bool operator==(const Variable& o) const { return raw == o.raw; }
    bool operator!=(const Variable& o) const { return raw != o.raw; }
};
static_assert(sizeof(Variable) == 8, "Variable is 8 bytes (variable.rs:197)");

template <class F>
struct LinearCombination {   // linear_combination.rs:15
    std::vector<std::pair<F, Variable>> terms;

    LinearCombination() = default;
    explicit LinearCombination(std::vector<std::pair<F, Variable>> t) : terms(std::move(t)) {}
    static LinearCombination zero() { return {}; }
    size_t len() const { return terms.size(); }

    // linear_combination.rs:53-82
    void compactify() {
        if (terms.size() <= 1) return;
        std::sort(terms.begin(), terms.end(), [](const auto& a, const auto& b) { return a.second < b.second; });
        size_t w = 0;
        for (size_t r = 1; r < terms.size(); r++) {
            if (terms[w].second == terms[r].second) terms[w].first = terms[w].first + terms[r].first;
            else terms[++w] = terms[r];
        }
        terms.resize(w + 1);
    }
    // linear_combination.rs:174-190: below 6 terms the linear scan never reports a hit (returns Err(idx))
    std::pair<bool, size_t> get_var_loc(const Variable& v) const {
        if (terms.size() < 6) {
            size_t found = 0;
            for (size_t i = 0; i < terms.size(); i++) {
                if (terms[i].second >= v) { found = i; break; }
                found++;
            }
            return {false, found};
        }
        size_t lo = 0, hi = terms.size();
        while (lo < hi) {
            size_t mid = (lo + hi) / 2;
            if (terms[mid].second < v) lo = mid + 1;
            else if (terms[mid].second == v) return {true, mid};
            else hi = mid;
        }
        return {false, lo};
    }
    // AddAssign<(F, Variable)>  linear_combination.rs:203-211
    LinearCombination& operator+=(const std::pair<F, Variable>& cv) {
        auto [hit, loc] = get_var_loc(cv.second);
        if (hit) terms[loc].first = terms[loc].first + cv.first;
        else terms.insert(terms.begin() + loc, cv);
        return *this;
    }
    LinearCombination operator+(const std::pair<F, Variable>& cv) const { LinearCombination r = *this; r += cv; return r; }
    LinearCombination operator+(const Variable& v) const { return *this + std::make_pair(F::one(), v); }
    LinearCombination operator-(const std::pair<F, Variable>& cv) const { return *this + std::make_pair(F::zero() - cv.first, cv.second); }
    LinearCombination operator-(const Variable& v) const { return *this - std::make_pair(F::one(), v); }
    LinearCombination operator*(const F& k) const { LinearCombination r = *this; for (auto& t : r.terms) t.first = t.first * k; return r; }
};

// lc!() forms (linear_combination.rs:19-30): lc<F>() empty; lc<F>({a, b}) = sum of variables; lc_pairs = sum of pairs
template <class F> LinearCombination<F> lc() { return {}; }
template <class F> LinearCombination<F> lc(std::initializer_list<Variable> vars) {
    LinearCombination<F> r;
    for (auto v : vars) r.terms.emplace_back(F::one(), v);
    r.compactify();
    return r;
}
template <class F> LinearCombination<F> lc_pairs(std::initializer_list<std::pair<F, Variable>> pairs) {
    LinearCombination<F> r;
    for (auto& p : pairs) r.terms.push_back(p);
    r.compactify();
    return r;
}

template <class F> using Matrix = std::vector<std::vector<std::pair<F, size_t>>>;   // utils/matrix.rs:4

template <class F>
std::vector<F> mat_vec_mul(const Matrix<F>& m, const std::vector<F>& v) {   // utils/matrix.rs:26-36
    std::vector<F> out;
    for (const auto& row : m) {
        F sum = F::zero();
        for (const auto& [val, col] : row) sum = sum + v[col] * val;
        out.push_back(sum);
    }
    return out;
}
template <class F>
Matrix<F> transpose(const Matrix<F>& m, size_t num_col) {   // utils/matrix.rs:8-23
    Matrix<F> t(num_col);
    for (size_t r = 0; r < m.size(); r++)
        for (const auto& [val, col] : m[r]) t[col].emplace_back(val, r);
    return t;
}

struct SynthesisMode {   // gr1cs/mod.rs:75-90
    bool setup = false, construct_matrices = true, generate_lc_assignments = true;
    static SynthesisMode Setup() { return {true, true, false}; }
    static SynthesisMode Prove(bool construct_matrices, bool generate_lc_assignments) { return {false, construct_matrices, generate_lc_assignments}; }
};
enum class OptimizationGoal { None, Constraints, Weight };   // gr1cs/mod.rs:96-106

template <class F>
class ConstraintSystem {
public:
    using LC = LinearCombination<F>;
    using Lazy = std::function<F()>;
    using LazyLc = std::function<LC()>;

    ConstraintSystem() {   // constraint_system.rs:109-139
        instance_assignment_.push_back(F::one());
        lcs_.push_back({});
        lc_assignment_.push_back(F::zero());
    }
    // -- counters (constraint_system.rs:210-230)
    size_t num_constraints() const { return constraints_.size(); }
    size_t num_instance_variables() const { return num_instance_; }
    size_t num_witness_variables() const { return num_witness_; }
    size_t num_variables() const { return num_instance_ + num_witness_; }

    void set_mode(SynthesisMode m) { mode_ = m; }
    bool is_in_setup_mode() const { return mode_.setup; }
    bool should_construct_matrices() const { return mode_.setup || mode_.construct_matrices; }
    bool should_generate_lc_assignments() const { return !mode_.setup && mode_.generate_lc_assignments; }
    void set_optimization_goal(OptimizationGoal g) {   // :563-566 asserts is_new
        if (!(num_instance_ == 1 && num_witness_ == 0 && constraints_.empty() && lcs_.size() == 1))
            throw std::logic_error("set_optimization_goal on a non-empty constraint system");
        goal_ = g;
    }
    OptimizationGoal optimization_goal() const { return goal_; }

    // -- allocation (constraint_system.rs:591-617): the closure is only evaluated outside setup mode
    Variable new_input_variable(const Lazy& f) {
        size_t idx = num_instance_++;
        if (!is_in_setup_mode()) instance_assignment_.push_back(f());
        return Variable::instance(idx);
    }
    Variable new_witness_variable(const Lazy& f) {
        size_t idx = num_witness_++;
        if (!is_in_setup_mode()) witness_assignment_.push_back(f());
        return Variable::witness(idx);
    }
    Variable new_lc(const LazyLc& f) { return new_lc_helper(f); }   // :523-532

    // -- constraints (constraint_system.rs:323-353, 431-438)
    void enforce_r1cs_constraint(const LazyLc& a, const LazyLc& b, const LazyLc& c) {
        if (should_construct_matrices()) {
            Variable va = new_lc_helper(a), vb = new_lc_helper(b), vc = new_lc_helper(c);
            constraints_.push_back({va, vb, vc});
        }
    }

    // -- assignments (constraint_system.rs:193-206)
    const std::vector<F>& instance_assignment() const {
        if (is_in_setup_mode()) throw SynthesisFailure(SynthesisError::AssignmentMissing);
        return instance_assignment_;
    }
    const std::vector<F>& witness_assignment() const {
        if (is_in_setup_mode()) throw SynthesisFailure(SynthesisError::AssignmentMissing);
        return witness_assignment_;
    }
    std::optional<F> assigned_value(Variable v) const {   // assignment.rs:26-35
        switch (v.kind()) {
            case VarKind::Zero: return F::zero();
            case VarKind::One: return F::one();
            case VarKind::Instance: return v.payload() < instance_assignment_.size() ? std::optional<F>(instance_assignment_[v.payload()]) : std::nullopt;
            case VarKind::Witness: return v.payload() < witness_assignment_.size() ? std::optional<F>(witness_assignment_[v.payload()]) : std::nullopt;
            default: return v.payload() < lc_assignment_.size() ? std::optional<F>(lc_assignment_[v.payload()]) : std::nullopt;
        }
    }

    // -- finalize: inline_all_lcs (constraint_system.rs:691-758)
    void finalize() {
        if (!should_construct_matrices()) return;
        bool any_used = false;
        for (const auto& l : lcs_) for (const auto& t : l) any_used |= t.second.is_lc();
        if (!any_used) return;   // early return leaves LCs untouched (:722-725)
        std::vector<std::vector<std::pair<F, Variable>>> inlined;
        inlined.reserve(lcs_.size());
        LC out;
        for (const auto& l : lcs_) {
            for (const auto& [coeff, var] : l) {
                if (auto li = var.get_lc_index()) {
                    const auto& sub = inlined[*li];   // already transformed: guaranteed by ordering
                    if (coeff == F::one()) out.terms.insert(out.terms.end(), sub.begin(), sub.end());
                    else for (const auto& [c, v] : sub) if (!v.is_zero() && !c.is_zero()) out.terms.emplace_back(coeff * c, v);
                } else {
                    out.terms.emplace_back(coeff, var);
                }
            }
            out.compactify();
            inlined.push_back(out.terms);
            out.terms.clear();
        }
        lcs_ = std::move(inlined);
    }

    // -- export (constraint_system.rs:768-804; predicate/mod.rs:207-217): [A, B, C]
    LC get_lc(Variable v) const {
        if (v.is_zero()) return {};
        if (v.is_lc()) return LC(lcs_[v.payload()]);
        return LC({{F::one(), v}});
    }
    std::vector<std::pair<F, size_t>> make_row(const LC& l) const {
        std::vector<std::pair<F, size_t>> row;
        for (const auto& [coeff, var] : l.terms) {
            if (coeff.is_zero() || var.is_zero()) continue;
            row.emplace_back(coeff, *var.get_variable_index(num_instance_));
        }
        return row;
    }
    std::vector<Matrix<F>> to_matrices() const {
        std::vector<Matrix<F>> m(3);
        for (const auto& cons : constraints_)
            for (int k = 0; k < 3; k++) m[k].push_back(make_row(get_lc(cons[k])));
        return m;
    }

    // -- satisfaction (constraint_system.rs:652-687; predicate/mod.rs:185-204), R1CS: x0*x1 - x2 == 0
    std::optional<size_t> which_is_unsatisfied() const {
        if (is_in_setup_mode()) throw SynthesisFailure(SynthesisError::AssignmentMissing);
        for (size_t i = 0; i < constraints_.size(); i++) {
            F x[3];
            for (int k = 0; k < 3; k++) {
                Variable v = constraints_[i][k];
                auto val = assigned_value(v);
                if (!val) {
                    F acc = F::zero();
                    for (const auto& [c, u] : get_lc(v).terms) acc = acc + c * *assigned_value(u);
                    val = acc;
                }
                x[k] = *val;
            }
            if (!(x[0] * x[1] - x[2]).is_zero()) return i;
        }
        return std::nullopt;
    }
    bool is_satisfied() const { return !which_is_unsatisfied().has_value(); }

private:
    Variable new_lc_helper(const LazyLc& f) {   // constraint_system.rs:472-519
        if (!(should_construct_matrices() || should_generate_lc_assignments())) return Variable::symbolic_lc(lcs_.size());
        LC l = f();
        const auto& t = l.terms;
        if (t.empty() || (t.size() == 1 && t[0].second.is_zero())) return Variable::symbolic_lc(0);
        if (t.size() == 1 && t[0].first == F::one()) return t[0].second;
        size_t idx = lcs_.size();
        lcs_.push_back(t);
        if (should_generate_lc_assignments()) {   // assignment.rs:40-52
            F acc = F::zero();
            for (const auto& [c, v] : t) acc = acc + c * *assigned_value(v);
            lc_assignment_.push_back(acc);
        }
        return Variable::symbolic_lc(idx);
    }

    size_t num_instance_ = 1, num_witness_ = 0;
    std::vector<F> instance_assignment_, witness_assignment_, lc_assignment_;
    std::vector<std::vector<std::pair<F, Variable>>> lcs_;
    std::vector<std::array<Variable, 3>> constraints_;
    SynthesisMode mode_ = SynthesisMode::Prove(true, true);   // constraint_system.rs:128-131
    OptimizationGoal goal_ = OptimizationGoal::None;
};

// Shared handle with a `None` variant (constraint_system_ref.rs:26-34).
template <class F>
class ConstraintSystemRef {
public:
    ConstraintSystemRef() = default;   // None
    static ConstraintSystemRef new_ref() { ConstraintSystemRef r; r.cs_ = std::make_shared<ConstraintSystem<F>>(); return r; }
    bool is_none() const { return !cs_; }
    ConstraintSystem<F>& inner() const { if (!cs_) throw SynthesisFailure(SynthesisError::MissingCS); return *cs_; }
    ConstraintSystem<F>* operator->() const { return &inner(); }
    Variable new_input_variable(const typename ConstraintSystem<F>::Lazy& f) const { return inner().new_input_variable(f); }
    Variable new_witness_variable(const typename ConstraintSystem<F>::Lazy& f) const { return inner().new_witness_variable(f); }
    Variable new_lc(const typename ConstraintSystem<F>::LazyLc& f) const { return inner().new_lc(f); }
    // constraint_system_ref.rs:235-250: a no-op returning Ok when matrices are not being constructed
    void enforce_r1cs_constraint(const typename ConstraintSystem<F>::LazyLc& a, const typename ConstraintSystem<F>::LazyLc& b,
                                 const typename ConstraintSystem<F>::LazyLc& c) const { inner().enforce_r1cs_constraint(a, b, c); }
    void finalize() const { inner().finalize(); }
    bool is_satisfied() const { return inner().is_satisfied(); }
    std::vector<Matrix<F>> to_matrices() const { return inner().to_matrices(); }
    size_t num_constraints() const { return inner().num_constraints(); }
    size_t num_instance_variables() const { return inner().num_instance_variables(); }
    size_t num_witness_variables() const { return inner().num_witness_variables(); }
    void set_mode(SynthesisMode m) const { inner().set_mode(m); }
    void set_optimization_goal(OptimizationGoal g) const { inner().set_optimization_goal(g); }
private:
    std::shared_ptr<ConstraintSystem<F>> cs_;
};

// gr1cs/mod.rs:54-61
template <class F>
struct ConstraintSynthesizer {
    virtual ~ConstraintSynthesizer() = default;
    virtual void generate_constraints(ConstraintSystemRef<F> cs) = 0;
};

}  // namespace gr1cs
}  // namespace ark_relations
