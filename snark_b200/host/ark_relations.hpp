// C++ mirror of the ark-relations GR1CS builder API -- the host side that sits ABOVE the C ABI.
//
// The reference is Rust and keeps constraint synthesis on the host ("Rust host code owns constraint
// synthesis and witness assignment exactly as the reference does"); this image has no Rust toolchain, so
// the same interface is provided in C++ with the reference's names, argument meaning and error behaviour,
// for hosts that are C++ and so that tests/native/host_relations_test.cpp can read like the reference's own
// tests (relations/src/gr1cs/tests/mod.rs).  Generic polynomial predicates are supported as in the reference; Groth16
// consumes the R1CS predicate's matrices.
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
//   FieldInterner, InternedField      relations/src/gr1cs/field_interner.rs:13-69
//   LcMap                             relations/src/gr1cs/lc_map.rs:51-56,87-215 (the flat storage b2s_r1cs_upload_lcmap ingests)
//   PolynomialPredicate               relations/src/gr1cs/predicate/polynomial_constraint.rs:16-73
//   PredicateConstraintSystem         relations/src/gr1cs/predicate/mod.rs:81-217
//   InstanceOutliner, outline_*       relations/src/gr1cs/instance_outliner.rs:17-80; constraint_system.rs:807-863
//   Namespace / ns                    relations/src/gr1cs/namespace.rs:9-52 (tracing spans are not mirrored)
//   Sr1csAdapter                      relations/src/sr1cs/mod.rs:18-265
//
// F is any field type with zero()/one()/+/-/*/==/is_zero() -- in this repository b2s::Fp<...> compiled for the
// host.  Coefficient arithmetic here is the builder's own (a handful of additions/multiplications per
// constraint), exactly as in the reference; the prover hot path never runs on the CPU.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <unordered_map>
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
    void negate_in_place() { for (auto& t : terms) t.first = F::zero() - t.first; }   // :163-166
    LinearCombination operator-() const { LinearCombination r = *this; r.negate_in_place(); return r; }
    // LC (+|-) LC: a merge of two variable-sorted term lists (linear_combination.rs:300-343 and the impls below it)
    LinearCombination operator+(const LinearCombination& o) const {
        if (o.terms.empty()) return *this;
        if (terms.empty()) return o;
        return merge(o, false);
    }
    LinearCombination operator-(const LinearCombination& o) const {
        if (o.terms.empty()) return *this;
        if (terms.empty()) return -o;
        return merge(o, true);
    }
    // From<Variable> / From<(F, Variable)> (linear_combination.rs:126-147)
    static LinearCombination from(const Variable& v) { return v.is_zero() ? LinearCombination() : LinearCombination({{F::one(), v}}); }
    static LinearCombination from(const F& c, const Variable& v) { return (c.is_zero() || v.is_zero()) ? LinearCombination() : LinearCombination({{c, v}}); }
    // lc_diff!(a, b) (linear_combination.rs:32-38, 107-113)
    static LinearCombination diff_vars(const Variable& a, const Variable& b) {
        if (a == b) return {};
        return LinearCombination({{F::one(), a}, {F::zero() - F::one(), b}});
    }

private:
    LinearCombination merge(const LinearCombination& o, bool subtract) const {
        LinearCombination r;
        size_t i = 0, j = 0;
        auto other = [&](size_t k) { return subtract ? F::zero() - o.terms[k].first : o.terms[k].first; };
        while (i < terms.size() && j < o.terms.size()) {
            if (o.terms[j].second < terms[i].second) { r.terms.emplace_back(other(j), o.terms[j].second); j++; }
            else if (terms[i].second < o.terms[j].second) r.terms.push_back(terms[i++]);
            else { r.terms.emplace_back(subtract ? terms[i].first - o.terms[j].first : terms[i].first + o.terms[j].first, terms[i].second); i++; j++; }
        }
        for (; i < terms.size(); i++) r.terms.push_back(terms[i]);
        for (; j < o.terms.size(); j++) r.terms.emplace_back(other(j), o.terms[j].second);
        return r;
    }
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

template <class F> LinearCombination<F> lc_diff(const Variable& a, const Variable& b) { return LinearCombination<F>::diff_vars(a, b); }

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

using Label = std::string;
inline const Label R1CS_PREDICATE_LABEL = "R1CS";     // polynomial_constraint.rs:68-69
inline const Label SR1CS_PREDICATE_LABEL = "SR1CS";   // polynomial_constraint.rs:71-73

// field_interner.rs:13-69.  One pool entry per DISTINCT coefficient: vec[0] = ONE, vec[1] = -ONE, ids index vec; ONE is
// always id 0 (which is also what the SpMV kernel keys its "skip the multiplication" on).  Values are compared by their
// in-memory bytes (canonical Montgomery form, so equal elements have equal bytes).
struct InternedField {
    uint32_t id;
    bool operator==(const InternedField& o) const { return id == o.id; }
};
template <class F>
class FieldInterner {
public:
    FieldInterner() { intern(F::one()); intern(F::zero() - F::one()); }
    InternedField get_or_intern(const F& value) {
        if (value == F::one()) return {0};
        auto it = map_.find(key(value));
        return it != map_.end() ? InternedField{it->second} : intern(value);
    }
    std::optional<F> value(InternedField id) const { return id.id < vec_.size() ? std::optional<F>(vec_[id.id]) : std::nullopt; }
    const std::vector<F>& vec() const { return vec_; }   // the pool, in the layout b2s_r1cs_upload_lcmap takes
private:
    static std::string key(const F& v) { return std::string(reinterpret_cast<const char*>(&v), sizeof(F)); }
    InternedField intern(const F& v) {
        const uint32_t id = uint32_t(vec_.size());
        map_[key(v)] = id;
        vec_.push_back(v);
        return {id};
    }
    std::unordered_map<std::string, uint32_t> map_;
    std::vector<F> vec_;
};

// lc_map.rs:51-215.  All linear combinations in three flat arrays: LC i is vars[offsets[i] .. offsets[i+1]) with the
// interned coefficients at the same positions.
template <class F>
class LcMap {
public:
    using Term = std::pair<F, Variable>;
    void push(const std::vector<Term>& lc, FieldInterner<F>& interner) {
        for (const auto& t : lc) { coeffs_.push_back(interner.get_or_intern(t.first)); vars_.push_back(t.second); }
        offsets_.push_back(uint64_t(vars_.size()));
    }
    size_t num_lcs() const { return offsets_.size() - 1; }
    size_t total_lc_size() const { return vars_.size(); }
    // (coefficient id, variable) pairs of LC `idx`, or nullopt past the end (lc_map.rs:188-204)
    std::optional<std::vector<std::pair<InternedField, Variable>>> get(size_t idx) const {
        if (idx >= num_lcs()) return std::nullopt;
        std::vector<std::pair<InternedField, Variable>> out;
        for (uint64_t e = offsets_[idx]; e < offsets_[idx + 1]; e++) out.emplace_back(coeffs_[e], vars_[e]);
        return out;
    }
    // to_non_interned_lc (lc_map.rs:58-63)
    std::vector<Term> get_lc(size_t idx, const FieldInterner<F>& interner) const {
        std::vector<Term> out;
        for (uint64_t e = offsets_.at(idx); e < offsets_.at(idx + 1); e++) out.emplace_back(*interner.value(coeffs_[e]), vars_[e]);
        return out;
    }
    // iter(): f(lc index, first term, one-past-last term) over every LC
    template <class Fn> void for_each_lc(Fn&& f) const { for (size_t i = 0; i < num_lcs(); i++) f(i, offsets_[i], offsets_[i + 1]); }
    // lc_vars_iter_mut(): f gets each LC's variables as a mutable [begin, end) range
    template <class Fn> void lc_vars_iter_mut(Fn&& f) {
        for (size_t i = 0; i < num_lcs(); i++) f(vars_.data() + offsets_[i], vars_.data() + offsets_[i + 1]);
    }
    const std::vector<Variable>& vars() const { return vars_; }
    const std::vector<InternedField>& coeffs() const { return coeffs_; }
    const std::vector<uint64_t>& offsets() const { return offsets_; }
private:
    std::vector<Variable> vars_;
    std::vector<InternedField> coeffs_;
    std::vector<uint64_t> offsets_{0};
};
static_assert(sizeof(InternedField) == 4, "coefficient ids are 4 bytes (lc_map.rs: 4 B/nnz)");

template <class F> class ConstraintSystem;

// A sparse multivariate polynomial; the predicate holds iff it evaluates to zero (polynomial_constraint.rs:16-66).
// terms: (coefficient, [(argument index, exponent), ...]).
template <class F>
struct PolynomialPredicate {
    using Monomial = std::vector<std::pair<size_t, size_t>>;
    using Terms = std::vector<std::pair<F, Monomial>>;
    size_t num_vars = 0;
    Terms terms;
    PolynomialPredicate() = default;
    PolynomialPredicate(size_t arity, Terms t) : num_vars(arity), terms(std::move(t)) {}
    F eval(const std::vector<F>& x) const {
        if (x.size() < num_vars) throw std::logic_error("PolynomialPredicate::eval: too few arguments");
        F acc = F::zero();
        for (const auto& [coeff, mono] : terms) {
            F t = coeff;
            for (const auto& [var, power] : mono)
                for (size_t k = 0; k < power; k++) t = t * x[var];
            acc = acc + t;
        }
        return acc;
    }
    bool is_satisfied(const std::vector<F>& x) const { return eval(x).is_zero(); }
    size_t arity() const { return num_vars; }
    size_t degree() const {   // largest total degree of a term
        size_t d = 0;
        for (const auto& [coeff, mono] : terms) { size_t t = 0; for (const auto& vp : mono) t += vp.second; d = std::max(d, t); }
        return d;
    }
};

// The constraints enforced under one predicate, stored column-wise: argument_lcs[k][i] is the k-th argument of the
// i-th constraint (predicate/mod.rs:81-94).
template <class F>
class PredicateConstraintSystem {
public:
    using Terms = typename PolynomialPredicate<F>::Terms;
    static PredicateConstraintSystem new_polynomial_predicate_cs(size_t arity, Terms terms) {   // :108-112
        PredicateConstraintSystem p;
        p.predicate_ = PolynomialPredicate<F>(arity, std::move(terms));
        p.argument_lcs_.assign(arity, {});
        return p;
    }
    static PredicateConstraintSystem new_r1cs() {   // :116-121   x0 * x1 - x2
        const F one = F::one(), minus_one = F::zero() - F::one();
        return new_polynomial_predicate_cs(3, {{one, {{0, 1}, {1, 1}}}, {minus_one, {{2, 1}}}});
    }
    static PredicateConstraintSystem new_sr1cs_predicate() {   // :124-129   x0^2 - x1
        const F one = F::one(), minus_one = F::zero() - F::one();
        return new_polynomial_predicate_cs(2, {{one, {{0, 2}}}, {minus_one, {{1, 1}}}});
    }
    size_t get_arity() const { return predicate_.arity(); }
    size_t num_constraints() const { return num_constraints_; }
    const std::vector<std::vector<Variable>>& get_constraints() const { return argument_lcs_; }
    const PolynomialPredicate<F>& get_predicate() const { return predicate_; }

    // predicate/mod.rs:156-174.  As upstream: the arguments are zipped with the columns BEFORE the arity check, so a
    // short list leaves its partial push behind and fails with ArityMismatch, and surplus arguments are dropped.
    void enforce_constraint(const std::vector<Variable>& constraint) {
        size_t arity = 0;
        for (; arity < constraint.size() && arity < argument_lcs_.size(); arity++) argument_lcs_[arity].push_back(constraint[arity]);
        if (arity != get_arity()) throw SynthesisFailure(SynthesisError::ArityMismatch);
        num_constraints_++;
    }
    std::vector<Variable> constraint(size_t i) const {
        std::vector<Variable> c;
        for (const auto& col : argument_lcs_) c.push_back(col[i]);
        return c;
    }
    std::optional<size_t> which_constraint_is_unsatisfied(const ConstraintSystem<F>& cs) const;   // :185-204
    std::vector<Matrix<F>> to_matrices(const ConstraintSystem<F>& cs) const;                      // :207-217

private:
    std::vector<std::vector<Variable>> argument_lcs_;
    size_t num_constraints_ = 0;
    PolynomialPredicate<F> predicate_;
};

// instance_outliner.rs:17-26: which predicate ties the copies to the instances, and how.
template <class F>
struct InstanceOutliner {
    Label pred_label;
    std::function<void(ConstraintSystem<F>&, const std::vector<Variable>&)> func;
};

template <class F>
class ConstraintSystem {
public:
    using LC = LinearCombination<F>;
    using Lazy = std::function<F()>;
    using LazyLc = std::function<LC()>;

    ConstraintSystem() {   // constraint_system.rs:109-139: One is instance 0, LC 0 is the empty LC, R1CS is registered
        instance_assignment_.push_back(F::one());
        lc_map_.push({}, interner_);
        lc_assignment_.push_back(F::zero());
        register_predicate(R1CS_PREDICATE_LABEL, PredicateConstraintSystem<F>::new_r1cs());
    }
    // -- counters (constraint_system.rs:210-236)
    size_t num_constraints() const { size_t n = 0; for (const auto& kv : predicates_) n += kv.second.num_constraints(); return n; }
    size_t num_instance_variables() const { return num_instance_; }
    size_t num_witness_variables() const { return num_witness_; }
    size_t num_variables() const { return num_instance_ + num_witness_; }
    size_t num_predicates() const { return predicates_.size(); }

    // -- predicates (constraint_system.rs:146-191, 620-642)
    void register_predicate(const Label& label, PredicateConstraintSystem<F> p) { predicates_.insert_or_assign(label, std::move(p)); }
    void remove_predicate(const Label& label) { predicates_.erase(label); }
    bool has_predicate(const Label& label) const { return predicates_.count(label) != 0; }
    std::optional<size_t> get_predicate_num_constraints(const Label& label) const {
        auto it = predicates_.find(label);
        return it == predicates_.end() ? std::nullopt : std::optional<size_t>(it->second.num_constraints());
    }
    std::optional<size_t> get_predicate_arity(const Label& label) const {
        auto it = predicates_.find(label);
        return it == predicates_.end() ? std::nullopt : std::optional<size_t>(it->second.get_arity());
    }
    std::map<Label, size_t> get_all_predicates_num_constraints() const {
        std::map<Label, size_t> m; for (const auto& kv : predicates_) m[kv.first] = kv.second.num_constraints(); return m;
    }
    std::map<Label, size_t> get_all_predicate_arities() const {
        std::map<Label, size_t> m; for (const auto& kv : predicates_) m[kv.first] = kv.second.get_arity(); return m;
    }
    const std::map<Label, PredicateConstraintSystem<F>>& predicates() const { return predicates_; }

    void set_mode(SynthesisMode m) { mode_ = m; }
    bool is_in_setup_mode() const { return mode_.setup; }
    bool should_construct_matrices() const { return mode_.setup || mode_.construct_matrices; }
    bool should_generate_lc_assignments() const { return !mode_.setup && mode_.generate_lc_assignments; }
    void set_optimization_goal(OptimizationGoal g) {   // :563-566 asserts is_new (:554-559)
        if (!(num_instance_ == 1 && num_witness_ == 0 && num_constraints() == 0 && num_lcs_ == 1))
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

    // -- constraints (constraint_system.rs:241-451)
    void enforce_constraint(const Label& label, const std::vector<LazyLc>& lcs) {
        auto it = predicates_.find(label);
        if (it == predicates_.end()) throw SynthesisFailure(SynthesisError::PredicateNotFound);
        if (!should_construct_matrices()) return;
        std::vector<Variable> args;
        for (const auto& f : lcs) args.push_back(new_lc_helper(f));   // new_constraint_lc :455-461
        it->second.enforce_constraint(args);
    }
    void enforce_constraint_arity_2(const Label& l, const LazyLc& a, const LazyLc& b) { enforce_constraint(l, {a, b}); }
    void enforce_constraint_arity_3(const Label& l, const LazyLc& a, const LazyLc& b, const LazyLc& c) { enforce_constraint(l, {a, b, c}); }
    void enforce_constraint_arity_4(const Label& l, const LazyLc& a, const LazyLc& b, const LazyLc& c, const LazyLc& d) { enforce_constraint(l, {a, b, c, d}); }
    void enforce_constraint_arity_5(const Label& l, const LazyLc& a, const LazyLc& b, const LazyLc& c, const LazyLc& d, const LazyLc& e) {
        enforce_constraint(l, {a, b, c, d, e});
    }
    void enforce_r1cs_constraint(const LazyLc& a, const LazyLc& b, const LazyLc& c) { enforce_constraint_arity_3(R1CS_PREDICATE_LABEL, a, b, c); }
    void enforce_sr1cs_constraint(const LazyLc& a, const LazyLc& b) { enforce_constraint_arity_2(SR1CS_PREDICATE_LABEL, a, b); }

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
    // value of a constraint argument: the stored assignment, else the LC evaluated term by term (predicate/mod.rs:190-197)
    F argument_value(Variable v) const {
        if (auto val = assigned_value(v)) return *val;
        F acc = F::zero();
        for (const auto& [c, u] : get_lc(v).terms) {
            auto x = assigned_value(u);
            if (!x) throw std::logic_error("variable is not assigned; did you run cs.finalize()?");
            acc = acc + c * *x;
        }
        return acc;
    }

    // -- finalize (constraint_system.rs:691-707): inline, then outline the instances if asked to
    void finalize() {
        inline_all_lcs();
        if (instance_outliner_) {
            InstanceOutliner<F> o = std::move(*instance_outliner_);
            instance_outliner_.reset();
            if (has_predicate(o.pred_label)) {
                try { perform_instance_outlining(o); } catch (const SynthesisFailure&) {}   // `let _ =` upstream
            }
        }
    }
    void inline_all_lcs() {   // :717-758
        if (!should_construct_matrices()) return;
        bool any_used = false;
        for (const Variable& v : lc_map_.vars()) any_used |= v.is_lc();
        if (!any_used) return;   // early return leaves LCs untouched (:722-725)
        LcMap<F> inlined;
        LC out;
        for (size_t i = 0; i < lc_map_.num_lcs(); i++) {
            for (const auto& [coeff, var] : lc_map_.get_lc(i, interner_)) {
                if (auto li = var.get_lc_index()) {
                    const auto sub = inlined.get_lc(*li, interner_);   // already transformed: guaranteed by ordering
                    if (coeff == F::one()) out.terms.insert(out.terms.end(), sub.begin(), sub.end());
                    else for (const auto& [c, v] : sub) if (!v.is_zero() && !c.is_zero()) out.terms.emplace_back(coeff * c, v);
                } else {
                    out.terms.emplace_back(coeff, var);
                }
            }
            out.compactify();
            inlined.push(out.terms, interner_);
            out.terms.clear();
        }
        lc_map_ = std::move(inlined);
    }

    // -- instance outlining (constraint_system.rs:807-863)
    void set_instance_outliner(InstanceOutliner<F> o) { instance_outliner_ = std::move(o); }
    bool should_outline_instances() const { return instance_outliner_.has_value(); }
    void perform_instance_outlining(const InstanceOutliner<F>& outliner) {
        std::vector<Variable> instance_to_witness;
        const Variable one_witness = new_witness_variable([] { return F::one(); });
        instance_to_witness.push_back(one_witness);
        const std::vector<F> inst = instance_assignment_;
        for (size_t i = 1; i < num_instance_; i++)
            instance_to_witness.push_back(new_witness_variable([&] {
                if (i >= inst.size()) throw SynthesisFailure(SynthesisError::AssignmentMissing);
                return inst[i];
            }));
        // rewritten in place: the terms keep their positions, so rows may come out unsorted (the C ABI allows that)
        lc_map_.lc_vars_iter_mut([&](Variable* v, Variable* end) {
            for (; v != end; ++v) {
                if (v->is_instance()) *v = instance_to_witness[v->payload()];
                else if (v->is_one()) *v = one_witness;
            }
        });
        outliner.func(*this, instance_to_witness);
    }

    // -- export (constraint_system.rs:768-804): label -> one matrix per predicate argument
    LC get_lc(Variable v) const {
        if (v.is_zero()) return {};
        if (v.is_lc()) return LC(lc_map_.get_lc(v.payload(), interner_));
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
    std::map<Label, std::vector<Matrix<F>>> to_matrices() const {
        std::map<Label, std::vector<Matrix<F>>> m;
        for (const auto& kv : predicates_) m[kv.first] = kv.second.to_matrices(*this);
        return m;
    }

    // -- flat storage, as b2s_r1cs_upload_lcmap takes it (include/b200snark.h)
    const LcMap<F>& lc_map() const { return lc_map_; }
    const FieldInterner<F>& field_interner() const { return interner_; }

    // -- satisfaction (constraint_system.rs:652-687): "<label> - <index>" of the first failing constraint, predicates
    // visited in label order (the form upstream reports when no ConstraintLayer trace is installed)
    std::optional<std::string> which_is_unsatisfied() const {
        if (is_in_setup_mode()) throw SynthesisFailure(SynthesisError::AssignmentMissing);
        for (const auto& kv : predicates_)
            if (auto i = kv.second.which_constraint_is_unsatisfied(*this)) return kv.first + " - " + std::to_string(*i);
        return std::nullopt;
    }
    bool is_satisfied() const { return !which_is_unsatisfied().has_value(); }

private:
    Variable new_lc_helper(const LazyLc& f) {   // constraint_system.rs:472-519
        if (!(should_construct_matrices() || should_generate_lc_assignments())) return Variable::symbolic_lc(num_lcs_++);   // :465-469
        LC l = f();
        const auto& t = l.terms;
        if (t.empty() || (t.size() == 1 && t[0].second.is_zero())) return Variable::symbolic_lc(0);
        if (t.size() == 1 && t[0].first == F::one()) return t[0].second;
        size_t idx = num_lcs_++;
        lc_map_.push(t, interner_);
        if (should_generate_lc_assignments()) {   // assignment.rs:40-52
            F acc = F::zero();
            for (const auto& [c, v] : t) acc = acc + c * *assigned_value(v);
            lc_assignment_.push_back(acc);
        }
        return Variable::symbolic_lc(idx);
    }

    size_t num_instance_ = 1, num_witness_ = 0, num_lcs_ = 1;
    std::vector<F> instance_assignment_, witness_assignment_, lc_assignment_;
    LcMap<F> lc_map_;              // `pub lc_map` upstream (#[doc(hidden)], constraint_system.rs:85-86)
    FieldInterner<F> interner_;    // private upstream (constraint_system.rs:88)
    std::map<Label, PredicateConstraintSystem<F>> predicates_;   // BTreeMap upstream: iteration in label order
    std::optional<InstanceOutliner<F>> instance_outliner_;
    SynthesisMode mode_ = SynthesisMode::Prove(true, true);   // constraint_system.rs:128-131
    OptimizationGoal goal_ = OptimizationGoal::None;
};

template <class F>
std::optional<size_t> PredicateConstraintSystem<F>::which_constraint_is_unsatisfied(const ConstraintSystem<F>& cs) const {
    std::vector<F> x(argument_lcs_.size(), F::zero());
    for (size_t i = 0; i < num_constraints_; i++) {
        for (size_t k = 0; k < argument_lcs_.size(); k++) x[k] = cs.argument_value(argument_lcs_[k][i]);
        if (!predicate_.is_satisfied(x)) return i;
    }
    return std::nullopt;
}
template <class F>
std::vector<Matrix<F>> PredicateConstraintSystem<F>::to_matrices(const ConstraintSystem<F>& cs) const {
    std::vector<Matrix<F>> m(get_arity());
    for (size_t i = 0; i < num_constraints_; i++)
        for (size_t k = 0; k < argument_lcs_.size(); k++) m[k].push_back(cs.make_row(cs.get_lc(argument_lcs_[k][i])));
    return m;
}

// instance_outliner.rs:40-60: one_w * one_w = One, then one_w * w_i = x_i for every instance variable
template <class F>
void outline_r1cs(ConstraintSystem<F>& cs, const std::vector<Variable>& instance_witness_map) {
    const Variable one = instance_witness_map[0];
    cs.enforce_r1cs_constraint([&] { return lc<F>({one}); }, [&] { return lc<F>({one}); }, [&] { return lc<F>({Variable::One()}); });
    for (size_t i = 1; i < instance_witness_map.size(); i++) {
        const Variable w = instance_witness_map[i];
        cs.enforce_r1cs_constraint([&] { return lc<F>({one}); }, [&] { return lc<F>({w}); }, [&] { return lc<F>({Variable::instance(i)}); });
    }
}
// instance_outliner.rs:63-80: (x_i - w_i)^2 = 0 for every instance variable, the constant included
template <class F>
void outline_sr1cs(ConstraintSystem<F>& cs, const std::vector<Variable>& instance_witness_map) {
    for (size_t i = 0; i < instance_witness_map.size(); i++) {
        const Variable w = instance_witness_map[i];
        cs.enforce_sr1cs_constraint([&] { return lc_diff<F>(Variable::instance(i), w); }, [] { return lc<F>(); });
    }
}

// Shared handle with a `None` variant (constraint_system_ref.rs:26-34).
template <class F>
class ConstraintSystemRef {
public:
    using CS = ConstraintSystem<F>;
    ConstraintSystemRef() = default;   // None
    static ConstraintSystemRef new_ref() { ConstraintSystemRef r; r.cs_ = std::make_shared<CS>(); return r; }
    bool is_none() const { return !cs_; }
    ConstraintSystemRef or_(const ConstraintSystemRef& other) const { return is_none() ? other : *this; }   // :457-462
    CS& inner() const { if (!cs_) throw SynthesisFailure(SynthesisError::MissingCS); return *cs_; }
    CS* operator->() const { return &inner(); }
    Variable new_input_variable(const typename CS::Lazy& f) const { return inner().new_input_variable(f); }
    Variable new_witness_variable(const typename CS::Lazy& f) const { return inner().new_witness_variable(f); }
    Variable new_lc(const typename CS::LazyLc& f) const { return inner().new_lc(f); }
    // constraint_system_ref.rs:144-270: no-ops returning Ok when matrices are not being constructed
    void enforce_constraint(const Label& l, const std::vector<typename CS::LazyLc>& lcs) const { inner().enforce_constraint(l, lcs); }
    void enforce_constraint_arity_2(const Label& l, const typename CS::LazyLc& a, const typename CS::LazyLc& b) const { inner().enforce_constraint_arity_2(l, a, b); }
    void enforce_constraint_arity_3(const Label& l, const typename CS::LazyLc& a, const typename CS::LazyLc& b, const typename CS::LazyLc& c) const {
        inner().enforce_constraint_arity_3(l, a, b, c);
    }
    void enforce_constraint_arity_4(const Label& l, const typename CS::LazyLc& a, const typename CS::LazyLc& b, const typename CS::LazyLc& c,
                                    const typename CS::LazyLc& d) const { inner().enforce_constraint_arity_4(l, a, b, c, d); }
    void enforce_constraint_arity_5(const Label& l, const typename CS::LazyLc& a, const typename CS::LazyLc& b, const typename CS::LazyLc& c,
                                    const typename CS::LazyLc& d, const typename CS::LazyLc& e) const { inner().enforce_constraint_arity_5(l, a, b, c, d, e); }
    void enforce_r1cs_constraint(const typename CS::LazyLc& a, const typename CS::LazyLc& b, const typename CS::LazyLc& c) const { inner().enforce_r1cs_constraint(a, b, c); }
    void enforce_sr1cs_constraint(const typename CS::LazyLc& a, const typename CS::LazyLc& b) const { inner().enforce_sr1cs_constraint(a, b); }
    void register_predicate(const Label& l, PredicateConstraintSystem<F> p) const { inner().register_predicate(l, std::move(p)); }
    void remove_predicate(const Label& l) const { inner().remove_predicate(l); }
    bool has_predicate(const Label& l) const { return cs_ && cs_->has_predicate(l); }   // :403-406
    size_t num_predicates() const { return inner().num_predicates(); }
    void finalize() const { if (cs_) cs_->finalize(); }                 // :435-439: None is a no-op
    void inline_all_lcs() const { if (cs_) cs_->inline_all_lcs(); }
    bool is_satisfied() const { return inner().is_satisfied(); }
    std::optional<std::string> which_is_unsatisfied() const { return inner().which_is_unsatisfied(); }
    std::optional<F> assigned_value(Variable v) const { return cs_ ? cs_->assigned_value(v) : std::nullopt; }
    std::map<Label, std::vector<Matrix<F>>> to_matrices() const { return inner().to_matrices(); }
    size_t num_constraints() const { return inner().num_constraints(); }
    size_t num_instance_variables() const { return inner().num_instance_variables(); }
    size_t num_witness_variables() const { return inner().num_witness_variables(); }
    size_t num_variables() const { return inner().num_variables(); }
    void set_mode(SynthesisMode m) const { inner().set_mode(m); }
    bool is_in_setup_mode() const { return cs_ && cs_->is_in_setup_mode(); }
    bool should_construct_matrices() const { return cs_ && cs_->should_construct_matrices(); }
    void set_optimization_goal(OptimizationGoal g) const { inner().set_optimization_goal(g); }
    OptimizationGoal optimization_goal() const { return cs_ ? cs_->optimization_goal() : OptimizationGoal::Constraints; }   // :305-309
    void set_instance_outliner(InstanceOutliner<F> o) const { inner().set_instance_outliner(std::move(o)); }
    bool should_outline_instances() const { return cs_ && cs_->should_outline_instances(); }
private:
    std::shared_ptr<CS> cs_;
};

// namespace.rs:9-52.  Upstream a namespace is a tracing span around a clone of the handle; the span is not mirrored.
template <class F>
struct Namespace {
    ConstraintSystemRef<F> inner;
    std::string name;
    ConstraintSystemRef<F> cs() const { return inner; }
    void leave_namespace() {}
};
template <class F> Namespace<F> ns(const ConstraintSystemRef<F>& cs, std::string name) { return {cs, std::move(name)}; }

// gr1cs/mod.rs:54-61
template <class F>
struct ConstraintSynthesizer {
    virtual ~ConstraintSynthesizer() = default;
    virtual void generate_constraints(ConstraintSystemRef<F> cs) = 0;
};

}  // namespace gr1cs

// R1CS -> Square-R1CS (the Groth-Maller17 shape), relations/src/sr1cs/mod.rs:18-265.  a*b = c becomes
// (a + b)^2 = 4c + s and (a - b)^2 = s with one fresh witness s per row; the old instance variables become
// witnesses and are tied to fresh instance variables by (x_old - x_new)^2 = 0.
namespace sr1cs {
using namespace gr1cs;

template <class F>
struct Sr1csAdapter {
    using Row = std::vector<std::pair<F, size_t>>;
    using LC = LinearCombination<F>;

    // sr1cs/mod.rs:24-56: inner product of a matrix row with the assignment; no multiplication for unit coefficients
    static F evaluate_constraint(const Row& terms, const std::vector<F>& assignment) {
        const F one = F::one();
        F sum = F::zero();
        for (const auto& [coeff, index] : terms) sum = sum + (coeff == one ? assignment[index] : assignment[index] * coeff);
        return sum;
    }

    // :122-184 -- shape only (Setup mode); needs a system whose only predicate is R1CS
    static ConstraintSystemRef<F> r1cs_to_sr1cs(const ConstraintSystemRef<F>& cs) {
        if (cs.num_predicates() != 1) throw std::logic_error("r1cs_to_sr1cs: exactly one predicate expected");
        return convert(cs.inner(), false);
    }
    // :192-264 -- with the assignment carried over; the result is finalized
    static ConstraintSystemRef<F> r1cs_to_sr1cs_with_assignment(ConstraintSystem<F>& cs) {
        auto out = convert(cs, true);
        cs.set_mode(SynthesisMode::Prove(true, true));
        out.finalize();
        return out;
    }

private:
    static ConstraintSystemRef<F> convert(ConstraintSystem<F>& cs, bool with_assignment) {
        const auto all = cs.to_matrices();
        const auto& m = all.at(R1CS_PREDICATE_LABEL);
        const size_t num_public = cs.num_instance_variables();
        std::vector<F> z;
        if (with_assignment) {
            z = cs.instance_assignment();
            z.insert(z.end(), cs.witness_assignment().begin(), cs.witness_assignment().end());
        }
        auto out = ConstraintSystemRef<F>::new_ref();
        out.remove_predicate(R1CS_PREDICATE_LABEL);
        out.register_predicate(SR1CS_PREDICATE_LABEL, PredicateConstraintSystem<F>::new_sr1cs_predicate());
        if (with_assignment) out.set_optimization_goal(OptimizationGoal::Constraints);
        else out.set_mode(SynthesisMode::Setup());
        std::map<size_t, Variable> public_vars, witness_vars;   // old column -> new witness, in order of first use
        const F one = F::one();
        auto value_of = [&](size_t col) { return with_assignment ? z[col] : one; };
        // a row of the old matrix as an LC over the new variables, with its value
        auto translate = [&](const Row& row) {
            std::pair<LC, F> r{LC(), F::zero()};
            for (const auto& term : row) {
                const F coeff = term.first;
                const size_t col = term.second;
                Variable v = Variable::One();
                if (col != 0) {
                    auto& table = col < num_public ? public_vars : witness_vars;
                    auto it = table.find(col);
                    if (it == table.end()) it = table.emplace(col, out.new_witness_variable([&] { return value_of(col); })).first;
                    v = it->second;
                }
                r.first.terms.emplace_back(coeff, v);   // collected as is: neither sorted nor merged upstream
                r.second = r.second + coeff * value_of(col);
            }
            return r;
        };
        for (size_t i = 0; i < m[0].size() && i < m[1].size() && i < m[2].size(); i++) {
            auto [a, a_val] = translate(m[0][i]);
            auto [b, b_val] = translate(m[1][i]);
            auto [c, c_val] = translate(m[2][i]);
            (void)c_val;
            const F d = a_val - b_val;
            const Variable square = out.new_witness_variable([&] { return with_assignment ? d * d : one; });
            for (auto& t : c.terms) { t.first = t.first + t.first; t.first = t.first + t.first; }
            const LC a_lc = a, b_lc = b, c_lc = c;
            out.enforce_sr1cs_constraint([&] { return a_lc + b_lc; }, [&] { return c_lc + square; });
            out.enforce_sr1cs_constraint([&] { return a_lc - b_lc; }, [&] { return lc<F>({square}); });
        }
        for (const auto& kv : public_vars) {
            const Variable old_var = kv.second;
            F value = one;
            if (with_assignment) {
                auto v = out.assigned_value(old_var);
                if (!v) throw SynthesisFailure(SynthesisError::AssignmentMissing);
                value = *v;
            }
            const Variable new_var = out.new_input_variable([&] { return value; });
            out.enforce_sr1cs_constraint([&] { return lc_diff<F>(old_var, new_var); }, [] { return lc<F>(); });
        }
        return out;
    }
};
}  // namespace sr1cs
}  // namespace ark_relations
