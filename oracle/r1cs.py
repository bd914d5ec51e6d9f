"""R1CS constraint system, matrix export and SpMV for the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Restates, in plain Python big-ints, the part of ark-relations that produces the inputs of the
prover hot path and the in-tree matrix x witness products:

  * Variable ordering / column index   relations/src/utils/variable.rs:8-14,105-113
  * LinearCombination (+, compactify)  relations/src/utils/linear_combination.rs:53-82,174-210
  * ConstraintSystem builder           relations/src/gr1cs/constraint_system.rs:109-139,323-353,
                                       472-532,591-617
  * finalize / inline_all_lcs          constraint_system.rs:691-758
  * instance outlining                 constraint_system.rs:807-863, instance_outliner.rs:40-60
  * to_matrices / get_lc / make_row    constraint_system.rs:768-804, predicate/mod.rs:207-217
  * is_satisfied (R1CS: x0*x1 - x2)    predicate/mod.rs:115-120,185-204
  * mat_vec_mul                        relations/src/utils/matrix.rs:26-36
  * evaluate_constraint                relations/src/sr1cs/mod.rs:24-56
  * DummyCircuit                       relations/src/sr1cs/mod.rs:296-317
  * BenchCircuit (shape only)          relations/examples/bench.rs:22-83

Pinned by the reference's golden matrices: circuit2.rs:21-43 and circuit1.rs:28-61
(tests/test_oracle_py.py).
"""
from .params import Curve

# Variable = (tag, payload); tags per variable.rs:8-14.  Tuple order == reference `Ord`.
ZERO, ONE, INSTANCE, WITNESS, LC = 0, 1, 2, 3, 4
V_ZERO = (ZERO, 0)
V_ONE = (ONE, 0)


def instance(i):
    return (INSTANCE, i)


def witness(i):
    return (WITNESS, i)


def symbolic_lc(i):
    return (LC, i)


def variable_index(v, witness_offset):
    """variable.rs:105-113."""
    if v[0] == ONE:
        return 0
    if v[0] == INSTANCE:
        return v[1]
    if v[0] == WITNESS:
        return v[1] + witness_offset
    return None


class SynthesisError(Exception):
    """utils/error.rs:5-21 (variant carried as a string)."""


class LinearCombination:
    """linear_combination.rs:15 — a Vec<(F, Variable)>; coefficients are ints mod r."""

    def __init__(self, r, terms=None):
        self.r = r
        self.t = list(terms) if terms else []

    def copy(self):
        return LinearCombination(self.r, self.t)

    def get_var_loc(self, var):
        """linear_combination.rs:174-190: for < 6 terms the linear scan never reports a hit."""
        if len(self.t) < 6:
            found = 0
            for i, (_, v) in enumerate(self.t):
                if v >= var:
                    found = i
                    break
                found += 1
            return (False, found)
        lo, hi = 0, len(self.t)
        while lo < hi:
            mid = (lo + hi) // 2
            if self.t[mid][1] < var:
                lo = mid + 1
            elif self.t[mid][1] > var:
                hi = mid
            else:
                return (True, mid)
        return (False, lo)

    def add_term(self, coeff, var):
        """AddAssign<(F, Variable)> linear_combination.rs:203-211."""
        found, loc = self.get_var_loc(var)
        if found:
            self.t[loc] = ((self.t[loc][0] + coeff) % self.r, var)
        else:
            self.t.insert(loc, (coeff % self.r, var))
        return self

    def __add__(self, other):
        out = self.copy()
        if isinstance(other, tuple) and len(other) == 2 and isinstance(other[1], tuple):
            return out.add_term(other[0], other[1])
        return out.add_term(1, other)  # a bare Variable

    def compactify(self):
        """linear_combination.rs:53-82 (sort is unstable upstream; sums are order-free)."""
        if len(self.t) <= 1:
            return
        self.t.sort(key=lambda e: e[1])
        out = [self.t[0]]
        for c, v in self.t[1:]:
            if out[-1][1] == v:
                out[-1] = ((out[-1][0] + c) % self.r, v)
            else:
                out.append((c, v))
        self.t = out


def lc(r, *vars_or_pairs):
    """The `lc!` macro (linear_combination.rs:19-30): sum_vars / from_sum_coeff_vars, compactified."""
    out = LinearCombination(r)
    for item in vars_or_pairs:
        if isinstance(item[1], tuple):
            out.t.append((item[0] % r, item[1]))
        else:
            out.t.append((1, item))
    out.compactify()
    return out


class ConstraintSystem:
    """R1CS-only restatement of gr1cs::ConstraintSystem (default Prove mode, matrices on)."""

    def __init__(self, curve: Curve, setup_mode=False):
        self.r = curve.r
        self.setup_mode = setup_mode
        self.instance_assignment = [1]           # constraint_system.rs:121
        self.witness_assignment = []
        self.num_instance_variables = 1
        self.num_witness_variables = 0
        self.lcs = [[]]                           # lc 0 == zero LC (constraint_system.rs:112)
        self.lc_assignment = [0]
        self.constraints = []                     # R1CS predicate: list of (a_var, b_var, c_var)
        # other predicates (predicate/mod.rs:81-94): label -> dict(arity, terms, constraints)
        self.predicates = {}
        self.instance_outliner = None             # (pred_label, func) -- instance_outliner.rs:17-26

    # -- allocation (constraint_system.rs:591-617) ------------------------
    def new_input_variable(self, f):
        idx = self.num_instance_variables
        self.num_instance_variables += 1
        if not self.setup_mode:
            self.instance_assignment.append(f() % self.r)
        return instance(idx)

    def new_witness_variable(self, f):
        idx = self.num_witness_variables
        self.num_witness_variables += 1
        if not self.setup_mode:
            self.witness_assignment.append(f() % self.r)
        return witness(idx)

    # -- LCs (constraint_system.rs:472-532) -------------------------------
    def assigned_value(self, v):
        """assignment.rs:26-35."""
        if v[0] == ZERO:
            return 0
        if v[0] == ONE:
            return 1
        if v[0] == INSTANCE:
            return self.instance_assignment[v[1]]
        if v[0] == WITNESS:
            return self.witness_assignment[v[1]]
        return self.lc_assignment[v[1]] if v[1] < len(self.lc_assignment) else None

    def new_lc(self, lcomb: LinearCombination):
        t = lcomb.t
        if len(t) == 0 or (len(t) == 1 and t[0][1] == V_ZERO):
            return symbolic_lc(0)
        if len(t) == 1 and t[0][0] == 1:
            return t[0][1]
        idx = len(self.lcs)
        self.lcs.append(list(t))
        if not self.setup_mode:
            # assignment.rs:40-52 eval_lc
            acc = 0
            for c, v in t:
                acc += c * self.assigned_value(v)
            self.lc_assignment.append(acc % self.r)
        return symbolic_lc(idx)

    def enforce_r1cs_constraint(self, a: LinearCombination, b: LinearCombination, c: LinearCombination):
        self.constraints.append((self.new_lc(a), self.new_lc(b), self.new_lc(c)))

    def num_constraints(self):
        return len(self.constraints) + sum(len(p["constraints"]) for p in self.predicates.values())

    # -- generic polynomial predicates (predicate/mod.rs:96-174, polynomial_constraint.rs:16-66) ------
    def register_predicate(self, label, arity, terms):
        """terms: [(coeff, [(argument index, exponent), ...])]; satisfied iff the polynomial is 0."""
        self.predicates[label] = {"arity": arity, "terms": terms, "constraints": []}

    def enforce_constraint(self, label, lcs):
        if label == "R1CS":
            return self.enforce_r1cs_constraint(*lcs)
        if label not in self.predicates:
            raise SynthesisError("PredicateNotFound")
        pred = self.predicates[label]
        if len(lcs) != pred["arity"]:
            raise SynthesisError("ArityMismatch")
        pred["constraints"].append(tuple(self.new_lc(l) for l in lcs))

    def _lc_value(self, v):
        val = self.assigned_value(v)
        if val is None:
            val = sum(c * self.assigned_value(x) for c, x in self.get_lc(v)) % self.r
        return val

    def to_matrices_all(self):
        """BTreeMap<Label, Vec<Matrix>> of constraint_system.rs:768-774 (labels in lexicographic order)."""
        out = {"R1CS": self.to_matrices()}
        for label, pred in self.predicates.items():
            mats = [[] for _ in range(pred["arity"])]
            for cons in pred["constraints"]:
                for k, v in enumerate(cons):
                    mats[k].append(self.make_row(self.get_lc(v)))
            out[label] = mats
        return dict(sorted(out.items()))

    # -- finalize (constraint_system.rs:691-707) --------------------------
    def finalize(self):
        self.inline_all_lcs()
        if self.instance_outliner is not None:
            label, func = self.instance_outliner
            self.instance_outliner = None
            if label == "R1CS" or label in self.predicates:
                self.perform_instance_outlining(func)

    def set_instance_outliner(self, pred_label, func):
        """constraint_system.rs:807-809."""
        self.instance_outliner = (pred_label, func)

    def perform_instance_outlining(self, func):
        """constraint_system.rs:826-863: a witness copy of ONE and of every instance variable; every stored LC is
        rewritten to use the copies (bare single-variable constraint arguments are not LCs and stay as they are);
        `func` then ties copies to instances."""
        one_w = self.new_witness_variable(lambda: 1)
        imap = [one_w]
        inst = list(self.instance_assignment)
        for i in range(1, self.num_instance_variables):
            imap.append(self.new_witness_variable(lambda i=i: inst[i]))
        for row in self.lcs:
            for k, (c, v) in enumerate(row):
                if v[0] == INSTANCE:
                    row[k] = (c, imap[v[1]])
                elif v[0] == ONE:
                    row[k] = (c, one_w)
        func(self, imap)

    def inline_all_lcs(self):
        """constraint_system.rs:717-758."""
        if not any(v[0] == LC for row in self.lcs for _, v in row):
            return
        inlined = []
        for row in self.lcs:
            out = LinearCombination(self.r)
            for coeff, var in row:
                if var[0] == LC:
                    sub = inlined[var[1]]
                    if coeff == 1:
                        out.t.extend(sub)
                    else:
                        out.t.extend(
                            (coeff * c % self.r, v) for c, v in sub if v != V_ZERO and c != 0
                        )
                else:
                    out.t.append((coeff, var))
            out.compactify()
            inlined.append(out.t)
        self.lcs = inlined

    # -- export (constraint_system.rs:768-804) ----------------------------
    def get_lc(self, var):
        if var == V_ZERO:
            return []
        if var[0] == LC:
            return list(self.lcs[var[1]])
        return [(1, var)]

    def make_row(self, terms):
        off = self.num_instance_variables
        return [(c, variable_index(v, off)) for c, v in terms if c != 0 and v != V_ZERO]

    def to_matrices(self):
        """[A, B, C], each a list of rows [(coeff, col)] (predicate/mod.rs:207-217)."""
        mats = [[], [], []]
        for cons in self.constraints:
            for k in range(3):
                mats[k].append(self.make_row(self.get_lc(cons[k])))
        return mats

    # -- satisfaction (predicate/mod.rs:185-204) --------------------------
    def which_is_unsatisfied(self):
        if self.setup_mode:
            raise SynthesisError("AssignmentMissing")
        for i, cons in enumerate(self.constraints):
            vals = []
            for v in cons:
                val = self.assigned_value(v)
                if val is None:
                    val = sum(c * self.assigned_value(x) for c, x in self.get_lc(v)) % self.r
                vals.append(val)
            if (vals[0] * vals[1] - vals[2]) % self.r != 0:
                return ("R1CS", i)
        for label in sorted(self.predicates):
            pred = self.predicates[label]
            for i, cons in enumerate(pred["constraints"]):
                x = [self._lc_value(v) for v in cons]
                acc = 0
                for coeff, mono in pred["terms"]:
                    t = coeff
                    for idx, e in mono:
                        t = t * pow(x[idx], e, self.r)
                    acc += t
                if acc % self.r != 0:
                    return (label, i)
        return None

    def is_satisfied(self):
        return self.which_is_unsatisfied() is None

    def to_lcmap(self):
        """The flat storage `to_matrices()` reads, as the reference holds it: LcMap {offsets, vars, coeffs}
        (lc_map.rs:51-56), the interner's value vector (field_interner.rs:19-45: [0] = ONE, [1] = -ONE, then values in
        order of first use, ONE always id 0) and the R1CS predicate's argument_lcs (predicate/mod.rs:81-94).
        Variables are the raw u64 of variable.rs:4-14 (tag << 61 | index).  Returns a dict of plain lists."""
        raw = lambda v: (v[0] << 61) | v[1]
        pool, ids = [1, self.r - 1], {1: 0, self.r - 1: 1}
        offsets, vars_, coeffs = [0], [], []
        for row in self.lcs:
            for c, v in row:
                c %= self.r
                if c not in ids:
                    ids[c] = len(pool)
                    pool.append(c)
                coeffs.append(ids[c])
                vars_.append(raw(v))
            offsets.append(len(vars_))
        args = [[raw(cons[k]) for cons in self.constraints] for k in range(3)]
        return {"offsets": offsets, "vars": vars_, "coeffs": coeffs, "pool": pool, "args": args}

    def z(self):
        """instance || witness (sr1cs/mod.rs:199-200)."""
        return self.instance_assignment + self.witness_assignment


def outline_r1cs(cs, instance_witness_map):
    """instance_outliner.rs:40-60: one * one = One, then one * w_i = x_i for every instance variable."""
    r = cs.r
    one = instance_witness_map[0]
    cs.enforce_r1cs_constraint(lc(r, one), lc(r, one), lc(r, V_ONE))
    for i, w in list(enumerate(instance_witness_map))[1:]:
        cs.enforce_r1cs_constraint(lc(r, one), lc(r, w), lc(r, instance(i)))


# ---------------------------------------------------------------------------------------------
# matrix x vector (the in-tree SpMV definitions)
# ---------------------------------------------------------------------------------------------
def mat_vec_mul(r, matrix, vector):
    """utils/matrix.rs:26-36."""
    return [sum(vector[col] * val for val, col in row) % r for row in matrix]


def evaluate_constraint(r, terms, assignment):
    """sr1cs/mod.rs:24-56 (skips the multiply when coeff == 1)."""
    acc = 0
    for coeff, idx in terms:
        acc += assignment[idx] if coeff == 1 else assignment[idx] * coeff
    return acc % r


# ---------------------------------------------------------------------------------------------
# circuits
# ---------------------------------------------------------------------------------------------
def circuit2(curve: Curve, a, b, c):
    """gr1cs/tests/circuit2.rs:47-60."""
    r = curve.r
    cs = ConstraintSystem(curve)
    va = cs.new_input_variable(lambda: a)
    vb = cs.new_witness_variable(lambda: b)
    vc = cs.new_witness_variable(lambda: c)
    L = lambda: LinearCombination(r)
    cs.enforce_r1cs_constraint(L() + va, L() + (2, vb), L() + vc)
    d = cs.new_lc(L() + va + vb)
    cs.enforce_r1cs_constraint(L() + va, L() + d, L() + d)
    e = cs.new_lc(L() + d + d)
    cs.enforce_r1cs_constraint(L() + V_ONE, L() + e, L() + e)
    return cs


def circuit1(curve: Curve, x, w):
    """gr1cs/tests/circuit1.rs:63-164: three polynomial predicates; x = (x1..x5), w = (w1..w8)."""
    r = curve.r
    cs = ConstraintSystem(curve)
    xv = [cs.new_input_variable(lambda v=v: v) for v in x]
    wv = [cs.new_witness_variable(lambda v=v: v) for v in w]
    x1, x2, x3, x4, x5 = xv
    w1, w2, w3, w4, w5, w6, _w7, w8 = wv
    cs.register_predicate("poly-predicate-A", 4, [(1, [(0, 1), (1, 1)]), (3, [(2, 2)]), (r - 1, [(3, 1)])])
    cs.register_predicate("poly-predicate-B", 3, [(7, [(1, 1)]), (1, [(0, 3)]), (r - 1, [(2, 1)])])
    cs.register_predicate("poly-predicate-C", 3, [(1, [(0, 1), (1, 1)]), (r - 1, [(2, 1)])])
    L = lambda: LinearCombination(r)
    cs.enforce_constraint("poly-predicate-A", [L() + x1, L() + x2, L() + x3, L() + w4])
    cs.enforce_constraint("poly-predicate-B", [L() + x4, L() + w1, L() + w5])
    cs.enforce_constraint("poly-predicate-B", [L() + w5, L() + w6, L() + w8])
    cs.enforce_constraint("poly-predicate-C", [L() + w2, L() + w3, L() + w6])
    cs.enforce_constraint("poly-predicate-C", [L() + w5 + w4, L() + w8, L() + x5])
    return cs


CIRCUIT1_GOLDEN = {  # circuit1.rs:28-61
    "R1CS": [[], [], []],
    "poly-predicate-A": [[[(1, 1)]], [[(1, 2)]], [[(1, 3)]], [[(1, 9)]]],
    "poly-predicate-B": [[[(1, 4)], [(1, 10)]], [[(1, 6)], [(1, 11)]], [[(1, 10)], [(1, 13)]]],
    "poly-predicate-C": [[[(1, 7)], [(1, 9), (1, 10)]], [[(1, 8)], [(1, 13)]], [[(1, 11)], [(1, 5)]]],
}
CIRCUIT1_SAT = ((1, 2, 3, 0, 1255254), (4, 2, 5, 29, 28, 10, 57, 22022))       # tests/mod.rs:19-33
CIRCUIT1_UNSAT = ((4, 2, 3, 0, 1255254), (4, 2, 5, 29, 28, 10, 57, 22022))     # tests/mod.rs:57-71

CIRCUIT2_GOLDEN = [  # circuit2.rs:21-43
    [[(1, 1)], [(1, 1)], [(1, 0)]],
    [[(2, 2)], [(1, 1), (1, 2)], [(2, 1), (2, 2)]],
    [[(1, 3)], [(1, 1), (1, 2)], [(2, 1), (2, 2)]],
]


def dummy_circuit(curve: Curve, a, b, num_variables, num_constraints):
    """sr1cs/mod.rs:296-317 through the builder (small sizes)."""
    r = curve.r
    cs = ConstraintSystem(curve)
    va = cs.new_witness_variable(lambda: a)
    vb = cs.new_witness_variable(lambda: b)
    vc = cs.new_input_variable(lambda: a * b % r)
    for _ in range(num_variables - 3):
        cs.new_witness_variable(lambda: a)
    for _ in range(num_constraints - 1):
        cs.enforce_r1cs_constraint(lc(r, va), lc(r, vb), lc(r, vc))
    cs.enforce_r1cs_constraint(lc(r), lc(r), lc(r))
    return cs


def dummy_circuit_direct(curve: Curve, a, b, num_variables, num_constraints):
    """Same R1CS as `dummy_circuit`, emitted directly as (matrices, z_instance, z_witness) so large
    sizes need no synthesis.  Column map: 0 = One, 1 = c (instance), 2 = a, 3 = b, 4.. = a copies."""
    r = curve.r
    n = num_constraints
    A = [[(1, 2)] for _ in range(n - 1)] + [[]]
    B = [[(1, 3)] for _ in range(n - 1)] + [[]]
    C = [[(1, 1)] for _ in range(n - 1)] + [[]]
    inst = [1, a * b % r]
    wit = [a % r, b % r] + [a % r] * (num_variables - 3)
    return [A, B, C], inst, wit


class XorShift64:
    """Deterministic PRNG shared by oracle, tests and bench (NOT the reference's StdRng)."""

    def __init__(self, seed):
        self.s = (seed ^ 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF or 1

    def next(self):
        s = self.s
        s ^= (s << 13) & 0xFFFFFFFFFFFFFFFF
        s ^= s >> 7
        s ^= (s << 17) & 0xFFFFFFFFFFFFFFFF
        self.s = s
        return s

    def below(self, n):
        return self.next() % n

    def field(self, r):
        v = 0
        for _ in range(4):
            v = (v << 64) | self.next()
        return v % r


def bench_circuit(curve: Curve, num_constraints, seed=0):
    """BenchCircuit shape (examples/bench.rs:22-83): rows of 1..10 unit-coefficient terms drawn
    from the last <= 10 allocated variables, every other A row extended by an inlined LC of the same
    size, C rows a single variable; 3 new witnesses per constraint.  The PRNG differs from the
    reference's StdRng, so only the *shape* matches.  Unlike the reference bench (which never checks
    satisfaction and points C at an existing variable), the C row is the first of the three fresh
    witnesses, assigned the product A_i(z) * B_i(z), so the system is satisfiable and usable for proofs."""
    r = curve.r
    cs = ConstraintSystem(curve)
    rng = XorShift64(seed)
    vals = [rng.field(r) for _ in range(3)]
    variables = [cs.new_witness_variable(lambda v=v: v) for v in vals]
    L = lambda: LinearCombination(r)
    for i in range(num_constraints):
        cur = min(len(variables), 10)
        lower, upper = len(variables) - cur, len(variables)
        na = 1 + rng.below(10)
        a_i = LinearCombination(r, [(1, variables[lower + rng.below(cur)]) for _ in range(na)])
        nb = 1 + rng.below(10)
        b_i = LinearCombination(r, [(1, variables[lower + rng.below(cur)]) for _ in range(nb)])
        if i % 2 == 0:
            extra = LinearCombination(r, [(1, variables[lower + rng.below(cur)]) for _ in range(na)])
            ev = cs.new_lc(extra)
            a_i = a_i + ev
        # value of A_i * B_i under the current assignment -> a fresh witness on the C side
        av = sum(c * _val(cs, v) for c, v in a_i.t) % r
        bv = sum(c * _val(cs, v) for c, v in b_i.t) % r
        prod = av * bv % r
        v1 = cs.new_witness_variable(lambda p=prod: p)
        cs.enforce_r1cs_constraint(a_i, b_i, L() + v1)
        v2 = cs.new_witness_variable(lambda: vals[0])
        v3 = cs.new_witness_variable(lambda: vals[0])
        variables.extend([v1, v2, v3])
    return cs


def _val(cs, v):
    val = cs.assigned_value(v)
    return val
