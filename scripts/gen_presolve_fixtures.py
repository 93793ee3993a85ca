"""Generates tests/golden/presolve_changes.json from the reference's own presolve tests
(/root/reference/src/data/linear_program/general_form/presolve/test/changes.rs): every `#[test]` there builds a
`GeneralForm` literal, optionally edits a few fields, and asserts the exact `Changes` (or the `Err(..)`) that
`compute_presolve_changes` returns.  The Rust literals are rewritten to Python expressions mechanically and
evaluated; no reference code is copied, only its test data (run in the build container, where /root/reference is).

    python scripts/gen_presolve_fixtures.py
"""
import json
import os
import re
import sys
from fractions import Fraction as F

SRC = "/root/reference/src/data/linear_program/general_form/presolve/test/changes.rs"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                   "presolve_changes.json")


def matching(text, start, open_ch, close_ch):
    depth = 0
    for k in range(start, len(text)):
        if text[k] == open_ch:
            depth += 1
        elif text[k] == close_ch:
            depth -= 1
            if depth == 0:
                return k
    raise ValueError("unbalanced")


def rewrite_vec(text):
    """vec![a, b] -> [a, b];  vec![x; n] -> REP(x, n)"""
    while True:
        m = re.search(r"vec!\[", text)
        if not m:
            return text
        start = m.end() - 1
        end = matching(text, start, "[", "]")
        inner = text[start + 1:end]
        depth, semi = 0, -1
        for k, ch in enumerate(inner):
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
            elif ch == ";" and depth == 0:
                semi = k
        if semi >= 0:
            new = "REP(" + inner[:semi] + ", " + inner[semi + 1:] + ")"
        else:
            new = "LIST(" + inner + ")"
        text = text[:m.start()] + new + text[end + 1:]


def rewrite_struct(text, name, ctor):
    while True:
        m = re.search(name + r"\s*\{", text)
        if not m:
            return text
        start = m.end() - 1
        end = matching(text, start, "{", "}")
        text = text[:m.start()] + ctor + "(" + text[start + 1:end] + ")" + text[end + 1:]


def rewrite_map_blocks(text):
    """{ let mut m = {}; m.insert(k, v); ...; m } -> MAP((k, v), ...)"""
    while True:
        m = re.search(r"\{\s*let mut (\w+) = \{\};", text)
        if not m:
            return text
        end = matching(text, m.start(), "{", "}")
        name = m.group(1)
        inner = text[m.end():end]
        pairs = []
        pos = 0
        while True:
            k = inner.find(name + ".insert(", pos)
            if k < 0:
                break
            a = k + len(name) + len(".insert")
            b = matching(inner, a, "(", ")")
            pairs.append("(" + inner[a + 1:b] + ")")
            pos = b
        text = text[:m.start()] + "MAP(" + ", ".join(pairs) + ")" + text[end + 1:]


def to_python(body):
    body = re.sub(r"//[^\n]*", "", body)
    body = re.sub(r"R32!\(", "F(", body)
    body = body.replace("GeneralForm::<_>::new", "GeneralForm_new").replace("GeneralForm::new", "GeneralForm_new")
    body = body.replace("ColumnMajor::from_test_data(&", "CM(").replace("ColumnMajor::from_test_data(", "CM(")
    body = body.replace("DenseVector::new(", "DV(").replace("DenseVector::from_test_data(", "DV(")
    body = body.replace("HashMap::default()", "{}").replace("HashMap::new()", "{}")
    body = body.replace(".into_iter().collect()", "").replace(".to_string()", "")
    for a, b in (("RangedConstraintRelation::Less", '"L"'), ("RangedConstraintRelation::Greater", '"G"'),
                 ("RangedConstraintRelation::Equal", '"E"'), ("RangedConstraintRelation::Range(", 'RANGE('),
                 ("Objective::Maximize", '"maximize"'), ("Objective::Minimize", '"minimize"'),
                 ("VariableType::Continuous", '"continuous"'), ("VariableType::Integer", '"integer"'),
                 ("BoundDirection::Lower", "0"), ("BoundDirection::Upper", "1"),
                 ("RemovedVariable::Solved(", "SOLVED("), ("LinearProgramType::Infeasible", '"infeasible"'),
                 ("LinearProgramType::Unbounded", '"unbounded"'), ("false", "False"), ("true", "True")):
        body = body.replace(a, b)
    body = rewrite_map_blocks(body)
    body = rewrite_struct(body, r"RemovedVariable::FunctionOfOthers", "FUNCTION")
    body = rewrite_struct(body, r"Variable", "VARIABLE")
    body = rewrite_struct(body, r"Changes", "CHANGES")
    body = rewrite_vec(body)
    body = re.sub(r"(?<![:\w])(\w+):(?!:)", r"\1=", body)            # struct fields -> keyword arguments
    return body


class Var:
    def __init__(self, variable_type, cost, lower_bound, upper_bound, shift, flipped):
        self.variable_type, self.cost, self.lower_bound, self.upper_bound = variable_type, cost, lower_bound, upper_bound

    def copy(self):
        return Var(self.variable_type, self.cost, self.lower_bound, self.upper_bound, 0, False)


class GF:
    def __init__(self, objective, cm, types, b, variables, names, fixed_cost):
        self.objective, (self.rows, self.ncols), self.types, self.b = objective, cm, list(types), list(b)
        self.variables, self.names, self.fixed_cost = variables, names, fixed_cost


def REP(x, n):
    return [x.copy() if isinstance(x, Var) else x for _ in range(n)]


ENV = dict(F=F, Some=lambda x: x, None_=None, LIST=lambda *a: list(a), REP=REP,
           VARIABLE=lambda **kw: Var(**kw), CM=lambda rows, n: ([list(r) for r in rows], n),
           DV=lambda v, n=None: [F(x) for x in v], RANGE=lambda r: ("R", r), SOLVED=lambda v: ("solved", v),
           FUNCTION=lambda constant, coefficients: ("function", constant, [list(t) for t in coefficients]),
           MAP=lambda *pairs: {k: v for k, v in pairs}, CHANGES=lambda **kw: kw, Ok=lambda x: ("ok", x), Err=lambda x: ("err", x), GeneralForm_new=GF)


def enc(x):
    if isinstance(x, F):
        return f"{x.numerator}/{x.denominator}"
    if isinstance(x, bool) or x is None or isinstance(x, (int, str)):
        return x
    if isinstance(x, dict):
        return [[enc(k), enc(v)] for k, v in sorted(x.items(), key=lambda t: str(t[0]))]
    return [enc(v) for v in x]


def main():
    text = open(SRC).read()
    cases = []
    for m in re.finditer(r"#\[test\]\s*fn (\w+)\(\)\s*\{", text):
        name = m.group(1)
        end = matching(text, m.end() - 1, "{", "}")
        body = text[m.end():end]
        am = re.search(r"assert_eq!\(", body)
        a_end = matching(body, am.end() - 1, "(", ")")
        setup, assertion = body[:am.start()], body[am.end():a_end]
        # the setup: `let [mut] initial = <expr>;` possibly wrapped in a block with field edits
        setup = re.sub(r"let\s+(mut\s+)?initial\s*=\s*\{", "", setup)
        setup = re.sub(r"\binitial\s*\n?\s*\};", "", setup)
        setup = re.sub(r"let\s+(mut\s+)?initial\s*=", "initial =", setup)
        py = to_python(setup)
        stmts = [s.strip() for s in re.split(r";\s*\n", py) if s.strip()]
        env = dict(ENV)
        for s in stmts:
            s = s.rstrip(";")
            s = "\n".join(line.strip() for line in s.splitlines())       # the literals span lines
            exec(s.replace("\n", " "), env)
        gf = env["initial"]
        # the assertion: `initial.compute_presolve_changes(), <expected>`
        apy = to_python(assertion)
        expected_src = apy[apy.index("compute_presolve_changes()") + len("compute_presolve_changes()"):].strip()
        expected_src = expected_src.lstrip(",").strip().rstrip(",")
        kind, value = eval(" ".join(line.strip() for line in expected_src.splitlines()), env)
        dense = [[F(v) for v in row] for row in gf.rows]
        cases.append(dict(
            name=name, objective=gf.objective, rows=enc(dense), ncols=gf.ncols, constraint_types=enc(gf.types),
            b=enc([F(v) for v in gf.b]), fixed_cost=enc(F(gf.fixed_cost)),
            variables=[dict(cost=enc(F(v.cost)), lower=enc(v.lower_bound), upper=enc(v.upper_bound)) for v in gf.variables],
            expect=dict(kind=kind, value=enc(value) if kind == "ok" else value)))
    with open(OUT, "w") as f:
        json.dump(dict(source="relp src/data/linear_program/general_form/presolve/test/changes.rs", cases=cases), f, indent=1)
    print(f"{len(cases)} cases -> {OUT}")


if __name__ == "__main__":
    sys.exit(main())
