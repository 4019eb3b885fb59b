"""Unit tests of oracle/tla_eval.py itself (the evaluator that runs the reference's VSR.tla): they need no reference file,
so the tool that pins the oracle is checked on every machine."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import tla_eval as T  # noqa: E402

MODULE = r"""
---------------------------- MODULE Toy ----------------------------
EXTENDS Naturals, Sequences
CONSTANTS N,     \* a number
          Red, Green
VARIABLES x, bag,
          log
(* a block comment (* nested *) with == and /\ inside *)
vars == << x, bag, log >>
Inc(m) == IF m \in DOMAIN bag
          THEN [bag EXCEPT ![m] = @ + 1]
          ELSE bag @@ (m :> 1)
Init ==
    /\ x = 0
    /\ bag = <<>>
    /\ log = <<>>
Step ==
    /\ x < N
    /\ \E c \in {Red, Green}, k \in 1..2 :
        /\ x' = x + k
        /\ bag' = Inc([colour |-> c, n |-> k])
        /\ LET e == [v |-> x, c |-> c]
           IN \/ /\ k = 1
                 /\ log' = Append(log, e)
              \/ /\ k = 2
                 /\ UNCHANGED log
Reset ==
    /\ x >= N
    /\ x' = 0
    /\ bag' = <<>>
    /\ UNCHANGED << log >>
    /\ Len(log) < 2
Next == \/ Step
        \/ Reset
Small == x <= N + 1
Pick == CHOOSE m \in DOMAIN bag : ~\E m1 \in DOMAIN bag : bag[m1] > bag[m]
=====================================================================
"""


def make(n=3):
    m = T.Module(MODULE)
    return T.Evaluator(m, {"N": n, "Red": T.ModelValue("Red"), "Green": T.ModelValue("Green")})


def test_module_structure():
    ev = make()
    assert ev.m.variables == ["x", "bag", "log"] and ev.m.constants == ["N", "Red", "Green"]
    assert {"Init", "Step", "Reset", "Next", "Inc", "Small", "vars"} <= set(ev.m.defs)


def test_init_and_successors_follow_tlc_assignment_rules():
    ev = make()
    (init,) = ev.initial_states()
    assert init == {"x": 0, "bag": T.EMPTY, "log": T.EMPTY}
    succ = ev.successors(init)
    assert len(succ) == 4 and {a for a, _ in succ} == {"Step"}  # 2 colours x 2 increments, Reset disabled
    by_x = sorted(sp["x"] for _, sp in succ)
    assert by_x == [1, 1, 2, 2]
    for _, sp in succ:
        (msg,) = sp["bag"].d
        assert sp["bag"].d[msg] == 1 and msg.d["n"] == sp["x"]
        assert len(sp["log"].d) == (1 if sp["x"] == 1 else 0)  # the inner disjunction: Append or UNCHANGED


def test_bag_counts_and_except_at():
    ev = make()
    st = ev.initial_states()[0]
    for _ in range(2):
        st = [sp for _, sp in ev.successors(st) if sp["x"] - st["x"] == 1 and next(iter(sp["bag"].d)).d["colour"].name == "Red"][0]
    (msg,) = st["bag"].d
    assert st["bag"].d[msg] == 2 and T.fmt(st["log"]) == "<<[v |-> 0, c |-> Red], [v |-> 1, c |-> Red]>>"


def test_bfs_counts_and_invariant():
    ev = make(2)
    r = T.bfs(ev, view=None, invariant="Small")
    assert r["violation_depth"] == 0 and r["distinct"] > 4 and r["generated"] > r["distinct"]
    r2 = T.bfs(make(2), view=None, invariant="Small", max_depth=3)
    assert r2["level_sizes"] == r["level_sizes"][:3]


def test_choose_reports_ambiguity():
    ev = make()
    st = ev.initial_states()[0]
    a = [sp for _, sp in ev.successors(st)][0]
    b = [sp for _, sp in ev.successors(a) if len(sp["bag"].d) == 2][0]
    ev.s, ev.sp, ev.choose_log = b, None, []
    ev.lookup("Pick", {})
    assert ev.choose_log == [2]  # two records with the same (maximal) count: the pick is this module's, not TLC's


@pytest.mark.parametrize("text,value", [
    ("1 + 2 * 3", 7), ("7 \\div 2", 3), ("7 % 3", 1), ("{1, 2} \\union {3}", frozenset({1, 2, 3})), ("1..3 \\ {2}", frozenset({1, 3})),
    ("Len(<<4, 5, 6>>)", 3), ("Append(<<1>>, 2)[2]", 2), ("SubSeq(<<1, 2, 3>>, 2, 3)", T.seq([2, 3])),
    ("[i \\in 1..2 |-> i * i][2]", 4), ("[a |-> 1, b |-> 2].b", 2), ("DOMAIN [a |-> 1]", frozenset({"a"})),
    ("[[a |-> 1, b |-> [c |-> 2]] EXCEPT !.b.c = @ + 5].b.c", 7), ("(1 :> 2 @@ 1 :> 3)[1]", 2),
    ("Cardinality({x \\in 1..10 : x % 2 = 0})", 5), ("{x * 2 : x \\in 1..3}", frozenset({2, 4, 6})),
    ("\\A x \\in {} : FALSE", True), ("\\E x \\in 1..3, y \\in 1..3 : x + y = 6", True), ("IF 1 < 2 THEN 10 ELSE 20", 10),
    ("LET f(a) == a + 1 IN f(f(1))", 3), ("CHOOSE x \\in 1..5 : x * x = 16", 4), ("<<>> = [i \\in {} |-> 0]", True),
    ("Quantify({1, 2, 3}, LAMBDA v : v > 1)", 2), ("Cardinality(Permutations({1, 2, 3}))", 6), ("~(TRUE /\\ FALSE) => TRUE", True),
])
def test_expressions(text, value):
    ev = make()
    ev.s = {}
    assert ev.ev(T.parse_expression(text), {}) == value


def test_junction_lists_end_where_the_column_says():
    e = T.parse_expression("/\\ 1 = 1\n/\\ \\/ 2 = 3\n   \\/ /\\ 4 = 4\n      /\\ 5 = 5\n/\\ 6 = 6")
    assert e[0] == "and" and len(e[1]) == 3 and e[1][1][0] == "or" and len(e[1][1][1]) == 2 and e[1][1][1][1][0] == "and"
    ev = make()
    ev.s = {}
    assert ev.ev(e, {}) is True


def test_errors_are_errors():
    ev = make()
    ev.s = {}
    with pytest.raises(T.EvalError, match="no field"):
        ev.ev(T.parse_expression("[a |-> 1].b"), {})
    with pytest.raises(T.EvalError, match="domain"):
        ev.ev(T.parse_expression("<<1, 2>>[3]"), {})
    with pytest.raises(T.EvalError, match="CHOOSE"):
        ev.ev(T.parse_expression("CHOOSE x \\in {1} : x > 1"), {})


def test_printer_round_trips_through_the_parser():
    text = "aux |-> (v1 :> TRUE @@ v2 :> FALSE),\nmsgs |-> ([type |-> PrepareMsg, dest |-> 2] :> 1),\nlog |-> <<<<>>, <<[op |-> v1]>>>>,\nreps |-> 1..3,\ns |-> {}"
    st = T.parse_state_record(text)
    assert T.fmt(st["reps"]) == "1..3" and T.fmt(st["log"]) == "<<<<>>, <<[op |-> v1]>>>>" and st["s"] == frozenset()
    again = T.parse_state_record(",\n".join("%s |-> %s" % (k, T.fmt(v)) for k, v in st.items()))
    assert again == st
