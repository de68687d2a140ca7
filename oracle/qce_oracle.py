"""oracle/qce_oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by the product).

CPU restatement (numpy) of the reference's per-query path: parser, predicate
arranger, filter operator, join operator with its mid-result state machine,
bystander re-join and projection checksums.  Every function cites the reference
code it restates as path:line under /root/reference.

Pinning: the reference ships no tests or golden vectors (no tests/ directory,
Makefile:11-12,44; workload files git-ignored, .gitignore:4-8).  This oracle is
pinned against *outputs of the reference itself*: `oracle/_ref/queries` is the
unmodified reference compiled by oracle/Makefile, and tests/test_oracle_vs_ref.py
+ tests/golden/*.json (made by tests/golden/make_golden.py) compare stdout
byte-for-byte on the parity-defined query class.

Tie order: the reference's sort is an unstable randomized quicksort below 4096
tuples (src/quicksort.c); this restatement uses a stable sort.  Results are
comparable exactly where the reference is tie-invariant (checked with the
alt-rand probe, oracle/altrand.c).
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

U64 = np.uint64

CLASSIC_JOIN, JOIN_SORT_LHS, JOIN_SORT_RHS, SCAN_JOIN, DO_NOTHING = 1, 2, 3, 4, 5  # src/join.h:15-19


class ReferenceAbort(Exception):
    """The reference would exit(EXIT_FAILURE) / crash here (stderr text in args)."""


# --------------------------------------------------------------------------- parsing
@dataclass
class Predicate:
    """src/structs.h:28-33.  type 0 = join (second = (binding, column)), 1 = filter."""
    type: int
    first: Tuple[int, int]
    second: object  # (binding, column) for joins, int constant for filters
    op: str

    def text(self) -> str:
        if self.type == 0:
            return f"{self.first[0]}.{self.first[1]}{self.op}{self.second[0]}.{self.second[1]}"
        return f"{self.first[0]}.{self.first[1]}{self.op}{self.second}"


@dataclass
class Query:
    relations: List[int]
    predicates: List[Predicate]
    selects: List[Tuple[int, int]]


_LINE = re.compile(r"^([0-9 ]+)\|+([0-9.=<>&]+)\|+([0-9. ]+)")
_JOIN = re.compile(r"^\s*([+-]?\d+)\.([+-]?\d+)(.)([+-]?\d+)\.([+-]?\d+)")
_FILT = re.compile(r"^\s*(\d+)\.(\d+)(.)(\d+)")


def parse_query(line: str) -> Optional[Query]:
    """parser()/parse_* of src/parsing.c:4-148: `rels|preds|selects`; lines that
    start with 'F' are batch separators and skipped (src/parsing.c:127)."""
    if line.startswith("F"):
        return None
    m = _LINE.match(line)
    if not m:
        return None
    rels = [int(t) for t in m.group(1).split(" ") if t != ""]
    preds: List[Predicate] = []
    for tok in m.group(2).split("&"):
        if tok == "":
            break
        j = _JOIN.match(tok)
        if j:  # 5-field "%d.%d%c%d.%d" (src/parsing.c:50)
            preds.append(Predicate(0, (int(j.group(1)), int(j.group(2))), (int(j.group(4)), int(j.group(5))), j.group(3)))
            continue
        f = _FILT.match(tok)
        if f:  # 4-field "%u.%u%c%u": constant is a uint32 (src/parsing.c:64-70)
            preds.append(Predicate(1, (int(f.group(1)), int(f.group(2))), int(f.group(4)) & 0xFFFFFFFF, f.group(3)))
    sels = []
    for tok in m.group(3).split(" "):
        if tok == "":
            break
        a, b = tok.split(".")
        sels.append((int(a), int(b)))
    return Query(rels, preds, sels)


def parse_batch(text: str) -> List[Query]:
    out = []
    for line in text.splitlines():
        q = parse_query(line)
        if q is not None:
            out.append(q)
    return out


# --------------------------------------------------------------------------- arranger
def _operands(p: Predicate):
    """The two (binding, column) operands is_match() compares (src/pred_arrange.c:29-48).
    A filter's `second` is a 4-byte constant that the reference reads as a 16-byte
    relation_column (heap over-read); with a zeroed heap tail that is
    (constant, 0), which is what is modelled here."""
    if p.type == 0:
        return p.first, p.second
    return p.first, (p.second, 0)


def _is_match(a: Predicate, b: Predicate) -> bool:
    a1, a2 = _operands(a)
    b1, b2 = _operands(b)
    return a1 == b1 or a1 == b2 or a2 == b1 or a2 == b2


def arrange_predicates(q: Query) -> None:
    """arrange_predicates = group_filters + group_matches, src/pred_arrange.c:50-93
    (including the off-by-one: position 0 is never examined by group_filters)."""
    p = q.predicates
    n = len(p)
    index = 0
    for i in range(1, n):  # group_filters, :70-86
        if p[i].type == 1:
            swaps = i
            for _ in range(i - index):
                p[swaps], p[swaps - 1] = p[swaps - 1], p[swaps]
                swaps -= 1
            index += 1
    i = index  # group_matches, :50-68
    while i < n - 1:
        cur_pos = i
        swapped = False
        for j in range(i + 1, n):
            # `current` is a pointer to slot i: a swap that moves slot i changes it
            if _is_match(p[cur_pos], p[j]):
                index += 1
                if index != j:
                    p[index], p[j] = p[j], p[index]
                swapped = True
        if swapped:
            i += index - i
        else:
            i += 1


# --------------------------------------------------------------------------- primitives
def _cmp(values: np.ndarray, op: str, c: int) -> np.ndarray:
    c = U64(c)
    if op == "=":
        return values == c
    if op == ">":
        return values > c
    if op == "<":
        return values < c
    raise ReferenceAbort("Wrong operator")  # src/filter.c:28,58


def filter_scan(col: np.ndarray, op: str, c: int) -> np.ndarray:
    """exec_filter_rel_no_exists, src/filter.c:37-64."""
    return np.nonzero(_cmp(col, op, c))[0].astype(U64)


def filter_refine(ids: np.ndarray, col: np.ndarray, op: str, c: int) -> np.ndarray:
    """exec_filter_rel_exists, src/filter.c:3-35 (order-preserving removal)."""
    return ids[_cmp(col[ids], op, c)]


def build_tuples(col: np.ndarray, ids: Optional[np.ndarray] = None):
    """allocate_relation (src/join.c:122-142) / allocate_relation_mid_results (:96-120)."""
    if ids is None:
        return col.copy(), np.arange(len(col), dtype=U64)
    return col[ids], ids.copy()


def sort_tuples(keys: np.ndarray, payloads: np.ndarray):
    """iterative_sort, src/join.c:5-94: ascending by key (stable here)."""
    order = np.argsort(keys, kind="stable")
    return keys[order], payloads[order]


def _is_sorted(keys: np.ndarray) -> bool:
    return len(keys) < 2 or bool(np.all(keys[:-1] <= keys[1:]))


def _merge_serial(kR, pR, kS, pS):
    """The literal pointer walk of join_relations, src/join.c:342-377 -- defined
    even when an input is not sorted.  Pure Python: small inputs only."""
    outR, outS = [], []
    pr, s_start, nR, nS = 0, 0, len(kR), len(kS)
    while pr < nR and s_start < nS:
        ps, flag = s_start, False
        while ps < nS:
            if kR[pr] < kS[ps]:
                break
            if kR[pr] > kS[ps]:
                ps += 1
                if not flag:
                    s_start = ps
            else:
                outR.append(pR[pr])
                outS.append(pS[ps])
                flag = True
                ps += 1
        pr += 1
    return np.array(outR, dtype=U64), np.array(outS, dtype=U64)


def merge_join(kR, pR, kS, pS):
    """join_relations, src/join.c:325-392: all matching pairs, R-major."""
    if not (_is_sorted(kR) and _is_sorted(kS)):
        return _merge_serial(kR, pR, kS, pS)
    lb = np.searchsorted(kS, kR, side="left")
    ub = np.searchsorted(kS, kR, side="right")
    cnt = (ub - lb).astype(np.int64)
    total = int(cnt.sum())
    outR = np.repeat(pR, cnt)
    if total == 0:
        return outR.astype(U64), np.zeros(0, dtype=U64)
    starts = np.repeat(lb, cnt)
    offs = np.arange(total, dtype=np.int64) - np.repeat(np.cumsum(cnt) - cnt, cnt)
    return outR.astype(U64), pS[starts + offs].astype(U64)


def distinct_pairs(outR: np.ndarray, outS: np.ndarray):
    """non_duplicates of join_relations (Hashmap dedup, src/join.c:358-367).  The
    reference keeps first-seen order; consumers sort, so sorted order is used."""
    if len(outR) == 0:
        return outR.copy(), outS.copy()
    pairs = np.unique(np.stack([outR, outS], axis=1), axis=0)
    return pairs[:, 0].copy(), pairs[:, 1].copy()


def scan_join(kR, pR, kS, pS):
    """scan_join, src/join.c:395-423: positional compare over min(nR,nS)."""
    n = min(len(kR), len(kS))
    m = kR[:n] == kS[:n]
    return pR[:n][m], pS[:n][m]


def join_payloads(driver: np.ndarray, last: np.ndarray, edit: np.ndarray) -> np.ndarray:
    """join_payloads, src/join.c:426-484."""
    n = len(last)
    if len(edit) < n:
        raise ReferenceAbort("join_payloads reads past the bystander column (src/join.c:433)")
    kR, pR = sort_tuples(last[:n].copy(), edit[:n].copy())
    kS = np.sort(driver, kind="stable")
    outR, _ = merge_join(kR, pR, kS, np.zeros(len(kS), dtype=U64))
    return outR


def checksum(col: np.ndarray, ids: np.ndarray) -> int:
    """print_sums inner loop, src/utilities.c:215-219 (wraps modulo 2^64)."""
    with np.errstate(over="ignore"):
        return int(np.sum(col[ids], dtype=U64))


# --------------------------------------------------------------------------- state machine
@dataclass
class MidResult:
    """src/structs.h:44-49."""
    relation: int
    binding: int
    last_column_sorted: int
    ids: np.ndarray


@dataclass
class QueryState:
    entities: List[List[MidResult]] = field(default_factory=list)
    stdout: List[str] = field(default_factory=list)


def _exists(st: QueryState, relation: int, binding: int):
    """relation_exists, src/utilities.c:164-181: newest entity first, first hit."""
    for e in range(len(st.entities) - 1, -1, -1):
        for j, mr in enumerate(st.entities[e]):
            if mr.relation == relation and mr.binding == binding:
                return e, j
    return None


def _exists_current(ent: List[MidResult], relation: int, binding: int) -> int:
    """relation_exists_current, src/utilities.c:183-194: last hit in one entity."""
    found = -1
    for j, mr in enumerate(ent):
        if mr.relation == relation and mr.binding == binding:
            found = j
    return found


def execute_filter(st: QueryState, pred: Predicate, relations: Sequence[int], db) -> None:
    """execute_filter, src/filter.c:66-100."""
    rel = relations[pred.first[0]]
    col = db[rel][pred.first[1]]
    if not st.entities:
        st.entities.append([])
    hit = _exists(st, rel, pred.first[0])
    if hit is not None:
        mr = st.entities[hit[0]][hit[1]]
        mr.ids = filter_refine(mr.ids, col, pred.op, pred.second)
        st.stdout.append(f"{len(mr.ids)}\n")  # src/filter.c:32
    else:
        st.entities[-1].append(MidResult(rel, pred.first[0], -1, filter_scan(col, pred.op, pred.second)))


def _fix_all(st: QueryState, res, hit, relR: int, relS: int, new: MidResult, mode: int) -> None:
    """fix_all_mid_results, src/join.c:486-505."""
    ent = st.entities[hit[0]]
    update = ent[hit[1]]
    no_dup = res["distinct"][mode]
    for edit in ent:
        if edit.relation != relR and edit.relation != relS:  # by relation id, not binding (:495)
            edit.ids = join_payloads(no_dup, update.ids, edit.ids)
    ent[hit[1]] = new


def execute_join(st: QueryState, pred: Predicate, relations: Sequence[int], db) -> None:
    """execute_join (src/join.c:630-679) with build_relations (:152-292) and
    update_mid_results (:507-628)."""
    lb, lcol = pred.first
    rb, rcol = pred.second
    lrel, rrel = relations[lb], relations[rb]
    if lrel == rrel and lcol == rcol:  # :161-163, bindings ignored
        return
    if not st.entities:
        st.entities.append([])
    cur = st.entities[-1]
    li = _exists_current(cur, lrel, lb)
    ri = _exists_current(cur, rrel, rb)

    def from_mid(mr: MidResult, rel: int, col: int):
        return build_tuples(db[rel][col], mr.ids)

    def from_base(rel: int, col: int):
        return build_tuples(db[rel][col])

    if li != -1 and ri == -1:  # :181-220
        L = from_mid(cur[li], lrel, lcol)
        hit = _exists(st, rrel, rb)
        if hit is None:
            R = from_base(rrel, rcol)
            if cur[li].last_column_sorted == lcol:
                kind = JOIN_SORT_RHS
            else:
                cur[li].last_column_sorted = lcol
                kind = CLASSIC_JOIN
        else:
            other = st.entities[hit[0]][hit[1]]
            R = from_mid(other, rrel, rcol)
            ls = cur[li].last_column_sorted == lcol
            rs = other.last_column_sorted == rcol
            kind = SCAN_JOIN if (ls and rs) else JOIN_SORT_RHS if ls else JOIN_SORT_LHS if rs else CLASSIC_JOIN
    elif li != -1 and ri != -1:  # :221-228
        L = from_mid(cur[li], lrel, lcol)
        R = from_mid(cur[ri], rrel, rcol)
        kind = SCAN_JOIN
    elif li == -1 and ri != -1:  # :229-269 (note the swapped constants)
        R = from_mid(cur[ri], rrel, rcol)
        hit = _exists(st, lrel, lb)
        if hit is None:
            L = from_base(lrel, lcol)
            if cur[ri].last_column_sorted == rcol:
                kind = JOIN_SORT_LHS
            else:
                cur[ri].last_column_sorted = rcol
                kind = CLASSIC_JOIN
        else:
            other = st.entities[hit[0]][hit[1]]
            L = from_mid(other, lrel, lcol)
            mid = cur[ri]
            if mid.last_column_sorted == rcol and other.last_column_sorted == lcol:
                kind = SCAN_JOIN
            elif mid.last_column_sorted == rcol:
                kind = JOIN_SORT_RHS  # :258-259
            elif mid.last_column_sorted == lcol:
                kind = JOIN_SORT_LHS  # :261-262
            else:
                kind = CLASSIC_JOIN
    else:  # :270-285
        st.entities.append([])
        R = from_base(rrel, rcol)
        L = from_base(lrel, lcol)
        kind = CLASSIC_JOIN if (rrel != lrel or lb != rb) else SCAN_JOIN

    # execute_join :639-663
    if kind == CLASSIC_JOIN:
        L, R = sort_tuples(*L), sort_tuples(*R)
    elif kind == JOIN_SORT_LHS:
        L = sort_tuples(*L)
    elif kind == JOIN_SORT_RHS:
        R = sort_tuples(*R)
    if kind == SCAN_JOIN:
        outL, outR = scan_join(L[0], L[1], R[0], R[1])
        res = {"out": (outL, outR), "distinct": (None, None)}
    else:
        outL, outR = merge_join(L[0], L[1], R[0], R[1])
        res = {"out": (outL, outR), "distinct": distinct_pairs(outL, outR)}

    # update_mid_results :507-628
    newL = MidResult(lrel, lb, lcol, outL)
    newR = MidResult(rrel, rb, rcol, outR)

    def push_or_fix(new: MidResult, mode: int, must_exist: bool, fix: bool):
        hit = _exists(st, new.relation, new.binding)
        if hit is None:
            if must_exist:
                raise ReferenceAbort("Something went really wrong")  # :563,:601
            st.entities[-1].append(new)
        elif fix:
            _fix_all(st, res, hit, lrel, rrel, new, mode)
        else:
            st.entities[hit[0]][hit[1]] = new  # :555-558, :587-590

    if kind == CLASSIC_JOIN:
        push_or_fix(newL, 0, False, True)
        push_or_fix(newR, 1, False, True)
    elif kind == JOIN_SORT_LHS:
        push_or_fix(newL, 0, False, False)
        push_or_fix(newR, 1, True, True)
    elif kind == JOIN_SORT_RHS:
        push_or_fix(newR, 1, False, False)
        push_or_fix(newL, 0, True, True)
    else:  # SCAN_JOIN :607-627: only the payload arrays are swapped in
        for new in (newL, newR):
            hit = _exists(st, new.relation, new.binding)
            if hit is None:
                raise ReferenceAbort("Something went really wrong")  # :610,:620
            st.entities[hit[0]][hit[1]].ids = new.ids


def execute_query(q: Query, db) -> str:
    """execute_query + print_sums, src/utilities.c:197-224,258-287.  Returns the
    bytes the reference writes to stdout for this query."""
    arrange_predicates(q)
    st = QueryState()
    for p in q.predicates:
        if p.type == 1:
            execute_filter(st, p, q.relations, db)
        else:
            execute_join(st, p, q.relations, db)
    out = []
    for b, c in q.selects:
        rel = q.relations[b]
        hit = _exists(st, rel, b)
        if hit is None:
            raise ReferenceAbort("Something went really wrong...")  # :204-207
        mr = st.entities[hit[0]][hit[1]]
        out.append("NULL " if len(mr.ids) == 0 else f"{checksum(db[rel][c], mr.ids)} ")
    return "".join(st.stdout) + "".join(out) + "\n"


def run_batch(db, text: str) -> str:
    """execute_queries, src/utilities.c:289-300."""
    return "".join(execute_query(q, db) for q in parse_batch(text))
