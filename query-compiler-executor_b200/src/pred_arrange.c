/* pred_arrange.c -- host-side predicate arranger; tiny and serial, stays on
 * the CPU (BASELINE.json north_star).
 *
 * Behavioural restatement of /root/reference/src/pred_arrange.c:29-93.  The
 * execution order it produces selects the join kinds downstream, so the two
 * quirks of the reference are kept on purpose:
 *   1. the filter pass starts at position 1, so a filter written first is
 *      never counted and later filters are rotated in front of it (:70-86);
 *   2. the grouping pass therefore can start on a filter, whose constant is
 *      then compared as if it were a (binding, column) operand (:32-43).  With
 *      the parser's zero-extended 16-byte constant block that operand is
 *      (constant, 0) -- the value the reference reads on a clean heap.
 * Golden orders produced by the reference's own object files:
 * tests/golden/arrange.json. */
#include "pred_arrange.h"

typedef struct operand_pair {
    relation_column a, b;
} operand_pair;

static operand_pair operands_of(const predicate *p)
{
    operand_pair o;
    o.a = p->first;
    o.b = *(const relation_column *)p->second; /* 16 readable bytes for both kinds */
    return o;
}

static int same_operand(relation_column x, relation_column y)
{
    return x.relation == y.relation && x.column == y.column;
}

static int shares_operand(const predicate *p, const predicate *q)
{
    operand_pair u = operands_of(p), v = operands_of(q);
    return same_operand(u.a, v.a) || same_operand(u.a, v.b) || same_operand(u.b, v.a) || same_operand(u.b, v.b);
}

static void exchange(predicate *list, ssize_t i, ssize_t j)
{
    if (i == j) return;
    predicate t = list[i];
    list[i] = list[j];
    list[j] = t;
}

/* Rotate every filter found at positions 1..n-1 down to the front block;
 * returns the size of that block (position 0 is never inspected). */
static ssize_t front_load_filters(predicate *list, ssize_t n)
{
    ssize_t front = 0;
    for (ssize_t i = 1; i < n; i++) {
        if (list[i].type != 1) continue;
        for (ssize_t k = i; k > front; k--) exchange(list, k, k - 1);
        front++;
    }
    return front;
}

/* From `start`, pull every later predicate that shares an operand with the
 * one in slot i to the slot after the grouped prefix (a swap, not a stable
 * move; the slot-i predicate itself can be displaced when prefix < i). */
static void group_shared_operands(predicate *list, ssize_t n, ssize_t start)
{
    ssize_t prefix = start;
    ssize_t i = start;
    while (i < n - 1) {
        int moved = 0;
        for (ssize_t j = i + 1; j < n; j++) {
            if (shares_operand(&list[i], &list[j])) {
                exchange(list, ++prefix, j);
                moved = 1;
            }
        }
        i = moved ? prefix : i + 1;
    }
}

void arrange_predicates(query *qry)
{
    ssize_t n = (ssize_t)qry->predicates_size;
    ssize_t front = front_load_filters(qry->predicates, n);
    group_shared_operands(qry->predicates, n, front);
}
