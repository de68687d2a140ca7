/* oracle/arrange_probe.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Links the reference's own parser (/root/reference/src/parsing.c:118) and
 * predicate arranger (/root/reference/src/pred_arrange.c:88) and prints, for
 * every query line on stdin, the order in which the reference would execute
 * the predicates.  Used to pin the host-side re-implementation
 * (query-compiler-executor_b200/src/pred_arrange.c) and the numpy oracle
 * against the real thing (tests/test_host_logic.py, tests/golden/arrange.json).
 *
 * Output: one line per query, predicates separated by a single space, e.g.
 *   0.2<100 0.1=1.1 1.2=2.1
 */
#include <stdio.h>
#include "structs.h"
#include "parsing.h"
#include "pred_arrange.h"

int main(void)
{
    DArray *queries = parser();
    for (size_t i = 0; i < DArray_count(queries); i++) {
        query *q = (query *)DArray_get(queries, i);
        arrange_predicates(q);
        for (size_t j = 0; j < q->predicates_size; j++) {
            predicate *p = &q->predicates[j];
            if (p->type == 0) {
                relation_column *rc = (relation_column *)p->second;
                printf("%lu.%lu%c%lu.%lu", p->first.relation, p->first.column, p->operator,
                       rc->relation, rc->column);
            } else {
                printf("%lu.%lu%c%u", p->first.relation, p->first.column, p->operator,
                       *(uint32_t *)p->second);
            }
            putchar(j + 1 < q->predicates_size ? ' ' : '\n');
        }
    }
    return 0;
}
