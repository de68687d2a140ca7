/* tests/c/host_probe.c -- prints the predicate order the host layer's own
 * parser + arranger produce (same output format as oracle/arrange_probe.c, which
 * links the reference's objects).  No GPU needed. */
#include <stdio.h>

#include "DArray.h"
#include "parsing.h"
#include "pred_arrange.h"
#include "structs.h"

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
                printf("%lu.%lu%c%lu.%lu", p->first.relation, p->first.column, p->operator, rc->relation, rc->column);
            } else {
                printf("%lu.%lu%c%lu", p->first.relation, p->first.column, p->operator, *(uint64_t *)p->second);
            }
            putchar(j + 1 < q->predicates_size ? ' ' : '\n');
        }
        printf("#rels=%zu sels=%zu", q->relations_size, q->select_size);
        for (size_t j = 0; j < q->relations_size; j++) printf(" r%u", q->relations[j]);
        for (size_t j = 0; j < q->select_size; j++) printf(" s%lu.%lu", q->selects[j].relation, q->selects[j].column);
        putchar('\n');
    }
    return 0;
}
