/* pred_arrange.h -- predicate reordering (entry point of
 * /root/reference/src/pred_arrange.h:14). */
#ifndef QCE_PRED_ARRANGE_H
#define QCE_PRED_ARRANGE_H

#include "structs.h"

/* Reorders q->predicates in place: filters first, then joins that share an
 * operand made adjacent.  The order decides which join kinds fire, so this is
 * behaviour-equivalent to the reference, quirks included (see the .c file). */
void arrange_predicates(query *qry);

#endif /* QCE_PRED_ARRANGE_H */
