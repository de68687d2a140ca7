/* join.h -- the join operator (entry point of /root/reference/src/join.h:26). */
#ifndef QCE_JOIN_H
#define QCE_JOIN_H

#include <stdint.h>
#include <stdio.h>

#include "DArray.h"
#include "structs.h"
#include "utilities.h"

/* how a join predicate is executed (values of src/join.h:15-19) */
#define CLASSIC_JOIN 1  /* sort both sides, merge            */
#define JOIN_SORT_LHS 2 /* rhs already in key order          */
#define JOIN_SORT_RHS 3 /* lhs already in key order          */
#define SCAN_JOIN 4     /* positional compare of two columns */
#define DO_NOTHING 5    /* same relation id and same column  */

struct qce_rowids;

/* device-side result of one join: aligned row-id columns for the two sides and,
 * computed on demand, the two sides of the distinct (rowid,rowid) pairs */
typedef struct result {
    struct qce_rowids *results[2];
    struct qce_rowids *non_duplicates[2];
} join_result;

/* Bystander re-join elision (join.c): the scheduler runs a query in elided mode
 * first when its shape allows it and replays it faithfully when the join operator reports that
 * elided mode can no longer vouch for the reference's result. */
#define QCE_JOIN_UNSAFE (-2)
void qce_join_elide_begin(int on);
int qce_join_elide_unsafe(void);
int qce_join_elide_active(void);
void qce_join_elide_raise(void);
int qce_entity_was_rejoined(const DArray *entity);

/* Applies `a.x = b.y` to the query's entity list.  Returns 0, or -1 on error
 * (including a merge that would run over unsorted input, which the reference
 * executes with undefined results). */
int execute_join(predicate *pred, uint32_t *relations, DArray *metadata_arr, DArray *mid_results);

#endif /* QCE_JOIN_H */
