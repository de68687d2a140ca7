/* filter.h -- the filter operator (entry point of
 * /root/reference/src/filter.h:14). */
#ifndef QCE_FILTER_H
#define QCE_FILTER_H

#include <stdint.h>
#include <stdio.h>

#include "DArray.h"
#include "structs.h"
#include "utilities.h"

/* Applies `binding.column OP constant`.  A binding seen for the first time gets
 * a new mid_result in the LAST entity holding every matching row id
 * (ascending); a binding that already has a mid_result anywhere has its row-id
 * column narrowed in place and the surviving count is printed to stdout.
 * Returns 0, or -1 on error.  `relations` maps binding -> relation id;
 * `mid_results` is the query's entity list (DArray of DArray* of mid_result). */
int execute_filter(predicate *pred, uint32_t *relations, DArray *metadata_arr, DArray *mid_results);

/* stdout of the running query (count lines must precede its result line) */
void qce_set_query_stdout(FILE *out);
FILE *qce_query_stdout(void);

#endif /* QCE_FILTER_H */
