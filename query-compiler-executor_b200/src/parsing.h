/* parsing.h -- query text -> struct query (entry point of
 * /root/reference/src/parsing.h:10). */
#ifndef QCE_PARSING_H
#define QCE_PARSING_H

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "DArray.h"
#include "structs.h"

/* Reads every remaining line of stdin.  Returns a DArray of `query`; all
 * arrays inside a query are malloc-family blocks owned by the caller (the
 * driver frees relations / predicates[j].second / predicates / selects). */
DArray *parser(void);

/* Same, from any stream (used by the driver's batch mode and by tests). */
DArray *parser_from(FILE *in);

/* One line `r r r|p&p&p|s s` -> *q.  Returns 0, or -1 if the line does not
 * have the three sections. */
int parse_query_line(const char *line, query *q);

#endif /* QCE_PARSING_H */
