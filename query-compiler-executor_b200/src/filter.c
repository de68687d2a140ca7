/* filter.c -- filter operator on the GPU.
 *
 * Same contract as /root/reference/src/filter.c:66-100; the two loops it
 * dispatches to are kernels behind the C-ABI:
 *   exec_filter_rel_no_exists (:37-64)  -> qce_filter_scan
 *   exec_filter_rel_exists    (:3-35)   -> qce_filter_refine (+ the count line
 *                                          of :32, printed here)
 */
#include "filter.h"

#include "../../include/qce_b200.h"
#include "join.h"

static __thread FILE *g_query_out = NULL; /* one query per host thread (schedule.c) */

void qce_set_query_stdout(FILE *out) { g_query_out = out; }
FILE *qce_query_stdout(void) { return g_query_out ? g_query_out : stdout; }

static DArray *last_entity(DArray *entities)
{
    if (DArray_count(entities) == 0) {
        DArray *fresh = DArray_create(sizeof(mid_result), 4);
        if (fresh == NULL) return NULL;
        DArray_push(entities, &fresh);
    }
    return *(DArray **)DArray_last(entities);
}

int execute_filter(predicate *pred, uint32_t *relations, DArray *metadata_arr, DArray *mid_results_array)
{
    const uint64_t binding = pred->first.relation;
    const uint32_t rel = relations[binding];
    const uint32_t col = (uint32_t)pred->first.column;
    const uint64_t constant = *(const uint64_t *)pred->second;
    (void)metadata_arr; /* columns are addressed by (relation, column) on the device */

    DArray *current = last_entity(mid_results_array);
    check_mem(current);

    exists_info where = relation_exists(mid_results_array, rel, binding);
    if (where.index != -1) {
        DArray *entity = *(DArray **)DArray_get(mid_results_array, where.mid_result);
        mid_result *entry = (mid_result *)DArray_get(entity, where.index);
        uint64_t survivors = 0;
        if (qce_join_elide_active() && DArray_count(entity) > 1) {
            /* a refine narrows ONE column of the entity (src/filter.c:3-35): its siblings keep their length,
             * and what the reference does with such an entity afterwards depends on their order */
            qce_join_elide_raise();
            return -1;
        }
        check(qce_filter_refine(entry->payloads, rel, col, pred->operator, constant, &survivors) == 0,
              "Execution of filter failed! %s", qce_last_error());
        fprintf(qce_query_stdout(), "%d\n", (int)survivors);
    } else {
        mid_result entry;
        entry.relation = rel;
        entry.predicate_id = binding;
        entry.last_column_sorted = -1;
        entry.payloads = NULL;
        check(qce_filter_scan(rel, col, pred->operator, constant, &entry.payloads) == 0,
              "Execution of filter failed! %s", qce_last_error());
        DArray_push(current, &entry);
    }
    return 0;

error:
    return -1;
}
