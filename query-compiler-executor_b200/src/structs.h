/* structs.h -- host-side data model of the query path.
 *
 * The layouts that the driver touches (metadata / relation / tuple / query /
 * predicate / relation_column) are those of /root/reference/src/structs.h:7-42,
 * so the reference's main/queries_main.c links and frees them unchanged.
 * What differs is where the data lives:
 *   - relation.tuples is always NULL: the reference copies every column into
 *     an AoS tuple{key,payload=i} array (src/utilities.c:111-120); here the
 *     column is uploaded once to HBM as SoA uint64 and found by (relation,
 *     column) index through the engine (qce_upload_column).
 *   - mid_result.payloads is a device row-id column handle, not a DArray of
 *     calloc'ed uint64_t. */
#ifndef QCE_STRUCTS_H
#define QCE_STRUCTS_H

#include <stddef.h>
#include <stdint.h>

#include "DArray.h"

struct qce_rowids; /* include/qce_b200.h */

/* one (key, row id) pair; only the type survives on the host */
typedef struct tuple {
    uint64_t key;
    uint64_t payload;
} tuple;

/* one column of one relation */
typedef struct relation {
    tuple *tuples;       /* NULL: data is resident on the GPU */
    uint64_t num_tuples; /* rows */
} relation;

/* one relation file */
typedef struct metadata {
    uint64_t tuples;  /* rows */
    uint64_t columns; /* columns */
    relation **data;  /* data[c] describes column c */
} metadata;

/* `binding.column`: `relation` is the index into the query's relation list */
typedef struct relation_column {
    uint64_t relation;
    uint64_t column;
} relation_column;

/* type 0: first = second (second is a relation_column*)
 * type 1: first OP constant (second points at the uint32 constant, stored
 *         zero-extended in a 16-byte block, see parsing.c) */
typedef struct predicate {
    int8_t type;
    relation_column first;
    void *second;
    char operator;
} predicate;

typedef struct query {
    uint32_t *relations;
    size_t relations_size;
    predicate *predicates;
    size_t predicates_size;
    relation_column *selects;
    size_t select_size;
} query;

/* one binding's surviving row ids inside an entity of joined bindings */
typedef struct mid_result {
    uint64_t relation;          /* relation id */
    uint64_t predicate_id;      /* binding index inside the query */
    int32_t last_column_sorted; /* column whose key order the ids are in, or -1 */
    struct qce_rowids *payloads;/* device row-id column (owned) */
} mid_result;

#endif /* QCE_STRUCTS_H */
