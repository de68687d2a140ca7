/* parsing.c -- host-side query parser.
 *
 * Behavioural restatement of /root/reference/src/parsing.c:4-148, written as a
 * hand tokenizer instead of nested sscanf scansets:
 *   - a line is `relations|predicates|selects`; the relation section accepts
 *     [0-9 ], the predicate section [0-9.=<>&], the select section [0-9. ]
 *     (scansets of src/parsing.c:132); sections are at most 127 characters;
 *   - lines starting with 'F' separate batches and are skipped (:127);
 *   - a predicate `a.b OP c.d` is a join (type 0), `a.b OP k` a filter
 *     (type 1) whose constant is parsed as a uint32 (:64-70);
 *   - the element counts are (separators + 1), as in the reference (:6-13,
 *     :32-39, :92-99).
 * One deliberate difference: a filter constant is stored zero-extended in a
 * 16-byte zeroed block.  The reference stores 4 bytes and later reads 8
 * (src/filter.c:70) or 16 (src/pred_arrange.c:32) of them -- a heap over-read
 * whose result is the zero-extended value on a clean heap; here that result is
 * guaranteed. */
#define _GNU_SOURCE
#include "parsing.h"

#include <ctype.h>

#define SECTION_MAX 127

static size_t span_of(const char *s, const char *accept, char *out)
{
    size_t n = strspn(s, accept);
    size_t keep = n < SECTION_MAX ? n : SECTION_MAX;
    memcpy(out, s, keep);
    out[keep] = '\0';
    return n;
}

static size_t count_char(const char *s, char c)
{
    size_t n = 0;
    for (; *s; s++) n += (*s == c);
    return n;
}

/* unsigned decimal; *end is set past the digits; returns 0 if none */
static int read_number(const char *s, const char **end, uint64_t *value)
{
    if (!isdigit((unsigned char)*s)) return 0;
    uint64_t v = 0;
    while (isdigit((unsigned char)*s)) v = v * 10 + (uint64_t)(*s++ - '0');
    *value = v;
    *end = s;
    return 1;
}

static void parse_relation_list(const char *text, query *q)
{
    size_t slots = count_char(text, ' ') + 1;
    q->relations = CALLOC(slots, sizeof(uint32_t), uint32_t);
    q->relations_size = slots;
    size_t i = 0;
    const char *p = text;
    uint64_t v;
    while (i < slots && read_number(p, &p, &v)) {
        q->relations[i++] = (uint32_t)v;
        if (*p != ' ') break;
        p++;
    }
}

/* returns the number of well-formed predicates (== q->predicates_size for a valid section) */
static size_t parse_predicate_list(const char *text, query *q)
{
    size_t slots = count_char(text, '&') + 1;
    q->predicates = CALLOC(slots, sizeof(predicate), predicate);
    q->predicates_size = slots;
    size_t i = 0;
    const char *p = text;
    while (i < slots && *p) {
        const char *tok_end = strchr(p, '&');
        if (tok_end == NULL) tok_end = p + strlen(p);
        uint64_t a, b, c, d;
        const char *s = p;
        if (read_number(s, &s, &a) && *s == '.' && read_number(s + 1, &s, &b) && s < tok_end) {
            char op = *s++;
            if (read_number(s, &s, &c)) {
                predicate *pr = &q->predicates[i];
                pr->first.relation = a;
                pr->first.column = b;
                pr->operator = op;
                if (*s == '.' && read_number(s + 1, &s, &d)) {
                    relation_column *rc = CALLOC(1, sizeof(relation_column), relation_column);
                    rc->relation = c;
                    rc->column = d;
                    pr->type = 0;
                    pr->second = rc;
                } else {
                    /* 16 zeroed bytes, constant truncated to uint32 like "%u" */
                    uint64_t *k = CALLOC(2, sizeof(uint64_t), uint64_t);
                    k[0] = (uint32_t)c;
                    pr->type = 1;
                    pr->second = k;
                }
                i++;
            }
        }
        if (*tok_end != '&') break;
        p = tok_end + 1;
    }
    return i;
}

static void free_query_lists(query *q)
{
    for (size_t j = 0; j < q->predicates_size; j++) FREE(q->predicates[j].second);
    FREE(q->predicates);
    FREE(q->relations);
    FREE(q->selects);
    memset(q, 0, sizeof *q);
}

static void parse_select_list(const char *text, query *q)
{
    size_t slots = count_char(text, ' ') + 1;
    q->selects = CALLOC(slots, sizeof(relation_column), relation_column);
    q->select_size = slots;
    size_t i = 0;
    const char *p = text;
    uint64_t a, b;
    while (i < slots && read_number(p, &p, &a)) {
        b = 0;
        if (*p == '.') read_number(p + 1, &p, &b);
        q->selects[i].relation = a;
        q->selects[i].column = b;
        i++;
        if (*p != ' ') break;
        p++;
    }
}

int parse_query_line(const char *line, query *q)
{
    char rels[SECTION_MAX + 1], preds[SECTION_MAX + 1], sels[SECTION_MAX + 1];
    const char *p = line;
    size_t n = span_of(p, "0123456789 ", rels);
    if (n == 0) return -1;
    p += n;
    n = strspn(p, "|");
    if (n == 0) return -1;
    p += n;
    n = span_of(p, "0123456789.=<>&", preds);
    if (n == 0) return -1;
    p += n;
    n = strspn(p, "|");
    if (n == 0) return -1;
    p += n;
    n = span_of(p, "0123456789. ", sels);
    if (n == 0) return -1;
    memset(q, 0, sizeof *q);
    parse_relation_list(rels, q);
    const size_t parsed = parse_predicate_list(preds, q);
    parse_select_list(sels, q);
    /* Hardening: the reference sizes its arrays by separator counts and then executes
     * whatever sscanf left in them (src/parsing.c:32-39, :92-99).  A predicate section with an
     * unparsable element, or a binding index outside the relation list, is refused here with a
     * diagnostic instead of reaching the operators with a NULL operand / an out-of-range index. */
    int ok = parsed == q->predicates_size;
    for (size_t j = 0; ok && j < q->predicates_size; j++) {
        const predicate *pr = &q->predicates[j];
        if (pr->first.relation >= q->relations_size) ok = 0;
        if (pr->type == 0 && ((const relation_column *)pr->second)->relation >= q->relations_size) ok = 0;
    }
    for (size_t j = 0; ok && j < q->select_size; j++)
        if (q->selects[j].relation >= q->relations_size) ok = 0;
    if (!ok) {
        fprintf(stderr, "[ERROR] malformed query line skipped: %.*s\n", (int)strcspn(line, "\n"), line);
        free_query_lists(q);
        return -1;
    }
    return 0;
}

DArray *parser_from(FILE *in)
{
    DArray *queries = DArray_create(sizeof(query), 10);
    if (queries == NULL) return NULL;
    char *line = NULL;
    size_t cap = 0;
    while (getline(&line, &cap, in) != -1) {
        if (line[0] == 'F') continue;
        query q;
        if (parse_query_line(line, &q) != 0) continue; /* blank / malformed line */
        DArray_push(queries, &q);
    }
    free(line);
    return queries;
}

DArray *parser(void) { return parser_from(stdin); }
