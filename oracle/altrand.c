/* oracle/altrand.c -- TEST INFRASTRUCTURE ONLY.
 *
 * LD_PRELOAD override of libc rand() for the tie-order probe.
 * The reference's small-bucket sort is a randomized, unstable Hoare quicksort
 * driven by the unseeded libc rand() (/root/reference/src/quicksort.c:7-14).
 * Running the reference a second time with a *different* rand() stream changes
 * the order of equal keys and nothing else; any checksum that differs between
 * the two runs is reference-undefined and is excluded from the parity set.
 */
#include <stdint.h>

static uint64_t state = 0x9E3779B97F4A7C15ull;

int rand(void)
{
    /* xorshift64*: different stream from glibc's TYPE_3 additive generator */
    state ^= state >> 12;
    state ^= state << 25;
    state ^= state >> 27;
    return (int)((state * 0x2545F4914F6CDD1Dull) >> 33) & 0x7fffffff;
}
