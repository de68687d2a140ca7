// CPU test of the engine's HBM arena (csrc/qce_arena.hpp) with malloc-backed slabs: random
// alloc / free / shrink traffic, the free-list invariants after every step.  Test infrastructure.
#include <stdio.h>
#include <stdlib.h>

#include <random>
#include <vector>

#include "../../query-compiler-executor_b200/csrc/qce_arena.hpp"

typedef QceArena::u64 u64;
static int g_slabs = 0;
static void *slab_alloc(u64 bytes) { g_slabs++; return aligned_alloc(512, (size_t)bytes); }
static void slab_free(void *p) { free(p); }

#define CHECK(c)                                                              \
    do {                                                                      \
        if (!(c)) { fprintf(stderr, "arena_test:%d: %s\n", __LINE__, #c); exit(1); } \
    } while (0)

// live and free blocks tile every slab exactly; no two free blocks touch inside a slab; used() = sum of live
static void check_invariants(const QceArena &a)
{
    u64 live = 0, freeb = 0, slabs = 0;
    for (auto &s : a.slabs()) {
        slabs += s.second;
        u64 at = s.first;
        bool prev_free = false;
        while (at < s.first + s.second) {
            auto l = a.live_blocks().find(at);
            auto f = a.free_blocks().find(at);
            CHECK((l != a.live_blocks().end()) != (f != a.free_blocks().end())); // exactly one of them starts here
            if (l != a.live_blocks().end()) { live += l->second; at += l->second; prev_free = false; }
            else { CHECK(!prev_free); freeb += f->second; at += f->second; prev_free = true; }
        }
        CHECK(at == s.first + s.second);
    }
    CHECK(live == a.used());
    CHECK(live + freeb == slabs && slabs == a.reserved());
}

int main()
{
    QceArena a(slab_alloc, slab_free);
    a.min_slab = 1 << 20;
    std::mt19937_64 rng(7);
    std::vector<std::pair<void *, u64>> held;
    for (int step = 0; step < 20000; step++) {
        const int op = (int)(rng() % 10);
        if (op < 5 || held.empty()) {
            const u64 bytes = rng() % 3 == 0 ? rng() % (3 << 20) : rng() % 40000;
            void *p = nullptr;
            CHECK(a.alloc(&p, bytes) == 0 && p != nullptr && ((u64)p % 512) == 0);
            held.push_back({p, bytes});
        } else if (op < 8) {
            const size_t k = rng() % held.size();
            a.free(held[k].first);
            held[k] = held.back();
            held.pop_back();
        } else {
            const size_t k = rng() % held.size();
            const u64 to = held[k].second ? rng() % (held[k].second + 1) : 0; // anything from 0 to the current size
            a.shrink(held[k].first, to);
            held[k].second = to;
        }
        if (step % 64 == 0) check_invariants(a);
    }
    check_invariants(a);
    // a block never shrinks below what its owner still uses, and growing requests are ignored
    void *p = nullptr;
    CHECK(a.alloc(&p, 10000) == 0);
    const u64 before = a.used();
    a.shrink(p, 20000);
    CHECK(a.used() == before);
    a.shrink(p, 600);
    CHECK(a.used() == before - (10240 - 1024));
    a.free(p);
    a.free(p);       // twice: the second call finds nothing
    a.shrink(p, 0);  // not live any more
    for (auto &h : held) a.free(h.first);
    check_invariants(a);
    CHECK(a.used() == 0);
    // everything came back and coalesced: every slab is one free block again, so the largest slab can be
    // handed out whole without growing
    CHECK(a.free_blocks().size() == a.slabs().size());
    u64 largest = 0;
    for (auto &s : a.slabs()) largest = std::max(largest, s.second);
    const int slabs_before = g_slabs;
    CHECK(a.alloc(&p, largest) == 0 && g_slabs == slabs_before);
    a.free(p);
    // the steady loop of the single-pass join: two outputs sized by a guess, trimmed, other work, all freed --
    // the same addresses every step, no growth
    void *first_r = nullptr;
    for (int step = 0; step < 50; step++) {
        void *r, *s, *t;
        CHECK(a.alloc(&r, 250000) == 0 && a.alloc(&s, 250000) == 0);
        a.shrink(r, 225000); a.shrink(s, 225000);
        CHECK(a.alloc(&t, 20000) == 0);
        if (step == 0) first_r = r;
        CHECK(r == first_r);
        a.free(t); a.free(r); a.free(s);
        CHECK(a.used() == 0 && g_slabs == slabs_before);
    }
    a.release_all();
    CHECK(a.reserved() == 0);
    printf("arena ok\n");
    return 0;
}
