// qce_arena.hpp -- the engine's HBM arena: large slabs, a best-fit free list with coalescing on the host.
//
// Temporaries (tuple runs, ping-pong buffers, masks, look-back words, join scratch, row-id columns) are
// GB-sized and short-lived.  cudaMallocAsync's pool handled the steady single-query loop, but as soon as
// the pool ran tight it defragmented by remapping virtual ranges: single allocations took 40-340 ms.
// The engine therefore owns its memory.  All work of a context is ordered on one stream, so a block
// freed by the host may be handed out again immediately -- any kernel that still reads it was enqueued
// earlier.
//
// The slab source is injected (the engine passes cudaMalloc / cudaFree; tests/c/arena_test.cpp passes
// malloc / free and checks the free-list invariants on the CPU, shrink included).
#pragma once
#include <stdint.h>

#include <algorithm>
#include <map>
#include <set>
#include <utility>
#include <vector>

class QceArena {
  public:
    typedef unsigned long long u64;
    typedef void *(*SlabAlloc)(u64 bytes);   // nullptr when out of memory
    typedef void (*SlabFree)(void *);
    typedef void (*GrowHook)(const QceArena *self, u64 grew_by, u64 need, bool failed); // trace / error text
    static constexpr u64 kAlign = 512;
    u64 min_slab = 1ull << 30; // worker contexts (small queries) grow in smaller steps

    QceArena(SlabAlloc a, SlabFree f, GrowHook h = nullptr) : slab_alloc_(a), slab_free_(f), hook_(h) {}

    int alloc(void **out, u64 bytes)
    {
        bytes = (bytes + kAlign - 1) / kAlign * kAlign;
        if (bytes == 0) bytes = kAlign;
        auto it = by_size_.lower_bound({bytes, 0});
        if (it == by_size_.end()) {
            if (grow(bytes) != 0) return -1;
            it = by_size_.lower_bound({bytes, 0});
        }
        const u64 size = it->first, addr = it->second;
        by_size_.erase(it);
        by_addr_.erase(addr);
        if (size > bytes) insert_free(addr + bytes, size - bytes);
        live_[addr] = bytes;
        used_ += bytes;
        *out = (void *)addr;
        return 0;
    }
    void free(void *p)
    {
        if (!p) return;
        auto it = live_.find((u64)p);
        if (it == live_.end()) return; // not ours (adopted buffer)
        u64 addr = it->first, size = it->second;
        live_.erase(it);
        used_ -= size;
        // coalesce with the free neighbours (never across slab boundaries)
        auto next = by_addr_.find(addr + size);
        if (next != by_addr_.end() && !slab_starts_.count(addr + size)) {
            size += next->second;
            by_size_.erase({next->second, next->first});
            by_addr_.erase(next);
        }
        auto prev = by_addr_.lower_bound(addr);
        if (prev != by_addr_.begin()) {
            --prev;
            if (prev->first + prev->second == addr && !slab_starts_.count(addr)) {
                addr = prev->first;
                size += prev->second;
                by_size_.erase({prev->second, prev->first});
                by_addr_.erase(prev);
            }
        }
        insert_free(addr, size);
    }
    // give the tail of a live block back (an output sized by a guess, once its real size is known)
    void shrink(void *p, u64 bytes)
    {
        auto it = live_.find((u64)p);
        if (it == live_.end()) return;
        bytes = (bytes + kAlign - 1) / kAlign * kAlign;
        if (bytes == 0) bytes = kAlign;
        if (bytes >= it->second) return;
        u64 addr = it->first + bytes, size = it->second - bytes;
        used_ -= size;
        it->second = bytes;
        auto next = by_addr_.find(addr + size);
        if (next != by_addr_.end() && !slab_starts_.count(addr + size)) {
            size += next->second;
            by_size_.erase({next->second, next->first});
            by_addr_.erase(next);
        }
        insert_free(addr, size);
    }
    void release_all()
    {
        for (auto &s : slabs_) slab_free_((void *)s.first);
        slabs_.clear(); slab_starts_.clear(); by_addr_.clear(); by_size_.clear(); live_.clear();
        reserved_ = used_ = 0;
    }
    u64 reserved() const { return reserved_; }
    u64 used() const { return used_; }
    // introspection for the tests: live blocks (addr -> size), free blocks (addr -> size), slabs (addr, size)
    const std::map<u64, u64> &live_blocks() const { return live_; }
    const std::map<u64, u64> &free_blocks() const { return by_addr_; }
    const std::vector<std::pair<u64, u64>> &slabs() const { return slabs_; }

  private:
    int grow(u64 need)
    {
        // at least as much again as is already reserved, so the slab count stays small
        u64 bytes = std::max<u64>(std::max<u64>(need, min_slab), reserved_);
        void *p = slab_alloc_(bytes);
        if (!p && bytes > need) {
            bytes = need;
            p = slab_alloc_(bytes);
        }
        if (!p) {
            if (hook_) hook_(this, 0, need, true);
            return -1;
        }
        if (hook_) hook_(this, bytes, need, false);
        slabs_.push_back({(u64)p, bytes});
        slab_starts_.insert((u64)p);
        reserved_ += bytes;
        insert_free((u64)p, bytes);
        return 0;
    }
    void insert_free(u64 addr, u64 size)
    {
        by_addr_[addr] = size;
        by_size_.insert({size, addr});
    }
    SlabAlloc slab_alloc_;
    SlabFree slab_free_;
    GrowHook hook_;
    std::vector<std::pair<u64, u64>> slabs_;
    std::set<u64> slab_starts_;
    std::map<u64, u64> by_addr_;            // free blocks: addr -> size
    std::set<std::pair<u64, u64>> by_size_; // free blocks: (size, addr)
    std::map<u64, u64> live_;               // handed out: addr -> size
    u64 reserved_ = 0, used_ = 0;
};
