// gx_table.cuh -- lock-free open-addressing hash table keyed on packed canonical k-mers.
//
// Replaces the reference's "external sort + pre-clustered group" pair
// (hyracks-dataflow-std/.../sort/ExternalSortOperatorDescriptor.java:119-194 and
// .../group/preclustered/PreclusteredGroupWriter.java:76-136 driving
// genomix-hyracks/.../AggregateKmerAggregateFactory.java:93-144): equal keys meet in one slot, the
// aggregate (coverage sum, edge-set union) is folded in with atomics.
//
// Protocols (linear probing, capacity never above the load limit so a free slot always exists):
//   KW=1  64-bit atomicCAS on the key word;      value folded with atomicAdd (+ rare atomicOr)
//   KW=2  128-bit atomicCAS on the key pair;     same
//   KW>=3 claim the slot by CAS(val: 0 -> LOCK), write key words, publish val with release order;
//         readers load val with acquire order before they compare key words.
#pragma once
#include "gx_internal.cuh"

namespace gx {

__device__ __forceinline__ u64 ld_acquire(const u64* p) {
    u64 v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(u64* p, u64 v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ u64 ld_relaxed(const u64* p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ u32 ld_relaxed_u32(const u32* p) {
    u32 v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ld_relaxed_v2(const u64* p, u64& a, u64& b) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void ld_relaxed_v4(const u64* p, u64& a, u64& b, u64& c, u64& d) {
    // two 16-byte halves of one 32-byte sector, issued back to back
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%4];\n\tld.relaxed.gpu.global.v2.u64 {%2, %3}, [%4+16];"
                 : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p) : "memory");
}
__device__ __forceinline__ void st_relaxed(u64* p, u64 v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ bool cas128(u64* addr, u64 c0, u64 c1, u64 n0, u64 n1, u64& o0, u64& o1) {
    asm volatile(
        "{\n\t.reg .b128 c, n, o;\n\t"
        "mov.b128 c, {%2, %3};\n\t"
        "mov.b128 n, {%4, %5};\n\t"
        "atom.relaxed.gpu.global.cas.b128 o, [%6], c, n;\n\t"
        "mov.b128 {%0, %1}, o;\n\t}"
        : "=l"(o0), "=l"(o1)
        : "l"(c0), "l"(c1), "l"(n0), "l"(n1), "l"(addr)
        : "memory");
    return o0 == c0 && o1 == c1;
}

// Fold `add` (count in the low 48 bits) and `mask` (16 edge bits) into an occupied slot's value word.
// `seen` is any earlier snapshot of that value word: edge bits only ever get set, so bits already present in
// the snapshot need no atomicOr. Neither atomic's result is used, so both compile to fire-and-forget RED.
__device__ __forceinline__ void fold_value(u64* val, u64 add, u32 mask, u64 seen) {
    atomicAdd(val, add);
    const u64 m = (u64)mask << MASK_SHIFT;
    if ((seen & m) != m) atomicOr(val, m);
}

// Insert-or-aggregate. Returns the slot index; *is_new set when this call created the slot.
// A probe budget bounds the work of one upsert: the host sizes batches so that the load stays near GROW_LOAD, where
// probe sequences are a handful of slots; if its prediction of new keys was too low and the table fills up, the
// budget runs out, `capacity` is returned and the caller spills the record (spill_record) for the host to insert
// after growing the table. Never a hang, never a lost occurrence.
static constexpr u64 PROBE_BUDGET = 1ull << 14;

// What the first probe of an upsert reads from the home slot: the key words and the value word (KW <= 2) or the value
// word with acquire order (KW >= 3). Loading it ahead of time (probe_load) lets a thread keep the home-slot loads of
// several records in flight before it folds the first one.
template <int KW> struct Probe { u64 w[KW <= 2 ? KW + 1 : 1]; };

template <int KW>
__device__ __forceinline__ void probe_load(const u64* s, Probe<KW>& p) {
    if constexpr (KW == 1) {
        ld_relaxed_v2(s, p.w[0], p.w[1]);
    } else if constexpr (KW == 2) {
        u64 pad;
        ld_relaxed_v4(s, p.w[0], p.w[1], p.w[2], pad);
    } else {
        p.w[0] = ld_acquire(s + KW);
    }
}

// `hl` is the key's local hash (local_hash(hash_key(key), n_ranks)); the home slot is slot_of(hl, capacity).
// `first`, if given, holds what probe_load read from the home slot.
template <int KW>
__device__ __forceinline__ u64 table_upsert(u64* __restrict__ table, u64 capacity, u64 hl, const u64 (&key)[KW], u64 add,
                                            u32 mask, bool& is_new, const Probe<KW>* first = nullptr) {
    constexpr int SW = SlotTraits<KW>::WORDS;
    u64 slot = slot_of(hl, capacity);
    is_new = false;
    u64 budget = PROBE_BUDGET;
    bool pre = first != nullptr;
    if constexpr (KW == 1) {
        for (; budget; --budget, pre = false) {
            u64* s = table + slot * SW;
            u64 cur, seen;
            if (pre) { cur = first->w[0]; seen = first->w[1]; }
            else ld_relaxed_v2(s, cur, seen);  // key and value word: one 16-byte request
            if (cur == EMPTY_WORD) {
                // claim the slot and store the first occurrence's value with ONE 16-byte CAS (an empty slot's value is 0)
                u64 o0, o1;
                if (cas128(s, EMPTY_WORD, 0ull, key[0], add | ((u64)mask << MASK_SHIFT), o0, o1)) { is_new = true; return slot; }
                cur = o0; seen = o1;
            }
            if (cur == key[0]) { fold_value(s + 1, add, mask, seen); return slot; }
            if (++slot == capacity) slot = 0;
        }
    } else if constexpr (KW == 2) {
        for (; budget; --budget, pre = false) {
            u64* s = table + slot * SW;
            u64 c0, c1, seen, pad;
            if (pre) { c0 = first->w[0]; c1 = first->w[1]; seen = first->w[2]; }
            else ld_relaxed_v4(s, c0, c1, seen, pad);  // the whole 32-byte slot (one sector)
            if (c0 == EMPTY_WORD && c1 == EMPTY_WORD) {
                if (cas128(s, EMPTY_WORD, EMPTY_WORD, key[0], key[1], c0, c1)) { is_new = true; c0 = key[0]; c1 = key[1]; }
                seen = 0;
            }
            if (c0 == key[0] && c1 == key[1]) { fold_value(s + 2, add, mask, seen); return slot; }
            if (++slot == capacity) slot = 0;
        }
    } else {
        for (; budget; --budget, pre = false) {
            u64* s = table + slot * SW;
            u64* vp = s + KW;
            u64 v = pre ? first->w[0] : ld_acquire(vp);
            if (v == 0) {
                if (atomicCAS(vp, 0ull, VAL_LOCK) == 0ull) {
#pragma unroll
                    for (int i = 0; i < KW; ++i) st_relaxed(s + i, key[i]);
                    st_release(vp, add | ((u64)mask << MASK_SHIFT));
                    is_new = true;
                    return slot;
                }
                continue;  // somebody else claimed it: look again
            }
            if (v == VAL_LOCK) continue;  // being written: look again
            bool eq = true;
#pragma unroll
            for (int i = 0; i < KW; ++i) eq = eq && (ld_relaxed(s + i) == key[i]);
            if (eq) { fold_value(vp, add, mask, v); return slot; }
            if (++slot == capacity) slot = 0;
        }
    }
    return capacity;  // probe budget exhausted
}

// table_upsert for the records the staged fast path of the region upsert could not place (home slot taken by another
// key): the same probe sequence, but PW consecutive slots are loaded per round trip and examined in registers, because
// what these records pay for is L2 latency per probe, not bytes. KW <= 2.
template <int KW>
__device__ __forceinline__ u64 table_upsert_wide(u64* __restrict__ table, u64 capacity, u64 hl, const u64 (&key)[KW], u64 add,
                                                 u32 mask, bool& is_new) {
    static_assert(KW <= 2, "CAS-on-key protocols only");
    constexpr int SW = SlotTraits<KW>::WORDS;
    constexpr int PW = KW == 1 ? 4 : 2;   // 64 bytes per round trip
    u64 slot = slot_of(hl, capacity);
    is_new = false;
    for (u64 budget = PROBE_BUDGET; budget; --budget) {
        u64 w[PW][KW + 1];
#pragma unroll
        for (int j = 0; j < PW; ++j) {
            const u64 sj = slot + j < capacity ? slot + j : slot + j - capacity;   // wraps around the end of the table
            if constexpr (KW == 1) ld_relaxed_v2(table + sj * SW, w[j][0], w[j][1]);
            else { u64 pad; ld_relaxed_v4(table + sj * SW, w[j][0], w[j][1], w[j][2], pad); }
        }
#pragma unroll
        for (int j = 0; j < PW; ++j) {
            const u64 sj = slot + j < capacity ? slot + j : slot + j - capacity;
            u64* s = table + sj * SW;
            bool eq = true, empty = true;
#pragma unroll
            for (int i = 0; i < KW; ++i) { eq = eq && w[j][i] == key[i]; empty = empty && w[j][i] == EMPTY_WORD; }
            u64 seen = w[j][KW];
            if (empty) {
                u64 o0, o1;
                if constexpr (KW == 1) {
                    if (cas128(s, EMPTY_WORD, 0ull, key[0], add | ((u64)mask << MASK_SHIFT), o0, o1)) { is_new = true; return sj; }
                    eq = o0 == key[0]; seen = o1;          // lost the slot: to the same key, or to another one
                } else {
                    if (cas128(s, EMPTY_WORD, EMPTY_WORD, key[0], key[1], o0, o1)) { is_new = true; eq = true; }
                    else eq = o0 == key[0] && o1 == key[1];
                    seen = 0;
                }
            }
            if (eq) { fold_value(s + KW, add, mask, seen); return sj; }
        }
        slot += PW;
        if (slot >= capacity) slot -= capacity;
    }
    return capacity;  // probe budget exhausted
}

// Find an existing key (after all inserts are complete). Returns capacity if absent.
template <int KW>
__device__ __forceinline__ u64 table_find(const u64* __restrict__ table, u64 capacity, u64 hl, const u64 (&key)[KW]) {
    constexpr int SW = SlotTraits<KW>::WORDS;
    u64 slot = slot_of(hl, capacity);
    for (u64 probes = 0; probes < capacity; ++probes) {
        const u64* s = table + slot * SW;
        bool eq = true, empty;
        if constexpr (KW <= 2) {
            empty = true;
#pragma unroll
            for (int i = 0; i < KW; ++i) { u64 c = s[i]; eq = eq && (c == key[i]); empty = empty && (c == EMPTY_WORD); }
        } else {
            empty = (s[KW] == 0);
#pragma unroll
            for (int i = 0; i < KW; ++i) eq = eq && (s[i] == key[i]);
        }
        if (empty) return capacity;
        if (eq) return slot;
        if (++slot == capacity) slot = 0;
    }
    return capacity;
}

template <int KW>
__device__ __forceinline__ bool slot_occupied(const u64* s) {
    if constexpr (KW <= 2) {
        bool empty = true;
#pragma unroll
        for (int i = 0; i < KW; ++i) empty = empty && (s[i] == EMPTY_WORD);
        return !empty;
    } else {
        return s[KW] != 0;
    }
}

}  // namespace gx
