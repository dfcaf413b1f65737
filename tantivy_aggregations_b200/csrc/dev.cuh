// dev.cuh — device-side descriptors and primitives shared by every kernel.
//
// Column layout in HBM: the bit-packed payload of a tantivy fast-field column
// (LSB-first, value i in bits [i*num_bits, (i+1)*num_bits)), WITHOUT the 16-byte header, in a
// 256-byte aligned allocation that is zero-padded to a whole number of 2048-value tiles plus 16
// bytes, so any aligned 8/16-byte read that touches a valid value stays inside the allocation.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tagg.h"

#define TAGG_MAX_COLS 16      // distinct device columns a plan may reference (multi fields take 2)
#define TAGG_MAX_FILTERS 8    // FILTER nodes per plan
#define TAGG_MAX_NODES 64     // nodes per plan
#define TAGG_MAX_DEPTH 12     // nesting depth of the tree walker
#define TAGG_MAX_SCOPES 16
#define TAGG_MAX_ROOT_SLOTS 16
#define TAGG_TILE_DOCS 2048   // column allocations are padded to this many values

struct DevColumn {
    const uint64_t* words;  // packed payload viewed as little-endian u64 words
    uint64_t min_value;
    uint64_t max_value;     // min_value + amplitude
    uint64_t mask;
    uint64_t n_values;
    uint32_t num_bits;
    uint32_t kind;
};

// FastFieldReader::get -> code (tantivy BitUnpacker::get + min_value, restated for aligned loads)
__device__ __forceinline__ uint64_t col_get(const DevColumn& c, uint64_t i) {
    if (c.num_bits == 0) return c.min_value;
    uint64_t bit = i * c.num_bits;
    uint64_t w = bit >> 6;
    uint32_t sh = (uint32_t)bit & 63u;
    uint64_t lo = __ldg(c.words + w);
    uint64_t v = lo >> sh;
    if (sh + c.num_bits > 64) {
        uint64_t hi = __ldg(c.words + w + 1);
        v |= (hi << 1) << (63u - sh);
    }
    return (v & c.mask) + c.min_value;
}

// ---- type codecs (tantivy common::{u64_to_f64, u64_to_i64}) ---------------------------------
__device__ __forceinline__ double code_to_f64(uint64_t c) {
    uint64_t bits = (c >> 63) ? (c ^ 0x8000000000000000ull) : ~c;
    return __longlong_as_double((long long)bits);
}
__device__ __forceinline__ uint64_t f64_to_code(double v) {
    uint64_t bits = (uint64_t)__double_as_longlong(v);
    return (bits >> 63) == 0 ? bits ^ 0x8000000000000000ull : ~bits;
}
// value bits in the column's natural type (u64 | i64 two's complement | f64 IEEE bits)
__device__ __forceinline__ uint64_t code_to_bits(uint32_t kind, uint64_t c) {
    if (kind == TAGG_U64) return c;
    if (kind == TAGG_F64) return (c >> 63) ? (c ^ 0x8000000000000000ull) : ~c;
    return c ^ 0x8000000000000000ull;
}

// histogram.rs:136-152 — exact IEEE: n = k - start; ord = floor(n / interval) as u64 (saturating).
// Returns false when the document is skipped (NaN or n < 0).
// `kind` is the key column's type: f64 in the reference; i64 / date keys (date_histogram, README.md:41 — seconds since the
// epoch) are converted to f64 first, exact below 2^53.
__device__ __forceinline__ bool hist_ord(uint64_t code, double start, double interval, uint64_t* ord, uint32_t kind = TAGG_F64) {
    double k = kind == TAGG_F64 ? code_to_f64(code) : kind == TAGG_U64 ? (double)code : (double)(long long)(code ^ 0x8000000000000000ull);
    if (k != k) return false;
    double n = __dsub_rn(k, start);
    if (n < 0.0) return false;
    double q = floor(__ddiv_rn(n, interval));
    uint64_t o;
    if (!(q == q) || q <= 0.0) o = 0;
    else if (q >= 18446744073709551616.0) o = ~0ull;
    else o = (uint64_t)q;
    *ord = o;
    return true;
}

__device__ __forceinline__ uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}

// ---- docsets ----------------------------------------------------------------------------------
enum { DS_ALL = 0, DS_BITSET = 1, DS_IDS = 2, DS_RANGE = 3 };
struct DevDocset {
    int32_t kind;
    int32_t col;            // DS_RANGE: device column slot
    const uint32_t* words;  // DS_BITSET: bit d = words[d>>5] >> (d&31) & 1
    const uint32_t* ids;    // DS_IDS
    uint64_t n;             // DS_IDS: id count
    uint64_t lo, hi;        // DS_RANGE
};

struct DevSegment {
    uint32_t max_doc;
    uint32_t has_deletes;
    const uint32_t* deleted;  // DeleteBitSet words
    DevDocset main;
    DevDocset filters[TAGG_MAX_FILTERS];
    DevColumn cols[TAGG_MAX_COLS];
};

__device__ __forceinline__ bool docset_test(const DevSegment& s, const DevDocset& d, uint32_t doc) {
    switch (d.kind) {
        case DS_ALL: return true;
        case DS_BITSET: return (__ldg(d.words + (doc >> 5)) >> (doc & 31)) & 1u;
        case DS_RANGE: {
            uint64_t c = col_get(s.cols[d.col], doc);
            return c >= d.lo && c <= d.hi;
        }
        default: return false;
    }
}

// ---- plan ---------------------------------------------------------------------------------------
struct DevNode {
    uint8_t op, kind, multi, pred;
    uint16_t col;        // device column slot (multi: idx column; vals column = col + 1)
    uint16_t end;        // index one past this node's sub-tree
    uint16_t scope;      // enclosing scope
    uint16_t own_scope;  // TERMS / HISTOGRAM: the scope this node keys
    uint16_t slot;       // leaf metrics: accumulator slot
    uint16_t aux;        // FILTER: filter index
    uint32_t skip;       // 1: this sub-tree was handled by a streaming launch, the tree walker steps over it
    const uint8_t* lut;  // PRED_LUT bitmap (device)
    double f0, f1;
    uint64_t u0, u1;
};

enum { SCOPE_DENSE = 0, SCOPE_HASH = 1 };
enum { ST_EMPTY = 0, ST_BUSY = 1, ST_READY = 2 };
struct DevScope {
    int32_t mode;
    int32_t parent;        // parent scope (-1 for root)
    uint64_t capacity;     // buckets addressable (dense: parent_cap * dom_size; hash: table size, pow2)
    uint64_t dom_min;      // dense: smallest key
    uint64_t dom_size;     // dense: keys per parent bucket
    uint8_t* present;      // dense: bucket touched
    uint64_t* keys;        // hash
    uint32_t* parents;     // hash
    uint32_t* state;       // hash
    unsigned long long* used;  // hash: claimed slots
};

#define INVALID_BUCKET 0xFFFFFFFFu

// multiplicative (Fibonacci) hashing: one 64-bit multiply in front of the probe load, slot = TOP bits of the product
__device__ __forceinline__ uint64_t hash_key(uint64_t key, uint32_t parent) {
    return (key ^ ((uint64_t)parent * 0xD6E8FEB86659FD93ull)) * 0x9E3779B97F4A7C15ull;
}

// `entry(key).or_insert_with(create_fruit)` (terms.rs:129-130, histogram.rs:148-149):
// bucket index of (parent bucket, key) in scope `sc`, created on first touch.
__device__ __forceinline__ uint32_t scope_lookup(uint32_t* overflow, const DevScope& sc, uint32_t parent, uint64_t key) {
    if (sc.mode == SCOPE_DENSE) {
        uint64_t rel = key - sc.dom_min;
        if (key < sc.dom_min || rel >= sc.dom_size) return INVALID_BUCKET;
        uint64_t idx = (uint64_t)parent * sc.dom_size + rel;
        if (!sc.present[idx]) sc.present[idx] = 1;
        return (uint32_t)idx;
    }
    uint64_t mask = sc.capacity - 1;
    uint64_t h = hash_key(key, parent) >> __clzll(mask);  // capacity is a power of two >= 1024
    const bool root = sc.parent == 0;  // buckets of a top-level scope all hang off parent bucket 0
    for (uint64_t probes = 0; probes <= mask;) {
        // fast pre-check on the key cell alone (a key, once written, never changes; an all-zero cell is
        // ambiguous — empty or key 0 — and takes the state-checked path below)
        if (root) {
            const uint64_t k = __ldcg(sc.keys + h);
            if (k != 0) {
                if (k == key) return (uint32_t)h;
                h = (h + 1) & mask;
                probes++;
                continue;
            }
        }
        uint32_t st = *((volatile uint32_t*)(sc.state + h));
        if (st == ST_READY) {
            if (*((volatile uint64_t*)(sc.keys + h)) == key && *((volatile uint32_t*)(sc.parents + h)) == parent)
                return (uint32_t)h;
            h = (h + 1) & mask;
            probes++;
            continue;
        }
        if (st == ST_EMPTY) {
            // keep the load factor <= 3/4: beyond that report overflow and let the host grow the table
            if (*((volatile unsigned long long*)sc.used) * 4ull >= sc.capacity * 3ull) break;
            uint32_t old = atomicCAS(sc.state + h, (uint32_t)ST_EMPTY, (uint32_t)ST_BUSY);
            if (old == ST_EMPTY) {
                sc.keys[h] = key;
                sc.parents[h] = parent;
                __threadfence();
                atomicExch(sc.state + h, (uint32_t)ST_READY);
                atomicAdd(sc.used, 1ull);
                return (uint32_t)h;
            }
        }
        // BUSY (or lost the race): re-read the same slot
    }
    atomicExch(overflow, 1u);
    return INVALID_BUCKET;
}


struct DevSlot {
    uint64_t* acc;   // per bucket of the enclosing scope.  MIN stores max(~code) so zero == empty identity
    uint8_t* seen;   // per bucket: Option is Some
    // f64 MIN / MAX on a column that holds NaN or both signed zeros (exact PartialOrd fold of minmax.rs:97-106, exec.cu
    // edge_scan): 3 * edge_cap cells of ~position (zero = none) — [first collected value | first -0.0 | first +0.0]
    uint64_t* edge;
    uint64_t edge_cap;
};

// f64 codes: the non-NaN values are exactly [code(-inf), code(+inf)]; the two zeros are adjacent
#define CODE_NEG_INF 0x000FFFFFFFFFFFFFull
#define CODE_POS_INF 0xFFF0000000000000ull
#define CODE_NEG_ZERO 0x7FFFFFFFFFFFFFFFull
#define CODE_POS_ZERO 0x8000000000000000ull
#define F64_NEG_ZERO_BITS 0x8000000000000000ull  // identity of every f64 sum cell: x + -0.0 == x for all x (sum.rs:95-102)
#define EDGE_POS_BITS 40                          // position of a collected value = segment index << 40 | doc / value index

struct DevPlan {
    uint32_t n_nodes, n_scopes, n_slots, n_root_slots;
    DevNode nodes[TAGG_MAX_NODES];
    DevScope scopes[TAGG_MAX_SCOPES];
    DevSlot slots[TAGG_MAX_NODES];
    uint16_t root_slot_nodes[TAGG_MAX_ROOT_SLOTS];  // node index of each register-accumulated root slot
    int16_t slot_root_index[TAGG_MAX_NODES];        // slot -> index in the per-thread root accumulators, -1 if none
    uint32_t* overflow;     // set when a hash scope ran out of room
    // PERCENTILES materialisation (generic path): values appended per slot
    uint64_t* pct_codes[4];
    uint32_t* pct_buckets[4];
    unsigned long long* pct_count[4];
    uint64_t pct_cap[4];
};
