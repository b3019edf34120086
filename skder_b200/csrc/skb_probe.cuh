// K2 -- per-reference seed index: two-choice bucketed hash table keyed by the seed k-mer, and the
// batched anchor lookup the chunk kernel's P1 runs against it.
//
// Stand-in for skani's seed map (k-mer -> positions) used by the pairwise estimator behind
// `skani triangle|dist|search` (reference call sites src/skDER/skder.py:16-18, :58-59, :119).
//
// Layout: buckets of 4 slots = one 32-byte sector.  A k-mer has two home buckets; an entry goes to the
// emptier one, so at load factor 0.5 a full bucket is rare and both-full practically never happens
// (then entries spill linearly after the first home).  A lookup therefore costs a FIXED two sector
// reads issued together -- no dependent probe chain and no lane waiting for another lane's chain,
// which is what bounded the linear-probing version (ncu: 37% of samples on the chain's scoreboard).
#pragma once
#include "skb_common.cuh"

namespace skb {

constexpr int PJ = 2;         // lookups in flight per lane in the batched probe
constexpr int STAGE_CAP = 8;  // hits staged per seed (max_mult upper bound)

__host__ __device__ inline uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32); }

__host__ __device__ inline void tab_homes(uint32_t kmer, uint32_t nb, uint32_t &b1, uint32_t &b2) {
    b1 = mulhi32(kmer * 0x9E3779B1u, nb);
    b2 = mulhi32((kmer ^ (kmer >> 15)) * 0x85EBCA77u + 0x165667B1u, nb);
    if (b2 == b1) b2 = (b1 + 1 == nb) ? 0 : b1 + 1;  // nb >= 2
}

__device__ __forceinline__ int genome_of(const uint64_t *__restrict__ off, int n, uint64_t i) {
    int lo = 0, hi = n - 1;  // last g with off[g] <= i
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (off[mid] <= i)
            lo = mid;
        else
            hi = mid - 1;
    }
    return lo;
}

// genome_of for a warp whose lanes hold consecutive, ascending i: lane 0 searches, the others walk on from there
// (a warp spans 32 records, a genome ~40,000).  All 32 lanes must call.
__device__ __forceinline__ int genome_of_warp(const uint64_t *__restrict__ off, int n, uint64_t i) {
    const uint64_t i0 = __shfl_sync(0xffffffffu, i, 0);
    int g = 0;
    if ((threadIdx.x & 31) == 0) g = genome_of(off, n, i0);
    g = __shfl_sync(0xffffffffu, g, 0);
    while (g + 1 < n && off[g + 1] <= i) g++;
    return g;
}

// one thread per seed record: insert into its genome's table (64-bit CAS; no deletions ever)
__global__ void tab_insert_kernel(const uint64_t *__restrict__ seeds, uint64_t n_seeds,
                                  const uint64_t *__restrict__ g_seed_off, int n_genomes, uint64_t *tab,
                                  const uint64_t *__restrict__ g_tab_off, const uint32_t *__restrict__ g_tab_buckets,
                                  uint64_t first /* records [first, n_seeds) are inserted */) {
    uint64_t i = first + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int g = genome_of_warp(g_seed_off, n_genomes, i < n_seeds ? i : n_seeds - 1);
    if (i >= n_seeds) return;
    const uint32_t nb = g_tab_buckets[g];
    volatile unsigned long long *T = reinterpret_cast<volatile unsigned long long *>(tab + g_tab_off[g]);
    const unsigned long long rec = seeds[i] & ~2ull;
    uint32_t b1, b2;
    tab_homes(seed_kmer(rec), nb, b1, b2);
    for (;;) {
        // occupancy of both homes from two 256-bit reads (slots fill in order); .cg: other threads are inserting
        unsigned long long a[BUCKET], c[BUCKET];
        asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];"
                     : "=l"(a[0]), "=l"(a[1]), "=l"(a[2]), "=l"(a[3]) : "l"(tab + g_tab_off[g] + (size_t)b1 * BUCKET) : "memory");
        asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];"
                     : "=l"(c[0]), "=l"(c[1]), "=l"(c[2]), "=l"(c[3]) : "l"(tab + g_tab_off[g] + (size_t)b2 * BUCKET) : "memory");
        int o1 = 0, o2 = 0;
#pragma unroll
        for (int k = 0; k < (int)BUCKET; k++) {
            o1 += a[k] != TAB_EMPTY;
            o2 += c[k] != TAB_EMPTY;
        }
        if (o1 == (int)BUCKET && o2 == (int)BUCKET) break;
        const uint32_t tb = o2 < o1 ? b2 : b1;
        const int slot = o2 < o1 ? o2 : o1;
        if (atomicCAS(const_cast<unsigned long long *>(&T[(size_t)tb * BUCKET + slot]), (unsigned long long)TAB_EMPTY, rec) ==
            TAB_EMPTY)
            return;
    }
    // both homes full: spill linearly after the first home, skipping the second
    uint32_t b = b1;
    for (;;) {
        b = (b + 1 == nb) ? 0 : b + 1;
        if (b == b2) continue;
        for (uint32_t j = 0; j < BUCKET; j++)
            if (atomicCAS(const_cast<unsigned long long *>(&T[(size_t)b * BUCKET + j]), (unsigned long long)TAB_EMPTY, rec) ==
                TAB_EMPTY)
                return;
    }
}

struct Bucket2 {
    ulonglong2 a0, a1, c0, c1;  // home 1 slots 0..3, home 2 slots 0..3
};
__device__ __forceinline__ Bucket2 empty_buckets() {
    Bucket2 B;
    B.a0 = B.a1 = B.c0 = B.c1 = make_ulonglong2(TAB_EMPTY, TAB_EMPTY);
    return B;
}
// One bucket = one 32-byte sector = ONE 256-bit load (LDG.E.256, sm_100).  Lookups are scattered, so every load
// instruction costs the L1 data pipe a wavefront per lane; with two 128-bit loads per bucket that pipe was the
// anchor kernel's limit (ncu: l1tex data-pipe wavefronts 99.6% of peak).  Buckets are 32-byte aligned (table
// offsets are multiples of BUCKET words, the allocation is 256-byte aligned).
__device__ __forceinline__ void ld_bucket(const uint64_t *__restrict__ p, ulonglong2 &lo, ulonglong2 &hi) {
    asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(lo.x), "=l"(lo.y), "=l"(hi.x), "=l"(hi.y) : "l"(p));
}
__device__ __forceinline__ Bucket2 load_buckets(const uint64_t *__restrict__ T, uint32_t b1, uint32_t b2) {
    Bucket2 B;
    ld_bucket(T + (size_t)b1 * BUCKET, B.a0, B.a1);
    ld_bucket(T + (size_t)b2 * BUCKET, B.c0, B.c1);
    return B;
}
__device__ __forceinline__ bool both_full(const Bucket2 &B) { return B.a1.y != TAB_EMPTY && B.c1.y != TAB_EMPTY; }

// entries holding `kmer` (an empty slot's k-mer field, all ones, is never a canonical k-mer)
__device__ __forceinline__ int tab_count(const uint64_t *__restrict__ T, uint32_t nb, uint32_t kmer, int cap) {
    uint32_t b1, b2;
    tab_homes(kmer, nb, b1, b2);
    const Bucket2 B = load_buckets(T, b1, b2);
    int c = (seed_kmer(B.a0.x) == kmer) + (seed_kmer(B.a0.y) == kmer) + (seed_kmer(B.a1.x) == kmer) +
            (seed_kmer(B.a1.y) == kmer) + (seed_kmer(B.c0.x) == kmer) + (seed_kmer(B.c0.y) == kmer) +
            (seed_kmer(B.c1.x) == kmer) + (seed_kmer(B.c1.y) == kmer);
    if (both_full(B)) {
        uint32_t b = b1;
        for (;;) {
            b = (b + 1 == nb) ? 0 : b + 1;
            if (b == b2) continue;
            ulonglong2 x0, x1;
            ld_bucket(T + (size_t)b * BUCKET, x0, x1);
            c += (seed_kmer(x0.x) == kmer) + (seed_kmer(x0.y) == kmer) + (seed_kmer(x1.x) == kmer) + (seed_kmer(x1.y) == kmer);
            if (x1.y == TAB_EMPTY || c >= cap) break;
        }
    }
    return c;
}

// flag seeds whose k-mer occurs more than max_mult times in their own genome (bit 1 of the record)
__global__ void rep_flag_kernel(uint64_t *seeds, uint64_t n_seeds, const uint64_t *__restrict__ g_seed_off,
                                int n_genomes, const uint64_t *__restrict__ tab,
                                const uint64_t *__restrict__ g_tab_off, const uint32_t *__restrict__ g_tab_buckets,
                                int max_mult, uint64_t first) {
    uint64_t i = first + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int g = genome_of_warp(g_seed_off, n_genomes, i < n_seeds ? i : n_seeds - 1);
    if (i >= n_seeds) return;
    const uint64_t s = seeds[i];
    const int c = tab_count(tab + g_tab_off[g], g_tab_buckets[g], seed_kmer(s), max_mult + 1);
    if (c > max_mult) seeds[i] = s | 2ull;
}

// staged hit: (ref_pos << 1) | strand relation
__device__ __forceinline__ uint32_t enc_hit(uint64_t e, uint64_t sd) {
    return (seed_pos(e) << 1) | (uint32_t)(seed_strand(e) != seed_strand(sd));
}

// One lookup of the batch: lane's seed `sd` with its two home buckets B already loaded.  Appends the
// seed's anchors (if it has 1..mult hits) after `base`; returns base + anchors of all 32 lanes.
__device__ __forceinline__ int probe_one(uint64_t sd, const Bucket2 &B, const uint64_t *__restrict__ T, uint32_t nb,
                                         int mult, int max_mult, int max_anchors, uint32_t *stage, uint64_t *anc,
                                         int base, int s, uint32_t cstart, int lane) {
    const uint32_t km = seed_kmer(sd);
    const uint64_t e8[8] = {B.a0.x, B.a0.y, B.a1.x, B.a1.y, B.c0.x, B.c0.y, B.c1.x, B.c1.y};
    // the k-mer is the top 30 bits of a record: compare on the high word only.  Flagged / out-of-range
    // lanes hold all-empty buckets, whose k-mer field (all ones) is never a canonical k-mer.
    const uint32_t kmhi = km << 2;
    int c = 0;
    uint64_t e1 = 0;  // the hit when there is exactly one
#pragma unroll
    for (int x = 0; x < 8; x++) {
        const bool hit = (((uint32_t)(e8[x] >> 32)) ^ kmhi) < 4u;
        c += hit;
        if (hit) e1 = e8[x];
    }
    const bool spilled = both_full(B);  // practically never: both homes full -> entries may have spilled
    const bool slow = spilled || c > 1;
    if (slow) {  // rare: repeats or spill -> stage all hits, sorted by ref position
        int cc = 0;
#pragma unroll
        for (int x = 0; x < 8; x++)
            if ((((uint32_t)(e8[x] >> 32)) ^ kmhi) < 4u) {
                if (cc < STAGE_CAP) stage[lane * STAGE_CAP + cc] = enc_hit(e8[x], sd);
                cc++;
            }
        if (spilled) {
            uint32_t b1, b2;
            tab_homes(km, nb, b1, b2);
            uint32_t b = b1;
            for (;;) {
                b = (b + 1 == nb) ? 0 : b + 1;
                if (b == b2) continue;
                ulonglong2 x0, x1;
                ld_bucket(T + (size_t)b * BUCKET, x0, x1);
                const uint64_t e4[4] = {x0.x, x0.y, x1.x, x1.y};
#pragma unroll
                for (int x = 0; x < 4; x++)
                    if (seed_kmer(e4[x]) == km) {
                        if (cc < STAGE_CAP) stage[lane * STAGE_CAP + cc] = enc_hit(e4[x], sd);
                        cc++;
                    }
                if (x1.y == TAB_EMPTY || cc > max_mult) break;
            }
        }
        c = cc;
        if (c > mult) c = 0;
        for (int x = 1; x < c; x++) {  // insertion sort, c <= STAGE_CAP
            const uint32_t v = stage[lane * STAGE_CAP + x];
            int y = x - 1;
            while (y >= 0 && stage[lane * STAGE_CAP + y] > v) {
                stage[lane * STAGE_CAP + y + 1] = stage[lane * STAGE_CAP + y];
                y--;
            }
            stage[lane * STAGE_CAP + y + 1] = v;
        }
    }
    const uint64_t lowbits = ((uint64_t)(seed_pos(sd) - cstart) << 17) | (uint64_t)s;
    const unsigned any_slow = __ballot_sync(0xffffffffu, slow);
    if (!any_slow) {  // common: every lane has 0 or 1 hit -> offsets from one ballot
        const unsigned has = __ballot_sync(0xffffffffu, c == 1);
        if (c == 1) {
            const int dst = base + __popc(has & ((1u << lane) - 1u));
            const uint32_t v = enc_hit(e1, sd);
            if (dst < max_anchors) __stcg(anc + dst, ((uint64_t)(v >> 1) << 32) | lowbits | ((uint64_t)(v & 1u) << 16));
        }
        return base + __popc(has);
    }
    int pre = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, pre, d);
        if (lane >= d) pre += u;
    }
    const int tot = __shfl_sync(0xffffffffu, pre, 31);
    pre -= c;
    if (!slow) {
        if (c == 1) {
            const uint32_t v = enc_hit(e1, sd);
            if (base + pre < max_anchors)
                __stcg(anc + base + pre, ((uint64_t)(v >> 1) << 32) | lowbits | ((uint64_t)(v & 1u) << 16));
        }
    } else {
        for (int x = 0; x < c; x++) {
            const int dst = base + pre + x;
            const uint32_t v = stage[lane * STAGE_CAP + x];
            if (dst < max_anchors) __stcg(anc + dst, ((uint64_t)(v >> 1) << 32) | lowbits | ((uint64_t)(v & 1u) << 16));
        }
    }
    return base + tot;
}

// first batch of a task's seed records (PJ x 32), one per lane and slot; out-of-range = flagged record
__device__ __forceinline__ void load_first_batch(const uint64_t *__restrict__ qs, int nseeds, int lane, uint64_t (&sdn)[PJ]) {
#pragma unroll
    for (int j = 0; j < PJ; j++) {
        const int s = 32 * j + lane;
        sdn[j] = s < nseeds ? qs[s] : 2ull;  // rep bit set = skip
    }
}

// sdn holds this task's first batch on entry and the NEXT task's first batch (qs_next/ns_next) on exit,
// so the only latency a task exposes is that of its bucket reads.
__device__ __forceinline__ int emit_anchors(const uint64_t *__restrict__ qs, int nseeds, uint32_t cstart,
                                            const uint64_t *__restrict__ T, uint32_t nb, int mult, int max_mult,
                                            int max_anchors, uint32_t *stage, uint64_t *anc, int lane,
                                            uint64_t (&sdn)[PJ], const uint64_t *__restrict__ qs_next, int ns_next) {
    int base = 0;
    for (int s0 = 0; s0 < nseeds; s0 += 32 * PJ) {
        uint64_t sd[PJ];
        Bucket2 B[PJ];
#pragma unroll
        for (int j = 0; j < PJ; j++) {
            sd[j] = sdn[j];
            if (!seed_rep(sd[j])) {
                uint32_t b1, b2;
                tab_homes(seed_kmer(sd[j]), nb, b1, b2);
                B[j] = load_buckets(T, b1, b2);
            } else
                B[j] = empty_buckets();
        }
        if (s0 + 32 * PJ < nseeds) {
#pragma unroll
            for (int j = 0; j < PJ; j++) {
                const int s = s0 + 32 * PJ + 32 * j + lane;
                sdn[j] = s < nseeds ? qs[s] : 2ull;
            }
        } else
            load_first_batch(qs_next, ns_next, lane, sdn);
#pragma unroll
        for (int j = 0; j < PJ; j++) {
            if (s0 + 32 * j >= nseeds) break;  // warp-uniform
            base = probe_one(sd[j], B[j], T, nb, mult, max_mult, max_anchors, stage, anc, base, s0 + 32 * j + lane,
                             cstart, lane);
        }
    }
    return base;
}

}  // namespace skb
