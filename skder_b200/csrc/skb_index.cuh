// K2 -- per-reference seed index (open-addressing hash table keyed by the seed k-mer), repeat
// flags, chunk tables; K3 -- marker prescreen on the inverted marker index.
//
// Stand-ins for skani's seed map (kmer -> positions) and marker screen (`-s`, passed by skDER at
// src/skDER/skder.py:17 via bin/skder:199-201).  Bit-exact against oracle finish_sketch() /
// ora_screen().
#pragma once
#include "skb_common.cuh"
#include "skb_probe.cuh"

namespace skb {

// chunk_begin: for every chunk (and one sentinel per genome) the index of its first seed
__global__ void chunk_begin_kernel(const uint64_t *__restrict__ seeds, const uint64_t *__restrict__ g_seed_off,
                                   const uint32_t *__restrict__ g_chunk_off, int n_genomes,
                                   const uint32_t *__restrict__ chunk_start, uint32_t *__restrict__ chunk_begin,
                                   uint32_t total_entries /* total chunks + n_genomes */, uint32_t first) {
    uint32_t t = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total_entries) return;
    // entry t belongs to genome g where g_chunk_off[g] + g <= t
    int lo = 0, hi = n_genomes - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (g_chunk_off[mid] + (uint32_t)mid <= t)
            lo = mid;
        else
            hi = mid - 1;
    }
    const int g = lo;
    const uint32_t local = t - (g_chunk_off[g] + g);
    const uint32_t nch = g_chunk_off[g + 1] - g_chunk_off[g];
    const uint64_t *S = seeds + g_seed_off[g];
    const uint32_t ns = (uint32_t)(g_seed_off[g + 1] - g_seed_off[g]);
    if (local == nch) {
        chunk_begin[t] = ns;
        return;
    }
    const uint32_t start = chunk_start[g_chunk_off[g] + local];
    uint32_t l = 0, h = ns;  // first seed with pos >= start
    while (l < h) {
        uint32_t mid = (l + h) >> 1;
        if (seed_pos(S[mid]) < start)
            l = mid + 1;
        else
            h = mid;
    }
    chunk_begin[t] = l;
}

// ---- markers -----------------------------------------------------------------------------------
// keys are sorted; flag the first of every run of equal (marker, gid) keys
__global__ void unique_flag_kernel(const uint64_t *__restrict__ keys, uint64_t n, uint32_t *__restrict__ flag) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}
__global__ void unique_scatter_kernel(const uint64_t *__restrict__ keys, uint64_t n,
                                      const uint32_t *__restrict__ flag, const uint32_t *__restrict__ pos,
                                      uint64_t *__restrict__ out, uint32_t *__restrict__ g_marker_cnt) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flag[i]) return;
    out[pos[i]] = keys[i];
    atomicAdd(&g_marker_cnt[keys[i] & GID_MASK], 1u);
}
// (marker << 22 | gid)  ->  (gid << 42 | marker): sorting these yields the per-genome sorted lists
__global__ void swap_key_kernel(const uint64_t *__restrict__ inv, uint64_t n, uint64_t *__restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t k = inv[i];
    out[i] = ((k & GID_MASK) << (2 * K_MARKER)) | (k >> GID_BITS);
}
// sorted (gid << 42 | marker) keys -> markers of the flagged (unique) ones, and per-genome counts
__global__ void unique_scatter_swapped_kernel(const uint64_t *__restrict__ keys, uint64_t n,
                                              const uint32_t *__restrict__ flag, const uint32_t *__restrict__ pos,
                                              uint64_t *__restrict__ out, uint32_t *__restrict__ g_marker_cnt) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flag[i]) return;
    out[pos[i]] = keys[i] & MASK_MARKER;
    atomicAdd(&g_marker_cnt[keys[i] >> (2 * K_MARKER)], 1u);
}
__global__ void strip_gid_kernel(uint64_t *keys, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] &= MASK_MARKER;
}

// K3a (triangle): every run of equal markers in the inverted index contributes +1 to each pair of
// genomes in the run.  Thread i owns entry i and pairs it with the later entries of its run, so
// row = smaller genome id.  Rows are dealt to partitions in zig-zag order (row_owner, skb_common.cuh).
// The count matrix holds the partition's local rows [row0, row0 + n_rows) only (row tiles: skb_api.cu).
__global__ void screen_runs_kernel(const uint64_t *__restrict__ inv, uint64_t n, uint32_t *cnt, uint32_t n_genomes,
                                   int part, int n_parts, uint32_t row0, uint32_t n_rows) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t k = inv[i];
    const uint32_t gi = (uint32_t)(k & GID_MASK);
    if ((int)row_owner(gi, (uint32_t)n_parts) != part) return;
    const uint32_t rl = row_local(gi, (uint32_t)n_parts) - row0;  // unsigned: rows before the tile wrap past n_rows
    if (rl >= n_rows) return;
    const uint64_t m = k >> GID_BITS;
    uint32_t *row = cnt + (size_t)rl * n_genomes;
    for (uint64_t j = i + 1; j < n; j++) {
        const uint64_t kj = inv[j];
        if ((kj >> GID_BITS) != m) break;
        atomicAdd(row + (uint32_t)(kj & GID_MASK), 1u);
    }
}

// K3b: threshold the count matrix and compact the surviving pairs (a << 32 | b), a < b
__global__ void screen_compact_kernel(const uint32_t *__restrict__ cnt, uint32_t n_genomes, uint32_t n_rows_local,
                                      uint32_t row0, int part, int n_parts, const uint32_t *__restrict__ g_marker_cnt,
                                      double cutoff_scale /* screen^21, <=0: everything passes */,
                                      unsigned long long *pairs, unsigned long long *n_pairs,
                                      unsigned long long cap) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t total = (uint64_t)n_rows_local * n_genomes;
    bool pass = false;
    uint32_t a = 0, b = 0;
    if (t < total) {
        const uint32_t rl = (uint32_t)(t / n_genomes);
        b = (uint32_t)(t % n_genomes);
        a = row_global(row0 + rl, (uint32_t)part, (uint32_t)n_parts);
        if (a < n_genomes && b > a) {
            if (cutoff_scale <= 0.0)
                pass = true;
            else {
                const uint32_t ma = g_marker_cnt[a], mb = g_marker_cnt[b];
                const double cutoff = cutoff_scale * (double)(ma < mb ? ma : mb);
                pass = (double)cnt[t] > cutoff;
            }
        }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, pass);
    if (!bal) return;
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(n_pairs, (unsigned long long)__popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (pass) {
        const unsigned long long idx = base + __popc(bal & ((1u << lane) - 1));
        if (idx < cap) pairs[idx] = ((unsigned long long)a << 32) | b;
    }
}

// K3c (rect: dist / search): one thread per marker of a query genome; the marker's run in the
// inverted index is found by binary search and every reference genome in it gets +1.
__global__ void screen_rect_kernel(const uint64_t *__restrict__ inv, uint64_t n_inv,
                                   const uint64_t *__restrict__ markers, const uint64_t *__restrict__ g_marker_off,
                                   const int32_t *__restrict__ queries, int n_queries,
                                   const uint64_t *__restrict__ q_marker_prefix /* [n_queries+1] */,
                                   const int32_t *__restrict__ ref_slot /* [n_genomes], -1 = not a reference */,
                                   uint32_t *cnt, uint32_t n_refs) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= q_marker_prefix[n_queries]) return;
    const int qs = genome_of(q_marker_prefix, n_queries, t);
    const int32_t q = queries[qs];
    const uint64_t m = markers[g_marker_off[q] + (t - q_marker_prefix[qs])];
    const uint64_t key0 = m << GID_BITS;
    uint64_t l = 0, h = n_inv;
    while (l < h) {
        uint64_t mid = (l + h) >> 1;
        if (inv[mid] < key0)
            l = mid + 1;
        else
            h = mid;
    }
    for (; l < n_inv; l++) {
        const uint64_t k = inv[l];
        if ((k >> GID_BITS) != m) break;
        const uint32_t g = (uint32_t)(k & GID_MASK);
        const int32_t rs = ref_slot[g];
        if (rs >= 0 && (int32_t)g != q) atomicAdd(cnt + (size_t)qs * n_refs + rs, 1u);
    }
}
__global__ void screen_rect_compact_kernel(const uint32_t *__restrict__ cnt, const int32_t *__restrict__ refs,
                                           uint32_t n_refs, const int32_t *__restrict__ queries, uint32_t n_queries,
                                           const uint32_t *__restrict__ g_marker_cnt, double cutoff_scale,
                                           unsigned long long *pairs, unsigned long long *n_pairs,
                                           unsigned long long cap) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool pass = false;
    uint32_t a = 0, b = 0;
    if (t < (uint64_t)n_refs * n_queries) {
        const uint32_t qs = (uint32_t)(t / n_refs), rs = (uint32_t)(t % n_refs);
        a = (uint32_t)refs[rs];
        b = (uint32_t)queries[qs];
        if (a != b) {
            if (cutoff_scale <= 0.0)
                pass = true;
            else {
                const uint32_t ma = g_marker_cnt[a], mb = g_marker_cnt[b];
                pass = (double)cnt[t] > cutoff_scale * (double)(ma < mb ? ma : mb);
            }
        }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, pass);
    if (!bal) return;
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(n_pairs, (unsigned long long)__popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (pass) {
        const unsigned long long idx = base + __popc(bal & ((1u << lane) - 1));
        if (idx < cap) pairs[idx] = ((unsigned long long)a << 32) | b;
    }
}

// pairs whose reference genome (the one with more seeds; ties: b -- pair_setup_kernel's rule) lies in [own0, own1)
__global__ void pair_owned_flag_kernel(const uint64_t *__restrict__ g_seed_off, const unsigned long long *__restrict__ pairs,
                                       int64_t n_pairs, uint32_t n_genomes, uint32_t own0, uint32_t own1,
                                       uint32_t *__restrict__ flag) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pairs) return;
    const uint32_t a = (uint32_t)(pairs[t] >> 32), b = (uint32_t)(pairs[t] & 0xffffffffu);
    uint32_t f = 0;
    if (a < n_genomes && b < n_genomes) {
        const uint64_t nsa = g_seed_off[a + 1] - g_seed_off[a], nsb = g_seed_off[b + 1] - g_seed_off[b];
        const uint32_t r = nsb < nsa ? a : b;
        f = r >= own0 && r < own1;
    }
    flag[t] = f;
}
__global__ void pair_keep_kernel(const unsigned long long *__restrict__ pairs, int64_t n_pairs, const uint32_t *__restrict__ flag,
                                 const uint32_t *__restrict__ pos, unsigned long long *__restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_pairs && flag[t]) out[pos[t]] = pairs[t];
}

// explicit pair list: |Ma ∩ Mb| by sorted-list intersection, one warp per pair (binary search of
// a's markers in b's list, ballot/popcount).  Independent of the inverted index: used to
// cross-check it.
__global__ void shared_markers_kernel(const uint64_t *__restrict__ markers,
                                      const uint64_t *__restrict__ g_marker_off, const uint32_t *__restrict__ pa,
                                      const uint32_t *__restrict__ pb, int64_t n_pairs, long long *out) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_pairs) return;
    const uint64_t *A = markers + g_marker_off[pa[w]];
    const uint64_t *B = markers + g_marker_off[pb[w]];
    const uint64_t na = g_marker_off[pa[w] + 1] - g_marker_off[pa[w]];
    const uint64_t nb = g_marker_off[pb[w] + 1] - g_marker_off[pb[w]];
    long long c = 0;
    for (uint64_t i0 = 0; i0 < na; i0 += 32) {
        const uint64_t i = i0 + lane;
        bool hit = false;
        if (i < na) {
            const uint64_t m = A[i];
            uint64_t l = 0, h = nb;
            while (l < h) {
                uint64_t mid = (l + h) >> 1;
                if (B[mid] < m)
                    l = mid + 1;
                else
                    h = mid;
            }
            hit = l < nb && B[l] == m;
        }
        c += __popc(__ballot_sync(0xffffffffu, hit));
    }
    if (lane == 0) out[w] = c;
}

}  // namespace skb
