// K4 -- fused seed anchoring + per-chunk chaining + chain selection + ANI/AF, one CTA per genome pair.
//
// Stands in for skani's pairwise estimator behind `skani triangle|dist|search`
// (reference call sites src/skDER/skder.py:16-18, :58-59, :119).  Integer results (anchors, seeds,
// spans, chain count) are bit-exact against oracle/skani_oracle.c ora_pair(); ANI/AF are the same
// IEEE double expressions (device pow() may differ from glibc's in the last ulp).
//
// Work split: a CTA of 16 warps takes one pair; each warp takes query chunks (20 kb windows of the
// query genome) round-robin.  Per chunk the warp
//   1. streams the chunk's position-ordered seeds (coalesced 8-byte records), probes the reference's
//      hash index (L2-resident: all ~50 members of a clade are compared against the same tables),
//      and drops the anchors (ref_pos, query_pos, strand relation) into its shared-memory slab;
//   2. bitonic-sorts the anchors by (ref_pos, query_pos);
//   3. runs the chaining DP: anchor i on all lanes, lane l scores predecessor i-1-l, one
//      REDUX (__reduce_max_sync) picks the best predecessor (ties: nearest);
//   4. keeps the best end of every DP tree with >= min_anchors / min_score, at most chunk_cap per chunk.
// The CTA then orders all candidates by (score desc, chunk, ordinal), resolves the greedy
// non-overlap selection in parallel, accumulates per-chunk anchors/seeds and clipped spans, and
// reduces ANI = sum(S_c * (A_c/S_c)^(1/15)) / sum(S_c), AF = span / genome length.
#pragma once
#include "skb_common.cuh"
#include "skb_index.cuh"

namespace skb {

constexpr int ANI_WARPS = 16;
constexpr int ANI_THREADS = ANI_WARPS * 32;
constexpr int MAXA = 512;   // anchors per chunk
constexpr int MAXP = 1024;  // chain candidates per pair
constexpr int STAGE = 8;    // max_mult upper bound (anchors staged per seed)

struct AniParams {
    int32_t band_bp, max_gap, anchor_score, min_anchors, min_score, max_mult, max_chunk_chains;
    int32_t ovl_num, ovl_den, span_ext, min_chunk_seeds;
    double debias_a, debias_g;
};

struct __align__(16) WarpSlab {
    uint64_t key[MAXA];   // (ref_pos << 32) | (query_pos << 1) | rev
    int32_t f[MAXA];      // DP score
    uint16_t root[MAXA];  // first anchor of the best chain ending here
    uint16_t cnt[MAXA];   // anchors in that chain
    uint32_t bor[MAXA];   // best-of-root / staging area (MAXA*4 = 32 lanes * STAGE * 8 bytes)
};
static_assert(sizeof(WarpSlab) == 10240, "slab size");
static_assert(MAXA * 4 == 32 * STAGE * 8, "staging area must fit in bor[]");

struct __align__(16) Cand {
    uint32_t q0, q1, r0, r1;
    uint32_t chunk;
    uint16_t score, n_anchors;
    uint16_t n_seeds;
    uint8_t ordinal, rev;
    uint32_t pad;
};
static_assert(sizeof(Cand) == 32, "cand size");

constexpr size_t ANI_SMEM_SLABS = sizeof(WarpSlab) * ANI_WARPS;           // 163840
constexpr size_t ANI_SMEM_CANDS = sizeof(Cand) * MAXP;                    // 32768
constexpr size_t ANI_SMEM_BYTES = ANI_SMEM_SLABS + ANI_SMEM_CANDS + 256;  // + control block
// after the chunk phase the slab area is reused: sort keys | state | chunk accumulators
constexpr size_t FIN_KEYS_OFF = 0;                        // uint64_t[MAXP]
constexpr size_t FIN_STATE_OFF = FIN_KEYS_OFF + 8 * MAXP; // uint8_t[MAXP]
constexpr size_t FIN_ACC_OFF = FIN_STATE_OFF + MAXP;      // uint32_t A[nch], S[nch]
constexpr uint32_t MAX_CHUNKS_PER_GENOME = (uint32_t)((ANI_SMEM_SLABS - FIN_ACC_OFF) / 8);

struct AniCtl {
    int n_cand;
    int next_pair;
    int unresolved;
    int used;
    unsigned long long span_q, span_r, a_tot, s_tot;
    double sw, sx;
};

// warp-level bitonic sort of n (<= MAXA) 64-bit keys in shared memory, ascending
__device__ __forceinline__ void warp_sort_keys(uint64_t *key, int n, int lane) {
    int m = 1;
    while (m < n) m <<= 1;
    for (int i = n + lane; i < m; i += 32) key[i] = ~0ull;
    __syncwarp();
    for (int k = 2; k <= m; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < (m >> 1); t += 32) {
                // t-th compare-exchange of this stage: i has bit j clear
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i | j;
                const bool up = (i & k) == 0;
                const uint64_t a = key[i], b = key[p];
                if ((a > b) == up) {
                    key[i] = b;
                    key[p] = a;
                }
            }
            __syncwarp();
        }
    }
}

// One query chunk against one reference index.  Appends up to chunk_cap candidates to cands[].
__device__ __forceinline__ void process_chunk(const AniParams &prm, WarpSlab &w, int lane,
                                              const uint64_t *__restrict__ qs /* chunk's seeds */, int nseeds,
                                              uint32_t chunk_id, uint32_t chunk_start,
                                              const uint64_t *__restrict__ T, uint32_t tmask, int tbits,
                                              int chunk_cap, Cand *cands, int *n_cand) {
    if (nseeds <= 0) return;
    uint64_t *stage = reinterpret_cast<uint64_t *>(w.bor);  // [32][STAGE]
    // ---- 1. anchors.  Optimistic pass with the full multiplicity cap; per-level tallies tell
    //         whether a lower cap is needed to fit MAXA (oracle: halve until it fits).
    int mult = prm.max_mult;
    int n = 0;
    for (int attempt = 0; attempt < 2; attempt++) {
        int tally[4] = {0, 0, 0, 0};  // anchors if the cap were max_mult >> lv
        int base = 0;
        for (int s0 = 0; s0 < nseeds; s0 += 32) {
            const int s = s0 + lane;
            int c = 0;
            uint64_t sd = 0;
            if (s < nseeds) {
                sd = qs[s];
                if (!seed_rep(sd)) {
                    const uint32_t km = seed_kmer(sd);
                    uint32_t h = tab_slot(km, tbits);
                    for (;;) {
                        const uint64_t e = __ldg(T + h);
                        if (e == TAB_EMPTY) break;
                        if (seed_kmer(e) == km) {
                            if (c < STAGE)
                                stage[lane * STAGE + c] = ((uint64_t)seed_pos(e) << 32) |
                                                          ((uint64_t)seed_pos(sd) << 1) |
                                                          (uint64_t)(seed_strand(e) != seed_strand(sd));
                            c++;
                            if (c > prm.max_mult) break;
                        }
                        h = (h + 1) & tmask;
                    }
                }
            }
            if (attempt == 0) {
#pragma unroll
                for (int lv = 0; lv < 4; lv++)
                    if (c >= 1 && c <= (prm.max_mult >> lv)) tally[lv] += c;
            }
            if (c > mult) c = 0;
            int pre = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, pre, d);
                if (lane >= d) pre += t;
            }
            const int tot = __shfl_sync(0xffffffffu, pre, 31);
            pre -= c;
            for (int t = 0; t < c; t++) {
                const int dst = base + pre + t;
                if (dst < MAXA) w.key[dst] = stage[lane * STAGE + t];
            }
            base += tot;
        }
        __syncwarp();
        if (attempt == 1) {
            n = base;
            break;
        }
#pragma unroll
        for (int lv = 0; lv < 4; lv++) tally[lv] = (int)__reduce_add_sync(0xffffffffu, (unsigned)tally[lv]);
        if (tally[0] <= MAXA) {
            n = tally[0];
            break;
        }
        int lv = 1;
        while (lv < 4 && (prm.max_mult >> lv) >= 1 && tally[lv] > MAXA) lv++;
        if (lv >= 4 || (prm.max_mult >> lv) < 1) return;  // nothing fits
        mult = prm.max_mult >> lv;
    }
    if (n < prm.min_anchors) return;

    // ---- 2. sort by (ref_pos, query_pos)
    warp_sort_keys(w.key, n, lane);

    // ---- 3. chaining DP (integer scores)
    for (int i = 0; i < n; i++) {
        const uint64_t ki = w.key[i];
        const uint32_t ri = (uint32_t)(ki >> 32), qi = (uint32_t)(ki & 0xffffffffu) >> 1;
        const uint32_t revi = (uint32_t)(ki & 1);
        const int j = i - 1 - lane;
        unsigned packed = 0;
        if (j >= 0) {
            const uint64_t kj = w.key[j];
            const uint32_t rj = (uint32_t)(kj >> 32), qj = (uint32_t)(kj & 0xffffffffu) >> 1;
            const uint32_t dr = ri - rj;
            if (dr <= (uint32_t)prm.band_bp && dr != 0 && (uint32_t)(kj & 1) == revi) {
                const int dq = revi ? (int)qj - (int)qi : (int)qi - (int)qj;
                if (dq > 0) {
                    int gap = (int)dr - dq;
                    gap = gap < 0 ? -gap : gap;
                    if (gap <= prm.max_gap) {
                        const int cand = w.f[j] + prm.anchor_score - gap;
                        if (cand > prm.anchor_score) packed = ((unsigned)cand << 5) | (unsigned)(31 - lane);
                    }
                }
            }
        }
        const unsigned best = __reduce_max_sync(0xffffffffu, packed);
        if (lane == 0) {
            if (best) {
                const int bj = i - 1 - (31 - (int)(best & 31));
                w.f[i] = (int)(best >> 5);
                w.root[i] = w.root[bj];
                w.cnt[i] = w.cnt[bj] + 1;
            } else {
                w.f[i] = prm.anchor_score;
                w.root[i] = (uint16_t)i;
                w.cnt[i] = 1;
            }
        }
        __syncwarp();
    }

    // ---- 4. best end of every DP tree (ties: lowest index)
    for (int i = lane; i < n; i += 32) w.bor[i] = 0;
    __syncwarp();
    for (int i = lane; i < n; i += 32) atomicMax(&w.bor[w.root[i]], ((uint32_t)w.f[i] << 9) | (uint32_t)(MAXA - 1 - i));
    __syncwarp();
    // candidates owned by this lane: ends i == lane (mod 32) that qualify
    uint32_t mine = 0;  // bit t <-> i = lane + 32 t
    for (int i = lane, t = 0; i < n; i += 32, t++) {
        const uint32_t pk = ((uint32_t)w.f[i] << 9) | (uint32_t)(MAXA - 1 - i);
        if (w.bor[w.root[i]] == pk && w.cnt[i] >= prm.min_anchors && w.f[i] >= prm.min_score) mine |= 1u << t;
    }
    for (int round = 0; round < chunk_cap; round++) {
        // lane-local best: smallest (16383 - score, q0 - chunk_start, r0)
        uint64_t bk = ~0ull;
        int bi = -1;
        for (uint32_t mm = mine; mm; mm &= mm - 1) {
            const int t = __ffs(mm) - 1, i = lane + 32 * t;
            const int rt = w.root[i];
            const uint64_t ke = w.key[i], kr = w.key[rt];
            const uint32_t qe = (uint32_t)(ke & 0xffffffffu) >> 1, qr = (uint32_t)(kr & 0xffffffffu) >> 1;
            const uint32_t q0 = qe < qr ? qe : qr;
            const uint64_t k = ((uint64_t)(16383 - w.f[i]) << 47) | ((uint64_t)(q0 - chunk_start) << 32) |
                               (uint64_t)(uint32_t)(kr >> 32);
            if (k < bk) {
                bk = k;
                bi = i;
            }
        }
        const uint32_t hi = __reduce_min_sync(0xffffffffu, (uint32_t)(bk >> 32));
        if (hi == 0xffffffffu) break;  // no candidate left in any lane
        const uint32_t lo = __reduce_min_sync(0xffffffffu, (uint32_t)(bk >> 32) == hi ? (uint32_t)bk : 0xffffffffu);
        const unsigned who = __ballot_sync(0xffffffffu, bk == (((uint64_t)hi << 32) | lo));
        const int src = __ffs(who) - 1;
        const int wi = __shfl_sync(0xffffffffu, bi, src);
        const int rt = w.root[wi];
        const uint64_t ke = w.key[wi], kr = w.key[rt];
        const uint32_t qe = (uint32_t)(ke & 0xffffffffu) >> 1, qr = (uint32_t)(kr & 0xffffffffu) >> 1;
        const uint32_t q0 = qe < qr ? qe : qr, q1 = qe < qr ? qr : qe;
        // query seeds inside [q0, q1]
        int cs = 0;
        for (int s0 = 0; s0 < nseeds; s0 += 32) {
            const int s = s0 + lane;
            bool in = false;
            if (s < nseeds) {
                const uint32_t p = seed_pos(qs[s]);
                in = p >= q0 && p <= q1;
            }
            cs += __popc(__ballot_sync(0xffffffffu, in));
        }
        if (lane == src) {
            mine &= ~(1u << ((wi - lane) >> 5));
            const int idx = atomicAdd(n_cand, 1);
            if (idx < MAXP) {
                Cand c;
                c.q0 = q0;
                c.q1 = q1;
                c.r0 = (uint32_t)(kr >> 32);
                c.r1 = (uint32_t)(ke >> 32);
                c.chunk = chunk_id;
                c.score = (uint16_t)w.f[wi];
                c.n_anchors = w.cnt[wi];
                c.n_seeds = (uint16_t)cs;
                c.ordinal = (uint8_t)round;
                c.rev = (uint8_t)(ke & 1);
                c.pad = 0;
                cands[idx] = c;
            }
        }
    }
    __syncwarp();
}

__device__ __forceinline__ double block_sum(double v, double *scratch /* [ANI_WARPS] */, int tid) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    __syncthreads();
    if ((tid & 31) == 0) scratch[tid >> 5] = v;
    __syncthreads();
    double r = 0;
    for (int i = 0; i < ANI_WARPS; i++) r += scratch[i];  // fixed order: deterministic
    return r;
}

__global__ void __launch_bounds__(ANI_THREADS, 1)
ani_pair_kernel(DbView db, AniParams prm, const unsigned long long *__restrict__ pairs, int64_t n_pairs,
                PairOut *__restrict__ out, int *work_counter) {
    extern __shared__ __align__(16) unsigned char smem[];
    WarpSlab *slabs = reinterpret_cast<WarpSlab *>(smem);
    Cand *cands = reinterpret_cast<Cand *>(smem + ANI_SMEM_SLABS);
    AniCtl *ctl = reinterpret_cast<AniCtl *>(smem + ANI_SMEM_SLABS + ANI_SMEM_CANDS);
    double *red = reinterpret_cast<double *>(smem + ANI_SMEM_SLABS + ANI_SMEM_CANDS + 96);  // [16]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (;;) {
        __syncthreads();
        if (tid == 0) ctl->next_pair = atomicAdd(work_counter, 1);
        __syncthreads();
        const int64_t pi = ctl->next_pair;
        if (pi >= n_pairs) return;
        const uint32_t ga = (uint32_t)(pairs[pi] >> 32), gb = (uint32_t)(pairs[pi] & 0xffffffffu);
        const uint64_t nsa = db.g_seed_off[ga + 1] - db.g_seed_off[ga];
        const uint64_t nsb = db.g_seed_off[gb + 1] - db.g_seed_off[gb];
        const int swapped = nsb < nsa;  // query = genome with fewer seeds (ties: a)
        const uint32_t gq = swapped ? gb : ga, gr = swapped ? ga : gb;
        const uint64_t *qseeds = db.seeds + db.g_seed_off[gq];
        const uint32_t choff = db.g_chunk_off[gq];
        const uint32_t nch = db.g_chunk_off[gq + 1] - choff;
        const uint32_t *cbeg = db.chunk_begin + choff + gq;
        const uint64_t *T = db.tab + db.g_tab_off[gr];
        const int tbits = db.g_tab_bits[gr];
        const uint32_t tmask = (1u << tbits) - 1;

        int overflow = 0;
        int chunk_cap = prm.max_chunk_chains;
        if (nch > MAX_CHUNKS_PER_GENOME) {
            overflow = 1;
            chunk_cap = 0;
        }
        // ---- chunk phase (retry with a halved per-chunk cap if the pair overflows MAXP)
        while (chunk_cap > 0) {
            if (tid == 0) ctl->n_cand = 0;
            __syncthreads();
            for (uint32_t ch = warp; ch < nch; ch += ANI_WARPS) {
                const uint32_t sb = cbeg[ch], se = cbeg[ch + 1];
                process_chunk(prm, slabs[warp], lane, qseeds + sb, (int)(se - sb), ch, db.chunk_start[choff + ch], T,
                              tmask, tbits, chunk_cap, cands, &ctl->n_cand);
            }
            __syncthreads();
            if (ctl->n_cand <= MAXP) break;
            chunk_cap >>= 1;
            if (chunk_cap == 0) overflow = 1;
            __syncthreads();
        }
        const int nc = (chunk_cap > 0) ? ctl->n_cand : 0;
        __syncthreads();

        // ---- finalize: slab area is free now
        uint64_t *skey = reinterpret_cast<uint64_t *>(smem + FIN_KEYS_OFF);
        uint8_t *state = smem + FIN_STATE_OFF;
        uint32_t *accA = reinterpret_cast<uint32_t *>(smem + FIN_ACC_OFF);
        uint32_t *accS = accA + nch;
        int m = 1;
        while (m < nc) m <<= 1;
        for (int i = tid; i < m; i += ANI_THREADS) {
            if (i < nc) {
                const Cand &c = cands[i];
                skey[i] = ((uint64_t)(16383 - c.score) << 48) | ((uint64_t)c.chunk << 16) |
                          ((uint64_t)c.ordinal << 12) | (uint64_t)i;
            } else
                skey[i] = ~0ull;
            state[i] = 0;
        }
        for (uint32_t i = tid; i < 2 * nch; i += ANI_THREADS) accA[i] = 0;
        if (tid == 0) {
            ctl->span_q = ctl->span_r = ctl->a_tot = ctl->s_tot = 0;
            ctl->used = 0;
        }
        __syncthreads();
        // block bitonic sort of skey[0..m)
        for (int k = 2; k <= m; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = tid; t < (m >> 1); t += ANI_THREADS) {
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const int p = i | j;
                    const bool up = (i & k) == 0;
                    const uint64_t a = skey[i], b = skey[p];
                    if ((a > b) == up) {
                        skey[i] = b;
                        skey[p] = a;
                    }
                }
                __syncthreads();
            }
        }
        // greedy non-overlap selection, resolved in parallel rounds.
        // state: 0 unknown, 1 accepted, 2 rejected.  Candidate at sorted rank t is accepted iff no
        // ACCEPTED candidate of lower rank overlaps more than ovl_num/ovl_den of ITS length on query or ref.
        for (;;) {
            __syncthreads();
            if (tid == 0) ctl->unresolved = 0;
            __syncthreads();
            for (int t = tid; t < nc; t += ANI_THREADS) {
                if (state[t]) continue;
                const Cand &c = cands[(int)(skey[t] & 0xfff)];
                const long long lq = (long long)c.q1 - c.q0 + 1, lr = (long long)c.r1 - c.r0 + 1;
                int verdict = 1;
                for (int u = 0; u < t; u++) {
                    const uint8_t su = ((volatile uint8_t *)state)[u];
                    if (su == 2) continue;
                    const Cand &d = cands[(int)(skey[u] & 0xfff)];
                    const long long oq = (long long)(c.q1 < d.q1 ? c.q1 : d.q1) - (long long)(c.q0 > d.q0 ? c.q0 : d.q0) + 1;
                    const long long orr = (long long)(c.r1 < d.r1 ? c.r1 : d.r1) - (long long)(c.r0 > d.r0 ? c.r0 : d.r0) + 1;
                    const bool blocks = (oq > 0 && oq * prm.ovl_den > lq * prm.ovl_num) ||
                                        (orr > 0 && orr * prm.ovl_den > lr * prm.ovl_num);
                    if (!blocks) continue;
                    if (su == 1) {
                        verdict = 2;
                        break;
                    }
                    verdict = 0;  // blocked by an undecided candidate: wait
                }
                if (verdict)
                    ((volatile uint8_t *)state)[t] = (uint8_t)verdict;
                else
                    ctl->unresolved = 1;
            }
            __syncthreads();
            if (!ctl->unresolved) break;
        }
        // accumulate accepted chains
        const uint32_t rcoff = db.g_ctg_off[gr];
        const int nrc = (int)(db.g_ctg_off[gr + 1] - rcoff);
        int n_acc_local = 0;
        for (int t = tid; t < nc; t += ANI_THREADS) {
            if (state[t] != 1) continue;
            n_acc_local++;
            const Cand &c = cands[(int)(skey[t] & 0xfff)];
            atomicAdd(&accA[c.chunk], (uint32_t)c.n_anchors);
            atomicAdd(&accS[c.chunk], (uint32_t)c.n_seeds);
            const long long e = prm.span_ext, k1 = K_SEED - 1;
            const long long cs = db.chunk_start[choff + c.chunk], ce = cs + db.chunk_len[choff + c.chunk] - 1;
            long long a0 = (long long)c.q0 - k1 - e, a1 = (long long)c.q1 + e;
            a0 = a0 < cs ? cs : a0;
            a1 = a1 > ce ? ce : a1;
            int lo = 0, hi = nrc - 1;  // reference contig holding r0
            while (lo < hi) {
                int mid = (lo + hi + 1) >> 1;
                if (db.ctg_pstart[rcoff + mid] <= c.r0)
                    lo = mid;
                else
                    hi = mid - 1;
            }
            const long long rs = db.ctg_pstart[rcoff + lo], re = rs + db.ctg_len[rcoff + lo] - 1;
            long long b0 = (long long)c.r0 - k1 - e, b1 = (long long)c.r1 + e;
            b0 = b0 < rs ? rs : b0;
            b1 = b1 > re ? re : b1;
            atomicAdd(&ctl->span_q, (unsigned long long)(a1 - a0 + 1));
            atomicAdd(&ctl->span_r, (unsigned long long)(b1 - b0 + 1));
            atomicAdd(&ctl->a_tot, (unsigned long long)c.n_anchors);
            atomicAdd(&ctl->s_tot, (unsigned long long)c.n_seeds);
        }
        if (n_acc_local) atomicAdd(&ctl->used, n_acc_local);  // chains accepted
        __syncthreads();
        // per-chunk ANI, seed-weighted mean
        double sw = 0, sx = 0;
        int used = 0;
        for (uint32_t ch = tid; ch < nch; ch += ANI_THREADS) {
            const uint32_t A = accA[ch], S = accS[ch];
            if ((int)S < prm.min_chunk_seeds || A == 0) continue;
            double ratio = (double)A / (double)S;
            if (ratio > 1.0) ratio = 1.0;
            const double x = pow(ratio, 1.0 / (double)K_SEED);
            sw += (double)S;
            sx += (double)S * x;
            used++;
        }
        sw = block_sum(sw, red, tid);
        sx = block_sum(sx, red, tid);
        const double usedd = block_sum((double)used, red, tid);
        if (tid == 0) {
            PairOut o;
            o.ani = o.ani_raw = -1.0;
            o.af_q = o.af_r = 0.0;
            o.n_anchors = (int64_t)ctl->a_tot;
            o.n_seeds = (int64_t)ctl->s_tot;
            o.span_q = (int64_t)ctl->span_q;
            o.span_r = (int64_t)ctl->span_r;
            o.n_chains = ctl->used;
            o.n_chunks_used = (int)usedd;
            o.swapped = swapped;
            o.overflow = overflow;
            if (usedd > 0 && sw > 0) {
                const double mean = sx / sw;
                o.ani_raw = mean;
                double afq = (double)o.span_q / (double)db.g_total_len[gq];
                double afr = (double)o.span_r / (double)db.g_total_len[gr];
                o.af_q = afq > 1.0 ? 1.0 : afq;
                o.af_r = afr > 1.0 ? 1.0 : afr;
                const double x = 100.0 * (1.0 - mean);
                double ani = 1.0;
                if (x > 0.0) {
                    ani = 1.0 - prm.debias_a * pow(x, prm.debias_g) / 100.0;
                    if (ani < 0.0) ani = 0.0;
                }
                o.ani = ani > 1.0 ? 1.0 : ani;
            }
            out[pi] = o;
        }
    }
}

// K5 -- edge compaction: keep pairs with an estimate and max(AF) >= min_af; percent units
__global__ void edge_compact_kernel(const unsigned long long *__restrict__ pairs, const PairOut *__restrict__ po,
                                    int64_t n_pairs, double min_af /* fraction */, skb_edge *edges,
                                    unsigned long long *n_edges) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool keep = false;
    skb_edge e;
    if (t < n_pairs) {
        const PairOut &o = po[t];
        const double afa = o.swapped ? o.af_r : o.af_q, afb = o.swapped ? o.af_q : o.af_r;
        if (o.ani >= 0.0 && (afa >= min_af || afb >= min_af)) {
            keep = true;
            e.a = (uint32_t)(pairs[t] >> 32);
            e.b = (uint32_t)(pairs[t] & 0xffffffffu);
            e.ani = o.ani * 100.0;
            e.af_a = afa * 100.0;
            e.af_b = afb * 100.0;
        }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (!bal) return;
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(n_edges, (unsigned long long)__popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) edges[base + __popc(bal & ((1u << lane) - 1))] = e;
}

}  // namespace skb
