// K4 -- seed anchoring + per-chunk chaining (chunk kernel), chain selection + ANI/AF (finalize kernel).
//
// Stands in for skani's pairwise estimator behind `skani triangle|dist|search`
// (reference call sites src/skDER/skder.py:16-18, :58-59, :119).  Integer results (anchors, seeds,
// spans, chain count) are bit-exact against oracle/skani_oracle.c ora_pair(); ANI/AF are the same
// IEEE double expressions (device pow() may differ from glibc's in the last ulp).
//
// A task = (surviving pair, 20 kb query chunk).  Three kernels, each at the occupancy its bottleneck
// wants; anchors and DP results go through a scratch array in HBM (written once, read once: the
// 2*16*A term of the roofline model).
//   K4a anchor_kernel (warp per task, many warps/SM): streams the chunk's position-ordered seeds
//      (coalesced 8-byte records), probes the reference's bucketed hash index (L2-resident: a
//      clade's members all hit the same tables) and writes the anchors in QUERY order -- the order
//      the DP wants, so there is no sort.
//   K4b chain_kernel (thread per task): the chaining DP with its 16-anchor look-back window in
//      registers; no shuffles, no shared memory, ~11 instructions per (anchor, predecessor), 32 tasks
//      per warp instruction.
//   K4c ends_kernel (warp per task): best end of every DP tree with >= min_anchors / min_score; the
//      chunk's top `max_chunk_chains` by (score, q0, r0) go to the task's fixed candidate slots.
// finalize_kernel: one CTA per pair gathers the candidates, orders them by (score desc, chunk,
//   ordinal), resolves the greedy non-overlap selection in parallel rounds, sums anchors, seeds and the
//   symmetrically clipped spans of the accepted chains, and reduces ANI = ((A - 2n) / (S - 2n))^(1/15) (n chains,
//   end anchors left out), AF = span / genome length.  Pairs with more than MAXP candidates run the same code on
//   global scratch (finalize_kernel<true>).
#pragma once
#include "skb_common.cuh"
#include "skb_index.cuh"

namespace skb {

constexpr int LB = 16;                        // DP look-back in anchors
constexpr int DP_UNR = 4;                     // anchors per DP iteration (code size vs register moves)
constexpr int DIAG_SLACK = 64;                // chain_kernel short cut: diagonal drift followed without a rebuild
constexpr int NEG_F = -(1 << 24);             // DP score of an empty window slot
constexpr int ANC_THREADS = 256;              // K4a: 8 warps = 8 tasks per CTA pass
constexpr int DP_THREADS = 128;               // K4b: one task per thread
constexpr int END_THREADS = 256;              // K4c: 8 warps = 8 tasks per CTA pass
constexpr int MAXA = 256;                     // anchors per chunk
constexpr int MAXP = 4096;                    // chain candidates per pair the shared-memory finalize kernel takes
constexpr int STAGE = 8;                      // max_mult upper bound (hits staged per seed)
constexpr int ENDS_K = 8;                     // qualifying DP trees per chunk tracked inside chain_kernel
constexpr int SLOTS = 8;                      // candidate slots per task (max_chunk_chains upper bound)
constexpr int FIN_THREADS = 256;

struct AniParams {
    int32_t band_bp, max_gap, anchor_score, min_anchors, min_score, max_mult, max_chunk_chains;
    int32_t ovl_num, ovl_den, span_ext;
    double debias_a, debias_g;
};

struct PairInfo {
    uint32_t q, r;  // query / reference genome ids
    uint32_t swapped;
    uint32_t nch;   // chunks of the query
};

// anchor record: ref_pos(32) | q_rel(15) | rev(1) | seed index in chunk(16)
__device__ __forceinline__ uint32_t an_r(uint64_t a) { return (uint32_t)(a >> 32); }
__device__ __forceinline__ uint32_t an_q(uint64_t a) { return ((uint32_t)a >> 17) & 0x7fffu; }
__device__ __forceinline__ uint32_t an_rev(uint64_t a) { return ((uint32_t)a >> 16) & 1u; }
__device__ __forceinline__ uint32_t an_sidx(uint64_t a) { return (uint32_t)a & 0xffffu; }
// DP result record: f(13) << 17 | root(8) << 9 | cnt(9)
__device__ __forceinline__ uint32_t rs_f(uint32_t x) { return x >> 17; }
__device__ __forceinline__ uint32_t rs_root(uint32_t x) { return (x >> 9) & 0xffu; }
__device__ __forceinline__ uint32_t rs_cnt(uint32_t x) { return x & 0x1ffu; }

struct __align__(16) Cand {
    uint32_t q0, q1, r0, r1;
    uint32_t chunk;
    uint16_t score, n_anchors;
    uint16_t n_seeds;
    uint8_t ordinal, rev;
    uint32_t pad;
};
static_assert(sizeof(Cand) == 32, "cand size");
static_assert(STAGE == STAGE_CAP, "stage layout");

// roles + sort key per pair (one thread per pair)
__global__ void pair_setup_kernel(DbView db, const unsigned long long *__restrict__ pairs, int64_t n_pairs,
                                  PairInfo *__restrict__ info, unsigned long long *__restrict__ sort_key,
                                  uint32_t *__restrict__ idx) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    const uint32_t ga = (uint32_t)(pairs[p] >> 32), gb = (uint32_t)(pairs[p] & 0xffffffffu);
    const uint64_t nsa = db.g_seed_off[ga + 1] - db.g_seed_off[ga];
    const uint64_t nsb = db.g_seed_off[gb + 1] - db.g_seed_off[gb];
    PairInfo pi;
    pi.swapped = nsb < nsa;  // query = genome with fewer seeds (ties: a)
    pi.q = pi.swapped ? gb : ga;
    pi.r = pi.swapped ? ga : gb;
    pi.nch = db.g_chunk_off[pi.q + 1] - db.g_chunk_off[pi.q];
    info[p] = pi;
    sort_key[p] = ((unsigned long long)pi.r << 32) | pi.q;  // pairs sharing a reference index run together (L2 reuse)
    idx[p] = (uint32_t)p;
}

// pairs in reference-major order: info_sorted[i] = info[perm[i]]
__global__ void pair_gather_kernel(const PairInfo *__restrict__ info, const uint32_t *__restrict__ perm, int64_t n_pairs,
                                   PairInfo *__restrict__ info_sorted, uint32_t *__restrict__ nch_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    const PairInfo pi = info[perm[i]];
    info_sorted[i] = pi;
    nch_out[i] = pi.nch;
}

// task descriptors: the (pair, chunk) decode is a chain of ~20 dependent loads, so it is done once,
// one thread per task, instead of by every warp in front of its two useful loads
struct __align__(16) TaskDesc {
    uint64_t seed_idx;  // first seed record of the chunk (index into db.seeds)
    uint64_t tab_idx;   // first slot of the reference's table (index into db.tab)
    uint32_t nseeds, cstart, nb, ch;
};
static_assert(sizeof(TaskDesc) == 32, "desc size");

// task_off: GLOBAL exclusive scan of the chunk counts, pointing at this batch's first pair; base = its value there
__global__ void task_setup_kernel(DbView db, const PairInfo *__restrict__ info, const uint32_t *__restrict__ task_off,
                                  uint32_t base, int64_t n_pairs, uint32_t n_tasks, TaskDesc *__restrict__ desc) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tasks) return;
    int64_t lo = 0, hi = n_pairs - 1;
    while (lo < hi) {
        const int64_t mid = (lo + hi + 1) >> 1;
        if (task_off[mid] - base <= t)
            lo = mid;
        else
            hi = mid - 1;
    }
    const PairInfo pi = info[lo];
    TaskDesc d;
    d.ch = t - (task_off[lo] - base);
    const uint32_t choff = db.g_chunk_off[pi.q];
    const uint32_t *cbeg = db.chunk_begin + choff + pi.q;
    const uint32_t sb = cbeg[d.ch];
    d.nseeds = cbeg[d.ch + 1] - sb;
    d.seed_idx = db.g_seed_off[pi.q] + sb;
    d.cstart = db.chunk_start[choff + d.ch];
    d.tab_idx = db.g_tab_off[pi.r];
    d.nb = db.g_tab_buckets[pi.r];
    desc[t] = d;
}

// ---- K4a: anchors.  One warp per task; the only dependent reads are descriptor -> seed records ->
// buckets, covered by the other resident warps.
#ifndef SKB_ANC_MIN_CTAS
#define SKB_ANC_MIN_CTAS 4
#endif
__global__ void __launch_bounds__(ANC_THREADS, SKB_ANC_MIN_CTAS)
anchor_kernel(DbView db, AniParams prm, const TaskDesc *__restrict__ desc, uint32_t n_tasks,
              uint64_t *__restrict__ anc_all, uint16_t *__restrict__ task_n, uint32_t *__restrict__ next_task) {
    __shared__ uint32_t stage_all[ANC_THREADS / 32][32 * STAGE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *stage = stage_all[warp];
    // Tasks are handed out in order from one counter, not strided by CTA: the grid is persistent and, next to the
    // consumer kernels of the previous batch, its CTAs become resident at different times; a CTA that starts late
    // must join the others on the reference tables that are in L2 NOW (tasks are reference-major), not work
    // through a fixed share milliseconds behind them (measured: 11 ms -> 54 ms per launch with a strided loop).
    for (;;) {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(next_task, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n_tasks) break;
        const TaskDesc d = desc[t];
        const int nseeds = (int)d.nseeds;
        int n = 0;
        if (nseeds > 0) {
            const uint64_t *qs = db.seeds + d.seed_idx;
            const uint64_t *T = db.tab + d.tab_idx;
            uint64_t *anc = anc_all + (size_t)t * MAXA;
            uint64_t sdn[PJ];
            // optimistic pass with the full multiplicity cap; if the chunk overflows MAXA, halve the cap
            // until it fits (oracle rule; rare)
            int mult = prm.max_mult;
            for (;;) {
                load_first_batch(qs, nseeds, lane, sdn);
                n = emit_anchors(qs, nseeds, d.cstart, T, d.nb, mult, prm.max_mult, MAXA, stage, anc, lane, sdn, qs, 0);
                if (n <= MAXA || mult == 1) break;
                mult >>= 1;
            }
            if (n > MAXA || n < prm.min_anchors) n = 0;
        }
        if (lane == 0) task_n[t] = (uint16_t)n;
        __syncwarp();
    }
}

// ---- K4b: chaining DP, one THREAD per task.  The look-back window lives in registers and ROTATES:
//   A[UNR + d] = d-th previous anchor (d = 0 nearest); the UNR anchors of one iteration sit in
//   A[UNR-1 .. 0]; after the iteration everything shifts by UNR.  (A 16x unrolled static window needs
//   no moves but is ~50 KB of code: ncu showed 64% of stalls on instruction fetch.)
//   Q = q_rel + (rev << 20) + 1: other strand relation => more than band apart, no extra test;
//       the +1 makes (Qi - Q[j]) == dq - 1, so one unsigned compare checks 0 < dq <= band
//   D = R - q with R = rev ? -ref_pos : ref_pos: gap = |D_i - D_j|, d_ref = (D_i - D_j) + dq
//   F = f + anchor_score
__global__ void __launch_bounds__(DP_THREADS, 4)
chain_kernel(AniParams prm, uint32_t n_tasks, const uint64_t *anc_all, const uint16_t *__restrict__ task_n,
             uint32_t *__restrict__ res_all, const TaskDesc *__restrict__ desc, Cand *__restrict__ cands,
             uint8_t *__restrict__ task_ncand, uint8_t *__restrict__ task_slow) {
    const uint32_t t = blockIdx.x * DP_THREADS + threadIdx.x;
    const int my_n = t < n_tasks ? (int)task_n[t] : 0;
    const int w_n = (int)__reduce_max_sync(0xffffffffu, (unsigned)my_n);
    if (w_n == 0) return;
    constexpr int UNR = DP_UNR;
    static_assert(UNR % 2 == 0 && MAXA % UNR == 0, "anchors are fetched 16 bytes at a time");
    int Q[LB + UNR], D[LB + UNR], F[LB + UNR];
    uint32_t RC[LB + UNR];
#pragma unroll
    for (int u = 0; u < LB + UNR; u++) {
        Q[u] = 0;
        D[u] = 0;
        F[u] = NEG_F;  // an empty slot can never win
        RC[u] = 0;
    }
    const uint32_t tt = t < n_tasks ? t : n_tasks - 1;  // idle lanes read a valid slab and write nothing
    const uint64_t *ap = anc_all + (size_t)tt * MAXA;
    uint32_t *rp = res_all + (size_t)tt * MAXA;
    const unsigned band = (unsigned)prm.band_bp;
    // chain ends, tracked on the fly: per DP tree (root) the best qualifying end as
    //   (f << 16) | (255 - i) << 8 | root   -- max over a tree = highest score, ties lowest index.
    // Anchors of a tree arrive mostly back to back, so the tree in progress sits in (cur_root, cur_e) and the
    // table of finished trees is touched only when the root changes.  More than ENDS_K qualifying trees in one
    // chunk (rare) hands the task to ends_kernel.  Valid because min_score > (min_anchors - 1) * anchor_score
    // (checked at context creation): an end that reaches min_score has min_anchors anchors, and if the tree's
    // best end does not qualify nothing in the tree does.
    __shared__ uint32_t tb_s[ENDS_K][DP_THREADS];  // dynamic indexing without local memory; column per thread
    uint32_t *const tb = &tb_s[0][threadIdx.x];
#pragma unroll
    for (int k = 0; k < ENDS_K; k++) tb[k * DP_THREADS] = 0;
    uint32_t cur_root = 0xffffffffu, cur_e = 0;
    bool slow = false;
    auto flush = [&](uint32_t e) {  // kept small on purpose: it is inlined once per unrolled anchor (I-cache)
        if (e == 0) return;
        int k = 0;
#pragma unroll 1
        for (; k < ENDS_K; k++) {  // entries fill front to back and never leave: first empty-or-same-root slot
            const uint32_t v = tb[k * DP_THREADS];
            if (v == 0 || ((v ^ e) & 0xffu) == 0) {
                tb[k * DP_THREADS] = v > e ? v : e;
                break;
            }
        }
        if (k == ENDS_K) slow = true;
    };
    int Mx = NEG_F, Dref = 0;  // short cut state, see below
    const int diag_lim = prm.max_gap + DIAG_SLACK;
    ulonglong2 vnext[UNR / 2];
#pragma unroll
    for (int x = 0; x < UNR / 2; x++) vnext[x] = __ldcg(reinterpret_cast<const ulonglong2 *>(ap) + x);
    for (int i0 = 0; i0 < w_n; i0 += UNR) {
        uint64_t av[UNR];
#pragma unroll
        for (int x = 0; x < UNR / 2; x++) {
            av[2 * x] = vnext[x].x;
            av[2 * x + 1] = vnext[x].y;
        }
        if (i0 + UNR < MAXA) {  // next iteration's anchors
#pragma unroll
            for (int x = 0; x < UNR / 2; x++) vnext[x] = __ldcg(reinterpret_cast<const ulonglong2 *>(ap + i0 + UNR) + x);
        }
        uint32_t outp[UNR];
#pragma unroll
        for (int x = 0; x < UNR; x++) {
            const uint64_t a = av[x];
            const int rev = (int)an_rev(a);
            const int qi = (int)an_q(a) + (rev << 20);
            const int Ri = rev ? -(int)an_r(a) : (int)an_r(a);
            const int Di = Ri - qi;
            int best = prm.anchor_score;
            uint32_t brc = (uint32_t)(i0 + x) << 9;  // own root, cnt 0 (+1 below)
            const int me = UNR - 1 - x;               // this anchor's slot
            const bool live = i0 + x < my_n;
            auto relax = [&](int sl) {
                const int dq1 = qi - Q[sl];  // dq - 1
                const int dd = Di - D[sl];
                const int dr1 = dd + dq1;    // d_ref - 1
                const int gap = dd < 0 ? -dd : dd;
                const int cand = F[sl] - gap;
                if ((unsigned)dq1 < band && dr1 >= 0 && gap <= prm.max_gap && cand > best) {
                    best = cand;
                    brc = RC[sl];
                }
            };
            relax(me + 1);  // d = 0, the nearest predecessor: ties keep it
            relax(me + 2);  // d = 1
            // Short cut.  Colinear anchors (the usual case) chain onto one of the two nearest predecessors with a score
            // no other one can reach; then the other 14 are not looked at.  (Two, because a chance anchor -- FracMinHash
            // samples the same 1/125 of k-mer space in both genomes, so ~1.7 % of a pair's anchors are random matches
            // on a random diagonal -- sits between two anchors of a chain, and the one after it must reach over it.)
            // Mx bounds what predecessors beyond the nearest two can offer: the largest F among those whose diagonal
            // lies within max_gap + DIAG_SLACK of Dref, where Dref follows this task's current diagonal to within
            // DIAG_SLACK.  An anchor further off than that is out of max_gap for this one, and one inside offers
            // F - gap <= Mx: if Mx <= best nothing replaces `best` (a later predecessor only wins with a strictly
            // larger candidate) -- the same result as the full scan.  An anchor off the current diagonal gets its own
            // bound (m2) from the window; the diagonal is adopted as the new Dref only when the previous anchor lies
            // on it too (contig end, indel, rearrangement -- not a chance anchor), so the old chain's high scores
            // neither keep the short cut off for the next 16 anchors nor does one stray anchor derail it.  Stale
            // entries that have left the window only make Mx too large (a missed short cut).
            long long dj = (long long)Di - Dref;  // 64-bit: opposite strands are up to 2^32 apart
            dj = dj < 0 ? -dj : dj;
            const bool jump = live && dj > DIAG_SLACK;
            bool settled = !live || Mx <= best;
            if (__any_sync(0xffffffffu, jump)) {
                int m2 = NEG_F;
#pragma unroll
                for (int d = 2; d < LB; d++) {
                    int dd = D[me + 1 + d] - Di;
                    dd = dd < 0 ? -dd : dd;
                    m2 = max(m2, dd <= diag_lim ? F[me + 1 + d] : NEG_F);
                }
                if (jump) {
                    settled = m2 <= best;
                    long long dp = (long long)Di - D[me + 1];
                    dp = dp < 0 ? -dp : dp;
                    if (dp <= DIAG_SLACK) {  // the previous anchor is on this diagonal too: follow it
                        Mx = m2;
                        Dref = Di;
                    }
                }
            }
            if (!__all_sync(0xffffffffu, settled)) {
#pragma unroll
                for (int d = 2; d < LB; d++) relax(me + 1 + d);
            }
            {  // the second nearest predecessor is beyond the nearest two of the next anchor
                int dd = D[me + 2] - Dref;
                dd = dd < 0 ? -dd : dd;
                Mx = max(Mx, dd <= diag_lim ? F[me + 2] : NEG_F);
            }
            const uint32_t rci = brc + 1;
            outp[x] = ((uint32_t)best << 17) | rci;
            if (live && best >= prm.min_score && (int)(rci & 0x1ffu) >= prm.min_anchors) {
                const uint32_t root = rci >> 9;
                const uint32_t e = ((uint32_t)best << 16) | ((uint32_t)(MAXA - 1 - (i0 + x)) << 8) | root;
                if (root == cur_root)
                    cur_e = cur_e > e ? cur_e : e;
                else {
                    flush(cur_e);
                    cur_root = root;
                    cur_e = e;
                }
            }
            Q[me] = qi + 1;
            D[me] = Di;
            F[me] = live ? best + prm.anchor_score : NEG_F;
            RC[me] = rci;
        }
        if (i0 < my_n) {
#pragma unroll
            for (int x = 0; x < UNR / 2; x++)
                __stcg(reinterpret_cast<uint2 *>(rp + i0) + x, make_uint2(outp[2 * x], outp[2 * x + 1]));
        }
#pragma unroll
        for (int u = LB + UNR - 1; u >= UNR; u--) {
            Q[u] = Q[u - UNR];
            D[u] = D[u - UNR];
            F[u] = F[u - UNR];
            RC[u] = RC[u - UNR];
        }
    }
    // ---- the chunk's top candidates, by (score desc, q0, r0), into the task's slots
    if (my_n == 0) return;
    flush(cur_e);
    if (slow) {
        task_slow[t] = 1;
        return;
    }
    uint64_t key[ENDS_K];
    int m = 0;
#pragma unroll
    for (int k = 0; k < ENDS_K; k++) {
        key[k] = ~0ull;
        const uint32_t e = tb[k * DP_THREADS];
        if (e) {
            const uint64_t ar = __ldcg(ap + (e & 0xffu));
            const uint64_t ae = __ldcg(ap + (MAXA - 1 - ((e >> 8) & 0xffu)));
            const uint32_t r0 = an_r(ar) < an_r(ae) ? an_r(ar) : an_r(ae);
            key[k] = ((uint64_t)(8191u - (e >> 16)) << 47) | ((uint64_t)an_q(ar) << 32) | (uint64_t)r0;
            m++;
        }
    }
    const TaskDesc d = desc[t];
    const int n_out = m < prm.max_chunk_chains ? m : prm.max_chunk_chains;
    for (int ord = 0; ord < n_out; ord++) {
        int bk = 0;
        uint64_t bkey = key[0];
#pragma unroll
        for (int k = 1; k < ENDS_K; k++)
            if (key[k] < bkey) {
                bkey = key[k];
                bk = k;
            }
#pragma unroll
        for (int k = 0; k < ENDS_K; k++)
            if (k == bk) key[k] = ~0ull;
        const uint32_t e = tb[bk * DP_THREADS];
        const int iend = MAXA - 1 - (int)((e >> 8) & 0xffu);
        const uint64_t ar = __ldcg(ap + (e & 0xffu)), ae = __ldcg(ap + iend);
        const uint32_t x = __ldcg(rp + iend);  // written above by this thread
        Cand c;
        c.q0 = d.cstart + an_q(ar);
        c.q1 = d.cstart + an_q(ae);
        c.r0 = an_r(ar) < an_r(ae) ? an_r(ar) : an_r(ae);
        c.r1 = an_r(ar) < an_r(ae) ? an_r(ae) : an_r(ar);
        c.chunk = d.ch;
        c.score = (uint16_t)(e >> 16);
        c.n_anchors = (uint16_t)rs_cnt(x);
        c.n_seeds = (uint16_t)(an_sidx(ae) - an_sidx(ar) + 1);
        c.ordinal = (uint8_t)ord;
        c.rev = (uint8_t)an_rev(ae);
        c.pad = 0;
        cands[(size_t)t * SLOTS + ord] = c;
    }
    if (n_out) task_ncand[t] = (uint8_t)n_out;
}

// ---- K4c: chain ends (fallback for chunks with more than ENDS_K qualifying trees).  One warp per task: best end of every DP tree (ties: lowest index) with
// >= min_anchors / min_score; the chunk's top `max_chunk_chains` by (score, q0, r0) go to the task's
// fixed candidate slots.
__global__ void __launch_bounds__(END_THREADS)
ends_kernel(AniParams prm, uint32_t n_tasks, const uint64_t *__restrict__ anc_all, const uint32_t *__restrict__ res_all,
            const uint16_t *__restrict__ task_n, const TaskDesc *__restrict__ desc, const uint8_t *__restrict__ task_slow,
            Cand *__restrict__ cands, uint8_t *__restrict__ task_ncand) {
    __shared__ uint32_t bor_all[END_THREADS / 32][MAXA];
    __shared__ __align__(16) uint32_t lst_all[END_THREADS / 32][SLOTS * 8 + SLOTS * 2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *bor = bor_all[warp];
    const uint32_t warps_total = gridDim.x * (END_THREADS / 32);
    for (uint32_t t = blockIdx.x * (END_THREADS / 32) + warp; t < n_tasks; t += warps_total) {
        if (!task_slow[t]) continue;  // chain_kernel already wrote this task's candidates
        const int n = task_n[t];
        if (n == 0) continue;
        const uint32_t ch_k = desc[t].ch, cstart_k = desc[t].cstart;
        const uint64_t *anc = anc_all + (size_t)t * MAXA;
        const uint32_t *res = res_all + (size_t)t * MAXA;
        uint32_t xs[MAXA / 32];
#pragma unroll
        for (int u = 0; u < MAXA / 32; u++) {
            const int i = lane + 32 * u;
            xs[u] = i < n ? __ldcg(res + i) : 0u;
            if (i < n) bor[i] = 0;
        }
        __syncwarp();
#pragma unroll
        for (int u = 0; u < MAXA / 32; u++) {
            const int i = lane + 32 * u;
            if (i < n) atomicMax(&bor[rs_root(xs[u])], (rs_f(xs[u]) << 8) | (uint32_t)(MAXA - 1 - i));
        }
        __syncwarp();
        uint32_t mine = 0;  // bit u <-> i = lane + 32 u
#pragma unroll
        for (int u = 0; u < MAXA / 32; u++) {
            const int i = lane + 32 * u;
            const uint32_t x = xs[u];
            if (i < n && bor[rs_root(x)] == ((rs_f(x) << 8) | (uint32_t)(MAXA - 1 - i)) &&
                (int)rs_cnt(x) >= prm.min_anchors && (int)rs_f(x) >= prm.min_score)
                mine |= 1u << u;
        }
        __syncwarp();
        const int my_cnt = __popc(mine);
        const int total = (int)__reduce_add_sync(0xffffffffu, (unsigned)my_cnt);
        if (total == 0) continue;
        if (total <= prm.max_chunk_chains) {
            // common case: every qualifying end becomes a candidate.  All lanes fetch their ends'
            // anchors at once, publish (candidate, key) to shared memory, then `total` lanes rank the
            // keys for the ordinals.
            int pre = my_cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, pre, d);
                if (lane >= d) pre += u;
            }
            pre -= my_cnt;
            Cand *lst = reinterpret_cast<Cand *>(lst_all[warp]);                  // [SLOTS]
            uint64_t *lkey = reinterpret_cast<uint64_t *>(lst_all[warp] + SLOTS * 8);  // [SLOTS]
#pragma unroll
            for (int u = 0; u < MAXA / 32; u++) {
                if ((mine >> u) & 1u) {
                    const int i = lane + 32 * u;
                    const uint32_t x = xs[u];
                    const uint64_t ar = __ldcg(anc + rs_root(x)), ae = __ldcg(anc + i);
                    Cand c;
                    c.q0 = cstart_k + an_q(ar);
                    c.q1 = cstart_k + an_q(ae);
                    c.r0 = an_r(ar) < an_r(ae) ? an_r(ar) : an_r(ae);
                    c.r1 = an_r(ar) < an_r(ae) ? an_r(ae) : an_r(ar);
                    c.chunk = ch_k;
                    c.score = (uint16_t)rs_f(x);
                    c.n_anchors = (uint16_t)rs_cnt(x);
                    c.n_seeds = (uint16_t)(an_sidx(ae) - an_sidx(ar) + 1);
                    c.ordinal = 0;
                    c.rev = (uint8_t)an_rev(ae);
                    c.pad = 0;
                    lst[pre] = c;
                    lkey[pre] = ((uint64_t)(8191u - rs_f(x)) << 47) | ((uint64_t)an_q(ar) << 32) | (uint64_t)c.r0;
                    pre++;
                }
            }
            __syncwarp();
            if (lane < total) {
                const uint64_t key = lkey[lane];
                int ord = 0;
                for (int j = 0; j < total; j++) ord += lkey[j] < key;  // keys are unique
                Cand c = lst[lane];
                c.ordinal = (uint8_t)ord;
                cands[(size_t)t * SLOTS + ord] = c;
            }
            if (lane == 0) task_ncand[t] = (uint8_t)total;
            __syncwarp();
            continue;
        }
        // rare: more qualifying ends than slots -> take the best max_chunk_chains, one per round
        int n_out = 0;
        for (int rnd = 0; rnd < prm.max_chunk_chains; rnd++) {
            uint64_t bk = ~0ull;
            int bi = -1;
            for (uint32_t mm = mine; mm; mm &= mm - 1) {
                const int u = __ffs(mm) - 1, i = lane + 32 * u;
                const uint32_t x = __ldcg(res + i);
                const uint64_t ar = __ldcg(anc + rs_root(x)), ae = __ldcg(anc + i);
                const uint32_t r0 = an_r(ar) < an_r(ae) ? an_r(ar) : an_r(ae);
                const uint64_t key = ((uint64_t)(8191u - rs_f(x)) << 47) | ((uint64_t)an_q(ar) << 32) | (uint64_t)r0;
                if (key < bk) {
                    bk = key;
                    bi = i;
                }
            }
            const uint32_t khi = __reduce_min_sync(0xffffffffu, (uint32_t)(bk >> 32));
            if (khi == 0xffffffffu) break;  // no candidate left
            const uint32_t klo = __reduce_min_sync(0xffffffffu, (uint32_t)(bk >> 32) == khi ? (uint32_t)bk : 0xffffffffu);
            if (bk == (((uint64_t)khi << 32) | klo)) {  // keys are unique: exactly one lane
                mine &= ~(1u << ((bi - lane) >> 5));
                const uint32_t x = __ldcg(res + bi);
                const uint64_t ar = __ldcg(anc + rs_root(x)), ae = __ldcg(anc + bi);
                Cand c;
                c.q0 = cstart_k + an_q(ar);
                c.q1 = cstart_k + an_q(ae);
                c.r0 = an_r(ar) < an_r(ae) ? an_r(ar) : an_r(ae);
                c.r1 = an_r(ar) < an_r(ae) ? an_r(ae) : an_r(ar);
                c.chunk = ch_k;
                c.score = (uint16_t)rs_f(x);
                c.n_anchors = (uint16_t)rs_cnt(x);
                c.n_seeds = (uint16_t)(an_sidx(ae) - an_sidx(ar) + 1);
                c.ordinal = (uint8_t)n_out;
                c.rev = (uint8_t)an_rev(ae);
                c.pad = 0;
                cands[(size_t)t * SLOTS + n_out] = c;
            }
            n_out++;
        }
        if (lane == 0 && n_out) task_ncand[t] = (uint8_t)n_out;
        __syncwarp();
    }
}

__device__ __forceinline__ double block_sum(double v, double *scratch /* [FIN_THREADS/32] */, int tid) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    __syncthreads();
    if ((tid & 31) == 0) scratch[tid >> 5] = v;
    __syncthreads();
    double r = 0;
    for (int i = 0; i < FIN_THREADS / 32; i++) r += scratch[i];  // fixed order: deterministic
    return r;
}

struct FinCtl {
    int n_cand, unresolved;
    double red[FIN_THREADS / 32];
};

// Candidates a pair brings to the selection (sum of its chunks' candidate counts).  Pairs with at most MAXP go
// through the shared-memory finalize kernel, whose launch is sized for the largest of them in the batch
// (max_small); the others -- a query of thousands of contigs against a close relative -- are listed for the
// global-memory instance (finalize_kernel<true>): no pair is ever dropped or capped.
__global__ void pair_ncand_kernel(const PairInfo *__restrict__ info, const uint32_t *__restrict__ task_off, uint32_t base,
                                  int64_t n_pairs, const uint8_t *__restrict__ task_ncand, uint32_t *__restrict__ pair_nc,
                                  uint32_t *ctl /* [0] max over small pairs, [1] number of big pairs */,
                                  uint32_t *__restrict__ big_list, uint32_t *__restrict__ big_nc) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    const uint32_t t0 = task_off[p] - base, nch = info[p].nch;
    uint32_t nc = 0;
    for (uint32_t ch = 0; ch < nch; ch++) nc += task_ncand[t0 + ch];
    pair_nc[p] = nc;
    if (nc <= (uint32_t)MAXP)
        atomicMax(&ctl[0], nc);
    else {
        const uint32_t k = atomicAdd(&ctl[1], 1u);
        big_list[k] = (uint32_t)p;
        big_nc[k] = nc;
    }
}

constexpr size_t FIN_BYTES_PER_CAND = sizeof(Cand) + 8 + 1;  // candidate + sort key + state

// sort key of a candidate: score desc, chunk, ordinal within the chunk (= (q0, r0) order), then its slot
//   (8191 - score)(13) << 51 | chunk(23) << 28 | ordinal(4) << 24 | slot(24)
constexpr uint32_t FIN_IDX_MASK = 0xffffffu;
constexpr uint32_t MAX_CHUNKS_PER_GENOME = 1u << 23;

// BIG = false: one CTA per pair of the batch, candidates / keys / states in dynamic shared memory sized for `cap`
//   candidates (a power of two >= the batch's largest small pair): Cand[cap] | u64 key[cap] | u8 state[cap].
// BIG = true: one CTA per listed pair, the same three arrays in global scratch at big_off[blockIdx.x] (bytes),
//   capacity big_cap[blockIdx.x].
template <bool BIG>
__global__ void __launch_bounds__(FIN_THREADS)
finalize_kernel(DbView db, AniParams prm, const PairInfo *__restrict__ info, const uint32_t *__restrict__ task_off,
                uint32_t base, int64_t n_pairs, const Cand *__restrict__ gcands, const uint8_t *__restrict__ task_ncand,
                const uint32_t *__restrict__ pair_nc, const uint32_t *__restrict__ perm, PairOut *__restrict__ out,
                uint32_t cap, const uint32_t *__restrict__ big_list, const unsigned long long *__restrict__ big_off,
                const uint32_t *__restrict__ big_cap, unsigned char *__restrict__ big_scratch) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ FinCtl ctl;
    const int tid = threadIdx.x;
    int64_t p = blockIdx.x;
    unsigned char *buf = smem;
    if (BIG) {
        p = big_list[blockIdx.x];
        cap = big_cap[blockIdx.x];
        buf = big_scratch + big_off[blockIdx.x];
    }
    if (p >= n_pairs) return;
    const int nc = (int)pair_nc[p];
    if (!BIG && nc > MAXP) return;  // finalize_kernel<true> does this pair
    Cand *cands = reinterpret_cast<Cand *>(buf);
    uint64_t *skey = reinterpret_cast<uint64_t *>(buf + sizeof(Cand) * (size_t)cap);
    uint8_t *state = buf + (sizeof(Cand) + 8) * (size_t)cap;
    const PairInfo pi = info[p];
    const uint32_t nch = pi.nch, t0 = task_off[p] - base;
    const uint32_t choff = db.g_chunk_off[pi.q];
    if (tid == 0) ctl.n_cand = 0;
    __syncthreads();
    for (uint32_t ch = tid; ch < nch; ch += FIN_THREADS) {
        const int c = task_ncand[t0 + ch];
        if (c) {
            const int at = atomicAdd(&ctl.n_cand, c);
            for (int x = 0; x < c; x++) cands[at + x] = gcands[(size_t)(t0 + ch) * SLOTS + x];
        }
    }
    __syncthreads();
    int m = 1;
    while (m < nc) m <<= 1;
    for (int i = tid; i < m; i += FIN_THREADS) {
        if (i < nc) {
            const Cand &c = cands[i];
            skey[i] = ((uint64_t)(8191u - c.score) << 51) | ((uint64_t)c.chunk << 28) | ((uint64_t)c.ordinal << 24) |
                      (uint64_t)i;
        } else
            skey[i] = ~0ull;
        state[i] = 0;
    }
    __syncthreads();
    for (int k = 2; k <= m; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (m >> 1); t += FIN_THREADS) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int q = i | j;
                const bool up = (i & k) == 0;
                const uint64_t a = skey[i], b = skey[q];
                if ((a > b) == up) {
                    skey[i] = b;
                    skey[q] = a;
                }
            }
            __syncthreads();
        }
    }
    // greedy non-overlap selection, resolved in parallel rounds.
    // state: 0 unknown, 1 accepted, 2 rejected.  The candidate at sorted rank t is accepted iff no
    // ACCEPTED candidate of lower rank overlaps more than ovl_num/ovl_den of ITS length on query or ref.
    for (;;) {
        __syncthreads();
        if (tid == 0) ctl.unresolved = 0;
        __syncthreads();
        // rank t costs t comparisons: give every thread one cheap and one expensive rank (tt, nc-1-tt)
        for (int w0 = tid; 2 * w0 < nc; w0 += FIN_THREADS) {
            for (int side = 0; side < 2; side++) {
                const int t = side ? nc - 1 - w0 : w0;
                if (side && t == w0) break;
                if (state[t]) continue;
                const uint4 cr = *reinterpret_cast<const uint4 *>(&cands[(int)(skey[t] & FIN_IDX_MASK)]);  // q0 q1 r0 r1
                const long long lq = (long long)cr.y - cr.x + 1, lr = (long long)cr.w - cr.z + 1;
                int verdict = 1;
                for (int u = 0; u < t; u++) {
                    const uint4 dr = *reinterpret_cast<const uint4 *>(&cands[(int)(skey[u] & FIN_IDX_MASK)]);
                    // intervals that do not even touch (the usual case) cannot block: two compares each
                    const bool tq = dr.x <= cr.y && cr.x <= dr.y, tr = dr.z <= cr.w && cr.z <= dr.w;
                    if (!tq && !tr) continue;
                    const uint8_t su = ((volatile uint8_t *)state)[u];
                    if (su == 2) continue;
                    const long long oq = (long long)(cr.y < dr.y ? cr.y : dr.y) - (long long)(cr.x > dr.x ? cr.x : dr.x) + 1;
                    const long long orr = (long long)(cr.w < dr.w ? cr.w : dr.w) - (long long)(cr.z > dr.z ? cr.z : dr.z) + 1;
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
                    ctl.unresolved = 1;
            }
        }
        __syncthreads();
        if (!ctl.unresolved) break;
    }
    // accumulate accepted chains: anchors / seeds (pooled), spans with the symmetric clipped extension
    const uint32_t rcoff = db.g_ctg_off[pi.r];
    const int nrc = (int)(db.g_ctg_off[pi.r + 1] - rcoff);
    int n_acc_local = 0;
    double l_span_q = 0, l_span_r = 0, l_a = 0, l_s = 0;  // exact in double (< 2^53); reduced below, no atomics
    for (int t = tid; t < nc; t += FIN_THREADS) {
        if (state[t] != 1) continue;
        n_acc_local++;
        const Cand &c = cands[(int)(skey[t] & FIN_IDX_MASK)];
        const long long e = prm.span_ext, k1 = K_SEED - 1;
        const long long cs = db.chunk_start[choff + c.chunk], ce = cs + db.chunk_len[choff + c.chunk] - 1;
        int lo = 0, hi = nrc - 1;  // reference contig holding r0
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (db.ctg_pstart[rcoff + mid] <= c.r0)
                lo = mid;
            else
                hi = mid - 1;
        }
        const long long rs = db.ctg_pstart[rcoff + lo], re = rs + db.ctg_len[rcoff + lo] - 1;
        const long long a0 = (long long)c.q0 - k1, a1 = c.q1, b0 = (long long)c.r0 - k1, b1 = c.r1;
        const long long room_ql = a0 - cs, room_qr = ce - a1, room_rlo = b0 - rs, room_rhi = re - b1;
        // an extension stops where either genome runs out; a reverse chain's left query end faces the reference's high end
        long long el = c.rev ? room_rhi : room_rlo, er = c.rev ? room_rlo : room_rhi;
        el = el < room_ql ? el : room_ql;
        er = er < room_qr ? er : room_qr;
        el = el < e ? el : e;
        er = er < e ? er : e;
        el = el < 0 ? 0 : el;
        er = er < 0 ? 0 : er;
        l_span_q += (double)(a1 - a0 + 1 + el + er);
        l_span_r += (double)(b1 - b0 + 1 + el + er);
        l_a += (double)c.n_anchors;
        l_s += (double)c.n_seeds;
    }
    const double t_span_q = block_sum(l_span_q, ctl.red, tid), t_span_r = block_sum(l_span_r, ctl.red, tid);
    const double t_a = block_sum(l_a, ctl.red, tid), t_s = block_sum(l_s, ctl.red, tid);
    const double t_acc = block_sum((double)n_acc_local, ctl.red, tid);
    if (tid == 0) {
        PairOut o;
        o.ani = o.ani_raw = -1.0;
        o.af_q = o.af_r = 0.0;
        o.n_anchors = (int64_t)t_a;
        o.n_seeds = (int64_t)t_s;
        o.span_q = (int64_t)t_span_q;
        o.span_r = (int64_t)t_span_r;
        o.n_chains = (int)t_acc;
        o.swapped = (int32_t)pi.swapped;
        // the two end anchors of every chain are anchors by construction: left out of both counts (oracle ora_pair)
        const int64_t a_in = o.n_anchors - 2 * (int64_t)o.n_chains, s_in = o.n_seeds - 2 * (int64_t)o.n_chains;
        if (a_in > 0 && s_in > 0) {
            double ratio = (double)a_in / (double)s_in;
            if (ratio > 1.0) ratio = 1.0;
            const double mean = pow(ratio, 1.0 / (double)K_SEED);
            o.ani_raw = mean;
            double afq = (double)o.span_q / (double)db.g_total_len[pi.q];
            double afr = (double)o.span_r / (double)db.g_total_len[pi.r];
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
        out[perm[p]] = o;
    }
}

// K5 -- edge compaction: keep pairs with an estimate and max(AF) >= min_af; percent units.  Order-preserving
// (flag -> exclusive scan -> scatter): the pair list is sorted by (a, b), so the edge list comes out sorted too.
__device__ __forceinline__ bool edge_kept(const PairOut &o, double min_af) {
    const double afa = o.swapped ? o.af_r : o.af_q, afb = o.swapped ? o.af_q : o.af_r;
    return o.ani >= 0.0 && (afa >= min_af || afb >= min_af);
}
__global__ void edge_flag_kernel(const PairOut *__restrict__ po, int64_t n_pairs, double min_af /* fraction */,
                                 uint32_t *__restrict__ flag) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_pairs) flag[t] = edge_kept(po[t], min_af) ? 1u : 0u;
}
__global__ void edge_scatter_kernel(const unsigned long long *__restrict__ pairs, const PairOut *__restrict__ po,
                                    int64_t n_pairs, const uint32_t *__restrict__ flag, const uint32_t *__restrict__ pos,
                                    skb_edge *__restrict__ edges, unsigned long long *n_edges) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pairs) return;
    if (t == n_pairs - 1) *n_edges = (unsigned long long)pos[t] + flag[t];
    if (!flag[t]) return;
    const PairOut &o = po[t];
    skb_edge e;
    e.a = (uint32_t)(pairs[t] >> 32);
    e.b = (uint32_t)(pairs[t] & 0xffffffffu);
    e.ani = o.ani * 100.0;
    e.af_a = (o.swapped ? o.af_r : o.af_q) * 100.0;
    e.af_b = (o.swapped ? o.af_q : o.af_r) * 100.0;
    edges[pos[t]] = e;
}

// bookkeeping: sum over pairs of the query genome's seed count and of the chained anchors (roofline bytes)
__global__ void pair_sums_kernel(const PairInfo *__restrict__ info, const PairOut *__restrict__ po, int64_t n_pairs,
                                 const uint64_t *__restrict__ g_seed_off, unsigned long long *sums /* [2] */) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long sq = 0, an = 0;
    if (t < n_pairs) {
        const uint32_t q = info[t].q;
        sq = g_seed_off[q + 1] - g_seed_off[q];
        an = (unsigned long long)po[t].n_anchors;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        sq += __shfl_down_sync(0xffffffffu, sq, d);
        an += __shfl_down_sync(0xffffffffu, an, d);
    }
    if ((threadIdx.x & 31) == 0 && (sq | an)) {
        atomicAdd(&sums[0], sq);
        atomicAdd(&sums[1], an);
    }
}

}  // namespace skb
