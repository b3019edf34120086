/*
 * oracle/skani_oracle.c -- TEST INFRASTRUCTURE ONLY (see skani_oracle.h for scope and parity status).
 *
 * Restates, on the CPU and in plain C, the path skDER runs through the external `skani` binary
 * (/root/reference/src/skDER/skder.py:16-18 triangle, :103/:119 sketch+search, :58-59 dist).
 * Each function names the step of the published skani method it follows; the reference repo itself
 * holds no arithmetic for this path.
 */
#include "skani_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <zlib.h>

#define SEED_KMER(s) ((uint32_t)((s) >> 34))
#define SEED_POS(s) ((uint32_t)(((s) >> 2) & 0xffffffffu))
#define SEED_REP(s) ((int)(((s) >> 1) & 1))
#define SEED_STRAND(s) ((int)((s) & 1))

struct ora_sketch {
    int32_t n_contigs;
    int64_t *contig_len;
    uint32_t *contig_off; /* padded genome coordinate of contig base 0 */
    int64_t total_len;
    char *first_name;
    int64_t n_seeds;
    uint64_t *seeds;  /* position-ordered */
    uint64_t *kidx;   /* same records sorted by (kmer, pos): the reference-side index */
    int64_t n_markers;
    uint64_t *markers; /* sorted unique */
    int32_t n_chunks;
    int64_t *chunk_seed_begin; /* n_chunks+1 */
    uint32_t *chunk_start;     /* padded genome coordinate of first base of chunk */
    uint32_t *chunk_len;
};

void ora_default_params(ora_params_t *p) {
    p->k = 15;
    p->marker_k = 21;
    p->c = 125;
    p->marker_c = 1000;
    p->min_contig_len = 500;
    p->contig_pad = 4096;
    p->chunk_len = 20000;
    p->band_bp = 2500;
    p->max_gap = 300;
    p->lookback = 16;
    p->anchor_score = 20;
    p->min_anchors = 3;
    p->min_score = 45;
    p->max_mult = 8;
    p->max_chunk_anchors = 256;
    p->max_chunk_chains = 8;
    p->ovl_num = 1;
    p->ovl_den = 2;
    p->span_ext = 170;
    p->role_rule = 0;
}

/* Invertible 64-bit mix used by skani/minimap2 for k-mer hashing (Appendix A of SURVEY.md).
 * The first line is ~(key + (key << 21)): Rust's unary `!` binds looser than the method call. */
uint64_t ora_mm_hash64(uint64_t key) {
    key = ~(key + (key << 21));
    key = key ^ (key >> 24);
    key = (key + (key << 3)) + (key << 8);
    key = key ^ (key >> 14);
    key = (key + (key << 2)) + (key << 4);
    key = key ^ (key >> 28);
    key = key + (key << 31);
    return key;
}

static const uint8_t *base_code_table(void) {
    static uint8_t t[256];
    static int init = 0;
    if (!init) {
        memset(t, 0, sizeof t); /* non-ACGT -> 0 ('A'), as skani's BYTE_TO_SEQ */
        t['C'] = t['c'] = 1;
        t['G'] = t['g'] = 2;
        t['T'] = t['t'] = 3;
        init = 1;
    }
    return t;
}

typedef struct {
    uint64_t *v;
    int64_t n, cap;
} u64vec;
static void push(u64vec *a, uint64_t x) {
    if (a->n == a->cap) {
        a->cap = a->cap ? a->cap * 2 : 4096;
        a->v = (uint64_t *)realloc(a->v, (size_t)a->cap * 8);
    }
    a->v[a->n++] = x;
}
static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

/* FracMinHash sketching of one contig (skani `fmh_seeds`): one pass of a rolling 21-mer window;
 * the seed 15-mer is the window's last 15 bases, so both k-mers end at base i. */
static void sketch_contig(const uint8_t *seq, int64_t len, uint32_t off, const ora_params_t *p, u64vec *seeds,
                          u64vec *markers) {
    const uint8_t *code = base_code_table();
    const int mk = p->marker_k, k = p->k;
    const uint64_t mmask = (~0ULL) >> (64 - 2 * mk);
    const uint64_t smask = (~0ULL) >> (64 - 2 * k);
    const uint64_t thr_seed = UINT64_MAX / p->c;
    const uint64_t thr_marker = UINT64_MAX / p->marker_c;
    uint64_t f = 0, r = 0;
    for (int64_t i = 0; i < len; i++) {
        uint64_t b = code[seq[i]];
        f = ((f << 2) | b) & mmask;
        r = (r >> 2) | ((3 - b) << (2 * (mk - 1)));
        if (i < mk - 1) continue;
        uint64_t fs = f & smask;
        uint64_t rs = r >> (2 * (mk - k));
        int fwd = fs < rs;
        uint64_t cs = fwd ? fs : rs;
        if (ora_mm_hash64(cs) < thr_seed)
            push(seeds, (cs << 34) | ((uint64_t)(off + (uint32_t)i) << 2) | (uint64_t)fwd);
        uint64_t cm = f < r ? f : r;
        if (ora_mm_hash64(cm) < thr_marker) push(markers, cm);
    }
}

static void finish_sketch(ora_sketch_t *s, u64vec *seeds, u64vec *markers, const ora_params_t *p) {
    /* markers: sorted set */
    qsort(markers->v, (size_t)markers->n, 8, cmp_u64);
    int64_t m = 0;
    for (int64_t i = 0; i < markers->n; i++)
        if (i == 0 || markers->v[i] != markers->v[i - 1]) markers->v[m++] = markers->v[i];
    s->markers = markers->v;
    s->n_markers = m;
    /* reference-side index: records sorted by (kmer, pos); flag k-mers repeated in this genome */
    s->n_seeds = seeds->n;
    s->seeds = seeds->v;
    s->kidx = (uint64_t *)malloc((size_t)(seeds->n ? seeds->n : 1) * 8);
    memcpy(s->kidx, s->seeds, (size_t)seeds->n * 8);
    qsort(s->kidx, (size_t)seeds->n, 8, cmp_u64);
    /* own multiplicity -> rep flag on both copies */
    for (int64_t i = 0; i < s->n_seeds;) {
        int64_t j = i;
        while (j < s->n_seeds && SEED_KMER(s->kidx[j]) == SEED_KMER(s->kidx[i])) j++;
        if (j - i > p->max_mult)
            for (int64_t t = i; t < j; t++) s->kidx[t] |= 2;
        i = j;
    }
    /* propagate rep flag to the position-ordered copy (binary search by full record) */
    for (int64_t i = 0; i < s->n_seeds; i++) {
        uint64_t key = s->seeds[i];
        int64_t lo = 0, hi = s->n_seeds;
        while (lo < hi) {
            int64_t mid = (lo + hi) >> 1;
            if ((s->kidx[mid] & ~2ULL) < key)
                lo = mid + 1;
            else
                hi = mid;
        }
        if (lo < s->n_seeds && (s->kidx[lo] & 2)) s->seeds[i] |= 2;
    }
    /* chunk table: every contig is cut into chunk_len windows */
    int32_t nch = 0;
    for (int32_t c = 0; c < s->n_contigs; c++) nch += (int32_t)((s->contig_len[c] + p->chunk_len - 1) / p->chunk_len);
    s->n_chunks = nch;
    s->chunk_seed_begin = (int64_t *)malloc((size_t)(nch + 1) * 8);
    s->chunk_start = (uint32_t *)malloc((size_t)(nch ? nch : 1) * 4);
    s->chunk_len = (uint32_t *)malloc((size_t)(nch ? nch : 1) * 4);
    int32_t ch = 0;
    int64_t si = 0;
    for (int32_t c = 0; c < s->n_contigs; c++) {
        for (int64_t st = 0; st < s->contig_len[c]; st += p->chunk_len, ch++) {
            int64_t ln = s->contig_len[c] - st < p->chunk_len ? s->contig_len[c] - st : p->chunk_len;
            s->chunk_start[ch] = s->contig_off[c] + (uint32_t)st;
            s->chunk_len[ch] = (uint32_t)ln;
            s->chunk_seed_begin[ch] = si;
            uint32_t end = s->chunk_start[ch] + (uint32_t)ln;
            while (si < s->n_seeds && SEED_POS(s->seeds[si]) < end) si++;
        }
    }
    s->chunk_seed_begin[nch] = si;
}

ora_sketch_t *ora_sketch_contigs(const char *const *seqs, const int64_t *lens, int n, const ora_params_t *p) {
    ora_sketch_t *s = (ora_sketch_t *)calloc(1, sizeof *s);
    s->contig_len = (int64_t *)malloc((size_t)(n ? n : 1) * 8);
    s->contig_off = (uint32_t *)malloc((size_t)(n ? n : 1) * 4);
    u64vec seeds = {0}, markers = {0};
    uint32_t off = 0;
    for (int i = 0; i < n; i++) {
        if (lens[i] < p->min_contig_len || lens[i] <= 0) continue;
        int c = s->n_contigs++;
        s->contig_len[c] = lens[i];
        s->contig_off[c] = off;
        s->total_len += lens[i];
        sketch_contig((const uint8_t *)seqs[i], lens[i], off, p, &seeds, &markers);
        off += (uint32_t)lens[i] + (uint32_t)p->contig_pad;
    }
    s->first_name = strdup("");
    finish_sketch(s, &seeds, &markers, p);
    return s;
}

/* FASTA ingest (plain or gzip; zlib's gzread handles both). Contigs below min_contig_len are
 * dropped, as skani's file reader does; Ref_name/Query_name is the first kept record's header. */
ora_sketch_t *ora_sketch_file(const char *path, const ora_params_t *p) {
    gzFile g = gzopen(path, "rb");
    if (!g) return NULL;
    size_t cap = 1 << 24, n = 0;
    char *buf = (char *)malloc(cap);
    for (;;) {
        if (n + (1 << 20) > cap) {
            cap *= 2;
            buf = (char *)realloc(buf, cap);
        }
        int got = gzread(g, buf + n, 1 << 20);
        if (got <= 0) break;
        n += (size_t)got;
    }
    gzclose(g);
    /* split records */
    int ncap = 1024, nrec = 0;
    char **seqs = (char **)malloc(sizeof(char *) * (size_t)ncap);
    int64_t *lens = (int64_t *)malloc(8 * (size_t)ncap);
    char **names = (char **)malloc(sizeof(char *) * (size_t)ncap);
    size_t i = 0;
    char *w = buf; /* compact sequence in place */
    while (i < n) {
        if (buf[i] == '>') {
            size_t j = i + 1;
            while (j < n && buf[j] != '\n') j++;
            size_t e = j;
            while (e > i + 1 && (buf[e - 1] == '\r')) e--;
            if (nrec == ncap) {
                ncap *= 2;
                seqs = (char **)realloc(seqs, sizeof(char *) * (size_t)ncap);
                lens = (int64_t *)realloc(lens, 8 * (size_t)ncap);
                names = (char **)realloc(names, sizeof(char *) * (size_t)ncap);
            }
            names[nrec] = strndup(buf + i + 1, e - (i + 1));
            seqs[nrec] = w;
            lens[nrec] = 0;
            nrec++;
            i = j + 1;
        } else {
            char ch = buf[i++];
            if (ch == '\n' || ch == '\r' || ch == ' ' || ch == '\t') continue;
            if (nrec == 0) continue; /* junk before first header */
            *w++ = ch;
            lens[nrec - 1]++;
        }
    }
    ora_sketch_t *s = ora_sketch_contigs((const char *const *)seqs, lens, nrec, p);
    free(s->first_name);
    s->first_name = NULL;
    for (int r = 0; r < nrec; r++) {
        if (!s->first_name && lens[r] >= p->min_contig_len && lens[r] > 0) s->first_name = strdup(names[r]);
        free(names[r]);
    }
    if (!s->first_name) s->first_name = strdup("");
    free(names);
    free(seqs);
    free(lens);
    free(buf);
    return s;
}

void ora_sketch_free(ora_sketch_t *s) {
    if (!s) return;
    free(s->contig_len);
    free(s->contig_off);
    free(s->first_name);
    free(s->seeds);
    free(s->kidx);
    free(s->markers);
    free(s->chunk_seed_begin);
    free(s->chunk_start);
    free(s->chunk_len);
    free(s);
}

int64_t ora_n_seeds(const ora_sketch_t *s) { return s->n_seeds; }
int64_t ora_n_markers(const ora_sketch_t *s) { return s->n_markers; }
int64_t ora_total_len(const ora_sketch_t *s) { return s->total_len; }
int32_t ora_n_contigs(const ora_sketch_t *s) { return s->n_contigs; }
int32_t ora_n_chunks(const ora_sketch_t *s) { return s->n_chunks; }
const uint64_t *ora_seeds(const ora_sketch_t *s) { return s->seeds; }
const uint64_t *ora_markers(const ora_sketch_t *s) { return s->markers; }
const int64_t *ora_contig_lens(const ora_sketch_t *s) { return s->contig_len; }
const char *ora_first_name(const ora_sketch_t *s) { return s->first_name; }

/* Marker prescreen (skani `-s`): shared markers must exceed screen^21 * min(|Ma|,|Mb|). */
int64_t ora_screen(const ora_sketch_t *a, const ora_sketch_t *b, double screen, const ora_params_t *p, int *pass) {
    int64_t i = 0, j = 0, shared = 0;
    while (i < a->n_markers && j < b->n_markers) {
        if (a->markers[i] < b->markers[j])
            i++;
        else if (a->markers[i] > b->markers[j])
            j++;
        else {
            shared++;
            i++;
            j++;
        }
    }
    int64_t mn = a->n_markers < b->n_markers ? a->n_markers : b->n_markers;
    double cutoff = pow(screen, (double)p->marker_k) * (double)mn;
    if (pass) *pass = (screen <= 0.0) ? 1 : ((double)shared > cutoff);
    return shared;
}

typedef struct {
    uint32_t r, q;
    int32_t rev;
} anchor_t;
static int cmp_anchor(const void *a, const void *b) {
    const anchor_t *x = (const anchor_t *)a, *y = (const anchor_t *)b;
    if (x->q != y->q) return x->q < y->q ? -1 : 1;
    if (x->r != y->r) return x->r < y->r ? -1 : 1;
    return 0;
}
typedef struct {
    ora_chain_t c;
} cand_t;
static int cmp_cand(const void *a, const void *b) {
    const ora_chain_t *x = (const ora_chain_t *)a, *y = (const ora_chain_t *)b;
    if (x->score != y->score) return x->score > y->score ? -1 : 1;
    if (x->chunk != y->chunk) return x->chunk < y->chunk ? -1 : 1;
    if (x->q0 != y->q0) return x->q0 < y->q0 ? -1 : 1;
    if (x->r0 != y->r0) return x->r0 < y->r0 ? -1 : 1;
    return 0;
}

static int64_t kidx_lower(const ora_sketch_t *s, uint32_t kmer) {
    int64_t lo = 0, hi = s->n_seeds;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (SEED_KMER(s->kidx[mid]) < kmer)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

/* anchors of one query chunk against the reference index, honouring the multiplicity cap */
static int chunk_anchors(const ora_sketch_t *q, const ora_sketch_t *r, int32_t ch, int mult, anchor_t *out, int cap,
                         int *overflow) {
    int n = 0;
    *overflow = 0;
    for (int64_t i = q->chunk_seed_begin[ch]; i < q->chunk_seed_begin[ch + 1]; i++) {
        uint64_t s = q->seeds[i];
        if (SEED_REP(s)) continue;
        uint32_t km = SEED_KMER(s);
        int64_t lo = kidx_lower(r, km), hi = lo;
        while (hi < r->n_seeds && SEED_KMER(r->kidx[hi]) == km) hi++;
        if (hi == lo || hi - lo > mult) continue;
        for (int64_t t = lo; t < hi; t++) {
            if (n == cap) {
                *overflow = 1;
                return n;
            }
            out[n].q = SEED_POS(s);
            out[n].r = SEED_POS(r->kidx[t]);
            out[n].rev = SEED_STRAND(s) != SEED_STRAND(r->kidx[t]);
            n++;
        }
    }
    return n;
}

static int64_t seeds_in_span(const ora_sketch_t *q, int32_t ch, uint32_t q0, uint32_t q1) {
    int64_t n = 0;
    for (int64_t i = q->chunk_seed_begin[ch]; i < q->chunk_seed_begin[ch + 1]; i++) {
        uint32_t pz = SEED_POS(q->seeds[i]);
        n += (pz >= q0 && pz <= q1);
    }
    return n;
}

/* Substitute for skani's learned ANI debiasing (a gradient-boosted model embedded in the skani
 * binary, on by default for c >= 70; its weights are not available here).  Divergence is mapped
 * by a power law fitted ONCE, openly, to the 561 golden pairs of
 * test_case/skder_gtdb_results/Skani_Triangle_Edge_Output.txt (fit script: oracle/fit_debias.py):
 *   100-ANI_reported = DEBIAS_A * (100-ANI_raw)^DEBIAS_G      [percent units] */
#define DEBIAS_A 1.438713
#define DEBIAS_G 0.905292
double ora_debias(double ani_raw) {
    double x = 100.0 * (1.0 - ani_raw);
    if (x <= 0.0) return 1.0;
    double y = DEBIAS_A * pow(x, DEBIAS_G);
    double a = 1.0 - y / 100.0;
    return a < 0.0 ? 0.0 : a;
}

int ora_pair(const ora_sketch_t *a, const ora_sketch_t *b, const ora_params_t *p, ora_pair_result_t *out,
             ora_chain_t *chains_out, int max_chains, int *n_chains_out) {
    memset(out, 0, sizeof *out);
    out->ani = out->ani_raw = -1.0;
    if (n_chains_out) *n_chains_out = 0;
    /* role assignment: results must not depend on argument order */
    int swap;
    if (a->n_seeds != b->n_seeds)
        swap = p->role_rule == 0 ? (b->n_seeds < a->n_seeds) : (b->n_seeds > a->n_seeds);
    else
        swap = 0;
    const ora_sketch_t *q = swap ? b : a, *r = swap ? a : b;
    out->swapped = swap;

    const int MAXA = p->max_chunk_anchors;
    anchor_t *an = (anchor_t *)malloc(sizeof(anchor_t) * (size_t)MAXA);
    int32_t *f = (int32_t *)malloc(4 * (size_t)MAXA);
    int32_t *root = (int32_t *)malloc(4 * (size_t)MAXA);
    int32_t *cnt = (int32_t *)malloc(4 * (size_t)MAXA);
    int32_t *best_end = (int32_t *)malloc(4 * (size_t)MAXA);
    size_t ccap = 1024, nc = 0;
    ora_chain_t *cands = (ora_chain_t *)malloc(sizeof(ora_chain_t) * ccap);

    const int chunk_cap = p->max_chunk_chains;
    nc = 0;
    for (int32_t ch = 0; ch < q->n_chunks; ch++) {
        /* anchors; if the chunk overflows, halve the multiplicity cap until it fits */
        int mult = p->max_mult, ovf = 0, n = 0;
        for (;;) {
            n = chunk_anchors(q, r, ch, mult, an, MAXA, &ovf);
            if (!ovf || mult == 0) break;
            mult >>= 1;
        }
        if (ovf) n = 0;
        if (n < p->min_anchors) continue;
        qsort(an, (size_t)n, sizeof(anchor_t), cmp_anchor);
        /* banded chaining DP in QUERY order (anchors sorted by (query pos, ref pos)); integer scores.
         * Predecessor j of i: same strand relation, 0 < dq <= band_bp, ref moves the matching way
         * (dr > 0), |dq - dr| <= max_gap, at most `lookback` anchors back; ties: nearest j. */
        for (int i = 0; i < n; i++) {
            int32_t best = p->anchor_score, bj = -1;
            for (int j = i - 1; j >= 0 && j >= i - p->lookback; j--) {
                uint32_t dq = an[i].q - an[j].q;
                if (dq > (uint32_t)p->band_bp) break;
                if (an[j].rev != an[i].rev || dq == 0) continue;
                int64_t dr = an[i].rev ? (int64_t)an[j].r - (int64_t)an[i].r : (int64_t)an[i].r - (int64_t)an[j].r;
                if (dr <= 0) continue;
                int64_t gap = dr - (int64_t)dq;
                if (gap < 0) gap = -gap;
                if (gap > p->max_gap) continue;
                int32_t cand = f[j] + p->anchor_score - (int32_t)gap;
                if (cand > best) {
                    best = cand;
                    bj = j;
                }
            }
            f[i] = best;
            if (bj < 0) {
                root[i] = i;
                cnt[i] = 1;
            } else {
                root[i] = root[bj];
                cnt[i] = cnt[bj] + 1;
            }
        }
        /* one chain per DP tree: its best-scoring end (ties: lowest index) */
        for (int i = 0; i < n; i++) best_end[i] = -1;
        for (int i = 0; i < n; i++) {
            int rt = root[i];
            if (best_end[rt] < 0 || f[i] > f[best_end[rt]]) best_end[rt] = i;
        }
        size_t first = nc;
        for (int rt = 0; rt < n; rt++) {
            int e = best_end[rt];
            if (e < 0) continue;
            if (cnt[e] < p->min_anchors || f[e] < p->min_score) continue;
            ora_chain_t c;
            c.chunk = ch;
            c.n_anchors = cnt[e];
            c.score = f[e];
            c.rev = an[e].rev;
            c.q0 = an[rt].q;
            c.q1 = an[e].q;
            c.r0 = an[rt].r < an[e].r ? an[rt].r : an[e].r;
            c.r1 = an[rt].r < an[e].r ? an[e].r : an[rt].r;
            c.n_seeds = (int32_t)seeds_in_span(q, ch, c.q0, c.q1);
            if (nc == ccap) {
                ccap *= 2;
                cands = (ora_chain_t *)realloc(cands, sizeof(ora_chain_t) * ccap);
            }
            cands[nc++] = c;
        }
        /* keep the max_chunk_chains best candidates of this chunk */
        if (nc - first > (size_t)chunk_cap) {
            qsort(cands + first, nc - first, sizeof(ora_chain_t), cmp_cand);
            nc = first + (size_t)chunk_cap;
        }
    }
    /* pair-level selection: best score first; reject a chain overlapping an accepted one by more
     * than ovl_num/ovl_den of its own length on the reference or on the query */
    qsort(cands, nc, sizeof(ora_chain_t), cmp_cand);
    size_t na = 0;
    for (size_t i = 0; i < nc; i++) {
        ora_chain_t *c = &cands[i];
        int64_t lq = (int64_t)c->q1 - c->q0 + 1, lr = (int64_t)c->r1 - c->r0 + 1;
        int ok = 1;
        for (size_t j = 0; j < na && ok; j++) {
            ora_chain_t *d = &cands[j];
            int64_t oq = (int64_t)(c->q1 < d->q1 ? c->q1 : d->q1) - (int64_t)(c->q0 > d->q0 ? c->q0 : d->q0) + 1;
            int64_t orr = (int64_t)(c->r1 < d->r1 ? c->r1 : d->r1) - (int64_t)(c->r0 > d->r0 ? c->r0 : d->r0) + 1;
            if (oq > 0 && oq * p->ovl_den > lq * p->ovl_num) ok = 0;
            if (orr > 0 && orr * p->ovl_den > lr * p->ovl_num) ok = 0;
        }
        if (ok) cands[na++] = *c;
    }
    /* ANI: anchors over query seeds inside the accepted chains, pooled over the pair.  The two END anchors of a
     * chain are anchors by construction (they define the span the seeds are counted in), so they are left out of
     * both counts -- otherwise short chains (fragmented assemblies) bias the ratio upward.  ANI_raw = ratio^(1/k).
     * AF: covered bases = the k-mer span first..last anchor, extended by span_ext on both sides (the homology
     * boundary lies about one anchor spacing beyond the outermost anchors).  An extension is cut where EITHER
     * genome runs out -- the query chunk or the reference contig, at the matching end (for a reverse chain the
     * query's left end faces the reference's high end) -- and applies to both spans alike. */
    int64_t span_q = 0, span_r = 0, At = 0, St = 0;
    for (size_t i = 0; i < na; i++) {
        const ora_chain_t *c = &cands[i];
        const int64_t k1 = p->k - 1, e = p->span_ext;
        const int64_t cs = q->chunk_start[c->chunk], ce = cs + q->chunk_len[c->chunk] - 1;
        int32_t lo = 0, hi = r->n_contigs - 1; /* contig holding r0 */
        while (lo < hi) {
            int32_t mid = (lo + hi + 1) >> 1;
            if (r->contig_off[mid] <= c->r0)
                lo = mid;
            else
                hi = mid - 1;
        }
        const int64_t rs = r->contig_off[lo], re = rs + r->contig_len[lo] - 1;
        const int64_t a0 = (int64_t)c->q0 - k1, a1 = c->q1, b0 = (int64_t)c->r0 - k1, b1 = c->r1;
        const int64_t room_ql = a0 - cs, room_qr = ce - a1, room_rlo = b0 - rs, room_rhi = re - b1;
        int64_t el = c->rev ? room_rhi : room_rlo, er = c->rev ? room_rlo : room_rhi;
        if (room_ql < el) el = room_ql;
        if (room_qr < er) er = room_qr;
        if (e < el) el = e;
        if (e < er) er = e;
        if (el < 0) el = 0;
        if (er < 0) er = 0;
        span_q += a1 - a0 + 1 + el + er;
        span_r += b1 - b0 + 1 + el + er;
        At += c->n_anchors;
        St += c->n_seeds;
    }
    out->n_chains = (int32_t)na;
    out->n_anchors_total = At;
    out->n_seeds_total = St;
    out->span_q = span_q;
    out->span_r = span_r;
    if (St - 2 * (int64_t)na > 0 && At - 2 * (int64_t)na > 0) {
        double ratio = (double)(At - 2 * (int64_t)na) / (double)(St - 2 * (int64_t)na);
        if (ratio > 1.0) ratio = 1.0;
        const double mean = pow(ratio, 1.0 / (double)p->k);
        out->ani_raw = mean;
        double afq = (double)span_q / (double)q->total_len, afr = (double)span_r / (double)r->total_len;
        if (afq > 1.0) afq = 1.0;
        if (afr > 1.0) afr = 1.0;
        out->ani = ora_debias(mean);
        if (out->ani > 1.0) out->ani = 1.0;
        out->af_a = swap ? afr : afq;
        out->af_b = swap ? afq : afr;
    }
    if (chains_out) {
        int m = (int)na < max_chains ? (int)na : max_chains;
        memcpy(chains_out, cands, sizeof(ora_chain_t) * (size_t)m);
        if (n_chains_out) *n_chains_out = m;
    }
    free(an);
    free(f);
    free(root);
    free(cnt);
    free(best_end);
    free(cands);
    return out->ani_raw >= 0 ? 0 : 1;
}

/* ---------------------------------------------------------------------------------------------
 * All-vs-all on host threads: bench.py's CPU arm.  ANI/AF is ora_pair above, driven from C so that
 * the measurement holds no Python per-pair overhead.  The prescreen is done the way SURVEY.md
 * section 8 row a4 says skani does it on the CPU -- an inverted marker -> genomes index, so its cost
 * follows the SHARED markers, not pairs x sketch size: (marker << 22 | genome) keys are radix-sorted,
 * every run of equal markers adds 1 to the count of each genome pair in it, and the counts are
 * thresholded with ora_screen's rule.  The decisions are identical to calling ora_screen on every
 * pair (tests/test_oracle_golden.py checks that).  Work is handed out by an atomic counter.
 * ------------------------------------------------------------------------------------------- */
#include <pthread.h>
#include <time.h>

typedef struct {
    const ora_sketch_t *const *sk;
    int n;
    double screen, min_af;
    const ora_params_t *p;
    const uint64_t *keys;   /* sorted (marker << 22 | genome) */
    int64_t n_keys;
    uint32_t *cnt;          /* n*n shared-marker counts, [a*n+b] for a<b */
    uint8_t *pass;          /* n*n */
    const uint32_t *surv;   /* 2 ints per surviving pair */
    int64_t n_surv;
    int64_t next;           /* atomic work counter */
    int64_t n_edges;        /* atomic */
    int phase;
} tri_job_t;

#define TRI_GID_BITS 22
#define TRI_BLOCK 4096

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* LSD radix sort of 64-bit keys, 8 bits per pass, passes over constant bytes skipped */
static void radix_sort_u64(uint64_t *a, uint64_t *tmp, int64_t n) {
    for (int shift = 0; shift < 64; shift += 8) {
        int64_t h[257];
        memset(h, 0, sizeof(h));
        for (int64_t i = 0; i < n; i++) h[((a[i] >> shift) & 255) + 1]++;
        int constant = 0;
        for (int b = 0; b < 256; b++)
            if (h[b + 1] == n) constant = 1;
        if (constant) continue;
        for (int b = 0; b < 256; b++) h[b + 1] += h[b];
        for (int64_t i = 0; i < n; i++) tmp[h[(a[i] >> shift) & 255]++] = a[i];
        memcpy(a, tmp, sizeof(uint64_t) * (size_t)n);
    }
}

static void *tri_worker(void *arg) {
    tri_job_t *j = (tri_job_t *)arg;
    const uint64_t gmask = ((uint64_t)1 << TRI_GID_BITS) - 1;
    if (j->phase == 0) { /* runs of equal markers -> pair counts; a run belongs to the block it starts in */
        for (;;) {
            const int64_t b0 = __atomic_fetch_add(&j->next, TRI_BLOCK, __ATOMIC_RELAXED);
            if (b0 >= j->n_keys) break;
            int64_t i = b0;
            const int64_t b1 = b0 + TRI_BLOCK < j->n_keys ? b0 + TRI_BLOCK : j->n_keys;
            while (i < b1 && i > 0 && (j->keys[i] >> TRI_GID_BITS) == (j->keys[i - 1] >> TRI_GID_BITS)) i++;
            while (i < b1) {
                int64_t e = i + 1;
                while (e < j->n_keys && (j->keys[e] >> TRI_GID_BITS) == (j->keys[i] >> TRI_GID_BITS)) e++;
                for (int64_t x = i; x < e; x++)
                    for (int64_t y = x + 1; y < e; y++)
                        __atomic_fetch_add(&j->cnt[(size_t)(j->keys[x] & gmask) * j->n + (size_t)(j->keys[y] & gmask)], 1u,
                                           __ATOMIC_RELAXED);
                i = e;
            }
        }
    } else if (j->phase == 1) { /* threshold the counts, ora_screen's rule */
        const double s21 = pow(j->screen, (double)j->p->marker_k);
        for (;;) {
            const int64_t a = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
            if (a >= j->n) break;
            for (int b = (int)a + 1; b < j->n; b++) {
                const int64_t ma = j->sk[a]->n_markers, mb = j->sk[b]->n_markers;
                const double cutoff = s21 * (double)(ma < mb ? ma : mb);
                j->pass[(size_t)a * j->n + b] = (j->screen <= 0.0) ? 1 : ((double)j->cnt[(size_t)a * j->n + b] > cutoff);
            }
        }
    } else {
        int64_t edges = 0;
        for (;;) {
            const int64_t i = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
            if (i >= j->n_surv) break;
            ora_pair_result_t r;
            ora_pair(j->sk[j->surv[2 * i]], j->sk[j->surv[2 * i + 1]], j->p, &r, NULL, 0, NULL);
            if (r.ani >= 0 && (r.af_a >= j->min_af || r.af_b >= j->min_af)) edges++;
        }
        __atomic_fetch_add(&j->n_edges, edges, __ATOMIC_RELAXED);
    }
    return NULL;
}

static void tri_run(tri_job_t *j, int threads) {
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    j->next = 0;
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, tri_worker, j);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
}

int64_t ora_triangle(const ora_sketch_t *const *sk, int n, double screen, double min_af, const ora_params_t *p,
                     int threads, int64_t *n_edges, uint8_t *pass_out, double *t_index, double *t_count, double *t_ani) {
    tri_job_t j;
    memset(&j, 0, sizeof(j));
    if (threads < 1) threads = 1;
    if (n >= (1 << TRI_GID_BITS)) return -1;
    j.sk = sk;
    j.n = n;
    j.screen = screen;
    j.min_af = min_af;
    j.p = p;
    double t0 = now_s();
    int64_t nk = 0;
    for (int g = 0; g < n; g++) nk += sk[g]->n_markers;
    uint64_t *keys = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(nk + 1));
    uint64_t *tmp = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(nk + 1));
    nk = 0;
    for (int g = 0; g < n; g++)
        for (int64_t i = 0; i < sk[g]->n_markers; i++) keys[nk++] = (sk[g]->markers[i] << TRI_GID_BITS) | (uint64_t)g;
    radix_sort_u64(keys, tmp, nk);
    free(tmp);
    double t1 = now_s();
    j.keys = keys;
    j.n_keys = nk;
    j.cnt = (uint32_t *)calloc((size_t)n * (size_t)n + 1, sizeof(uint32_t));
    j.pass = (uint8_t *)calloc((size_t)n * (size_t)n + 1, 1);
    j.phase = 0;
    tri_run(&j, threads);
    j.phase = 1;
    tri_run(&j, threads);
    int64_t ns = 0;
    for (int a = 0; a < n; a++)
        for (int b = a + 1; b < n; b++) ns += j.pass[(size_t)a * n + b];
    uint32_t *surv = (uint32_t *)malloc(sizeof(uint32_t) * 2 * (size_t)(ns + 1));
    int64_t k = 0;
    for (int a = 0; a < n; a++)
        for (int b = a + 1; b < n; b++)
            if (j.pass[(size_t)a * n + b]) {
                surv[2 * k] = (uint32_t)a;
                surv[2 * k + 1] = (uint32_t)b;
                k++;
            }
    double t2 = now_s();
    j.surv = surv;
    j.n_surv = ns;
    j.phase = 2;
    tri_run(&j, threads);
    double t3 = now_s();
    if (n_edges) *n_edges = j.n_edges;
    if (pass_out) memcpy(pass_out, j.pass, (size_t)n * (size_t)n);
    if (t_index) *t_index = t1 - t0;
    if (t_count) *t_count = t2 - t1;
    if (t_ani) *t_ani = t3 - t2;
    free(surv);
    free(j.pass);
    free(j.cnt);
    free(keys);
    return ns;
}

/* ---------------------------------------------------------------------------------------------
 * One query against a database on host threads: what `skani search q -d db` computes for one call of the
 * low_mem_greedy loop (reference src/skDER/skder.py:116-133).  Every database genome is screened against the query
 * (ora_screen) and, if it passes, compared (ora_pair); rows with an estimate and max(AF) >= min_af are kept.
 * out_ref / out_ani / out_af_ref / out_af_query need room for n entries.  Returns the number of rows.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    const ora_sketch_t *const *sk;
    int n, q;
    double screen, min_af;
    const ora_params_t *p;
    int64_t next;
    int32_t *out_ref;
    double *out_ani, *out_af_ref, *out_af_query;
    int64_t n_out;
} search_job_t;

static void *search_worker(void *arg) {
    search_job_t *j = (search_job_t *)arg;
    for (;;) {
        const int64_t r = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
        if (r >= j->n) break;
        int pass = 0;
        ora_screen(j->sk[r], j->sk[j->q], j->screen, j->p, &pass);
        if (!pass) continue;
        ora_pair_result_t res;
        ora_pair(j->sk[r], j->sk[j->q], j->p, &res, NULL, 0, NULL);
        if (res.ani < 0 || (res.af_a < j->min_af && res.af_b < j->min_af)) continue;
        const int64_t k = __atomic_fetch_add(&j->n_out, 1, __ATOMIC_RELAXED);
        j->out_ref[k] = (int32_t)r;
        j->out_ani[k] = res.ani;
        j->out_af_ref[k] = res.af_a;
        j->out_af_query[k] = res.af_b;
    }
    return NULL;
}

int64_t ora_search(const ora_sketch_t *const *sk, int n, int q, double screen, double min_af, const ora_params_t *p,
                   int threads, int32_t *out_ref, double *out_ani, double *out_af_ref, double *out_af_query) {
    search_job_t j;
    memset(&j, 0, sizeof(j));
    j.sk = sk;
    j.n = n;
    j.q = q;
    j.screen = screen;
    j.min_af = min_af;
    j.p = p;
    j.out_ref = out_ref;
    j.out_ani = out_ani;
    j.out_af_ref = out_af_ref;
    j.out_af_query = out_af_query;
    if (threads < 1) threads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, search_worker, &j);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
    return j.n_out;
}
