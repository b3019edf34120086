/*
 * oracle/skani_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C99) of the all-vs-all ANI/AF path that skDER reaches through
 * `skani triangle` / `skani search` / `skani dist`
 *   (/root/reference/src/skDER/skder.py:16-18, :58-59, :103, :119; cidder.py:362-363).
 *
 * The arithmetic of that path lives in the third-party Rust program `skani`
 * (bluenote-1577/skani; version NOT pinned by the reference: skDER_env.yml:12,
 * bioconda_recipe/meta.yaml:28) whose source is absent from /root/reference. This file therefore
 * restates the PUBLISHED method (Shaw & Yu, Nat. Methods 2023, cited at reference README.md:436):
 * FracMinHash seeds (k=15, c=125) and markers (k=21, c=1000), marker-containment prescreen,
 * seed anchoring, banded chaining per 20 kb query chunk, ANI = (anchors/seeds inside the chains)^(1/k),
 * AF = chained span / genome length. Details skani does not publish (and its learned ANI
 * debiasing model, whose weights are unavailable) are stated in DESIGN.md section 3.
 *
 * PARITY STATUS: pinned against the reference's golden triangle/dist outputs only
 * (test_case/skder_results, test_case/skder_gtdb_results, test_case/cidder_results) at the
 * residuals recorded in tests/golden/ORACLE_VS_GOLDEN.md.  k-mer hashes, sketches and prescreen
 * decisions: PARITY UNPINNED vs skani (no fixture exposes them).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may use this.
 */
#ifndef SKANI_ORACLE_H
#define SKANI_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int32_t k;              /* seed k-mer length (15) */
    int32_t marker_k;       /* marker k-mer length (21) */
    uint64_t c;             /* seed compression: keep hash < 2^64/c (125) */
    uint64_t marker_c;      /* marker compression (1000) */
    int32_t min_contig_len; /* contigs shorter than this are ignored (500) */
    int32_t contig_pad;     /* virtual gap between contigs in genome coordinates (4096) */
    int32_t chunk_len;      /* query chunk length (20000) */
    int32_t band_bp;        /* chaining look-back on the query, bp (2500) */
    int32_t max_gap;        /* max |d_ref - d_query| between chained anchors (300) */
    int32_t lookback;       /* chaining look-back in anchors (16) */
    int32_t anchor_score;   /* score per chained anchor (20) */
    int32_t min_anchors;    /* anchors a chain needs (3) */
    int32_t min_score;      /* score a chain needs (45) */
    int32_t max_mult;       /* seeds whose k-mer occurs more often in either genome are skipped */
    int32_t max_chunk_anchors; /* anchors kept per chunk (256) */
    int32_t max_chunk_chains;  /* chain candidates kept per chunk (8) */
    int32_t ovl_num;        /* chain rejected if overlap*ovl_den > ovl_num*own_length ... */
    int32_t ovl_den;        /* ... with an accepted chain on the reference or the query */
    int32_t span_ext;       /* bases each accepted chain is extended by on both sides, clipped (170; scanned, see fit_debias.py) */
    int32_t role_rule;      /* 0: query = fewer seeds (ties: lower index) ; 1: query = more seeds */
} ora_params_t;

/* one accepted chain, for diagnostics and calibration */
typedef struct {
    int32_t chunk;      /* query chunk id */
    int32_t n_anchors;
    int32_t n_seeds;    /* query seeds inside [q0,q1] */
    int32_t score;
    uint32_t q0, q1;    /* query span, genome coordinates (padded) */
    uint32_t r0, r1;    /* reference span */
    int32_t rev;
} ora_chain_t;

typedef struct {
    double ani;         /* final ANI in [0,1] (after debias), <0 if no estimate */
    double ani_raw;     /* before debias */
    double af_a;        /* aligned fraction of genome a in [0,1] */
    double af_b;
    int32_t n_chains;
    int32_t swapped;    /* 1 if b was the query */
    int64_t n_anchors_total;
    int64_t n_seeds_total;
    int64_t span_q, span_r;
} ora_pair_result_t;

typedef struct ora_sketch ora_sketch_t;

void ora_default_params(ora_params_t *p);
uint64_t ora_mm_hash64(uint64_t key);

/* FASTA (plain or .gz) -> sketch.  Returns NULL on failure. */
ora_sketch_t *ora_sketch_file(const char *path, const ora_params_t *p);
/* from an in-memory list of contigs (ASCII bases) */
ora_sketch_t *ora_sketch_contigs(const char *const *seqs, const int64_t *lens, int n, const ora_params_t *p);
void ora_sketch_free(ora_sketch_t *s);

int64_t ora_n_seeds(const ora_sketch_t *s);
int64_t ora_n_markers(const ora_sketch_t *s);
int64_t ora_total_len(const ora_sketch_t *s);
int32_t ora_n_contigs(const ora_sketch_t *s);
int32_t ora_n_chunks(const ora_sketch_t *s);
const uint64_t *ora_seeds(const ora_sketch_t *s);     /* position-ordered packed seeds */
const uint64_t *ora_markers(const ora_sketch_t *s);   /* sorted unique canonical 21-mers */
const int64_t *ora_contig_lens(const ora_sketch_t *s);
const char *ora_first_name(const ora_sketch_t *s);

/* prescreen: returns shared marker count; *pass = shared > screen^marker_k * min(|Ma|,|Mb|) */
int64_t ora_screen(const ora_sketch_t *a, const ora_sketch_t *b, double screen, const ora_params_t *p, int *pass);

/* ANI/AF of one pair.  chains (may be NULL) receives up to max_chains accepted chains. */
int ora_pair(const ora_sketch_t *a, const ora_sketch_t *b, const ora_params_t *p,
             ora_pair_result_t *out, ora_chain_t *chains, int max_chains, int *n_chains_out);

/* All pairs a<b of n sketches on `threads` host threads (bench.py's CPU arm): prescreen through an inverted marker
 * index (same decisions as ora_screen on every pair), ANI/AF for the survivors, count the edges an
 * `skani triangle --min-af` would print (min_af in [0,1]).  Returns the number of surviving pairs (-1: too many
 * genomes).  pass_out (may be NULL) receives the n*n decision matrix ([a*n+b], a<b).  Wall seconds: t_index = key
 * sort (scales with genomes), t_count = run counting + threshold (scales with shared markers, i.e. with the
 * surviving pairs), t_ani = ANI/AF (scales with the surviving pairs). */
int64_t ora_triangle(const ora_sketch_t *const *sk, int n, double screen, double min_af, const ora_params_t *p,
                     int threads, int64_t *n_edges, uint8_t *pass_out, double *t_index, double *t_count, double *t_ani);

/* One query (sk[q]) against all n sketches on `threads` host threads: one `skani search` call of skDER's low_mem_greedy
 * loop (reference src/skDER/skder.py:119).  Rows (database genome, ANI, AF of the database genome, AF of the query), in
 * no particular order, for pairs that pass the screen, have an estimate and max(AF) >= min_af.  Returns the row count. */
int64_t ora_search(const ora_sketch_t *const *sk, int n, int q, double screen, double min_af, const ora_params_t *p,
                   int threads, int32_t *out_ref, double *out_ani, double *out_af_ref, double *out_af_query);

/* learned-debias substitute: maps raw ANI (+features) to reported ANI */
double ora_debias(double ani_raw);

#ifdef __cplusplus
}
#endif
#endif
