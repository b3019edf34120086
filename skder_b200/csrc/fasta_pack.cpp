// Host-side FASTA ingest and 2-bit packing for the B200 ANI/AF engine.
//
// Stands in for skani's file reader on every path named in skDER's `-l` / `--rl` / `--ql` list
// files (reference src/skDER/skder.py:16, :58, :103) and for the positional FASTA of
// `skani search` (skder.py:119).  Also computes the assembly N50 exactly as the reference's
// util.n50_calc (src/skDER/util.py:686-724), since the packer sees every record length anyway.
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/skani_b200.h"

namespace {

constexpr uint8_t SKIP = 4;  // blanks inside a sequence line are dropped, not packed
struct CodeTable {
    uint8_t t[256];
    CodeTable() {
        std::memset(t, 0, sizeof t);  // anything that is not C/G/T (either case) packs as A
        t[(int)'C'] = t[(int)'c'] = 1;
        t[(int)'G'] = t[(int)'g'] = 2;
        t[(int)'T'] = t[(int)'t'] = 3;
        t[(int)'\r'] = t[(int)' '] = t[(int)'\t'] = SKIP;
    }
};
const CodeTable kCode;

struct Packer {
    std::vector<uint64_t> words;
    uint64_t cur = 0;
    int64_t n_bases = 0;
    inline void put(uint8_t code) {
        cur |= (uint64_t)code << (2 * (n_bases & 31));
        if ((++n_bases & 31) == 0) {
            words.push_back(cur);
            cur = 0;
        }
    }
    // n bytes of one sequence line.  Returns the number of bases packed (blanks are dropped).  The common case -- no
    // blank in the piece -- goes 32 bytes -> one word without a branch per base; a piece with blanks is redone
    // base by base.
    int64_t put_line(const uint8_t *s, int64_t n) {
        const size_t w0 = words.size();
        const uint64_t cur0 = cur;
        const int64_t nb0 = n_bases;
        uint8_t flags = 0;
        int64_t k = 0;
        const int sh = 2 * (int)(n_bases & 31);
        for (; k + 32 <= n; k += 32) {
            uint64_t w = 0;
#pragma GCC unroll 32
            for (int x = 0; x < 32; x++) {
                const uint8_t c = kCode.t[s[k + x]];
                flags |= c;
                w |= (uint64_t)(c & 3) << (2 * x);
            }
            words.push_back(cur | (w << sh));
            cur = sh ? w >> (64 - sh) : 0;
        }
        n_bases += k;
        if (k < n) {  // the last r < 32 bytes: one partial word, merged the same way
            const int r = (int)(n - k);
            uint64_t w = 0;
            for (int x = 0; x < r; x++) {
                const uint8_t c = kCode.t[s[k + x]];
                flags |= c;
                w |= (uint64_t)(c & 3) << (2 * x);
            }
            cur |= w << sh;
            if (sh + 2 * r >= 64) {
                words.push_back(cur);
                cur = sh ? w >> (64 - sh) : 0;
            }
            n_bases += r;
        }
        if (!(flags & SKIP)) return n;
        words.resize(w0);  // rare: blanks inside the line
        cur = cur0;
        n_bases = nb0;
        int64_t kept = 0;
        for (k = 0; k < n; k++) {
            const uint8_t c = kCode.t[s[k]];
            if (c & SKIP) continue;
            put(c);
            kept++;
        }
        return kept;
    }
    void finish() {
        if (n_bases & 31) words.push_back(cur);
        cur = 0;
        if (words.size() & 1) words.push_back(0);  // 16-byte multiple for 128-bit device loads
        if (words.empty()) {
            words.push_back(0);
            words.push_back(0);
        }
    }
};

// N50 as reference util.n50_calc: records with empty sequence are skipped, n2 = floor(total/2),
// first cumulative sum >= n2 over lengths sorted descending.
int64_t n50_of(std::vector<int64_t> lens) {
    lens.erase(std::remove(lens.begin(), lens.end(), (int64_t)0), lens.end());
    if (lens.empty()) return 0;
    std::sort(lens.begin(), lens.end(), std::greater<int64_t>());
    int64_t total = 0;
    for (int64_t l : lens) total += l;
    int64_t n2 = total / 2, c = 0;
    for (int64_t l : lens) {
        c += l;
        if (c >= n2) return l;
    }
    return lens.back();
}

skb_packed *make_packed(Packer &pk, std::vector<int64_t> &kept, const std::string &name,
                        const std::vector<int64_t> &all_lens) {
    pk.finish();
    skb_packed *p = (skb_packed *)std::calloc(1, sizeof(skb_packed));
    if (!p) return nullptr;
    p->n_words = (int64_t)pk.words.size();
    p->words = (uint64_t *)std::malloc(sizeof(uint64_t) * pk.words.size());
    std::memcpy(p->words, pk.words.data(), sizeof(uint64_t) * pk.words.size());
    p->n_bases = pk.n_bases;
    p->n_contigs = (int32_t)kept.size();
    p->contig_lens = (int64_t *)std::malloc(sizeof(int64_t) * std::max<size_t>(kept.size(), 1));
    if (!kept.empty()) std::memcpy(p->contig_lens, kept.data(), sizeof(int64_t) * kept.size());
    p->first_name = strdup(name.c_str());
    p->n50 = n50_of(all_lens);
    int64_t t = 0;
    for (int64_t l : all_lens) t += l;
    p->total_bases_all = t;
    return p;
}

}  // namespace

extern "C" {

int skb_pack_fasta(const char *path, int32_t min_contig_len, skb_packed **out) {
    if (!path || !out) return SKB_EINVAL;
    *out = nullptr;
    gzFile g = gzopen(path, "rb");  // transparently reads plain files too
    if (!g) return SKB_EIO;
    gzbuffer(g, 1 << 20);
    std::vector<char> buf(1 << 22);
    // Records are packed straight into the output; a record that turns out shorter than
    // min_contig_len is rolled back.
    Packer pk;
    pk.words.reserve(1 << 18);
    std::vector<int64_t> kept, all_lens;
    std::string first_name, cur_name;
    bool have_first = false, in_header = false, in_record = false;
    int64_t rec_len = 0, rec_start_bases = 0;
    size_t rec_start_words = 0;
    uint64_t rec_start_cur = 0;
    auto close_record = [&]() {
        if (!in_record) return;
        all_lens.push_back(rec_len);
        if (rec_len >= min_contig_len && rec_len > 0) {
            kept.push_back(rec_len);
            if (!have_first) {
                first_name = cur_name;
                have_first = true;
            }
        } else {  // roll back
            pk.words.resize(rec_start_words);
            pk.cur = rec_start_cur;
            pk.n_bases = rec_start_bases;
        }
    };
    bool at_line_start = true;
    for (;;) {
        int got = gzread(g, buf.data(), (unsigned)buf.size());
        if (got < 0) {
            gzclose(g);
            return SKB_EIO;
        }
        if (got == 0) break;
        const char *p = buf.data(), *const end = p + got;
        while (p < end) {
            if (in_header) {  // the rest of the header line
                const char *nl = (const char *)std::memchr(p, '\n', (size_t)(end - p));
                cur_name.append(p, nl ? nl : end);
                if (!nl) break;
                in_header = false;
                at_line_start = true;
                while (!cur_name.empty() && cur_name.back() == '\r') cur_name.pop_back();
                p = nl + 1;
                continue;
            }
            if (at_line_start && *p == '>') {
                close_record();
                in_record = true;
                in_header = true;
                cur_name.clear();
                rec_len = 0;
                rec_start_bases = pk.n_bases;
                rec_start_words = pk.words.size();
                rec_start_cur = pk.cur;
                p++;
                continue;
            }
            // (the rest of) a sequence line; a '>' that is not first on its line is an ordinary non-ACGT byte
            const char *nl = (const char *)std::memchr(p, '\n', (size_t)(end - p));
            const char *e = nl ? nl : end;
            if (in_record && e > p) rec_len += pk.put_line((const uint8_t *)p, (int64_t)(e - p));  // else: text before the first header
            at_line_start = nl != nullptr;
            p = nl ? nl + 1 : end;
        }
    }
    gzclose(g);
    close_record();
    *out = make_packed(pk, kept, first_name, all_lens);
    return *out ? SKB_OK : SKB_ENOMEM;
}

int skb_pack_contigs(const char *const *seqs, const int64_t *lens, int32_t n, int32_t min_contig_len,
                     skb_packed **out) {
    if (!out || n < 0) return SKB_EINVAL;
    Packer pk;
    std::vector<int64_t> kept, all_lens;
    for (int32_t i = 0; i < n; i++) {
        all_lens.push_back(lens[i]);
        if (lens[i] < min_contig_len || lens[i] <= 0) continue;
        kept.push_back(lens[i]);
        const uint8_t *s = (const uint8_t *)seqs[i];
        for (int64_t j = 0; j < lens[i]; j++) pk.put(kCode.t[s[j]]);
    }
    *out = make_packed(pk, kept, "", all_lens);
    return *out ? SKB_OK : SKB_ENOMEM;
}

int skb_pack_fasta_many(const char *const *paths, int32_t n, int32_t min_contig_len, int32_t n_threads,
                        skb_packed **out) {
    if (!paths || !out || n < 0) return SKB_EINVAL;
    if (n_threads < 1) n_threads = 1;
    std::atomic<int32_t> next(0), failures(0);
    auto work = [&]() {
        for (;;) {
            int32_t i = next.fetch_add(1);
            if (i >= n) break;
            out[i] = nullptr;
            if (skb_pack_fasta(paths[i], min_contig_len, &out[i]) != SKB_OK) failures.fetch_add(1);
        }
    };
    std::vector<std::thread> th;
    for (int32_t t = 1; t < std::min(n_threads, std::max(n, 1)); t++) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
    return failures.load();
}

void skb_packed_free(skb_packed *p) {
    if (!p) return;
    std::free(p->words);
    std::free(p->contig_lens);
    std::free(p->first_name);
    std::free(p);
}

}  // extern "C"
