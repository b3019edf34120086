"""Deterministic synthetic genome sets of the shapes BASELINE.json names.

A clade = one random ancestor; each member = the ancestor with point substitutions at rate d_i,
five random 10 kb deletions and one 20 kb random insertion, cut into n_c contigs
(n_c ~ logUniform[1, 300]).  Different clades are unrelated.  PRNG: numpy PCG64.
"""
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _mutate(anc, d, rng, n_del=5, del_len=10000, ins_len=20000):
    g = anc.copy()
    n = len(g)
    nsub = rng.binomial(n, d)
    if nsub:
        pos = rng.integers(0, n, nsub)
        g[pos] = (g[pos] + rng.integers(1, 4, nsub).astype(np.uint8)) & 3
    if n > 20 * del_len:
        keep = np.ones(n, bool)
        for s in rng.integers(0, n - del_len, n_del):
            keep[s:s + del_len] = False
        g = g[keep]
        at = int(rng.integers(0, len(g)))
        g = np.concatenate([g[:at], rng.integers(0, 4, ins_len, dtype=np.uint8), g[at:]])
    return g


def _cut(g, n_contigs, rng, min_len=1000):
    if n_contigs <= 1 or len(g) < 4 * min_len:
        return [g]
    cuts = np.sort(rng.integers(min_len, len(g) - min_len, n_contigs - 1))
    cuts = np.concatenate([[0], cuts, [len(g)]])
    return [g[a:b] for a, b in zip(cuts[:-1], cuts[1:]) if b - a >= min_len]


def one_clade(c, per_clade, length, seed, d_lo=0.0005, d_hi=0.025, max_contigs=300, length_hi=None):
    """Members of clade c as lists of contig byte strings.  Every clade has its own PCG64 stream
    (seed, c), so clades can be generated in any order or in parallel."""
    rng = np.random.Generator(np.random.PCG64([seed, c]))
    L = length if length_hi is None else int(rng.integers(length, length_hi))
    anc = rng.integers(0, 4, L, dtype=np.uint8)
    out = []
    for _ in range(per_clade):
        d = rng.uniform(d_lo, d_hi)
        g = _mutate(anc, d, rng)
        nc = int(np.exp(rng.uniform(0, np.log(max_contigs))))
        out.append([ACGT[x].tobytes() for x in _cut(g, nc, rng)])
    return out


def clade_genomes(n_clades, per_clade, length, seed, d_lo=0.0005, d_hi=0.025, max_contigs=300, length_hi=None):
    """Yield (clade, member, [contig byte strings]) in a fixed order."""
    for c in range(n_clades):
        for m, contigs in enumerate(one_clade(c, per_clade, length, seed, d_lo, d_hi, max_contigs, length_hi)):
            yield c, m, contigs


CONFIGS = {
    # name: (n_clades, per_clade, length, length_hi, d_lo, d_hi, seed)
    "tiny": (3, 4, 200_000, None, 0.0005, 0.025, 20261017),
    "config2": (20, 50, 5_000_000, None, 0.0005, 0.025, 20261017 + 2),
    "config3": (100, 50, 5_000_000, None, 0.0005, 0.025, 20261017 + 3),
    "config4": (200, 100, 2_800_000, None, 0.0005, 0.01, 20261017 + 4),
    "config5": (500, 100, 2_000_000, 8_000_000, 0.0005, 0.025, 20261017 + 5),
}


def config_genomes(name, n_clades=None, per_clade=None):
    """Genomes of a named configuration; n_clades / per_clade override its shape (bounded samples)."""
    nc, per, L, Lhi, dlo, dhi, seed = CONFIGS[name]
    return clade_genomes(n_clades or nc, per_clade or per, L, seed, dlo, dhi, 300, Lhi)


def config_packed(name, pack, n_clades=None, per_clade=None, threads=None, clade_offset=0):
    """Generate a configuration clade-parallel on host threads and pack each genome with `pack`
    (a callable taking the list of contig byte strings).  Returns the packed genomes in fixed order."""
    import os
    from concurrent.futures import ThreadPoolExecutor

    nc, per, L, Lhi, dlo, dhi, seed = CONFIGS[name]
    nc, per = n_clades or nc, per_clade or per

    def work(c):
        return [pack(g) for g in one_clade(c + clade_offset, per, L, seed, dlo, dhi, 300, Lhi)]

    with ThreadPoolExecutor(threads or min(32, os.cpu_count() or 1)) as ex:
        res = list(ex.map(work, range(nc)))
    return [g for clade in res for g in clade]


def write_fasta(path, contigs, name="ctg", width=80):
    with open(path, "wb") as f:
        for i, s in enumerate(contigs):
            f.write(b">%s_%d\n" % (name.encode(), i + 1))
            for j in range(0, len(s), width):
                f.write(s[j:j + width])
                f.write(b"\n")
