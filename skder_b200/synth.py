"""Deterministic synthetic genome sets of the shapes BASELINE.json names.

A clade = one random ancestor; each member = the ancestor with point substitutions at rate d_i,
five random 10 kb deletions and one 20 kb random insertion, cut into n_c contigs
(n_c ~ logUniform[1, 300]).  Different clades are unrelated.  PRNG: numpy PCG64.
"""
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _small_indels(g, rate, rng, max_len=12):
    """Insertions / deletions of 1..max_len bases at `rate` per base (half each): chains pick up non-zero gaps."""
    n = len(g)
    k = rng.binomial(n, rate)
    if k == 0:
        return g
    pos = np.sort(rng.integers(1, n - max_len - 1, k))
    lens = np.minimum(rng.geometric(0.4, k), max_len)
    is_ins = rng.random(k) < 0.5
    pieces, at = [], 0
    for p_, l_, ins in zip(pos.tolist(), lens.tolist(), is_ins.tolist()):
        if p_ < at:
            continue
        pieces.append(g[at:p_])
        if ins:
            pieces.append(rng.integers(0, 4, l_, dtype=np.uint8))
            at = p_
        else:
            at = p_ + l_
    pieces.append(g[at:])
    return np.concatenate(pieces)


def _add_repeats(anc, rng, families=((1500, 6), (1200, 4), (900, 12), (2500, 3))):
    """IS-element-like repeat families pasted over the ancestor: (length, copies).  Copy numbers straddle the
    multiplicity cap (8): multi-hit seeds, staged probes and, for the 12-copy family, repeat-flagged seeds."""
    g = anc.copy()
    n = len(g)
    for length, copies in families:
        if n < 40 * length:
            continue
        unit = rng.integers(0, 4, length, dtype=np.uint8)
        for at in rng.integers(0, n - length, copies).tolist():
            g[at:at + length] = unit
    return g


def _mutate(anc, d, rng, n_del=5, del_len=10000, ins_len=20000, indel_rate=0.0):
    g = anc.copy()
    n = len(g)
    nsub = rng.binomial(n, d)
    if nsub:
        pos = rng.integers(0, n, nsub)
        g[pos] = (g[pos] + rng.integers(1, 4, nsub).astype(np.uint8)) & 3
    if indel_rate > 0:
        g = _small_indels(g, indel_rate, rng)
        n = len(g)
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


def one_clade(c, per_clade, length, seed, d_lo=0.0005, d_hi=0.025, max_contigs=300, length_hi=None, realistic=False):
    """Members of clade c as lists of contig byte strings.  Every clade has its own PCG64 stream
    (seed, c), so clades can be generated in any order or in parallel.  realistic: the ancestor carries repeat families
    and members also differ by small indels (one per eight substitutions)."""
    rng = np.random.Generator(np.random.PCG64([seed, c]))
    L = length if length_hi is None else int(rng.integers(length, length_hi))
    anc = rng.integers(0, 4, L, dtype=np.uint8)
    if realistic:
        anc = _add_repeats(anc, rng)
    out = []
    for _ in range(per_clade):
        d = rng.uniform(d_lo, d_hi)
        g = _mutate(anc, d, rng, indel_rate=d / 8 if realistic else 0.0)
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
    # config3 with repeat families in every ancestor and small indels between members (REALISTIC below): the slow paths
    # of the probe (multi-hit seeds, repeat flags) and of the chaining DP (non-zero gaps) are on the timed path
    "config3r": (100, 50, 5_000_000, None, 0.0005, 0.025, 20261017 + 33),
    "tinyr": (3, 4, 200_000, None, 0.0005, 0.025, 20261017 + 34),
    "tiny4": (4, 6, 150_000, None, 0.0002, 0.004, 20261017 + 35),  # config4's shape in small: close relatives, search path
    "config4": (200, 100, 2_800_000, None, 0.0005, 0.01, 20261017 + 4),
    "config5": (500, 100, 2_000_000, 8_000_000, 0.0005, 0.025, 20261017 + 5),
}


REALISTIC = {"config3r", "tinyr"}


def clade_of(name, c, per_clade=None):
    """Members of clade c of a named configuration (per_clade overrides the clade size: bounded samples)."""
    nc, per, L, Lhi, dlo, dhi, seed = CONFIGS[name]
    return one_clade(c, per_clade or per, L, seed, dlo, dhi, 300, Lhi, realistic=name in REALISTIC)


def config_genomes(name, n_clades=None, per_clade=None):
    """Genomes of a named configuration; n_clades / per_clade override its shape (bounded samples)."""
    nc = CONFIGS[name][0]
    for c in range(n_clades or nc):
        for m, contigs in enumerate(clade_of(name, c, per_clade)):
            yield c, m, contigs


def config_packed(name, pack, n_clades=None, per_clade=None, threads=None, clade_offset=0):
    """Generate a configuration clade-parallel on host threads and pack each genome with `pack`
    (a callable taking the list of contig byte strings).  Returns the packed genomes in fixed order."""
    import os
    from concurrent.futures import ThreadPoolExecutor

    nc, per, L, Lhi, dlo, dhi, seed = CONFIGS[name]
    nc, per = n_clades or nc, per_clade or per

    def work(c):
        return [pack(g) for g in clade_of(name, c + clade_offset, per)]

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
