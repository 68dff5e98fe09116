"""Seeded synthetic metagenomes for the abawaca hot path.

There is no network, so every workload is synthetic: scaffolds drawn from
synthetic genomes with genome-specific oligonucleotide composition, synthetic
per-sample coverage (read records), and synthetic single-copy-gene (SCG)
placements.  The same object can be

  * handed to the CUDA path / the flat-array oracle as flat numpy arrays, or
  * written as the text files the reference binaries read (FASTA, SAM,
    gene2scg, scg.list) -- see `write_reference_inputs`.

Scaffolds are emitted in byte-wise name order, because both reference programs
order scaffolds through `std::map<std::string, ...>`
(abawaca-build.cpp:482,596; ScafDpData.cpp:59,91).
"""
from __future__ import annotations

import dataclasses
import os

import numpy as np

MASTER_SEED = 20261018
READ_DTYPE = np.dtype([("scaf", "<u4"), ("pos0", "<u4"), ("len", "<u4"), ("flag_nsnps", "<u4")])

_TRIMERS = np.array([[ord("ACGT"[(t >> 4) & 3]), ord("ACGT"[(t >> 2) & 3]), ord("ACGT"[t & 3])] for t in range(64)], dtype=np.uint8)


@dataclasses.dataclass
class Metagenome:
    names: list            # scaffold names, byte-wise sorted
    genome: np.ndarray     # [S] genome of each scaffold
    offsets: np.ndarray    # [S+1] uint64 offsets into `seq`
    seq: np.ndarray        # uint8 ASCII, concatenated scaffolds
    reads: list            # per sample: structured array READ_DTYPE, in "SAM order"
    scg_names: list        # the SCG vocabulary (scg.list)
    gene2scg: list         # (gene name "<scaf>_<idx>", scg name)
    read_len: int
    seed: int

    @property
    def nscaf(self):
        return len(self.names)

    def scaffold(self, i):
        return self.seq[int(self.offsets[i]):int(self.offsets[i + 1])]

    def scg_masks(self, scaf_index_of_name=None):
        """uint64 [S][W] bit masks of SCG names per scaffold (sets, SCGdb.h:25)."""
        idx = {n: i for i, n in enumerate(self.names)} if scaf_index_of_name is None else scaf_index_of_name
        gidx = {n: i for i, n in enumerate(sorted(set(s for _, s in self.gene2scg)))}
        W = max(1, (len(gidx) + 63) // 64)
        masks = np.zeros((len(idx), W), dtype=np.uint64)
        for gene, scg in self.gene2scg:
            scaf = gene[:gene.rfind("_")]
            if scaf in idx:
                g = gidx[scg]
                masks[idx[scaf], g // 64] |= np.uint64(1) << np.uint64(g % 64)
        return masks


def make_metagenome(n_scaffolds, n_samples, n_genomes, seed, *, min_len=4000, mean_extra=6000, max_len=200000,
                    read_len=150, cov_lo=0.5, cov_hi=16.0, n_run_frac=0.005, n_scg=51, scg_p=0.9,
                    bad_read_frac=0.03, q6_reads=False, shuffle_reads=False, with_reads=True,
                    gc_lo=0.25, gc_hi=0.75, tri_sigma=0.6, shard=None, bimodal_frac=0.0, bimodal_factor=4.0) -> Metagenome:
    """shard: when given, the genome models (composition, coverage) still come from `seed` alone, but the scaffolds, reads and gene
    placements are drawn from (seed, shard): several shards are then pieces of ONE community, which is how bench.py builds the
    N-GPU workload (every rank generates only its own scaffolds).
    bimodal_frac: that share of the genomes gets two coverage modes (half of their scaffolds are covered bimodal_factor times deeper in every
    sample) while their single-copy genes stay spread over all their scaffolds: a clean separation INSIDE one genome, which only the SCG test of
    ClusterQuality::is_split_better (disjoint SCG sets on the two sides) turns down.  Drawn from its own generator: 0 leaves every other set as it was."""
    rng = np.random.default_rng(seed)
    # genome models
    gc = rng.uniform(gc_lo, gc_hi, n_genomes)
    base_p = np.stack([(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2], axis=1)  # A C G T
    tri_p = np.empty((n_genomes, 64))
    for t in range(64):
        tri_p[:, t] = base_p[:, (t >> 4) & 3] * base_p[:, (t >> 2) & 3] * base_p[:, t & 3]
    tri_p *= np.exp(rng.normal(0.0, tri_sigma, (n_genomes, 64)))
    tri_p /= tri_p.sum(axis=1, keepdims=True)
    cov = np.exp(rng.uniform(np.log(cov_lo), np.log(cov_hi), (n_genomes, max(n_samples, 1))))
    tag = ""
    if shard is not None:
        rng = np.random.default_rng([seed, 7919 + int(shard)])
        tag = f"r{int(shard)}_"

    genome = rng.integers(0, n_genomes, n_scaffolds)
    lengths = np.minimum(min_len + rng.exponential(mean_extra, n_scaffolds).astype(np.int64), max_len)
    names = [f"{tag}g{g}_scaffold_{i}" for i, g in enumerate(genome)]
    order = sorted(range(n_scaffolds), key=lambda i: names[i].encode())
    names = [names[i] for i in order]
    genome = genome[order]
    lengths = lengths[order]
    offsets = np.zeros(n_scaffolds + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lengths)
    seq = np.empty(int(offsets[-1]), dtype=np.uint8)
    ntri = (lengths + 2) // 3
    for g in range(n_genomes):
        idx = np.nonzero(genome == g)[0]
        if idx.size == 0:
            continue
        total = int(ntri[idx].sum())
        tri = rng.choice(64, size=total, p=tri_p[g]).astype(np.uint8)
        bases = _TRIMERS[tri].reshape(-1)
        o = 0
        for i in idx:
            L = int(lengths[i])
            seq[int(offsets[i]):int(offsets[i]) + L] = bases[o:o + L]
            o += int(ntri[i]) * 3
    # a few runs of N
    n_with_n = int(round(n_run_frac * n_scaffolds))
    if n_with_n > 0:
        for i in rng.choice(n_scaffolds, size=n_with_n, replace=False):
            L = int(lengths[i])
            run = int(rng.integers(10, 101))
            p = int(rng.integers(0, L - run))
            seq[int(offsets[i]) + p:int(offsets[i]) + p + run] = ord("N")

    cov_mult = np.ones(n_scaffolds)
    if bimodal_frac > 0:
        rng2 = np.random.default_rng([seed, 424243] + ([int(shard)] if shard is not None else []))
        bimodal = rng2.random(n_genomes) < bimodal_frac
        deep = rng2.random(n_scaffolds) < 0.5
        cov_mult = np.where(bimodal[genome] & deep, bimodal_factor, 1.0)
    # reads, one record stream per sample, scaffold-major with random positions
    reads = []
    if with_reads:
        for j in range(n_samples):
            lam = cov[genome, j] * cov_mult * lengths / read_len
            n = np.floor(lam + rng.random(n_scaffolds)).astype(np.int64)
            if q6_reads and j == n_samples - 1:
                extra = (lengths + 999) // 1000
            else:
                extra = np.zeros(n_scaffolds, dtype=np.int64)
            tot = n + extra
            rec = np.zeros(int(tot.sum()), dtype=READ_DTYPE)
            scaf_of = np.repeat(np.arange(n_scaffolds, dtype=np.uint32), tot)
            start_of = np.repeat(np.cumsum(tot) - tot, tot)
            k = np.arange(rec.size) - start_of           # index of the read within its scaffold
            L_of = lengths[scaf_of]
            n_of = n[scaf_of]
            span = np.maximum(L_of - read_len, 0) + 1
            pos = (rng.random(rec.size) * span).astype(np.int64)
            is_extra = k >= n_of
            pos = np.where(is_extra, np.minimum((k - n_of) * 1000, np.maximum(L_of - read_len, 0)), pos)
            flag = np.zeros(rec.size, dtype=np.uint32)
            nsnps = np.zeros(rec.size, dtype=np.uint32)
            if bad_read_frac > 0:
                u = rng.random(rec.size)
                b = bad_read_frac / 3
                flag = np.where((u < b) & ~is_extra, 0x100, flag)            # secondary alignment
                flag = np.where((u >= b) & (u < 2 * b) & ~is_extra, 0x4, flag)  # unmapped
                nsnps = np.where((u >= 2 * b) & (u < 3 * b) & ~is_extra, 16, nsnps)
                nsnps = np.where((u >= 3 * b) & (u < 4 * b), 3, nsnps)     # a few tolerated mismatches
            rec["scaf"] = scaf_of
            rec["pos0"] = pos.astype(np.uint32)
            rec["len"] = read_len
            rec["flag_nsnps"] = (flag & 0xFFFF) | (nsnps.astype(np.uint32) << 16)
            if shuffle_reads:
                rec = rec[rng.permutation(rec.size)]
            reads.append(rec)

    # single-copy genes
    scg_names = [f"SCG{i + 1:02d}" for i in range(n_scg)]
    gene2scg = []
    gene_counter = {}
    for g in range(n_genomes):
        idx = np.nonzero(genome == g)[0]
        if idx.size == 0:
            continue
        for s in range(n_scg):
            if rng.random() < scg_p:
                i = int(idx[rng.integers(0, idx.size)])
                gene_counter[i] = gene_counter.get(i, 0) + 1
                gene2scg.append((f"{names[i]}_{gene_counter[i]}", scg_names[s]))
    return Metagenome(names, genome, offsets, seq, reads, scg_names, gene2scg, read_len, seed)


CONFIGS = {
    # BASELINE.json configs -> generator arguments
    "cfg1": dict(n_scaffolds=2000, n_samples=3, n_genomes=8, seed=MASTER_SEED + 1),
    "cfg2": dict(n_scaffolds=50000, n_samples=10, n_genomes=32, seed=MASTER_SEED + 2),
    "cfg3": dict(n_scaffolds=500000, n_samples=20, n_genomes=128, seed=MASTER_SEED + 3),
    "cfg4": dict(n_scaffolds=1000000, n_samples=50, n_genomes=256, seed=MASTER_SEED + 4),
}


def write_reference_inputs(mg: Metagenome, directory: str):
    """Write FASTA + one SAM per sample + gene2scg + scg.list, the text inputs of the reference binaries.

    SAM lines have the >= 11 tab fields `ReadMapping(const char*)` needs
    (ReadMapping.cpp:36-69).  Mismatches are encoded in MD:Z so that
    `num_snps()` equals the record's nsnps.
    """
    os.makedirs(directory, exist_ok=True)
    fa = os.path.join(directory, "assembly.fa")
    with open(fa, "w") as f:
        for i, name in enumerate(mg.names):
            s = mg.scaffold(i).tobytes().decode()
            f.write(f">{name}\n")
            for o in range(0, len(s), 80):
                f.write(s[o:o + 80] + "\n")
    sams = []
    for j, rec in enumerate(mg.reads):
        path = os.path.join(directory, f"sample{j:02d}.sam")
        sams.append(path)
        with open(path, "w") as f:
            f.write("@HD\tVN:1.0\tSO:unsorted\n")
            for r, (scaf, pos0, ln, fn) in enumerate(rec.tolist()):
                flag, nsnps = fn & 0xFFFF, fn >> 16
                seq = "A" * ln
                qual = "I" * ln
                if nsnps == 0:
                    md = f"MD:Z:{ln}"
                else:  # nsnps isolated mismatches at the start of the read
                    md = "MD:Z:0" + "".join("C0" for _ in range(nsnps - 1)) + f"C{ln - nsnps}"
                rname = "*" if (flag & 0x4) else mg.names[scaf]
                pos1 = 0 if (flag & 0x4) else pos0 + 1
                cigar = "*" if (flag & 0x4) else f"{ln}M"
                tail = "" if (flag & 0x4) else "\t" + md
                f.write(f"r{j}_{r}/1\t{flag}\t{rname}\t{pos1}\t42\t{cigar}\t*\t0\t0\t{seq}\t{qual}{tail}\n")
    g2s = os.path.join(directory, "genes.scg")
    with open(g2s, "w") as f:
        for gene, scg in mg.gene2scg:
            f.write(f"{gene}\t{scg}\n")
    lst = os.path.join(directory, "scg.list")
    with open(lst, "w") as f:
        for s in mg.scg_names:
            f.write(s + "\n")
    return dict(fasta=fa, sams=sams, gene2scg=g2s, scg_list=lst)
