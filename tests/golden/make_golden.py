#!/usr/bin/env python
"""Generate the committed golden fixtures from the UNMODIFIED reference (oracle/_ref).

Runs only where /root/reference was available to build oracle/_ref (the development
container).  The GPU box has neither, so tests read the fixtures written here:

  kat_*.json          known-answer vectors: hand-made FASTA/SAM edge cases -> reference outputs
  kat_coverage_order  windows of round lengths: the order of the reads decides the third decimal of the coverage (quirk Q5)
  tiny_clean.*        400 scaffolds / 3 samples / 4 well separated genomes
  tiny_noisy.*        700 scaffolds / 4 samples / 14 similar genomes (imperfect scores, SCG filter active)

For the synthetic sets the inputs are regenerated from the seed by abawaca_b200.synth; the
fixture stores sha256 digests of the generated arrays so that generator drift is detected.

usage: python tests/golden/make_golden.py
"""
import gzip
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from abawaca_b200 import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")

SETS = {
    "tiny_clean": dict(n_scaffolds=400, n_samples=3, n_genomes=4, seed=123, q6_reads=True, shuffle_reads=True),
    "tiny_noisy": dict(n_scaffolds=700, n_samples=4, n_genomes=14, seed=77, q6_reads=True, shuffle_reads=True,
                       gc_lo=0.40, gc_hi=0.60, tri_sigma=0.12, cov_lo=2.0, cov_hi=8.0, mean_extra=9000),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def digests(mg):
    return dict(seq=sha(mg.seq), offsets=sha(mg.offsets), reads=[sha(r) for r in mg.reads],
                names=hashlib.sha256("\n".join(mg.names).encode()).hexdigest(),
                gene2scg=hashlib.sha256(repr(mg.gene2scg).encode()).hexdigest())


def run_reference(workdir, paths, this_sample_index=0, strategies=("sensspec", "splitscafs")):
    build = os.path.join(workdir, "build")
    out = os.path.join(workdir, "out")
    shutil.rmtree(build, ignore_errors=True)
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(build)
    glob = os.path.join(workdir, "sample*.sam")
    subprocess.run([os.path.join(REF, "abawaca-build"), "-f", paths["fasta"], "-o", build, "-s", glob, "-c", paths["sams"][this_sample_index]],
                   check=True, stderr=subprocess.DEVNULL)
    env = dict(os.environ, ABW_SCG_LIST=paths["scg_list"])
    subprocess.run([os.path.join(REF, "abawaca"), "-u", build, "-o", out, "-c", paths["gene2scg"], "-p", "8"], check=True, env=env,
                   stderr=subprocess.DEVNULL)
    res = {}
    for s in strategies:
        tsv = os.path.join(workdir, f"ref_{s}.tsv")
        subprocess.run([os.path.join(REF, "ref_search"), os.path.join(build, "abawaca.names"), paths["fasta"], os.path.join(build, "abawaca.info"),
                        os.path.join(build, "abawaca.lrn"), paths["gene2scg"], paths["scg_list"], s, "8", tsv], check=True)
        res[s] = open(tsv).read()
    raw = os.path.join(workdir, "raw.tsv")
    subprocess.run([os.path.join(REF, "ref_features"), paths["fasta"], raw] + paths["sams"], check=True)
    return build, out, res, raw


def gz_copy(src, dst):
    with open(src, "rb") as f, gzip.GzipFile(dst, "wb", mtime=0) as g:
        g.write(f.read())


def make_synthetic(name, args):
    mg = synth.make_metagenome(**args)
    with tempfile.TemporaryDirectory() as wd:
        paths = synth.write_reference_inputs(mg, wd)
        build, out, res, raw = run_reference(wd, paths)
        gz_copy(os.path.join(build, "abawaca.lrn"), os.path.join(HERE, f"{name}.lrn.gz"))
        gz_copy(os.path.join(build, "abawaca.names"), os.path.join(HERE, f"{name}.names.gz"))
        gz_copy(os.path.join(build, "abawaca.info"), os.path.join(HERE, f"{name}.info.gz"))
        # un-truncated coverage columns (k-mer columns are implied by the .lrn and the KAT vectors)
        rows = []
        for line in open(raw):
            if line.startswith("SEG"):
                rows.append([float(x) for x in line.rstrip("\n").split("\t")[3 + 180:]])
        np.save(os.path.join(HERE, f"{name}.rawcov.npy"), np.array(rows, dtype=np.float64))
        log = [l.split("]", 1)[1].strip() for l in open(os.path.join(out, "log")) if "Best separation" in l]
        meta = dict(args=args, digests=digests(mg), sample_names=paths["sams"], this_sample=0,
                    scaf2cluster=open(os.path.join(out, "scaf2cluster.txt")).read(),
                    summary=open(os.path.join(out, "summary.txt")).read(),
                    best_separation_log=log, ref_search=res)
        meta["sample_names"] = [os.path.basename(p) for p in paths["sams"]]
        with open(os.path.join(HERE, f"{name}.json"), "w") as f:
            json.dump(meta, f, indent=1)
    print(name, "done:", len(log), "splits")


def kat_sequences():
    rng = np.random.default_rng(5)

    def rnd(n, alphabet="ACGT"):
        return "".join(alphabet[i] for i in rng.integers(0, len(alphabet), n))
    seqs = [
        ("k01", "ACGT" * 600),
        ("k02", "AACCGGTT" * 550),
        ("k03", "ACGTN" * 1000),
        ("k04", "acgtR" * 900),
        ("k05", "A" * 4001),
        ("k06", rnd(150)),                                   # one short segment
        ("k07", "N" * 50 + rnd(5000) + "N" * 30 + rnd(3000)),  # N runs at the start and inside
        ("k08", "N" * 7),                                    # no non-N base: every character closes a segment
        ("k09", rnd(4000)),                                  # exactly two windows
        ("k10", rnd(3999)),                                  # one window of 3999
        ("k11", rnd(6100, "ACGTacgtRYKMSWryk")),             # IUPAC + lower case
        ("k12", "ACG"),                                      # shorter than k=4
        ("k13", "T"),
        ("k14", rnd(2100) + "N" * 2500 + rnd(2300)),         # a long N run inside a segment
        ("k15", rnd(12345)),
    ]
    return seqs


def make_kat():
    seqs = kat_sequences()
    with tempfile.TemporaryDirectory() as wd:
        fa = os.path.join(wd, "assembly.fa")
        with open(fa, "w") as f:
            for n, s in seqs:
                f.write(f">{n} some description\n")
                for o in range(0, len(s), 60):
                    f.write(s[o:o + 60] + "\n")
        # sample00: systematic reads so that every segment is hit (keeps the reference away from its
        # unchecked read past the dimensions vector, quirk Q6); sample01: the survey's five reads + edge cases
        lines0, lines1 = ["@HD\tVN:1.0"], ["@HD\tVN:1.0"]

        def sam(name, flag, rname, pos1, ln, md=None):
            md = md or f"MD:Z:{ln}"
            return f"{name}\t{flag}\t{rname}\t{pos1}\t42\t{ln}M\t*\t0\t0\t{'A' * ln}\t{'I' * ln}\t{md}"
        r = 0
        for n, s in seqs:
            L = len(s)
            ln = min(100, L)
            for pos1 in sorted(set([1, 2] + list(range(1, max(L - ln + 2, 2), 333)) + [max(L - ln + 1, 1)])):
                if pos1 + ln - 1 <= L + 50:
                    lines0.append(sam(f"s{r}/1", 0, n, pos1, ln))
                    lines1.append(sam(f"t{r}/1", 16, n, pos1, ln))
                    r += 1
        lines1 += [
            sam("r1/1", 0, "k02", 1, 100),
            sam("r2/1", 0, "k02", 2151, 100),
            sam("r3/1", 256, "k02", 10, 100),
            "r4/1\t4\t*\t0\t0\t*\t*\t0\t0\t" + "A" * 100 + "\t" + "I" * 100,
            sam("r5/1", 0, "k02", 20, 100, "MD:Z:0" + "C0" * 15 + "C84"),      # 16 mismatches: dropped
            sam("r6/1", 0, "k02", 30, 100, "MD:Z:0" + "C0" * 14 + "C85"),      # 15 mismatches: kept
            sam("r7/1", 0, "k05", 3950, 100),                                   # hangs over the dropped tail
            sam("r8/1", 0, "k05", 4001, 1),                                     # only in the dropped tail
            sam("r9/1", 0, "k15", 12300, 100),                                  # hangs over the scaffold end
            sam("r10/1", 0, "nosuchscaffold", 5, 100),
            sam("r11/1", 0, "k09", 1951, 100),                                  # straddles the two windows of k09
            sam("r12/1", 1024, "k09", 1951, 100),                               # duplicate flag is NOT filtered
        ]
        s0 = os.path.join(wd, "sample00.sam")
        s1 = os.path.join(wd, "sample01.sam")
        open(s0, "w").write("\n".join(lines0) + "\n")
        open(s1, "w").write("\n".join(lines1) + "\n")
        build = os.path.join(wd, "build")
        os.makedirs(build)
        subprocess.run([os.path.join(REF, "abawaca-build"), "-f", fa, "-o", build, "-s", os.path.join(wd, "sample*.sam"), "-c", s0], check=True,
                       stderr=subprocess.DEVNULL)
        raw = os.path.join(wd, "raw.tsv")
        subprocess.run([os.path.join(REF, "ref_features"), fa, raw, s0, s1], check=True)
        reads = []
        for path in (s0, s1):
            rr = []
            for line in open(path):
                if line.startswith("@"):
                    continue
                f = line.rstrip("\n").split("\t")
                md = [x for x in f[11:] if x.startswith("MD:Z:")]
                nsnps = sum(1 for c in md[0][5:] if c.isalpha()) if md else 0
                rr.append([f[2], int(f[3]), len(f[9]), int(f[1]), nsnps])
            reads.append(rr)
        kat = dict(seqs=seqs, reads=reads, sample_names=["sample00.sam", "sample01.sam"],
                   lrn=open(os.path.join(build, "abawaca.lrn")).read(),
                   names=open(os.path.join(build, "abawaca.names")).read(),
                   info=open(os.path.join(build, "abawaca.info")).read(),
                   raw=open(raw).read())
        with gzip.GzipFile(os.path.join(HERE, "kat_features.json.gz"), "wb", mtime=0) as g:
            g.write(json.dumps(kat).encode())
    print("kat done")


COVERAGE_ORDER_ARGS = dict(seed=11, nscaf=60, nsamples=2, depth=25, cover_all_n=True)


def make_coverage_order():
    """Round window lengths (2000, 2048, 2500, 3125 ...): for many windows 1000 * (sum of overlaps) / length is an integer, so the order in which
    the reference adds fl(overlap / length) decides the third decimal (quirk Q5).  Inputs come from tests/golden_util.coverage_edge_workload;
    the fixture keeps the reference's .lrn and its un-truncated coverage columns."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from golden_util import coverage_edge_workload
    seq, offsets, reads = coverage_edge_workload(**COVERAGE_ORDER_ARGS)
    nscaf = offsets.size - 1
    names = ["e%04d" % i for i in range(nscaf)]
    with tempfile.TemporaryDirectory() as wd:
        fa = os.path.join(wd, "assembly.fa")
        with open(fa, "w") as f:
            for i, n in enumerate(names):
                s = seq[int(offsets[i]):int(offsets[i + 1])].tobytes().decode()
                f.write(f">{n}\n")
                for o in range(0, len(s), 80):
                    f.write(s[o:o + 80] + "\n")
        sams = []
        for j, r in enumerate(reads):
            path = os.path.join(wd, f"sample{j:02d}.sam")
            with open(path, "w") as f:
                f.write("@HD\tVN:1.0\n")
                for k in range(r.size):
                    sc, pos0, ln, fn = int(r["scaf"][k]), int(r["pos0"][k]), int(r["len"][k]), int(r["flag_nsnps"][k])
                    flag, nsnps = fn & 0xFFFF, fn >> 16
                    md = "MD:Z:" + "0C" * nsnps + str(ln - nsnps)
                    f.write(f"q{k}/1\t{flag}\t{names[sc]}\t{pos0 + 1}\t42\t{ln}M\t*\t0\t0\t{'A' * ln}\t{'I' * ln}\t{md}\n")
            sams.append(path)
        build = os.path.join(wd, "build")
        os.makedirs(build)
        subprocess.run([os.path.join(REF, "abawaca-build"), "-f", fa, "-o", build, "-s", os.path.join(wd, "sample*.sam"), "-c", sams[0]], check=True,
                       stderr=subprocess.DEVNULL, stdout=subprocess.DEVNULL)
        raw = os.path.join(wd, "raw.tsv")
        subprocess.run([os.path.join(REF, "ref_features"), fa, raw] + sams, check=True)
        rawcov = [[float(x) for x in line.rstrip("\n").split("\t")[3 + 180:]] for line in open(raw) if line.startswith("SEG")]
        fix = dict(args=COVERAGE_ORDER_ARGS, digests=dict(seq=sha(seq), offsets=sha(offsets), reads=[sha(r) for r in reads]),
                   lrn=open(os.path.join(build, "abawaca.lrn")).read(), info=open(os.path.join(build, "abawaca.info")).read(), rawcov=rawcov)
        with gzip.GzipFile(os.path.join(HERE, "kat_coverage_order.json.gz"), "wb", mtime=0) as g:
            g.write(json.dumps(fix).encode())
    print("coverage order done:", len(rawcov), "windows")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "coverage_order":
        make_coverage_order()
        sys.exit(0)
    if not os.path.exists(os.path.join(REF, "abawaca")):
        sys.exit("oracle/_ref is not built: run `make -C oracle ref` where /root/reference is mounted")
    make_kat()
    make_coverage_order()
    for name, args in SETS.items():
        make_synthetic(name, args)
