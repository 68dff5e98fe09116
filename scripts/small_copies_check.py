"""Drop-in command lines with ABW_SMALL_COPIES=kernel against the unmodified reference binaries on a small synthetic set.

  python scripts/small_copies_check.py prepare DIR     # here (CPU): inputs + reference outputs (needs oracle/_ref)
  (on the GPU box)  cd DIR; ABW_SMALL_COPIES=kernel ../abawaca_b200/bin/abawaca-build -f assembly.fa -o OUT/build -s 'sample*.sam' -c sample00.sam
                    ABW_SCG_LIST=$PWD/scg.list ABW_SMALL_COPIES=kernel ../abawaca_b200/bin/abawaca -u OUT/build -o OUT/out -c genes.scg -p 8
  python scripts/small_copies_check.py compare DIR OUT # here again: byte comparison (cluster dumps as sets of rows: the reference writes them
                                                       # in the iteration order of an unordered_set, ClusterWriter.cpp:92-98)

Round 1: 150 scaffolds, 2 genomes, 2 samples, seed 777 -> every file identical (both binaries need about 4 s each, CUDA start-up included)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def prepare(wd):
    from abawaca_b200 import synth
    os.makedirs(wd, exist_ok=True)
    mg = synth.make_metagenome(150, 2, 2, 777, q6_reads=True)
    synth.write_reference_inputs(mg, wd)
    ref = os.path.join(ROOT, "oracle", "_ref")
    os.makedirs(os.path.join(wd, "ref_build"), exist_ok=True)
    os.makedirs(os.path.join(wd, "ref_out"), exist_ok=True)
    subprocess.run([os.path.join(ref, "abawaca-build"), "-f", "assembly.fa", "-o", "ref_build", "-s", "sample*.sam", "-c", "sample00.sam"], cwd=wd, check=True,
                   stdout=subprocess.DEVNULL)
    subprocess.run([os.path.join(ref, "abawaca"), "-u", "ref_build", "-o", "ref_out", "-c", "genes.scg", "-p", "8"], cwd=wd, check=True, stdout=subprocess.DEVNULL,
                   env=dict(os.environ, ABW_SCG_LIST=os.path.join(wd, "scg.list")))
    return 0


def compare(wd, out):
    rd = lambda p: open(p, "rb").read()                     # noqa: E731
    bad = []
    for f in ("abawaca.names", "abawaca.info"):
        if rd(os.path.join(wd, "ref_build", f)) != rd(os.path.join(out, "build", f)):
            bad.append(f)
    a, b = rd(os.path.join(wd, "ref_build", "abawaca.lrn")).split(b"\n"), rd(os.path.join(out, "build", "abawaca.lrn")).split(b"\n")
    if a[:3] + a[4:] != b[:3] + b[4:]:                      # line 4 names the SAM paths
        bad.append("abawaca.lrn")
    for f in ("scaf2cluster.txt", "summary.txt", "dp2cluster.txt"):
        if rd(os.path.join(wd, "ref_out", f)) != rd(os.path.join(out, "out", f)):
            bad.append(f)
    for f in sorted(os.listdir(os.path.join(wd, "ref_out", "clusters"))):
        x, y = rd(os.path.join(wd, "ref_out", "clusters", f)), rd(os.path.join(out, "out", "clusters", f))
        if (sorted(x.split(b"\n")) != sorted(y.split(b"\n"))) if f.endswith(".lrn") else (x != y):
            bad.append("clusters/" + f)
    for f in sorted(os.listdir(os.path.join(wd, "ref_out", "final-clusters"))):
        if rd(os.path.join(wd, "ref_out", "final-clusters", f)) != rd(os.path.join(out, "out", "final-clusters", f)):
            bad.append("final-clusters/" + f)
    print("identical" if not bad else "DIFFERENT: " + ", ".join(bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(prepare(sys.argv[2]) if sys.argv[1] == "prepare" else compare(sys.argv[2], sys.argv[3]))
