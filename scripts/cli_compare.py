"""T3 of SURVEY.md section 8d: wall time of the full command lines (text in, text out) of the reference binaries (oracle/_ref, -O2)
and of the drop-in programs (abawaca_b200/bin) on the same bounded sample of configs[1]; checks that the bins are identical.

  python scripts/cli_compare.py [scaffolds_per_genome] [genomes]
"""
import json, os, shutil, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from abawaca_b200 import synth


def run(cmd, env=None):
    t0 = time.perf_counter()
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=env)
    return time.perf_counter() - t0


def main():
    per_genome = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    genomes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    mg = synth.make_metagenome(**synth.CONFIGS["cfg2"], q6_reads=True)
    sample = bench.reference_sample(mg, per_genome=per_genome, genomes=genomes)
    wd = tempfile.mkdtemp(prefix="abw_cli_")
    try:
        paths = synth.write_reference_inputs(sample, wd)
        env = dict(os.environ, ABW_SCG_LIST=paths["scg_list"])
        ncpu = max(1, min(40, os.cpu_count() or 1))
        out = {}
        for tag, bindir in (("reference", os.path.join(ROOT, "oracle", "_ref")), ("b200", os.path.join(ROOT, "abawaca_b200", "bin"))):
            b, o = os.path.join(wd, "build_" + tag), os.path.join(wd, "out_" + tag)
            os.makedirs(b)
            tb = run([os.path.join(bindir, "abawaca-build"), "-f", paths["fasta"], "-o", b, "-s", os.path.join(wd, "sample*.sam"), "-c", paths["sams"][0]])
            ts = run([os.path.join(bindir, "abawaca"), "-u", b, "-o", o, "-c", paths["gene2scg"], "-p", str(ncpu)], env=env)
            out[tag] = dict(build_s=round(tb, 3), bin_s=round(ts, 3), scaffolds_per_s=round(sample.nscaf / (tb + ts), 1),
                            scaf2cluster=open(os.path.join(o, "scaf2cluster.txt")).read(), lrn=open(os.path.join(b, "abawaca.lrn")).read().split("\n", 4)[4])
        same = out["reference"]["scaf2cluster"] == out["b200"]["scaf2cluster"] and out["reference"]["lrn"] == out["b200"]["lrn"]
        for t in out.values():
            del t["scaf2cluster"], t["lrn"]
        print(json.dumps(dict(sample=f"{sample.nscaf} scaffolds, {sum(r.size for r in sample.reads)} reads in {len(sample.reads)} SAM files, {os.cpu_count()} host cores",
                              identical_lrn_and_bins=same, **out)))
    finally:
        shutil.rmtree(wd, ignore_errors=True)


main()
