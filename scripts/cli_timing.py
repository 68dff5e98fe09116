"""Where the time of the drop-in command lines goes: python scripts/cli_timing.py [per_genome] [genomes]  (a sample of the configs[1] workload as text files)"""
import os, subprocess, sys, tempfile, time, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from abawaca_b200 import synth
per_genome = int(sys.argv[1]) if len(sys.argv) > 1 else 400
genomes = int(sys.argv[2]) if len(sys.argv) > 2 else 8
mg = synth.make_metagenome(**synth.CONFIGS["cfg2"], q6_reads=True)
sample = bench.reference_sample(mg, per_genome, genomes)
wd = tempfile.mkdtemp(prefix="abw_cli_")
paths = synth.write_reference_inputs(sample, wd)
mine = os.path.join(bench.ROOT, "abawaca_b200", "bin")
env = dict(os.environ, ABW_SCG_LIST=paths["scg_list"], ABW_TIMING="1")
for rep in range(2):
    build, out = os.path.join(wd, "b"), os.path.join(wd, "o")
    shutil.rmtree(build, ignore_errors=True); shutil.rmtree(out, ignore_errors=True); os.makedirs(build)
    t0 = time.perf_counter()
    r1 = subprocess.run([os.path.join(mine, "abawaca-build"), "-f", paths["fasta"], "-o", build, "-s", os.path.join(wd, "sample*.sam"), "-c", paths["sams"][0]], env=env, capture_output=True, text=True)
    t1 = time.perf_counter()
    r2 = subprocess.run([os.path.join(mine, "abawaca"), "-u", build, "-o", out, "-c", paths["gene2scg"], "-p", "16"], env=env, capture_output=True, text=True)
    t2 = time.perf_counter()
    print(f"rep {rep}: {sample.nscaf} scaffolds  abawaca-build {t1 - t0:.3f} s  abawaca {t2 - t1:.3f} s")
    for l in (r1.stderr + r2.stderr).splitlines():
        if "[abw timing]" in l:
            print("   ", l)
shutil.rmtree(wd, ignore_errors=True)
