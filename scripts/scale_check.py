"""Size-independent checks at sizes the oracle cannot reach quickly (BASELINE.json configs[2]/[3] are 10-20x configs[1]):
    python scripts/scale_check.py [n_scaffolds] [n_samples] [n_genomes]
builds the features and bins a synthetic community on ONE GPU and verifies properties that hold at any size:
  * every feature value is a multiple of 0.001, k-mer columns in [0, 1], coverage >= 0;
  * (reported, not required: it depends on how separable the synthetic genomes are) whether every bin holds scaffolds of one synthetic genome;
  * idempotence: the largest bin, searched on its own, is terminal.
Prints one JSON line with sizes, timings and the outcome of the checks."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from abawaca_b200 import capi, pipeline, synth

n_scaf = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
n_samp = int(sys.argv[2]) if len(sys.argv) > 2 else 4
n_gen = int(sys.argv[3]) if len(sys.argv) > 3 else 64
t0 = time.perf_counter()
mg = synth.make_metagenome(n_scaf, n_samp, n_gen, synth.MASTER_SEED + 33, q6_reads=True)
t_gen = time.perf_counter() - t0
ctx = capi.Context(0)
out = {"scaffolds": n_scaf, "samples": n_samp, "genomes": n_gen, "bases": int(mg.seq.size), "reads": int(sum(r.size for r in mg.reads)), "generate_s": round(t_gen, 1)}
for it in range(2):
    t0 = time.perf_counter()
    fb = pipeline.build_features(ctx, mg.seq, mg.offsets, mg.reads, this_sample=0)
    ctx.synchronize()
    t1 = time.perf_counter()
    keep, dp2scaf, T, kept = pipeline.search_problem_from_counts(np.diff(fb.seg_first_host().astype(np.int64)))
    length = np.diff(mg.offsets.astype(np.int64)).astype(np.uint64)[kept]
    mask = mg.scg_masks()[kept] if it == 0 else mask
    row_of_dp = None if keep.all() else np.nonzero(keep)[0].astype(np.uint64)
    res = pipeline.search(ctx, fb.d_rows, dp2scaf, T, length, mask, layout=capi.LAYOUT_ROWMAJOR, values_on_device=True, nrows=fb.nseg, D=fb.ncols, ld=fb.ncols, row_of_dp=row_of_dp)
    t2 = time.perf_counter()
    if it == 1:
        out.update(windows=int(fb.nseg), dims=int(fb.ncols), features_ms=round(1000 * (t1 - t0), 1), search_ms=round(1000 * (t2 - t1), 1),
                   scaffolds_per_s_from_host_buffers=round(n_scaf / (t2 - t0), 1), clusters_evaluated=len(res.recs), levels=int(res.profile.levels))
        rows = fb.rows_host()
    fb.close()
bins = res.scaf2cluster
genome = mg.genome[kept]
ids = sorted(set(bins.tolist()) - {0})
pure = all(len(set(genome[bins == b].tolist())) == 1 for b in ids)
majority = sum(int(np.bincount(genome[bins == b]).max()) for b in ids)
out.update(bins=len(ids), unbinned_scaffolds=int((bins == 0).sum()), every_bin_one_genome=bool(pure), scaffolds_in_the_majority_genome_of_their_bin=round(majority / max(1, int((bins != 0).sum())), 4),
           values_are_milli=bool(np.array_equal(rows, np.round(rows * 1000) / 1000)), kmer_in_unit_interval=bool((rows[:, :179] >= 0).all() and (rows[:, :179] <= 1).all()),
           coverage_nonnegative=bool((rows[:, 179:] >= 0).all()))
b = int(np.bincount(bins)[1:].argmax()) + 1
sel_scaf = np.nonzero(bins == b)[0]
sel_dp = np.nonzero(np.isin(dp2scaf, sel_scaf))[0]
remap = np.full(T.size, -1, dtype=np.int64)
remap[sel_scaf] = np.arange(sel_scaf.size)
vals = np.ascontiguousarray(rows[keep][sel_dp].T)
sub = pipeline.search(ctx, vals, remap[dp2scaf[sel_dp]].astype(np.uint32), T[sel_scaf], length[sel_scaf], mask[sel_scaf])
out["largest_bin_is_terminal_on_its_own"] = bool(len(sub.recs) == 1 and sub.recs[0].split == 0)
print(json.dumps(out))
