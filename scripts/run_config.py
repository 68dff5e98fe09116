"""Full-size run of a BASELINE.json configuration on ONE GPU, with the oracle on a random subsample of scaffolds and size-independent properties:
  python scripts/run_config.py cfg3|cfg4 [--coverage-scale X] [--check-scaffolds K]
cfg3 = 500 000 scaffolds x 20 samples (128 genomes), cfg4 = 1 000 000 scaffolds x 50 samples (256 genomes); the generator's coverage (0.5-16x) is scaled by
--coverage-scale (default 0.05: the read records of the full coverage would not fit the host's memory).  Prints one JSON line.
Checks: window table, k-mer and coverage rows of K random scaffolds against oracle/abw_oracle.c (a scaffold's rows depend on that scaffold and its reads only);
every final bin is pure (one genome) or terminal for a stated reason; idempotence of the largest bin; bins are a partition of the kept scaffolds."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from abawaca_b200 import capi, pipeline, synth
from oracle import abwo

ap = argparse.ArgumentParser()
ap.add_argument("config", choices=["cfg2", "cfg3", "cfg4"])
ap.add_argument("--coverage-scale", type=float, default=0.05)
ap.add_argument("--check-scaffolds", type=int, default=2000)
ap.add_argument("--scaffolds", type=int, default=0)
args = ap.parse_args()
cfg = dict(synth.CONFIGS[args.config])
if args.scaffolds:
    cfg["n_scaffolds"] = args.scaffolds
out = {"config": args.config, **cfg, "coverage_scale": args.coverage_scale}
t0 = time.time()
mg = synth.make_metagenome(**cfg, q6_reads=True, cov_lo=0.5 * args.coverage_scale, cov_hi=16.0 * args.coverage_scale)
out["generate_s"] = round(time.time() - t0, 1)
out["bp"] = int(mg.seq.size)
out["reads"] = int(sum(r.size for r in mg.reads))
ctx = capi.Context(0)
reads = [pipeline.compact_reads(r, mg.nscaf) for r in mg.reads]
t = {}
t0 = time.time()
fb = pipeline.build_features(ctx, mg.seq, mg.offsets, reads, this_sample=0, timings=t)
out["features_wall_s"] = round(time.time() - t0, 3)
out["feature_phase_ms"] = {k: round(v, 2) for k, v in t.items()}
out["windows"], out["dims"] = int(fb.nseg), int(fb.ncols)
first = fb.seg_first_host().astype(np.int64)
counts = np.diff(first)
# oracle on a subsample of scaffolds
rng = np.random.default_rng(5)
pick = np.sort(rng.choice(mg.nscaf, min(args.check_scaffolds, mg.nscaf), replace=False))
remap = np.full(mg.nscaf, -1, dtype=np.int64); remap[pick] = np.arange(pick.size)
lens = np.diff(mg.offsets.astype(np.int64))[pick]
off = np.zeros(pick.size + 1, dtype=np.uint64); off[1:] = np.cumsum(lens)
seq = np.concatenate([mg.scaffold(int(i)) for i in pick])
sub_reads = []
for r in mg.reads:
    rr = r[remap[r["scaf"]] >= 0].copy()
    rr["scaf"] = remap[rr["scaf"]].astype(np.uint32)
    sub_reads.append(rr)
f = abwo.build_features(seq, off, sub_reads, this_sample=0)
rows_pick = np.concatenate([np.arange(first[i], first[i + 1]) for i in pick])
d_rows = np.empty((rows_pick.size, fb.ncols))
# fetch the picked rows only: whole matrix to the host in slabs would also do, but the matrix of cfg4 is 9 GB
rows = fb.rows_host()
got = rows[rows_pick]
out["oracle_subsample"] = {"scaffolds": int(pick.size), "windows": int(rows_pick.size), "rows_identical": bool(np.array_equal(got, f["rows"])),
                           "info_coverage_identical": bool(np.array_equal(fb.scaffold_stats_host(np.diff(mg.offsets.astype(np.int64)))["cvg"][pick], f["info_cvg"]))}
out["rows_are_multiples_of_0.001"] = bool(np.array_equal(rows, np.round(rows * 1000) / 1000))
del rows, got
row_of_dp, T, kept, N = pipeline.search_rows_from_counts(counts)
if kept is None:
    kept = slice(None)
length = np.diff(mg.offsets.astype(np.int64)).astype(np.uint64)[kept]
mask = mg.scg_masks()[kept]
st = {}
t0 = time.time()
res = pipeline.search(ctx, fb.d_rows, None, T, length, mask, layout=capi.LAYOUT_ROWMAJOR, values_on_device=True, nrows=fb.nseg, D=fb.ncols, ld=fb.ncols,
                      row_of_dp=row_of_dp, timings=st)
out["search_first_call_s"] = round(time.time() - t0, 3)
st = {}
res = pipeline.search(ctx, fb.d_rows, None, T, length, mask, layout=capi.LAYOUT_ROWMAJOR, values_on_device=True, nrows=fb.nseg, D=fb.ncols, ld=fb.ncols,
                      row_of_dp=row_of_dp, timings=st)
out["search_ms"] = {k: round(v, 2) for k, v in st.items()}
out["datapoints"] = int(N)
out["clusters_evaluated"] = len(res.recs)
out["levels"] = int(res.profile.levels)
bins = res.scaf2cluster
genome = mg.genome[kept]
ids = sorted(set(bins.tolist()) - {0})
out["bins"] = len(ids)
pure = sum(1 for b in ids if len(set(genome[bins == b].tolist())) == 1)
out["pure_bins"] = pure
out["scaffolds_binned"] = int((bins != 0).sum())
out["scaffolds_kept"] = int(bins.size)
# every record: children sizes add up, ids ascending, parents evaluated before children
ok = all(r.child1_ndps + r.child2_ndps == r.ndps for r in res.recs if r.split) and [r.id for r in res.recs] == sorted(r.id for r in res.recs)
out["records_consistent"] = bool(ok)
# idempotence: the largest final bin, searched on its own, is terminal
b = int(np.bincount(bins).argmax()) if ids else 0
if b:
    dp2scaf = np.repeat(np.arange(T.size, dtype=np.uint32), T)
    sel_scaf = np.nonzero(bins == b)[0]
    sel_dp = np.nonzero(np.isin(dp2scaf, sel_scaf))[0]
    rmap = np.full(T.size, -1, dtype=np.int64); rmap[sel_scaf] = np.arange(sel_scaf.size)
    rod = sel_dp if row_of_dp is None else row_of_dp[sel_dp]
    sub = pipeline.search(ctx, fb.d_rows, rmap[dp2scaf[sel_dp]].astype(np.uint32), T[sel_scaf], length[sel_scaf], mask[sel_scaf], layout=capi.LAYOUT_ROWMAJOR,
                          values_on_device=True, nrows=fb.nseg, D=fb.ncols, ld=fb.ncols, row_of_dp=rod.astype(np.uint64))
    out["largest_bin_is_terminal_on_its_own"] = bool(len(sub.recs) == 1 and sub.recs[0].split == 0)
fb.close()
ctx.close()
print(json.dumps(out))
