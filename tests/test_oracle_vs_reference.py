"""CPU tests: the flat-array oracle (oracle/abw_oracle.c) against the committed golden vectors that were
produced by the UNMODIFIED reference (tests/golden/make_golden.py).  This is what pins the oracle."""
import gzip
import json
import os

import numpy as np
import pytest

from golden_util import GOLDEN, load_set, parse_lrn_text, parse_ref_search, search_problem, compare_cluster_records


def _kat():
    return json.loads(gzip.open(os.path.join(GOLDEN, "kat_features.json.gz"), "rb").read())


def _kat_arrays(kat):
    """Scaffolds in byte-wise name order + read records per sample, as the flat API wants them."""
    seqs = sorted(kat["seqs"], key=lambda x: x[0].encode())
    names = [n for n, _ in seqs]
    idx = {n: i for i, n in enumerate(names)}
    offsets = np.zeros(len(seqs) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(s) for _, s in seqs])
    seq = np.frombuffer("".join(s for _, s in seqs).encode(), dtype=np.uint8)
    from oracle import abwo
    reads = []
    for rr in kat["reads"]:
        rec = np.zeros(len(rr), dtype=abwo.READ_DTYPE)
        for i, (rname, pos1, ln, flag, nsnps) in enumerate(rr):
            rec[i] = (idx.get(rname, 0xFFFFFFFF), (pos1 - 1) & 0xFFFFFFFF, ln, (flag & 0xFFFF) | (nsnps << 16))
        reads.append(rec)
    return names, seq, offsets, reads


def test_kat_features_match_reference(oracle):
    kat = _kat()
    names, seq, offsets, reads = _kat_arrays(kat)
    f = oracle.build_features(seq, offsets, reads, this_sample=0, want_raw=True)
    heads, vals = parse_lrn_text(kat["lrn"])
    assert heads[:179] == [n for n in oracle.dim_names() if n != "A"]
    assert vals.shape == f["rows"].shape
    assert np.array_equal(vals, f["rows"])
    # un-truncated doubles (ref_features dump): k-mer columns and coverage
    raw = np.array([[float(x) for x in l.split("\t")[3:]] for l in kat["raw"].splitlines() if l.startswith("SEG")])
    assert np.array_equal(raw, f["raw"])
    # .names: coordinates and non-N counts
    recs = [l.split("\t") for l in kat["names"].splitlines()[1:]]
    k = 0
    per = {}
    for i, r in enumerate(recs):
        s = int(f["seg_scaf"][i])
        per[s] = per.get(s, 0) + 1
        ln = int(f["seg_end"][i] - f["seg_start"][i] + 1)
        assert r[1] == f"{names[s]}_{per[s]}"
        assert r[2] == f"{names[s]}:({int(f['seg_start'][i])}, {int(f['seg_end'][i])}), {int(f['seg_nonN'][i])}/{ln} non-Ns bps"
    # .info
    for i, l in enumerate(kat["info"].splitlines()):
        nm, ln, cvg, gc, Ns = l.split("\t")
        assert nm == names[i] and int(ln) == int(offsets[i + 1] - offsets[i])
        assert "%.3f" % f["info_cvg"][i] == cvg and "%.3f" % f["info_gc"][i] == gc and int(Ns) == f["info_Ns"][i]


def test_kat_survey_known_answers(oracle):
    """SURVEY.md section 8c table (recorded from the compiled reference)."""
    def segs(s):
        st, en = oracle.segment(np.frombuffer(s.encode(), dtype=np.uint8))
        return list(zip(st.tolist(), en.tolist()))
    assert segs("ACGT" * 600) == [(1, 2400)]
    assert segs("AACCGGTT" * 550) == [(1, 2200), (2201, 4400)]
    assert segs("ACGTN" * 1000) == [(1, 2499), (2500, 4999)]
    assert segs("A" * 4001) == [(1, 2000), (2001, 4000)]
    up, bad = oracle.validate_upper(np.frombuffer(("acgtR" * 900).encode(), dtype=np.uint8))
    assert bad == -1 and segs(up.tobytes().decode()) == [(1, 2250), (2251, 4500)]
    assert oracle.validate_upper(np.frombuffer(b"ACGnT", dtype=np.uint8))[1] == 3
    names = oracle.dim_names()
    f, _, _ = oracle.kmer_features(np.frombuffer(("ACGT" * 600).encode(), dtype=np.uint8))
    t = {n: oracle.lib().abwo_trunc3(float(v)) for n, v in zip(names, f)}
    assert (t["C"], t["AA"], t["AC"]) == (0.5, 0.0, 0.5)
    f, _, _ = oracle.kmer_features(np.frombuffer(("ACGTN" * 1000)[:2499].encode(), dtype=np.uint8))
    assert oracle.lib().abwo_trunc3(float(f[names.index("AC")])) == 0.666
    f, _, _ = oracle.kmer_features(np.frombuffer(b"A" * 2000, dtype=np.uint8))
    assert f[names.index("AA")] == 1.0
    assert len(names) == 180 and names[:12] == ["A", "C", "AA", "AC", "AG", "AT", "CA", "CC", "CG", "GA", "GC", "TA"] and names[-1] == "TTAA"


@pytest.mark.parametrize("name", ["tiny_clean", "tiny_noisy", "cfg1"])
def test_features_match_reference_files(oracle, name):
    g = load_set(name)
    mg = g["mg"]
    f = oracle.build_features(mg.seq, mg.offsets, mg.reads, this_sample=0, want_raw=True)
    heads, vals = parse_lrn_text(g["lrn_text"])
    assert np.array_equal(vals, f["rows"])
    assert np.array_equal(g["rawcov"], f["raw"][:, 180:])
    info = [l.split("\t") for l in g["info_text"].splitlines()]
    assert [x[0] for x in info] == mg.names
    assert ["%.3f" % v for v in f["info_cvg"]] == [x[2] for x in info]
    assert ["%.3f" % v for v in f["info_gc"]] == [x[3] for x in info]
    assert [int(x[4]) for x in info] == f["info_Ns"].tolist()


@pytest.mark.parametrize("name", ["tiny_clean", "tiny_noisy", "cfg1"])
@pytest.mark.parametrize("strategy", [0, 1])
def test_search_matches_reference(oracle, name, strategy):
    g = load_set(name)
    prob = search_problem(name)
    S = oracle.Search(prob["values"], prob["dp2scaf"], prob["T"], prob["len"], prob["scgmask"])
    recs, dp2c, s2c = S.run(strategy=strategy)
    ref_clusters, ref_bins = parse_ref_search(g["meta"]["ref_search"]["sensspec" if strategy == 0 else "splitscafs"])
    assert compare_cluster_records(ref_clusters, recs, strategy) == []
    assert [b for _, b in ref_bins] == s2c.tolist()
    assert [n for n, _ in ref_bins] == prob["scaf_names"]
    if strategy == 0:
        # the real `abawaca` binary's scaf2cluster.txt (abawaca.cpp:205-210)
        lines = [l.split("\t") for l in g["meta"]["scaf2cluster"].splitlines()]
        assert [int(x[1]) for x in lines] == s2c.tolist()


def test_search_threads_do_not_change_result(oracle):
    prob = search_problem("tiny_noisy")
    S = oracle.Search(prob["values"], prob["dp2scaf"], prob["T"], prob["len"], prob["scgmask"])
    r1, _, s1 = S.run(nthreads=1)
    r8, _, s8 = S.run(nthreads=8)
    assert s1.tolist() == s8.tolist()
    assert [(r.best.dim, r.best.value, r.best.a, r.best.b) for r in r1] == [(r.best.dim, r.best.value, r.best.a, r.best.b) for r in r8]


def test_three_decimal_coverage_from_the_integer_sum(oracle):
    """The shortcut abw_coverage takes for three-decimal output (k_cov_quotient), restated in numpy: wherever it claims
    that the truncation cannot depend on the order of the reads, floor(1000 * A / len) / 1000 must be the sequential value of the oracle; and
    on this workload (round window lengths) the windows it leaves to the in-order path must include some where the naive quotient is wrong."""
    from golden_util import coverage_edge_workload
    seq, offsets, reads = coverage_edge_workload()
    f = oracle.build_features(seq, offsets, reads, this_sample=0)
    seg_scaf, seg_start, seg_end = f["seg_scaf"].astype(np.int64), f["seg_start"].astype(np.int64), f["seg_end"].astype(np.int64)
    nscaf = offsets.size - 1
    first = np.searchsorted(seg_scaf, np.arange(nscaf + 1))
    claimed = wrong_if_naive = undecided = 0
    for j, r in enumerate(reads):
        flag, nsnps = r["flag_nsnps"] & 0xFFFF, r["flag_nsnps"] >> 16
        ok = ~(((flag & 0x4) != 0) | (nsnps > 15) | ((flag & 0x100) != 0)) & (r["scaf"] < nscaf)
        A = np.zeros(seg_scaf.size, dtype=np.int64)
        for sc, s, ln in zip(r["scaf"][ok].astype(np.int64), r["pos0"][ok].astype(np.int64), r["len"][ok].astype(np.int64)):
            e = s + ln - 1
            for g in range(first[sc], first[sc + 1]):
                if s > seg_end[g]:
                    continue
                if e < seg_start[g]:
                    break
                A[g] += min(e, seg_end[g]) - max(s, seg_start[g]) + 1
        ln = seg_end - seg_start + 1
        m, rem = np.divmod(1000 * A, ln)
        bound = 2000.0 * (A + 2) * A * 2.0 ** -53
        safe = (A == 0) | ((rem != 0) & (bound < np.minimum(rem, ln - rem)))
        got = f["rows"][:, 179 + j]
        assert np.array_equal((m / 1000.0)[safe], got[safe])
        claimed += int(safe.sum())
        undecided += int((~safe).sum())
        wrong_if_naive += int(((m / 1000.0) != got)[~safe].sum())
    assert claimed > 0 and undecided > 100 and wrong_if_naive > 0


def _coverage_order_fixture():
    import hashlib
    from golden_util import coverage_edge_workload
    fix = json.loads(gzip.open(os.path.join(GOLDEN, "kat_coverage_order.json.gz"), "rb").read())
    seq, offsets, reads = coverage_edge_workload(**fix["args"])
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()   # noqa: E731
    assert fix["digests"] == dict(seq=sha(seq), offsets=sha(offsets), reads=[sha(r) for r in reads]), "generator drift: regenerate with make_golden.py coverage_order"
    heads, vals = parse_lrn_text(fix["lrn"])
    return fix, seq, offsets, reads, vals


def test_coverage_order_windows_match_reference(oracle):
    """Windows of round lengths, produced by the unmodified reference: the third decimal depends on the order in which the reference adds the
    reads (quirk Q5).  Pins the oracle's sequential sum on exactly the windows where abw_coverage's integer shortcut does not apply."""
    fix, seq, offsets, reads, vals = _coverage_order_fixture()
    f = oracle.build_features(seq, offsets, reads, this_sample=0, want_raw=True)
    assert f["rows"].shape == vals.shape and np.array_equal(f["rows"], vals)
    rawcov = np.array(fix["rawcov"], dtype=np.float64)
    assert np.array_equal(f["raw"][:, 180:], rawcov)
    for i, l in enumerate(fix["info"].splitlines()):
        assert "%.3f" % f["info_cvg"][i] == l.split("\t")[2]
    # the fixture does contain windows whose exact value is a multiple of 0.001, and some where truncating the exact quotient would be wrong
    ln = (f["seg_end"] - f["seg_start"] + 1).astype(np.int64)
    undecided = wrong = 0
    for j in range(rawcov.shape[1]):
        A = np.rint(rawcov[:, j] * ln).astype(np.int64)
        m, rem = np.divmod(1000 * A, ln)
        undecided += int(((rem == 0) & (A > 0)).sum())
        wrong += int(((m / 1000.0) != vals[:, 179 + j]).sum())
    assert undecided > 50 and wrong > 0
