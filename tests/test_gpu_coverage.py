"""GPU parity of the coverage stage (abw_coverage / abw_coverage_batch, csrc/coverage.cu) against the flat-array oracle, which adds fl(overlap / length)
in SAM order like Scaf_segment::add_mapped_read (abawaca-build.cpp:177-185): both record formats, one call per sample against one call for all samples,
reads in scaffold order / coordinate order / shuffled, and a randomized property test of the integer-sum shortcut (k_cov_quotient) over window lengths
from 1 to 10^6 -- the one place where bit-exactness rests on an inequality instead of on doing the same operations."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from abawaca_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def _variants(ctx, mg_seq, offsets, reads, this_sample, params=None):
    """rows and .info coverage through: one batched call (full records), one call per sample, compact host-filtered records"""
    from abawaca_b200 import pipeline
    nscaf = offsets.size - 1
    lengths = np.diff(offsets.astype(np.int64))
    out = {}
    for name, kw, rd in (("batch", {}, reads), ("per_sample", dict(per_sample_calls=True), reads),
                         ("compact", {}, [pipeline.compact_reads(r, nscaf, (params.max_snps if params else 15)) for r in reads])):
        fb = pipeline.build_features(ctx, mg_seq, offsets, rd, this_sample=this_sample, params=params, **kw)
        out[name] = (fb.rows_host(), fb.scaffold_stats_host(lengths)["cvg"])
        fb.close()
    return out


@pytest.mark.parametrize("order", ["scaffold", "coordinate", "shuffled"])
def test_record_formats_and_batching_match_the_oracle(ctx, oracle, order):
    from abawaca_b200 import synth
    mg = synth.make_metagenome(1500, 3, 5, 77, shuffle_reads=(order == "shuffled"), n_run_frac=0.05)
    rng = np.random.default_rng(5)
    reads = [r.copy() for r in mg.reads]
    if order == "coordinate":                               # what samtools sort leaves: by scaffold, then by position
        reads = [r[np.lexsort((r["pos0"], r["scaf"]))] for r in reads]
    reads[1]["len"] = rng.integers(30, 400, reads[1].size)  # trimmed reads: per-read lengths (compact format with len16)
    reads[2]["scaf"][::97] = 0xFFFFFFFF                     # reference name not in the assembly
    f = oracle.build_features(mg.seq, mg.offsets, reads, this_sample=1)
    for name, (rows, cvg) in _variants(ctx, mg.seq, mg.offsets, reads, 1).items():
        assert np.array_equal(rows, f["rows"]), name
        assert np.array_equal(cvg, f["info_cvg"]), name


def test_compact_records_with_one_length_use_eight_bytes_per_read():
    from abawaca_b200 import capi, pipeline, synth
    mg = synth.make_metagenome(200, 2, 2, 3)
    c = pipeline.compact_reads(mg.reads[0], mg.nscaf)
    flag, nsnps = mg.reads[0]["flag_nsnps"] & 0xFFFF, mg.reads[0]["flag_nsnps"] >> 16
    assert c.fmt == capi.READS_COMPACT and c.len16 is None and c.length == 150 and c.nbytes == 8 * c.n
    assert c.n == int((((flag & 0x104) == 0) & (nsnps <= 15)).sum())


@pytest.mark.parametrize("window,seed", [(1, 1), (7, 2), (1000, 3), (2000, 4), (2048, 5), (3125, 6), (65536, 7), (250000, 8), (1000000, 9)])
def test_integer_sum_shortcut_on_random_window_lengths(ctx, oracle, window, seed):
    """Random scaffolds cut with window sizes from 1 to 10^6 (window lengths then lie in [window, 2 window)), random read multisets with random lengths
    in random order, deep enough that sums reach 10^7: every value must equal the oracle's in-order sum truncated to three decimals."""
    from abawaca_b200 import capi, pipeline
    rng = np.random.default_rng(1000 + seed)
    nscaf = 40 if window >= 65536 else 120
    lens = np.maximum(1, (window * rng.uniform(0.4, 6.0, nscaf)).astype(np.int64))
    lens = np.minimum(lens, 4_000_000)
    if window <= 7:
        lens = rng.integers(1, 400, nscaf)
    offsets = np.zeros(nscaf + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lens)
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(offsets[-1]))].copy()
    for i in range(0, nscaf, 9):                            # runs of N move the window boundaries
        L = int(lens[i])
        if L > 20:
            p = int(rng.integers(0, L - 10))
            seq[int(offsets[i]) + p:int(offsets[i]) + p + int(rng.integers(1, max(2, L // 10)))] = ord("N")
    reads = []
    for j in range(2):
        recs = []
        for i in range(nscaf):
            L = int(lens[i])
            k = int(rng.integers(0, 60)) if j == 0 else int(rng.integers(0, max(2, min(4000, 40 * L // 100))))
            r = np.zeros(k, dtype=capi.READ_DTYPE)
            r["scaf"] = i
            r["pos0"] = rng.integers(0, L + 3, k)
            r["len"] = rng.integers(1, max(2, min(2 * window + 50, 60000)), k) if j == 0 else rng.choice([100, 150, 250], k)
            r["flag_nsnps"] = np.where(rng.random(k) < 0.03, 0x100, 0) | (rng.integers(0, 18, k).astype(np.uint32) << 16)
            recs.append(r)
        r = np.concatenate(recs)
        rng.shuffle(r)
        reads.append(r)
    p, po = capi.default_params(), oracle.default_params()
    p.window_size = po.window_size = window
    f = oracle.build_features(seq, offsets, reads, this_sample=0, params=po)
    for name, (rows, cvg) in _variants(ctx, seq, offsets, reads, 0, params=p).items():
        assert np.array_equal(rows[:, 179:], f["rows"][:, 179:]), (name, window)
        assert np.array_equal(cvg, f["info_cvg"]), name


def test_raw_coverage_in_both_formats(ctx, oracle):
    from abawaca_b200 import capi, pipeline, synth
    mg = synth.make_metagenome(400, 2, 3, 12, shuffle_reads=True, n_run_frac=0.1)
    f = oracle.build_features(mg.seq, mg.offsets, mg.reads, this_sample=0, want_raw=True)
    for rd in (mg.reads, [pipeline.compact_reads(r, mg.nscaf) for r in mg.reads]):
        fb = pipeline.build_features(ctx, mg.seq, mg.offsets, rd, this_sample=0, kind=capi.FEAT_RAW, skip_A=False)
        assert np.array_equal(fb.rows_host(), f["raw"])
        fb.close()


def test_empty_samples_and_empty_assembly(ctx, oracle):
    from abawaca_b200 import capi, pipeline, synth
    mg = synth.make_metagenome(50, 2, 2, 8)
    reads = [mg.reads[0], np.zeros(0, dtype=capi.READ_DTYPE)]
    f = oracle.build_features(mg.seq, mg.offsets, reads, this_sample=1)
    for name, (rows, cvg) in _variants(ctx, mg.seq, mg.offsets, reads, 1).items():
        assert np.array_equal(rows, f["rows"]) and np.array_equal(cvg, f["info_cvg"]), name


def test_a_ticket_survives_a_synchronize(ctx):
    """abw_h2d_async tickets are monotonic: a ticket retired by abw_ctx_synchronize means "already complete" (it used to be rejected, and could
    alias a later copy's event)."""
    import torch
    a = torch.arange(1 << 16, dtype=torch.int32).pin_memory().numpy()
    d1, d2 = ctx.alloc(a.nbytes), ctx.alloc(a.nbytes)
    t1 = ctx.h2d_async(d1, a)
    ctx.synchronize()
    ctx.wait_h2d(t1)                                        # retired, not unknown
    t2 = ctx.h2d_async(d2, a)
    assert t2 > t1
    ctx.wait_h2d(t1)
    ctx.wait_h2d(t2)
    back = np.zeros_like(a)
    ctx.to_host(back, d2)
    assert np.array_equal(back, a)
    with pytest.raises(Exception):
        ctx.wait_h2d(t2 + 5)
    ctx.free(d1); ctx.free(d2)


def test_rows_as_integer_thousandths(ctx, oracle):
    from abawaca_b200 import pipeline, synth
    mg = synth.make_metagenome(300, 3, 3, 21)
    fb = pipeline.build_features(ctx, mg.seq, mg.offsets, mg.reads, this_sample=0)
    rows = fb.rows_host()
    k16, k32, bad = fb.rows_milli_host()
    assert bad == 0
    assert k16.dtype == np.uint16 and k16.shape == (fb.nseg, 179) and k32.dtype == np.uint32 and k32.shape == (fb.nseg, 3)
    assert np.array_equal(k16 / 1000.0, rows[:, :179]) and np.array_equal(k32 / 1000.0, rows[:, 179:])
    fb.close()
