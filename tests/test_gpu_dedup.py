"""GPU: tsk_dedup_* (duplicate collapsing by UMI + the UMI sort) against the restatement of
tailseq-dedup-perfect, the reference binary's committed outputs and, where oracle/_ref is built,
the binary itself.  Both text outputs must be identical byte for byte."""
import os

import numpy as np
import pytest

from oracle import dedup_oracle
from tailseeker_b200 import dedup

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "dedup_perfect.npz")


def _run(data, sort_by_umi=False, input_order=False):
    cols = dedup.parse_taginfo(data)
    hi, lo = dedup.pack_umis(cols["umi"])
    d = dedup.Deduplicator(max(1, len(hi)))
    d.upload(hi, lo, cols["flags"], cols["polya"])
    d.run(sort_by_umi=sort_by_umi)
    rec, grp = d.records(), d.groups()
    res = dedup.format_outputs(cols, rec, grp, input_order=input_order, lost=d.lost())
    d.close()
    return res, rec, grp


@pytest.mark.parametrize("case", ["mixed", "sparse", "one_group"])
def test_golden(case):
    g = np.load(GOLD)
    (out, trace), rec, grp = _run(g[case + "_in"].tobytes())
    assert out == g[case + "_out"].tobytes()
    assert trace == g[case + "_trace"].tobytes()


@pytest.mark.parametrize("kw", [dict(n=20000, seed=21, n_umis=5000, big_groups=(3000, 257, 256, 255, 33, 32, 31)),
                                dict(n=30000, seed=22),
                                dict(n=1, seed=23), dict(n=2, seed=24, n_umis=1), dict(n=33, seed=25, n_umis=1),
                                dict(n=5000, seed=26, n_umis=2, umi_len=4),
                                dict(n=9000, seed=27, n_umis=40, umi_len=42)])
def test_matches_oracle(kw):
    data = dedup_oracle.synth_taginfo(**kw)
    (out, trace), rec, grp = _run(data)
    want_out, want_trace = dedup_oracle.dedup_perfect(data)
    assert out == want_out
    assert trace == want_trace
    n = len(rec["group"])
    assert sorted(rec["trace_slot"].tolist()) == list(range(n))
    assert int(grp["ndup"].sum()) == n


def test_many_epochs_in_one_group():
    """Priorities that keep rising (every record opens a new queue) and long equal runs."""
    bits = [0x10, 0x80, 0x02, 0x800, 0x200, 0x400, 0x40, 0x20]          # cleared one by one -> rising priority
    flags, polya = [], []
    f = sum(bits)
    rng = np.random.default_rng(3)
    for b in [0] + bits:
        f &= ~b
        for _ in range(int(rng.integers(1, 90))):
            flags.append(f | int(rng.integers(0, 2)) * 0x4)              # 0x4 does not change the priority
            polya.append(int(rng.integers(-1, 200)))
        flags.append(sum(bits))                                           # a low one in between: dropped at once
        polya.append(5)
    lines = [b"a1101\t%d\t%d\t%d\tT\tACGTACGT\n" % (i + 1, fl, pa) for i, (fl, pa) in enumerate(zip(flags, polya))]
    data = b"".join(lines) + b"a1101\t9999\t1\t7\t\tTTTTTTTT\n"
    (out, trace), _, _ = _run(data)
    assert (out, trace) == dedup_oracle.dedup_perfect(data)


@pytest.mark.skipif(not dedup_oracle.ref_available(), reason="oracle/_ref not built")
def test_matches_reference_binary(tmp_path):
    data = dedup_oracle.synth_taginfo(n=200000, seed=31, n_umis=60000, big_groups=(20000, 5000, 1200, 600))
    (out, trace), _, _ = _run(data)
    ref_out, ref_trace = dedup_oracle.run_reference(data, str(tmp_path))
    assert out == ref_out
    assert trace == ref_trace


def _uniform_groups(sizes, seed, lowfrac=0.1):
    rng = np.random.default_rng(seed)
    out, c = [], 0
    for gi, m in enumerate(sizes):
        umi = b"A" * (gi + 1)
        for _ in range(m):
            c += 1
            fl = 1 if rng.random() > lowfrac else 0x11
            out.append(b"a1101\t%d\t%d\t%d\tT\t%s\n" % (c, fl, int(rng.integers(0, 200)), umi))
    return b"".join(out)


@pytest.mark.parametrize("sizes", [(2046,), (2047,), (2048,), (2049,), (5000,), (3000, 2500, 100, 4000),
                                   (10000, 8000), (1500, 3071, 3072, 3073, 7000)])
def test_queue_growth_loses_the_arriving_record(sizes):
    """More than 2047 records of one priority in one group: the reference's queue grows and the
    record that arrives at the full queue continues as zeros (tqueue_expand(), :142-166)."""
    data = _uniform_groups(sizes, seed=sum(sizes))
    (out, trace), _, _ = _run(data)
    want_out, want_trace = dedup_oracle.dedup_perfect(data)
    assert out == want_out
    assert trace == want_trace
    if max(sizes) >= 2500:                      # 10 % of the records have a lower priority
        assert b"\n\t0\t" in trace


def test_sort_by_umi_replaces_the_external_sort():
    """Unsorted input (tiles concatenated in (tile, cluster) order) with sort_by_umi: the same
    groups, representatives and trace as the reference fed by `sort -k6,6 -k1,1 -k2,2n`, and the
    stdout lines in input order are the pipeline's second `sort -k1,1 -k2,2n`."""
    raw = dedup_oracle.synth_taginfo(n=40000, seed=41, n_umis=9000, big_groups=(2000, 300), sort=False)
    lines = raw.split(b"\n")[:-1]
    key = lambda ln: (ln.split(b"\t")[5], ln.split(b"\t")[0], int(ln.split(b"\t")[1]))
    sorted_data = b"\n".join(sorted(lines, key=key)) + b"\n"
    want_out, want_trace = dedup_oracle.dedup_perfect(sorted_data)
    (out, trace), rec, grp = _run(raw, sort_by_umi=True, input_order=True)
    assert trace == want_trace
    resorted = sorted(want_out.split(b"\n")[:-1], key=lambda ln: (ln.split(b"\t")[0], int(ln.split(b"\t")[1])))
    assert out == b"\n".join(resorted) + b"\n"
    # and the sorted input without sorting gives the same thing in group order
    (out2, trace2), _, _ = _run(sorted_data)
    assert out2 == want_out and trace2 == want_trace


def test_full_size_properties():
    """16 M records (one MiSeq run's worth of tags for one sample), keys made on the host:
    group count, group sizes and the order agree with NumPy; every record gets one trace line."""
    n = 16_000_000
    rng = np.random.default_rng(7)
    pool = rng.integers(0, 1 << 60, 9_000_000, dtype=np.uint64)
    pick = rng.integers(0, len(pool), n)
    hi = pool[pick] & np.uint64((1 << 63) - 1)
    lo = (pool[pick] * np.uint64(0x9E3779B97F4A7C15)) & np.uint64((1 << 63) - 1) & ~np.uint64(0xFFFFFF)
    flags = rng.integers(0, 4096, n).astype(np.int32)
    polya = rng.integers(-1, 231, n).astype(np.int32)
    d = dedup.Deduplicator(n)
    d.upload(hi, lo, flags, polya)
    ng = d.run(sort_by_umi=True)
    rec, grp = d.records(), d.groups()
    order = rec["order"].astype(np.int64)
    assert (np.sort(order) == np.arange(n)).all()
    kh, kl = hi[order], lo[order]
    nondecreasing = (kh[1:] > kh[:-1]) | ((kh[1:] == kh[:-1]) & (kl[1:] >= kl[:-1]))
    assert nondecreasing.all()
    same = (kh[1:] == kh[:-1]) & (kl[1:] == kl[:-1])
    assert (order[1:][same] > order[:-1][same]).all()                    # stable
    assert ng == int((~same).sum()) + 1
    assert (np.diff(rec["group"].astype(np.int64)) == (~same)).all()
    assert int(grp["ndup"].sum()) == n
    slots = np.sort(rec["trace_slot"])
    assert (slots == np.arange(n)).all()
    # representatives: of the highest priority in their group, nearest to the group's mean
    prio = np.array([dedup_oracle.priority(f) for f in range(4096)])[flags[order]]
    gmax = np.zeros(ng, np.int64)
    np.maximum.at(gmax, rec["group"], prio)
    assert (prio[grp["rep"]] == gmax).all()
    kept = np.bincount(rec["group"], weights=(prio == gmax[rec["group"]]), minlength=ng).astype(np.int64)
    assert (kept == grp["kept"]).all()
    assert (np.bincount(rec["group"], minlength=ng) == grp["ndup"]).all()
    dropped = rec["trace_value"] == -3
    assert (dropped == (prio < gmax[rec["group"]])).all()
    d.close()


def test_errors():
    from tailseeker_b200 import _lib
    d = dedup.Deduplicator(16)
    with pytest.raises(_lib.TskError):
        d.upload(np.zeros(17, np.uint64), np.zeros(17, np.uint64), np.zeros(17, np.int32), np.zeros(17, np.int32))
    d.upload(np.zeros(0, np.uint64), np.zeros(0, np.uint64), np.zeros(0, np.int32), np.zeros(0, np.int32))
    assert d.run() == 0
    d.close()


# ---- the drop-in program (C host + CUDA library) against the reference binary -----------------
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUR_DEDUP = os.path.join(ROOT, "tailseeker_b200", "bin", "tailseq-dedup-perfect")


@pytest.mark.skipif(not dedup_oracle.ref_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("kw", [dict(n=60000, seed=51, n_umis=15000, big_groups=(4000, 300, 64)),
                                dict(n=1, seed=52), dict(n=300, seed=53, n_umis=1)])
def test_program_matches_reference_binary(tmp_path, kw):
    assert os.access(OUR_DEDUP, os.X_OK), "host programs not built (__graft_entry__.build())"
    data = dedup_oracle.synth_taginfo(**kw)
    (tmp_path / "a").mkdir(); (tmp_path / "b").mkdir()
    ref = dedup_oracle.run_reference(data, str(tmp_path / "a"))
    our = dedup_oracle.run_reference(data, str(tmp_path / "b"), binary=OUR_DEDUP)
    assert our[0] == ref[0]
    assert our[1] == ref[1]


@pytest.mark.skipif(not dedup_oracle.ref_available(), reason="oracle/_ref not built")
def test_program_reproduces_lost_records(tmp_path):
    data = _uniform_groups((3000, 2500, 100, 4000), seed=9)
    (tmp_path / "a").mkdir(); (tmp_path / "b").mkdir()
    ref = dedup_oracle.run_reference(data, str(tmp_path / "a"))
    our = dedup_oracle.run_reference(data, str(tmp_path / "b"), binary=OUR_DEDUP)
    assert our == ref and b"\n\t0\t" in ref[1]


def test_program_sort_mode_and_edge_cases(tmp_path):
    import subprocess
    raw = dedup_oracle.synth_taginfo(n=30000, seed=61, n_umis=7000, big_groups=(900,), sort=False)
    lines = raw.split(b"\n")[:-1]
    key = lambda ln: (ln.split(b"\t")[5], ln.split(b"\t")[0], int(ln.split(b"\t")[1]))
    want_out, want_trace = dedup_oracle.dedup_perfect(b"\n".join(sorted(lines, key=key)) + b"\n")
    resorted = sorted(want_out.split(b"\n")[:-1], key=lambda ln: (ln.split(b"\t")[0], int(ln.split(b"\t")[1])))
    out, trace = dedup_oracle.run_reference(raw, str(tmp_path), binary=OUR_DEDUP, extra_args=("--sort",))
    assert out == b"\n".join(resorted) + b"\n"
    assert trace == want_trace
    # no trace argument, empty input, a line that does not parse (the rest before it is still processed)
    r = subprocess.run([OUR_DEDUP], input=b"a1101\t1\t1\t30\tTT\tACGT\na1101\t2\t1\t50\tG\tACGT\n", capture_output=True)
    assert r.returncode == 0 and r.stdout == b"a1101\t1\t1\t40\tTT\t2\n"
    r = subprocess.run([OUR_DEDUP, str(tmp_path / "e.gz")], input=b"", capture_output=True)
    assert r.returncode == 0 and r.stdout == b""
    import gzip
    assert gzip.open(str(tmp_path / "e.gz")).read() == b""
    r = subprocess.run([OUR_DEDUP], input=b"a1101\t1\t1\t30\tTT\tACGT\nbroken line\na1101\t9\t1\t5\t\tCCCC\n", capture_output=True)
    assert r.returncode == 0 and r.stdout == b"a1101\t1\t1\t30\tTT\t1\n" and b"Could not parse line 1" in r.stderr
