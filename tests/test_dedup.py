"""CPU: the restatement of tailseq-dedup-perfect (oracle/dedup_oracle.py) against the reference
binary's committed outputs (tests/golden/dedup_perfect.npz, made by tests/golden/make_golden.py dedup)
and, where oracle/_ref is built, against the binary itself; UMI keys; host-side parsing."""
import os

import numpy as np
import pytest

from oracle import dedup_oracle
from tailseeker_b200 import dedup

GOLD = os.path.join(os.path.dirname(__file__), "golden", "dedup_perfect.npz")


@pytest.mark.parametrize("case", ["mixed", "sparse", "one_group"])
def test_oracle_matches_golden(case):
    g = np.load(GOLD)
    out, trace = dedup_oracle.dedup_perfect(g[case + "_in"].tobytes())
    assert out == g[case + "_out"].tobytes()
    assert trace == g[case + "_trace"].tobytes()


@pytest.mark.skipif(not dedup_oracle.ref_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("kw", [dict(n=4000, seed=1, n_umis=900, big_groups=(500, 33, 32)),
                                dict(n=3000, seed=2), dict(n=1, seed=3), dict(n=2, seed=4, n_umis=1),
                                dict(n=800, seed=5, n_umis=3, umi_len=6)])
def test_oracle_matches_reference_binary(tmp_path, kw):
    data = dedup_oracle.synth_taginfo(**kw)
    assert dedup_oracle.run_reference(data, str(tmp_path)) == dedup_oracle.dedup_perfect(data)


@pytest.mark.skipif(not dedup_oracle.ref_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("sizes", [(2047,), (2048,), (2049,), (5000,), (3000, 2500, 100, 4000)])
def test_oracle_reproduces_the_queue_growth_quirk(tmp_path, sizes):
    """tqueue_expand() (:142-166) loses the record that arrives at a full queue."""
    rng = np.random.default_rng(sum(sizes))
    lines, c = [], 0
    for gi, m in enumerate(sizes):
        for _ in range(m):
            c += 1
            lines.append(b"a1101\t%d\t%d\t%d\tT\t%s\n" % (c, 1 if rng.random() > 0.1 else 0x11,
                                                            int(rng.integers(0, 200)), b"A" * (gi + 1)))
    data = b"".join(lines)
    assert dedup_oracle.run_reference(data, str(tmp_path)) == dedup_oracle.dedup_perfect(data)


def test_priority_table():
    # spot values of calculate_tag_prority() (:87-105)
    assert dedup_oracle.priority(0) == 1024 + 512 + 256 + 64 + 32 + 16 + 8 + 4
    assert dedup_oracle.priority(0x0009) == dedup_oracle.priority(0) + 128 + 2
    assert dedup_oracle.priority(0x0010) == dedup_oracle.priority(0) - 1024
    assert dedup_oracle.priority(0x1000 | 0x2000 | 0x0100 | 0x0004) == dedup_oracle.priority(0)


def test_umi_keys_order_like_strcmp():
    rng = np.random.default_rng(5)
    umis = []
    for _ in range(400):
        ln = int(rng.integers(0, 43))
        umis.append(bytes(rng.choice(np.frombuffer(b"ACGNT", np.uint8), ln)))
    umis += [b"", b"A", b"AA", b"T" * 42, b"A" * 21, b"A" * 22, b"A" * 21 + b"C"]
    hi, lo = dedup.pack_umis(umis)
    by_key = sorted(range(len(umis)), key=lambda i: (int(hi[i]), int(lo[i]), i))
    by_str = sorted(range(len(umis)), key=lambda i: (umis[i], i))
    assert by_key == by_str
    assert len({(int(a), int(b)) for a, b in zip(hi, lo)}) == len(set(umis))
    with pytest.raises(ValueError):
        dedup.pack_umis([b"A" * 43])
    with pytest.raises(ValueError):
        dedup.pack_umis([b"ACGU"])


def test_parse_taginfo_columns():
    data = b"a1101\t7\t9\t30\tTT\tACGT\nb2101\t12\t17\t-1\t\tNNNN\n"
    c = dedup.parse_taginfo(data)
    assert c["tile"] == [b"a1101", b"b2101"] and c["cluster"].tolist() == [7, 12]
    assert c["flags"].tolist() == [9, 17] and c["polya"].tolist() == [30, -1]
    assert c["mods"] == [b"TT", b""] and c["umi"] == [b"ACGT", b"NNNN"]
    assert c["head"] == [b"a1101\t7\t9\t30\tTT", b"b2101\t12\t17\t-1\t"]


def test_c_umi_packing_matches_the_python_mirror():
    """tsk_dedup_pack_umi() is a host function of the library: no GPU needed."""
    import ctypes
    from tailseeker_b200 import _lib
    L = _lib.load()
    rng = np.random.default_rng(9)
    umis = [bytes(rng.choice(np.frombuffer(b"ACGNT", np.uint8), int(rng.integers(0, 43)))) for _ in range(300)]
    hi, lo = dedup.pack_umis(umis)
    key = (ctypes.c_uint64 * 2)()
    for i, u in enumerate(umis):
        assert L.tsk_dedup_pack_umi(u, len(u), key) == 0
        assert (key[0], key[1]) == (int(hi[i]), int(lo[i])), u
    assert L.tsk_dedup_pack_umi(b"A" * 43, 43, key) == -1 and b"42" in L.tsk_last_error()
    assert L.tsk_dedup_pack_umi(b"ACGU", 4, key) == -1 and b"ACGNT" in L.tsk_last_error()


def test_dedup_has_no_cpu_path(tmp_path):
    """Without a B200 the library entry points and the drop-in program fail with a message."""
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from tailseeker_b200 import _lib
    with pytest.raises(_lib.TskError):
        dedup.Deduplicator(16)
    prog = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tailseeker_b200", "bin",
                        "tailseq-dedup-perfect")
    if os.access(prog, os.X_OK):
        r = subprocess.run([prog], input=b"a1101\t1\t1\t30\tTT\tACGT\n", capture_output=True)
        assert r.returncode != 0 and r.stdout == b"" and b"no CPU path" in r.stderr
