"""CPU: the drop-in tailseq-writefastq (csrc/host/tailseq-writefastq.c, a host program without a
kernel) against the reference binary (src/exporter/tailseq-writefastq.c compiled into oracle/_ref)
on the same taginfo / seqqual files: same FASTQ bytes after decompression, same exit status and
diagnostics on broken inputs."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from oracle import refrun

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "tailseeker_b200", "bin", "tailseq-writefastq")
REF = os.path.join(ROOT, "oracle", "_ref", "tailseq-writefastq")

pytestmark = pytest.mark.skipif(not (refrun.available() and os.path.exists(REF)),
                                reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def tiles(tmp_path_factory):
    """Two tiles through the reference importer; per sample a merged taginfo in the pipeline's
    post-dedup layout `tile cluster flags polyA mods duplicates` holding a subset of the clusters."""
    from tailseeker_b200.config import stock_config
    from tailseeker_b200.synth import make_tile
    work = str(tmp_path_factory.mktemp("wfq"))
    base = stock_config("miseq")
    rng = np.random.default_rng(11)
    merged = {s.name: [] for s in base.samples}
    for k, tileno in enumerate((1101, 2104)):
        cfg = base.replace(tile=tileno)
        tile = make_tile(cfg, 2500, seed=200 + k, spikein_weight=1.0, unknown_frac=0.05)
        out = refrun.run_import(cfg, tile, workdir=work)
        for s in cfg.samples:
            for line in out["taginfo"][s.name].decode().splitlines():
                cluster, flags, polya, mods, _umi = line.split("\t")
                if rng.random() < 0.7:
                    merged[s.name].append("%s\t%s\t%s\t%s\t%s\t%d\n" % (cfg.tile_id, cluster, flags, polya, mods,
                                                                      int(rng.integers(1, 40))))
    paths = {}
    for name, lines in merged.items():
        paths[name] = os.path.join(work, "taginfo", "%s.txt.gz" % name)
        with gzip.open(paths[name], "wt") as f:
            f.writelines(lines)
    return work, base, paths, merged


@pytest.fixture(autouse=True)
def host_writers(monkeypatch, request):
    """The CPU suite runs the program's host output path (TSK_HOST_WRITERS=1: zlib members); its default,
    the GPU's BGZF deflater, is tests/test_gpu_writefastq.py, which re-runs these tests on a GPU box."""
    if "gpu" not in request.keywords:
        monkeypatch.setenv("TSK_HOST_WRITERS", "1")


def _run(binary, work, taginfo, sample, tag, extra=()):
    r5, r3 = os.path.join(work, "%s_%s_R5.fastq.gz" % (sample, tag)), os.path.join(work, "%s_%s_R3.fastq.gz" % (sample, tag))
    r = subprocess.run([binary, "--taginfo", taginfo, "--seqqual", "seqqual/%s_@tile@.txt.gz" % sample,
                        "--fastq5", r5, "--fastq3", r3, *extra], cwd=work, capture_output=True)
    texts = []
    for p in (r5, r3):
        try:
            with gzip.open(p, "rb") as f:
                texts.append(f.read())
        except (OSError, EOFError):
            texts.append(None)
    return r, texts


@pytest.mark.parametrize("extra", [(), ("--fastq-id-verbose",), ("--fastq-id-verbose", "--threads", "4"),
                                   ("-t", "3")])
def test_fastq_matches_reference(tiles, extra):
    assert os.access(OURS, os.X_OK), "host programs not built (__graft_entry__.build())"
    work, cfg, paths, merged = tiles
    os.environ["TSK_GZ_LEVEL"] = "1"
    total = 0
    for s in cfg.samples:
        a, ta = _run(REF, work, paths[s.name], s.name, "ref", extra)
        b, tb = _run(OURS, work, paths[s.name], s.name, "our", extra)
        assert a.returncode == 0 and b.returncode == 0, (s.name, a.stderr, b.stderr)
        assert ta == tb, s.name
        assert ta[0].count(b"\n") == 4 * len(merged[s.name])
        total += len(merged[s.name])
    assert total > 2000


def test_many_pieces(tiles, tmp_path):
    """More text than one flush: the same record repeated is refused by both (the join consumes a
    seqqual line per taginfo line), so inflate the text with a long list of distinct clusters instead."""
    work, cfg, paths, merged = tiles
    name = max(merged, key=lambda k: len(merged[k]))
    sq = os.path.join(work, "seqqual", "%s_zz9.txt.gz" % name)
    rng = np.random.default_rng(3)
    n = 12000
    with gzip.open(sq, "wt", compresslevel=1) as f:
        for c in range(n):
            s5 = "".join(rng.choice(list("ACGTN"), 51)); s3 = "".join(rng.choice(list("ACGT"), 100))
            f.write("%d\t%s\t%s\t%s\t%s\n" % (c, s5, "F" * 51, s3, "#" * 100))
    tg = str(tmp_path / "big.txt.gz")
    with gzip.open(tg, "wt", compresslevel=1) as f:
        for c in range(0, n, 1):
            f.write("zz9\t%d\t%d\t%d\t%s\t%d\n" % (c, c % 4096, c % 200, "TT"[:c % 3], 1 + c % 7))
    a, ta = _run(REF, work, tg, name, "bigref", ("--fastq-id-verbose",))
    b, tb = _run(OURS, work, tg, name, "bigour", ("--fastq-id-verbose", "--threads", "2"))
    assert a.returncode == 0 and b.returncode == 0
    assert ta == tb and len(ta[1]) > 3 * (1 << 20) // 2


@pytest.mark.parametrize("case,message", [
    ("not_subset", b"The taginfo file must be a subset of seqqual."),
    ("past_end", b"Unexpected format in a seqqual file."),
    ("no_tile_file", b"Failed to open"),
    ("garbage", b"Failed to parse a line"),
    ("no_args", b"One or more required arguments are not specified."),
    ("empty", b"")])
def test_errors_match_reference(tiles, tmp_path, case, message):
    work, cfg, paths, merged = tiles
    name = cfg.samples[0].name
    lines = list(merged[name])
    tile_id = lines[0].split("\t")[0]
    if case == "not_subset":
        lines[5] = lines[4]                                   # the same cluster twice
    elif case == "past_end":
        lines.insert(len([l for l in lines if l.startswith(tile_id)]), "%s\t99999999\t0\t0\t\t1\n" % tile_id)
    elif case == "no_tile_file":
        lines.append("nosuchtile\t1\t0\t0\t\t1\n")
    elif case == "garbage":
        lines.insert(3, "no tabs at all\n")
    elif case == "empty":
        lines = []
    tg = str(tmp_path / "t.txt.gz")
    with gzip.open(tg, "wt") as f:
        f.writelines(lines)
    if case == "no_args":
        a = subprocess.run([REF, "--taginfo", tg], capture_output=True)
        b = subprocess.run([OURS, "--taginfo", tg], capture_output=True)
    else:
        a, ta = _run(REF, work, tg, name, "eref")
        b, tb = _run(OURS, work, tg, name, "eour")
        if case == "empty":
            assert ta == tb == [b"", b""]
    assert a.returncode == b.returncode == (0 if case == "empty" else 255)
    assert message in a.stderr and message in b.stderr
