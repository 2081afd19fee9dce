"""tailseq-writefastq with its default output side -- the GPU's BGZF deflater (tsk_bgzf_deflate) --
against the reference binary: the tests of test_writefastq.py re-run without TSK_HOST_WRITERS, plus a
check that the files really are BGZF."""
import os

import pytest

import test_writefastq as T
from test_gpu_encode import bgzf_blocks

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.access(T.REF, os.X_OK), reason="oracle/_ref not built")]

tiles = T.tiles


@pytest.mark.parametrize("extra", [(), ("--fastq-id-verbose", "--threads", "4")])
def test_fastq_matches_reference_gpu(tiles, extra):
    T.test_fastq_matches_reference(tiles, extra)
    work, cfg, paths, merged = tiles
    name = cfg.samples[0].name
    raw = open(os.path.join(work, "%s_our_R3.fastq.gz" % name), "rb").read()
    blocks = bgzf_blocks(raw)
    assert blocks and blocks[-1] == b"" and len(blocks) >= 2          # data blocks + the end-of-file block


def test_many_pieces_gpu(tiles, tmp_path):
    T.test_many_pieces(tiles, tmp_path)


@pytest.mark.parametrize("case,message", [("not_subset", b"The taginfo file must be a subset of seqqual."),
                                          ("no_args", b"One or more required arguments are not specified."), ("empty", b"")])
def test_errors_match_reference_gpu(tiles, tmp_path, case, message):
    T.test_errors_match_reference(tiles, tmp_path, case, message)
