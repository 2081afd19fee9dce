"""GPU: the CUDA path through the C ABI against golden outputs of the unmodified reference."""
import numpy as np
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["importer_miseq", "importer_hiseq_keep"])
def test_importer_and_ruler_match_reference_files(name):
    from tailseeker_b200.importer import TileProcessor, invert_colormatrix
    g = gu.load(name)
    cfg = gu.fixture_config(name)
    proc = TileProcessor(cfg)
    try:
        proc.upload(g["bcl"], g["cif"], invert_colormatrix(gu.matrix_of(g)))
        proc.process()
        S = proc.params.num_samples
        pos, neg = proc.hist()
        gu.check_importer_outputs(cfg, g, proc.calls(), {s: proc.records(s) for s in range(S)}, pos, neg,
                                  proc.counters())
        toks = bytes(g["cutoffs_text"]).decode().split("\t")[1:]
        packed = proc.cutoffs_to_packed([float(t) for t in toks if t.strip()])

        def ruler_fn(recs, _):
            proc.ruler_reset(cfg.total_cycles)
            lens = proc.ruler_records(recs, packed)
            return lens, proc.ruler_hist()

        gu.check_ruler_outputs(cfg, g, ruler_fn)
    finally:
        proc.close()
