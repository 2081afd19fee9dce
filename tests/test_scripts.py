"""Script surfaces of the path (CPU part): the INI that the reference's own
scripts/generate-signalproc-conf.py renders (tests/golden/generated_signalproc.ini, made by
tests/golden/make_generated_ini.py) is read alike by both INI parsers and says what
stock_config("miseq") says; the SNAKEMAKE_PARAMS reader; the batched cutoff bookkeeping."""
import json
import os
import subprocess

import numpy as np
import pytest

import golden_util as gu
from test_estimator import _sigdists

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "tailseeker_b200", "bin", "tsk-ini-dump")
REF = os.path.join(ROOT, "oracle", "_ref", "ini-dump")
GENERATED = os.path.join(ROOT, "tests", "golden", "generated_signalproc.ini")


def _triples(binary, path):
    out = subprocess.run([binary, path], capture_output=True, check=True).stdout.decode()
    return [tuple(ln.split("\x1f")) for ln in out.splitlines()]


@pytest.mark.skipif(not os.access(OURS, os.X_OK), reason="host programs not built")
def test_generated_ini_says_what_the_mirror_says(tmp_path):
    from tailseeker_b200.config import stock_config
    gen = _triples(OURS, GENERATED)
    if os.access(REF, os.X_OK):
        assert gen == _triples(REF, GENERATED)
    p = tmp_path / "mirror.ini"
    p.write_text(stock_config("miseq", phix_match_name="PhiX").to_ini())
    mine = _triples(OURS, str(p))
    skip = {("source", "data-dir"), ("source", "threep-colormatrix"), ("options", "threads"),
            ("options", "read-buffer-size")}
    a = {(s, n): v for s, n, v in gen if (s, n) not in skip and s not in ("output", "alternative_calls")}
    b = {(s, n): v for s, n, v in mine if (s, n) not in skip and s not in ("output", "alternative_calls")}
    assert set(a) == set(b), (set(a) ^ set(b))
    for k in a:
        try:
            assert float(a[k]) == float(b[k]), k
        except ValueError:
            assert a[k] == b[k], k
    # sample sections in the generator's order (experimental, then spike-ins, each sorted)
    order = [s for s, n, v in gen if s.startswith("sample:") and n == "index"]
    assert order == [s for s, n, v in mine if s.startswith("sample:") and n == "index"] and len(order) == 11
    # the output templates the pipeline expects
    outs = {n: v for s, n, v in gen if s == "output"}
    assert outs["seqqual"] == "scratch/seqqual/{name}_a1101.txt.gz" and outs["signal-dists"].count("{posneg}") == 1


def test_params_reader(tmp_path, monkeypatch):
    from tailseeker_b200.scripts import _params
    pj = tmp_path / "p.json"
    pj.write_text(json.dumps({"input": [["signals", "a.sigpack"], ["signals", "b.sigpack"], ["score_cutoffs", "c.txt"], [None, "x"]],
                              "output": [["sigdists", "o.sigdists"]], "threads": 3, "wildcards": [["tile", "a1101"]],
                              "params": [["CONF", {"k": {"v": 1}}], ["total_cycles", 308]]}))
    monkeypatch.setenv("SNAKEMAKE_PARAMS", str(pj))
    v = _params.load()
    assert "SNAKEMAKE_PARAMS" not in os.environ
    assert v["threads"] == 3 and v["input"].signals == ["a.sigpack", "b.sigpack"] and v["input"].score_cutoffs == "c.txt"
    assert list(v["input"]) == ["a.sigpack", "b.sigpack", "c.txt", "x"] and v["input"][3] == "x"
    assert v["params"]["CONF"]["k"]["v"] == 1 and v["params"].total_cycles == 308 and v["wildcards"].tile == "a1101"
    with pytest.raises(AttributeError):
        v["input"].nothing


def test_batched_bookkeeping_equals_per_tile():
    """One stacked fit call == one call per tile: same cutoffs text as the reference script's golden output."""
    from oracle import estimator_oracle
    from tailseeker_b200 import estimator
    g = gu.load("estimator_two_tiles")
    sig = _sigdists(g)
    T = int(g["shape"][0])
    calls = []

    def fit(p, n):
        calls.append(p.shape)
        return estimator_oracle.fit_cutoffs(p, n, 300, 0.1, 0.0, (0.4, 0.6))

    cut, reports = estimator.calculate_optimal_cutoffs_batched(fit, sig, ["a1102", "a1101"], "run", T, 300)
    assert calls == [(3 * T, int(g["shape"][1]))]
    assert estimator.format_cutoffs(cut, T) == bytes(g["text"]).decode()
    assert estimator.format_bases(reports) == bytes(g["bases"]).decode()
