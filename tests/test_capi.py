"""CPU: the C-ABI shared library loads and exports every symbol include/tailseeker_b200.h declares;
the ctypes mirrors have the C layout.  No compute calls (no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tailseeker_b200.h")


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tsk_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from tailseeker_b200 import _lib, build
    build.build()
    lib = _lib.load()
    declared = _declared_symbols()
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), "libtailseeker_b200.so does not export %s" % name
    assert sorted(_lib.SYMBOLS) == declared


def test_no_gpu_means_loud_failure():
    """There is no CPU fallback: without a B200 every compute entry point fails with a message."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from tailseeker_b200 import _lib
    from tailseeker_b200.config import stock_config
    from tailseeker_b200.importer import TileProcessor
    assert _lib.load().tsk_probe(0) == -1
    with pytest.raises(_lib.TskError):
        TileProcessor(stock_config("miseq"))


def test_programs_fail_loudly_without_gpu(tmp_path):
    """The drop-in importer (one or several tiles per process) has no CPU path either: it reads the tile,
    then stops with the library's message and the reference's failure status (-1 -> 255)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from tailseeker_b200.config import stock_config
    from tailseeker_b200.synth import make_tile, write_tile_files
    exe = os.path.join(ROOT, "tailseeker_b200", "bin", "tailseq-import")
    assert os.access(exe, os.X_OK), "host programs not built (__graft_entry__.build())"
    cfg = stock_config("miseq")
    tile = make_tile(cfg, 50, seed=1)
    data = str(tmp_path / "data")
    mpath = write_tile_files(tile, cfg, data)
    for sub in ("seqqual", "taginfo", "signals", "sigdists-r00", "stats"):
        os.makedirs(str(tmp_path / sub))
    ini = str(tmp_path / "t.ini")
    with open(ini, "w") as f:
        f.write(cfg.replace(data_dir=data, threep_colormatrix=mpath).to_ini())
    r = subprocess.run([exe, ini, ini], cwd=str(tmp_path), capture_output=True)
    assert r.returncode == 255
    assert b"no CPU path" in r.stderr and b"Finished." not in r.stdout
    assert r.stdout.count(b"Opening input sources") == 1          # the chain stops at the first tile


def test_ctypes_layout_matches_header(tmp_path):
    from tailseeker_b200.config import CCall, CParams, CSample
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "%s"\n'
                   'int main(void){printf("%%zu %%zu %%zu %%zu %%zu %%zu\\n", sizeof(tsk_params), sizeof(tsk_sample),'
                   'sizeof(tsk_call), sizeof(tsk_counters), offsetof(tsk_params, samples),'
                   'offsetof(tsk_params, t_intensity_score));return 0;}\n' % HEADER)
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-o", str(exe), str(src)])
    vals = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert vals[0] == ctypes.sizeof(CParams)
    assert vals[1] == ctypes.sizeof(CSample)
    assert vals[2] == ctypes.sizeof(CCall) == 16
    assert vals[3] == 24
    assert vals[4] == CParams.samples.offset
    assert vals[5] == CParams.t_intensity_score.offset


def test_config_derived_values():
    """compute_derived_values() mirror: list order, absolute delimiter positions, dump lengths."""
    from tailseeker_b200.config import stock_config
    cfg = stock_config("miseq")
    sl = cfg.sample_list()
    # LIFO of the generator's order (experimental then spike-ins, each sorted by name)
    assert [s["name"] for s in sl][:2] == ["Unknown", "spkN"] and sl[-1]["name"] == "Control1"
    exp, spk = sl[-1], next(s for s in sl if s["name"] == "spk128")
    assert exp["delimiter_pos"] == 57 + 15 and exp["signal_dump_length"] == 231
    assert spk["delimiter_pos"] == 57 + 10 and spk["signal_dump_length"] == 182
    assert sl[0]["signal_dump_length"] == 252          # pseudo-sample quirk (delimiter_pos = -1)
    assert cfg.min_bases_passes() == 14
    p = cfg.to_cparams()
    assert abs(p.maximum_entropy - 1.3862943649291992) < 1e-7
    assert p.t_intensity_score[200] == 1.0 and 0 < p.t_intensity_score[0] < 1e-4
    ini = cfg.to_ini()
    assert "[sample:Control1]" in ini and "umi-start:2 = 58" in ini and "limit-threep-processing = 182" in ini
