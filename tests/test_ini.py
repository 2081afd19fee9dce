"""The importer's INI reader (csrc/host/ini.c, written for this library) against the reference's
vendored parser (oracle/_ref/ini-dump = src/contrib/ini.c compiled where it lies): same
(section, name, value) triples on awkward files and on random line soups, and on the INI the
configuration mirror renders.  Without oracle/_ref the awkward file is checked against the
committed triples."""
import os
import random
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "tailseeker_b200", "bin", "tsk-ini-dump")
REF = os.path.join(ROOT, "oracle", "_ref", "ini-dump")

AWKWARD = (
    b"\xef\xbb\xbf[source]\n"
    b"data-dir = /data/run 1 ; the run folder\n"
    b"  ; an indented comment, not a continuation\n"
    b"  continued text ; kept as is\n"
    b"laneid=a\n"
    b"\n"
    b"\tmore of laneid\n"
    b"# hash comment\n"
    b"[read_format] ; trailing comment\n"
    b"[broken ;] section\n"
    b"[ spaced name ]\n"
    b"fivep-start = 1\r\n"
    b"name with spaces   =   value with  spaces   \n"
    b"empty =\n"
    b"=novalue-name\n"
    b"semi=;not a comment\n"
    b"semi2 = a;b ;c\n"
    b"eq = a=b=c\n"
    b"no equals sign here\n"
    b"key ;= commented out before the equals sign\n"
    b"[unterminated\n"
    b"[sample:" + b"x" * 70 + b"]\n"
    b"y" * 70 + b" = long name\n"
    b"  continuation of a long name\n"
    b"long = " + b"v" * 1100 + b"\n"
    b"after = 1\n"
    b"last = no newline")

WANT = None


def _dump(binary, path):
    r = subprocess.run([binary, path], capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    return r.stdout


def _have(binary):
    return os.access(binary, os.X_OK)


@pytest.mark.skipif(not _have(OURS), reason="host programs not built")
def test_awkward_file(tmp_path):
    p = tmp_path / "a.ini"
    p.write_bytes(AWKWARD)
    ours = _dump(OURS, str(p))
    golden = os.path.join(ROOT, "tests", "golden", "ini_awkward.triples")
    if _have(REF):
        ref = _dump(REF, str(p))
        assert ours == ref
        if os.environ.get("TSK_WRITE_GOLDEN"):
            open(golden, "wb").write(ref)
    assert ours == open(golden, "rb").read()
    # the cases above are really exercised
    text = ours.decode("latin-1")
    assert "source\x1fdata-dir\x1f/data/run 1\n" in text
    assert "source\x1fdata-dir\x1fcontinued text ; kept as is\n" in text
    assert "semi\x1f;not a comment" in text and "semi2\x1fa;b\n" in text


@pytest.mark.skipif(not (_have(OURS) and _have(REF)), reason="needs oracle/_ref/ini-dump")
def test_random_line_soups(tmp_path):
    rng = random.Random(12)
    atoms = ["[", "]", "=", ";", "#", " ", "\t", "a", "bc", "name", "1.5", "\r", " ;", "= ", "[s]", "\x0c"]
    for i in range(300):
        lines = []
        for _ in range(rng.randint(1, 25)):
            ln = "".join(rng.choice(atoms) for _ in range(rng.randint(0, 12)))
            if rng.random() < 0.03:
                ln += "z" * rng.randint(1000, 2100)
            lines.append(ln)
        p = tmp_path / "r.ini"
        p.write_bytes(("\n".join(lines) + ("\n" if rng.random() < 0.8 else "")).encode("latin-1"))
        assert _dump(OURS, str(p)) == _dump(REF, str(p)), "soup %d" % i


@pytest.mark.skipif(not (_have(OURS) and _have(REF)), reason="needs oracle/_ref/ini-dump")
def test_rendered_configuration(tmp_path):
    from tailseeker_b200.config import stock_config
    for layout in ("miseq", "hiseq"):
        p = tmp_path / (layout + ".ini")
        p.write_text(stock_config(layout, phix_match_name="PhiX").to_ini())
        a, b = _dump(OURS, str(p)), _dump(REF, str(p))
        assert a == b and a.count(b"\n") > 60
