"""Regenerates tests/golden/generated_signalproc.ini: the per-tile INI that the REFERENCE's own
scripts/generate-signalproc-conf.py (:224-239) renders for templates/miseq-v2.yaml (+ the conf/
includes it names), run from /root/reference under oracle/snakemake_stub.  Only run where the
reference tree exists; the output (generated text, not reference source) is what the tests use.
    python tests/golden/make_generated_ini.py"""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
STUB = os.path.join(ROOT, "oracle", "snakemake_stub")

CODE = r'''
import json, sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
from tailseeker.configurations import Configurations
conf = Configurations(%r, open(%r))
tile = "a1101"
tileinfo = {tile: dict(id=tile, source="a", laneid="a", lane=1, tile=1101, type="MiSeq-V2",
                       datadir="@DATADIR@", intensitiesdir="@DATADIR@")}
json.dump({"input": [], "output": [["sigproc_conf", %r]], "threads": 4, "wildcards": [["tile", tile]],
           "params": [["tileinfo", tileinfo], ["conf", conf.confdata], ["exp_samples", conf.exp_samples],
                      ["spikein_samples", conf.spikein_samples]]}, open(%r, "w"))
'''


def main():
    work = tempfile.mkdtemp(prefix="tskini_")
    ini, pj = os.path.join(work, "tile.ini"), os.path.join(work, "params.json")
    code = CODE % (STUB, REF, REF, os.path.join(REF, "templates", "miseq-v2.yaml"), ini, pj)
    subprocess.check_call([sys.executable, "-c", code], cwd=work)
    env = dict(os.environ, SNAKEMAKE_PARAMS=pj, PYTHONPATH=STUB + ":" + REF)
    subprocess.check_call([sys.executable, os.path.join(REF, "scripts", "generate-signalproc-conf.py")], env=env, cwd=work)
    text = open(ini).read()
    out = os.path.join(ROOT, "tests", "golden", "generated_signalproc.ini")
    open(out, "w").write(text)
    print(out, len(text), "bytes")


if __name__ == "__main__":
    main()
