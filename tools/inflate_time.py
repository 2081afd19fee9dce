#!/usr/bin/env python
"""Device-side inflate of a tile's base-call files (.bcl.gz, gzip level 6) against zlib on the host
cores: wall time of handing a 1.07 M-cluster tile over row by row with the 308 base-call rows as
plain arrays, and with all of them as compressed files.
    python tools/inflate_time.py [clusters]"""
import gzip
import json
import os
import struct
import sys
import time
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tailseeker_b200.config import stock_config                              # noqa: E402
from tailseeker_b200.importer import TileProcessor, invert_colormatrix       # noqa: E402
from tailseeker_b200.synth_torch import make_tile_device                     # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_070_000
    cfg = stock_config("miseq", phix_match_name="PhiX")
    bcl, cif, mtext, pitch = make_tile_device(cfg, n, seed=5, device="cuda", phix_frac=0.6)
    hb = np.ascontiguousarray(bcl[:, :n].cpu().numpy())
    hc = np.ascontiguousarray(cif[:, :, :n].cpu().numpy())
    del bcl, cif
    inv = invert_colormatrix(np.array([np.float32(x) for x in mtext.split()], dtype=np.float32))
    with ThreadPoolExecutor(os.cpu_count()) as ex:
        raws = list(ex.map(lambda c: gzip.compress(struct.pack("<I", n) + hb[c].tobytes(), 6), range(hb.shape[0])))
        t0 = time.perf_counter()
        outs = list(ex.map(lambda r: zlib.decompress(r, 31), raws))
        host_s = time.perf_counter() - t0
    assert all(o[4:] == hb[c].tobytes() for c, o in enumerate(outs))
    gz_rows = {c: (raws[c], 4 + n, 4) for c in range(hb.shape[0])}
    p = TileProcessor(cfg)
    res = {}
    for name, rows in (("arrays", None), ("gz", gz_rows)):
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            p.upload_rows(hb, hc, inv, gz_rows=rows)
            best = min(best, time.perf_counter() - t0)
            p.process()
        res[name] = best
    calls = p.calls()
    p.upload(hb, hc, inv); p.process()
    assert (p.calls() == calls).all()
    p.close()
    comp = sum(len(r) for r in raws)
    print(json.dumps({"clusters": n, "bcl_bytes": int(hb.nbytes), "gz_bytes": comp, "ratio": comp / hb.nbytes,
                      "rowwise_upload_arrays_s": res["arrays"], "rowwise_upload_all_bcl_gz_s": res["gz"],
                      "device_inflate_s": res["gz"] - res["arrays"] + hb.nbytes / 12e9,
                      "host_zlib_inflate_s": host_s, "host_threads": os.cpu_count()}, indent=1))


if __name__ == "__main__":
    main()
