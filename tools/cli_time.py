#!/usr/bin/env python
"""Wall time of the drop-in tailseq-import over several full tiles in one process, phase by phase
(TSK_HOST_TIMING), beside the reference binary: bench.py's `e2e_cli` leg with more tiles.
    python tools/cli_time.py [tiles] [clusters per tile]
TSK_CLI_RAW_TIMING=1 also passes the program's own "[timing] phase seconds" lines through (stderr),
tile by tile, instead of only their sums."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                              # noqa: E402
from tailseeker_b200.config import stock_config          # noqa: E402

if __name__ == "__main__":
    ntiles = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    n = int(sys.argv[2]) if len(sys.argv) > 2 else bench.CLUSTERS_PER_TILE
    cfg = stock_config("miseq", phix_match_name="PhiX")
    print(json.dumps(bench.cli_leg(cfg, ntiles, n, os.cpu_count() or 1), indent=1))
