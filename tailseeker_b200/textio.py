"""Host-side writers of the importer's output formats, driven by the per-cluster calls.

Mirrors (same bytes) the reference's text/binary writers:
  taginfo line    write_taginfo_entry()            src/importer/spotanalyzer.c:270-307
  modifications   get_modification_sequence()      src/importer/spotanalyzer.c:247-267
  seqqual line    write_seqqual_entry()            src/importer/spotanalyzer.c:206-244,310-352
  sigpack header  write_output_file_headers()      src/importer/tailseq-import.c:98-119
  sigdists        write_signal_samples_dists()     src/importer/tailseq-import.c:268-300
  stats CSV       write_demultiplexing_statistics() src/importer/tailseq-import.c:685-714
  ruler stdout    output_corrected_polya_measurements() src/polyaruler/polyaruler.c:316-388
The compute (calls, records, histograms, counters) comes from the CUDA library; this module
only formats.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Optional

import numpy as np

from .config import ImporterConfig, PAFLAG_MEASURED_FROM_FLUORESCENCE

_LETTERS = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMPLEMENT = {ord("A"): "T", ord("C"): "G", ord("G"): "C", ord("T"): "A"}


def decode_basecalls(bcl: np.ndarray):
    """format_basecalls(), bclreader.c:146-166, for whole arrays.
    Returns (seq, qual) as uint8 [N][T] ASCII arrays."""
    b = bcl.T
    nocall = b == 0
    seq = np.where(nocall, np.uint8(ord("N")), _LETTERS[b & 3])
    qual = np.where(nocall, np.uint8(2 + 33), (b >> 2) + np.uint8(33)).astype(np.uint8)
    return np.ascontiguousarray(seq), np.ascontiguousarray(qual)


def format_taginfo(cfg: ImporterConfig, sample: dict, calls: np.ndarray, seq: np.ndarray,
                   firstclusterno: int = 0) -> bytes:
    rows = np.where((calls["sample"] == sample["numindex"]) & (calls["written"] == 1))[0]
    out = []
    umi = sample["umi"]
    for n in rows:
        c = calls[n]
        s = seq[n]
        mods = ""
        tm = int(c["terminal_mods"])
        if tm > 0:
            de = int(c["delimiter_end"])
            mods = "".join(_COMPLEMENT.get(int(ch), "N") for ch in s[de:de + tm][::-1])
        key = b"".join(s[st:st + ln].tobytes() for st, ln in umi if ln > 0)
        out.append(b"%d\t%d\t%d\t%s\t%s\n" % (firstclusterno + int(n), int(c["flags"]),
                                             int(c["polya_len"]), mods.encode(), key))
    return b"".join(out)


def format_seqqual(cfg: ImporterConfig, sample: dict, calls: np.ndarray, seq: np.ndarray,
                   qual: np.ndarray, firstclusterno: int = 0) -> bytes:
    rows = np.where((calls["sample"] == sample["numindex"]) & (calls["written"] == 1))[0]
    f0, fl = cfg.fivep_start - 1, cfg.fivep_length
    ts, tl = cfg.threep_start - 1, cfg.threep_length
    outlen = min(cfg.threep_seqqual_output_length, tl)
    out = []
    for n in rows:
        de = int(calls[n]["delimiter_end"])
        if de >= 0:
            start3, len3 = de, ts + tl - de
        else:
            start3, len3 = ts, tl
        len3 = min(len3, outlen)
        s, q = seq[n], qual[n]
        out.append(b"%d\t%s\t%s\t%s\t%s\n" % (
            firstclusterno + int(n), s[f0:f0 + fl].tobytes(), q[f0:f0 + fl].tobytes(),
            s[start3:start3 + len3].tobytes(), q[start3:start3 + len3].tobytes()))
    return b"".join(out)


def sigpack_bytes(total_clusters: int, dump_len: int, records: np.ndarray) -> bytes:
    """Uncompressed .sigpack content.  `records` u32 [n][2+dump_len]; withdrawn slots
    (negative valid_cycle_count) are skipped."""
    records = np.asarray(records, dtype=np.uint32).reshape(-1, 2 + dump_len)
    valid = (records[:, 1] >> 16).astype(np.uint16).view(np.int16) >= 0
    return struct.pack("<III", 4, total_clusters, dump_len) + records[valid].astype("<u4").tobytes()


def format_stats_csv(cfg: ImporterConfig, counters: np.ndarray) -> str:
    lines = ["name,index,max_index_mm,delim,delim_pos,max_delim_mm,"
             "cln_no_index_mm,cln_1_index_mm,cln_2+_index_mm,"
             "cln_fp_mm,cln_qc_fail,cln_no_delim\n"]
    for s, c in zip(cfg.sample_list(), counters):
        mm0, mm1, mm2, nodelim, fpmm, qcfail = (int(v) for v in c)
        lines.append("%s,%s,%d,%s,%d,%d,%d,%d,%d,%d,%d,%d\n" % (
            s["name"], s["index"], s["maximum_index_mismatches"], s["delimiter"],
            s["delimiter_pos"], s["maximum_delimiter_mismatches"],
            mm0, mm1, mm2, fpmm, qcfail, nodelim))
    return "".join(lines)


def format_ruler_output(tile_id: str, taginfo: bytes, polya_by_cluster: Dict[int, int]) -> bytes:
    """output_corrected_polya_measurements(), polyaruler.c:316-388: prefix the tile id, and
    for measured clusters replace polyA and OR in PAFLAG_MEASURED_FROM_FLUORESCENCE."""
    tid = tile_id.encode()
    out = []
    for line in taginfo.splitlines():
        fields = line.split(b"\t")
        cl = int(fields[0])
        m = polya_by_cluster.get(cl, -1)
        if m < 0:
            out.append(tid + b"\t" + line + b"\n")
        else:
            flags = int(fields[1]) | PAFLAG_MEASURED_FROM_FLUORESCENCE
            out.append(b"%s\t%d\t%d\t%d\t%s\t%s\n" % (tid, cl, flags, m, fields[3], fields[4]))
    return b"".join(out)
