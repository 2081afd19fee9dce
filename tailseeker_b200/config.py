"""Host-side mirror of the reference's importer configuration.

Mirrors, with the same names and the same derived values,
  * the INI schema emitted by ``scripts/generate-signalproc-conf.py:31-221`` and parsed by
    ``src/importer/parseconfig.c:39-479``,
  * the defaults of ``parseconfig.c:482-534`` and ``conf/defaults.conf``,
  * ``compute_derived_values()`` (``parseconfig.c:537-671``): sample list order
    (Unknown, PhiX, then ``[sample:*]`` sections in reverse file order), absolute
    ``delimiter_pos``, ``signal_dump_length``, ``min_bases_passes``,
  * ``precalc_score_tables()`` (``signalproc.c:54-82``) evaluated with the HOST libm
    (``logf``/``expf`` through ctypes) because the device must not evaluate them.

The result is packed into the POD ``tsk_params`` of ``include/tailseeker_b200.h``.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import dataclasses
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

TSK_ABI_VERSION = 3
TSK_MAX_SAMPLES = 64
TSK_MAX_SEQ = 32
TSK_T_SCORE_BINS = 200
TSK_SCORE_MAX = (1 << 24) - 1

PAFLAG_POLYA_DETECTED = 0x0001
PAFLAG_DELIMITER_HAS_MISMATCH = 0x0002
PAFLAG_HAVE_3P_MODIFICATION = 0x0004
PAFLAG_MEASURED_FROM_FLUORESCENCE = 0x0008
PAFLAG_BARCODE_HAS_MISMATCHES = 0x0010
PAFLAG_DELIMITER_IS_SHIFTED = 0x0020
PAFLAG_DARKCYCLE_EXISTS = 0x0040
PAFLAG_DELIMITER_NOT_FOUND = 0x0080
PAFLAG_BALANCER_CALL_QUALITY_BAD = 0x0100
PAFLAG_BALANCER_BIASED = 0x0200
PAFLAG_BALANCER_SIGNAL_BAD = 0x0400
PAFLAG_DARKCYCLE_OVER_THRESHOLD = 0x0800


# --------------------------------------------------------------------------------------
# ctypes mirror of include/tailseeker_b200.h
# --------------------------------------------------------------------------------------
class CSample(ctypes.Structure):
    _fields_ = [
        ("index", ctypes.c_char * TSK_MAX_SEQ),
        ("delimiter", ctypes.c_char * TSK_MAX_SEQ),
        ("fingerprint", ctypes.c_char * TSK_MAX_SEQ),
        ("maximum_index_mismatches", ctypes.c_int32),
        ("delimiter_pos", ctypes.c_int32),
        ("delimiter_length", ctypes.c_int32),
        ("maximum_delimiter_mismatches", ctypes.c_int32),
        ("fingerprint_pos", ctypes.c_int32),
        ("fingerprint_length", ctypes.c_int32),
        ("maximum_fingerprint_mismatches", ctypes.c_int32),
        ("umi_ranges_count", ctypes.c_int32),
        ("limit_threep_processing", ctypes.c_int32),
        ("signal_dump_length", ctypes.c_int32),
    ]


class CParams(ctypes.Structure):
    _fields_ = [
        ("abi_version", ctypes.c_int32),
        ("total_cycles", ctypes.c_int32),
        ("index_start", ctypes.c_int32),
        ("index_length", ctypes.c_int32),
        ("threep_start", ctypes.c_int32),
        ("threep_length", ctypes.c_int32),
        ("keep_no_delimiter", ctypes.c_int32),
        ("keep_low_quality_balancer", ctypes.c_int32),
        ("has_control", ctypes.c_int32),
        ("control_first_cycle", ctypes.c_int32),
        ("control_read_length", ctypes.c_int32),
        ("balancer_start", ctypes.c_int32),
        ("balancer_length", ctypes.c_int32),
        ("balancer_min_quality", ctypes.c_int32),
        ("balancer_min_bases_passes", ctypes.c_int32),
        ("balancer_minimum_occurrence", ctypes.c_int32),
        ("balancer_num_positive_samples", ctypes.c_int32),
        ("balancer_num_negative_samples", ctypes.c_int32),
        ("seed_trigger_polya_length", ctypes.c_int32),
        ("negative_sample_polya_length", ctypes.c_int32),
        ("max_cctr_scan_left_space", ctypes.c_int32),
        ("max_cctr_scan_right_space", ctypes.c_int32),
        ("required_cdf_contrast", ctypes.c_float),
        ("polya_boundary_pos", ctypes.c_int32),
        ("polya_sampling_gap", ctypes.c_int32),
        ("dist_sampling_bins", ctypes.c_int32),
        ("fair_sampling_fingerprint_length", ctypes.c_int32),
        ("fair_sampling_hash_space_size", ctypes.c_int32),
        ("fair_sampling_max_count", ctypes.c_int32),
        ("weights_polyA", ctypes.c_int32 * 5),
        ("weights_nonA", ctypes.c_int32 * 5),
        ("max_terminal_modifications", ctypes.c_int32),
        ("min_polya_length", ctypes.c_int32),
        ("sigproc_trigger_polya_length", ctypes.c_int32),
        ("dark_cycles_threshold", ctypes.c_float),
        ("max_dark_cycles", ctypes.c_int32),
        ("maximum_entropy", ctypes.c_float),
        ("t_intensity_score", ctypes.c_float * (TSK_T_SCORE_BINS + 1)),
        ("num_samples", ctypes.c_int32),
        ("samples", CSample * TSK_MAX_SAMPLES),
    ]


TSK_ENCODE_MAX_UMI = 8


class CEncodeSample(ctypes.Structure):
    """tsk_encode_sample (include/tailseeker_b200.h)."""
    _fields_ = [("want_seqqual", ctypes.c_int32), ("want_taginfo", ctypes.c_int32), ("want_signal", ctypes.c_int32),
                ("umi_count", ctypes.c_int32),
                ("umi_start", ctypes.c_int32 * TSK_ENCODE_MAX_UMI), ("umi_length", ctypes.c_int32 * TSK_ENCODE_MAX_UMI)]


class CEncodeLayout(ctypes.Structure):
    """tsk_encode_layout (include/tailseeker_b200.h)."""
    _fields_ = [("fivep_start", ctypes.c_int32), ("fivep_length", ctypes.c_int32),
                ("threep_seqqual_output_length", ctypes.c_int32),
                ("samples", CEncodeSample * TSK_MAX_SAMPLES)]


class CCall(ctypes.Structure):
    _fields_ = [
        ("sample", ctypes.c_int16),
        ("flags", ctypes.c_uint16),
        ("delimiter_end", ctypes.c_int16),
        ("polya_len", ctypes.c_int16),
        ("terminal_mods", ctypes.c_int16),
        ("written", ctypes.c_uint8),
        ("emitted", ctypes.c_uint8),
        ("record_slot", ctypes.c_uint32),
    ]


CALL_DTYPE = np.dtype([
    ("sample", "<i2"), ("flags", "<u2"), ("delimiter_end", "<i2"), ("polya_len", "<i2"),
    ("terminal_mods", "<i2"), ("written", "u1"), ("emitted", "u1"), ("record_slot", "<u4"),
])
assert CALL_DTYPE.itemsize == 16 and ctypes.sizeof(CCall) == 16

COUNTER_FIELDS = ("clusters_mm0", "clusters_mm1", "clusters_mm2plus",
                  "clusters_nodelim", "clusters_fpmismatch", "clusters_qcfailed")


# --------------------------------------------------------------------------------------
# host libm (the reference evaluates logf/expf with the host's libm; so must we)
# --------------------------------------------------------------------------------------
_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.logf.restype = ctypes.c_float
_libm.logf.argtypes = [ctypes.c_float]
_libm.expf.restype = ctypes.c_float
_libm.expf.argtypes = [ctypes.c_float]


def host_logf(x: float) -> np.float32:
    return np.float32(_libm.logf(ctypes.c_float(float(x))))


def host_expf(x: float) -> np.float32:
    return np.float32(_libm.expf(ctypes.c_float(float(x))))


def calculate_maximum_entropy() -> np.float32:
    """signalproc.c:39-64: -sum p*logf(p) over four p = 1/4, float accumulation."""
    res = np.float32(0.0)
    p = np.float32(1.0 / 4)
    for _ in range(4):
        res = np.float32(res - np.float32(p * host_logf(p)))
    return res


def precalc_t_intensity_score(k: float, center: float) -> np.ndarray:
    """signalproc.c:67-82 with the same float/double mix."""
    kf, cf = np.float32(k), np.float32(center)
    table = np.empty(TSK_T_SCORE_BINS + 1, dtype=np.float32)
    for i in range(TSK_T_SCORE_BINS):
        vpoint = np.float32((1.0 * (0.5 + i)) / TSK_T_SCORE_BINS)
        arg = np.float32(np.float32(-kf) * np.float32(vpoint - cf))
        table[i] = np.float32(1.0 / (1.0 + float(host_expf(arg))))
    table[TSK_T_SCORE_BINS] = np.float32(1.0)
    return table


# --------------------------------------------------------------------------------------
# configuration dataclasses (INI names, with '-' -> '_')
# --------------------------------------------------------------------------------------
@dataclass
class SampleConf:
    """One ``[sample:NAME]`` section (parseconfig.c:388-438). Starts are 1-based as in the INI;
    ``delimiter_start`` is relative to the 3' read, ``umi``/``fingerprint_start`` absolute."""
    name: str
    index: str
    maximum_index_mismatch: int = 0
    delimiter_seq: Optional[str] = None
    delimiter_start: int = 0
    maximum_delimiter_mismatch: int = 0
    fingerprint_seq: Optional[str] = None
    fingerprint_start: int = 0
    maximum_fingerprint_mismatch: int = 0
    umi: List[Tuple[int, int]] = field(default_factory=list)   # (start 1-based, length)
    limit_threep_processing: int = 0


@dataclass
class ImporterConfig:
    # [source]
    data_dir: str = "."
    laneid: str = "a"
    lane: int = 1
    tile: int = 1101
    threep_colormatrix: str = ""
    # [read_format] (1-based starts, as written in the INI)
    total_cycles: int = 308
    fivep_start: int = 1
    fivep_length: int = 51
    index_start: int = 52
    index_length: int = 6
    threep_start: int = 58
    threep_length: int = 251
    threep_seqqual_output_length: int = 100
    # [options]
    keep_no_delimiter: int = 0
    keep_low_quality_balancer: int = 0
    threads: int = 1
    read_buffer_size: int = 2147483648
    # [output]
    seqqual: str = "seqqual/{name}_{tile}.txt.gz"
    taginfo: str = "taginfo/{name}_{tile}.txt.gz"
    signal: str = "signals/{name}_{tile}.sigpack"
    signal_dists: str = "sigdists-r00/{posneg}_{tile}.sigdists"
    stats: str = "stats/signal-proc-{tile}.csv"
    # [alternative_calls]: (first cycle, 1-based; FASTQ path) in INI order (parseconfig.c:153-185)
    alternative_calls: Tuple[Tuple[int, str], ...] = ()
    # [control]
    phix_match_name: str = ""          # "" = no control (parseconfig.c:576)
    phix_match_start: int = 6
    phix_match_length: int = 40
    # [balancer]  (conf/defaults.conf:33-40)
    balancer_start: int = 1
    balancer_length: int = 20
    balancer_minimum_occurrence: int = 2
    balancer_minimum_quality: int = 25
    balancer_minimum_qcpass_percent: float = 70
    balancer_num_positive_samples: int = 2
    balancer_num_negative_samples: int = 4
    # [polyA_seeder]  (conf/defaults.conf:69-85)
    seed_trigger_polya_length: int = 10
    negative_sample_polya_length: int = 0
    max_cctr_scan_left_space: int = 20
    max_cctr_scan_right_space: int = 20
    required_cdf_contrast: float = 0.35
    polya_boundary_pos: int = 120
    polya_sampling_gap: int = 3
    dist_sampling_bins: int = 1000
    fair_sampling_fingerprint_length: int = 30
    fair_sampling_hash_space_size: int = 1048576
    fair_sampling_max_count: int = 5
    # [polyA_finder]  (conf/defaults.conf:46-60)
    polyA_weights: Dict[str, int] = field(
        default_factory=lambda: dict(T=2, A=-9, C=-9, G=-9, N=-1))
    nonA_weights: Dict[str, int] = field(
        default_factory=lambda: dict(T=-1, A=0, C=-4, G=-4, N=0))
    minimum_polya_length: int = 5
    maximum_modifications: int = 20
    signal_analysis_trigger: int = 8
    # [polyA_ruler]  (conf/defaults.conf:62-69)
    dark_cycles_threshold: float = 10
    maximum_dark_cycles: int = 5
    t_intensity_k: float = 17.0
    t_intensity_center: float = 0.6
    downhill_extension_weight: float = 0.49
    signal_resampling_gap: float = 0.1
    # estimator (conf/defaults.conf:76-80)
    cutoff_score_search_low: float = 0.0
    cutoff_score_search_high: Tuple[float, ...] = (0.4, 0.6)
    minimum_spots_for_dist_sampling: int = 300
    kde_bandwidth_for_dist: float = 0.1
    # samples in INI file order
    samples: List[SampleConf] = field(default_factory=list)

    # ---------------------------------------------------------------- INI round trip
    def to_ini(self) -> str:
        """Same sections/keys/order as generate-signalproc-conf.py:31-221."""
        o = []
        o.append("[source]\ndata-dir = %s\nlaneid = %s\nlane = %d\ntile = %d\n"
                 "threep-colormatrix = %s\n" % (self.data_dir, self.laneid, self.lane,
                                               self.tile, self.threep_colormatrix))
        o.append("[read_format]\ntotal-cycles = %d\nfivep-start = %d\nfivep-length = %d\n"
                 "index-start = %d\nindex-length = %d\nthreep-start = %d\nthreep-length = %d\n"
                 % (self.total_cycles, self.fivep_start, self.fivep_length, self.index_start,
                    self.index_length, self.threep_start, self.threep_length))
        o.append("[options]\nkeep-no-delimiter = %d\nkeep-low-quality-balancer = %d\n"
                 "threads = %d\nread-buffer-size = %d\n"
                 % (self.keep_no_delimiter, self.keep_low_quality_balancer, self.threads,
                    self.read_buffer_size))
        o.append("[output]\nseqqual = %s\ntaginfo = %s\nsignal = %s\nsignal-dists = %s\n"
                 "stats = %s\n" % (self._tile(self.seqqual), self._tile(self.taginfo),
                                   self._tile(self.signal), self._tile(self.signal_dists),
                                   self._tile(self.stats)))
        o.append("[alternative_calls]\n" +
                 "".join("%d = %s\n" % (first, path) for first, path in self.alternative_calls))
        if self.phix_match_name:
            o.append("[control]\nphix-match-name = %s\nphix-match-start = %d\n"
                     "phix-match-length = %d\n" % (self.phix_match_name, self.phix_match_start,
                                                   self.phix_match_length))
        o.append("[balancer]\nstart = %d\nlength = %d\nminimum-occurrence = %d\n"
                 "minimum-quality = %d\nminimum-qcpass-percent = %s\n"
                 "num-positive-samples = %d\nnum-negative-samples = %d\n"
                 % (self.balancer_start, self.balancer_length, self.balancer_minimum_occurrence,
                    self.balancer_minimum_quality, _num(self.balancer_minimum_qcpass_percent),
                    self.balancer_num_positive_samples, self.balancer_num_negative_samples))
        o.append("[polyA_seeder]\nseed-trigger-polya-length = %d\n"
                 "negative-sample-polya-length = %d\nmax-cctr-scan-left-space = %d\n"
                 "max-cctr-scan-right-space = %d\nrequired-cdf-contrast = %s\n"
                 "polya-boundary-pos = %d\npolya-sampling-gap = %d\ndist-sampling-bins = %d\n"
                 "fair-sampling-fingerprint-length = %d\nfair-sampling-hash-space-size = %d\n"
                 "fair-sampling-max-count = %d\n"
                 % (self.seed_trigger_polya_length, self.negative_sample_polya_length,
                    self.max_cctr_scan_left_space, self.max_cctr_scan_right_space,
                    _num(self.required_cdf_contrast), self.polya_boundary_pos,
                    self.polya_sampling_gap, self.dist_sampling_bins,
                    self.fair_sampling_fingerprint_length, self.fair_sampling_hash_space_size,
                    self.fair_sampling_max_count))
        pw, nw = self.polyA_weights, self.nonA_weights
        o.append("[polyA_finder]\n" +
                 "".join("polyA-weight-%s = %d\n" % (b, pw[b]) for b in "TACGN") +
                 "".join("nonA-weight-%s = %d\n" % (b, nw[b]) for b in "TACGN") +
                 "minimum-polya-length = %d\nmaximum-modifications = %d\n"
                 "signal-analysis-trigger = %d\n"
                 % (self.minimum_polya_length, self.maximum_modifications,
                    self.signal_analysis_trigger))
        o.append("[polyA_ruler]\ndark-cycles-threshold = %s\nmaximum-dark-cycles = %d\n"
                 "t-intensity-k = %s\nt-intensity-center = %s\n"
                 % (_num(self.dark_cycles_threshold), self.maximum_dark_cycles,
                    _num(self.t_intensity_k), _num(self.t_intensity_center)))
        for s in self.samples:
            sec = "[sample:%s]\nindex = %s\nmaximum-index-mismatch = %d\n" % (
                s.name, s.index, s.maximum_index_mismatch)
            if s.delimiter_seq is not None:
                sec += ("delimiter-seq = %s\ndelimiter-start = %d\n"
                        "maximum-delimiter-mismatch = %d\n"
                        % (s.delimiter_seq, s.delimiter_start, s.maximum_delimiter_mismatch))
            for i, (start, length) in enumerate(s.umi):
                sec += "umi-start:%d = %d\numi-length:%d = %d\n" % (i + 1, start, i + 1, length)
            if s.fingerprint_seq is not None:
                sec += ("fingerprint-seq = %s\nfingerprint-start = %d\n"
                        "maximum-fingerprint-mismatch = %d\n"
                        % (s.fingerprint_seq, s.fingerprint_start,
                           s.maximum_fingerprint_mismatch))
            if s.limit_threep_processing:
                sec += "limit-threep-processing = %d\n" % s.limit_threep_processing
            o.append(sec)
        return "\n".join(o)

    def _tile(self, template: str) -> str:
        return template.replace("{tile}", "%s%04d" % (self.laneid, self.tile))

    @property
    def tile_id(self) -> str:
        return "%s%04d" % (self.laneid, self.tile)

    # ---------------------------------------------------------------- derived values
    def sample_list(self) -> List[dict]:
        """Sample list in the reference's order with derived fields
        (parseconfig.c:550-641)."""
        ts = self.threep_start - 1
        out: List[dict] = []

        def pseudo(name):
            return dict(name=name, index="X" * self.index_length, delimiter="",
                        delimiter_pos=-1, delimiter_length=0,
                        maximum_index_mismatches=self.index_length,
                        maximum_delimiter_mismatches=-1, fingerprint="", fingerprint_pos=0,
                        fingerprint_length=0, maximum_fingerprint_mismatches=0,
                        umi=[], limit_threep_processing=0, has_streams=False)

        out.append(pseudo("Unknown"))
        if self.phix_match_name:
            out.append(pseudo(self.phix_match_name))
        for s in reversed(self.samples):
            d = dict(name=s.name, index=s.index,
                     maximum_index_mismatches=s.maximum_index_mismatch,
                     delimiter=s.delimiter_seq, maximum_delimiter_mismatches=
                     s.maximum_delimiter_mismatch,
                     # memset-0 struct: delimiter_pos stays 0 when the key is absent
                     delimiter_pos=(s.delimiter_start - 1) if s.delimiter_seq is not None else 0,
                     delimiter_length=len(s.delimiter_seq) if s.delimiter_seq is not None else 0,
                     fingerprint=s.fingerprint_seq or "",
                     fingerprint_pos=(s.fingerprint_start - 1) if s.fingerprint_seq else 0,
                     fingerprint_length=len(s.fingerprint_seq) if s.fingerprint_seq else 0,
                     maximum_fingerprint_mismatches=s.maximum_fingerprint_mismatch,
                     umi=[(st - 1, ln) for st, ln in s.umi],
                     limit_threep_processing=s.limit_threep_processing,
                     has_streams=not s.index.startswith("X"))
            out.append(d)
        for i, d in enumerate(out):
            d["numindex"] = i
            dump = self.threep_length
            if d["delimiter"] is not None:
                dump -= d["delimiter_pos"] + d["delimiter_length"]
            if 0 < d["limit_threep_processing"] < dump:
                dump = d["limit_threep_processing"]
            d["signal_dump_length"] = dump
            if d["delimiter_pos"] >= 0:
                d["delimiter_pos"] += ts
            if d["delimiter"] is None:
                d["delimiter"] = ""
        return out

    def to_encode_layout(self, taginfo: bool = True) -> "CEncodeLayout":
        """What tsk_encode_configure() needs besides the parameters: the 5' read, the seqqual
        output length and every sample's files and UMI ranges (parseconfig.c:244-262,491)."""
        lay = CEncodeLayout()
        lay.fivep_start = self.fivep_start - 1
        lay.fivep_length = self.fivep_length
        lay.threep_seqqual_output_length = self.threep_seqqual_output_length
        for i, d in enumerate(self.sample_list()):
            es = lay.samples[i]
            es.want_seqqual = es.want_signal = 1 if d["has_streams"] else 0
            es.want_taginfo = 1 if (d["has_streams"] and taginfo) else 0
            es.umi_count = len(d["umi"])
            for u, (st, ln) in enumerate(d["umi"]):
                es.umi_start[u], es.umi_length[u] = st, ln
        return lay

    def min_bases_passes(self) -> int:
        frac = np.float32(float(self.balancer_minimum_qcpass_percent) * 0.01)
        return int(np.float32(self.balancer_length) * frac)

    def to_cparams(self) -> CParams:
        p = CParams()
        p.abi_version = TSK_ABI_VERSION
        p.total_cycles = self.total_cycles
        p.index_start = self.index_start - 1
        p.index_length = self.index_length
        p.threep_start = self.threep_start - 1
        p.threep_length = self.threep_length
        p.keep_no_delimiter = self.keep_no_delimiter
        p.keep_low_quality_balancer = self.keep_low_quality_balancer
        p.has_control = 1 if self.phix_match_name else 0
        p.control_first_cycle = self.phix_match_start - 1
        p.control_read_length = self.phix_match_length
        p.balancer_start = self.balancer_start - 1
        p.balancer_length = self.balancer_length
        p.balancer_min_quality = self.balancer_minimum_quality
        p.balancer_min_bases_passes = self.min_bases_passes()
        p.balancer_minimum_occurrence = self.balancer_minimum_occurrence
        p.balancer_num_positive_samples = self.balancer_num_positive_samples
        p.balancer_num_negative_samples = self.balancer_num_negative_samples
        p.seed_trigger_polya_length = self.seed_trigger_polya_length
        p.negative_sample_polya_length = self.negative_sample_polya_length
        p.max_cctr_scan_left_space = self.max_cctr_scan_left_space
        p.max_cctr_scan_right_space = self.max_cctr_scan_right_space
        p.required_cdf_contrast = float(np.float32(self.required_cdf_contrast))
        p.polya_boundary_pos = self.polya_boundary_pos
        p.polya_sampling_gap = self.polya_sampling_gap
        p.dist_sampling_bins = self.dist_sampling_bins
        p.fair_sampling_fingerprint_length = self.fair_sampling_fingerprint_length
        p.fair_sampling_hash_space_size = self.fair_sampling_hash_space_size
        p.fair_sampling_max_count = self.fair_sampling_max_count & 0xFF
        for i, b in enumerate("ACGTN"):
            p.weights_polyA[i] = self.polyA_weights[b]
            p.weights_nonA[i] = self.nonA_weights[b]
        p.max_terminal_modifications = self.maximum_modifications
        p.min_polya_length = self.minimum_polya_length
        p.sigproc_trigger_polya_length = self.signal_analysis_trigger
        p.dark_cycles_threshold = float(np.float32(self.dark_cycles_threshold))
        p.max_dark_cycles = self.maximum_dark_cycles
        p.maximum_entropy = float(calculate_maximum_entropy())
        table = precalc_t_intensity_score(self.t_intensity_k, self.t_intensity_center)
        for i in range(TSK_T_SCORE_BINS + 1):
            p.t_intensity_score[i] = float(table[i])
        slist = self.sample_list()
        if len(slist) > TSK_MAX_SAMPLES:
            raise ValueError("too many samples")
        p.num_samples = len(slist)
        for i, d in enumerate(slist):
            s = p.samples[i]
            for key in ("index", "delimiter", "fingerprint"):
                v = d[key].encode("ascii")
                if len(v) >= TSK_MAX_SEQ:
                    raise ValueError("%s too long" % key)
                setattr(s, key, v)
            s.maximum_index_mismatches = d["maximum_index_mismatches"]
            s.delimiter_pos = d["delimiter_pos"]
            s.delimiter_length = d["delimiter_length"]
            s.maximum_delimiter_mismatches = d["maximum_delimiter_mismatches"]
            s.fingerprint_pos = d["fingerprint_pos"]
            s.fingerprint_length = d["fingerprint_length"]
            s.maximum_fingerprint_mismatches = d["maximum_fingerprint_mismatches"]
            s.umi_ranges_count = len(d["umi"])
            s.limit_threep_processing = d["limit_threep_processing"]
            s.signal_dump_length = d["signal_dump_length"]
        return p

    def replace(self, **kw) -> "ImporterConfig":
        return dataclasses.replace(self, **kw)


def _num(v) -> str:
    """YAML scalars are pasted into the INI with str.format (generate-signalproc-conf.py)."""
    return repr(v) if isinstance(v, float) else str(v)


# --------------------------------------------------------------------------------------
# stock configurations (conf/defaults-miseq.conf, defaults-hiseq.conf,
# spikeins-standard.conf, templates/miseq-v2.yaml)
# --------------------------------------------------------------------------------------
SPIKEINS_STANDARD = [   # name, index, trimming length, design poly(A) length
    ("spkN", "ATCACG", 62, 0), ("spk8", "CGATGT", 62, 8), ("spk16", "TTAGGC", 70, 16),
    ("spk32", "TGACCA", 86, 32), ("spk64", "ACAGTG", 118, 64), ("spk118", "CAGATC", 172, 118),
    ("spk128", "ACTTGA", 182, 128),
]
EXPERIMENTAL_MISEQ_V2 = [("Control1", "GCCAAT"), ("Control2", "TGATAC"),
                         ("KD1", "ATGAGC"), ("KD2", "CAGGCG")]


def stock_config(layout: str = "miseq", spikeins: bool = True, **overrides) -> ImporterConfig:
    """The configuration generate-signalproc-conf.py renders for templates/miseq-v2.yaml
    (layout='miseq': R3 = cycles 58-308) or defaults-hiseq.conf (layout='hiseq': 59-309)."""
    if layout == "miseq":
        total, r3 = 308, 58
    elif layout == "hiseq":
        total, r3 = 309, 59
    else:
        raise ValueError(layout)
    samples = []
    # sample sections in the generator's order: experimental, then spike-ins, each sorted by name
    # (generate-signalproc-conf.py:176, configurations.py:121-131)
    for name, index in sorted(EXPERIMENTAL_MISEQ_V2):
        samples.append(SampleConf(name=name, index=index, maximum_index_mismatch=2,
                                  delimiter_seq="GTCAG", delimiter_start=16,
                                  maximum_delimiter_mismatch=1,
                                  umi=[(21, 15), (r3, 15)]))
    if spikeins:
        for name, index, trim, _ in sorted(SPIKEINS_STANDARD):
            samples.append(SampleConf(name=name, index=index, maximum_index_mismatch=3,
                                      delimiter_seq="GTCAG", delimiter_start=11,
                                      maximum_delimiter_mismatch=1,
                                      fingerprint_seq="ACAGTAGCTC", fingerprint_start=r3,
                                      maximum_fingerprint_mismatch=1,
                                      limit_threep_processing=trim))
    cfg = ImporterConfig(total_cycles=total, threep_start=r3, threep_length=251,
                         samples=samples)
    return dataclasses.replace(cfg, **overrides)
