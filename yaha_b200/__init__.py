"""yaha_b200 -- B200-native drop-in for the alignment hot path of yaha 0.1.83.

This package is a thin ctypes binding over the C ABI in ``include/yaha_b200.h``
(``libyaha_b200.so``, hand-written sm_100a CUDA).  There is NO CPU fallback: importing works
anywhere, but creating an :class:`Aligner` requires the compiled library and an sm_100 GPU and
raises otherwise.

Reference seams replaced (file:line under the reference's ``src/``):
  * seed lookup loop                         Query.c:365-412
  * findFragmentsSort / processFragmentsGapped region scan   Math.h:554-555, QueryMatch.c:52-303
  * findAGSAlignment[Banded] / findAGS{Forward,Backward}Extension   Math.h:401-408, SW.cpp:462-1208
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import refio

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libyaha_b200.so")

YA_OK, YA_E_ARG, YA_E_CUDA, YA_E_CAPACITY, YA_E_STATE, YA_E_INTERNAL = 0, 1, 2, 3, 4, 5
DP_FULL, DP_BANDED, DP_EXT_FWD, DP_EXT_BWD = 0, 1, 2, 3


class Params(C.Structure):
    """POD copy of the scoring/seeding fields of AlignmentArgs_t (Math.h:281-304)."""
    _fields_ = [(n, C.c_int32) for n in
                ("wordLen", "maxHits", "bandWidth", "maxGap", "maxIntron", "minMatch",
                 "GOCost", "GECost", "RCost", "MScore", "XCutoff", "minExtLength")]

    @classmethod
    def defaults(cls, word_len=15, max_hits=650, bw=5, max_gap=50, min_match=25,
                 goc=5, gec=2, rc=3, ms=1, x=25) -> "Params":
        """AlignArgs.c:48-87 defaults plus the derived values of AlignArgs.c:108-169."""
        ln, sc, target = 1, 0, min(rc, goc + gec)
        while sc <= target:
            sc += ms
            ln += 1
        return cls(word_len, max_hits, bw, max_gap, max_gap, min_match, goc, gec, rc, ms, x, ln)


FRAG_DT = np.dtype([("startRefOff", "<u4"), ("startQueryOff", "<u2"), ("endQueryOff", "<u2"),
                    ("hitCount", "<u2"), ("refLen", "<u2")])
STRAND_DT = np.dtype([("first", "<u4"), ("n_frags", "<u4"), ("n_frags_all", "<u4"), ("total_hits", "<u4")])
JOB_DT = np.dtype([("rOff", "<u4"), ("read", "<u4"), ("rLen", "<u2"), ("qOff", "<u2"), ("qLen", "<u2"),
                   ("kind", "u1"), ("strand", "u1")])
RES_DT = np.dtype([("score", "<i4"), ("addedQLen", "<u2"), ("addedRLen", "<u2"), ("ops_off", "<u4"), ("ops_n", "<u4")])
OP_DT = np.dtype([("length", "<u2"), ("opcode", "u1"), ("pad", "u1")])


class _ReadBatch(C.Structure):
    _fields_ = [("n_reads", C.c_int32), ("codes", C.c_void_p), ("offsets", C.c_void_p)]


class _FragBatch(C.Structure):
    _fields_ = [("frags_cap", C.c_size_t), ("strands", C.c_void_p), ("frags", C.c_void_p),
                ("region", C.c_void_p), ("n_frags", C.c_size_t), ("frags_needed", C.c_size_t)]


class _ClumpBatch(C.Structure):
    _fields_ = [("maxDesert", C.c_int32), ("minNonOverlap", C.c_int32), ("cap", C.c_size_t), ("clump_first", C.c_void_p),
                ("clump_count", C.c_void_p), ("clumps", C.c_void_p), ("path", C.c_void_p), ("n_clumps", C.c_size_t), ("n_path", C.c_size_t)]


CLUMP_DT = np.dtype([("first", np.uint32), ("n", np.uint16), ("matchedBases", np.uint16)])


class Counters(C.Structure):
    _fields_ = [("probes", C.c_uint64), ("hits", C.c_uint64), ("frags_all", C.c_uint64), ("frags_out", C.c_uint64),
                ("dp_jobs", C.c_uint64), ("dp_cells", C.c_uint64), ("ms_seed", C.c_double), ("ms_dp", C.c_double),
                ("ms_traceback", C.c_double), ("launches", C.c_uint64), ("ext_cells", C.c_uint64), ("ms_ext", C.c_double),
                ("ext_launches", C.c_uint64), ("ms_lookup", C.c_double), ("ms_finish", C.c_double), ("reads_finished", C.c_uint64),
                ("reads_handed_back", C.c_uint64), ("text_bytes", C.c_uint64)]


class OutParams(C.Structure):
    """ya_out_params: the AlignmentArgs_t fields the tail of the per-read path reads (defaults: AlignArgs.c:48-169)."""
    _fields_ = [("maxDesert", C.c_int32), ("minNonOverlap", C.c_int32), ("minRawScore", C.c_int32), ("minIdentity", C.c_float),
                ("OQC", C.c_int32), ("FBS", C.c_int32), ("OQCMinNonOverlap", C.c_int32), ("BPCost", C.c_int32), ("maxBPLog", C.c_int32),
                ("FBS_PSLength", C.c_float), ("FBS_PSScore", C.c_float), ("hardClip", C.c_int32), ("fastq", C.c_int32)]

    @classmethod
    def defaults(cls, min_match=25, max_desert=50, fbs=False, oqc=True, hard_clip=True, fastq=False):
        return cls(max_desert, min_match, min_match, 0.9, int(oqc), int(fbs), min_match, 5, 5, 0.9, 0.9, int(hard_clip), int(fastq))


class _TextBatch(C.Structure):
    _fields_ = [("n_reads", C.c_int32), ("chars", C.c_void_p), ("offsets", C.c_void_p), ("quals", C.c_void_p), ("ids", C.c_void_p),
                ("id_off", C.c_void_p), ("text", C.c_void_p), ("text_cap", C.c_size_t), ("text_off", C.c_void_p), ("status", C.c_void_p),
                ("text_len", C.c_size_t), ("text_needed", C.c_size_t), ("n_handed_back", C.c_int32)]


EXPORTS = ("ya_open", "ya_open_build", "ya_index_sizes", "ya_index_download", "ya_open_peer", "ya_open_shared", "ya_close", "ya_last_error", "ya_set_params", "ya_set_stream",
           "ya_reads_upload", "ya_seed_frags", "ya_form_clumps", "ya_prepare_clumps", "ya_sw_batch", "ya_sw_fetch_ops", "ya_perfect_ext", "ya_host_alloc", "ya_host_free", "ya_get_counters",
           "ya_measure_int32_peak", "ya_measure_gather_peak", "ya_set_output", "ya_align_batch", "ya_align_fetch_text", "ya_peer_direct", "ya_get_ext_intervals", "ya_bind_thread", "ya_set_priority")

_lib = None


def load_library() -> C.CDLL:
    """dlopen libyaha_b200.so (no GPU needed for this) and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(yaha_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    lib.ya_open.restype = vp
    lib.ya_open.argtypes = [C.c_int, C.POINTER(Params), vp, C.c_size_t, vp, C.c_size_t, vp, C.c_size_t, C.c_uint32]
    lib.ya_open_build.restype = vp
    lib.ya_open_build.argtypes = [C.c_int, C.POINTER(Params), vp, C.c_size_t, vp, vp, C.c_int, C.c_uint32, C.c_uint32]
    lib.ya_index_sizes.argtypes = [vp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    lib.ya_index_download.argtypes = [vp, vp, vp]
    lib.ya_open_peer.restype = vp
    lib.ya_open_peer.argtypes = [C.c_int, vp]
    lib.ya_open_shared.restype = vp
    lib.ya_open_shared.argtypes = [vp]
    lib.ya_close.restype = None
    lib.ya_close.argtypes = [vp]
    lib.ya_last_error.restype = C.c_char_p
    lib.ya_last_error.argtypes = [vp]
    lib.ya_set_params.argtypes = [vp, C.POINTER(Params)]
    lib.ya_set_stream.argtypes = [vp, vp]
    lib.ya_reads_upload.argtypes = [vp, C.POINTER(_ReadBatch)]
    lib.ya_seed_frags.argtypes = [vp, C.POINTER(_FragBatch)]
    lib.ya_form_clumps.argtypes = [vp, C.POINTER(_ClumpBatch)]
    lib.ya_sw_batch.argtypes = [vp, vp, C.c_int, vp, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.ya_sw_fetch_ops.argtypes = [vp, vp, C.c_size_t]
    lib.ya_host_alloc.restype = vp
    lib.ya_host_alloc.argtypes = [C.c_size_t]
    lib.ya_host_free.restype = None
    lib.ya_host_free.argtypes = [vp]
    lib.ya_perfect_ext.argtypes = [vp, vp, C.c_int, vp]
    lib.ya_get_counters.argtypes = [vp, C.POINTER(Counters)]
    lib.ya_measure_int32_peak.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.ya_measure_gather_peak.argtypes = [vp, C.POINTER(C.c_double)]
    lib.ya_set_output.argtypes = [vp, C.POINTER(OutParams), C.c_int, C.POINTER(C.c_char_p), vp, vp]
    lib.ya_align_batch.argtypes = [vp, C.POINTER(_TextBatch)]
    lib.ya_align_fetch_text.argtypes = [vp, vp, C.c_size_t]
    lib.ya_peer_direct.argtypes = [vp]
    _lib = lib
    return lib


class YahaError(RuntimeError):
    pass


class Aligner:
    """One GPU context with a resident index (`ya_ctx`)."""

    def __init__(self, nib2: refio.Nib2, index: "refio.Index | None", params: Params | None = None, device: int = 0,
                 peer_of: "Aligner | None" = None, build_max_hits: int = 65525, build_skip: int = 1):
        self.lib = load_library()
        self.nib2, self.index = nib2, index
        if params is None:
            params = Params.defaults(word_len=index.word_len, max_hits=min(650, index.max_hits))
        self.params = params
        if peer_of is not None:
            self.ctx = self.lib.ya_open_peer(device, peer_of.ctx)
        elif index is None:
            # build the index on the device from the .nib2 bases (Index.c:49-331 replacement)
            bases = np.ascontiguousarray(nib2.bases)
            st = np.ascontiguousarray(nib2.starts, dtype=np.uint32)
            ln = np.ascontiguousarray(nib2.lengths, dtype=np.uint32)
            self.ctx = self.lib.ya_open_build(device, C.byref(params), bases.ctypes.data, len(bases),
                                              st.ctypes.data, ln.ctypes.data, len(st), build_max_hits, build_skip)
        else:
            so = np.ascontiguousarray(index.so)
            roa = np.ascontiguousarray(index.roa)
            bases = np.ascontiguousarray(nib2.bases)
            self.ctx = self.lib.ya_open(device, C.byref(params), so.ctypes.data, len(so), roa.ctypes.data, len(roa),
                                        bases.ctypes.data, len(bases), nib2.max_roff)
        if not self.ctx:
            raise YahaError("ya_open failed: " + self.lib.ya_last_error(None).decode())
        self.n_reads = 0

    @classmethod
    def from_files(cls, nib2_path: str, index_path: str, device: int = 0, **param_overrides) -> "Aligner":
        nib, idx = refio.load_nib2(nib2_path), refio.load_index(index_path)
        mh = min(param_overrides.pop("max_hits", 650), idx.max_hits)
        p = Params.defaults(word_len=idx.word_len, max_hits=mh, **param_overrides)
        return cls(nib, idx, p, device)

    def download_index(self, max_hits: int = 65525) -> refio.Index:
        """Copy the device-resident index back (e.g. one built by ya_open_build)."""
        a, b = C.c_size_t(0), C.c_size_t(0)
        self._check(self.lib.ya_index_sizes(self.ctx, C.byref(a), C.byref(b)))
        so = np.zeros(a.value, dtype=np.uint32)
        roa = np.zeros(max(b.value, 1), dtype=np.uint32)
        self._check(self.lib.ya_index_download(self.ctx, so.ctypes.data, roa.ctypes.data))
        return refio.Index(self.params.wordLen, max_hits, int(b.value), so, roa[:b.value])

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.ya_close(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != YA_OK:
            raise YahaError(f"yaha_b200 error {rc}: {self.lib.ya_last_error(self.ctx).decode()}")

    def set_params(self, p: Params):
        self._check(self.lib.ya_set_params(self.ctx, C.byref(p)))
        self.params = p

    def set_stream(self, cuda_stream_ptr: int | None):
        self._check(self.lib.ya_set_stream(self.ctx, cuda_stream_ptr))

    def upload_reads(self, codes: np.ndarray, offsets: np.ndarray):
        """codes: concatenated forward 4-bit codes (uint8); offsets: uint64[n+1]."""
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        b = _ReadBatch(len(offsets) - 1, codes.ctypes.data, offsets.ctypes.data)
        self._check(self.lib.ya_reads_upload(self.ctx, C.byref(b)))
        self.n_reads = len(offsets) - 1

    def upload_read_list(self, reads: list[np.ndarray]):
        offs = np.zeros(len(reads) + 1, dtype=np.uint64)
        if reads:
            offs[1:] = np.cumsum([len(r) for r in reads])
        codes = np.concatenate(reads) if reads else np.zeros(0, np.uint8)
        self.upload_reads(codes, offs)

    def seed_frags(self, frags_cap: int | None = None):
        """Stage 1+2 for the uploaded batch -> (strands[2n], frags, region)."""
        n = self.n_reads
        strands = np.zeros(2 * n, dtype=STRAND_DT)
        cap = frags_cap if frags_cap is not None else max(1024, 64 * n)
        while True:
            frags = np.zeros(cap, dtype=FRAG_DT)
            region = np.zeros(cap, dtype=np.uint32)
            fb = _FragBatch(cap, strands.ctypes.data, frags.ctypes.data, region.ctypes.data, 0, 0)
            rc = self.lib.ya_seed_frags(self.ctx, C.byref(fb))
            if rc == YA_E_CAPACITY:
                cap = int(fb.frags_needed) + 16
                continue
            self._check(rc)
            return strands, frags[:fb.n_frags], region[:fb.n_frags]

    def form_clumps(self, n_frags: int, max_desert: int = 50, min_non_overlap: int = 25):
        """Row N1 for the batch seed_frags() has just processed -> (clump_first[2n], clump_count[2n], clumps, path)."""
        n = self.n_reads
        cap = max(16, int(n_frags))
        first = np.zeros(2 * n, dtype=np.uint32)
        count = np.zeros(2 * n, dtype=np.uint32)
        clumps = np.zeros(cap, dtype=CLUMP_DT)
        path = np.zeros(cap, dtype=FRAG_DT)
        cb = _ClumpBatch(max_desert, min_non_overlap, cap, first.ctypes.data, count.ctypes.data, clumps.ctypes.data, path.ctypes.data, 0, 0)
        self._check(self.lib.ya_form_clumps(self.ctx, C.byref(cb)))
        return first, count, clumps, path

    def sw_batch(self, jobs: np.ndarray, ops_cap: int | None = None):
        """Stage 3 for a structured array of JOB_DT -> (results RES_DT[n], ops OP_DT[total])."""
        jobs = np.ascontiguousarray(jobs, dtype=JOB_DT)
        n = len(jobs)
        res = np.zeros(n, dtype=RES_DT)
        cap = ops_cap if ops_cap is not None else max(1024, 16 * n)
        while True:
            ops = np.zeros(cap, dtype=OP_DT)
            need = C.c_size_t(0)
            rc = self.lib.ya_sw_batch(self.ctx, jobs.ctypes.data, n, res.ctypes.data, ops.ctypes.data, cap, C.byref(need))
            if rc == YA_E_CAPACITY:               # res is complete; fetch the ops the device kept (no DP repeated)
                ops = np.zeros(int(need.value), dtype=OP_DT)
                rc = self.lib.ya_sw_fetch_ops(self.ctx, ops.ctypes.data, len(ops))
            self._check(rc)
            return res, ops[:need.value]

    def set_output(self, out: "OutParams | None" = None):
        """ya_set_output: output flags + the sequence table of the loaded .nib2 (names, starts, lengths)."""
        out = out or OutParams.defaults(min_match=self.params.minMatch)
        names = (C.c_char_p * len(self.nib2.names))(*[n.encode() if isinstance(n, str) else bytes(n) for n in self.nib2.names])
        st = np.ascontiguousarray(self.nib2.starts, dtype=np.uint32)
        ln = np.ascontiguousarray(self.nib2.lengths, dtype=np.uint32)
        self._check(self.lib.ya_set_output(self.ctx, C.byref(out), len(st), names, st.ctypes.data, ln.ctypes.data))

    def align_batch(self, reads: "list[tuple[str, bytes]]", quals: "list[bytes] | None" = None):
        """ya_align_batch: reads as (id, characters) -> (SAM text of the reads finished on the device as bytes, per-read text
        offsets uint64[n+1], status uint8[n]: 1 = handed back)."""
        n = len(reads)
        offs = np.zeros(n + 1, dtype=np.uint64)
        idoff = np.zeros(n + 1, dtype=np.uint32)
        if n:
            offs[1:] = np.cumsum([len(s) for _, s in reads])
            idoff[1:] = np.cumsum([len(i) for i, _ in reads])
        chars = np.frombuffer(b"".join(bytes(s) for _, s in reads) + b"\0", dtype=np.uint8).copy()
        ids = np.frombuffer("".join(i for i, _ in reads).encode() + b"\0", dtype=np.uint8).copy()
        q = np.frombuffer(b"".join(quals) + b"\0", dtype=np.uint8).copy() if quals is not None else None
        cap = 4 * int(offs[-1]) + 1024 * n + 4096
        text = np.zeros(cap, dtype=np.uint8)
        toff = np.zeros(n + 1, dtype=np.uint64)
        status = np.zeros(max(n, 1), dtype=np.uint8)
        tb = _TextBatch(n, chars.ctypes.data, offs.ctypes.data, q.ctypes.data if q is not None else None, ids.ctypes.data,
                        idoff.ctypes.data, text.ctypes.data, cap, toff.ctypes.data, status.ctypes.data, 0, 0, 0)
        rc = self.lib.ya_align_batch(self.ctx, C.byref(tb))
        if rc == YA_E_CAPACITY:
            text = np.zeros(int(tb.text_needed) + 16, dtype=np.uint8)
            rc = self.lib.ya_align_fetch_text(self.ctx, text.ctypes.data, len(text))
        self._check(rc)
        self.n_reads = n
        return text[:tb.text_len].tobytes(), toff, status[:n]

    def perfect_ext(self, jobs: np.ndarray) -> np.ndarray:
        jobs = np.ascontiguousarray(jobs, dtype=JOB_DT)
        out = np.zeros(len(jobs), dtype=np.uint16)
        self._check(self.lib.ya_perfect_ext(self.ctx, jobs.ctypes.data, len(jobs), out.ctypes.data))
        return out

    def counters(self) -> Counters:
        c = Counters()
        self._check(self.lib.ya_get_counters(self.ctx, C.byref(c)))
        return c

    def gather_peak(self) -> float:
        g = C.c_double(0)
        self._check(self.lib.ya_measure_gather_peak(self.ctx, C.byref(g)))
        return g.value

    def int32_peak(self) -> tuple[float, float]:
        a, m = C.c_double(0), C.c_double(0)
        self._check(self.lib.ya_measure_int32_peak(self.ctx, C.byref(a), C.byref(m)))
        return a.value, m.value
