"""Shared helpers for the test-suite: ctypes binding of the CPU oracle (oracle/liboracle.so),
parser for the reference dump format (oracle/dump_shim.c) and parameter defaults.

The oracle is test infrastructure; nothing under yaha_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import gzip
import os
import re
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "yaha")
REF_DUMP_BIN = os.path.join(ORACLE_DIR, "_ref", "yaha_dump")
GOLDEN = os.path.join(ROOT, "tests", "golden")


class YaParams(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("wordLen", "maxHits", "bandWidth", "maxGap", "maxIntron", "minMatch",
                 "GOCost", "GECost", "RCost", "MScore", "XCutoff", "minExtLength")]


class YaFrag(C.Structure):
    _fields_ = [("startRefOff", C.c_uint32), ("startQueryOff", C.c_uint16),
                ("endQueryOff", C.c_uint16), ("hitCount", C.c_uint16), ("refLen", C.c_uint16)]


class YaOp(C.Structure):
    _fields_ = [("length", C.c_uint16), ("opcode", C.c_uint8), ("pad", C.c_uint8)]


FRAG_DT = np.dtype([("startRefOff", "<u4"), ("startQueryOff", "<u2"), ("endQueryOff", "<u2"),
                    ("hitCount", "<u2"), ("refLen", "<u2")])
OP_DT = np.dtype([("length", "<u2"), ("opcode", "u1"), ("pad", "u1")])

KIND_OF_CHAR = {"F": 0, "B": 1, "E": 2, "R": 3}


def default_params(word_len=15, max_hits=650, bw=5, max_gap=50, min_match=25,
                   goc=5, gec=2, rc=3, ms=1, x=25) -> YaParams:
    """Defaults of AlignArgs.c:48-87 with the derived values of AlignArgs.c:108-169."""
    ln, sc, target = 1, 0, min(rc, goc + gec)
    while sc <= target:
        sc += ms
        ln += 1
    return YaParams(word_len, max_hits, bw, max_gap, max_gap, min_match, goc, gec, rc, ms, x, ln)


_oracle = None


def oracle():
    """Load (building if necessary) oracle/liboracle.so."""
    global _oracle
    if _oracle is not None:
        return _oracle
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("oracle_dp.c", "oracle_seed.c", "oracle_clumps.c", "oracle.h")]
    srcs += [os.path.join(ROOT, "yaha_b200", "csrc", h) for h in ("form_clumps.h", "prepare_clumps.h", "assemble_clumps.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
    lib = C.CDLL(so)
    u8p, u32p, i32p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_int)
    lib.orc_dp.restype = C.c_int
    lib.orc_dp.argtypes = [C.POINTER(YaParams), C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.c_uint32,
                           C.c_int, C.c_int, C.c_int, i32p, i32p, C.c_void_p, C.c_int, i32p,
                           C.POINTER(C.c_int64)]
    lib.orc_seed_lookup.restype = C.c_uint32
    lib.orc_seed_lookup.argtypes = [C.POINTER(YaParams), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.orc_find_frags.restype = C.c_int
    lib.orc_find_frags.argtypes = [C.POINTER(YaParams), C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                   C.c_int, C.c_void_p, C.c_int]
    lib.orc_regions.restype = C.c_int
    lib.orc_regions.argtypes = [C.POINTER(YaParams), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.orc_perfect.restype = C.c_int
    lib.orc_perfect.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int]
    lib.orc_encode.restype = None
    lib.orc_encode.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.orc_form_clumps.restype = C.c_int
    lib.orc_form_clumps.argtypes = [C.c_int] * 9 + [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    _oracle = lib
    return lib


def ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def ops_to_str(ops: np.ndarray) -> str:
    if len(ops) == 0:
        return "-"
    return "".join(f"{int(o['length'])}{chr(int(o['opcode']))}" for o in ops)


_OPS_RE = re.compile(r"(\d+)([MRID])")


def str_to_ops(s: str):
    return [] if s == "-" else [(int(n), c) for n, c in _OPS_RE.findall(s)]


def oracle_dp(p: YaParams, bases: np.ndarray, max_roff: int, codes: np.ndarray, kind: int,
              roff: int, rlen: int, qoff: int, qlen: int):
    """Returns (score, addedQ, addedR, ops_string, cells)."""
    lib = oracle()
    cap = 2 * (qlen + rlen + 64) + 8
    ops = np.zeros(cap, dtype=OP_DT)
    aq, ar, n = C.c_int(0), C.c_int(0), C.c_int(0)
    cells = C.c_int64(0)
    score = lib.orc_dp(C.byref(p), ptr(bases), max_roff, ptr(codes), kind, roff, rlen, qoff, qlen,
                       C.byref(aq), C.byref(ar), ptr(ops), cap, C.byref(n), C.byref(cells))
    return score, aq.value, ar.value, ops_to_str(ops[:n.value]), cells.value


def oracle_seed_frags(p: YaParams, so: np.ndarray, roa: np.ndarray, codes: np.ndarray):
    """Stage 1 + 2 for one strand.  Returns (sOffset, count, total, frags, region, keep)."""
    lib = oracle()
    L = len(codes)
    m = L - p.wordLen + 1
    soff = np.zeros(max(m, 1), dtype=np.uint32)
    cnt = np.zeros(max(m, 1), dtype=np.uint32)
    if m <= 0:
        return soff[:0], cnt[:0], 0, np.zeros(0, FRAG_DT), np.zeros(0, np.uint32), np.zeros(0, np.uint8)
    total = lib.orc_seed_lookup(C.byref(p), ptr(so), ptr(codes), L, ptr(soff), ptr(cnt))
    cap = int(total) + 4 * L + 16
    frags = np.zeros(cap, dtype=FRAG_DT)
    nf = 0
    if total:
        nf = lib.orc_find_frags(C.byref(p), ptr(roa), len(roa), ptr(soff), ptr(cnt), m, ptr(frags), cap)
        assert nf >= 0
    frags = frags[:nf]
    region = np.zeros(nf, dtype=np.uint32)
    keep = np.zeros(nf, dtype=np.uint8)
    if nf:
        lib.orc_regions(C.byref(p), ptr(frags), nf, ptr(region), ptr(keep))
    return soff, cnt, int(total), frags, region, keep


def parse_dump(path: str, kinds: str = "DSGC"):
    """Yield parsed records of a dump file written by oracle/dump_shim.c."""
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as f:
        for line in f:
            t = line[0]
            if t not in kinds:
                continue
            w = line.split()
            if t == "D":
                yield ("D", w[1], w[2], int(w[3]), int(w[4]), int(w[5]), int(w[6]), int(w[7]),
                       int(w[8]), int(w[9]), int(w[10]), w[11])
            elif t == "S":
                ent = [tuple(int(x) for x in e.split(":")) for e in w[5:]]
                yield ("S", w[1], int(w[2]), int(w[3]), ent)
            elif t == "G":
                fr = [tuple(int(x) for x in e.split(":")) for e in w[4:]]
                yield ("G", w[1], int(w[2]), fr)
            elif t == "C":
                n = int(w[3])
                i = 4
                clumps = []
                for _ in range(n):
                    rev, nf = int(w[i]), int(w[i + 1])
                    i += 2
                    clumps.append((rev, [tuple(int(x) for x in e.split(":")) for e in w[i:i + nf]]))
                    i += nf
                yield ("C", w[1], int(w[2]), clumps)
