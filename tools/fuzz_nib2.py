#!/usr/bin/env python
"""Differential fuzzing of `-g` (genome compression, host/indexer.cpp) against the unmodified reference's compressFile
(Compress.c:140-331): random FASTA files with varying line widths, blank lines, CRLF, lower case, IUPAC codes, N runs, stray
characters, descriptions, missing final newline; the .nib2 files must be byte-identical.  CPU only (the host program on the mock
of the ABI writes the .nib2 before it asks the device for the index).   usage: tools/fuzz_nib2.py --seeds 0:300"""
import argparse
import hashlib
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "yaha")
MOCK = os.path.join(ROOT, "tests", "_build", "yaha_host_mock")


def draw_fasta(rng):
    out = bytearray()
    nseq = int(rng.integers(1, 6))
    alphabet = np.frombuffer(b"ACGT", dtype=np.uint8)
    odd = np.frombuffer(b"NNNNnRYKMSWBDHVXacgt.-*", dtype=np.uint8)
    eol = b"\r\n" if rng.integers(0, 6) == 0 else b"\n"
    for k in range(nseq):
        name = f"seq{k}".encode()
        if rng.integers(0, 2):
            name += b" some description " + str(int(rng.integers(0, 1000))).encode()
        if rng.integers(0, 8) == 0:
            name += b"\t>x@+"
        out += b">" + name + eol
        n = int(rng.choice([0, 1, 7, 60, 61, 500, 5000, 20000])) if rng.integers(0, 4) == 0 else int(rng.integers(50, 30000))
        seq = alphabet[rng.integers(0, 4, size=n)].copy()
        for _ in range(int(rng.integers(0, 4))):
            if n > 10:
                a = int(rng.integers(0, n - 1)); b = min(n, a + int(rng.integers(1, 300)))
                seq[a:b] = ord("N") if rng.integers(0, 2) else ord("n")
        if rng.integers(0, 3) == 0 and n:
            k2 = int(rng.integers(1, 20))
            seq[rng.integers(0, n, size=k2)] = odd[rng.integers(0, len(odd), size=k2)]
        if rng.integers(0, 4) == 0 and n:
            a = int(rng.integers(0, n)); b = int(rng.integers(a, n + 1))
            s2 = seq[a:b]; up = (s2 >= 65) & (s2 <= 90); s2[up] += 32
        width = int(rng.choice([50, 60, 70, 80, 1000]))
        pos = 0
        while pos < n:
            w = width if rng.integers(0, 10) else int(rng.integers(1, 200))
            out += bytes(seq[pos:pos + w]) + eol
            pos += w
            if rng.integers(0, 40) == 0:
                out += eol                                   # blank line
    if rng.integers(0, 5) == 0 and out.endswith(eol) and n > 0:
        del out[-len(eol):]                                  # no final newline (never behind an id line: the reference then
                                                             # searches for the newline past the end of its mapping, Compress.c:272)
    return bytes(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", default="0:100")
    ap.add_argument("--keep", default="/tmp/fuzz_nib2_fail")
    a = ap.parse_args()
    lo, hi = (int(x) for x in a.seeds.split(":"))
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "mock"), "SAN="])
    nbad = ncmp = 0
    for seed in range(lo, hi):
        rng = np.random.default_rng(seed)
        fa = draw_fasta(rng)
        tmp = tempfile.mkdtemp(prefix=f"nib{seed}_")
        try:
            for side in ("want", "got"):
                os.mkdir(f"{tmp}/{side}")
                open(f"{tmp}/{side}/g.fa", "wb").write(fa)
            r = subprocess.run([REF, "-g", "g.fa", "-L", "11"], cwd=tmp + "/want", capture_output=True, text=True, timeout=300)
            m = subprocess.run([MOCK, "-g", "g.fa", "-L", "11"], cwd=tmp + "/got", capture_output=True, text=True, timeout=300)
            wantf, gotf = tmp + "/want/g.nib2", tmp + "/got/g.nib2"
            finished = "Finished compressing" in (r.stdout + r.stderr)      # (the reference may still crash while it forms the index)
            if not os.path.exists(wantf) or (r.returncode != 0 and not finished):
                status = "reference failed rc=%d %r" % (r.returncode, r.stderr.strip()[-120:])
                # the host must not succeed silently where the reference refuses
                print("skip", seed, status, "| host rc", m.returncode, repr(m.stderr.strip()[-120:]), flush=True)
                continue
            ncmp += 1
            ok = os.path.exists(gotf) and hashlib.sha256(open(gotf, "rb").read()).digest() == hashlib.sha256(open(wantf, "rb").read()).digest()
            print("ok  " if ok else "FAIL", seed, len(fa), "bytes", flush=True)
            if not ok:
                nbad += 1
                print("   host stderr:", m.stderr.strip()[-300:])
                os.makedirs(a.keep, exist_ok=True)
                shutil.copy(tmp + "/want/g.fa", f"{a.keep}/seed{seed}.fa")
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    print(f"{hi - lo} cases, {ncmp} compared, {nbad} failed")
    return 1 if nbad else 0


if __name__ == "__main__":
    sys.exit(main())
