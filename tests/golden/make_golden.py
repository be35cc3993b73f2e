"""Regenerates tests/golden/small/ from the UNMODIFIED reference (oracle/_ref/, built by
oracle/Makefile from /root/reference).  Run in the build container only:

    make -C oracle ref && python tests/golden/make_golden.py

Outputs (all gzip'd, deterministic seeds):
  ref.fa.gz, reads.fa.gz      synthetic 300 kbp / 3-sequence reference and 635 reads incl. the
                              edge cases of yaha_b200.synth.edge_reads
  dump_bw5.txt.gz             every call through the hot-path seams with default flags
  dump_bw10.txt.gz            same with -BW 10 -G 100 (D and G records only)
  out_bw5.sam.gz, out_bw10.sam.gz   the reference's SAM output (-t 1)
  out_{fbs,nooqc,fastq_oss,blast8}.sam.gz   more reference outputs (-FBS Y, -OQC N, FASTQ input with -oss, -o8)
  flag_sweep.json             digest of the reference's SAM for every flag set of tests/hostcases.py FLAG_SWEEP (--only-flag-sweep)
  index_variants.json         sha256 of reference-built indexes for several -L / -S / -H on two references (--only-index-variants)
  files.sha256                digests of the reference-built ref.nib2 and index
"""
import gzip, hashlib, os, shutil, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from yaha_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "small")
REF = os.path.join(ROOT, "oracle", "_ref")


def gz(src, dst):
    with open(src, "rb") as f, gzip.GzipFile(dst, "wb", mtime=0) as g:
        shutil.copyfileobj(f, g)


def chimera_reads(ref, seed, n_reads=60):
    """Reads glued from 2-6 distant loci, each locus piece (60-110 bases either side) with a block of 44-49 substituted bases in its middle: every
    piece becomes one clump whose score dips below zero inside, i.e. a clump that has to be SPLIT -- several per read,
    which is what sends the host through its child-fiber scoring (host/pipeline.cpp runAsChildren)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    out = []
    for r in range(n_reads):
        pieces = []
        for _ in range(int(rng.integers(2, 7))):
            s = int(rng.integers(2000, len(ref) - 3000))
            a, b = int(rng.integers(60, 111)), int(rng.integers(60, 111))
            blk = int(rng.integers(44, 50))
            left = ref[s:s + a]
            mid = synth._BASES[(np.searchsorted(synth._BASES, ref[s + a:s + a + blk]) + rng.integers(1, 4, size=blk)) % 4]
            right = ref[s + a + blk:s + a + blk + b]
            p = np.concatenate([synth.mutate(left, 0.03, rng), mid, synth.mutate(right, 0.03, rng)])
            if rng.integers(0, 3) == 0:
                p = synth._COMP[p[::-1]]
            pieces.append(p)
        out.append((f"chim{r}", np.concatenate(pieces)))
    return out


def only_chimera():
    """Adds chimera.fa.gz / out_chimera.sam.gz without touching the other goldens."""
    tmp = tempfile.mkdtemp()
    ref = synth.random_reference(300_000, 4242)
    bounds = [0, 120_000, 200_003, 300_000]
    synth.write_fasta(tmp + "/ref.fa", [(f"chr{k + 1}", ref[bounds[k]:bounds[k + 1]]) for k in range(3)])
    subprocess.check_call([REF + "/yaha", "-g", "ref.fa", "-L", "11", "-S", "1"], cwd=tmp)
    synth.write_reads(tmp + "/chimera.fa", chimera_reads(ref, 17))
    subprocess.check_call([REF + "/yaha", "-x", "ref.X11_01_65525S", "-q", "chimera.fa", "-osh", "out_chimera.sam", "-t", "1"], cwd=tmp)
    gz(tmp + "/out_chimera.sam", OUT + "/out_chimera.sam.gz")
    gz(tmp + "/chimera.fa", OUT + "/chimera.fa.gz")
    shutil.rmtree(tmp)


def multi_reads(ref, seed, n=220):
    import numpy as np
    """Reads glued from 2-7 clean pieces (0-2 % error: no piece has to be split) so that Optimal Query Coverage has real
    work on reads the device finishes itself (csrc/finish_reads.h): pieces from different loci, strands and sequences,
    pieces that overlap in the query (the tail of one locus repeated at the head of the next: accurate overlap scoring),
    the same locus twice (duplicates), a short piece inside a long one (subsumed), equal-scoring copies (the RNG tie break
    of the sort), secondaries that -FBS keeps."""
    rng = np.random.default_rng(seed)
    L = len(ref)
    out = []
    for i in range(n):
        pieces, tags = [], []
        k = int(rng.integers(2, 8))
        prev = None
        for j in range(k):
            mode = int(rng.integers(0, 6))
            ln = int(rng.integers(40, 260))
            if mode == 0 and prev is not None:                         # same locus again (duplicate / equal keys)
                s, ln2, st = prev
                p = ref[s:s + ln2]
            elif mode == 1 and prev is not None:                       # overlaps the previous piece in the reference
                s0, ln0, st = prev
                s = max(0, s0 + ln0 - int(rng.integers(10, 40))); p = ref[s:s + ln]
            elif mode == 2 and prev is not None:                       # a short piece from inside the previous one
                s0, ln0, st = prev
                ln = max(30, ln0 // 3); s = s0 + int(rng.integers(0, max(1, ln0 - ln))); p = ref[s:s + ln]
            else:
                s = int(rng.integers(0, L - ln)); p = ref[s:s + ln]
            st = int(rng.integers(0, 2))
            prev = (s, len(p), st)
            if st:
                p = synth._COMP[p[::-1]]
            pieces.append(synth.mutate(p, float(rng.choice([0.0, 0.0, 0.01, 0.02])), rng))
            tags.append(f"{s}{'-' if st else '+'}")
        out.append((f"multi{i}_" + "_".join(tags), np.concatenate(pieces)))
    return out


def only_multi():
    """Adds multi.fa.gz / out_multi.sam.gz / out_multi_fbs.sam.gz / out_multi_mno.sam.gz."""
    tmp = tempfile.mkdtemp()
    with gzip.open(OUT + "/ref.fa.gz", "rb") as f, open(tmp + "/ref.fa", "wb") as o:
        o.write(f.read())
    ref = synth.random_reference(300_000, 4242)
    subprocess.check_call([REF + "/yaha", "-g", "ref.fa", "-L", "11", "-S", "1"], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    synth.write_reads(tmp + "/multi.fa", multi_reads(ref, 23))
    for tag, flags in (("multi", []), ("multi_fbs", ["-FBS", "Y", "-PRL", "0.5", "-PSS", "0.5"]), ("multi_mno", ["-MNO", "5", "-BP", "2", "-MGDP", "9", "-M", "15"])):
        subprocess.check_call([REF + "/yaha", "-x", "ref.X11_01_65525S", "-q", "multi.fa", "-osh", f"out_{tag}.sam", "-t", "1"] + flags, cwd=tmp,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        gz(f"{tmp}/out_{tag}.sam", f"{OUT}/out_{tag}.sam.gz")
    gz(tmp + "/multi.fa", OUT + "/multi.fa.gz")
    shutil.rmtree(tmp)


def only_flag_sweep():
    """Writes flag_sweep.json: for every flag set of tests/hostcases.py FLAG_SWEEP the line count and sha256 of the reference's
    SAM (reads.fa, -osh, -t 1) without its @PG line."""
    import json
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import hostcases as H
    tmp = tempfile.mkdtemp()
    for name in ("ref.fa", "reads.fa"):
        with gzip.open(f"{OUT}/{name}.gz", "rb") as f, open(f"{tmp}/{name}", "wb") as o:
            o.write(f.read())
    subprocess.check_call([REF + "/yaha", "-g", "ref.fa", "-L", "11", "-S", "1"], cwd=tmp)
    table = {}
    for flags in H.FLAG_SWEEP:
        subprocess.check_call([REF + "/yaha", "-x", "ref.X11_01_65525S", "-q", "reads.fa", "-osh", "o.sam", "-t", "1"] + flags, cwd=tmp,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        table[" ".join(flags)] = H.digest(H.sam_lines(open(tmp + "/o.sam").read()))
    with open(OUT + "/flag_sweep.json", "w") as f:
        json.dump(table, f, indent=0, sort_keys=True)
        f.write("\n")
    shutil.rmtree(tmp)


INDEX_VARIANTS = [(11, 1, 65525), (11, 3, 65525), (10, 4, 65525), (9, 1, 20), (8, 2, 7), (11, 11, 65525), (11, 5, 3)]   # (-L, -S, -H)


def only_index_variants():
    """Writes index_variants.json: sha256 of the index files the reference builds for synth.n_rich_reference() and for the
    golden reference with several -L / -S / -H (tests/test_formats.py rebuilds them with yaha_b200.refio)."""
    import json
    tmp = tempfile.mkdtemp()
    synth.write_fasta(tmp + "/nrich.fa", synth.n_rich_reference())
    with gzip.open(OUT + "/ref.fa.gz", "rb") as f, open(tmp + "/small.fa", "wb") as o:
        o.write(f.read())
    table = {}
    for stem in ("nrich", "small"):
        for L, S, H in INDEX_VARIANTS:
            subprocess.check_call([REF + "/yaha", "-g", stem + ".fa", "-L", str(L), "-S", str(S), "-H", str(H)], cwd=tmp,
                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            name = f"{stem}.X{L:02d}_{S:02d}_{H:05d}S"
            table[name] = hashlib.sha256(open(f"{tmp}/{name}", "rb").read()).hexdigest()
            os.remove(f"{tmp}/{name}")
        table[stem + ".nib2"] = hashlib.sha256(open(f"{tmp}/{stem}.nib2", "rb").read()).hexdigest()
    with open(OUT + "/index_variants.json", "w") as f:
        json.dump(table, f, indent=0, sort_keys=True)
        f.write("\n")
    shutil.rmtree(tmp)


def only_weird_fastq():
    """Adds weird.fq.gz / out_weird_fastq.sam.gz: FASTQ reader corner cases (Query.c:102-228) -- description after the id and on the
    '+' line, CRLF, multi-line sequence and quality, '@' inside a quality line, quality shorter than the sequence (skipped with the
    reference's warning), lower case, a read shorter than the word length, no final newline."""
    tmp = tempfile.mkdtemp()
    for name in ("ref.fa", "reads.fa"):
        with gzip.open(f"{OUT}/{name}.gz", "rb") as f, open(f"{tmp}/{name}", "wb") as o:
            o.write(f.read())
    subprocess.check_call([REF + "/yaha", "-g", "ref.fa", "-L", "11", "-S", "1"], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    lines = open(tmp + "/reads.fa").read().split(">")[1:8]
    rs = [(l.split("\n", 1)[0], l.split("\n", 1)[1].replace("\n", "")) for l in lines]

    def qual(n, k):
        return "".join(chr(33 + (i * 7 + k) % 40) for i in range(n))
    with open(tmp + "/weird.fq", "w") as f:
        n, s = rs[0]
        f.write(f"@{n} desc with spaces\n{s}\n+{n} again\n{qual(len(s), 1)}\n")
        n, s = rs[1]
        f.write(f"@{n}\r\n{s}\r\n+\r\n{qual(len(s), 2)}\r\n")
        n, s = rs[2]
        h, q = len(s) // 3, qual(len(s), 3)
        f.write(f"@{n}\n{s[:h]}\n{s[h:]}\n+\n{q[:h]}\n{q[h:2 * h]}\n{q[2 * h:]}\n")
        n, s = rs[3]
        q = list(qual(len(s), 4))
        q[50] = "@"
        f.write(f"@{n}\n{s}\n+\n{''.join(q)}\n")
        n, s = rs[4]
        f.write(f"@{n}\n{s}\n+\n{qual(len(s) - 5, 5)}\n")
        n, s = rs[5]
        f.write(f"@{n}\n{s.lower()}\n+\n{qual(len(s), 6)}\n")
        f.write("@tiny\nACGTACG\n+\nIIIIIII\n")
        n, s = rs[6]
        f.write(f"@{n}\n{s}\n+\n{qual(len(s), 7)}")
    subprocess.check_call([REF + "/yaha", "-x", "ref.X11_01_65525S", "-q", "weird.fq", "-oss", "out_weird_fastq.sam", "-t", "1"], cwd=tmp,
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    gz(tmp + "/out_weird_fastq.sam", OUT + "/out_weird_fastq.sam.gz")
    gz(tmp + "/weird.fq", OUT + "/weird.fq.gz")
    shutil.rmtree(tmp)


def only_weird_headers():
    """Adds weird2.fa.gz / out_weird2.sam.gz: FASTA id lines that contain the record markers themselves ('>' '+' '@'
    inside or at the start of the description).  The reference reads an id line to its newline whatever it holds
    (Query.c:111-135); only sequence lines end at the next '>' (Query.c:137-158)."""
    tmp = tempfile.mkdtemp()
    for name in ("ref.fa", "reads.fa"):
        with gzip.open(f"{OUT}/{name}.gz", "rb") as f, open(f"{tmp}/{name}", "wb") as o:
            o.write(f.read())
    subprocess.check_call([REF + "/yaha", "-g", "ref.fa", "-L", "11", "-S", "1"], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    lines = open(tmp + "/reads.fa").read().split(">")[10:17]
    rs = [(l.split("\n", 1)[0], l.split("\n", 1)[1].replace("\n", "")) for l in lines]
    with open(tmp + "/weird2.fa", "w") as f:
        f.write(f">{rs[0][0]}\n{rs[0][1]}\n")
        f.write(f">{rs[1][0]} desc>foo\n{rs[1][1]}\n")
        f.write(f">>{rs[2][0]}\n{rs[2][1][:120]}\n{rs[2][1][120:]}\n")
        f.write(f">{rs[3][0]}+plus @at >\n{rs[3][1]}\n")
        f.write(f">{rs[4][0]}>\r\n{rs[4][1]}\r\n")
        f.write(f">{rs[5][0]}\n{rs[5][1]}\n")
        f.write(f">{rs[6][0]} last>one\n{rs[6][1]}")
    subprocess.check_call([REF + "/yaha", "-x", "ref.X11_01_65525S", "-q", "weird2.fa", "-osh", "out_weird2.sam", "-t", "1"], cwd=tmp,
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    gz(tmp + "/out_weird2.sam", OUT + "/out_weird2.sam.gz")
    gz(tmp + "/weird2.fa", OUT + "/weird2.fa.gz")
    shutil.rmtree(tmp)


def only_empty_first():
    """Adds weird3.fa.gz / weird3.fq.gz and their outputs: an EMPTY record in front.  A record without bases ends the run like
    EOF (the loop of processQueries runs while queryLen > 0, Query.c:306) -- but the first record readNextQuery returns gets a
    second chance (Query.c:304 after main's own first read, Query.c:637), also behind records that were skipped as too short.
    The empty record further down ends the run: the reads behind it are never aligned.  Found by tools/fuzz_parity.py --weird."""
    tmp = tempfile.mkdtemp()
    for name in ("ref.fa", "reads.fa"):
        with gzip.open(f"{OUT}/{name}.gz", "rb") as f, open(f"{tmp}/{name}", "wb") as o:
            o.write(f.read())
    subprocess.check_call([REF + "/yaha", "-g", "ref.fa", "-L", "11", "-S", "1"], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    lines = open(tmp + "/reads.fa").read().split(">")[20:32]
    rs = [(l.split("\n", 1)[0], l.split("\n", 1)[1].replace("\n", "")) for l in lines]
    with open(tmp + "/weird3.fa", "w") as f:
        f.write(f">tooshort\nACGTAC\n")                       # skipped with a warning: not a returned record
        f.write(f">empty_first\n")                            # the first record returned, empty: read once more
        for k in range(0, 6):
            f.write(f">{rs[k][0]}\n{rs[k][1]}\n")
        f.write(f">empty_again\n\n")                          # ends the run
        for k in range(6, 12):
            f.write(f">{rs[k][0]}\n{rs[k][1]}\n")
    with open(tmp + "/weird3.fq", "w") as f:
        f.write("@empty_first\n\n+\n\n")
        for k in range(0, 6):
            f.write(f"@{rs[k][0]}\n{rs[k][1]}\n+\n{'H' * len(rs[k][1])}\n")
        f.write("@empty_again\n\n+\n\n")
        for k in range(6, 12):
            f.write(f"@{rs[k][0]}\n{rs[k][1]}\n+\n{'H' * len(rs[k][1])}\n")
    for q, flag, out in (("weird3.fa", "-osh", "out_weird3.sam"), ("weird3.fq", "-oss", "out_weird3_fastq.sam")):
        subprocess.check_call([REF + "/yaha", "-x", "ref.X11_01_65525S", "-q", q, flag, out, "-t", "1"], cwd=tmp,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        n = sum(1 for l in open(f"{tmp}/{out}") if not l.startswith("@"))
        assert n >= 6, (q, n)
        gz(f"{tmp}/{out}", f"{OUT}/{out}.gz")
        gz(f"{tmp}/{q}", f"{OUT}/{q}.gz")
    shutil.rmtree(tmp)


def main():
    if "--only-empty-first" in sys.argv:
        only_empty_first()
        return
    if "--only-multi" in sys.argv:
        only_multi()
        return
    if "--only-weird-headers" in sys.argv:
        only_weird_headers()
        return
    if "--only-weird-fastq" in sys.argv:
        only_weird_fastq()
        return
    if "--only-index-variants" in sys.argv:
        only_index_variants()
        return
    if "--only-chimera" in sys.argv:
        only_chimera()
        return
    if "--only-flag-sweep" in sys.argv:
        only_flag_sweep()
        return
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp()
    ref = synth.random_reference(300_000, 4242)
    bounds = [0, 120_000, 200_003, 300_000]
    synth.write_fasta(tmp + "/ref.fa", [(f"chr{k + 1}", ref[bounds[k]:bounds[k + 1]]) for k in range(3)])
    reads = list(synth.simulate_reads(ref, 300, 400, 0.05, 99))
    reads += list(synth.simulate_reads(ref, 150, 500, 0.10, 100))
    reads += synth.edge_reads(ref, 7, n_each=25, seq_bounds=bounds)
    synth.write_reads(tmp + "/reads.fa", reads)
    subprocess.check_call([REF + "/yaha", "-g", "ref.fa", "-L", "11", "-S", "1"], cwd=tmp)
    idx = "ref.X11_01_65525S"
    for bw, g, mask in ((5, 50, 15), (10, 100, 5)):
        env = dict(os.environ, YAHA_DUMP=f"{tmp}/dump_bw{bw}.txt", YAHA_DUMP_MASK=str(mask))
        subprocess.check_call([REF + "/yaha_dump", "-x", idx, "-q", "reads.fa", "-osh", f"out_bw{bw}.sam", "-t", "1",
                               "-BW", str(bw), "-G", str(g)], cwd=tmp, env=env)
        gz(f"{tmp}/dump_bw{bw}.txt", f"{OUT}/dump_bw{bw}.txt.gz")
        gz(f"{tmp}/out_bw{bw}.sam", f"{OUT}/out_bw{bw}.sam.gz")
    # extra flag sets and FASTQ input (multi-line records, lower case, a long id with spaces)
    sub = reads[:60] + reads[450:520]
    with open(tmp + "/reads.fq", "w") as f:
        for k, (name, seq) in enumerate(sub):
            s = seq.tobytes().decode()
            if k % 3 == 1:
                s = s.lower()
            q = "".join(chr(33 + (i * 7 + k) % 40) for i in range(len(s)))
            if k % 4 == 2:            # multi-line sequence and quality
                h = len(s) // 2
                f.write(f"@{name} extra words here\n{s[:h]}\n{s[h:]}\n+{name}\n{q[:h]}\n{q[h:]}\n")
            else:
                f.write(f"@{name}\n{s}\n+\n{q}\n")
    for tag, flags, rd in (("fbs", ["-FBS", "Y"], "reads.fa"), ("nooqc", ["-OQC", "N"], "reads.fa"),
                           ("fastq_oss", ["-oss"], "reads.fq"), ("blast8", ["-o8"], "reads.fa")):
        outflag = flags[0] if flags[0] in ("-oss", "-o8") else "-osh"
        extra = [] if flags[0] in ("-oss", "-o8") else flags
        subprocess.check_call([REF + "/yaha", "-x", idx, "-q", rd, outflag, f"out_{tag}.sam", "-t", "1"] + extra, cwd=tmp)
        gz(f"{tmp}/out_{tag}.sam", f"{OUT}/out_{tag}.sam.gz")
    # reader corner cases (Query.c:102-228): long id with spaces, blank line, CRLF, '>' in the middle of a
    # sequence line, a read over 32 000 bases (skipped), lower case, exactly 32 000 bases, no final newline
    rs = [(n, q.tobytes().decode()) for n, q in reads[:5]]
    with open(tmp + "/weird.fa", "w") as f:
        f.write(f">{rs[0][0]} with spaces " + "x" * 250 + "\n" + rs[0][1][:100] + "\n\n" + rs[0][1][100:] + "\n")
        f.write(f">{rs[1][0]}\r\n" + rs[1][1][:150] + "\r\n" + rs[1][1][150:] + "\r\n")
        f.write(f">{rs[2][0]}\n" + rs[2][1][:200] + ">inline_break\n" + rs[2][1][200:] + "\n")
        big = ref[1000:34000].tobytes().decode()
        f.write(">toolong\n" + "\n".join(big[i:i + 70] for i in range(0, len(big), 70)) + "\n")
        f.write(f">{rs[3][0]}\n" + rs[3][1].lower() + "\n")
        f.write(">exact32000\n" + ref[50000:82000].tobytes().decode() + "\n")
        f.write(f">{rs[4][0]}\n" + rs[4][1])
    subprocess.check_call([REF + "/yaha", "-x", idx, "-q", "weird.fa", "-osh", "out_weird.sam", "-t", "1"], cwd=tmp)
    gz(tmp + "/out_weird.sam", OUT + "/out_weird.sam.gz")
    gz(tmp + "/weird.fa", OUT + "/weird.fa.gz")
    gz(tmp + "/reads.fq", OUT + "/reads.fq.gz")
    gz(tmp + "/ref.fa", OUT + "/ref.fa.gz")
    gz(tmp + "/reads.fa", OUT + "/reads.fa.gz")
    with open(OUT + "/files.sha256", "w") as f:
        for name in ("ref.nib2", idx):
            f.write(hashlib.sha256(open(f"{tmp}/{name}", "rb").read()).hexdigest() + "  " + name + "\n")
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
