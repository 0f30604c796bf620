"""The C++ command-line host (build/test): option handling, messages and exit codes against the reference
program, and the host-side record/admission/index-file logic against the reference's own functions.
None of this needs a GPU."""
import os
import subprocess

import numpy as np
import pytest

import oracle_py
from datasets import SKETCH, dataset

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "build", "test")
REF = oracle_py.REF_BIN

pytestmark = pytest.mark.skipif(not os.path.exists(OURS), reason="build/test not built (python __graft_entry__.py)")
needs_ref = pytest.mark.skipif(not os.path.exists(REF), reason="reference binary oracle/_ref/ref_test not present")


def run(exe, *args):
    p = subprocess.run([exe] + list(args), capture_output=True, text=True)
    return p.returncode, p.stdout.replace(exe, "PROG"), p.stderr.replace(exe, "PROG")


@needs_ref
@pytest.mark.parametrize("args", [["-h"], ["--help"], ["-o", "bogus"], ["-o", "quant", "a", "b"], ["-o", "index", "a"],
                                  ["-k", "21,,31", "-o", "nope"], ["-o", "quant"], ["a", "b"]])
def test_cli_conformance(args):
    assert run(OURS, *args) == run(REF, *args)


@needs_ref
def test_unknown_option_exit_code_and_help():
    rc, out, err = run(OURS, "-z")
    rrc, rout, rerr = run(REF, "-z")
    assert rc == rrc == 1 and out == rout


TRICKY_FASTQ = (b"@r1 first\nACGTACGTACGTACGTACGTACGTACGTACGTACGT\n+\nIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII\n"
                b"\n"                                                 # blank line between records is skipped
                b"junk line that is not a header\n"
                b"@r2\nACGTNACGTACGTACGTACGTACGTACGTACGTACGT\n+\nIIII\n"          # N: refused
                b"@r3\nacgtacgtacgtacgtacgtacgtacgtacgtacgt\n+\nIIII\n"           # lower case: refused
                b"@r4\nACGTACGT\n+\nIIIIIIII\n"                                    # shorter than k: refused
                b"@r5\nTTTTACGTACGTACGTACGTACGTACGTACGTACGTAAAA\n+\n@IIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII\n"  # quality starts with @
                b"@r1 first\nGGGGACGTACGTACGTACGTACGTACGTACGTACGTCCCC\n+\nIIII\n"  # duplicate id: last wins
                b"@r6\nACGTACGTACGTACGTACGTACGTACGTACGTACGT\r\n+\nIIII\n"         # \\r makes it invalid
                b"@r7\nCCCCACGTACGTACGTACGTACGTACGTACGTACGTGGGG")                  # no trailing newline, truncated record


@needs_ref
def test_fastq_admission_matches_reference(tmp_path):
    fq = tmp_path / "t.fq"
    fq.write_bytes(TRICKY_FASTQ)
    rc, out, err = run(OURS, "-k", "21,31", "-o", "selftest-admit", str(fq))
    assert rc == 0
    lines = out.strip().split("\n")
    mine = dict(l.split("\t") for l in lines[1:])
    r = oracle_py.RefOracle([21, 31])
    n = r.fastq(str(fq), SKETCH)
    assert n == len(mine)
    for rid, seq in mine.items():
        sk = r.read_sketch(rid.encode(), 31)
        assert sk is not None, rid
        assert sk.tolist() == oracle_py.RefOracle.sketch_of(seq.encode(), 31, SKETCH).tolist()
    assert set(mine) == {"r1 first", "r5", "r7"}
    assert mine["r1 first"].startswith("GGGG")


@needs_ref
def test_fasta_quirks_and_index_roundtrip(tmp_path, sqb):
    d = dataset()
    fa = tmp_path / "t.fa"
    with open(fa, "wb") as f:
        f.write(b">tA some description\nACGTACGTAC\nGTACGTACGTACGTACGTACGTACGTACGTACGTACGT\n")
        f.write(b">tN\nACGTNNACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT\n")      # dropped: invalid character
        f.write(b">tA dup\nTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTT\n")          # duplicate id: first wins
        f.write(b"\n>tS\nACGTACGT\n")                                              # shorter than k: kept, no sketch
        for nm, s in zip(d["names"][:6], d["tseqs"][:6]):
            f.write(b">" + nm.encode() + b" x\n" + s + b"\n")
        f.write(b">tLast\nACGTACGTACGTNNNNacgtacgtacgtacgtacgtacgtacgtacgtacgtacgtGATTACAGATTACAGATTACAGATTACAGATTACA\n")
    rc, out, err = run(OURS, "-o", "selftest-fasta", str(fa))
    mine = [l.split("\t") for l in out.strip().split("\n")]
    # the reference's view of the same file: build its index and read it back
    idx = tmp_path / "ref.idx"
    subprocess.run([REF, "-k", "21,31", "-o", "index", str(fa), str(idx)], check=True, capture_output=True)
    ks, names, seqs, postings = sqb.index_io.read_index(str(idx))
    assert sorted(n for n, _ in mine) == sorted(names)
    assert dict((n, s.encode()) for n, s in mine) == dict(zip(names, seqs))
    assert "tN" not in names and "tS" in names and "tLast" in names
    # our reader on the reference-written file, and our writer read back by the reference
    rc, out, err = run(OURS, "-o", "selftest-index", str(idx), str(tmp_path / "copy.idx"))
    assert rc == 0 and "ks 21 31" in out and "T %d" % len(names) in out
    r = oracle_py.RefOracle([1])
    assert r.load_index_file(str(tmp_path / "copy.idx")) == [21, 31]
    assert sorted(r.transcript_names()) == sorted(names)
    order = {n: i for i, n in enumerate(r.transcript_names())}
    for k in (21, 31):
        keys, off, tids = r.get_postings(k)
        k0, o0, t0 = postings[k]
        assert keys.tolist() == k0.tolist() and off.tolist() == o0.tolist()
        ref_names = [[r.transcript_names()[t] for t in tids[int(off[i]):int(off[i + 1])]] for i in range(len(keys))]
        my_names = [sorted(names[t] for t in t0[int(o0[i]):int(o0[i + 1])]) for i in range(len(k0))]
        assert [sorted(x) for x in ref_names] == my_names


def test_unopenable_inputs(tmp_path):
    rc, out, err = run(OURS, "-o", "quant", str(tmp_path / "none.idx"), str(tmp_path / "none.fq"), str(tmp_path / "o.csv"))
    assert "Unable to open file for reading" in err and "Loading index completed" in out
    assert rc != 0 and "Could not open FASTQ file" in err  # uncaught runtime_error -> terminate, like upstream


def test_streaming_scan_equals_whole_file_scan(tmp_path):
    """quant mode's streaming scanner + admission + 2-bit packing sees the records the whole-file scan sees"""
    rng = np.random.default_rng(8)
    recs = []
    for i in range(400):
        n = int(rng.integers(20, 300))
        s = bytes(rng.choice(list(b"ACGT"), n).tolist())
        if i % 17 == 0:
            s = s[:5] + b"N" + s[6:]
        if i % 29 == 0:
            s = s.lower()
        recs.append(b"@id%d extra\n" % i + s + b"\n+\n" + b"I" * n + b"\n")
        if i % 50 == 0:
            recs.append(b"\nstray line\n")
    fq = tmp_path / "s.fq"
    fq.write_bytes(b"".join(recs))
    rc, out, _ = run(OURS, "-k", "21,31", "-o", "selftest-admit", str(fq))
    rc2, out2, _ = run(OURS, "-k", "21,31", "-o", "selftest-stream", str(fq))
    assert rc == 0 and rc2 == 0
    whole = [l.split("\t")[1] for l in out.strip().split("\n")[1:]]
    stream = out2.strip().split("\n")
    assert stream[0] == out.split("\n")[0] + " unique 1"
    assert stream[1:] == whole
    # duplicate ids are reported (quant mode then takes the exact whole-file path)
    fq.write_bytes(TRICKY_FASTQ)
    rc3, out3, _ = run(OURS, "-k", "21,31", "-o", "selftest-stream", str(fq))
    assert rc3 == 0 and out3.split("\n")[0].endswith("unique 0")


def test_parallel_segment_scan_equals_sequential_scan(tmp_path):
    """quant mode scans the FASTQ in segments cut at line starts (line counts taken in parallel); on a regular
    four-line file every segmentation gives the records of the sequential scan, an irregular file is reported"""
    rng = np.random.default_rng(21)
    recs = []
    for i in range(700):
        n = int(rng.integers(1, 400))
        s = bytes(rng.choice(list(b"ACGT"), n).tolist())
        if i % 13 == 0:
            s = s[: n // 2] + b"N" + s[n // 2 + 1:]
        q = bytes(rng.choice(list(b"@+IJ#"), n).tolist())  # quality lines that start with '@' or '+'
        recs.append(b"@r%d/1 x\n" % i + s + b"\n+anything\n" + q + b"\n")
    fq = tmp_path / "p.fq"
    for tail in (b"", b"@last\nACGTACGTACGTACGTACGTACGTACGTACGTACGT", b"@cut\n"):  # complete, no final newline, header only
        fq.write_bytes(b"".join(recs) + tail)
        rc, seq_out, _ = run(OURS, "-k", "21,31", "-o", "selftest-stream", str(fq))
        assert rc == 0
        for target in (1, 97, 1000, 50_000, 10_000_000):
            rc2, out2, _ = run(OURS, "-k", "21,31", "-o", "selftest-segments", str(fq), str(target))
            assert rc2 == 0
            first, rest = out2.split("\n", 1)
            assert first.endswith("regular 1"), (target, first)
            assert rest == seq_out, target
    # irregular files: a blank line between records, a header that lost its '@', a three-line record
    for bad in (b"".join(recs[:300]) + b"\n" + b"".join(recs[300:]),
                b"".join(recs[:300]) + recs[300][1:] + b"".join(recs[301:]),
                b"".join(recs[:300]) + b"@x\nACGT\n+\n" + b"".join(recs[300:])):
        fq.write_bytes(bad)
        flagged = 0
        for target in (1000, 10_000_000):
            rc3, out3, _ = run(OURS, "-k", "21,31", "-o", "selftest-segments", str(fq), str(target))
            assert rc3 == 0
            flagged += out3.split("\n")[0].endswith("regular 0")
        assert flagged == 2
