"""End to end through the drop-in command line on the GPU: `build/test -o index` + `-o quant` against the
reference PROGRAM (oracle/_ref/ref_test) on the same FASTA/FASTQ, including index files exchanged both ways."""
import os
import subprocess

import pytest

import oracle_py
from datasets import dataset
from test_cli_host import TRICKY_FASTQ

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "build", "test")
REF = oracle_py.REF_BIN

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (os.path.exists(OURS) and os.path.exists(REF)), reason="binaries not built")]

MARKERS = ["Index loaded from", "Loading index completed", "Loading read completed", "Sparse chaining completed",
           "EM estimation completed", "Read assignment completed", "Output written to"]


def read_csv(path):
    lines = open(path).read().splitlines()
    assert lines[0] == "Name,NumReads,EM_Abundance"
    return {l.split(",")[0]: (float(l.split(",")[1]), float(l.split(",")[2])) for l in lines[1:]}


def assert_csv_equal(a, b, what=""):
    assert set(a) == set(b), what
    for k in a:
        # both programs print 6 significant digits
        assert a[k][0] == pytest.approx(b[k][0], rel=2e-5), (what, k)
        assert a[k][1] == pytest.approx(b[k][1], rel=2e-5), (what, k)


def write_inputs(tmp_path, d, extra_fastq=b""):
    fa, fq = str(tmp_path / "t.fa"), str(tmp_path / "r.fq")
    with open(fa, "wb") as f:
        for nm, s in zip(d["names"], d["tseqs"]):
            f.write(b">" + nm.encode() + b" gene\n")
            for i in range(0, len(s), 70):
                f.write(s[i:i + 70] + b"\n")
    with open(fq, "wb") as f:
        for i, s in enumerate(d["reads"]):
            f.write(b"@q%d/1\n" % i + s + b"\n+\n" + b"F" * len(s) + b"\n")
        f.write(extra_fastq)
    return fa, fq


@pytest.mark.parametrize("klist", ["31", "21,25,31"])
def test_cli_matches_reference_program(gpu_lib, tmp_path, klist):
    d = dataset(n_genes=60, n_reads=1500, seed=13)
    fa, fq = write_inputs(tmp_path, d, TRICKY_FASTQ)
    p = {n: str(tmp_path / n) for n in ("ours.idx", "ref.idx", "oo.csv", "rr.csv", "or.csv", "ro.csv")}
    o = subprocess.run([OURS, "-k", klist, "-o", "index", fa, p["ours.idx"]], capture_output=True, text=True)
    assert o.returncode == 0 and "Index built in" in o.stdout and "Index saved to " + p["ours.idx"] in o.stdout
    subprocess.run([REF, "-k", klist, "-o", "index", fa, p["ref.idx"]], check=True, capture_output=True)
    # quant ignores -k (main.cpp:174): pass a wrong one on purpose
    runs = [(OURS, "ours.idx", "oo.csv"), (REF, "ref.idx", "rr.csv"), (OURS, "ref.idx", "or.csv"), (REF, "ours.idx", "ro.csv")]
    for exe, idx, csv in runs:
        r = subprocess.run([exe, "-k", "99", "-o", "quant", p[idx], fq, p[csv]], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        pos = [r.stdout.find(m) for m in MARKERS]
        assert all(x >= 0 for x in pos) and pos == sorted(pos), r.stdout
    ref = read_csv(p["rr.csv"])
    assert len(ref) > 20
    for csv in ("oo.csv", "or.csv", "ro.csv"):
        assert_csv_equal(read_csv(p[csv]), ref, csv)


def test_cli_report_and_default_mode(gpu_lib, tmp_path):
    d = dataset(n_genes=30, n_reads=400, seed=5)
    fa, fq = write_inputs(tmp_path, d)
    idx, csv, rep = str(tmp_path / "i.idx"), str(tmp_path / "o.csv"), str(tmp_path / "rep.json")
    subprocess.run([OURS, "-o", "index", fa, idx], check=True, capture_output=True)
    # default mode is quant (main.cpp:214)
    r = subprocess.run([OURS, "--report", rep, idx, fq, csv], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    import json
    j = json.load(open(rep))
    assert j["reads_admitted"] == 400 and j["kernel_launches"] > 0 and j["transcripts"] == len(d["names"])


def test_index_sidecar_cache(gpu_lib, tmp_path):
    d = dataset(n_genes=30, n_reads=500, seed=8)
    fa, fq = write_inputs(tmp_path, d)
    idx = str(tmp_path / "i.idx")
    subprocess.run([OURS, "-k", "21,31", "-o", "index", fa, idx], check=True, capture_output=True)
    out = {}
    for tag, extra in (("plain", []), ("make", ["--index-cache"]), ("use", ["--index-cache"])):
        csv = str(tmp_path / (tag + ".csv"))
        r = subprocess.run([OURS] + extra + ["-o", "quant", idx, fq, csv], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert "Index loaded from " + idx in r.stdout
        out[tag] = open(csv).read()
    assert os.path.exists(idx + ".sqidx")
    assert out["plain"] == out["make"] == out["use"]
    # a changed index file invalidates the sidecar (size/mtime mismatch): rebuild the index with another k
    subprocess.run([OURS, "-k", "25", "-o", "index", fa, idx], check=True, capture_output=True)
    csv2 = str(tmp_path / "k25.csv")
    ref2 = str(tmp_path / "k25_ref.csv")
    subprocess.run([OURS, "--index-cache", "-o", "quant", idx, fq, csv2], check=True, capture_output=True)
    subprocess.run([REF, "-o", "quant", idx, fq, ref2], check=True, capture_output=True)
    assert_csv_equal(read_csv(csv2), read_csv(ref2))


def test_repeated_key_in_index_file_last_wins(gpu_lib, sqb, tmp_path):
    """the reference's loader does mapping[kmer] = vec (src/data_io.cpp:297): of a key that a hand-made index file
    repeats, the last record counts; and duplicate read ids take the exact whole-file path"""
    import numpy as np
    d = dataset(n_genes=30, n_reads=600, seed=21)
    fa, fq = write_inputs(tmp_path, d, TRICKY_FASTQ)
    idx, idx2 = str(tmp_path / "a.idx"), str(tmp_path / "b.idx")
    subprocess.run([OURS, "-k", "31", "-o", "index", fa, idx], check=True, capture_output=True)
    ks, names, seqs, post = sqb.index_io.read_index(idx)
    keys, off, tids = post[31]
    # append every 5th key once more with another posting list (the first transcript only)
    ek, eo, et = list(keys), list(off), list(tids)
    for i in range(0, len(keys), 5):
        ek.append(int(keys[i]))
        et.append(0)
        eo.append(len(et))
    post2 = {31: (np.asarray(ek, dtype=np.uint32), np.asarray(eo, dtype=np.uint64), np.asarray(et, dtype=np.uint32))}
    sqb.index_io.write_index(idx2, ks, names, seqs, post2)
    a, b = str(tmp_path / "ours.csv"), str(tmp_path / "ref.csv")
    rep = str(tmp_path / "rep.json")
    r = subprocess.run([OURS, "--report", rep, "-o", "quant", idx2, fq, a], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    subprocess.run([REF, "-o", "quant", idx2, fq, b], check=True, capture_output=True)
    assert_csv_equal(read_csv(a), read_csv(b))
    import json
    assert json.load(open(rep))["duplicate_id_path"] is True  # TRICKY_FASTQ repeats the id r1


def test_cli_long_reads_chunked_by_bases(gpu_lib, tmp_path):
    """long reads: the ingest pipeline cuts chunks by bases as well as by reads"""
    d = dataset(n_genes=60, n_reads=150, long_reads=(1000, 10000), err=0.05, exon_median=400, seed=9)
    fa, fq = write_inputs(tmp_path, d)
    idx = str(tmp_path / "i.idx")
    subprocess.run([OURS, "-k", "31", "-o", "index", fa, idx], check=True, capture_output=True)
    a, b = str(tmp_path / "ours.csv"), str(tmp_path / "ref.csv")
    r = subprocess.run([OURS, "-o", "quant", idx, fq, a], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    subprocess.run([REF, "-o", "quant", idx, fq, b], check=True, capture_output=True)
    assert_csv_equal(read_csv(a), read_csv(b))
