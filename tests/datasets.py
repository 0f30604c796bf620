"""Small seeded datasets shared by the tests (made with the package's torch generators on the CPU)."""
import functools

import numpy as np

from _sqpkg import sqb

SKETCH = float(np.float32(0.05))


def _split(codes, off):
    s = sqb.synth.codes_to_ascii(codes)
    off = off.tolist()
    return [s[off[i]:off[i + 1]] for i in range(len(off) - 1)]


@functools.lru_cache(maxsize=None)
def dataset(n_genes=40, n_reads=400, read_len=150, seed=7, err=0.005, long_reads=None, exon_median=150):
    tx = sqb.synth.make_transcriptome(n_genes, seed=seed, exon_median=exon_median)
    tseqs = _split(tx["codes"], tx["t_off"])
    names = sqb.synth.transcript_names(len(tseqs), tx["gene"])
    reads, tids = [], []
    for ch in sqb.synth.simulate_reads(tx, n_reads, read_len, seed=seed + 1, err=err, long_reads=long_reads):
        reads += _split(ch["codes"], ch["r_off"])
        tids += ch["tid"].tolist()
    return {"tseqs": tseqs, "names": names, "reads": reads, "read_tid": tids}


def csr_to_lists(off, tid, score):
    return [sorted(zip(tid[int(off[r]):int(off[r + 1])].tolist(), score[int(off[r]):int(off[r + 1])].tolist()))
            for r in range(len(off) - 1)]
