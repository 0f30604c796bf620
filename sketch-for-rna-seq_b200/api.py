"""Host-side mirror of the reference's functions on the quant path, routed through the CUDA library.

Same names, argument meaning and result shapes as the reference (paths relative to its checkout):
  createSketch_FracMinhash_direct   src/sketch.cpp:24
  build_kmer_to_transcript_map      src/sketch.cpp:51   (as CSR postings)
  process_fastq_single_pass         src/main.cpp:107    (admission + packing; sketching happens on the GPU)
  sparse_chain                      src/sparse_chaining.cpp:29
  estimate_isoform_abundance_em     src/isoform_assignment.cpp:9
  assign_reads_to_isoforms          src/isoform_assignment.cpp:70
  output_to_csv                     src/data_io.cpp:133
  quantification                    src/main.cpp:165
Every function needs a CUDA device; none has a CPU fallback.
"""
import numpy as np

from . import index_io, packing
from .capi import Engine

SKETCH_SIZE = float(np.float32(0.05))  # const float sketch_size = 0.05f, src/main.cpp:43
CHAIN_FRACTION = 0.9                   # src/main.cpp:185
EM_ITERATIONS = 20                     # src/main.cpp:188
EM_TOLERANCE = 0.01


def createSketch_FracMinhash_direct(sequence, k, fraction=SKETCH_SIZE, device=0):
    """set of 32-bit forward ntHash values <= (uint32)(UINT32_MAX*fraction) over all k-mers of `sequence`"""
    seq = sequence.encode() if isinstance(sequence, str) else bytes(sequence)
    if len(seq) < k:
        raise ValueError("sequence length (%d) is smaller than k (%d)" % (len(seq), k))
    words, off, ln = packing.pack_reads([seq])
    with Engine([k], 1, sketch_fraction=fraction, device=device) as e:
        _, hashes = e.sketch(words, off, ln)
    return set(int(h) for h in hashes)


def build_kmer_to_transcript_map(sequences, ks, fraction=SKETCH_SIZE, device=0):
    """sequences: list of bytes, dense ids = positions.  -> dict k -> (keys, off, tids) with sorted distinct ids.
    Sequences shorter than any k get no sketch at all (src/main.cpp:66-75)."""
    idx = [i for i, s in enumerate(sequences) if len(s) >= max(ks)]
    out = {}
    if not idx:
        return {k: (np.zeros(0, np.uint32), np.zeros(1, np.uint64), np.zeros(0, np.uint32)) for k in ks}
    words, off, ln = packing.pack_reads([sequences[i] for i in idx])
    with Engine(ks, len(sequences), sketch_fraction=fraction, device=device) as e:
        for ki, k in enumerate(ks):
            out[k] = e.build_postings(ki, words, off, ln, np.asarray(idx, dtype=np.uint32))
    return out


def process_fastq_single_pass(fastq_file, effective_kmer_lengths):
    """Record scan and admission exactly like src/main.cpp:120-148: a non-empty line starting with '@' opens a
    record (id = rest of the line), the next three lines are sequence, '+', quality; reads with a non-ACGT
    character or shorter than max k are dropped; later duplicates of an id replace earlier ones.
    -> (ids, seqs) of the admitted reads."""
    reads = {}
    with open(fastq_file, "rb") as f:
        lines = f.read().split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()  # getline does not produce a trailing empty line
    i, n = 0, len(lines)
    while i < n:
        line = lines[i]
        i += 1
        if not line or line[:1] != b"@":
            continue
        rid = line[1:]
        seq = lines[i] if i < n else b""
        i += 3
        if not packing.admit(seq, effective_kmer_lengths):
            continue
        reads[rid] = seq
    return list(reads.keys()), list(reads.values())


class Quantifier:
    """The state quantification() threads through its calls, kept on the GPU."""

    def __init__(self, ks, names, postings, sketch_fraction=SKETCH_SIZE, chain_fraction=CHAIN_FRACTION, device=0):
        self.ks = list(ks)
        self.names = list(names)
        self.engine = Engine(self.ks, len(self.names), sketch_fraction=sketch_fraction,
                             chain_fraction=chain_fraction, device=device)
        for ki, k in enumerate(self.ks):
            if k in postings:
                keys, off, tids = postings[k]
                self.engine.load_index(ki, keys, off, tids)

    def close(self):
        self.engine.close()

    def sparse_chain(self, seqs):
        """push admitted read sequences; returns per-read candidate lists [(transcript name, score)] in
        (score desc) order for everything pushed so far"""
        if seqs:
            words, off, ln = packing.pack_reads(seqs)
            self.engine.push_reads(words, off, ln)
        roff, tid, score = self.engine.candidates()
        return [[(self.names[int(tid[j])], int(score[j])) for j in range(int(roff[r]), int(roff[r + 1]))]
                for r in range(len(roff) - 1)]

    def estimate_and_assign(self, max_iterations=EM_ITERATIONS, convergence_threshold=EM_TOLERANCE, R_total=0):
        """-> (pi dict over ALL transcripts, read_counts dict over transcripts with a NumReads entry)"""
        pi, nr, present, iters = self.engine.finish(R_total, max_iterations, convergence_threshold)
        self.iterations = iters
        pi_d = {self.names[i]: float(pi[i]) for i in range(len(self.names))}
        rc_d = {self.names[i]: float(nr[i]) for i in range(len(self.names)) if present[i]}
        return pi_d, rc_d


def sparse_chain(read_seqs, postings, names, kmer_lengths, fraction=CHAIN_FRACTION, sketch_fraction=SKETCH_SIZE):
    q = Quantifier(kmer_lengths, names, postings, sketch_fraction, fraction)
    try:
        return q.sparse_chain(list(read_seqs))
    finally:
        q.close()


def output_to_csv(filename, read_counts, pi, names):
    """src/data_io.cpp:133-152; default ostream double formatting == '%g'"""
    with open(filename, "w") as f:
        f.write("Name,NumReads,EM_Abundance\n")
        for nm in names:
            if nm in read_counts and nm in pi:
                f.write("%s,%g,%g\n" % (nm, read_counts[nm], pi[nm]))


def quantification(index_path, reads_path, output_path, kmer_lengths=None, log=print):
    """src/main.cpp:165-197; the k list always comes from the index (the -k option is ignored in quant mode)"""
    ks, names, _seqs, postings = index_io.read_index(index_path)
    log("Index loaded from " + index_path)
    log("Loading index completed")
    ids, seqs = process_fastq_single_pass(reads_path, ks)
    log("Loading read completed")
    q = Quantifier(ks, names, postings)
    try:
        if seqs:
            words, off, ln = packing.pack_reads(seqs)
            q.engine.push_reads(words, off, ln)
        q.engine.sync()
        log("Sparse chaining completed")
        pi, rc = q.estimate_and_assign()
        log("EM estimation completed")
        log("Read assignment completed")
    finally:
        q.close()
    output_to_csv(output_path, rc, pi, names)
    log("Output written to " + output_path)
    return pi, rc
