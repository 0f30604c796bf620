"""Seeded synthetic transcriptomes and reads of the shapes named in BASELINE.json / SURVEY.md section 8(d).

Written with torch tensor ops so the same code makes the small CPU fixtures of the tests and, on a CUDA
device, the human-scale bench workload in seconds.  This is data generation, not the product path.

Everything is forward-strand (the reference hashes the forward strand only, src/sketch.cpp:33) and errors
are substitutions, so every read stays ACGT-only and is admitted.
"""
import math

import numpy as np
import torch


def _gen(seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def _segment_starts(lengths):
    off = torch.zeros(lengths.numel() + 1, dtype=torch.int64, device=lengths.device)
    off[1:] = torch.cumsum(lengths, 0)
    return off


def _ragged_arange(lengths, off=None):
    """for segments of the given lengths: (segment id per element, position inside the segment)"""
    if off is None:
        off = _segment_starts(lengths)
    seg = torch.repeat_interleave(torch.arange(lengths.numel(), device=lengths.device), lengths)
    within = torch.arange(int(off[-1]), device=lengths.device) - off[:-1][seg]
    return seg, within


def make_transcriptome(n_genes, seed=7, device="cpu", mean_isoforms=4.0, exons=(8, 12), exon_median=150,
                       exon_sigma=0.8, exon_min=20, exon_cap=5000, keep_p=0.8):
    """Genes made of random exons; isoforms are exon subsets, so k-mer sharing looks like a transcriptome
    (mean posting degree 2-5).  Returns dict(codes uint8[total], t_off int64[T+1], gene int64[T])."""
    dev = torch.device(device)
    g = _gen(seed, dev)
    G = int(n_genes)
    n_ex = torch.randint(exons[0], exons[1] + 1, (G,), generator=g, device=dev)
    ex_start = _segment_starts(n_ex)
    NE = int(ex_start[-1])
    ex_len = torch.exp(math.log(exon_median) + exon_sigma * torch.randn(NE, generator=g, device=dev))
    ex_len = ex_len.round().clamp(exon_min, exon_cap).to(torch.int64)
    ex_off = _segment_starts(ex_len)
    pool = torch.randint(0, 4, (int(ex_off[-1]),), generator=g, device=dev, dtype=torch.uint8)
    # isoforms per gene: 1 + geometric with mean (mean_isoforms-1)
    p = 1.0 / max(mean_isoforms, 1.0)
    u = torch.rand(G, generator=g, device=dev).clamp_min(1e-12)
    n_iso = 1 + torch.floor(torch.log(u) / math.log(1.0 - p + 1e-12)).to(torch.int64).clamp(0, 40)
    iso_gene = torch.repeat_interleave(torch.arange(G, device=dev), n_iso)
    T = iso_gene.numel()
    # (isoform, exon slot) grid, each exon kept with keep_p, slot 0 always kept
    iso_nex = n_ex[iso_gene]
    iso_of, slot = _ragged_arange(iso_nex)
    keep = torch.rand(iso_of.numel(), generator=g, device=dev) < keep_p
    keep |= slot == 0
    iso_k = iso_of[keep]
    exon_k = ex_start[:-1][iso_gene[iso_k]] + slot[keep]
    seg_len = ex_len[exon_k]
    t_len = torch.zeros(T, dtype=torch.int64, device=dev).index_add_(0, iso_k, seg_len)
    t_off = _segment_starts(t_len)
    seg, within = _ragged_arange(seg_len)
    codes = pool[ex_off[:-1][exon_k][seg] + within]
    return {"codes": codes, "t_off": t_off, "gene": iso_gene}


def transcript_names(T, gene=None):
    """ids in the style of SURVEY 8(d): ENST%07d.%d|G%05d (<= 32 chars)"""
    if gene is None:
        return ["ENST%07d.1" % i for i in range(T)]
    gene = gene.tolist() if hasattr(gene, "tolist") else list(gene)
    return ["ENST%07d.%d|G%05d" % (i, 1 + i % 3, gene[i]) for i in range(T)]


def simulate_reads(tx, n_reads, read_len=150, seed=11, err=0.005, long_reads=None, expr_sigma=1.5, chunk=1 << 21):
    """Forward-strand reads with substitution errors.  read_len: fixed length, or long_reads=(lo, hi) for
    log-uniform lengths clipped to the transcript.  Expression is log-normal per transcript.
    Yields dicts(codes uint8[total], r_off int64[n+1], tid int64[n]) in chunks."""
    codes, t_off = tx["codes"], tx["t_off"]
    dev = codes.device
    g = _gen(seed, dev)
    t_len = t_off[1:] - t_off[:-1]
    T = t_len.numel()
    min_len = read_len if long_reads is None else long_reads[0]
    w = torch.exp(expr_sigma * torch.randn(T, generator=g, device=dev, dtype=torch.float64))
    w = torch.where(t_len >= min_len, w, torch.zeros_like(w))
    if float(w.sum()) <= 0:
        raise ValueError("no transcript is long enough for the requested reads")
    cdf = torch.cumsum(w / w.sum(), 0)
    done = 0
    while done < n_reads:
        n = min(chunk, n_reads - done)
        u = torch.rand(n, generator=g, device=dev, dtype=torch.float64)
        t = torch.searchsorted(cdf, u).clamp(max=T - 1)
        # searchsorted may land on a zero-weight transcript at a cdf plateau edge: move to the next eligible one
        bad = t_len[t] < min_len
        if bad.any():
            elig = torch.nonzero(t_len >= min_len).flatten()
            t[bad] = elig[torch.searchsorted(elig, t[bad]).clamp(max=elig.numel() - 1)]
        if long_reads is None:
            rl = torch.full((n,), read_len, dtype=torch.int64, device=dev)
        else:
            lo, hi = long_reads
            rl = torch.exp(math.log(lo) + (math.log(hi) - math.log(lo)) *
                           torch.rand(n, generator=g, device=dev)).to(torch.int64)
            rl = torch.minimum(rl, t_len[t])
        start = (torch.rand(n, generator=g, device=dev, dtype=torch.float64) *
                 (t_len[t] - rl + 1).to(torch.float64)).to(torch.int64)
        start = torch.minimum(start, t_len[t] - rl)
        r_off = _segment_starts(rl)
        seg, within = _ragged_arange(rl, r_off)
        rc = codes[t_off[:-1][t][seg] + start[seg] + within]
        if err > 0:
            m = torch.rand(rc.numel(), generator=g, device=dev) < err
            sub = torch.randint(1, 4, (rc.numel(),), generator=g, device=dev, dtype=torch.uint8)
            rc = torch.where(m, (rc + sub) & 3, rc)
        yield {"codes": rc, "r_off": r_off, "tid": t}
        done += n


def pack_ragged(codes, off, align=4):
    """uint8 codes of segments [off[i], off[i+1]) -> (words int32-as-uint32 tensor, base_off uint32, len uint32),
    each segment starting at the next multiple of `align` bases.  Word count padded to a multiple of 4 (+4)."""
    dev = codes.device
    ln = off[1:] - off[:-1]
    n = ln.numel()
    padded = (ln + (align - 1)) // align * align
    base_off = torch.zeros(n, dtype=torch.int64, device=dev)
    if n > 1:
        base_off[1:] = torch.cumsum(padded[:-1], 0)
    total = int(base_off[-1] + ln[-1]) if n else 0
    n_words = ((total + 15) // 16 + 3) // 4 * 4 + 4
    if bool((padded == ln).all()):
        flat = torch.zeros(n_words * 16, dtype=torch.uint8, device=dev)
        flat[:codes.numel()] = codes
    else:
        flat = torch.zeros(n_words * 16, dtype=torch.uint8, device=dev)
        seg, within = _ragged_arange(ln, off)
        flat[base_off[seg] + within] = codes
    words = torch.zeros(n_words, dtype=torch.int64, device=dev)
    f = flat.view(n_words, 16).to(torch.int64)
    for j in range(16):
        words |= f[:, j] << (2 * j)
    words = words.to(torch.int32) if False else (words & 0xFFFFFFFF)
    # store as int32 bit pattern (torch has no uint32 arithmetic); numpy view gives uint32
    words = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)
    return words, base_off.to(torch.int32), ln.to(torch.int32)


def to_u32(t):
    """int32 torch tensor -> uint32 numpy view (host copy)"""
    return t.detach().cpu().contiguous().numpy().view(np.uint32)


def codes_to_ascii(codes):
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=codes.device)
    return lut[codes.to(torch.int64)].cpu().numpy().tobytes()


def write_fasta(path, names, tx, desc=" synthetic transcript", width=0):
    seq = codes_to_ascii(tx["codes"])
    off = tx["t_off"].tolist()
    with open(path, "wb") as f:
        for i, nm in enumerate(names):
            f.write(b">" + nm.encode() + desc.encode() + b"\n")
            s = seq[off[i]:off[i + 1]]
            if width:
                for j in range(0, len(s), width):
                    f.write(s[j:j + width] + b"\n")
            else:
                f.write(s + b"\n")


def write_fastq(path, reads, first_id=0, prefix="r", mode="wb"):
    """reads: one chunk dict from simulate_reads"""
    seq = codes_to_ascii(reads["codes"])
    off = reads["r_off"].tolist()
    with open(path, mode) as f:
        for i in range(len(off) - 1):
            s = seq[off[i]:off[i + 1]]
            f.write(b"@" + prefix.encode() + str(first_id + i).encode() + b"\n" + s + b"\n+\n" + b"I" * len(s) + b"\n")
    return len(off) - 1
