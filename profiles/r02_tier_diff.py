"""Diagnostic: candidate lists of the config-2 workload under the three vote tiers (automatic, window kernel for
every read, general kernel for every read) must be identical; prints the first reads where they are not."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from _sqpkg import sqb
syn = sqb.synth
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
ks = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [31]
scale = float(np.float32(float(sys.argv[3]) if len(sys.argv) > 3 else 0.05))
long_reads = len(sys.argv) > 4 and sys.argv[4] == "long"
tx = syn.make_transcriptome(62500, seed=7, device="cuda:0")
T = tx["t_off"].numel() - 1
tlen = tx["t_off"][1:] - tx["t_off"][:-1]
keep = torch.nonzero(tlen >= max(ks)).flatten()
words, boff, ln = syn.pack_ragged(tx["codes"], tx["t_off"], align=4)
chunks = []
sim = dict(long_reads=(1000, 10000), err=0.05, chunk=1 << 16) if long_reads else dict(read_len=150, err=0.005, chunk=1 << 21)
for ch in syn.simulate_reads(tx, n_reads, seed=1000, **sim):
    w, b, l = syn.pack_ragged(ch["codes"], ch["r_off"], align=4)
    chunks.append((w, b, l, int(ch["r_off"][-1])))
res = {}
post = None
for tier in (2, 0, 1):
    e = sqb.Engine(ks, T, sketch_fraction=scale)
    if post is None:
        post = {}
        for ki, k in enumerate(ks):
            post[k] = e.build_postings(ki, syn.to_u32(words), syn.to_u32(boff[keep].contiguous()), syn.to_u32(ln[keep].contiguous()),
                                       keep.cpu().numpy().astype(np.uint32))
    for ki, k in enumerate(ks):
        e.load_index(ki, *post[k])
    e.set_option("vote_tier", tier)
    for w, b, l, nb in chunks:
        e.push_reads_device(w.data_ptr(), w.numel(), b.data_ptr(), l.data_ptr(), l.numel(), nb + 4 * l.numel())
        e.sync()
    res[tier] = e.candidates()
    st = e.stats()
    print("tier", tier, "pairs", st["pairs"], "mid", st["mid_reads"], "slow", st["slow_reads"], flush=True)
    e.close()
off2, tid2, sc2 = res[2]
for tier in (0, 1):
    off, tid, sc = res[tier]
    cnt, cnt2 = np.diff(off.astype(np.int64)), np.diff(off2.astype(np.int64))
    bad = np.nonzero(cnt != cnt2)[0]
    same_len = np.nonzero(cnt == cnt2)[0]
    if len(bad) == 0 and np.array_equal(tid, tid2) and np.array_equal(sc, sc2):
        print("tier", tier, "== general: identical")
        continue
    # reads with equal counts but different content
    if np.array_equal(off, off2):
        diff = np.nonzero((tid != tid2) | (sc != sc2))[0]
        rd = np.unique(np.searchsorted(off, diff, side="right") - 1)
    else:
        rd = bad
    print("tier", tier, "differs from general in", len(rd), "reads (count mismatches: %d)" % len(bad))
    for r in rd[:12]:
        a, b = int(off[r]), int(off[r + 1])
        a2, b2 = int(off2[r]), int(off2[r + 1])
        print("  read", int(r), "tier:", list(zip(tid[a:b].tolist(), sc[a:b].tolist())), " general:", list(zip(tid2[a2:b2].tolist(), sc2[a2:b2].tolist())))
