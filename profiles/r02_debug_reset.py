"""Diagnostic: results after sq_reset_reads + a second pass must equal those of a fresh engine."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
from _sqpkg import sqb
import oracle_py
from datasets import dataset, SKETCH
d = dataset(n_genes=60, n_reads=1500, seed=13)
port = oracle_py.PortOracle()
ks = [31]
thr = port.threshold(SKETCH)
post = port.postings_from_sequences(d["tseqs"], ks, thr)
T = len(d["names"])
def run(eng, reads):
    eng.push_reads(*sqb.packing.pack_reads(reads))
    off, tid, sc = eng.candidates()
    pi, nr, pr, it = eng.finish(0, 20, 0.01)
    return off, tid, sc, pi, nr, pr, eng.stats()
e = sqb.Engine(ks, T, sketch_fraction=SKETCH)
e.load_index(0, *post[31])
extra = [b"ACGT" * 10, b"GGGG" + b"ACGT" * 8 + b"CCCC"]
a = run(e, d["reads"] + extra)
e.reset_reads()
b = run(e, d["reads"])
e.close()
f = sqb.Engine(ks, T, sketch_fraction=SKETCH)
f.load_index(0, *post[31])
c = run(f, d["reads"])
f.close()
print("after reset vs fresh: cand equal", np.array_equal(b[0], c[0]) and np.array_equal(b[1], c[1]) and np.array_equal(b[2], c[2]),
      "pi max rel", float(np.max(np.abs(b[3] - c[3]) / c[3])), "nr max abs", float(np.max(np.abs(b[4] - c[4]))),
      "classes", b[6]["em_classes"], c[6]["em_classes"], "class pairs", b[6]["em_class_pairs"], c[6]["em_class_pairs"])
_, ooff, otid, osc, R = port.chain_batch(ks, thr, 0.9, post, d["reads"])
opi, _ = port.em(ooff, otid, osc, R, T)
print("candidates fresh vs oracle equal:", np.array_equal(c[0], ooff) and np.array_equal(c[1], otid) and np.array_equal(c[2], osc))
for tier in (1, 2):
    g = sqb.Engine(ks, T, sketch_fraction=SKETCH)
    g.load_index(0, *post[31])
    g.set_option("vote_tier", tier)
    x = run(g, d["reads"])
    g.set_option("exact_classes", 1)
    pix, nrx, _, _ = g.finish(0, 20, 0.01)
    g.close()
    print("tier", tier, "cand equal oracle:", np.array_equal(x[0], ooff) and np.array_equal(x[1], otid) and np.array_equal(x[2], osc),
          "pi max rel (table path):", float(np.max(np.abs(x[3] - opi) / opi)), " (sort path):", float(np.max(np.abs(pix - opi) / opi)))
print("fresh vs oracle: pi max rel", float(np.max(np.abs(c[3] - opi) / opi)), " after-reset vs oracle:", float(np.max(np.abs(b[3] - opi) / opi)))
