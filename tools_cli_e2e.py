"""File-to-CSV timing of the drop-in CLI (and optionally the reference program) on a synthetic data set.
usage: python tools_cli_e2e.py <n_genes> <n_reads> [--ref N]   (writes under /tmp/sq_e2e)"""
import json, os, subprocess, sys, time
sys.path[:0] = [os.path.dirname(os.path.abspath(__file__))]
from _sqpkg import sqb
import torch

genes, n_reads = int(sys.argv[1]), int(sys.argv[2])
ref_n = int(sys.argv[sys.argv.index("--ref") + 1]) if "--ref" in sys.argv else 0
d = "/tmp/sq_e2e"
os.makedirs(d, exist_ok=True)
dev = "cuda:0" if torch.cuda.is_available() else "cpu"
t0 = time.time()
tx = sqb.synth.make_transcriptome(genes, seed=7, device=dev)
T = tx["t_off"].numel() - 1
names = sqb.synth.transcript_names(T)
txc = {k: v.cpu() for k, v in tx.items()}
sqb.synth.write_fasta(d + "/tx.fa", names, txc, width=70)
done = 0
for ch in sqb.synth.simulate_reads(tx, n_reads, 150, seed=5, chunk=1 << 20):
    chc = {k: v.cpu() for k, v in ch.items()}
    done += sqb.synth.write_fastq(d + "/reads.fq", chc, first_id=done, prefix="read", mode="ab" if done else "wb")
print("data: T=%d reads=%d fasta=%.0f MB fastq=%.0f MB (%.1fs)" % (T, done, os.path.getsize(d + "/tx.fa") / 1e6,
                                                                   os.path.getsize(d + "/reads.fq") / 1e6, time.time() - t0))
exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "build", "test")
t0 = time.time(); subprocess.run([exe, "-k", "31", "-o", "index", d + "/tx.fa", d + "/tx.idx"], check=True); t_index = time.time() - t0
for rep in range(2):
    t0 = time.time()
    subprocess.run([exe, "--index-cache", "--report", d + "/rep.json", "-o", "quant", d + "/tx.idx", d + "/reads.fq", d + "/out.csv"], check=True, stdout=subprocess.DEVNULL, env=dict(os.environ, SQ_TRACE="1"))
    t_q = time.time() - t0
    print("ours: index %.2fs quant wall %.2fs -> %.0f reads/s" % (t_index, t_q, done / t_q), json.load(open(d + "/rep.json")))
if ref_n:
    ref = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle", "_ref", "ref_test")
    subprocess.run("head -n %d %s/reads.fq > %s/reads_ref.fq" % (4 * ref_n, d, d), shell=True, check=True)
    t0 = time.time(); out = subprocess.run([ref, "-o", "quant", d + "/tx.idx", d + "/reads_ref.fq", d + "/ref.csv"], check=True, capture_output=True, text=True); t_r = time.time() - t0
    print("reference program on %d reads: wall %.2fs (incl. its index load)" % (ref_n, t_r))
