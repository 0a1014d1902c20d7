"""Regenerates tests/golden/ by running the UNMODIFIED reference (oracle/_ref/kma, built from
/root/reference by oracle/Makefile.ref) on a small seeded synthetic data set.

    python tests/golden/make_golden.py

Outputs (gz where large): db.fsa, db.{comp.b,length.b,seq.b,name} (kma index), reads.fq,
s1.bin (kma -s1 stage-1 stream), s2.bin (kma -1t1 -s2 stage-2 stream).
"""
import gzip
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from kma_b200 import synth  # noqa: E402

KMA = os.path.join(ROOT, "oracle", "_ref", "kma")


def main():
    d = tempfile.mkdtemp()
    names, seqs = synth.gene_db(1234, n_families=40, n_variants=5, len_lo=300, len_hi=900)
    synth.write_fasta(f"{d}/db.fsa", names, seqs)
    reads = synth.short_reads(99, seqs, 3000, L=150, sub=0.01, n_rate=0.002, junk_frac=0.05)
    synth.write_fastq(f"{d}/reads.fq", reads)
    subprocess.run([KMA, "index", "-i", "db.fsa", "-o", "db"], cwd=d, check=True, stderr=subprocess.DEVNULL)
    with open(f"{d}/s1.bin", "wb") as fo:
        subprocess.run([KMA, "-i", "reads.fq", "-o", "o", "-t_db", "db", "-1t1", "-s1"], cwd=d, check=True, stdout=fo,
                       stderr=subprocess.DEVNULL)
    with open(f"{d}/s2.bin", "wb") as fo:
        subprocess.run([KMA, "-i", "reads.fq", "-o", "o", "-t_db", "db", "-1t1", "-s2"], cwd=d, check=True, stdout=fo,
                       stderr=subprocess.DEVNULL)
    for f in ["db.fsa", "db.comp.b", "db.seq.b", "reads.fq", "s1.bin", "s2.bin"]:
        with open(f"{d}/{f}", "rb") as fi, gzip.GzipFile(os.path.join(HERE, f + ".gz"), "wb", mtime=0) as fo:
            shutil.copyfileobj(fi, fo)
    for f in ["db.length.b", "db.name"]:
        shutil.copy(f"{d}/{f}", os.path.join(HERE, f))
    shutil.rmtree(d)


if __name__ == "__main__":
    main()
