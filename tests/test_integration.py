"""GATB-core's own SortingCountAlgorithm<span> running on the GPU path (integration/): tools written against the reference's
public API -- the reference's examples/kmer/kmer12.cpp UNCHANGED, and integration/dsk_tool.cpp -- are linked once against the
reference's instantiation (CPU) and once against integration/SortingCountAlgorithmGPU.cpp + libgatb_b200.so (GPU), and must
produce byte-identical solid partitions (default processor chain -> Partition<Count>), histogram, statistics and custom
ICountProcessor call sequences.  The binaries are built in this container (make -C integration, needs /root/reference) and
travel to the GPU box with the snapshot."""
import filecmp
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "integration", "_build")
REF = "/root/reference/gatb-core/src"


def build():
    if os.path.isdir(REF):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "-j8", "all"], check=True)
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "integration"), "-j8", "all"], check=True)
    missing = [t for t in ("dsk_tool_cpu", "dsk_tool_gpu", "kmer12_gpu") if not os.path.exists(os.path.join(BUILD, t))]
    if missing:
        pytest.skip("integration/_build lacks %s (built where /root/reference exists)" % missing)


def write_fasta(oracle, path, n, L, seed, with_n=False, fastq=False):
    codes = oracle.synth_reads(seed, n * L // 30, 0, n, L).reshape(n, L)
    rng = np.random.default_rng(seed)
    with open(path, "wb") as f:
        for i, r in enumerate(codes):
            s = bytearray(oracle.codes_to_ascii(r))
            if with_n and i % 7 == 0:
                s[int(rng.integers(0, L))] = ord("N")
            if fastq:
                f.write(b"@r%d\n" % i + bytes(s) + b"\n+\n" + bytes(rng.integers(33, 74, len(s), dtype=np.uint8)) + b"\n")
            elif i % 11 == 0:                                 # multi-line records, like real FASTA files
                f.write(b">r%d some comment\n" % i + bytes(s[:60]) + b"\n" + bytes(s[60:]) + b"\n")
            else:
                f.write(b">r%d\n" % i + bytes(s) + b"\n")


def run_tool(tool, fasta, out, k, extra, abundance_min=2, env=None, storage="hdf5"):
    cmd = [os.path.join(BUILD, tool), "-in", fasta, "-kmer-size", str(k), "-abundance-min", str(abundance_min), "-out", out,
           "-out-dir", os.path.dirname(out), "-out-tmp", os.path.dirname(out), "-storage-type", storage, "-verbose", "0"] + extra
    return subprocess.run(cmd, capture_output=True, text=True, cwd=os.path.dirname(out), env=dict(os.environ, **(env or {})))


def test_integration_builds_and_has_no_cpu_fallback(oracle, tmp_path):
    import torch
    build()
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu tests")
    fa = str(tmp_path / "r.fa")
    write_fasta(oracle, fa, 2000, 100, 3)
    r = run_tool("dsk_tool_gpu", fa, str(tmp_path / "g"), 21, [])
    assert r.returncode != 0 and "no CPU fallback" in (r.stdout + r.stderr)
    # the reference build of the same tool agrees with the oracle port on the number of distinct / solid k-mers
    r = run_tool("dsk_tool_cpu", fa, str(tmp_path / "c"), 21, ["-nb-cores", "2"])
    assert os.path.getsize(str(tmp_path / "c.h5dump.txt")) > 0
    assert r.returncode == 0, r.stdout + r.stderr
    info = dict(l.split() for l in open(str(tmp_path / "c.info.txt")))
    seqs = [l.strip() for l in open(fa, "rb") if not l.startswith(b">")]
    seqs = []
    cur = b""
    for l in open(fa, "rb"):
        if l.startswith(b">"):
            if cur:
                seqs.append(cur)
            cur = b""
        else:
            cur += l.strip()
    seqs.append(cur)
    want = oracle.dsk(seqs, 21, 8, np.zeros(4 ** 8, np.uint16), 1, abundance_min=2)
    assert int(info["kmers_nb_distinct"]) == int(want["stats"][2]) and int(info["kmers_nb_solid"]) == int(want["stats"][3])


# the GPU build reads a plain FASTA / FASTQ file itself and parses it on the device; "_bank_iterator" cases force the other
# route (the reference's own bank iterator feeding ASCII batches), which is what gzip files, albums and in-memory banks take
CASES = [("k21_cores4", 21, 100, ["-nb-cores", "4"], False),
         ("k31_fastq", 31, 150, ["-nb-cores", "4"], True),
         ("k31_bank_iterator", 31, 150, ["-nb-cores", "2"], True),
         ("k31_default", 31, 150, [], False),
         ("k31_three_passes_with_N", 31, 150, ["-max-disk", "1", "-nb-cores", "2"], True),
         ("k31_m8_small_memory", 31, 150, ["-minimizer-size", "8", "-max-memory", "100", "-nb-cores", "3"], True),
         ("k63", 63, 250, ["-nb-cores", "4"], True),
         ("k47_solidity_range", 47, 150, ["-abundance-max", "20", "-nb-cores", "2"], False)]


@pytest.mark.gpu
@pytest.mark.parametrize("name,k,L,extra,with_n", CASES, ids=[c[0] for c in CASES])
def test_gatb_tool_on_gpu_equals_reference_build(oracle, tmp_path, name, k, L, extra, with_n):
    build()
    fastq = name.endswith("fastq")
    fa = str(tmp_path / ("reads.fq" if fastq else "reads.fa"))
    write_fasta(oracle, fa, 20000, L, 100 + k, with_n, fastq)
    outs = {}
    for tool in ("dsk_tool_cpu", "dsk_tool_gpu"):
        d = tmp_path / tool
        d.mkdir()
        storage = "file" if name == "k21_cores4" else "hdf5"            # .h5 is the reference's default output; one case keeps the file storage
        r = run_tool(tool, fa, str(d / "x"), k, extra, env={"GATB_GPU_NO_TEXT_PARSER": "1"} if name.endswith("bank_iterator") else None, storage=storage)
        assert r.returncode == 0, tool + ": " + r.stdout + r.stderr
        outs[tool] = str(d / "x")
        if storage == "hdf5":
            assert open(str(d / "x.h5"), "rb").read(8) == b"\x89HDF\r\n\x1a\n"            # a real HDF5 file
    # .h5dump.txt: the .h5 re-opened from disk with the reference's Storage classes (what Graph::load consumes): every dsk/solid/<p>
    # dataset, the dsk attributes, the histogram and the Repartitor table -- identical dataset by dataset
    for suffix in (".solid.txt", ".histo.txt", ".info.txt", ".all.txt") + ((".h5dump.txt",) if storage == "hdf5" else ()):
        a, b = outs["dsk_tool_cpu"] + suffix, outs["dsk_tool_gpu"] + suffix
        assert os.path.getsize(a) > 0
        assert filecmp.cmp(a, b, shallow=False), "%s differs between the reference build and the GPU build (%s)" % (suffix, name)


@pytest.mark.gpu
def test_reference_example_kmer12_unchanged_on_gpu(oracle, tmp_path):
    # examples/kmer/kmer12.cpp compiled from the reference tree as is: a custom CountProcessorAbstract<span> added with
    # addProcessor(); it prints the number of k-mers whose count is in [abundance-min, 2^30)
    build()
    fa = str(tmp_path / "reads.fa")
    write_fasta(oracle, fa, 20000, 150, 5)
    r = subprocess.run([os.path.join(BUILD, "kmer12_gpu"), "-in", fa, "-kmer-size", "31", "-abundance-min", "3", "-out-dir", str(tmp_path),
                        "-out-tmp", str(tmp_path)], capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout + r.stderr
    c = run_tool("dsk_tool_cpu", fa, str(tmp_path / "c"), 31, [], abundance_min=3)
    assert c.returncode == 0, c.stdout + c.stderr
    info = dict(l.split() for l in open(str(tmp_path / "c.info.txt")))
    assert ("%s" % info["kmers_nb_solid"]) in r.stdout, r.stdout
