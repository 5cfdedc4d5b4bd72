"""Row f3: the Repartitor table.  The sampling pass as the device runs it (gatb_core_b200/csrc/k_repart_core.cuh, compiled for the host
by tests/cpp/test_repart_core.cpp) + the host distribution (repart_host.h) against the UNMODIFIED reference's RepartitorAlgorithm
(oracle/_ref), and -- on a GPU -- gatb_gpu_repartition against the same."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def synth_seqs(oracle, seed, n, L, p_n=0.0):
    codes = oracle.synth_reads(seed, n * L // 30, 0, n, L)
    seqs = [oracle.codes_to_ascii(r) for r in codes.reshape(n, L)]
    if p_n:
        rng = np.random.default_rng(seed)
        out = []
        for s in seqs:
            a = np.frombuffer(s, np.uint8).copy()
            a[rng.random(L) < p_n] = ord("N")
            out.append(a.tobytes())
        seqs = out
    return seqs


def write_fasta(path, seqs):
    with open(path, "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b">r%d\n" % i + s + b"\n")


# the reference samples max(5 % of the estimated reads, 10^6) super-k-mers (RepartitionAlgorithm.cpp:451): with 100000 reads of 150 nt
# (~14 super-k-mers each) the iteration is cancelled inside the bank, with the smaller inputs it runs to the end
CASES = [(31, 10, 13, 100000, 150, 0.0), (21, 8, 4, 3000, 100, 0.01), (31, 10, 64, 5000, 150, 0.0), (25, 6, 7, 2000, 120, 0.02),
         (63, 10, 9, 3000, 250, 0.01), (32, 8, 5, 2000, 150, 0.0), (47, 9, 16, 2000, 200, 0.0)]


def host_table(tmp_path, seqs, k, m, nparts, to_see):
    src = os.path.join(ROOT, "tests", "cpp", "test_repart_core.cpp")
    exe = os.path.join(ROOT, "tests", "cpp", "test_repart_core")
    inc = os.path.join(ROOT, "gatb_core_b200", "csrc")
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wno-unknown-pragmas", "-I" + inc, src, "-o", exe], check=True)
    txt, out = os.path.join(tmp_path, "seqs.txt"), os.path.join(tmp_path, "table.bin")
    with open(txt, "wb") as f:
        f.write(b"\n".join(seqs) + b"\n")
    r = subprocess.run([exe, txt, str(k), str(m), str(nparts), str(to_see), out], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return np.fromfile(out, np.uint16), r.stdout


@pytest.mark.parametrize("k,m,nparts,n,L,p_n", CASES)
def test_sampling_and_distribution_on_the_host_equal_the_reference(oracle, reference, tmp_path, k, m, nparts, n, L, p_n):
    seqs = synth_seqs(oracle, 5 + k, n, L, p_n)
    fa = os.path.join(tmp_path, "in.fa")
    write_fasta(fa, seqs)
    want = reference.repartition(fa, k, m, nparts)
    got, log = host_table(str(tmp_path), seqs, k, m, nparts, 1000000)
    assert len(got) == len(want) == 4 ** m
    assert (got == want).all(), (log, int((got != want).sum()))
    if n >= 100000:
        assert int(log.split()[0]) < n, "the sampling cutoff was expected inside the bank: " + log


@pytest.mark.gpu
@pytest.mark.parametrize("k,m,nparts,n,L,p_n", CASES)
def test_gpu_repartition_equals_the_reference(oracle, reference, tmp_path, k, m, nparts, n, L, p_n):
    import gatb_core_b200
    from test_gpu_parity import pack_seqs
    gpu = gatb_core_b200.GatbGpu(0)
    seqs = synth_seqs(oracle, 5 + k, n, L, p_n)
    fa = os.path.join(tmp_path, "in.fa")
    write_fasta(fa, seqs)
    want = reference.repartition(fa, k, m, nparts)
    packed, offs, mask = pack_seqs(oracle, seqs)
    params = gpu.make_params(k, m, nb_partitions=nparts)
    table, info = gpu.repartition(packed, offs, len(seqs), params, 1000000, n_mask=mask)
    assert (table == want).all(), (info, int((table != want).sum()))
    if n >= 100000:
        assert info[0] < n
    # and the table drives a count that equals the reference's own run with its own table
    gpu.close()
