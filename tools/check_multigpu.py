#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/check_multigpu.py : the N-GPU sharded count, merged on rank 0, must equal the
single-GPU count of the same reads (bit-exact), for k=31 (Kmer<32>) and k=63 (Kmer<64>)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gatb_core_b200  # noqa: E402
from gatb_core_b200 import multigpu  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gpu = gatb_core_b200.GatbGpu(local)
    for k, L, n in ((31, 150, 400000), (63, 250, 100000)):
        m, nparts = 10, 5
        n_global = n * world
        genome = n_global * L // 30
        repart = (np.arange(4 ** m) * 2654435761 % nparts).astype(np.uint16)
        params = gpu.make_params(k, m, nb_partitions=nparts, abundance_min=2, read_len=L)
        nbytes = (n * L + 3) // 4
        reads = torch.zeros(nbytes + 64, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()      # torch's zero fill vs the library's own stream
        gpu.synth_reads_dev(7, genome, rank * n, n, L, reads.data_ptr())
        gpu.synchronize()
        res, stats = multigpu.count_distributed(gpu, params, reads.data_ptr(), n, n_global, n_global * (L - k + 1), rank, world, repart=repart, route=True)
        mine = gpu.result_to_host(res, params)
        bloom, bloom_bits = multigpu.bloom_distributed(gpu, "neighbor", k, res, stats["kmers_nb_solid"], world)
        bloom = bloom.cpu().numpy()
        gpu.result_free(res)
        if k == 31:     # the same without the second exchange: per-rank ascending runs of disjoint k-mers, merged on the host
            res_b, _ = multigpu.count_distributed(gpu, params, reads.data_ptr(), n, n_global, n_global * (L - k + 1), rank, world, repart=repart, route=False)
            runs = gpu.result_to_host(res_b, params)
            gpu.result_free(res_b)
            box = [None] * world
            dist.all_gather_object(box, runs["parts"])
            if rank == 0:
                unrouted = [multigpu.merge_sorted_runs([b[key] for b in box]) for key in range(nparts)]
        gathered = [None] * world
        dist.all_gather_object(gathered, {"parts": mine["parts"], "hist": mine["histogram"]})
        if rank == 0:
            allr = torch.zeros((n_global * L + 3) // 4 + 64, dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            assert (n * L) % 4 == 0
            gpu.synth_reads_dev(7, genome, 0, n_global, L, allr.data_ptr())
            gpu.synchronize()
            single = gpu.count_dev(allr.data_ptr(), None, n_global, params, repart=repart)
            want = gpu.result_to_host(single, params)
            gpu.result_free(single)
            for key in range(nparts):
                # second exchange: the partition is whole and ascending on its owner rank, every other rank holds nothing of it
                for r in range(world):
                    if r != key % world:
                        assert len(gathered[r]["parts"][key][0]) == 0, (k, key, r)
                lo, hi, cn = gathered[key % world]["parts"][key]
                wlo, whi, wcn = want["parts"][key]
                assert len(lo) == len(wlo) and (lo == wlo).all() and (hi == whi).all() and (cn == wcn).all(), (k, key)
                if k == 31:
                    ulo, uhi, ucn = unrouted[key]
                    assert len(ulo) == len(wlo) and (ulo == wlo).all() and (ucn == wcn).all(), ("unrouted", key)
            hist = sum(g["hist"].astype(np.int64) for g in gathered)
            assert (hist == want["histogram"].astype(np.int64)).all()
            assert stats["kmers_nb_distinct"] == want["stats"]["kmers_nb_distinct"]
            # Bloom filter of ALL solid k-mers on one GPU == per-rank filters gathered and OR-ed
            lo1 = np.concatenate([want["parts"][key][0] for key in range(nparts)])
            hi1 = np.concatenate([want["parts"][key][1] for key in range(nparts)]) if k > 31 else None
            size, nh = gpu.bloom_params(k, len(lo1))
            b1, bits1 = gpu.bloom("neighbor", size, nh, k, lo1, hi1)
            assert bits1 == bloom_bits and len(b1) == len(bloom) and (b1 == bloom).all(), "distributed Bloom filter differs"
            print("k=%d: %d GPUs == 1 GPU: %d distinct, %d solid k-mers, histogram identical, Bloom filter (%d bytes) identical" % (k, world, stats["kmers_nb_distinct"], stats["kmers_nb_solid"], len(bloom)))
        dist.barrier()
    gpu.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
