"""Multi-GPU k-mer counting: one process per GPU, torch.distributed (NCCL over NVLink) for the single exchange step.

The path shards naturally: a canonical k-mer's device bin is a pure function of the k-mer, so bins are
independent units.  Every rank partitions ITS reads into the same nb1 coarse bins (k1), the coarse-bin regions are moved
to their owners with one all-to-all (coarse bin b belongs to rank b // bins_per_rank; regions are contiguous in the
partition buffer), and each owner fine-splits, counts and sorts the bins it gathered (k2a, k2b, k3).  The reference has
no collective at all: its only exchange medium are the SuperKmerBinFiles temp files
(/root/reference/gatb-core/src/gatb/tools/storage/impl/Storage.cpp:310-347); this all-to-all replaces them.

Everything that touches torch.distributed here works on CPU tensors with the gloo backend too (tests/test_multigpu_host.py).
"""
import json
import os
import time

import numpy as np
import torch
import torch.distributed as dist


def owner_of_bin(b, bins_per_rank):
    return b // bins_per_rank


def exchange_bins(bins, cursors, world):
    """All-to-all of the partition buffers.

    bins        uint8  [nb1 * cap * record_bytes]   (nb1 = world * bins_per_rank; bin regions contiguous per owner)
    cursors     int32  [nb1]
    Returns (recv_bins [world, bpr*cap*rb], recv_cursors [world, bpr]):
    row s = what source rank s produced for the bins THIS rank owns.  (The fine split counts its own bins from the
    records, so nothing but the records and their cursors has to travel.)
    """
    recv_bins = torch.empty_like(bins)
    recv_cur = torch.empty_like(cursors)
    if world == 1:
        recv_bins.copy_(bins)
        recv_cur.copy_(cursors)
    else:
        dist.all_to_all_single(recv_bins, bins)
        dist.all_to_all_single(recv_cur, cursors)
    return recv_bins.view(world, -1), recv_cur.view(world, -1)


def merge_sorted_runs(runs):
    """Host-side merge of per-rank results for ONE partition key: runs = [(lo, hi, counts), ...], each ascending by
    (hi, lo) and holding DISJOINT k-mers (a k-mer lives in exactly one device bin, hence on one rank).
    Returns the ascending concatenation -- what ICountProcessor::process must see for that partition."""
    lo = np.concatenate([r[0] for r in runs])
    hi = np.concatenate([r[1] for r in runs])
    cn = np.concatenate([r[2] for r in runs])
    order = np.lexsort((lo, hi))
    return lo[order], hi[order], cn[order]


def count_distributed(gpu, params, d_reads, n_reads_local, n_reads_global, total_kmers_global, rank, world, repart=None,
                      d_offsets=None, timers=None):
    """One distributed counting pass.  Returns (device Result of this rank's bins, stats dict with GLOBAL sums)."""
    dev = torch.device("cuda", gpu.device)
    geom = gpu.plan(params, total_kmers_global, n_reads_global, world)
    t = {}
    while True:
        nb1, cap, rb, fb = geom.nb1, geom.cap, geom.record_bytes, geom.fine_bits
        bins = torch.empty(nb1 * cap * rb, dtype=torch.uint8, device=dev)
        cursors = torch.zeros(nb1, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        t0 = time.time()
        st = gpu.partition_into(params, geom, d_reads, d_offsets, n_reads_local, bins.data_ptr(), cursors.data_ptr())
        t["partition"] = time.time() - t0
        flag = torch.tensor([st[3], int(cursors.max().item())], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if int(flag[0].item()) == 0:
            break
        geom.cap = (int(flag[1].item()) + 15) & ~15         # a bin overflowed somewhere: every rank re-runs with the global demand
        del bins, cursors
    t0 = time.time()
    recv_bins, recv_cur = exchange_bins(bins, cursors, world)
    del bins, cursors
    torch.cuda.synchronize()
    t["exchange"] = time.time() - t0
    bpr = geom.bins_per_rank
    src_bins = [recv_bins[s].data_ptr() for s in range(world)]
    src_cur = [recv_cur[s].data_ptr() for s in range(world)]
    gathered = int(recv_cur.clamp(max=geom.cap).sum().item())
    tot = torch.tensor([st[0], st[2]], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
    avg_len = float(tot[0].item()) / max(float(tot[1].item()), 1.0)       # k-mers per record, whole job
    # k-mers in the bins this rank owns: estimate with 25 % head-room, never above the hard bound
    kmers_bound = int(min(gathered * geom.maxlen, gathered * avg_len * 1.25 + 65536))
    t0 = time.time()
    res = gpu.count_bins(params, geom, src_bins, src_cur, bpr, kmers_bound, repart=repart)
    t["count"] = time.time() - t0
    t["count_kernels"] = [float(x) for x in res.kernel_seconds][:5]
    t["overflow_kmers"] = int(res.stats[11])
    t["count_stages"] = [float(x) for x in res.seconds][:7]
    t["overflow_bins"] = int(res.stats[8])
    t["bins"] = int(res.stats[7])
    sums = torch.tensor([st[0], st[1], int(res.stats[2]), int(res.stats[3]), st[2], int(res.n_items)], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(sums)
    stats = dict(zip(["kmers_nb_valid", "kmers_nb_invalid", "kmers_nb_distinct", "kmers_nb_solid", "records", "items"],
                     [int(x) for x in sums.tolist()]))
    stats["exchanged_bytes_per_rank"] = int(geom.nb1 * geom.cap * geom.record_bytes * (world - 1) // world)
    if timers is not None:
        timers.update(t)
    return res, stats


_PINNED = {}


def result_to_pinned(gpu, res, params):
    """Device Result -> pinned host tensors (cached, grow-only) with one copy per array; returns numpy views."""
    W = 1 if params.kmer_size < 32 else 2
    n, nk = int(res.n_items), int(res.n_keys)

    def buf(name, count, dtype):
        t = _PINNED.get(name)
        if t is None or t.numel() < max(count, 1):
            t = torch.empty(max(count, 1) + max(count, 1) // 8, dtype=dtype, pin_memory=True)
            _PINNED[name] = t
        return t[:count].numpy()
    out = {"part_offsets": buf("offs", nk + 1, torch.int64), "histogram": buf("hist", params.histo_max + 1, torch.int64),
           "kmers_lo": buf("lo", n, torch.int64), "counts": buf("cnt", n, torch.int32)}
    gpu.d2h(out["part_offsets"], res.part_offsets)
    gpu.d2h(out["histogram"], res.histogram)
    if n:
        gpu.d2h(out["kmers_lo"], res.kmers_lo)
        gpu.d2h(out["counts"], res.counts)
        if W == 2:
            out["kmers_hi"] = buf("hi", n, torch.int64)
            gpu.d2h(out["kmers_hi"], res.kmers_hi)
    out["n_items"] = n
    return out


def bench(args, rank, world, local):
    """bench.py --gpus N (N>1): weak scaling, args.reads reads per GPU out of one genome sized for all of them."""
    import gatb_core_b200
    from bench import K, M, L, ABUNDANCE_MIN, COVERAGE, SEED, METRIC, UNIT, ClockSampler
    gpu = gatb_core_b200.GatbGpu(local)
    dev = torch.device("cuda", local)
    n = args.reads
    n_global = n * world
    genome = n_global * L // COVERAGE
    nbytes = (n * L + 3) // 4
    reads = torch.zeros(nbytes + 64, dtype=torch.uint8, device=dev)
    gpu.synth_reads_dev(SEED, genome, rank * n, n, L, reads.data_ptr())
    gpu.synchronize()
    params = gpu.make_params(K, M, abundance_min=ABUNDANCE_MIN, read_len=L)
    total_kmers = n_global * (L - K + 1)

    def step(timers=None):
        res, stats = count_distributed(gpu, params, reads.data_ptr(), n, n_global, total_kmers, rank, world, timers=timers)
        gpu.result_free(res)
        return stats

    for _ in range(args.warmup):
        stats = step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dist.barrier()
    torch.cuda.synchronize()
    launches0 = gpu.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    timers = {}
    for _ in range(args.steps):
        stats = step(timers)
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    el = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
    dist.all_reduce(el, op=dist.ReduceOp.MAX)
    per_step = float(el.item()) / args.steps
    launches = gpu.kernel_launches - launches0

    # ---- end to end: pinned host reads -> device, results -> pinned host, every step ----
    h_reads = torch.empty(nbytes + 64, dtype=torch.uint8, pin_memory=True)
    h_reads.copy_(reads)
    torch.cuda.synchronize()
    e2e = []
    d2h_bytes = 0
    for i in range(1 + args.steps):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.time()
        reads.copy_(h_reads, non_blocking=True)
        torch.cuda.synchronize()
        t1 = time.time()
        res, stats = count_distributed(gpu, params, reads.data_ptr(), n, n_global, total_kmers, rank, world)
        t2 = time.time()
        host = result_to_pinned(gpu, res, params)
        gpu.result_free(res)
        if rank == 0 and os.environ.get("GATB_BENCH_DEBUG"):
            print("e2e step %d: h2d %.1f ms, count %.1f ms, d2h %.1f ms" % (i, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (time.time() - t2) * 1e3), file=__import__("sys").stderr)
        d2h_bytes = int(host["n_items"]) * 12 + (10001 + 2) * 8
        el2 = torch.tensor([time.time() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(el2, op=dist.ReduceOp.MAX)
        if i >= 1:
            e2e.append(float(el2.item()))
    e2e_step = sum(e2e) / len(e2e)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        line = {"metric": METRIC, "value": stats["kmers_nb_distinct"] / per_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u64", "data": "synthetic",
                "config": {"workload": "k=31, %d synthetic 150bp reads (%d per GPU), %dxB200, minimizer buckets sharded via NCCL all-to-all, m=10, abundance-min=2" % (n_global, n, world),
                           "reads": n_global, "genome_nt": genome, "coverage": COVERAGE, "error_rate": 0.01,
                           "l2": "per-GPU inputs (%.1f GB packed reads) far exceed the 126 MB L2" % (nbytes / 1e9)},
                "input_bases_per_s": n_global * L / per_step, "kmer_occurrences_per_s": stats["kmers_nb_valid"] / per_step,
                "distinct": stats["kmers_nb_distinct"], "solid": stats["kmers_nb_solid"], "records": stats["records"],
                "stage_seconds_rank0_last_step": timers, "exchanged_bytes_per_rank": stats["exchanged_bytes_per_rank"],
                "e2e": {"value": stats["kmers_nb_distinct"] / e2e_step, "unit": UNIT, "h2d_bytes_per_step": nbytes * world,
                        "d2h_bytes_per_step": d2h_bytes * world, "ms_per_step": e2e_step * 1e3},
                "gpu_launches": int(launches) * world, "clocks": clocks, "roofline": None, "cpu_baseline": None}
        print(json.dumps(line))
    gpu.close()
    dist.destroy_process_group()
