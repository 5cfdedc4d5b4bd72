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


def exchange_bins(bins, cursors, world, rank=0, used_bytes=None, wait=True):
    """All-to-all of the partition buffers.

    bins        uint8  [nb1 * cap * record_bytes]   (nb1 = world * bins_per_rank; one contiguous region per owner rank)
    cursors     int32  [nb1]
    used_bytes  optional list [world]: how much of region r is in use (a region is filled round by round from its
                start, kernels.h coarse_index, so everything past the last used round is slack that need not travel)
    Returns (pieces, recv_cursors [world, bpr]): pieces[s] is a uint8 tensor holding what source rank s produced for the
    bins THIS rank owns (its used prefix at least); pieces[rank] is a view of this rank's own region -- no copy.
    With wait=False the transfers are only started: a third value, the list of requests to wait for, is returned.
    (The fine split counts its own bins from the records, so nothing but the records and their cursors has to travel.)
    """
    region = bins.numel() // world
    recv_cur = torch.empty_like(cursors)
    if world == 1:
        recv_cur.copy_(cursors)
        return ([bins.view(world, -1)[0]], recv_cur.view(world, -1)) if wait else ([bins.view(world, -1)[0]], recv_cur.view(world, -1), [])
    dist.all_to_all_single(recv_cur, cursors)
    send = torch.tensor([region] * world if used_bytes is None else [int(x) for x in used_bytes], dtype=torch.int64, device=bins.device)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send)                       # how much every source is going to send me
    recv = [int(x) for x in recv.tolist()]
    send = [int(x) for x in send.tolist()]
    views = bins.view(world, -1)
    pieces, ops = [None] * world, []
    for s in range(world):
        if s == rank:
            pieces[s] = views[rank]
            continue
        pieces[s] = torch.empty(max(recv[s], 16), dtype=torch.uint8, device=bins.device)
        if recv[s]:
            ops.append(dist.P2POp(dist.irecv, pieces[s][:recv[s]], s))
        if send[s]:
            ops.append(dist.P2POp(dist.isend, views[s][:send[s]], s))
    reqs = dist.batch_isend_irecv(ops) if ops else []
    if wait:
        for req in reqs:
            req.wait()
        return pieces, recv_cur.view(world, -1)
    return pieces, recv_cur.view(world, -1), reqs


def merge_sorted_runs(runs):
    """Host-side merge of per-rank results for ONE partition key: runs = [(lo, hi, counts), ...], each ascending by
    (hi, lo) and holding DISJOINT k-mers (a k-mer lives in exactly one device bin, hence on one rank).
    Returns the ascending concatenation -- what ICountProcessor::process must see for that partition."""
    lo = np.concatenate([r[0] for r in runs])
    hi = np.concatenate([r[1] for r in runs])
    cn = np.concatenate([r[2] for r in runs])
    order = np.lexsort((lo, hi))
    return lo[order], hi[order], cn[order]


def bloom_distributed(gpu, kind, k, res, n_solid_global, world):
    """Bloom filter of the solid k-mers of a distributed count (BloomAlgorithm::execute, kmer/impl/BloomAlgorithm.cpp:155-203):
    sized from the GLOBAL number of solid k-mers exactly like the reference (float32 product), every rank inserts the solid
    k-mers it owns into a private full-size bit array on its GPU (gatb_gpu_bloom_dev, byte-exact hash functions), the arrays are
    all-gathered over NCCL and OR-ed.  Returns (uint8 device tensor with the reference's byte layout, bit size).
    `res` is the device Result of count_distributed (emit range = solid range)."""
    import gatb_core_b200
    dev = torch.device("cuda", gpu.device)
    size, nh = gpu.bloom_params(k, n_solid_global)
    nbytes, bits = gpu.bloom_layout(kind, size)
    padded = (nbytes + 3) // 4 * 4
    mine = torch.zeros(padded, dtype=torch.uint8, device=dev)
    torch.cuda.current_stream().synchronize()
    n = int(res.n_items)
    gpu._check(gpu.L.gatb_gpu_bloom_dev(gpu.ctx, gatb_core_b200.BLOOM_KINDS[kind], size, nh, k, res.kmers_lo, res.kmers_hi, n, mine.data_ptr()))
    gpu.synchronize()
    if world > 1:
        allb = torch.empty(world * padded, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allb, mine)
        v = allb.view(world, padded).view(torch.int32)
        out = v[0].clone()
        for r in range(1, world):
            out |= v[r]
        mine = out.view(torch.uint8)
    return mine[:nbytes], bits


class _DevArray:
    """A device array owned by the library, as torch sees it (zero copy, __cuda_array_interface__)"""
    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr if count else 0, False), "version": 2}


def _as_tensor(ptr, count, dtype, dev):
    if count == 0:
        return torch.empty(0, dtype=dtype, device=dev)
    typestr = {torch.int64: "<i8", torch.int32: "<i4", torch.uint8: "|u1"}[dtype]
    return torch.as_tensor(_DevArray(ptr, count, typestr), device=dev)


def _count_and_route(gpu, params, geom, src_bins, src_cur, bpr, kmers_bound, repart, rank, world, dev, t):
    """count -> route by partition owner -> all-to-all of the emitted k-mers (values, counts, keys) -> sort of the owned partitions.
    The result of partition key k is ONE ascending sequence on rank k % world, like the reference's (every other rank holds nothing
    for k); histogram and statistics stay per rank (they combine by sum)."""
    W = 1 if params.kmer_size < 32 else 2
    res1, d_keys, send, cap = gpu.count_bins_routed(params, geom, src_bins, src_cur, bpr, kmers_bound, world, repart=repart)
    t0 = time.time()
    send_t = torch.tensor(send, dtype=torch.int64, device=dev)
    recv_t = torch.empty_like(send_t)
    dist.all_to_all_single(recv_t, send_t)
    recv = [int(x) for x in recv_t.tolist()]
    n_recv = sum(recv)
    starts = [sum(recv[:r]) for r in range(world)]
    # (element size in bytes: the 16-bit keys travel as bytes, NCCL has no 16-bit integer type)
    arrays = [(res1.kmers_lo, torch.int64, 1)] + ([(res1.kmers_hi, torch.int64, 1)] if W == 2 else []) + [(res1.counts, torch.int32, 1), (d_keys, torch.uint8, 2)]
    got, works = [], []
    for ptr, dt, sc in arrays:
        src = _as_tensor(ptr, cap * world * sc, dt, dev)      # one region of 'cap' items per destination rank
        dst = torch.empty(max(n_recv, 1) * sc, dtype=dt, device=dev)
        works.append(dist.all_to_all([dst[starts[r] * sc:(starts[r] + recv[r]) * sc] for r in range(world)],
                                     [src[r * cap * sc:(r * cap + send[r]) * sc] for r in range(world)], async_op=True))
        got.append(dst)
    for w in works:
        w.wait()
    torch.cuda.current_stream().synchronize()
    t["route_exchange"] = time.time() - t0
    t["routed_bytes"] = (sum(send) - send[rank]) * (8 * W + 6)
    d_hi = got[1].data_ptr() if W == 2 else None
    res = gpu.sort_routed(params, got[0].data_ptr(), d_hi, got[-2].data_ptr(), got[-1].data_ptr(), n_recv, world, rank)
    for i in range(len(res.stats)):
        res.stats[i] = res1.stats[i]
    for i in range(8):
        res.seconds[i] = res1.seconds[i]
    ks = [float(x) for x in res1.kernel_seconds]
    t["route_kernel"] = float(res1.kernel_seconds[3])
    t["sort_routed"] = float(res.kernel_seconds[3])
    t["sort_diag"] = [float(res.kernel_seconds[i]) for i in (5, 6, 7)]
    ks[3] += float(res.kernel_seconds[3])                 # routing kernels + sort
    ks[5:8] = t["sort_diag"]
    for i in range(8):
        res.kernel_seconds[i] = ks[i]
    res._keep = got                                       # (nothing of it is referenced by the result, but keep the order of frees simple)
    return res


N_PIECES = 4      # a rank partitions its reads in this many pieces: piece i travels while piece i+1 is being partitioned
MAX_SOURCES = 32  # GATB_GPU_MAX_SOURCES (include/gatb_gpu.h): ranks x pieces


def pieces_per_rank(world, n_reads_local):
    """Pieces a rank cuts its reads into (only the last piece's exchange is exposed, so more pieces hide more of it)"""
    if world <= 1 or n_reads_local < 1024:
        return 1
    npc = N_PIECES
    while npc > 1 and world * npc > MAX_SOURCES:
        npc //= 2
    return npc


def piece_bounds(n_reads_local, npc):
    """First read of every piece (multiples of 32 reads) and the end"""
    return [(n_reads_local * i // npc) & ~31 for i in range(npc)] + [n_reads_local]


def bind_to_gpu_numa(local):
    """Pins this process to the CPUs next to its GPU (NVML affinity) BEFORE it allocates pinned host memory: eight ranks whose
    staging buffers sit on the wrong socket share one inter-socket link for all their host<->device copies."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        n_words = (os.cpu_count() + 63) // 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def count_distributed(gpu, params, d_reads, n_reads_local, n_reads_global, total_kmers_global, rank, world, repart=None,
                      d_offsets=None, timers=None, ready=None, route=False):
    """One distributed counting pass.  Returns (device Result of this rank's bins, stats dict with GLOBAL sums).

    partition piece 0 -> [send piece 0 || partition piece 1] -> ... -> send the last piece -> count: every (source rank, piece)
    is one source of the owner's fine split (gatb_gpu_count_bins, at most 32 sources).
    ready: optional list of torch events, one per piece (piece_bounds): the reads of piece i are on the device once ready[i] has
    completed (the caller's host->device copies run on their own stream while the earlier pieces are being partitioned).
    route: second exchange (_count_and_route): every partition key ends up whole and ascending on rank key % world."""
    dev = torch.device("cuda", gpu.device)
    geom = gpu.plan(params, total_kmers_global, n_reads_global, world)
    npc = pieces_per_rank(world, n_reads_local)
    t = {"partition": 0.0, "exchange_wait": 0.0}
    nb1, rb, bpr, blk = geom.nb1, geom.record_bytes, geom.bins_per_rank, geom.coarse_blk
    if npc > 1:                                          # a piece holds 1/npc of the records of a bin: shrink the capacity with it
        geom.cap = (int(geom.cap / npc * 1.15) + 64 + blk - 1) // blk * blk
    firsts = piece_bounds(n_reads_local, npc)
    st_sum = [0, 0, 0, 0]
    all_pieces, all_cur, keep, used_total = [], [], [], 0
    pending = None
    t_begin = time.time()
    for i in range(npc):
        n_i = firsts[i + 1] - firsts[i]
        if ready is not None:
            ready[i].synchronize()
        while True:
            cap = geom.cap
            bins = torch.empty(nb1 * cap * rb, dtype=torch.uint8, device=dev)
            cursors = torch.zeros(nb1, dtype=torch.int32, device=dev)
            torch.cuda.current_stream().synchronize()    # torch's fills run on torch's stream, the library on its own
            t0 = time.time()
            st = gpu.partition_into(params, geom, d_reads, d_offsets, n_i, bins.data_ptr(), cursors.data_ptr(), first_read=firsts[i])
            t["partition"] += time.time() - t0
            flag = torch.tensor([st[3], int(cursors.max().item())], dtype=torch.int64, device=dev)
            if world > 1:
                dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            if int(flag[0].item()) == 0:
                break
            geom.cap = (int(flag[1].item()) + blk - 1) // blk * blk     # a bin overflowed somewhere: every rank re-runs this piece with the global demand
            del bins, cursors
        for k in range(4):
            st_sum[k] += st[k]
        if pending is not None:                          # the previous piece has had the whole partition of this one to travel
            t0 = time.time()
            for req in pending:
                req.wait()
            t["exchange_wait"] += time.time() - t0
        # region r is filled round by round (blk records per bin and round): only the rounds in use travel
        rounds = (cursors.view(world, bpr).max(dim=1).values.clamp(max=cap).to(torch.int64) + blk - 1) // blk
        used = (rounds * (bpr * blk * rb)).tolist()
        pieces, recv_cur, pending = exchange_bins(bins, cursors, world, rank, used, wait=False)
        used_total += int(sum(used) - used[rank])
        keep.append((bins, cursors, geom.cap))
        all_pieces.append(pieces)
        all_cur.append(recv_cur)
    t0 = time.time()
    for req in pending:
        req.wait()
    torch.cuda.synchronize()
    t["exchange_wait"] += time.time() - t0
    t["exchange"] = t["exchange_wait"]
    t["partition_and_exchange"] = time.time() - t_begin
    t["exchanged_bytes"] = used_total
    geom.cap = max(c for (_, _, c) in keep)              # the fine split only clamps cursors with it: the largest piece capacity serves all
    st = st_sum
    src_bins = [all_pieces[i][s].data_ptr() for i in range(npc) for s in range(world)]
    src_cur = [all_cur[i][s].data_ptr() for i in range(npc) for s in range(world)]
    recv_cur = torch.cat([c.reshape(-1) for c in all_cur])
    gathered = int(recv_cur.clamp(max=geom.cap).sum().item())
    tot = torch.tensor([st[0], st[2]], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
    avg_len = float(tot[0].item()) / max(float(tot[1].item()), 1.0)       # k-mers per record, whole job
    # k-mers in the bins this rank owns: estimate with 25 % head-room, never above the hard bound
    kmers_bound = int(min(gathered * geom.maxlen, gathered * avg_len * 1.25 + 65536))
    t0 = time.time()
    if world > 1 and route:
        res = _count_and_route(gpu, params, geom, src_bins, src_cur, bpr, kmers_bound, repart, rank, world, dev, t)
    else:
        res = gpu.count_bins(params, geom, src_bins, src_cur, bpr, kmers_bound, repart=repart)
    t["count"] = time.time() - t0
    t["count_kernels"] = [float(x) for x in res.kernel_seconds][:5]
    t["overflow_kmers"] = int(res.stats[11])
    t["count_stages"] = [float(x) for x in res.seconds][:7]
    t["overflow_bins"] = int(res.stats[8])
    t["owned_records"], t["owned_distinct"] = int(res.stats[4]), int(res.stats[2])
    t["bins"] = int(res.stats[7])
    sums = torch.tensor([st[0], st[1], int(res.stats[2]), int(res.stats[3]), st[2], int(res.n_items)], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(sums)
    stats = dict(zip(["kmers_nb_valid", "kmers_nb_invalid", "kmers_nb_distinct", "kmers_nb_solid", "records", "items"],
                     [int(x) for x in sums.tolist()]))
    stats["exchanged_bytes_per_rank"] = t["exchanged_bytes"]
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
    from bench import K, M, L, ABUNDANCE_MIN, COVERAGE, SEED, METRIC, UNIT, ClockSampler, measured_peak
    bind_to_gpu_numa(local)
    gpu = gatb_core_b200.GatbGpu(local)
    dev = torch.device("cuda", local)
    n = args.reads
    n_global = n * world
    genome = n_global * L // COVERAGE
    nbytes = (n * L + 3) // 4
    reads = torch.zeros(nbytes + 64, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()          # the zero fill runs on torch's stream, the library works on its own: order them
    gpu.synth_reads_dev(SEED, genome, rank * n, n, L, reads.data_ptr())
    gpu.synchronize()
    # the partitioning the reference would use for the WHOLE job (bench.reference_configuration: ConfigurationAlgorithm's
    # arithmetic + the reference's RepartitorAlgorithm on a sample), computed on rank 0 outside the timed region and broadcast
    from bench import reference_configuration
    total_kmers = n_global * (L - K + 1)
    if rank == 0:
        cfg = reference_configuration(gpu, n_global, args)
    else:
        cfg = None
    box = [cfg]
    dist.broadcast_object_list(box, src=0)
    nb_passes, nb_partitions, repart, repart_src = box[0]
    params = gpu.make_params(K, M, nb_partitions=nb_partitions, nb_passes=nb_passes, abundance_min=ABUNDANCE_MIN, read_len=L,
                             path_flags=args.path_flags, bin_load_pct=args.bin_load_pct, table_log2=args.table_log2, fine_bits=args.fine_bits, bin_target_pct=args.bin_target_pct)

    def step(timers=None):
        res, stats = count_distributed(gpu, params, reads.data_ptr(), n, n_global, total_kmers, rank, world, repart=repart, timers=timers, route=args.route)
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
    copy_stream = torch.cuda.Stream(device=dev)
    e2e = []
    d2h_bytes = 0
    for i in range(1 + args.steps):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.time()
        # the reads go up piece by piece on a copy stream; piece i is partitioned as soon as it has landed
        bounds = [b * L // 4 for b in piece_bounds(n, pieces_per_rank(world, n))]
        bounds[-1] = reads.numel()
        ready = []
        with torch.cuda.stream(copy_stream):
            for a, b in zip(bounds[:-1], bounds[1:]):
                reads[a:b].copy_(h_reads[a:b], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                ready.append(ev)
        t1 = time.time()
        res, stats = count_distributed(gpu, params, reads.data_ptr(), n, n_global, total_kmers, rank, world, repart=repart, ready=ready, route=args.route)
        t2 = time.time()
        host = result_to_pinned(gpu, res, params)
        gpu.result_free(res)
        if rank == 0 and os.environ.get("GATB_BENCH_DEBUG"):
            print("e2e step %d: h2d %.1f ms, count %.1f ms, d2h %.1f ms" % (i, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (time.time() - t2) * 1e3), file=__import__("sys").stderr)
        d2h_bytes = int(host["n_items"]) * 12 + (10001 + nb_passes * nb_partitions + 1) * 8
        el2 = torch.tensor([time.time() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(el2, op=dist.ReduceOp.MAX)
        if i >= 1:
            e2e.append(float(el2.item()))
    e2e_step = sum(e2e) / len(e2e)
    clocks = sampler.stop() if rank == 0 else None
    # ---- size-independent invariants of the full-size result (the host arrays of the last end-to-end step; outside every
    #      timed region): the oracle cannot run at this size, these properties can be checked ----
    local, err = [0, 0, 0, 0, 1], 0
    try:
        hist = np.asarray(host["histogram"], dtype=np.int64)
        ni = int(host["n_items"])
        cnt, lo = host["counts"][:ni], host["kmers_lo"][:ni]
        offs = np.asarray(host["part_offsets"], dtype=np.int64)
        desc = np.nonzero(lo[1:] <= lo[:-1])[0] + 1 if ni > 1 else np.zeros(0, np.int64)
        asc = bool(np.isin(desc, offs).all())                # ascending inside every partition key (k <= 31: one word per k-mer)
        sizes = np.diff(offs)
        owned = (np.arange(len(sizes)) % world) == rank      # second exchange: a partition lives whole on rank key % world
        whole = (not args.route) or bool((sizes[~owned] == 0).all())
        local = [int(hist.sum()), sum(c * int(hist[c]) for c in range(ABUNDANCE_MIN)) + int(cnt.sum(dtype=np.int64)),
                 int(hist[ABUNDANCE_MIN:].sum()), ni, 0 if (asc and whole) else 1]
    except Exception:                                        # never let the checker take the measurement down
        err = 1
    loc = torch.tensor(local + [err], dtype=torch.int64, device=dev)
    dist.all_reduce(loc)                                     # every rank gets here, whatever happened above
    loc = [int(x) for x in loc.tolist()]
    invariants = {"sum_hist_eq_distinct": loc[0] == stats["kmers_nb_distinct"], "occurrences_accounted": loc[1] == stats["kmers_nb_valid"],
                  "solid_is_histogram_tail": loc[2] == loc[3] == stats["kmers_nb_solid"], ("partitions_whole_and_ascending_on_their_owner" if args.route else "strictly_ascending_per_rank"): loc[4] == 0,
                  "checker_errors": loc[5]}
    invariants["all"] = all(v is True for k, v in invariants.items() if k != "checker_errors") and loc[5] == 0
    if rank == 0:
        # roofline of the dominant kernel on rank 0 (same accounting as the single-GPU line)
        occ_owned = stats["kmers_nb_valid"] / world
        s_alg = timers["owned_records"] * (1 + (K - 1) / 4.0 + 0.375) + occ_owned / 4.0
        kern = {"k1_superkmer_fast": (n * L / 4.0 + stats["records"] / world * (1 + (K - 1) / 4.0 + 0.375) + occ_owned / 4.0, timers["partition"]),
                "k2b_warp_bins": (s_alg + timers["owned_distinct"] * 12.0, timers["count_kernels"][2])}
        dom = max(kern, key=lambda name: kern[name][1])
        peak, peak_src = measured_peak()
        roofline = {"bound": "hbm", "kernel": dom, "rank": 0, "achieved": kern[dom][0] / kern[dom][1] / 1e9, "peak": peak, "unit": "GB/s",
                    "frac": kern[dom][0] / kern[dom][1] / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": kern[dom][0], "ms_per_launch": kern[dom][1] * 1e3}
        line = {"metric": METRIC, "value": stats["kmers_nb_distinct"] / per_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u64", "data": "synthetic",
                "config": {"workload": "k=31, %d synthetic 150bp reads (%d per GPU), %dxB200, minimizer buckets sharded via NCCL all-to-all, m=10, abundance-min=2, "
                                       "%d partitions x %d pass(es) (the reference's own configuration)%s" % (n_global, n, world, nb_partitions, nb_passes,
                                          "" if not args.route else "; second all-to-all: every partition whole and ascending on rank key %% %d" % world),
                           "reads": n_global, "nb_partitions": nb_partitions, "nb_passes": nb_passes, "repartitor": repart_src, "genome_nt": genome, "coverage": COVERAGE, "error_rate": 0.01,
                           "l2": "per-GPU inputs (%.1f GB packed reads) far exceed the 126 MB L2" % (nbytes / 1e9)},
                "input_bases_per_s": n_global * L / per_step, "kmer_occurrences_per_s": stats["kmers_nb_valid"] / per_step,
                "distinct": stats["kmers_nb_distinct"], "solid": stats["kmers_nb_solid"], "records": stats["records"],
                "stage_seconds_rank0_last_step": timers, "exchanged_bytes_per_rank": stats["exchanged_bytes_per_rank"],
                "e2e": {"value": stats["kmers_nb_distinct"] / e2e_step, "unit": UNIT, "h2d_bytes_per_step": nbytes * world,
                        "d2h_bytes_per_step": d2h_bytes * world, "ms_per_step": e2e_step * 1e3},
                "gpu_launches": int(launches) * world, "clocks": clocks, "roofline": roofline, "cpu_baseline": None,
                "invariants": invariants}
        args.emit(line)
    gpu.close()
    dist.destroy_process_group()
