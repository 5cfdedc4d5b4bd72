"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo processes exercise the all-to-all plumbing
(gatb_core_b200.multigpu.exchange_bins) and the merge of per-rank sorted runs.  No CUDA kernels run here."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gatb_core_b200 import multigpu


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bpr, cap, rb, fb = 3, 4, 16, 2
        nb1 = bpr * world
        # record (b, i) of source `rank` carries the bytes [rank, b, i, 0...]
        bins = torch.zeros(nb1, cap, rb, dtype=torch.uint8)
        for b in range(nb1):
            for i in range(cap):
                bins[b, i, 0], bins[b, i, 1], bins[b, i, 2] = rank, b, i
        cursors = torch.tensor([(b + rank) % (cap + 1) for b in range(nb1)], dtype=torch.int32)
        # rank r only uses (and sends) the first `used[r]` bytes of each region: a whole number of bins here
        used = [(1 + (rank + r) % bpr) * cap * rb for r in range(world)]
        pieces, rcur = multigpu.exchange_bins(bins.view(-1), cursors, world, rank, used)
        ok = len(pieces) == world and rcur.shape == (world, bpr)
        for s in range(world):
            nbin = bpr if s == rank else 1 + (s + rank) % bpr                   # own region: a view of everything
            ok &= pieces[s].numel() >= nbin * cap * rb
            piece = torch.zeros(bpr, cap, rb, dtype=torch.uint8)
            piece.view(-1)[:nbin * cap * rb] = pieces[s][:nbin * cap * rb]
            for lb in range(bpr):
                gb = rank * bpr + lb                                  # global id of the bin this rank owns
                assert multigpu.owner_of_bin(gb, bpr) == rank
                if lb < nbin:
                    ok &= bool((piece[lb, :, 0] == s).all() and (piece[lb, :, 1] == gb).all())
                ok &= int(rcur[s, lb]) == (gb + s) % (cap + 1)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_exchange_bins_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_exchange_bins_world1_is_identity():
    bins = torch.arange(4 * 3 * 16, dtype=torch.uint8)
    cur = torch.tensor([1, 2, 3, 0], dtype=torch.int32)
    a, b = multigpu.exchange_bins(bins, cur, 1)
    assert len(a) == 1 and (a[0].view(-1) == bins).all() and (b.view(-1) == cur).all()


def test_merge_sorted_runs():
    rng = np.random.default_rng(0)
    full = np.unique(rng.integers(0, 2 ** 62, 5000, dtype=np.uint64))
    hi = rng.integers(0, 4, len(full)).astype(np.uint64)
    order = np.lexsort((full, hi))
    full, hi = full[order], hi[order]
    cn = rng.integers(1, 100, len(full)).astype(np.int32)
    owner = rng.integers(0, 3, len(full))
    runs = [(full[owner == r], hi[owner == r], cn[owner == r]) for r in range(3)]
    lo2, hi2, cn2 = multigpu.merge_sorted_runs(runs)
    assert (lo2 == full).all() and (hi2 == hi).all() and (cn2 == cn).all()


def test_pieces_and_bounds():
    # a rank cuts its reads into pieces so that ranks x pieces stays within the sources a count accepts; pieces start on multiples of 32 reads
    assert multigpu.pieces_per_rank(1, 10 ** 8) == 1
    assert multigpu.pieces_per_rank(8, 100) == 1                        # tiny inputs are not cut
    for world in (2, 4, 8):
        npc = multigpu.pieces_per_rank(world, 10 ** 8)
        assert npc >= 1 and world * npc <= multigpu.MAX_SOURCES
        b = multigpu.piece_bounds(10 ** 8 + 17, npc)
        assert b[0] == 0 and b[-1] == 10 ** 8 + 17 and len(b) == npc + 1
        assert all(x % 32 == 0 for x in b[:-1]) and all(b[i] < b[i + 1] for i in range(npc))


def test_owner_of_routed_keys():
    # second exchange: key k lives on rank k % world as the (k // world)-th key of that rank; the offsets of all keys follow from the
    # local ones (the arithmetic gatb_gpu_sort_routed does on the host)
    for n_keys, world in ((5, 8), (176, 2), (2816, 8), (7, 3)):
        for rank in range(world):
            owned = [k for k in range(n_keys) if k % world == rank]
            n_local = (n_keys - rank + world - 1) // world if n_keys > rank else 0
            assert n_local == len(owned)
            sizes = np.arange(1, n_local + 1)
            loc = np.concatenate([[0], np.cumsum(sizes)]) if n_local else np.zeros(1, np.int64)
            offs = []
            for key in range(n_keys + 1):
                j = (key - rank + world - 1) // world if key > rank else 0
                offs.append(int(loc[min(j, n_local)]))
            got = np.diff(offs)
            for key in range(n_keys):
                assert got[key] == (sizes[key // world] if key % world == rank else 0), (n_keys, world, rank, key)
