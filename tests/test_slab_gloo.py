"""CPU test of the multi-GPU decomposition rules with world_size 2 over gloo.

Each rank keeps only the agents of its slab plus ghosts, advances them with the plain-C
oracle (grid mode, i.e. the kernels' neighbour order) and exchanges halo / migrating agents
with its neighbour exactly by the rules the CUDA runtime implements (openabl_b200.slab:
slab_layer, send_masks, owned_mask ⇔ k_slab_classify / exchange in abl_runtime.cu).  The
union of the slabs must equal the undecomposed run bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _exchange(rank, world, send_lo, send_hi, dtype):
    """Swaps variable-length record arrays with the lower / upper neighbour."""
    got = []
    for peer, payload in ((rank - 1, send_lo), (rank + 1, send_hi)):
        if peer < 0 or peer >= world:
            continue
        raw = torch.from_numpy(np.frombuffer(payload.tobytes(), dtype=np.uint8).copy())
        n_out = torch.tensor([len(payload)], dtype=torch.int64)
        n_in = torch.zeros(1, dtype=torch.int64)
        # lower rank sends first to avoid a deadlock of blocking gloo sends
        if rank < peer:
            dist.send(n_out, peer); dist.recv(n_in, peer)
        else:
            dist.recv(n_in, peer); dist.send(n_out, peer)
        buf = torch.zeros(int(n_in.item()) * dtype.itemsize, dtype=torch.uint8)
        if rank < peer:
            if len(raw): dist.send(raw, peer)
            if len(buf): dist.recv(buf, peer)
        else:
            if len(buf): dist.recv(buf, peer)
            if len(raw): dist.send(raw, peer)
        got.append(np.frombuffer(buf.numpy().tobytes(), dtype=dtype))
    return np.concatenate(got) if got else np.zeros(0, dtype=dtype)


def _rank_main(rank, world, port, n, steps, out_dir):
    import sys
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    from oracle import GRID, Oracle
    from openabl_b200.slab import owned_mask, send_masks, slab_layer, split_layers

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    o = Oracle(False)
    full = o.boids_init(n)
    # grid of boids2d: max_pos = sqrt(n / 500) (integer division), cell = interaction radius
    max_pos = np.sqrt(float(n // 500))
    cell = 0.05
    n_layers = int(np.ceil(max_pos / cell))
    bounds = split_layers(n_layers, world)
    begin, end = bounds[rank]

    rec_dt = np.dtype([("id", np.uint32), ("agent", full.dtype)])
    layer = slab_layer(full["pos"][:, 1], 0.0, cell, n_layers)
    mine = owned_mask(layer, begin, end)
    owned = np.zeros(int(mine.sum()), dtype=rec_dt)
    owned["id"] = np.nonzero(mine)[0]
    owned["agent"] = full[mine]

    def refresh(owned):
        """halo + migration exchange; returns the local population (owned + ghosts + arrivals)"""
        lay = slab_layer(owned["agent"]["pos"][:, 1], 0.0, cell, n_layers)
        to_lo, to_hi = send_masks(lay, bounds, rank, 1)
        arrivals = _exchange(rank, world, owned[to_lo], owned[to_hi], rec_dt)
        local = np.concatenate([owned, arrivals])
        return local[np.argsort(local["id"], kind="stable")]   # index order == id order

    local = refresh(owned)
    for _ in range(steps):
        state = np.ascontiguousarray(local["agent"])
        nxt = o.boids_run(state, 1, GRID, num_agents=n)      # all local agents; ghosts are discarded
        lay_before = slab_layer(local["agent"]["pos"][:, 1], 0.0, cell, n_layers)
        keep = owned_mask(lay_before, begin, end)            # results are valid for owned agents only
        owned = np.zeros(int(keep.sum()), dtype=rec_dt)
        owned["id"] = local["id"][keep]
        owned["agent"] = nxt[keep]
        local = refresh(owned)
    lay = slab_layer(local["agent"]["pos"][:, 1], 0.0, cell, n_layers)
    final = local[owned_mask(lay, begin, end)]
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), final)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_slabs_over_gloo_equal_undecomposed_run(tmp_path):
    import sys
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    from oracle import GRID, Oracle
    n, steps, world = 20000, 5, 2
    port = _free_port()
    mp.spawn(_rank_main, args=(world, port, n, steps, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npy" % r)) for r in range(world)]
    merged = np.concatenate(parts)
    merged = merged[np.argsort(merged["id"], kind="stable")]
    assert np.array_equal(merged["id"], np.arange(n, dtype=np.uint32)), "agents lost or duplicated"
    o = Oracle(False)
    want = o.boids_run(o.boids_init(n), steps, GRID)
    for f in want.dtype.names:
        assert np.array_equal(merged["agent"][f], want[f]), "member %s differs" % f


def _ids_main(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, REPO)
    from openabl_b200.slab import all_gather_ids, global_add_ranks
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    # parents of one adding step: rank 0 owns ids 5, 9, 40; rank 1 owns 7 only; then an empty round
    rounds = [([5, 9, 40], [7]), ([], []), ([], [3, 4])]
    next_id = 100
    log = []
    for parts in rounds:
        mine = np.array(parts[rank], dtype=np.uint32)
        ranks, total = global_add_ranks(all_gather_ids(dist, world, mine))
        log.append(((next_id + ranks[rank]).tolist(), total))
        next_id += total
    np.save(os.path.join(out_dir, "ids%d.npy" % rank), np.array(log, dtype=object), allow_pickle=True)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_new_agent_ids_are_resolved_across_ranks(tmp_path):
    """Run-time add() under slab decomposition: every rank numbers its new agents
    next_id + (rank of the parent among the parents of all ranks) and all ranks advance next_id
    by the global total — the numbering of an undecomposed run (k_append: parent-id order)."""
    world = 2
    port = _free_port()
    mp.spawn(_ids_main, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = [np.load(os.path.join(str(tmp_path), "ids%d.npy" % r), allow_pickle=True).tolist() for r in range(world)]
    assert got[0] == [[[100, 102, 103], 4], [[], 0], [[], 2]]
    assert got[1] == [[[101], 4], [[], 0], [[104, 105], 2]]


# ---- scalable upload: every rank uploads 1/N of the population by index, one all-to-all routes the
# ---- records to their owners (RankSlab.upload; rules of k_route_classify restated in numpy) -----------
def _route_rank_main(rank, world, port, n, out_dir):
    import sys
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    from oracle import Oracle
    from openabl_b200.slab import exchange_partitions, owner_of, slab_layer, split_layers

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    o = Oracle(False)
    pop = o.boids_init(n)                          # every process can rebuild the population; it touches only its part
    W = float(np.sqrt(n / 500.0)) if n else 1.0
    cell = 0.05 * (1.0 + 2.0 ** -20)               # ABL_CELL_PAD_F64
    layers = max(1, int(np.ceil(o.lib.oracle_fold6(W) / cell)))
    bounds = split_layers(layers, world)
    lo, hi = n * rank // world, n * (rank + 1) // world
    part = pop[lo:hi]
    owner = owner_of(slab_layer(part["pos"][:, 1], 0.0, cell, layers), bounds)
    order = np.argsort(owner, kind="stable")
    counts = [int((owner == d).sum()) for d in range(world)]
    transit = np.dtype([("rec", pop.dtype), ("id", np.uint32)])
    send = np.zeros(len(part), dtype=transit)
    send["rec"] = part[order]
    send["id"] = (lo + order).astype(np.uint32)
    raw = torch.from_numpy(np.frombuffer(send.tobytes(), dtype=np.uint8).copy()) if len(send) else torch.zeros(1, dtype=torch.uint8)
    recv, total = exchange_partitions(dist, rank, world, raw, counts, transit.itemsize)
    got = np.frombuffer(recv.numpy()[:total * transit.itemsize].tobytes(), dtype=transit)
    np.save(os.path.join(out_dir, "ids_%d.npy" % rank), got["id"])
    np.save(os.path.join(out_dir, "pos_%d.npy" % rank), got["rec"]["pos"])
    dist.destroy_process_group()


def test_scalable_upload_routes_every_agent_to_its_owner(tmp_path):
    import sys
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    from oracle import Oracle
    from openabl_b200.slab import owner_of, slab_layer, split_layers
    n, world = 20000, 2
    mp.spawn(_route_rank_main, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    o = Oracle(False)
    pop = o.boids_init(n)
    cell = 0.05 * (1.0 + 2.0 ** -20)
    layers = int(np.ceil(o.lib.oracle_fold6(float(np.sqrt(n / 500.0))) / cell))
    owner = owner_of(slab_layer(pop["pos"][:, 1], 0.0, cell, layers), split_layers(layers, world))
    seen = []
    for r in range(world):
        ids = np.load(os.path.join(str(tmp_path), "ids_%d.npy" % r))
        pos = np.load(os.path.join(str(tmp_path), "pos_%d.npy" % r))
        assert np.array_equal(np.sort(ids), np.nonzero(owner == r)[0]), "rank %d did not receive exactly its agents" % r
        assert np.array_equal(pos, pop["pos"][ids]), "records do not travel with their ids"
        seen.append(ids)
    assert len(np.concatenate(seen)) == n
