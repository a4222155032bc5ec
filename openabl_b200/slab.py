"""Slab decomposition harness: one model, several device runtimes.

  * `LocalSlabs`  — S slabs in ONE process on one GPU.  transport="staged": the runtime's
    in-process transport (abl_cuda_set_local_peers / exchange_begin / exchange_end);
    transport="direct": the same peer-memory transport the multi-GPU runs use
    (abl_cuda_halo_setup / halo_connect_local) — step kernels append their halo records to the
    neighbouring slab's receive area and nothing synchronises with the host.  This is how the
    single-GPU test tier checks that a decomposed run is bit-identical to an undecomposed one.
  * `RankSlab`    — one slab per process / GPU (torchrun); used by bench.py --gpus N.
    transport="direct" (default): receive areas are exchanged as CUDA IPC handles and written
    over NVLink by the step kernels; transport="nccl": grouped ncclSend/ncclRecv inside
    abl_cuda_step (abl_cuda_comm_init_nccl).

The partition is by whole cell layers along the slowest grid axis, balanced by layer count
(`split_layers`); ownership and ghosts are maintained by the runtime (abl_runtime.cu).
"""
import ctypes as C

import numpy as np

from .runtime import Runtime, check


def split_layers(n_layers, parts):
    """-> [(begin, end)] contiguous, as even as possible, every part non-empty."""
    if parts > n_layers:
        raise ValueError("more slabs (%d) than cell layers (%d)" % (parts, n_layers))
    bounds = [(n_layers * r) // parts for r in range(parts + 1)]
    return [(bounds[r], bounds[r + 1]) for r in range(parts)]


def merge_by_id(parts_ids, parts_records):
    """Concatenates per-slab downloads and orders them by agent id."""
    ids = np.concatenate(parts_ids)
    rec = np.concatenate(parts_records)
    order = np.argsort(ids, kind="stable")
    return ids[order], rec[order]


def halo_capacity(total_agents, n_layers, slack=4):
    """Records one halo message must hold: `slack` x the mean population of a cell layer."""
    return max(16384, int(slack * total_agents / max(n_layers, 1)) + 4096)


def global_add_ranks(parts):
    """parts: ascending parent-id arrays of the slabs (one adding step).  -> (per-slab ranks of
    the parents in the sorted union, total).  New agents are numbered next_id + rank, exactly
    as an undecomposed run numbers them (parent-id order)."""
    parts = [np.asarray(p, dtype=np.uint32) for p in parts]
    every = np.sort(np.concatenate(parts)) if parts else np.zeros(0, dtype=np.uint32)
    return [np.searchsorted(every, p).astype(np.uint32) for p in parts], int(len(every))


def all_gather_ids(dist, world, mine):
    """Variable-length all-gather of uint32 id arrays over torch.distributed (counts first, then
    padded payloads).  -> list of arrays, one per rank."""
    import torch
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    n = torch.tensor([len(mine)], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    width = max(max(counts), 1)
    buf = torch.zeros(width, dtype=torch.int64, device=dev)
    if len(mine):
        buf[:len(mine)] = torch.from_numpy(np.asarray(mine).astype(np.int64)).to(dev)
    every = [torch.zeros(width, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(every, buf)
    return [e[:c].cpu().numpy().astype(np.uint32) for e, c in zip(every, counts)]


def owner_of(layer, bounds):
    """Slab that owns cell layer `layer` (array) for contiguous bounds [(begin, end)] — the rule of
    k_route_classify (asset/cuda/abl_runtime.cu)."""
    starts = np.array([b for b, _ in bounds])
    return np.searchsorted(starts, layer, side="right") - 1


def exchange_partitions(dist, rank, world, send, counts, rec):
    """All-to-all of transit records.  send: uint8 tensor, the records this rank uploaded grouped by owning
    slab (counts[d] records of `rec` bytes for rank d).  -> (uint8 tensor with the records this rank owns, in
    source-rank order; their number)."""
    import torch
    dev = send.device
    mine = torch.tensor(counts, dtype=torch.int64, device=dev)
    every = [torch.zeros(world, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(every, mine)
    incoming = [int(e[rank].item()) for e in every]       # records rank r holds for me
    total = sum(incoming)
    recv = torch.empty(max(1, total * rec), dtype=torch.uint8, device=dev)
    dist.all_to_all_single(recv[:total * rec], send[:sum(counts) * rec],
                           output_split_sizes=[c * rec for c in incoming],
                           input_split_sizes=[c * rec for c in counts])
    return recv, total


ADDS = 2      # abl_model_step_flags bits
REMOVES = 1


def ring_neighbours(rank, world):
    """(lower, upper) slab of `rank`; slabs form a ring when there are more than two."""
    ring = world > 2
    lower = rank - 1 if rank > 0 else (world - 1 if ring else None)
    upper = rank + 1 if rank + 1 < world else (0 if ring else None)
    return lower, upper


class LocalSlabs:
    def __init__(self, model, n_slabs, transport="staged", halo_records=0, **rt_kw):
        self.model = model
        self.transport = transport
        self.halo_records = halo_records
        self.rts = []
        for _ in range(n_slabs):
            rt = Runtime(use_float=model.use_float, **rt_kw)
            check(model.lib.abl_model_setup(rt.handle), "abl_model_setup")
            self.rts.append(rt)
        layers = self.rts[0].slab_layers()
        self.bounds = split_layers(layers, n_slabs)
        for r, rt in enumerate(self.rts):
            rt.set_slab(self.bounds, r)
        self._connected = False
        if transport == "staged":
            for r, rt in enumerate(self.rts):
                # periodic worlds: the first and the last slab are neighbours too
                lo, hi = ring_neighbours(r, n_slabs)
                rt.set_local_peers(self.rts[lo] if lo is not None else None,
                                   self.rts[hi] if hi is not None else None)

    def _connect_direct(self, total_agents):
        m = self.model
        layers = self.rts[0].slab_layers()
        for t in range(m.n_types):
            cap = self.halo_records or halo_capacity(total_agents[t], layers)
            for rt in self.rts:
                rt.halo_setup(m.pool(t), cap)
            for r, rt in enumerate(self.rts):
                lo, hi = ring_neighbours(r, len(self.rts))
                rt.halo_connect_local(m.pool(t), self.rts[lo] if lo is not None else None,
                                      self.rts[hi] if hi is not None else None)
        self._connected = True

    def upload(self, host_arrays, scalable=False):
        """Every slab receives the whole population and keeps its own part; scalable=True: slab r uploads
        the r-th part by index and the records are routed between the slabs on the device (the path of
        RankSlab.upload and of abl_cuda_group_simulate, with torch copies standing in for the all-to-all)."""
        m = self.model
        if self.transport == "direct" and not self._connected:
            self._connect_direct([len(a) for a in host_arrays])
        if not scalable:
            for rt in self.rts:
                for t, arr in enumerate(host_arrays):
                    rt.upload(m.pool(t), np.ascontiguousarray(arr))
        else:
            import torch
            S = len(self.rts)
            for t, arr in enumerate(host_arrays):
                pool, n = m.pool(t), len(arr)
                rec = self.rts[0].transit_record_bytes(pool)
                sends, counts = [], []
                for r, rt in enumerate(self.rts):
                    lo, hi = n * r // S, n * (r + 1) // S
                    send = torch.empty(max(1, (hi - lo) * rec), dtype=torch.uint8, device="cuda")
                    counts.append(rt.partition_upload(pool, np.ascontiguousarray(arr[lo:hi]), lo, send.data_ptr(), S))
                    sends.append(send)
                for d, rt in enumerate(self.rts):
                    parts = []
                    for r in range(S):
                        before = sum(counts[r][:d]) * rec
                        parts.append(sends[r][before:before + counts[r][d] * rec])
                    recv = torch.cat(parts) if parts else torch.empty(0, dtype=torch.uint8, device="cuda")
                    torch.cuda.synchronize()
                    rt.adopt_records(pool, recv.data_ptr() if recv.numel() else 0, recv.numel() // rec, n)
        for t in range(m.n_types):
            self._exchange(m.pool(t))

    def _exchange(self, pool):
        if self.transport == "direct":
            for rt in self.rts:
                rt.exchange(pool)   # enqueues pack + publish + wait + unpack, returns at once
            return
        for rt in self.rts:
            rt.exchange_begin(pool)
        for rt in self.rts:
            rt.exchange_end(pool)

    def timestep(self):
        m = self.model
        for rt in self.rts:
            rt.begin_timestep()
        for s in range(m.n_steps):
            for rt in self.rts:
                check(m.lib.abl_model_run_step(rt.handle, s), "abl_model_run_step")
            if m.step_flags(s) & ADDS:
                # ids of new agents are global: rank of every parent among the parents of all slabs
                ranks, total = global_add_ranks([rt.pending_add_parents() for rt in self.rts])
                for rt, r in zip(self.rts, ranks):
                    rt.resolve_adds(r, total)
            if self.transport == "direct":
                continue   # abl_cuda_step / resolve_adds exchanged by themselves
            # the step's own pool may have changed: refresh ghosts / migrate
            pool = self.step_pool(s)
            self._exchange(pool)
        for rt in self.rts:
            rt.end_timestep()   # advances the counter the in-step RNG is keyed by

    def step_pool(self, s):
        # pools are registered in agent declaration order; the generated library reports the
        # pool of each step through abl_model_step_pool
        fn = self.model.lib.abl_model_step_pool
        fn.restype = C.c_int
        fn.argtypes = [C.c_int]
        return fn(s)

    def download(self, t):
        m = self.model
        ids = [rt.download_ids(m.pool(t)) for rt in self.rts]
        rec = [rt.download(m.pool(t), m.dtypes[t]) for rt in self.rts]
        return merge_by_id(ids, rec)

    def owned_counts(self, t):
        return [rt.pool_size(self.model.pool(t)) for rt in self.rts]

    def close(self):
        for rt in self.rts:
            rt.close()
        self.rts = []


class RankSlab:
    """One slab per process.  `dist` is torch.distributed (already initialised)."""

    def __init__(self, model, rank, world, dist=None, device=0, transport="direct", halo_records=0, **rt_kw):
        import torch
        self.model = model
        self.rank, self.world = rank, world
        self.dist = dist
        self.transport = transport if world > 1 else "none"
        self.halo_records = halo_records
        self._connected = False
        self.last_upload_bytes = 0      # host -> device bytes of the latest upload() on this rank
        self._dl = {}                   # page-locked download buffers, per agent type
        self.rt = model.create_runtime(device=device, **rt_kw)
        layers = self.rt.slab_layers()
        self.bounds = split_layers(layers, world)
        if self.transport == "nccl":
            uid = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                uid = torch.tensor(list(self.rt.nccl_unique_id()), dtype=torch.uint8)
            uid = uid.cuda(device)
            dist.broadcast(uid, src=0)
            self.rt.init_nccl(bytes(uid.cpu().tolist()), rank, world)
        self.rt.set_slab(self.bounds, rank)
        self._mutating = any(model.step_flags(s) for s in range(model.n_steps))
        if world > 1 and dist is not None:
            self._install_reduce_hook()

    def upload(self, host_arrays=None, scalable=True):
        """scalable (default): rank r uploads the r-th part of every agent array BY INDEX, the runtime
        groups the records by owning slab on the device (abl_cuda_partition_upload), one all-to-all over
        torch.distributed routes them, and every rank adopts what it owns — H2D traffic and host work per
        rank are 1/N of the population.  scalable=False: the round-1 path (every rank uploads everything
        and crops).  Either way one exchange fetches the ghosts."""
        import torch
        m = self.model
        arrays = [m.host_view(t) for t in range(m.n_types)] if host_arrays is None else list(host_arrays)
        if self.transport == "direct" and not self._connected:
            self._connect_direct([len(a) for a in arrays])
        import os
        if os.environ.get("ABL_UPLOAD_SCALABLE", "1") in ("0", ""):
            scalable = False
        if not scalable or self.world == 1 or self.dist is None:
            if host_arrays is None:
                m.upload_host()
            else:
                for t, arr in enumerate(arrays):
                    self.rt.upload(m.pool(t), np.ascontiguousarray(arr))
            self.last_upload_bytes = sum(a.nbytes for a in arrays)
        else:
            self.last_upload_bytes = 0
            dev = torch.device("cuda", torch.cuda.current_device())
            for t, arr in enumerate(arrays):
                pool = m.pool(t)
                if not self._has_position(t):
                    self.rt.upload(pool, np.ascontiguousarray(arr))
                    self.last_upload_bytes += arr.nbytes
                    continue
                n = len(arr)
                lo, hi = n * self.rank // self.world, n * (self.rank + 1) // self.world
                rec = self.rt.transit_record_bytes(pool)
                part = np.ascontiguousarray(arr[lo:hi])
                if host_arrays is None and len(part):
                    self.rt.pin_host(part)         # this rank's part of the model's own array: page-locked once, DMA from then on
                send = torch.empty(max(1, (hi - lo) * rec), dtype=torch.uint8, device=dev)
                counts = self.rt.partition_upload(pool, part, lo, send.data_ptr(), self.world)
                self.last_upload_bytes += part.nbytes
                recv, total = exchange_partitions(self.dist, self.rank, self.world, send, counts, rec)
                torch.cuda.synchronize()
                self.rt.adopt_records(pool, recv.data_ptr(), total, n)
        for t in range(m.n_types):
            self.rt.exchange(m.pool(t))

    def _has_position(self, t):
        return any(is_pos for _, _, is_pos in self.model.agents[t][1])

    def _connect_direct(self, total_agents):
        """Every rank allocates its receive areas and learns its ring neighbours' IPC handles."""
        import torch
        m = self.model
        layers = self.rt.slab_layers()
        lo, hi = ring_neighbours(self.rank, self.world)
        for t in range(m.n_types):
            cap = self.halo_records or halo_capacity(total_agents[t], layers)
            mine = torch.tensor(list(self.rt.halo_setup(m.pool(t), cap)), dtype=torch.uint8)
            every = [torch.zeros(64, dtype=torch.uint8) for _ in range(self.world)]
            if self.dist.get_backend() == "nccl":
                dev = torch.device("cuda", torch.cuda.current_device())
                every = [e.to(dev) for e in every]
                mine = mine.to(dev)
            self.dist.all_gather(every, mine)
            handles = [bytes(e.cpu().tolist()) for e in every]
            self.rt.halo_connect(m.pool(t), handles[lo] if lo is not None else None,
                                 handles[hi] if hi is not None else None)
        self.dist.barrier()
        self._connected = True

    def timestep(self):
        m = self.model
        if not self._mutating or self.world == 1:
            m.timestep()   # abl_cuda_step exchanges by itself (peer memory or NCCL)
            return
        # run-time add(): the step stays open until the parents' ids have been combined across ranks
        self.rt.begin_timestep()
        for s in range(m.n_steps):
            m.run_step(s)
            if m.step_flags(s) & ADDS:
                mine = self.rt.pending_add_parents()
                ranks, total = global_add_ranks(self._all_gather_ids(mine))
                self.rt.resolve_adds(ranks[self.rank], total)
        m.sequential_step()
        self.rt.end_timestep()

    def _all_gather_ids(self, mine):
        return all_gather_ids(self.dist, self.world, mine)

    def _install_reduce_hook(self):
        """count()/sum() of the sequential step become sums over all ranks."""
        import torch
        dist = self.dist
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")

        def hook(ints, reals):
            if ints:
                t = torch.tensor(ints, dtype=torch.int64, device=dev)
                dist.all_reduce(t)
                ints = t.cpu().tolist()
            if reals:
                t = torch.tensor(reals, dtype=torch.float64, device=dev)
                dist.all_reduce(t)
                reals = t.cpu().tolist()
            return ints, reals
        self.rt.set_reduce_hook(hook)

    def owned(self, t):
        return self.rt.pool_size(self.model.pool(t))

    def download_owned(self, t):
        """This rank's owned agents of type `t` (ascending id) into a page-locked buffer that is reused
        from call to call: the device -> host half of an end-to-end simulate()."""
        import torch
        m = self.model
        n = self.owned(t)
        buf = self._dl.get(t)
        need = max(1, n) * m.dtypes[t].itemsize
        if buf is None or buf.numel() < need:
            buf = torch.empty(need + need // 8, dtype=torch.uint8, pin_memory=True)
            self._dl[t] = buf
        out = buf.numpy()[:n * m.dtypes[t].itemsize].view(m.dtypes[t])
        got = self.rt.download_into(m.pool(t), out)
        return out[:got]


# ---- the decomposition rules, stated once more in numpy ------------------------------------
# (reference model of k_slab_classify / exchange in asset/cuda/abl_runtime.cu; exercised on CPU
# by tests/test_slab_gloo.py with world_size 2 over gloo)
def slab_layer(axis_pos, origin, cell, n_layers):
    """Cell layer along the slab axis; same formula as abl_cell_coord."""
    inv = axis_pos.dtype.type(1) / axis_pos.dtype.type(cell)
    layer = np.floor((axis_pos - axis_pos.dtype.type(origin)) * inv).astype(np.int64)
    return np.clip(layer, 0, n_layers - 1)


def send_masks(layer, bounds, me, ghost_layers=1):
    """Which formerly owned agents are copied to the lower / upper peer after a step.
    Still owned and within `ghost_layers` of a boundary -> ghost copy for that neighbour;
    now inside a neighbouring slab -> migration to it (the sender keeps its record, which
    becomes a ghost by key range).  With more than two slabs the peers form a ring, so agents
    teleported from one end of a periodic world to the other reach the slab over there."""
    n = len(bounds)
    begin, end = bounds[me]
    starts = np.array([b for b, _ in bounds])
    owner = np.searchsorted(starts, layer, side="right") - 1
    mine = owner == me
    lo = (mine & (me > 0) & (layer < begin + ghost_layers)) | (owner == me - 1)
    hi = (mine & (me < n - 1) & (layer >= end - ghost_layers)) | (owner == me + 1)
    if n > 2:
        if me == 0:
            lo = lo | (owner == n - 1)
        if me == n - 1:
            hi = hi | (owner == 0)
    return lo, hi


def owned_mask(layer, begin, end):
    return (layer >= begin) & (layer < end)
