"""Hash-partitioned multi-GPU PA-Star: one process per GPU (torchrun), one partition per GPU.

Replaces the reference's distribution layer for the successor path:
  worker_inner reconciliation      pastar/PAStar.cpp:366-396         -> per-destination outboxes filled on the device
  sender / receiver / decoder      pastar/pastar_functions/PAStarSender.cpp:11-112, PAStarReceiver.cpp:11-107,
                                   PAStarMessageProcesser.cpp:12-78  -> one NCCL all-to-all of fixed-width records
  check_stop's two allreduces      pastar/PAStar.cpp:502-519          -> one allreduce(min) of {min open f, best goal g}
  distributed backtrace            pastar_functions/PAStarDistributedBacktrace.cpp:18-214 -> owner lookups + allreduce

Per round, on every rank:   expand own frontier  ->  all-to-all(successors by owner)  ->  dedupe + push  ->  allreduce(min).
There is nothing in flight between rounds, so the reference's optimality-preserving stop (accept the goal only when
no open node anywhere has f < g_goal and no message is in flight) is exactly `min over ranks of min_open_f >= min over
ranks of best_goal_g`.

The engine is the CUDA context (PastarGPU).  The driver only moves bytes; torch.distributed is plumbing (NCCL on GPUs,
gloo in the CPU tests where a test-only engine stands in for the kernels).
"""
import os

import numpy as np

INT_MAX = 2**31 - 1


_SYMM_CACHE = {}


def _symm_buffer(symm, torch, dist, numel, dtype, device):
    """A symmetric-memory tensor + its rendezvous handle, cached per (shape, dtype, device, world).  Every rank creates
    its engines in the same order, so all ranks hit or miss together (the rendezvous is a collective)."""
    key = (int(numel), str(dtype), device.index, dist.get_world_size())
    if key not in _SYMM_CACHE:
        t = symm.empty(int(numel), dtype=dtype, device=device)
        _SYMM_CACHE[key] = (t, symm.rendezvous(t, dist.group.WORLD))
    return _SYMM_CACHE[key]


class CudaEngine:
    """Adapter: PastarGPU step-wise search -> the byte-tensor interface the driver speaks."""

    def __init__(self, gpu, n_parts, part, table_capacity=0, batch_target=0):
        import torch
        self.torch = torch
        self.g = gpu
        self.n_parts, self.part = n_parts, part
        # everything (kernels, NCCL, torch copies) is ordered on torch's current stream: the insert kernel must not
        # start before the all-to-all that fills its inbox has completed
        gpu.set_stream(torch.cuda.current_stream().cuda_stream)
        gpu.search_begin(n_parts, part, table_capacity, batch_target)
        self.xrec = gpu.xrec_stride()
        self.device = torch.device("cuda", torch.cuda.current_device())

    def _wrap(self, ptr, nbytes):
        class _Mem:  # zero-copy view of library-owned device memory
            __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
        return self.torch.as_tensor(_Mem(), device=self.device)

    def round(self, f_limit):
        """Expand this partition's frontier; returns the per-destination outboxes as uint8 tensors."""
        self.g.search_round(f_limit)
        out = []
        for dst in range(self.n_parts):
            if dst == self.part:
                out.append(self.torch.empty(0, dtype=self.torch.uint8, device=self.device))
                continue
            ptr, n = self.g.search_outbox(dst)
            out.append(self._wrap(ptr, n * self.xrec) if n else self.torch.empty(0, dtype=self.torch.uint8, device=self.device))
        return out

    def insert(self, buf):
        if buf.numel():
            self.g.search_insert_dev(buf.data_ptr(), buf.numel() // self.xrec)

    def status(self):
        return self.g.search_status()

    def lookup(self, pos):
        return self.g.search_lookup(pos)

    def empty(self, nbytes):
        return self.torch.empty(nbytes, dtype=self.torch.uint8, device=self.device)

    def end(self):
        self.g.search_end()


class CudaEngineP2P(CudaEngine):
    """Fused expansion + exchange, device-driven: the expand kernel stores remote successors straight into the owner's
    inbox over NVLink (peer-mapped symmetric memory) while it computes, a one-thread-per-destination kernel stores the
    record counts next to them, one cross-GPU barrier runs on the stream (symmetric-memory signal pads), and the insert
    kernels read their counts from device memory.  No collective and no host round trip per round; inboxes and count
    arrays are double-buffered so one barrier per round is enough.  Same records and insert kernel as CudaEngine."""

    async_rounds = True

    def __init__(self, gpu, n_parts, part, dist, table_capacity=0, batch_target=0, forward=True):
        import torch
        import torch.distributed._symmetric_memory as symm
        self.forward = forward
        self.torch, self.g, self.n_parts, self.part = torch, gpu, n_parts, part
        if n_parts > 16:
            raise ValueError("P2P mode supports up to 16 partitions")
        self.device = torch.device("cuda", torch.cuda.current_device())
        gpu.set_stream(torch.cuda.current_stream().cuda_stream)
        gpu.search_begin(n_parts, part, table_capacity, batch_target, p2p=2 if forward else 1)
        self.xrec = gpu.xrec_stride() if not forward else 8 * (2 if gpu.xrec_stride() == 24 else 3)
        self.region = gpu.search_region_bytes()                        # bytes one source may write into one inbox
        # peer-mapped buffers are kept for the life of the process and reused by later searches of the same shape: the
        # allocation + handle exchange (a collective over the store) costs ~0.4 s, more than a sub-second search
        self.inbox, self.hdl = _symm_buffer(symm, torch, dist, 2 * n_parts * self.region, torch.uint8, self.device)
        self.counts, self.hdl_c = _symm_buffer(symm, torch, dist, 2 * n_parts, torch.int64, self.device)
        self.counts.zero_()  # stamps of an earlier search must not be taken for this one's
        gpu.search_set_peers([int(self.hdl.buffer_ptrs[r]) for r in range(n_parts)])
        gpu.search_set_peer_counts([int(self.hdl_c.buffer_ptrs[r]) for r in range(n_parts)], 2)
        # data-flow synchronisation (default): the counts carry the round and the receiver waits for them on the device, so
        # there is no barrier kernel between a round's two halves; PG_DEVICE_SYNC=0 goes back to one signal-pad barrier
        self.device_sync = os.environ.get("PG_DEVICE_SYNC", "1") != "0"
        if self.device_sync:
            gpu.search_set_device_sync(True)
        self.rounds_done, self._graph, self._graph_key = 0, None, None
        self.sent = self._wrap64(gpu.search_outbox_counts_dev(), n_parts)   # this round's records per destination
        self.sent_total = torch.zeros(1, dtype=torch.int64, device=self.device)
        torch.cuda.synchronize()
        dist.barrier()

    def _wrap64(self, ptr, n):
        class _Mem:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}
        return self.torch.as_tensor(_Mem(), device=self.device)

    @property
    def bytes_sent(self):
        return int(self.sent_total.item()) * self.xrec

    def round_and_exchange(self, f_limit, dist):
        self.g.search_round_async(f_limit)     # remote successors + their counts are on their way to the owners
        self.sent_total += self.sent.sum()     # bookkeeping only (device side)
        if not self.device_sync:
            self.hdl.barrier(channel=0)        # every partition's stores of this round are complete and visible
        self.g.search_insert_inbox_async()     # (device sync: waits for the sources' counts first) dedupe + push; flips the buffers
        self.rounds_done += 1

    GRAPH_ROUNDS = 8

    def rounds(self, f_limit, dist, n):
        """n rounds.  With device-side synchronisation a round is the same launches with the same arguments every second
        round (the double-buffered inboxes alternate; the exchange stamp lives in the device control block), so groups
        of GRAPH_ROUNDS rounds are captured once as a CUDA graph and replayed: no launch gaps between the nine small
        launches of a round.  Per-launch profiling, a changed f limit and odd leftovers take the plain path."""
        torch = self.torch
        use_graph = (self.device_sync and n >= self.GRAPH_ROUNDS and not getattr(self.g, "profiling", False)
                     and os.environ.get("PG_NO_GRAPH") is None and self.rounds_done >= 2)
        done = 0
        if use_graph:
            if self._graph is None or self._graph_key != (f_limit, self.rounds_done & 1):
                if (self.rounds_done & 1) != 0:  # graphs start on an even round: the buffer halves repeat with period 2
                    self.round_and_exchange(f_limit, dist)
                    done += 1
                if n - done >= self.GRAPH_ROUNDS:
                    self._capture(f_limit, dist)
            while self._graph is not None and self._graph_key == (f_limit, self.rounds_done & 1) and n - done >= self.GRAPH_ROUNDS:
                self._graph.replay()
                self.rounds_done += self.GRAPH_ROUNDS
                self.g.search_note_rounds(self.GRAPH_ROUNDS)
                done += self.GRAPH_ROUNDS
        for _ in range(n - done):
            self.round_and_exchange(f_limit, dist)

    def _capture(self, f_limit, dist):
        torch = self.torch
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        graph = torch.cuda.CUDAGraph()
        before = self.rounds_done
        with torch.cuda.graph(graph, stream=side):
            self.g.set_stream(side.cuda_stream)   # the library launches on the capturing stream
            for _ in range(self.GRAPH_ROUNDS):
                self.round_and_exchange(f_limit, dist)
        self.g.set_stream(cur.cuda_stream)
        cur.wait_stream(side)
        # capturing launched nothing: the host-side round bookkeeping moves back (GRAPH_ROUNDS is even, so the buffer half
        # the library would use next is unchanged)
        self.rounds_done = before
        self.g.search_note_rounds(-self.GRAPH_ROUNDS)
        self._graph, self._graph_key = graph, (f_limit, before & 1)

    def status(self):
        self.g.search_sync()
        return self.g.search_status()


class PartitionedSearch:
    """The per-rank loop.  `dist` is torch.distributed (initialised by the caller), engine as above."""

    def __init__(self, engine, dist, seqs, owner_fn, max_expansions=0):
        import torch
        self.torch, self.e, self.dist = torch, engine, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.seqs = seqs
        self.owner_fn = owner_fn  # coord -> owning partition (Coord::get_id(world))
        self.max_expansions = max_expansions
        self.rounds = 0
        self.bytes_sent = 0
        self.rounds_per_status = 8  # one graph-replayed group (CudaEngineP2P.GRAPH_ROUNDS) between status exchanges

    def _dev(self):
        return getattr(self.e, "device", self.torch.device("cpu"))

    def exchange(self, outboxes):
        """All-to-all of variable-length record buffers: counts first, then the payload."""
        t, dist = self.torch, self.dist
        send_n = t.tensor([b.numel() for b in outboxes], dtype=t.int64, device=self._dev())
        recv_n = t.empty_like(send_n)
        dist.all_to_all_single(recv_n, send_n)
        recv_n = recv_n.tolist()
        send_l = send_n.tolist()
        inbox = self.e.empty(int(sum(recv_n)))
        # all_to_all_single with split sizes (the list form is not implemented by gloo); the outboxes live in one
        # allocation with gaps, so they are compacted into one send buffer first
        payload = t.cat(outboxes) if sum(send_l) else self.e.empty(0)
        dist.all_to_all_single(inbox, payload, output_split_sizes=recv_n, input_split_sizes=send_l)
        self.bytes_sent += int(sum(send_l))
        return inbox

    def step(self, f_limit=INT_MAX, rounds=1):
        """`rounds` rounds, then one status exchange; returns (global min open f, global best goal g, global counters)."""
        t, dist = self.torch, self.dist
        if hasattr(self.e, "rounds"):
            self.e.rounds(f_limit, dist, rounds)  # device-driven, groups of rounds replayed as a CUDA graph
            self.rounds += rounds
        else:
            for _ in range(rounds):
                if hasattr(self.e, "round_and_exchange"):
                    self.e.round_and_exchange(f_limit, dist)
                else:
                    outboxes = self.e.round(f_limit)
                    self.e.insert(self.exchange(outboxes))
                self.rounds += 1
        if hasattr(self.e, "round_and_exchange"):
            self.bytes_sent = self.e.bytes_sent
        mn, best, cnt = self.e.status()
        # one small collective for the stop test and the counters: min over ranks of {min open f, best goal g}
        # (PAStar.cpp:502-519) and the sums
        mine = t.tensor([mn, best, cnt["expansions"], cnt["generated"], cnt["pops"]], dtype=t.int64, device=self._dev())
        allv = t.empty(self.world * 5, dtype=t.int64, device=self._dev())
        dist.all_gather_into_tensor(allv, mine)
        allv = allv.view(self.world, 5)
        mins = allv[:, :2].min(dim=0).values.tolist()
        tot = allv[:, 2:].sum(dim=0).tolist()
        return int(mins[0]), int(mins[1]), [int(x) for x in tot]

    def run(self):
        """Search to the optimality-preserving stop (or the expansion budget).  Returns a dict on every rank."""
        best = INT_MAX
        # device-driven engines chain a few rounds between status exchanges: rounds past the optimum only pop nodes
        # with f >= g* on partitions that have not heard of the goal yet, which cannot change the result
        per = self.rounds_per_status if getattr(self.e, "async_rounds", False) else 1
        while True:
            mn, best, tot = self.step(best, per)
            if mn >= best or mn == INT_MAX:  # every open node everywhere has f >= g_goal; nothing in flight
                finished = best != INT_MAX
                break
            if self.max_expansions and tot[0] >= self.max_expansions:
                finished = False
                break
        res = {"finished": int(finished), "g": best if finished else -1, "expansions": tot[0], "generated": tot[1],
               "pops": tot[2], "rounds": self.rounds}
        if finished:
            res["rows"] = self.backtrace()
        return res

    def backtrace(self):
        """Walk parenti from the final coordinate; the owner of each coordinate answers (one allreduce per column)."""
        t, dist = self.torch, self.dist
        n = len(self.seqs)
        pos = [len(s) for s in self.seqs]
        cols = []
        while any(pos):
            hit = self.e.lookup(np.array(pos, dtype=np.uint16)) if self.owner_fn(pos) == self.rank else None
            v = t.tensor([1, hit[0], hit[1]] if hit else [0, -1, -1], dtype=t.int64, device=self._dev())
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
            if int(v[0]) != 1:
                raise RuntimeError("backtrace lost the parent chain at %s" % (pos,))
            mask = int(v[2])
            if mask <= 0:  # a zero move mask would never reach the origin
                raise RuntimeError("backtrace found an entry without a parent move at %s" % (pos,))
            cols.append(mask)
            pos = [p - ((mask >> i) & 1) for i, p in enumerate(pos)]
        rows = [[] for _ in range(n)]
        at = [0] * n
        for mask in reversed(cols):
            for i in range(n):
                if (mask >> i) & 1:
                    rows[i].append(self.seqs[i][at[i]])
                    at[i] += 1
                else:
                    rows[i].append("-")
        return ["".join(r) for r in rows]
