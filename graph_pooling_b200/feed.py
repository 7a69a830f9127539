"""Host -> device feed of the dense fp32 adjacency the reference hands the model (train.py:197-201).

End to end, a step of the large configuration is bound by PCIe, not by the GPU: 4.3 GB of {0,1} floats per
256 x 2048^2 batch take ~66 ms to cross PCIe against ~18 ms of GPU work.  ``HostAdjacencyFeed`` keeps the caller's
contract -- a dense float32 adjacency in (pinned) host memory, a different one every step -- and moves it smarter:

  * the host cores bit-pack a fraction of the batch's graphs (``gp_host_pack_adj_bits``: 32x smaller, exact for
    {0,1}; ~100 GB/s on the B200 box's 16 threads, i.e. faster than PCIe moves the raw floats),
  * the remaining graphs cross PCIe as fp32 AT THE SAME TIME (a copy stream),
  * ``gp_adj_prepare_x`` expands both parts into the bf16 operand on the device (``PreparedAdjacency``), which the
    tensor-core encoders accept in place of ``adj``.

The packed fraction is chosen from the two rates measured on the machine (``calibrate``).  Three stages overlap in
steady state: host pack of batch i+2, H2D of batch i+1, GPU compute of batch i (double-buffered host and device
staging).  An adjacency with entries outside {0,1} is reported by the packer; ``submit`` then raises, and the
caller falls back to the plain fp32 copy.
"""
import ctypes as C
import os
import time
from concurrent.futures import ThreadPoolExecutor

import torch

from . import engine_tc as T
from ._lib import load


class HostAdjacencyFeed:
    def __init__(self, B, N, device, packed_graphs=None, threads=None):
        self.B, self.N, self.dev = int(B), int(N), torch.device(device)
        self.ldb = (self.N + 7) // 8
        self.threads = int(threads or os.cpu_count() or 1)
        self.Bp = self.B // 2 if packed_graphs is None else int(packed_graphs)
        self._lib = load()
        self._alloc()

    def _alloc(self):
        B, N, Bp, dev = self.B, self.N, self.Bp, self.dev
        self.hbits = [torch.empty(max(Bp, 1), N, self.ldb, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.dbits = [torch.empty(max(Bp, 1), N, self.ldb, device=dev, dtype=torch.uint8) for _ in range(2)]
        self.dtail = [torch.empty(max(B - Bp, 1), N, N, device=dev) for _ in range(2)]
        self.pa = [T.PreparedAdjacency(B, N, dev) for _ in range(2)]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        # freed[slot]: the last CONSUMER of the slot's device staging buffers (gp_adj_prepare_x in prepared(), and the
        # step that read the caller's extra destinations -- consumed()) has finished; copy() waits for it before
        # overwriting them (write-after-read across the copy and compute streams)
        self.freed = [torch.cuda.Event(), torch.cuda.Event()]
        for e in self.ready + self.freed:
            e.record(torch.cuda.current_stream(dev))
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.pool = ThreadPoolExecutor(1)
        self.fut = [None, None]
        self._src = [None, None]

    def close(self):
        self.pool.shutdown(wait=True)

    # ---- the three stages ---------------------------------------------------------------------------------
    def _pack(self, adj_host, dst, cnt):
        bad = C.c_int(0)
        rc = self._lib.gp_host_pack_adj_bits(adj_host.data_ptr(), C.c_longlong(cnt * self.N), self.N, dst.data_ptr(),
                                             C.c_longlong(self.ldb), self.threads, C.addressof(bad))
        if rc != 0 or bad.value:
            raise ValueError('adjacency has entries outside {0,1}: the bit-packed feed does not apply')

    def submit(self, adj_host, slot):
        """Stage 1 (host threads, asynchronous): start packing graphs [0, Bp) of `adj_host` ([B,N,N] float32, pinned)
        into host staging buffer `slot`.  The buffer's previous H2D copy must have completed (it has: we wait)."""
        if adj_host.dtype != torch.float32 or adj_host.is_cuda or tuple(adj_host.shape) != (self.B, self.N, self.N):
            raise ValueError('expected a host float32 adjacency of shape [B,N,N]')
        self.ready[slot].synchronize()
        self._src[slot] = adj_host
        self.fut[slot] = self.pool.submit(self._pack, adj_host, self.hbits[slot], self.Bp) if self.Bp else None

    def copy(self, slot, extra=()):
        """Stage 2 (copy stream, asynchronous): H2D of the packed bits and of the fp32 tail of the batch submitted to
        `slot`; `extra` = [(dst_device_tensor, src_pinned_tensor)] rides along (features, labels).  The copies wait
        for the slot's previous consumer: prepared() marks the staging buffers as consumed by itself; a caller whose
        step reads `extra` destinations without a host synchronisation must call consumed(slot) after that step."""
        if self.fut[slot] is not None:
            self.fut[slot].result()
        adj_host = self._src[slot]
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.freed[slot])
            for dst, src in extra:
                dst.copy_(src, non_blocking=True)
            if self.Bp:
                self.dbits[slot].copy_(self.hbits[slot], non_blocking=True)
            if self.Bp < self.B:
                self.dtail[slot][:self.B - self.Bp].copy_(adj_host[self.Bp:], non_blocking=True)
            self.ready[slot].record(self.copy_stream)

    def prepared(self, slot, nb_dev=None):
        """Stage 3 (current stream): wait for the copy, expand both parts into the bf16 operand; returns the
        PreparedAdjacency to pass to model(x, adj, ...) / model.loss(..., adj, ...)."""
        torch.cuda.current_stream(self.dev).wait_event(self.ready[slot])
        pa = self.pa[slot]
        pa.reset()
        if self.Bp:
            pa.add(self.dbits[slot], 0, nb_dev, 'bits')
        if self.Bp < self.B:
            pa.add(self.dtail[slot][:self.B - self.Bp], self.Bp, nb_dev, 'f32')
        self.freed[slot].record(torch.cuda.current_stream(self.dev))
        return pa

    def consumed(self, slot):
        """Call on the compute stream after the step that read the slot's `extra` destinations (and the prepared
        operand): the next copy() into this slot is ordered behind it."""
        self.freed[slot].record(torch.cuda.current_stream(self.dev))

    # ---- choosing the packed fraction ---------------------------------------------------------------------------
    @staticmethod
    def measure_rates(adj_host, device, threads=None):
        """(seconds to bit-pack the whole batch on the host, seconds to copy it as fp32): the two competing paths."""
        B, N, _ = adj_host.shape
        f = HostAdjacencyFeed(B, N, device, packed_graphs=B, threads=threads)
        t0 = time.perf_counter()
        f._pack(adj_host, f.hbits[0], B)
        t_pack = time.perf_counter() - t0
        tmp = torch.empty(B, N, N, device=device)
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        tmp.copy_(adj_host, non_blocking=True)
        torch.cuda.synchronize(device)
        t_h2d = time.perf_counter() - t0
        f.close()
        return t_pack, t_h2d


class EdgeListFeed:
    """The plugin's own feed: per step the host hands over the batch's EDGE LISTS (pinned int32
    [E,2] graph-local ids + eptr [B+1]) and the features; 8 (int32 ids) or 4 (int16 ids) bytes per edge cross PCIe instead of 4 N^2 bytes per graph,
    and gp_adj_from_edges writes the bf16 operand on the device.  Double-buffered like HostAdjacencyFeed: H2D of
    step i+1 on a copy stream while step i computes; copy() waits for the slot's previous consumer."""

    def __init__(self, B, N, device, max_edges, edge_dtype=torch.int32):
        self.B, self.N, self.dev = int(B), int(N), torch.device(device)
        self.max_edges = int(max_edges)
        if edge_dtype not in (torch.int32, torch.int16) or (edge_dtype == torch.int16 and self.N > 32768):
            raise ValueError('edge_dtype: torch.int32, or torch.int16 for N <= 32768')
        self.edge_dtype = edge_dtype
        self.dedges = [torch.empty(self.max_edges, 2, device=self.dev, dtype=edge_dtype) for _ in range(2)]
        self.deptr = [torch.empty(self.B + 1, device=self.dev, dtype=torch.int32) for _ in range(2)]
        self.pa = [T.PreparedAdjacency(B, N, self.dev) for _ in range(2)]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.freed = [torch.cuda.Event(), torch.cuda.Event()]
        for e in self.ready + self.freed:
            e.record(torch.cuda.current_stream(self.dev))
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self._n = [0, 0]
        self._maxdeg = [0, 0]

    def copy(self, slot, edges_host, eptr_host, max_edges_per_graph, extra=()):
        """H2D (copy stream) of one batch's edge lists; `extra` = [(dst_device_tensor, src_pinned_tensor)]."""
        E_ = int(edges_host.shape[0])
        if edges_host.dtype != self.edge_dtype or eptr_host.dtype != torch.int32 or E_ > self.max_edges:
            raise ValueError('EdgeListFeed.copy: edge lists of the feed\'s edge_dtype, at most max_edges edges')
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.freed[slot])
            for dst, src in extra:
                dst.copy_(src, non_blocking=True)
            self.dedges[slot][:E_].copy_(edges_host, non_blocking=True)
            self.deptr[slot].copy_(eptr_host, non_blocking=True)
            self.ready[slot].record(self.copy_stream)
        self._n[slot], self._maxdeg[slot] = E_, int(max_edges_per_graph)

    def prepared(self, slot, undirected=True):
        """Current stream: wait for the copy, build the bf16 operand; returns the PreparedAdjacency for model(...)."""
        torch.cuda.current_stream(self.dev).wait_event(self.ready[slot])
        pa = self.pa[slot]
        pa.reset()
        pa.from_edges(self.dedges[slot][:max(self._n[slot], 1)], self.deptr[slot], self._maxdeg[slot], undirected)
        return pa

    def consumed(self, slot):
        self.freed[slot].record(torch.cuda.current_stream(self.dev))
