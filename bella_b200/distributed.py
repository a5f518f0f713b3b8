"""Row-sharded multi-GPU overlap SpGEMM: one process per GPU, `torch.distributed` for the plumbing.

Layout (the reference has no distributed code, so there is no call site to mirror):
  * A is row-sharded: rank r owns the reads [read_lo_r, read_hi_r) -- equivalently the columns of
    B = A^T of those reads -- as a *compressed column panel*: per nonzero the k-mer id with the strand
    bit in bit 31 (u32) and the position (u16), plus per read its k-mer count and length.
  * per batch ONE all-gather of the panels (NCCL over NVLink) gives every rank the whole B.
  * the transpose of B is split by k-mer range: rank r builds A's columns for its k-mers only and
    expands the kept products of those k-mers -- for every output column -- into a send buffer
    ordered by output column (bella_b200_mg_transpose / _mg_scatter);
  * output columns are independent (overlap.hpp:286-287) and owned in contiguous ranges balanced on
    the exact product counts (an all-gather of the per-column counts, 4 bytes per read and rank); one
    all-to-all moves every product (8 bytes) to the owner of its column;
  * the owner regroups what it received and runs the group + fold kernels on its columns
    (bella_b200_mg_finish).  Outputs are disjoint column ranges: no reduction, the caller concatenates.
`mode="nvlink"` (the default of bench.py) takes NCCL off the data path altogether: a rank owns the output columns of its own
reads, every rank maps the others' exchange buffers (torch symmetric memory = CUDA peer memory over NVLink / NVSwitch), level
1 of the transpose stores each nonzero straight into the coarse k-mer bucket on the rank that transposes that k-mer range,
the per-column counts and the products' blocks are pushed the same way, the exchange is planned by one scan and one kernel
on the device, and the only thing between the phases is a device-side barrier (three per step) -- see _step_nvlink.
(`mode="replicate"` keeps the first version: after the all-gather every rank transposes all reads at or
above its first column itself and no products are exchanged.)

Everything in this file is host-side plumbing on torch tensors (CPU tensors with gloo in the unit
tests, CUDA tensors with NCCL on the box); the arithmetic is in libbella_b200.so.
"""
import sys

import numpy as np
import torch
import torch.distributed as dist

HEADER_WORDS = 4          # int64: n_reads, nnz, read_lo, reserved
ALIGN = 16


def _up(x, a=ALIGN):
    return (x + a - 1) // a * a


def panel_layout(n_r, nnz_r):
    """Byte offsets of the sections of one packed panel. -> dict, total bytes"""
    off = {}
    o = HEADER_WORDS * 8
    off["rowids"] = o; o = _up(o + 4 * nnz_r)
    off["values"] = o; o = _up(o + 2 * nnz_r)
    off["counts"] = o; o = _up(o + 4 * n_r)
    off["read_len"] = o; o = _up(o + 4 * n_r)
    return off, o


def shard_bounds(B_colptr, world):
    """Read ranges owned by the ranks before the exchange: contiguous, about equal nnz."""
    n = len(B_colptr) - 1
    nnz = int(B_colptr[-1])
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(B_colptr, nnz * r // world, side="left")))
    cuts.append(n)
    return [min(max(c, 0), n) for c in cuts]


def pack_panel(inp, r0, r1):
    """Compressed column panel of reads [r0, r1) of `inp` (frontend.OverlapInputs) as one uint8 array."""
    if inp.n_kmers > 0x80000000:
        raise ValueError("panel format keeps the strand bit in bit 31 of the k-mer id: needs < 2^31 k-mers")
    j0, j1 = int(inp.B_colptr[r0]), int(inp.B_colptr[r1])
    n_r, nnz_r = r1 - r0, j1 - j0
    off, total = panel_layout(n_r, nnz_r)
    buf = np.zeros(total, dtype=np.uint8)
    buf[:HEADER_WORDS * 8].view(np.int64)[:] = [n_r, nnz_r, r0, 0]
    strand = np.unpackbits(inp.B_strand, bitorder="little")[j0:j1].astype(np.uint32)
    buf[off["rowids"]:off["rowids"] + 4 * nnz_r].view(np.uint32)[:] = inp.B_rowids[j0:j1] | (strand << 31)
    buf[off["values"]:off["values"] + 2 * nnz_r].view(np.uint16)[:] = inp.B_values[j0:j1]
    buf[off["counts"]:off["counts"] + 4 * n_r].view(np.uint32)[:] = np.diff(inp.B_colptr[r0:r1 + 1])
    buf[off["read_len"]:off["read_len"] + 4 * n_r].view(np.uint32)[:] = inp.read_len[r0:r1]
    return buf


def exchange_sizes(n_r, nnz_r, device):
    """Tiny all-gather of the panel shapes (once per sharding, not per batch). -> [(n_r, nnz_r)] per rank"""
    world = dist.get_world_size()
    mine = torch.tensor([n_r, nnz_r], dtype=torch.int64, device=device)
    allm = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allm, mine)
    return [(int(t[0]), int(t[1])) for t in allm]


def all_gather_panels(panel, max_bytes):
    """THE collective of the batch: every rank contributes its panel (padded to max_bytes). -> (world, max_bytes) uint8"""
    world = dist.get_world_size()
    assert panel.dtype == torch.uint8 and panel.numel() == max_bytes
    out = torch.empty((world, max_bytes), dtype=torch.uint8, device=panel.device)
    if dist.get_backend() == "nccl":
        dist.all_gather_into_tensor(out.view(-1), panel)
    else:
        dist.all_gather(list(out.unbind(0)), panel)
    return out


def unpack_panels(gathered, shapes):
    """(world, max_bytes) uint8 + per-rank (n_r, nnz_r) -> B of all reads as tensors on the same device:
    colptr int32 [n+1] (uint32 bit pattern), rowids int32 [nnz] (strand in bit 31), values int16 [nnz], read_len int32 [n]"""
    rows, vals, cnts, lens = [], [], [], []
    for r, (n_r, nnz_r) in enumerate(shapes):
        off, _ = panel_layout(n_r, nnz_r)
        b = gathered[r]
        rows.append(b[off["rowids"]:off["rowids"] + 4 * nnz_r].view(torch.int32))
        vals.append(b[off["values"]:off["values"] + 2 * nnz_r].view(torch.int16))
        cnts.append(b[off["counts"]:off["counts"] + 4 * n_r].view(torch.int32))
        lens.append(b[off["read_len"]:off["read_len"] + 4 * n_r].view(torch.int32))
    counts = torch.cat(cnts).to(torch.int64)
    colptr64 = torch.zeros(counts.numel() + 1, dtype=torch.int64, device=gathered.device)
    torch.cumsum(counts, 0, out=colptr64[1:])
    return {"colptr64": colptr64, "colptr": colptr64.to(torch.int32), "rowids": torch.cat(rows), "values": torch.cat(vals),
            "read_len": torch.cat(lens)}


RHO = 0.28      # cost of transposing one nonzero relative to one estimated product (measured on B200)


def column_ranges(colptr64, world, rho=RHO):
    """Contiguous output-column ranges [b_r, b_{r+1}) that equalise the modelled per-rank time
         rho * nnz(rows >= b_r)  +  sum_{i in range} len_i * (n-1-i)/(n-1)
    -- a rank transposes every read at or above its first column (rows below never pair with its
    columns) and expands the kept products of its own columns, whose number is estimated by the read
    length times the share of reads with a larger id (the strictly lower triangle makes low column ids
    heavier; the reference balances its stages on the nnz prefix, overlap.hpp:703-710).
    Deterministic in its inputs, so every rank computes the same bounds. -> list of world+1 ints"""
    cp = colptr64.detach().cpu().numpy().astype(np.float64)
    n = cp.size - 1
    if n == 0:
        return [0] * (world + 1)
    nnz = cp[-1]
    if nnz <= 0 or world == 1:
        return [n * r // world for r in range(world + 1)]
    lens = np.diff(cp)
    w = lens * (np.arange(n - 1, -1, -1, dtype=np.float64) / max(n - 1, 1))
    W = np.concatenate([[0.0], np.cumsum(w)])
    tail = nnz - cp                          # nnz of rows >= i, i = 0..n
    g = rho * tail - W                       # decreasing in i

    def assign(T):
        bounds = [n]
        hi = n
        for _ in range(world - 1):
            # smallest lo with rho*tail[lo] + W[hi] - W[lo] <= T
            lo = int(np.searchsorted(-g[:hi + 1], -(T - W[hi]), side="left"))
            lo = min(lo, hi)
            bounds.append(lo)
            hi = lo
        return bounds[::-1], rho * nnz + W[hi]      # bounds[1:], cost of rank 0 = columns [0, hi)

    lo_T, hi_T = 0.0, rho * nnz + W[-1]
    for _ in range(60):
        T = 0.5 * (lo_T + hi_T)
        _, c0 = assign(T)
        if c0 <= T:
            hi_T = T
        else:
            lo_T = T
    inner, _ = assign(hi_T)
    bounds = [0] + [int(x) for x in inner]
    for k in range(1, len(bounds)):
        bounds[k] = min(max(bounds[k], bounds[k - 1]), n)
    bounds[-1] = n
    return bounds


def kmer_ranges(n_kmers, world):
    """k-mer ranges transposed by the ranks (k-mer ids are hash-order ids: uniform)."""
    return [int(n_kmers * r // world) for r in range(world + 1)]


def _owner_bounds(pre, world):
    """pre = inclusive prefix of the per-column totals (float64 [n]) -> int64 [world+1] bounds, on pre's device"""
    n = pre.numel()
    targets = pre[-1] * torch.arange(1, world, dtype=torch.float64, device=pre.device) / world
    cuts = torch.searchsorted(pre, targets, right=False) + 1
    cuts = torch.cummax(torch.clamp(cuts, max=n), 0).values if world > 1 else cuts
    z = torch.zeros(1, dtype=torch.int64, device=pre.device)
    return torch.cat([z, cuts.to(torch.int64), z + n])


UNIT_OVERHEAD = 320     # products: what one more (non-empty) column costs the group + fold stage, measured on 4 x B200
                        # (two ranks with equal products, 7 k vs 24 k columns: 1.04 vs 1.34 ms)


def exchange_plan(counts_all, rank, fixed_bounds=None):
    """counts_all int32 [world][n]: every rank's per-column product counts.  One device->host copy.
    fixed_bounds: owner ranges given by the caller (route mode: a rank owns the columns of its own reads).
    -> (bounds list, in_splits, out_splits, segoff int64 [world][ncols+1], recvbase int64 [world], sendoff int64 [n+1])"""
    world, n = counts_all.shape
    dev = counts_all.device
    C = torch.zeros((world, n + 1), dtype=torch.int64, device=dev)
    torch.cumsum(counts_all, 1, out=C[:, 1:])
    if n == 0:
        z = torch.zeros((world, 1), dtype=torch.int64, device=dev)
        return [0] * (world + 1), [0] * world, [0] * world, z, torch.zeros(world, dtype=torch.int64, device=dev), C[rank]
    # owner ranges equalise  products + UNIT_OVERHEAD per non-empty column  (both prefix sums are monotone)
    if fixed_bounds is not None:
        b = torch.tensor(fixed_bounds, dtype=torch.int64, device=dev)
    else:
        tot_pre = C[:, 1:].sum(0)
        nonempty_pre = torch.cumsum((counts_all.sum(0) > 0).to(torch.int64), 0)
        b = _owner_bounds((tot_pre + UNIT_OVERHEAD * nonempty_pre).to(torch.float64), world)
    Cb = C[:, b]                                            # [world][world+1]
    host = torch.cat([b.view(1, -1), Cb]).cpu()             # the one synchronising copy
    bounds = [int(x) for x in host[0]]
    seg = host[1:, 1:] - host[1:, :-1]                      # seg[s][d] = products rank s sends to rank d
    in_splits = [int(x) for x in seg[:, rank]]
    out_splits = [int(x) for x in seg[rank, :]]
    lo, hi = bounds[rank], bounds[rank + 1]
    segoff = (C[:, lo:hi + 1] - C[:, lo:lo + 1]).contiguous()
    rb = [0]
    for x in in_splits[:-1]:
        rb.append(rb[-1] + x)
    recvbase = torch.tensor(rb, dtype=torch.int64, device=dev)
    return bounds, in_splits, out_splits, segoff, recvbase, C[rank]


class ShardedOverlapSpGEMM:
    """One instance per rank.  load_shard() once; step() per batch: all-gather + local SpGEMM of this rank's columns."""

    def __init__(self, device_index, mode="exchange"):
        from . import spgemm
        self.mode = mode
        self.dev = torch.device("cuda", device_index)
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.g = spgemm.OverlapSpGEMM(device_index)
        self.g.set_stream(torch.cuda.current_stream(self.dev).cuda_stream)
        self.panel = None
        self.keep = None
        import os
        self._profile = os.environ.get("BELLA_MG_PROFILE") == "1"
        self.prof = {}

    def load_shard(self, inp, pinned=False):
        """Take this rank's reads of `inp`, pack the panel and make it device resident."""
        if self.mode == "nvlink":
            # a rank owns the output columns of its own reads: shards balanced on the estimated products + the per-column
            # overhead of the group + fold stage (the same weights the exchange mode balances exactly, but up front)
            lens = np.diff(inp.B_colptr.astype(np.int64)).astype(np.float64)
            n = inp.n_reads
            w = 1.83 * lens * (np.arange(n - 1, -1, -1, dtype=np.float64) / max(n - 1, 1)) + UNIT_OVERHEAD * (lens > 0)
            if self.mode == "nvlink":
                # + what a rank pays per nonzero of its own reads (the route pass), in products: measured on 4 x B200, the route pass
                # costs 0.6 ms for 18 M nonzeros where group + fold costs 1.0 ms for 16 M products
                w = w + 0.53 * lens
            pre = np.cumsum(w)
            cuts = [0] + [int(np.searchsorted(pre, pre[-1] * r / self.world)) + 1 for r in range(1, self.world)] + [n]
            cuts = [min(max(c, 0), n) for c in np.maximum.accumulate(cuts)]
        else:
            cuts = shard_bounds(inp.B_colptr, self.world)
        self.cuts = cuts
        r0, r1 = cuts[self.rank], cuts[self.rank + 1]
        self.r0, self.r1 = r0, r1
        host = pack_panel(inp, r0, r1)
        n_r, nnz_r = r1 - r0, int(inp.B_colptr[r1]) - int(inp.B_colptr[r0])
        self.shapes = exchange_sizes(n_r, nnz_r, self.dev)
        self.max_bytes = max(panel_layout(a, b)[1] for a, b in self.shapes)
        self.n_kmers, self.kmer_size, self.bin_size = inp.n_kmers, inp.kmer_size, inp.bin_size
        self.host_panel = torch.zeros(self.max_bytes, dtype=torch.uint8)
        self.host_panel[:host.size] = torch.from_numpy(host)
        if pinned:
            self.host_panel = self.host_panel.pin_memory()
        self.panel = self.host_panel.to(self.dev)
        if self.mode == "nvlink":
            self._setup_nvlink(inp.n_reads, inp.nnz, n_r, nnz_r)
        if self.mode == "nvlink":
            off, _ = panel_layout(n_r, nnz_r)
            p = self.panel
            self.loc = {"rowids": p[off["rowids"]:off["rowids"] + 4 * nnz_r].view(torch.int32),
                        "values": p[off["values"]:off["values"] + 2 * nnz_r].view(torch.int16),
                        "counts": p[off["counts"]:off["counts"] + 4 * n_r].view(torch.int32),
                        "read_len": p[off["read_len"]:off["read_len"] + 4 * n_r].view(torch.int32)}
            self.nnz_local = nnz_r
        return r0, r1

    def _setup_nvlink(self, n, nnz_total, n_r, nnz_r):
        self._nnz_total = nnz_total
        """Exchange buffers of the NVLink mode, allocated ONCE per sharding as one block of symmetric memory (every rank maps
        every other rank's block); the layout is the same on all ranks."""
        import ctypes
        import torch.distributed._symmetric_memory as symm
        from . import spgemm
        world, dev = self.world, self.dev
        geom = np.zeros(8, dtype=np.uint32)
        rc = spgemm.lib().bella_b200_mg_geometry(self.n_kmers, nnz_total, world, geom.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"bella_b200_mg_geometry failed: {rc}")
        # a slot holds what ONE rank sends to ONE rank: sized for the rank with the most nonzeros (the shards are balanced on
        # products, not on nonzeros)
        max_nnz = max(b for _, b in self.shapes)
        geom[5] = (int(1.3 * max_nnz / world) + 2 * 4096 + 15) // 16 * 16
        self.geom = geom
        wshift, shift1, nb1, nb1_loc, kpr, cap1 = (int(x) for x in geom[:6])
        nsb = world                                              # one slot per source rank
        if not hasattr(self, "cap_recv"):
            # products a rank receives / sends (about flops / world; flops is about nnz for [l,u] = [2,8]); a step that finds them
            # too small says so on every rank (BELLA_B200_ERR_CAPACITY) and _step_nvlink grows them
            self.cap_recv = int(1.6 * nnz_total / world) + (1 << 20)
            self.cap_send = self.cap_recv
        lay, o = {}, 0
        for name, nbytes in (("E", nsb * cap1 * 8), ("K", nsb * cap1 * 4), ("CNT", (nsb + 2) * 8 * 4), ("CALL", (world * n + 48) * 4),
                             ("LEN0", (n + 16) * 4), ("LEN1", (n + 16) * 4), ("PROD", self.cap_recv * 8)):
            lay[name] = int(o)
            o = _up(int(o) + int(nbytes), 256)
        o = int(o)
        self.sym_layout, self.sym_bytes = lay, o
        self.sym = symm.empty(o, dtype=torch.uint8, device=dev)
        self.sym.zero_()
        self.sym_hdl = symm.rendezvous(self.sym, dist.group.WORLD)
        base = [int(p) for p in self.sym_hdl.buffer_ptrs]
        self.peer = {k: [b + off for b in base] for k, off in lay.items()}
        self.kr_lo = min(self.rank * kpr, self.n_kmers)
        self.kr_hi = min((self.rank + 1) * kpr, self.n_kmers)
        self.nnz_cap = int(1.15 * nnz_total / world) + 4096       # records this rank expects (k-mer ids are uniform)
        self.loc_i32 = lambda name, count: self.sym[lay[name]:lay[name] + 4 * count].view(torch.int32)
        # local work arrays (not symmetric)
        ncols = self.cuts[self.rank + 1] - self.cuts[self.rank]
        self.w_cnt = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
        self.w_scan = torch.zeros(world * n + 2, dtype=torch.int64, device=dev)
        self.w_sendoff = torch.zeros(n + 2, dtype=torch.int64, device=dev)
        self.w_segoff = torch.zeros(world * (ncols + 1) + 1, dtype=torch.int64, device=dev)
        self.w_recvbase = torch.zeros(world + 1, dtype=torch.int64, device=dev)
        self.w_push = torch.zeros(3 * world + 2, dtype=torch.int64, device=dev)
        self.w_send = torch.empty(self.cap_send + 2, dtype=torch.int64, device=dev)
        self.w_cuts = torch.tensor(self.cuts, dtype=torch.int32, device=dev)

    def _step_nvlink(self, fetch):
        """One batch without a collective on the data path (module docstring); every phase is one or two kernels of
        libbella_b200.so, the barriers are torch symmetric-memory barriers on the same stream.  The send / receive buffers
        are sized once; a batch that needs more is refused by every rank alike, the buffers grow and the batch runs again."""
        from .spgemm import BellaB200Error
        for _ in range(4):
            try:
                return self._step_nvlink_once(fetch)
            except BellaB200Error as e:
                if getattr(e, "code", 0) != -6:
                    raise
                need = [int(x) for x in self.w_push[3 * self.world:3 * self.world + 2].tolist()]
                if need[0] <= self.cap_recv and need[1] <= self.cap_send:
                    raise
                self.cap_recv = max(self.cap_recv, int(1.25 * need[0]) + 4096)
                self.cap_send = max(self.cap_send, int(1.25 * need[1]) + 4096)
                torch.cuda.synchronize(self.dev)
                self._setup_nvlink(self.cuts[-1], self._nnz_total, self.r1 - self.r0, self.nnz_local)
        raise RuntimeError("the exchange buffers did not converge")

    def _step_nvlink_once(self, fetch):
        if self._profile:
            import time
            torch.cuda.synchronize(self.dev)
            self._t0 = time.perf_counter()
        g, world, rank = self.g, self.world, self.rank
        r0, r1, n_r = self.r0, self.r1, self.r1 - self.r0
        n = self.cuts[-1]
        lay = self.sym_layout
        if not hasattr(self, "colptr_local"):
            # static per shard: the local colptr (the reads' k-mer counts travel with the panel)
            cp = torch.zeros(n_r + 1, dtype=torch.int64, device=self.dev)
            torch.cumsum(self.loc["counts"].to(torch.int64), 0, out=cp[1:])
            self.colptr_local = cp.to(torch.int32)
        colptr_global = self.colptr_local.data_ptr() - 4 * r0          # indexed by the GLOBAL read id
        # (two copies of the lengths, alternating: a fast rank posts the next batch's while a slow one still folds this batch's)
        self._flip = 1 - getattr(self, "_flip", 1)
        LEN = "LEN%d" % self._flip
        read_len = self.loc_i32(LEN, n)
        g.set_inputs_device(n, self.n_kmers, self.nnz_local, (colptr_global, self.loc["rowids"], self.loc["values"]), read_len, None,
                            self.kmer_size, self.bin_size)
        # read lengths of this rank's reads -> every rank (4 bytes per read), then level 1 of the transpose into the owners' buckets
        g.mg_post(n_r, self.loc["read_len"], world, r0, self.peer[LEN])
        g.mg_route_push(r0, r1, colptr_global, self.loc["rowids"], self.loc["values"], self.n_kmers, self.geom, world, rank,
                        self.peer["E"], self.peer["K"], self.peer["CNT"])
        self.sym_hdl.barrier(channel=0)
        self._tick("route_push + barrier")
        g.mg_transpose_coarse(self.kr_lo, self.kr_hi, self.geom, world, self.sym.data_ptr() + lay["E"], self.sym.data_ptr() + lay["K"],
                              self.sym.data_ptr() + lay["CNT"], self.nnz_cap, self.w_cnt)
        self._tick("transpose (level 2 + buckets)")
        g.mg_exchange(world, rank, self.w_cuts, self.w_cnt, self.peer["CALL"], 0)
        self.sym_hdl.barrier(channel=0)
        self._tick("counts + barrier")
        g.mg_exchange(world, rank, self.w_cuts, self.w_cnt, self.peer["CALL"], 1, counts_all=self.sym.data_ptr() + lay["CALL"], scan=self.w_scan,
                      cap_recv=self.cap_recv, cap_send=self.cap_send, sendoff=self.w_sendoff, segoff=self.w_segoff, recvbase=self.w_recvbase, push=self.w_push,
                      sendbuf=self.w_send, peer_recv=self.peer["PROD"])
        self.sym_hdl.barrier(channel=0)
        self._tick("plan + expand + push + barrier")
        g.mg_finish(r0, r1, world, self.sym.data_ptr() + lay["CALL"], self.w_segoff, self.w_recvbase, self.sym.data_ptr() + lay["PROD"])
        self._tick("mg_finish")
        if fetch:
            colptrC = g.get_colptr(pinned=True)
            res = g.numeric(pinned=True)
            return int(colptrC[r1 - r0]), g.result_flops(), (r0, r1), (colptrC[:r1 - r0 + 1],) + res
        g.numeric_device()
        return int(g.result_nnz()), g.result_flops(), (r0, r1)

    def upload(self):
        """e2e leg: host panel -> device inside the timed region."""
        self.panel.copy_(self.host_panel, non_blocking=True)

    def step(self, fetch=False):
        """-> (Z of this rank's columns, products, (col_lo, col_hi)[, host results when fetch=True])"""
        if self.mode == "nvlink":
            return self._step_nvlink(fetch)
        if self._profile:
            import time
            torch.cuda.synchronize(self.dev)
            self._t0 = time.perf_counter()
        # (an uneven all-gather straight into the final arrays was measured slower on 4 x B200 -- NCCL runs it as
        #  grouped broadcasts: 2.0-2.9 ms against 1.3 ms for the padded all-gather + one repacking copy)
        gathered = all_gather_panels(self.panel, self.max_bytes)
        B = unpack_panels(gathered, self.shapes)
        if self.mode == "exchange":
            return self._step_exchange(gathered, B, fetch)
        bounds = column_ranges(B["colptr64"], self.world)
        lo, hi = bounds[self.rank], bounds[self.rank + 1]
        n = B["read_len"].numel()
        nnz = B["rowids"].numel()
        self.keep = (gathered, B)
        self.g.set_inputs_device(n, self.n_kmers, nnz, (B["colptr"], B["rowids"], B["values"]), B["read_len"], None,
                                 self.kmer_size, self.bin_size)
        self.g.set_column_range(lo, hi)
        if fetch:       # the public host-facing calls: colptrC and the tuples of this rank's columns come back to the host
            flops, _, colptrC = self.g.symbolic(want_flopC=False, pinned=True)
            res = self.g.numeric(pinned=True)
            return int(colptrC[hi - lo]), flops, (lo, hi), (colptrC[:hi - lo + 1],) + res
        Z, flops = self.g.run_resident()
        return Z, flops, (lo, hi)

    def _tick(self, name):
        """BELLA_MG_PROFILE=1: per-stage wall times (device synchronised), kept in self.prof."""
        if not self._profile:
            return
        import time
        torch.cuda.synchronize(self.dev)
        now = time.perf_counter()
        self.prof[name] = self.prof.get(name, 0.0) + (now - self._t0) * 1e3
        self._t0 = now

    def _step_exchange(self, gathered, B, fetch):
        n, nnz = B["read_len"].numel(), B["rowids"].numel()
        g, dev = self.g, self.dev
        self._tick("allgather+unpack")
        g.set_inputs_device(n, self.n_kmers, nnz, (B["colptr"], B["rowids"], B["values"]), B["read_len"], None,
                            self.kmer_size, self.bin_size)
        kr = kmer_ranges(self.n_kmers, self.world)
        cnt_local = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        g.mg_transpose(kr[self.rank], kr[self.rank + 1], cnt_local)
        self._tick("mg_transpose")
        counts_all = torch.empty((self.world, n), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(counts_all.view(-1), cnt_local[:n].contiguous())
        bounds, in_splits, out_splits, segoff, recvbase, sendoff = exchange_plan(counts_all, self.rank)
        self._tick("counts+plan")
        send = torch.empty(max(sum(out_splits), 1), dtype=torch.int64, device=dev)
        g.mg_scatter(sendoff, send)
        self._tick("mg_scatter")
        recv = torch.empty(max(sum(in_splits), 1), dtype=torch.int64, device=dev)
        dist.all_to_all_single(recv[:sum(in_splits)], send[:sum(out_splits)], in_splits, out_splits)
        self._tick("all_to_all")
        lo, hi = bounds[self.rank], bounds[self.rank + 1]
        self.keep = (gathered, B, counts_all, segoff, recvbase, send, recv, cnt_local, sendoff)
        g.mg_finish(lo, hi, self.world, counts_all, segoff, recvbase, recv)
        self._tick("mg_finish")
        flops = sum(in_splits)
        if fetch:
            colptrC = g.get_colptr(pinned=True)
            res = g.numeric(pinned=True)
            return int(colptrC[hi - lo]), flops, (lo, hi), (colptrC[:hi - lo + 1],) + res
        g.numeric_device()
        return int(g.result_nnz()), flops, (lo, hi)

    def close(self):
        self.g.close()


def bench_multi(args, w, inp, rank, world, local, METRIC, UNIT, workload, ClockSampler, algorithmic_bytes, measured_peak):
    """bench.py body for N > 1 (launched under torchrun): strong scaling of the N=1 workload."""
    import json
    import time
    import os
    dev = torch.device("cuda", local)
    mode = os.environ.get("BELLA_MG_MODE", "nvlink")        # "exchange" / "replicate": the NCCL-based modes of round 1
    sh = ShardedOverlapSpGEMM(local, mode=mode)
    try:
        sh.load_shard(inp, pinned=True)
    except Exception as e:      # noqa: BLE001 -- e.g. peer memory cannot be mapped on this box: every rank fails alike
        if mode != "nvlink":
            raise
        print(f"[bench] rank {rank}: the NVLink mode could not be set up ({type(e).__name__}: {e}); falling back to the NCCL exchange mode",
              file=sys.stderr, flush=True)
        mode = "exchange"
        sh = ShardedOverlapSpGEMM(local, mode=mode)
        sh.load_shard(inp, pinned=True)
    stream = torch.cuda.current_stream(dev)

    def timed(fn, steps):
        torch.cuda.synchronize(dev)
        dist.barrier()
        torch.cuda.synchronize(dev)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ev0.record(stream)
        out = None
        for k in range(steps):
            out = fn()
            marks[k].record(stream)
        ev1.record(stream)
        torch.cuda.synchronize(dev)
        if rank == 0:       # per-step device times, for the log only
            per = [(ev0 if k == 0 else marks[k - 1]).elapsed_time(marks[k]) for k in range(steps)]
            print("[bench] per-step ms (rank 0): " + " ".join(f"{x:.2f}" for x in per), file=sys.stderr, flush=True)
        ms = torch.tensor([ev0.elapsed_time(ev1) / steps], dtype=torch.float64, device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0]), out

    for _ in range(max(args.warmup, 3) + 2):       # two more than asked: the first exchanges also set up NCCL's peer channels
        sh.step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches = [0]
    phases = np.zeros(4)

    def resident():
        r = sh.step()
        t = sh.g.timings()
        launches[0] += t["launches"]
        phases[:] += [t["transpose_ms"], t["scatter_ms"], t["group_fold_ms"], t["output_ms"]]
        return r

    ms, (Z, flops, rng) = timed(resident, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    tot = torch.tensor([Z, flops, launches[0]], dtype=torch.int64, device=dev)
    dist.all_reduce(tot)
    Zt, Ft, Lt = int(tot[0]), int(tot[1]), int(tot[2])

    # e2e: host panel -> device, exchange, SpGEMM, results of this rank's columns back to the host
    def e2e():
        sh.upload()
        return sh.step(fetch=True)[3]

    e2e_steps = max(2, min(args.steps, 5))
    e2e()
    e2e_ms, res = timed(e2e, e2e_steps)
    # what was computed: the checksum of every rank's tuples, summed, against the single-GPU result of the same workload
    # (rank 0 runs the single-GPU path once, outside the timed regions)
    from .checks import tuple_checksum
    cs, zc = tuple_checksum(rng[0], res[0], res[1:])
    allcs = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(allcs, torch.tensor([cs - (1 << 64) if cs >= (1 << 63) else cs, zc], dtype=torch.int64, device=dev))
    parity = None
    if rank == 0:
        from . import spgemm
        total = sum(int(t[0]) for t in allcs) % (1 << 64)
        zsum = sum(int(t[1]) for t in allcs)
        one = spgemm.overlap_spgemm(inp, device=local)
        ref_cs, ref_z = tuple_checksum(0, one["colptrC"], (one["rowids"], one["count"], one["posH"], one["posV"]))
        if total != ref_cs or zsum != ref_z:
            raise SystemExit(f"bench.py: PARITY FAILURE at {world} GPUs: checksum {total:016x} over {zsum} tuples, single GPU {ref_cs:016x} over {ref_z}")
        parity = f"checksum of the {zsum} tuples of all {world} ranks == single-GPU result of the same workload ({ref_cs:016x})"
        print("[bench] parity: " + parity, file=sys.stderr, flush=True)
    dist.barrier()
    d2h = torch.tensor([int(sum(a.nbytes for a in res))], dtype=torch.int64, device=dev)
    h2d = torch.tensor([int(sh.max_bytes)], dtype=torch.int64, device=dev)
    dist.all_reduce(d2h)
    dist.all_reduce(h2d)

    if rank == 0:
        peak, peak_src = measured_peak()
        alg = algorithmic_bytes(inp, Zt, Ft)
        ph = phases / args.steps
        line = {"metric": METRIC, "value": Zt / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "u16/u32", "data": "synthetic",
                "config": {"workload": workload},
                "workload_stats": {"n_kmers": inp.n_kmers, "nnz_A": inp.nnz, "products": Ft, "output_nnz": Zt, "l2": "inputs larger than L2, no flush",
                                   "parallelism": {"nvlink": f"row-sharded x{world}: a rank owns the columns of its reads; nonzeros stored into the k-mer owner's "
                                                             "buckets and product blocks pushed to the column owner over NVLink peer memory, three device barriers, no collective",
                                                   }.get(mode, f"row-sharded x{world}: all-gather of the B panel, transpose split by k-mer range, all-to-all of the products")},
                "parity": parity,
                "clocks": clocks,
                "e2e": {"value": Zt / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d[0]), "d2h_bytes_per_step": int(d2h[0]),
                        "ms_per_step": e2e_ms},
                "gpu_launches": Lt,
                "roofline": {"bound": "hbm", "kernel": "whole step (rank 0 phases below)", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak * world,
                             "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / (peak * world), "traffic": None, "peak_source": peak_src,
                             "rank0_phase_ms": {"transpose": float(ph[0]), "k_scatter": float(ph[1]), "k_group_fold": float(ph[2]),
                                                "output": float(ph[3])}},
                "cpu_baseline": None}
        print(json.dumps(line), flush=True)
    sh.close()
    dist.barrier()
    dist.destroy_process_group()
    return 0
