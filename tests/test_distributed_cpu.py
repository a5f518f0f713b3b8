"""CPU, world_size 2 and 3, gloo: the host-side logic of the row-sharded multi-GPU path
(bella_b200/distributed.py) -- panel packing, the all-gather, reconstruction of B, the balanced
column ranges -- and that the union of the per-rank column ranges reproduces the whole result
(each rank's range is evaluated with the CPU oracle here; on the box the same ranges go to the GPUs)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmpdir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from bella_b200 import distributed as bd, frontend as fe
    import oracle_lib as ol
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    inp = fe.synthetic(900, 3000, seed=21)
    cuts = bd.shard_bounds(inp.B_colptr, world)
    r0, r1 = cuts[rank], cuts[rank + 1]
    host = bd.pack_panel(inp, r0, r1)
    shapes = bd.exchange_sizes(r1 - r0, int(inp.B_colptr[r1]) - int(inp.B_colptr[r0]), torch.device("cpu"))
    assert sum(a for a, _ in shapes) == inp.n_reads and sum(b for _, b in shapes) == inp.nnz
    max_bytes = max(bd.panel_layout(a, b)[1] for a, b in shapes)
    panel = torch.zeros(max_bytes, dtype=torch.uint8)
    panel[:host.size] = torch.from_numpy(host)
    B = bd.unpack_panels(bd.all_gather_panels(panel, max_bytes), shapes)
    # every rank reconstructs the whole B
    strand = np.unpackbits(inp.B_strand, bitorder="little")[:inp.nnz].astype(np.uint32)
    np.testing.assert_array_equal(B["colptr"].numpy().view(np.uint32), inp.B_colptr)
    np.testing.assert_array_equal(B["rowids"].numpy().view(np.uint32), inp.B_rowids | (strand << 31))
    np.testing.assert_array_equal(B["values"].numpy().view(np.uint16), inp.B_values)
    np.testing.assert_array_equal(B["read_len"].numpy().view(np.uint32), inp.read_len)
    bounds = bd.column_ranges(B["colptr64"], world)
    assert bounds[0] == 0 and bounds[-1] == inp.n_reads and all(a <= b for a, b in zip(bounds[:-1], bounds[1:]))
    # this rank's columns (oracle stands in for the GPU here), checked against the whole result on rank 0
    want = ol.oracle_spgemm(inp)
    lo, hi = bounds[rank], bounds[rank + 1]
    z0, z1 = int(want.colptrC[lo]), int(want.colptrC[hi])
    mine = torch.tensor([z1 - z0, int(want.flopC[lo:hi].astype(np.int64).sum())], dtype=torch.int64)
    allm = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allm, mine)
    assert sum(int(t[0]) for t in allm) == want.nnz
    # modelled per-rank time (transposition of every read at or above the range + this range's products) is even
    flops = [int(t[1]) for t in allm]
    scale = sum(flops) / max(1.0, float(sum(np.diff(inp.B_colptr.astype(np.int64)) * np.arange(inp.n_reads - 1, -1, -1)) / (inp.n_reads - 1)))
    cost = [bd.RHO * scale * (inp.nnz - int(inp.B_colptr[bounds[r]])) + flops[r] for r in range(world)]
    assert max(cost) <= 1.3 * (sum(cost) / world), f"unbalanced ranges: {cost}"
    np.save(os.path.join(tmpdir, f"rows_{rank}.npy"), want.rowids[z0:z1])
    dist.barrier()
    if rank == 0:
        rows = np.concatenate([np.load(os.path.join(tmpdir, f"rows_{r}.npy")) for r in range(world)])
        np.testing.assert_array_equal(rows, want.rowids)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_exchange_and_ranges(world, tmp_path):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)


def test_column_ranges_degenerate():
    from bella_b200 import distributed as bd
    assert bd.column_ranges(torch.zeros(1, dtype=torch.int64), 4) == [0, 0, 0, 0, 0]
    b = bd.column_ranges(torch.arange(0, 11, dtype=torch.int64) * 0, 2)      # all-empty reads
    assert b[0] == 0 and b[-1] == 10


def test_exchange_plan_is_consistent_across_ranks():
    # every rank derives the same owner ranges from the all-gathered counts, what rank s sends to rank d is what d
    # expects from s, and the (source, column) segment offsets tile the receive buffer exactly
    from bella_b200 import distributed as bd
    torch.manual_seed(3)
    world, n = 4, 3000
    counts = torch.randint(0, 60, (world, n), dtype=torch.int32)
    counts[:, 100:400] = 0                                   # a stretch of empty columns
    plans = [bd.exchange_plan(counts, r) for r in range(world)]
    bounds = plans[0][0]
    assert bounds[0] == 0 and bounds[-1] == n and all(a <= b for a, b in zip(bounds[:-1], bounds[1:]))
    tot = counts.to(torch.int64).sum(0)
    cost = [int(tot[bounds[r]:bounds[r + 1]].sum()) + bd.UNIT_OVERHEAD * int((tot[bounds[r]:bounds[r + 1]] > 0).sum()) for r in range(world)]
    assert max(cost) <= 1.05 * (sum(cost) / world) + 2 * (60 * world + bd.UNIT_OVERHEAD)
    for r in range(world):
        b, ins, outs, segoff, recvbase, sendoff = plans[r]
        assert b == bounds
        lo, hi = bounds[r], bounds[r + 1]
        assert segoff.shape == (world, hi - lo + 1)
        for s in range(world):
            assert ins[s] == plans[s][2][r]                              # s -> r as seen from both sides
            assert int(segoff[s, -1]) == ins[s]
            np.testing.assert_array_equal(np.diff(segoff[s].numpy()), counts[s, lo:hi].numpy())
        assert recvbase.tolist() == [sum(ins[:s]) for s in range(world)]
        assert int(sendoff[-1]) == int(counts[r].sum()) and sendoff.numel() == n + 1


# ---- row f1: the alignment step shards by pairs, no collective (bella_b200/distributed_xdrop.py) -----------------------

class _EmulatedAligner:
    """stands in for bella_b200.xdrop.XdropAligner on a CPU box: the same device source under the lane emulator"""

    def set_reads(self, seqs, seq_off):
        from bella_b200 import frontend as fe
        self.inp = fe.OverlapInputs(n_reads=len(seq_off) - 1, n_kmers=0, nnz=0, A_colptr=None, A_rowids=None, A_values=None, A_strand=None,
                                    B_colptr=None, B_rowids=None, B_values=None, B_strand=None, read_len=None, kmer_size=17, seqs=seqs, seq_off=seq_off)

    def set_params(self, kmer_len, xdrop, ratiophi, delta, fixed):
        self.inp.kmer_size = kmer_len
        self.params = (xdrop, ratiophi, delta, fixed)

    def align(self, rows, cols, posH, posV):
        import test_xdrop_emu as emu
        x, phi, delta, fixed = self.params
        rc, out, _ = emu.emu_align(self.inp, rows, cols, posH, posV, x, 1, 64, phi, delta, fixed, warps=2)
        assert rc == 0
        return out


def _xdrop_worker(rank, world, port, tmpdir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from bella_b200 import distributed_xdrop as dx, frontend as fe
    import oracle_lib as ol
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    inp = fe.synthetic(100, 1000, coverage=12.0, seed=17)
    c = ol.oracle_spgemm(inp, want_aux=False)
    cols = np.repeat(np.arange(inp.n_reads, dtype=np.uint32), np.diff(c.colptrC.astype(np.int64)))
    n = min(c.nnz, 240)
    pairs = (c.rowids[:n], cols[:n], c.posH[:n], c.posV[:n])
    a = dx.ShardedXdropAligner(rank, world, aligner_factory=_EmulatedAligner)
    a.set_reads(inp.seqs, inp.seq_off)
    a.set_params(inp.kmer_size, 7, 0.5, 0.1, -1)
    lo, hi, out = a.align_my_share(*pairs)
    # the slices tile the batch and the union is the oracle's answer (gathered on every rank)
    parts = [None] * world
    dist.all_gather_object(parts, (lo, hi, out))
    parts.sort(key=lambda t: t[0])
    assert parts[0][0] == 0 and parts[-1][1] == n and all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
    got = np.concatenate([p[2] for p in parts])
    np.testing.assert_array_equal(got, ol.oracle_align_post(inp, *pairs, 7, 0.5, 0.1, -1))
    cost = dx.pair_costs(inp.seq_off, *pairs, inp.kmer_size)
    share = [cost[p[0]:p[1]].sum() for p in parts]
    assert max(share) <= 1.35 * (sum(share) / world)            # balanced on expected work, not on pair count
    dist.destroy_process_group()
    open(os.path.join(tmpdir, f"xok{rank}"), "w").close()


@pytest.mark.parametrize("world", [2, 3])
def test_alignment_shards_by_pairs_without_a_collective(world, tmp_path):
    port = _free_port()
    mp.spawn(_xdrop_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"xok{r}") for r in range(world))
