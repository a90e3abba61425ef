"""CPU tests of the multi-GPU host logic (sharding.py): the row partitioner, the claim that a panel's
SGT equals its slice of the whole graph's SGT (what makes sharded results bit-identical), and the
per-layer exchange on a world_size-2 gloo group."""
import os
import socket

import numpy as np
import pytest
import torch

import tcgnn_oracle as orc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("kind", ["uniform", "rmat"])
def test_partition_rows_properties(world, kind):
    from sharding import partition_rows
    n = 5003
    rp, ci = (orc.rmat_graph if kind == "rmat" else orc.random_graph)(n, 120000, seed=world)
    b = partition_rows(torch.from_numpy(rp), world)
    assert len(b) == world + 1 and b[0] == 0 and b[-1] == n
    assert all(x <= y for x, y in zip(b, b[1:]))
    assert all(x % 16 == 0 for x in b[:-1])
    nnz = [int(rp[b[g + 1]] - rp[b[g]]) for g in range(world)]
    assert sum(nnz) == len(ci)
    if kind == "uniform" and world > 1:
        assert max(nnz) <= 1.15 * len(ci) / world      # balanced by stored non-zeros


def test_partition_rows_more_ranks_than_windows():
    from sharding import partition_rows
    rp = torch.tensor([0, 1, 2, 3], dtype=torch.int32)
    b = partition_rows(rp, 4)
    assert b[0] == 0 and b[-1] == 3 and all(x <= y for x, y in zip(b, b[1:]))


@pytest.mark.parametrize("world", [2, 5])
def test_panel_sgt_is_slice_of_global_sgt(world):
    """Host SGT of each panel (TCGNN.preprocess_panel on CPU tensors) == slice of the global arrays."""
    from sharding import RowPanel
    n = 3001
    rp, ci = orc.rmat_graph(n, 70000, seed=31)
    bp, e2c, e2r, _ = orc.sgt(rp, ci, n)
    t_rp, t_ci = torch.from_numpy(rp), torch.from_numpy(ci)
    covered_rows = covered_edges = 0
    for rank in range(world):
        p = RowPanel(t_rp, t_ci, rank, world)
        r0, r1, e0, e1 = p.row_base, p.row_base + p.num_rows, p.edge_begin, p.edge_end
        assert r0 % 16 == 0
        w0 = r0 // 16
        assert np.array_equal(p.blockPartition.numpy(), bp[w0:w0 + (p.num_rows + 15) // 16])
        assert np.array_equal(p.edgeToColumn.numpy(), e2c[e0:e1])
        assert np.array_equal(p.edgeToRow.numpy(), e2r[e0:e1] - r0)
        assert np.array_equal(p.row_pointers.numpy(), rp[r0:r1 + 1] - rp[r0])
        assert np.array_equal(p.column_index.numpy(), ci[e0:e1])          # global column ids
        # slicing precomputed global SGT arrays gives the same panel
        q = RowPanel(t_rp, t_ci, rank, world, bounds=p.bounds,
                     sgt=(torch.from_numpy(bp), torch.from_numpy(e2c), torch.from_numpy(e2r)))
        for a, b_ in zip(p.graph, q.graph):
            assert torch.equal(a, b_)
        covered_rows += p.num_rows
        covered_edges += p.num_edges
    assert covered_rows == n and covered_edges == len(ci)


def _gloo_worker(rank, world, port, n, d, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    for p in (os.path.join(root, "tc-gnn_atc23_b200"), os.path.join(root, "oracle"), here):
        sys.path.insert(0, p)
    import torch.distributed as dist
    from sharding import RowPanel
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rp, ci = orc.rmat_graph(n, 30000, seed=77)
        x = torch.from_numpy(np.random.default_rng(5).standard_normal((n, d)).astype(np.float32))
        p = RowPanel(torch.from_numpy(rp), torch.from_numpy(ci), rank, world)
        x_local = x[p.row_base:p.row_base + p.num_rows].clone()
        x_all = p.all_gather(x_local)
        ok = torch.equal(x_all, x)
        # a second layer re-uses the buffer
        x_all2 = p.all_gather(x_local * 2)
        ok = ok and torch.equal(x_all2, x * 2) and x_all2.data_ptr() == x_all.data_ptr()
        # the sharded aggregation composed from panels equals the global oracle (compute stands in on the
        # checker here: the CUDA kernels are exercised by tests/test_gpu_sharding.py)
        y_local = orc.spmm(x_all2.numpy() / 2, p.row_pointers.numpy(), p.column_index.numpy()) if p.num_rows else None
        y_ref = orc.spmm(x.numpy(), rp, ci)[p.row_base:p.row_base + p.num_rows]
        ok = ok and (y_local is None or np.array_equal(y_local, y_ref))
        q.put((rank, bool(ok), p.bounds))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_all_gather_exchange():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, 1500, 24, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert res[0][2] == res[1][2]        # same boundaries on both ranks


def test_partition_on_tc_blocks_balances_tiles():
    """With the SGT tile counts the panels are balanced on TC blocks (+ a per-window cost), which is what the
    kernels' time follows; boundaries stay on window boundaries and cover every row once."""
    from sharding import partition_rows
    n = 40000
    rp, ci = orc.rmat_graph(n, 900000, seed=33)
    bp, _, _, _ = orc.sgt(rp, ci, n)
    for world in (2, 3, 8):
        b = partition_rows(torch.from_numpy(rp), world, block_partition=torch.from_numpy(bp))
        assert b[0] == 0 and b[-1] == n and all(x % 16 == 0 for x in b[1:-1]) and b == sorted(b)
        cost = [int((np.maximum(bp[b[i] // 16:(b[i + 1] + 15) // 16], 1) + 3).sum()) for i in range(world)]
        assert max(cost) <= 1.05 * (sum(cost) / world) + int(bp.max()) + 3
        nnz_b = partition_rows(torch.from_numpy(rp), world)
        tiles_nnz = [int(bp[nnz_b[i] // 16:(nnz_b[i + 1] + 15) // 16].sum()) for i in range(world)]
        tiles_tc = [int(bp[b[i] // 16:(b[i + 1] + 15) // 16].sum()) for i in range(world)]
        assert max(tiles_tc) <= max(tiles_nnz)


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_two_phase_exchange_chunks_tile_every_panel(world):
    """The balanced exchange scatters chunk j of panel g to rank j and lets rank j forward it: the chunks of a panel
    must tile it exactly, and the chunks held by one rank after phase A (one per panel) must be disjoint."""
    from sharding import RowPanel
    n = 10007
    rp, ci = orc.rmat_graph(n, 200000, seed=40 + world)
    t_rp, t_ci = torch.from_numpy(rp), torch.from_numpy(ci)
    bp, e2c, e2r, _ = orc.sgt(rp, ci, n)
    sgt = tuple(torch.from_numpy(a) for a in (bp, e2c, e2r))
    p = RowPanel(t_rp, t_ci, 0, world, sgt=sgt)          # TC-block-balanced bounds, sliced SGT
    covered = np.zeros(n, dtype=np.int32)
    for g in range(world):
        prev = p.bounds[g]
        for j in range(world):
            c0, c1 = p._chunk(g, j)
            assert c0 == prev and c1 >= c0
            covered[c0:c1] += 1
            prev = c1
        assert prev == p.bounds[g + 1]
    assert (covered == 1).all()
    for j in range(world):                               # what rank j holds after phase A
        held = [p._chunk(g, j) for g in range(world)]
        assert all(a[1] <= b[0] for a, b in zip(held, held[1:]))
        rows = sum(c1 - c0 for c0, c1 in held)
        assert abs(rows - n / world) <= world            # ~1/N of the matrix whatever the panel sizes


@pytest.mark.parametrize("world", [2, 3, 8])
def test_partition_with_send_cost_bounds_rows_and_compute(world):
    """With a per-row send cost the boundaries bound BOTH a panel's TC blocks and its rows (what it has to push to
    every peer): the largest panel is smaller than under pure compute balancing, nothing is lost or duplicated."""
    from sharding import partition_rows, window_send_cost
    n = 60000
    rp, ci = orc.rmat_graph(n, 1500000, seed=35)
    bp, _, _, _ = orc.sgt(rp, ci, n)
    t_rp, t_bp = torch.from_numpy(rp), torch.from_numpy(bp)
    b0 = partition_rows(t_rp, world, block_partition=t_bp)
    b1 = partition_rows(t_rp, world, block_partition=t_bp, send_cost_per_row=6.0)
    for b in (b0, b1):
        assert b[0] == 0 and b[-1] == n and b == sorted(b) and all(x % 16 == 0 for x in b[1:-1])
    cost = np.maximum(bp, 1) + 3
    sw = window_send_cost(t_rp, world, 6.0)
    assert len(sw) == len(bp) and sw.max() <= 6.0 * 16 + 1e-9

    def worst(b):
        comp = [float(cost[b[i] // 16:(b[i + 1] + 15) // 16].sum()) for i in range(world)]
        send = [float(sw[b[i] // 16:(b[i + 1] + 15) // 16].sum()) for i in range(world)]
        return max(max(comp), max(send)), comp, send

    w0, _, _ = worst(b0)
    w1, comp, send = worst(b1)
    assert w1 <= w0 + 1e-6                       # never worse than pure compute balancing on the combined bound
    # and within one window of the unreachable ideal where both sums split perfectly
    ideal = max(sum(comp) / world, sum(send) / world)
    assert w1 <= 1.5 * ideal + cost.max() + sw.max()


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("dense_fraction", [0.0, 2.0])
def test_source_subgraphs_sum_to_the_panel_product(world, dense_fraction):
    """The overlapped exchange computes Y_panel = sum_p A[panel, cols of p] . X[rows shipped by p]; the sub-graphs
    (whole source panel, ids rebased -- or packed referenced rows, ids remapped) must reproduce the panel product.
    Host logic only: the oracle's SpMM stands in for the kernels, the SGT is the product's host SGT."""
    from sharding import RowPanel
    n = 2503
    rp, ci = orc.rmat_graph(n, 40000, seed=61)
    x = np.random.default_rng(6).integers(-4, 5, size=(n, 8)).astype(np.float32)
    want = orc.spmm(x, rp, ci)
    t_rp, t_ci = torch.from_numpy(rp), torch.from_numpy(ci)
    for rank in range(world):
        p = RowPanel(t_rp, t_ci, rank, world)
        subs = p.build_source_subgraphs(dense_fraction=dense_fraction)
        y = np.zeros((p.num_rows, 8), dtype=np.float32)
        edges = 0
        for src, sb in enumerate(subs):
            b0, b1 = p.bounds[src], p.bounds[src + 1]
            s_rp, s_ci, s_bp, s_e2c, s_e2r = (t.numpy() for t in sb["graph"])
            edges += len(s_ci)
            if sb["dense"]:
                assert src == rank or dense_fraction == 0.0
                xs = x[b0:b1]
            else:
                ref = sb["ref_rows"].numpy()
                assert len(ref) == sb["n_src"] and np.all(np.diff(ref) > 0) and ref.min() >= 0 and ref.max() < b1 - b0
                xs = x[b0:b1][ref]                 # what gather_rows packs on the source rank
            assert sb["n_src"] == len(xs)
            if len(s_ci):
                assert s_ci.max() < len(xs)
                o_bp, o_e2c, o_e2r, _ = orc.sgt(s_rp, s_ci, p.num_rows)
                assert np.array_equal(s_bp, o_bp) and np.array_equal(s_e2c, o_e2c) and np.array_equal(s_e2r, o_e2r)
                y += orc.spmm(xs, s_rp, s_ci)[:p.num_rows] if len(xs) >= p.num_rows else _spmm_rect(xs, s_rp, s_ci)
        assert edges == p.num_edges
        assert np.array_equal(y, want[p.row_base:p.row_base + p.num_rows])


def _spmm_rect(xs, rp, ci):
    """Pattern SpMM of a rectangular CSR (more rows than source rows) in plain NumPy (checker)."""
    y = np.zeros((len(rp) - 1, xs.shape[1]), dtype=np.float32)
    rows = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
    key = np.unique(rows.astype(np.int64) * (ci.max() + 1 if len(ci) else 1) + ci)
    m = ci.max() + 1 if len(ci) else 1
    np.add.at(y, key // m, xs[key % m])
    return y


@pytest.mark.parametrize("world,sizes", [(4, None), (8, None), (8, [1, 1, 1, 1, 1, 1, 1]), (8, [7]), (5, [2, 2])])
def test_group_subgraphs_follow_the_arrival_order_and_sum_to_the_panel_product(world, sizes):
    """Consecutive arrivals are merged into one sub-graph whose columns index the concatenation of what those
    sources ship, in arrival order (the layout of the receive area): own product + group products == panel product."""
    from sharding import RowPanel
    n = 3203
    rp, ci = orc.rmat_graph(n, 60000, seed=71)
    x = np.random.default_rng(7).integers(-4, 5, size=(n, 4)).astype(np.float32)
    want = orc.spmm(x, rp, ci)
    t_rp, t_ci = torch.from_numpy(rp), torch.from_numpy(ci)
    os.environ["TCGNN_DENSE_FRACTION"] = "0.5"          # a mix of whole-panel and packed sources
    try:
        for rank in (0, world - 1, world // 2):
            p = RowPanel(t_rp, t_ci, rank, world)
            subs = p.build_source_subgraphs()
            shipped = {}
            for src, sb in enumerate(subs):
                b0, b1 = p.bounds[src], p.bounds[src + 1]
                shipped[src] = x[b0:b1] if sb["dense"] else x[b0:b1][sb["ref_rows"].numpy()]
            groups = p.build_group_subgraphs(sizes)
            order = p.arrival_order()
            assert [s for g in groups for s in g["sources"]] == order and order[0] == (rank - 1) % world
            y = _spmm_rect(shipped[rank], subs[rank]["graph"][0].numpy(), subs[rank]["graph"][1].numpy())
            for g in groups:
                xs = np.concatenate([shipped[s] for s in g["sources"]]) if g["sources"] else np.zeros((0, 4), np.float32)
                assert g["ncols"] == len(xs)
                g_rp, g_ci = g["graph"][0].numpy(), g["graph"][1].numpy()
                if len(g_ci):
                    o_bp, o_e2c, o_e2r, _ = orc.sgt(g_rp, g_ci, p.num_rows)
                    assert np.array_equal(g["graph"][2].numpy(), o_bp) and np.array_equal(g["graph"][3].numpy(), o_e2c)
                    y += _spmm_rect(xs, g_rp, g_ci)
            assert np.array_equal(y, want[p.row_base:p.row_base + p.num_rows])
    finally:
        os.environ.pop("TCGNN_DENSE_FRACTION", None)
