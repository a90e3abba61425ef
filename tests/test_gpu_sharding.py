"""Row-panel sharding on the GPU: panel kernels (tcgnn_plan_create_panel through TCGNN.panel_*) must
reproduce the single-GPU operators, first on one device (panels computed one after the other), then
with real ranks over NCCL when the box has >= 2 GPUs."""
import os
import socket

import numpy as np
import pytest

import tcgnn_oracle as orc
from _util import assert_normwise, features, sgt_arrays, to_dev

pytestmark = pytest.mark.gpu


def _full(rp, ci, n, x, w=None):
    import TCGNN
    bp, e2c, e2r = sgt_arrays(rp, ci, n)
    d_rp, d_ci, d_bp, d_e2c, d_e2r, d_x = to_dev(rp, ci, bp, e2c, e2r, x)
    g = (d_rp, d_ci, d_bp, d_e2c, d_e2r)
    y = TCGNN.forward(d_x, *g)[0]
    ef = TCGNN.forward_ef(d_x, *g)[0]
    return g, d_x, y, ef


@pytest.mark.parametrize("world", [2, 3, 8])
def test_panels_on_one_device_reproduce_full_graph(world):
    import torch
    from sharding import RowPanel
    n = 6007
    rp, ci = orc.rmat_graph(n, 150000, seed=41)
    xi = features(n, 96, seed=42, kind="ints")
    g, d_x, y_full, ef_full = _full(rp, ci, n, xi)
    ys, efs = [], []
    for rank in range(world):
        p = RowPanel(g[0], g[1], rank, world)           # device SGT of the panel (tcgnn_sgt_cuda_panel)
        assert np.array_equal(p.blockPartition.cpu().numpy(),
                              g[2].cpu().numpy()[p.row_base // 16:p.row_base // 16 + (p.num_rows + 15) // 16])
        ys.append(p.spmm(d_x))
        efs.append(p.sddmm(d_x))
    torch.cuda.synchronize()
    assert torch.equal(torch.cat(ys), y_full)            # integer features: exact whatever the slicing
    assert torch.equal(torch.cat(efs), ef_full)
    assert np.array_equal(y_full.cpu().numpy(), orc.spmm(xi, rp, ci))


def test_panel_weighted_and_float_features():
    import torch
    from sharding import RowPanel
    n = 4100
    rp, ci = orc.random_graph(n, 90000, seed=43)
    x = features(n, 64, seed=44)
    g, d_x, y_full, ef_full = _full(rp, ci, n, x)
    w = np.random.default_rng(45).standard_normal(len(ci)).astype(np.float32)
    d_w = torch.from_numpy(w).cuda()
    want = orc.spmm(x, rp, ci, w)
    scale = orc.spmm_abs(x, rp, ci, w)
    for rank in range(3):
        p = RowPanel(g[0], g[1], rank, 3)
        att = d_w[p.edge_begin:p.edge_end].reshape(1, -1).contiguous()
        y = p.spmm(d_x, att).cpu().numpy()
        sl = slice(p.row_base, p.row_base + p.num_rows)
        assert_normwise(y, want[sl], scale[sl], 1e-5, f"weighted panel {rank}")
        ef = p.sddmm(d_x)
        assert torch.equal(ef, ef_full[p.edge_begin:p.edge_end])     # one accumulation chain per edge: bit-identical


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _nccl_worker(rank, world, port, q):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    for p in (os.path.join(root, "tc-gnn_atc23_b200"), os.path.join(root, "oracle"), here):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    from sharding import RowPanel
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        n, d = 20011, 128
        rp, ci = orc.rmat_graph(n, 600000, seed=51)
        x = features(n, d, seed=52, kind="ints")
        t_rp, t_ci = torch.from_numpy(rp).cuda(), torch.from_numpy(ci).cuda()
        p = RowPanel(t_rp, t_ci, rank, world)
        x_local = torch.from_numpy(x[p.row_base:p.row_base + p.num_rows]).cuda()
        want = orc.spmm(x, rp, ci)[p.row_base:p.row_base + p.num_rows]
        ok = True
        # the exchange four ways: NCCL uneven all-gather, fused round + direct P2P push, fused round + two-phase
        # balanced push, auto (NVSwitch multicast when the group supports it); twice each, because the second
        # step reuses the symmetric buffer behind the barriers
        for mode in ("nccl", "p2p", "p2p2", "auto"):
            os.environ["TCGNN_EXCHANGE"] = mode
            for _ in range(2):
                y = p.spmm(p.all_gather(x_local, round_tf32=True), x_is_tf32=True)   # exchange, then panel SpMM
                ok = ok and np.array_equal(y.cpu().numpy(), want)
        # the overlapped exchange (per-source-panel partial products behind copy-engine pushes + flags), with every
        # source shipping its whole panel and with every source shipping packed referenced rows; several steps each
        # (double-buffered receive area, monotonic flags)
        os.environ["TCGNN_EXCHANGE"] = "overlap"
        for frac in ("0.0", "2.0"):
            os.environ["TCGNN_DENSE_FRACTION"] = frac
            po = RowPanel(t_rp, t_ci, rank, world, bounds=p.bounds)
            for step in range(4):
                xs = x_local * float(step + 1)
                y = po.aggregate(xs)
                ok = ok and np.array_equal(y.cpu().numpy(), want * float(step + 1))
            po.overlap_check()
            st = po.overlap_stats(d)
            ok = ok and st is not None and (st["packed_sources"] == world - 1 if frac == "2.0" else st["dense_sources"] == world - 1)
            if frac == "2.0":
                ok = ok and st["recv_rows"] <= st["full_gather_rows"]
        os.environ["TCGNN_EXCHANGE"] = "auto"
        y2, ef = p.agnn_aggregate(x_local, torch.full((1, 1), 0.5, device="cuda"))
        torch.cuda.synchronize()
        ef_want = orc.sddmm(x, rp, ci)[p.edge_begin:p.edge_end]
        ok = ok and np.array_equal(ef.cpu().numpy(), ef_want)
        yw = orc.spmm(x, rp, ci, orc.sddmm(x, rp, ci) * 0.5)[p.row_base:p.row_base + p.num_rows]
        ok = ok and np.allclose(y2.cpu().numpy(), yw, rtol=2e-3, atol=1e-2 * np.abs(yw).max())
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_two_ranks_over_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)


def test_prerounded_operands_give_identical_results():
    """tcgnn_round_tf32 + TCGNN_X_IS_TF32 (what the sharded path ships between GPUs) == rounding inside the op."""
    import torch
    import TCGNN
    n = 6000
    rp, ci = orc.rmat_graph(n, 150000, seed=31)
    bp, e2c, e2r, _ = orc.sgt(rp, ci, n)
    g = [torch.from_numpy(a).cuda() for a in (rp, ci, bp, e2c, e2r)]
    x = torch.randn(n, 96, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    xr = TCGNN.round_tf32(x)
    assert torch.equal(xr, torch.from_numpy(orc.tf32_rna(x.cpu().numpy())).cuda())
    # windows split over CTAs are combined with fp32 atomics (order varies run to run): compare the ops on
    # integer-valued data, where every partial sum is exact
    xi = torch.randint(-8, 9, (n, 96), device="cuda", generator=torch.Generator(device="cuda").manual_seed(4)).float()
    xir = TCGNN.round_tf32(xi)
    assert torch.equal(xir, xi)
    y0 = TCGNN.forward(xi, *g)[0]
    y1 = TCGNN.panel_forward(xir, 0, *g, x_is_tf32=True)[0]
    assert torch.equal(y0, y1)
    e0 = TCGNN.forward_ef(x, *g)[0]                     # SDDMM has no cross-CTA accumulation
    e1 = TCGNN.panel_forward_ef(xr, 0, *g, x_is_tf32=True)[0]
    assert torch.equal(e0, e1)
    w = torch.randint(-3, 4, (1, len(ci)), device="cuda").float()
    z0 = TCGNN.forward_AGNN(xi, g[0], g[1], w, *g[2:])[0]
    z1 = TCGNN.panel_forward_AGNN(xir, 0, g[0], g[1], w, *g[2:], x_is_tf32=True)[0]
    assert torch.equal(z0, z1)
    # and on random-normal data within the accumulation-order tolerance
    yr = TCGNN.panel_forward(xr, 0, *g, x_is_tf32=True)[0]
    assert_normwise(yr.cpu().numpy(), orc.spmm(x.cpu().numpy(), rp, ci), orc.spmm_abs(x.cpu().numpy(), rp, ci), 1e-5,
                    "pre-rounded SpMM")
