"""GPU SGT (tcgnn_sgt_cuda / TCGNN.preprocess_gpu) must be bit-exact with the reference's
preprocess: checked against the golden arrays produced by the reference itself and against the
oracle on graphs that exercise both device code paths (in-smem sort and hub-window bitmap)."""
import numpy as np
import pytest

import tcgnn_oracle as orc
from _util import golden_sgt_files, load_golden, to_dev

pytestmark = pytest.mark.gpu


def run_gpu_sgt(rp, ci, n, via="capi"):
    import torch
    d_rp, d_ci = to_dev(rp, ci)
    w = (n + 15) // 16
    bp = torch.full((w,), -7, dtype=torch.int32, device="cuda")
    e2c = torch.full((len(ci),), -7, dtype=torch.int32, device="cuda")
    e2r = torch.full((len(ci),), -7, dtype=torch.int32, device="cuda")
    if via == "capi":
        import tcgnn_capi
        total = tcgnn_capi.sgt_cuda(d_rp, d_ci, n, bp, e2c, e2r)
    else:
        import TCGNN
        TCGNN.preprocess_gpu(d_ci, d_rp, n, 16, 8, bp, e2c, e2r)
        total = None
    torch.cuda.synchronize()
    return bp.cpu().numpy(), e2c.cpu().numpy(), e2r.cpu().numpy(), total


@pytest.mark.parametrize("path", golden_sgt_files(), ids=lambda p: p.split("sgt_")[-1][:-4])
def test_gpu_sgt_matches_reference_golden(path):
    g = load_golden(path)
    n = int(g["num_nodes"])
    bp, e2c, e2r, total = run_gpu_sgt(g["row_pointers"], g["column_index"], n)
    assert np.array_equal(bp, g["blockPartition"])
    assert np.array_equal(e2c, g["edgeToColumn"])
    assert np.array_equal(e2r, g["edgeToRow"])
    assert total == int(g["tc_blocks_printed"])


def test_gpu_sgt_hub_windows_use_bitmap_path():
    n = 30000
    rng = np.random.default_rng(30)
    src = np.concatenate([np.full(20000, 5), np.full(9000, 12345), rng.integers(0, n, 200000)])
    dst = np.concatenate([rng.choice(n, 20000, replace=False), rng.choice(n, 9000, replace=False),
                          rng.integers(0, n, 200000)])
    rp, ci = orc.csr_from_edges(src, dst, n)
    want = orc.sgt(rp, ci, n)
    bp, e2c, e2r, total = run_gpu_sgt(rp, ci, n, via="module")
    assert np.array_equal(bp, want[0]) and np.array_equal(e2c, want[1]) and np.array_equal(e2r, want[2])


def test_gpu_sgt_mid_size_rmat():
    n = 100000
    rp, ci = orc.rmat_graph(n, 3_000_000, seed=31)
    want = orc.sgt(rp, ci, n)
    bp, e2c, e2r, total = run_gpu_sgt(rp, ci, n)
    assert np.array_equal(bp, want[0]) and np.array_equal(e2c, want[1]) and np.array_equal(e2r, want[2])
    assert total == want[3]
