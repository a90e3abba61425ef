#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the REFERENCE ITSELF (run in the build container only).

SGT goldens: outputs of the reference's own compiled `preprocess` (oracle/_ref/TCGNN_ref*.so,
built by oracle/build_ref.sh from /root/reference/TCGNN_conv, unmodified) on seeded graphs that
cover the edge cases of SURVEY.md 8a/A1: N % 16 in {0, 1, 15}, E == 0, empty rows, empty
windows, a hub row, unsorted rows with duplicated columns, cora- and citeseer-sized graphs.
The `TC_Blocks` total the reference prints (TCGNN.cpp:225) is captured from its stdout.

KAT goldens: the reference's two hand-checkable fixtures (gnn_conv.py:13-23 gen_test_tensor,
gnn_conv.py:61 ones_like), evaluated in exact integer arithmetic (every value is a small
integer, exact in TF32 and fp32, so the expected SpMM output is unique).

Usage: python tests/golden/make_golden.py      (needs oracle/_ref; GPU not needed)
"""
import os
import re
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import TCGNN_ref  # noqa: E402  (the reference extension)
import tcgnn_oracle as orc  # noqa: E402  (graph generators only)


def ref_preprocess(rp, ci, n):
    """Call the reference preprocess and capture the TC_Blocks line it printf()s."""
    nwin = (n + 15) // 16
    bp = torch.zeros(nwin + 1, dtype=torch.int32)  # +1: the reference writes one past the end when n%16==0
    e2c = torch.zeros(len(ci), dtype=torch.int32)
    e2r = torch.zeros(len(ci), dtype=torch.int32)
    sys.stdout.flush()
    with tempfile.TemporaryFile(mode="w+b") as tmp:
        saved = os.dup(1)
        os.dup2(tmp.fileno(), 1)
        try:
            TCGNN_ref.preprocess(torch.from_numpy(ci.copy()), torch.from_numpy(rp.copy()), n, 16, 8, bp, e2c, e2r)
            import ctypes
            ctypes.CDLL(None).fflush(None)
        finally:
            os.dup2(saved, 1)
            os.close(saved)
        tmp.seek(0)
        text = tmp.read().decode()
    m = re.search(r"TC_Blocks:\s*(-?\d+)", text)
    return bp[:nwin].numpy(), e2c.numpy(), e2r.numpy(), int(m.group(1))


def csr_unsorted_with_dups(n, e, seed):
    rng = np.random.default_rng(seed)
    deg = rng.multinomial(e, np.ones(n) / n)
    rp = np.zeros(n + 1, dtype=np.int32)
    rp[1:] = np.cumsum(deg)
    ci = rng.integers(0, n, size=e, dtype=np.int64).astype(np.int32)  # unsorted, duplicates likely
    return rp, ci


def graphs():
    yield "uniform_n37", *orc.random_graph(37, 200, seed=1), 37
    yield "uniform_n48_mod0", *orc.random_graph(48, 300, seed=2), 48
    yield "uniform_n49_mod1", *orc.random_graph(49, 300, seed=3), 49
    yield "uniform_n63_mod15", *orc.random_graph(63, 500, seed=4), 63
    yield "empty_e0_n37", np.zeros(38, np.int32), np.zeros(0, np.int32), 37
    yield "empty_e0_n32", np.zeros(33, np.int32), np.zeros(0, np.int32), 32
    # empty windows in the middle: only rows 0..15 and 64..79 have edges
    rng = np.random.default_rng(5)
    src = np.concatenate([rng.integers(0, 16, 120), rng.integers(64, 80, 150)])
    dst = rng.integers(0, 100, len(src))
    yield "empty_windows_n100", *orc.csr_from_edges(src, dst, 100), 100
    # hub row: row 3 is connected to everyone
    n = 300
    src = np.concatenate([np.full(n, 3), rng.integers(0, n, 900)])
    dst = np.concatenate([np.arange(n), rng.integers(0, n, 900)])
    yield "hub_n300", *orc.csr_from_edges(src, dst, n), n
    yield "unsorted_dups_n200", *csr_unsorted_with_dups(200, 3000, 6), 200
    yield "rmat_n1000", *orc.rmat_graph(1000, 12000, seed=7), 1000
    yield "cora_like", *orc.random_graph(2708, 10858, seed=0), 2708
    yield "citeseer_like", *orc.random_graph(3327, 9464, seed=0), 3327


def main():
    for name, rp, ci, n in graphs():
        bp, e2c, e2r, printed = ref_preprocess(rp, ci, n)
        np.savez_compressed(os.path.join(HERE, f"sgt_{name}.npz"), row_pointers=rp, column_index=ci,
                            num_nodes=np.int64(n), blockPartition=bp, edgeToColumn=e2c, edgeToRow=e2r,
                            tc_blocks_printed=np.int64(printed))
        print(f"sgt_{name}: N={n} E={len(ci)} W={len(bp)} TC_Blocks(printed)={printed} sum(bp)={int(bp.sum())}")

    # KATs (exact integer arithmetic)
    for name, rp, ci, n in [("kat_n500", *orc.random_graph(500, 6000, seed=11), 500),
                            ("kat_n2000", *orc.rmat_graph(2000, 30000, seed=12), 2000)]:
        rows = np.repeat(np.arange(n), np.diff(rp))
        deg = np.diff(rp).astype(np.int64)
        nbr_sum = np.zeros(n, dtype=np.int64)
        np.add.at(nbr_sum, rows, ci.astype(np.int64))
        assert nbr_sum.max() < 2 ** 24  # exactly representable in fp32; every addend < 2^11 is exact in TF32
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), row_pointers=rp, column_index=ci,
                            num_nodes=np.int64(n), degree=deg, neighbour_id_sum=nbr_sum)
        print(f"{name}: N={n} E={len(ci)}")


if __name__ == "__main__":
    main()
