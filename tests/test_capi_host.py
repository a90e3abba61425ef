"""CPU tests of the product's host side: the C-ABI library loads and exports every symbol that
include/tcgnn_b200.h declares, the host SGT (tcgnn_sgt_cpu, multi-threaded) is bit-exact with the
reference's preprocess goldens, argument errors come back as status codes (never exit), and the
`TCGNN` module exposes the reference's operator names.  No device compute is called here."""
import os
import re

import numpy as np
import pytest

import tcgnn_oracle as orc
from _util import ROOT, golden_sgt_files, load_golden


@pytest.fixture(scope="module")
def capi():
    import tcgnn_capi
    if not os.path.exists(tcgnn_capi.LIB_PATH):
        import importlib.util
        spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "tc-gnn_atc23_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build_library()
        mod.build_binding()
    return tcgnn_capi


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "tcgnn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tcgnn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(capi):
    L = capi.lib()
    names = declared_symbols()
    assert {"tcgnn_sgt_cpu", "tcgnn_sgt_cuda", "tcgnn_sgt_cuda_panel", "tcgnn_plan_create", "tcgnn_plan_create_panel",
            "tcgnn_plan_destroy", "tcgnn_plan_info", "tcgnn_spmm_f32", "tcgnn_sddmm_f32", "tcgnn_version",
            "tcgnn_status_string", "tcgnn_last_error", "tcgnn_launch_count", "tcgnn_debug_umma"} <= set(names)
    for nm in names:
        assert hasattr(L, nm), f"{nm} declared in include/tcgnn_b200.h but not exported"
    assert L.tcgnn_version() >= 100
    assert L.tcgnn_status_string(0) == b"ok"
    assert L.tcgnn_status_string(-1) == b"invalid argument"


@pytest.mark.parametrize("path", golden_sgt_files(), ids=lambda p: p.split("sgt_")[-1][:-4])
@pytest.mark.parametrize("threads", [1, 4])
def test_host_sgt_matches_reference_golden(capi, path, threads):
    g = load_golden(path)
    rp = np.ascontiguousarray(g["row_pointers"], dtype=np.int32)
    ci = np.ascontiguousarray(g["column_index"], dtype=np.int32)
    n = int(g["num_nodes"])
    bp = np.full((n + 15) // 16, -7, dtype=np.int32)
    e2c = np.full(len(ci), -7, dtype=np.int32)
    e2r = np.full(len(ci), -7, dtype=np.int32)
    total = capi.sgt_cpu(rp, ci, n, bp, e2c, e2r, threads=threads)
    assert np.array_equal(bp, g["blockPartition"])
    assert np.array_equal(e2c, g["edgeToColumn"])
    assert np.array_equal(e2r, g["edgeToRow"])
    assert total == int(g["tc_blocks_printed"])


def test_host_sgt_multithreaded_large_graph_matches_oracle(capi):
    n = 40000
    rp, ci = orc.rmat_graph(n, 1_200_000, seed=12)
    bp = np.zeros((n + 15) // 16, dtype=np.int32)
    e2c = np.zeros(len(ci), dtype=np.int32)
    e2r = np.zeros(len(ci), dtype=np.int32)
    total = capi.sgt_cpu(rp, ci, n, bp, e2c, e2r, threads=0)
    o_bp, o_e2c, o_e2r, o_total = orc.sgt(rp, ci, n)
    assert np.array_equal(bp, o_bp) and np.array_equal(e2c, o_e2c) and np.array_equal(e2r, o_e2r)
    assert total == o_total


def test_host_sgt_bad_arguments_return_status(capi):
    L = capi.lib()
    assert L.tcgnn_sgt_cpu(None, None, 10, 0, 16, 8, None, None, None, None, 1) == -1
    assert b"bad argument" in L.tcgnn_last_error()
    rp = np.zeros(5, dtype=np.int32)
    bp = np.zeros(1, dtype=np.int32)
    assert L.tcgnn_sgt_cpu(capi._ptr(rp), None, 4, 0, 0, 8, capi._ptr(bp), None, None, None, 1) == -1
    # plan / op entry points validate before touching the device
    assert L.tcgnn_plan_create(None, None, None, None, None, 0, 0, 0, None, None) == -1
    assert L.tcgnn_spmm_f32(None, None, 0, None, None, 0, 0, None) == -1
    assert L.tcgnn_plan_destroy(None) == 0


def test_module_exposes_reference_operator_surface(capi):
    """Names of /root/reference TCGNN_conv/TCGNN.cpp:260-272 (+ the north star's SDDMM_forward alias)."""
    import TCGNN
    for nm in ("preprocess", "preprocess_gpu", "forward", "forward_ef", "forward_AGNN", "backward", "backward_ef",
               "SDDMM_forward", "panel_forward", "panel_forward_ef", "panel_forward_AGNN", "preprocess_panel"):
        assert callable(getattr(TCGNN, nm)), nm


def test_module_preprocess_on_cpu_tensors_matches_golden(capi, capfd):
    """TCGNN.preprocess(edgeList, nodePointer, N, 16, 8, bp, e2c, e2r) -- the reference's call
    (main_tcgnn.py:50-54) -- on CPU tensors; also prints the reference's two log lines."""
    import torch
    import TCGNN
    g = load_golden([p for p in golden_sgt_files() if "cora_like" in p][0])
    n = int(g["num_nodes"])
    ci = torch.from_numpy(g["column_index"].astype(np.int32))
    rp = torch.from_numpy(g["row_pointers"].astype(np.int32))
    bp = torch.zeros((n + 15) // 16, dtype=torch.int32)
    e2c = torch.zeros(ci.numel(), dtype=torch.int32)
    e2r = torch.zeros(ci.numel(), dtype=torch.int32)
    TCGNN.preprocess(ci, rp, n, 16, 8, bp, e2c, e2r)
    out = capfd.readouterr().out
    assert np.array_equal(bp.numpy(), g["blockPartition"])
    assert np.array_equal(e2c.numpy(), g["edgeToColumn"])
    assert np.array_equal(e2r.numpy(), g["edgeToRow"])
    t = int(g["tc_blocks_printed"])
    assert f"TC_Blocks:\t{t}\nExp_Edges:\t{t * 128}\n" in out


def test_module_rejects_cpu_tensors_for_compute(capi):
    """Like the reference's CHECK_INPUT (TCGNN.cpp:54-56): a CPU tensor raises, nothing falls back."""
    import torch
    import TCGNN
    x = torch.zeros(16, 16)
    i = torch.zeros(17, dtype=torch.int32)
    e = torch.zeros(0, dtype=torch.int32)
    b = torch.ones(1, dtype=torch.int32)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        TCGNN.forward(x, i, e, b, e, e)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        TCGNN.forward_ef(x, i, e, b, e, e)


def test_optional_entry_points_validate_arguments_without_a_gpu(capi):
    """Argument checks of the round / push / *_ex entry points run before any CUDA call."""
    import ctypes as C
    L = capi.lib()
    assert L.tcgnn_round_tf32(None, 4, None, 4, 1, 4, None) == -1
    assert L.tcgnn_round_tf32_multicast(None, 4, None, 4, 1, 4, None) == -1
    assert b"tcgnn_round_tf32" in L.tcgnn_last_error()
    buf = (C.c_float * 64)()
    addr = C.addressof(buf)
    assert L.tcgnn_round_tf32(addr, 4, addr + 4, 4, 1, 4, None) == -1          # out not 16-byte aligned
    assert L.tcgnn_round_tf32(addr, 4, addr, 6, 1, 4, None) == -1              # ldo % 4 != 0
    assert L.tcgnn_push_rows(None, None, 0, None, None, 0, 4, None) == -1
    assert L.tcgnn_spmm_f32_ex(None, None, 0, None, None, 0, 0, 0, None) == -1
    assert L.tcgnn_sddmm_f32_ex(None, None, 0, None, 0, 0, None) == -1


def test_round2_entry_points_validate_arguments_without_a_gpu(capi):
    """tcgnn_agnn_f32 / tcgnn_csr_transpose / tcgnn_gather_rows / tcgnn_stream_wait_flag / the host-buffer entries
    reject bad arguments before any CUDA call; the module exposes their torch-level counterparts."""
    import ctypes as C
    import TCGNN
    L = capi.lib()
    assert L.tcgnn_agnn_f32(None, None, 0, None, None, 0, None, None, 0, 0, None) == -1
    assert L.tcgnn_sddmm_f32_host(None, None, 0, None, 0, None) == -1
    assert L.tcgnn_agnn_f32_host(None, None, 0, None, None, 0, None, 0, None) == -1
    assert L.tcgnn_csr_transpose(None, None, 4, 4, 0, None, None, None, None) == -1
    assert b"tcgnn_csr_transpose" in L.tcgnn_last_error()
    buf = (C.c_float * 64)()
    addr = C.addressof(buf)
    idx = (C.c_int32 * 4)()
    assert L.tcgnn_gather_rows(addr, 6, C.addressof(idx), 2, addr, None) == -1       # ld % 4 != 0
    assert L.tcgnn_gather_rows(addr + 4, 8, C.addressof(idx), 2, addr, None) == -1   # src not 16-byte aligned
    assert L.tcgnn_gather_rows(None, 8, None, 0, None, None) == 0                    # nothing to do
    assert L.tcgnn_stream_wait_flag(None, 1, 10, None, None) == -1
    for nm in ("forward_AGNN_fused", "forward_AGNN_tile", "backward_T", "backward_T_AGNN", "csr_transpose",
               "source_forward", "gather_rows", "stream_wait_flag", "forward_host", "forward_ef_host",
               "forward_AGNN_host"):
        assert callable(getattr(TCGNN, nm)), nm


def test_production_library_ignores_the_profiling_switches(capi):
    """TCGNN_ABLATE / TCGNN_TUNE / TCGNN_TRACE / TCGNN_PRESET are compiled out unless build.py --debug-switches
    (switches that produce wrong results must not be live in the shipped library)."""
    import subprocess
    lib = os.path.join(ROOT, "tc-gnn_atc23_b200", "libtcgnn_b200.so")
    strings = subprocess.run(["strings", "-a", lib], stdout=subprocess.PIPE, text=True).stdout
    # ... and so is the experimental register-gather SpMM (TCGNN_SPMM_TS: correct but 1.9x slower)
    for env in ("TCGNN_ABLATE", "TCGNN_TRACE_CTA", "TCGNN_PRESET", "TCGNN_SPMM_TS"):
        assert env not in strings, f"{env} is reachable in the production library"
