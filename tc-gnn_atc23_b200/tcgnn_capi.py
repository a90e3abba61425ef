"""ctypes view of the C ABI declared in include/tcgnn_b200.h (libtcgnn_b200.so).

The torch extension module `TCGNN` is the normal way in; this wrapper exists so tests and the
benchmark can drive the raw C entry points (device pointers + sizes) exactly as a foreign-language
host would.  There is no fallback of any kind: if the shared library is missing, loading raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtcgnn_b200.so")

_lib = None

i32p = C.POINTER(C.c_int32)
f32p = C.POINTER(C.c_float)


class TcgnnError(RuntimeError):
    pass


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TcgnnError(f"{LIB_PATH} is missing -- run `python tc-gnn_atc23_b200/build.py` (there is no fallback path)")
        L = C.CDLL(LIB_PATH)
        vp, i32, i64, u32, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64
        L.tcgnn_version.restype = C.c_int
        L.tcgnn_status_string.restype = C.c_char_p
        L.tcgnn_status_string.argtypes = [C.c_int]
        L.tcgnn_last_error.restype = C.c_char_p
        L.tcgnn_launch_count.restype = i64
        L.tcgnn_launch_count.argtypes = [C.c_int]
        sgt_args = [vp, vp, i32, i64, i32, i32, vp, vp, vp, C.POINTER(i64)]
        L.tcgnn_sgt_cpu.argtypes = sgt_args + [i32]
        L.tcgnn_sgt_cuda.argtypes = sgt_args + [vp]
        L.tcgnn_sgt_cuda_panel.argtypes = [vp, vp, i32, i32, i64, i32, i32, vp, vp, vp, C.POINTER(i64), vp]
        L.tcgnn_plan_create.argtypes = [vp, vp, vp, vp, vp, i32, i64, i32, vp, C.POINTER(vp)]
        L.tcgnn_plan_create_panel.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i64, i32, vp, C.POINTER(vp)]
        L.tcgnn_plan_destroy.argtypes = [vp]
        L.tcgnn_plan_info.argtypes = [vp, C.POINTER(i64)]
        L.tcgnn_spmm_f32.argtypes = [vp, vp, i64, vp, vp, i64, i32, vp]
        L.tcgnn_sddmm_f32.argtypes = [vp, vp, i64, vp, i32, vp]
        L.tcgnn_debug_umma.argtypes = [vp, i32, vp, i32, u64, u64, u32, i32, i32, i32, vp, i32, vp]
        L.tcgnn_spmm_f32_ex.argtypes = [vp, vp, i64, vp, vp, i64, i32, u32, vp]
        L.tcgnn_sddmm_f32_ex.argtypes = [vp, vp, i64, vp, i32, u32, vp]
        L.tcgnn_round_tf32.argtypes = [vp, i64, vp, i64, i64, i32, vp]
        L.tcgnn_spmm_f32_host.argtypes = [vp, vp, i64, vp, vp, i64, i32, vp]
        L.tcgnn_spmm_f32_host.restype = C.c_int
        L.tcgnn_push_rows.argtypes = [vp, vp, i32, vp, vp, i32, i64, vp]
        L.tcgnn_push_rows.restype = C.c_int
        L.tcgnn_round_tf32_multicast.argtypes = [vp, i64, vp, i64, i64, i32, vp]
        L.tcgnn_agnn_f32.argtypes = [vp, vp, i64, vp, vp, i64, vp, vp, i32, u32, vp]
        L.tcgnn_csr_transpose.argtypes = [vp, vp, i32, i32, i64, vp, vp, vp, vp]
        L.tcgnn_gather_rows.argtypes = [vp, i64, vp, i64, vp, vp]
        L.tcgnn_stream_wait_flag.argtypes = [vp, i32, i32, vp, vp]
        L.tcgnn_stream_wait_flag_dev.argtypes = [vp, vp, i32, vp, vp]
        L.tcgnn_stream_wait_flag_dev.restype = C.c_int
        L.tcgnn_sddmm_f32_host.argtypes = [vp, vp, i64, vp, i32, vp]
        L.tcgnn_agnn_f32_host.argtypes = [vp, vp, i64, vp, vp, i64, vp, i32, vp]
        for name in ("tcgnn_agnn_f32", "tcgnn_csr_transpose", "tcgnn_gather_rows", "tcgnn_stream_wait_flag",
                     "tcgnn_sddmm_f32_host", "tcgnn_agnn_f32_host"):
            getattr(L, name).restype = C.c_int
        for name in ("tcgnn_spmm_f32_ex", "tcgnn_sddmm_f32_ex", "tcgnn_round_tf32", "tcgnn_round_tf32_multicast"):
            getattr(L, name).restype = C.c_int
        L.tcgnn_debug_umma_bench.argtypes = [u64, u64, u32, i32, i32, i32, i32, i32, i32, i32, i32, C.POINTER(i64), vp]
        L.tcgnn_debug_umma_bench.restype = C.c_int
        for name in ("tcgnn_sgt_cpu", "tcgnn_sgt_cuda", "tcgnn_sgt_cuda_panel", "tcgnn_plan_create", "tcgnn_plan_create_panel", "tcgnn_plan_destroy", "tcgnn_plan_info",
                     "tcgnn_spmm_f32", "tcgnn_sddmm_f32", "tcgnn_debug_umma"):
            getattr(L, name).restype = C.c_int
        _lib = L
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        L = lib()
        raise TcgnnError(f"{what}: {L.tcgnn_status_string(status).decode()} -- {L.tcgnn_last_error().decode()}")


def _ptr(t):
    """data pointer of a torch tensor / numpy array (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def sgt_cpu(row_ptr, col_idx, num_nodes, block_partition, edge_to_col, edge_to_row, blk_h=16, blk_w=8, threads=0):
    """Host SGT on numpy int32 arrays (or CPU torch tensors); returns the printed TC_Blocks total."""
    total = C.c_int64(0)
    check(lib().tcgnn_sgt_cpu(_ptr(row_ptr), _ptr(col_idx), num_nodes, len(col_idx), blk_h, blk_w,
                              _ptr(block_partition), _ptr(edge_to_col), _ptr(edge_to_row), C.byref(total), threads),
          "tcgnn_sgt_cpu")
    return total.value


def sgt_cuda(row_ptr, col_idx, num_nodes, block_partition, edge_to_col, edge_to_row, blk_h=16, blk_w=8):
    total = C.c_int64(0)
    check(lib().tcgnn_sgt_cuda(_ptr(row_ptr), _ptr(col_idx), num_nodes, col_idx.numel(), blk_h, blk_w,
                               _ptr(block_partition), _ptr(edge_to_col), _ptr(edge_to_row), C.byref(total), _stream()),
          "tcgnn_sgt_cuda")
    return total.value


class Plan:
    """Owns a tcgnn_plan; keeps the five device tensors alive (the plan borrows their memory)."""

    def __init__(self, row_ptr, col_idx, block_partition, edge_to_col, edge_to_row, num_cols=None, row_base=0):
        self._tensors = (row_ptr, col_idx, block_partition, edge_to_col, edge_to_row)
        self.num_nodes = row_ptr.numel() - 1
        self.num_edges = col_idx.numel()
        self.num_cols = self.num_nodes if num_cols is None else num_cols
        self.row_base = row_base
        self._h = C.c_void_p()
        check(lib().tcgnn_plan_create_panel(_ptr(row_ptr), _ptr(col_idx), _ptr(block_partition), _ptr(edge_to_col),
                                            _ptr(edge_to_row), self.num_nodes, self.num_cols, row_base,
                                            self.num_edges, block_partition.numel(), _stream(), C.byref(self._h)),
              "tcgnn_plan_create_panel")

    def info(self):
        buf = (C.c_int64 * 8)()
        check(lib().tcgnn_plan_info(self._h, buf), "tcgnn_plan_info")
        keys = ("num_nodes", "num_edges", "num_windows", "num_tiles", "plan_bytes", "pairs", "device", "sms")
        return dict(zip(keys, list(buf)))

    def spmm(self, x, y, edge_weight=None, dim=None):
        dim = x.shape[1] if dim is None else dim
        check(lib().tcgnn_spmm_f32(self._h, _ptr(x), x.stride(0), _ptr(edge_weight), _ptr(y), y.stride(0), dim,
                                   _stream()), "tcgnn_spmm_f32")
        return y

    def sddmm(self, x, out, dim=None):
        dim = x.shape[1] if dim is None else dim
        check(lib().tcgnn_sddmm_f32(self._h, _ptr(x), x.stride(0), _ptr(out), dim, _stream()), "tcgnn_sddmm_f32")
        return out

    def spmm_ex(self, x, y, edge_weight=None, flags=0, dim=None):
        """tcgnn_spmm_f32_ex: flags = X_IS_TF32 | ACCUMULATE | W_TILE_ORDER."""
        dim = x.shape[1] if dim is None else dim
        check(lib().tcgnn_spmm_f32_ex(self._h, _ptr(x), x.stride(0), _ptr(edge_weight), _ptr(y), y.stride(0), dim,
                                      flags, _stream()), "tcgnn_spmm_f32_ex")
        return y

    def agnn(self, x, attention_w, y, att_tile_out=None, edge_out=None, flags=0, dim=None):
        """tcgnn_agnn_f32: fused SDDMM -> x attention_w -> weighted SpMM."""
        dim = x.shape[1] if dim is None else dim
        check(lib().tcgnn_agnn_f32(self._h, _ptr(x), x.stride(0), _ptr(attention_w), _ptr(y), y.stride(0),
                                   _ptr(att_tile_out), _ptr(edge_out), dim, flags, _stream()), "tcgnn_agnn_f32")
        return y

    def close(self):
        if self._h:
            lib().tcgnn_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def debug_umma(a_image, b_image, adesc, bdesc, idesc, ksteps, a_step, b_step, ncols=16):
    """Run the layout probe; images are numpy uint8/float32 arrays; returns float32 [128, ncols]."""
    import numpy as np
    a = np.ascontiguousarray(a_image).view(np.uint8).reshape(-1)
    b = np.ascontiguousarray(b_image).view(np.uint8).reshape(-1)
    out = np.zeros((128, ncols), dtype=np.float32)
    check(lib().tcgnn_debug_umma(_ptr(a), a.size, _ptr(b), b.size, adesc, bdesc, idesc, ksteps, a_step, b_step,
                                 _ptr(out), ncols, _stream()), "tcgnn_debug_umma")
    return out


X_IS_TF32, ACCUMULATE, W_TILE_ORDER = 1, 2, 4


def csr_transpose(row_ptr, col_idx, num_cols=None):
    """tcgnn_csr_transpose on device int32 tensors -> (row_ptr_t, col_idx_t, edge_map_t)."""
    import torch
    n = row_ptr.numel() - 1
    m = n if num_cols is None else num_cols
    e = col_idx.numel()
    rp_t = torch.empty(m + 1, dtype=torch.int32, device=row_ptr.device)
    ci_t = torch.empty(e, dtype=torch.int32, device=row_ptr.device)
    map_t = torch.empty(e, dtype=torch.int32, device=row_ptr.device)
    check(lib().tcgnn_csr_transpose(_ptr(row_ptr), _ptr(col_idx), n, m, e, _ptr(rp_t), _ptr(ci_t), _ptr(map_t),
                                    _stream()), "tcgnn_csr_transpose")
    return rp_t, ci_t, map_t


def launch_count(reset=False) -> int:
    return lib().tcgnn_launch_count(1 if reset else 0)
