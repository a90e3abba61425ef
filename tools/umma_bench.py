#!/usr/bin/env python3
"""tcgen05.mma.kind::tf32 issue-rate probe (GPU box): cycles per MMA for the shapes the kernels use."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tc-gnn_atc23_b200"))
import torch  # noqa: F401,E402  (initialises CUDA)
import tcgnn_capi  # noqa: E402

torch.zeros(1, device="cuda")
L = tcgnn_capi.lib()
SW_NONE, SW_128B_BASE32B, SW_128B = 0, 1, 2


def smem_desc(lbo, sbo, swizzle):
    return (((lbo >> 4) & 0x3FFF) << 16) | (((sbo >> 4) & 0x3FFF) << 32) | (1 << 46) | (swizzle << 61)


def idesc_tf32(m, n, a_mn, b_mn):
    return (1 << 4) | (2 << 7) | (2 << 10) | (int(a_mn) << 15) | (int(b_mn) << 16) | ((n >> 3) << 17) | ((m >> 4) << 24)


def run(name, adesc, bdesc, idesc, n_mma, n_acc, acc_stride, n_a, a_step, n_b, b_step, grid=1):
    out = (C.c_int64 * 2)()
    st = L.tcgnn_debug_umma_bench(adesc, bdesc, idesc, n_mma, n_acc, acc_stride, n_a, a_step, n_b, b_step, grid, out, None)
    if st != 0:
        print(name, "FAILED", L.tcgnn_last_error().decode())
        return
    print(f"{name:70s} n={n_mma:5d} issue {out[0] / n_mma:7.1f} cyc/MMA   complete {out[1] / n_mma:7.1f} cyc/MMA", flush=True)


if len(sys.argv) > 1 and sys.argv[1] == "ts":
    # A operand from tensor memory (adesc == 0): the operand fetch reads shared memory for the 512-byte B tile only
    b_k = smem_desc(128, 256, SW_NONE)
    a_mn = smem_desc(1024, 512, SW_128B_BASE32B)
    for grid in (1, 148):
        run(f"A in smem   M128 N16 K8 x8 grid {grid}", a_mn, b_k, idesc_tf32(128, 16, True, False), 2048, 0, 16, 8, 4096, 8, 512, grid=grid)
        run(f"A in TMEM   M128 N16 K8 x8 grid {grid}", 0, b_k, idesc_tf32(128, 16, False, False), 2048, 0, 16, 48, 0, 8, 512, grid=grid)
    run("A in TMEM   M128 N32 K8 x8", 0, b_k, idesc_tf32(128, 32, False, False), 2048, 0, 32, 48, 0, 8, 512)
    run("A in TMEM   M128 N64 K8 x8", 0, b_k, idesc_tf32(128, 64, False, False), 2048, 0, 64, 48, 0, 8, 512)
    sys.exit(0)

k128 = smem_desc(16, 1024, SW_128B)               # SDDMM: K-major 128B swizzle
a_mn = smem_desc(1024, 512, SW_128B_BASE32B)      # SpMM A: MN-major, 4 KB per K=8 tile
b_k = smem_desc(128, 256, SW_NONE)                # SpMM B: K-major 16x8
k128 = smem_desc(16, 1024, SW_128B)               # SDDMM: K-major 128B swizzle
for n in (2048,):
    run("spmm M128 N16 K8, unrolled x8 (production fast path)", a_mn, b_k, idesc_tf32(128, 16, True, False), n, 0, 16, 8, 4096, 8, 512)
    run("spmm M128 N16 K8, unrolled x8, grid 148", a_mn, b_k, idesc_tf32(128, 16, True, False), n, 0, 16, 8, 4096, 8, 512, grid=148)
    run("spmm M128 N64 K8, unrolled x8", a_mn, b_k, idesc_tf32(128, 64, True, False), n, 0, 64, 8, 4096, 8, 512)
    run("spmm M128 N256 K8, unrolled x8", a_mn, b_k, idesc_tf32(128, 256, True, False), n, 0, 256, 8, 4096, 8, 512)
    run("spmm M64 N16 K8, unrolled x8", a_mn, b_k, idesc_tf32(64, 16, True, False), n, 0, 16, 8, 4096, 8, 512)
    run("K-major sw128 M128 N16 K8, unrolled x8", k128, k128, idesc_tf32(128, 16, False, False), n, 0, 16, 8, 4096, 8, 512)
    for nacc in (1, 2, 4):
        run(f"spmm M128 N16 K8 A=MN-major sw128/32 acc x{nacc}", a_mn, b_k, idesc_tf32(128, 16, True, False), n, nacc, 16, 8, 4096, 8, 512)
    run("spmm same, grid 148", a_mn, b_k, idesc_tf32(128, 16, True, False), n, 1, 16, 8, 4096, 8, 512, grid=148)
    run("spmm M64 N16 K8 A=MN-major (half features)", a_mn, b_k, idesc_tf32(64, 16, True, False), n, 1, 16, 8, 4096, 8, 512)
    for nn in (32, 64, 128, 256):
        run(f"spmm M128 N{nn} K8 A=MN-major (B garbage layout)", a_mn, smem_desc(128, 256, SW_NONE), idesc_tf32(128, nn, True, False), n, 1, nn, 8, 4096, 1, 512)
    for nacc in (1, 2):
        run(f"sddmm M128 N16 K8 both K-major sw128 acc x{nacc} (32 B k-steps)", k128, k128, idesc_tf32(128, 16, False, False), n, nacc, 16, 4, 32, 4, 32)
    run("sddmm M128 N32 K8", k128, k128, idesc_tf32(128, 32, False, False), n, 1, 32, 4, 32, 4, 32)
    run("sddmm M128 N64 K8", k128, k128, idesc_tf32(128, 64, False, False), n, 1, 64, 4, 32, 4, 32)
    run("sddmm M128 N128 K8", k128, k128, idesc_tf32(128, 128, False, False), n, 1, 128, 4, 32, 4, 32)
