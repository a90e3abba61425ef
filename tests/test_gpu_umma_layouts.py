"""Pins, on real hardware, the shared-memory layouts and tcgen05 descriptor encodings the
production kernels rely on (tc-gnn_atc23_b200/csrc/spmm_tc.cu, sddmm_tc.cu) by running single
MMAs through `tcgnn_debug_umma` on operands built here with NumPy.  All operand values are small
integers, exact in TF32, so the comparison is exact."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SW_NONE, SW_128B_BASE32B, SW_128B, SW_32B = 0, 1, 2, 6


def smem_desc(lbo, sbo, swizzle):
    return (((lbo >> 4) & 0x3FFF) << 16) | (((sbo >> 4) & 0x3FFF) << 32) | (1 << 46) | (swizzle << 61)


def idesc_tf32(m, n, a_mn, b_mn):
    return (1 << 4) | (2 << 7) | (2 << 10) | (int(a_mn) << 15) | (int(b_mn) << 16) | ((n >> 3) << 17) | ((m >> 4) << 24)


def sw128(row, chunk16):
    return row * 128 + ((chunk16 ^ (row & 7)) << 4)


def sw128_base32(row, chunk16):
    """128-byte rows, 32-byte swizzle granule (Swizzle<2,5,2>): the only MN-major layout tf32 operands have."""
    return row * 128 + ((chunk16 ^ ((row & 3) << 1)) << 4)


def image_mn_major_sw128(At, lbo=1024, sbo=512):
    """At[f, k]: f = M index (128), k = K index (8).  Atoms of 32 features x 4 k-rows (4 x 128 B):
    feature blocks `lbo` bytes apart, the two k-halves `sbo` bytes apart."""
    img = np.zeros(4096 // 4, dtype=np.float32)
    for f in range(At.shape[0]):
        for k in range(8):
            off = (f // 32) * lbo + (k // 4) * sbo + sw128_base32(k % 4, (f % 32) // 4) + (f % 4) * 4
            img[off // 4] = At[f, k]
    return img


def image_k_major_sw128(A, rows):
    """A[m, k] with 32 k per 128-byte row; 8-row atoms 1024 B apart."""
    img = np.zeros(rows * 32, dtype=np.float32)
    for m in range(rows):
        for k in range(32):
            off = (m // 8) * 1024 + sw128(m % 8, k // 4) + (k % 4) * 4
            img[off // 4] = A[m, k]
    return img


def image_b_kmajor_noswizzle(B, lbo, sbo):
    img = np.zeros(512 // 4, dtype=np.float32)
    for n in range(16):
        for k in range(8):
            off = (n // 8) * sbo + (k // 4) * lbo + (n % 8) * 16 + (k % 4) * 4
            img[off // 4] = B[n, k]
    return img


def image_b_kmajor_sw32(B):
    img = np.zeros(512 // 4, dtype=np.float32)
    for n in range(16):
        for k in range(8):
            chunk = (k // 4) ^ (((n % 8) >> 2) & 1)
            off = (n // 8) * 256 + (n % 8) * 32 + chunk * 16 + (k % 4) * 4
            img[off // 4] = B[n, k]
    return img


def _report(name, ok, err):
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/umma_probe.txt", "a") as fh:
        fh.write(f"{name}: {'OK' if ok else 'MISMATCH'} max_abs_err={err}\n")


@pytest.fixture(scope="module")
def capi():
    import tcgnn_capi
    return tcgnn_capi


def test_spmm_layout_mn_major_a_kmajor_b(capi):
    """The SpMM configuration: A = gathered feature rows (MN-major, 128B rows with 32B swizzle granule
    -- SWIZZLE_128B_BASE32B, the only MN-major layout for tf32 -- feature blocks LBO=1024, k-halves SBO=512),
    B = 16x8 sparse tile (K-major, no swizzle, LBO=128 between K chunks, SBO=256 between 8-row groups)."""
    rng = np.random.default_rng(0)
    At = rng.integers(-8, 9, size=(128, 8)).astype(np.float32)
    B = rng.integers(-4, 5, size=(16, 8)).astype(np.float32)
    want = At @ B.T
    got = capi.debug_umma(image_mn_major_sw128(At), image_b_kmajor_noswizzle(B, 128, 256),
                          smem_desc(1024, 512, SW_128B_BASE32B), smem_desc(128, 256, SW_NONE),
                          idesc_tf32(128, 16, True, False), 1, 0, 0)
    err = float(np.abs(got - want).max())
    _report("spmm_layout(A MN-major SW128, B K-major none LBO128/SBO256)", err == 0.0, err)
    if err != 0.0:
        # diagnostics for the next iteration: try the alternative encodings and dump what came back
        alt = capi.debug_umma(image_mn_major_sw128(At), image_b_kmajor_noswizzle(B, 128, 256),
                              smem_desc(1024, 512, SW_128B_BASE32B), smem_desc(256, 128, SW_NONE),
                              idesc_tf32(128, 16, True, False), 1, 0, 0)
        _report("  alt B desc LBO256/SBO128", bool(np.array_equal(alt, want)), float(np.abs(alt - want).max()))
        alt = capi.debug_umma(image_mn_major_sw128(At), image_b_kmajor_sw32(B),
                              smem_desc(1024, 512, SW_128B_BASE32B), smem_desc(16, 256, SW_32B),
                              idesc_tf32(128, 16, True, False), 1, 0, 0)
        _report("  alt B SW32", bool(np.array_equal(alt, want)), float(np.abs(alt - want).max()))
        alt = capi.debug_umma(image_mn_major_sw128(At, 512, 2048), image_b_kmajor_noswizzle(B, 128, 256),
                              smem_desc(512, 2048, SW_128B_BASE32B), smem_desc(128, 256, SW_NONE),
                              idesc_tf32(128, 16, True, False), 1, 0, 0)
        _report("  alt A layout LBO512/SBO2048", bool(np.array_equal(alt, want)), float(np.abs(alt - want).max()))
        np.save("gpurun_out/umma_probe_spmm_alt_got.npy", alt)
        np.save("gpurun_out/umma_probe_spmm_got.npy", got)
        np.save("gpurun_out/umma_probe_spmm_want.npy", want)
    assert err == 0.0


def test_spmm_layout_identity_maps_lanes(capi):
    """TMEM lane == feature, TMEM column == window row: A row f holds f in k=0, B picks k=0 for n=3."""
    At = np.zeros((128, 8), np.float32)
    At[:, 0] = np.arange(128)
    At[:, 5] = 1000 + np.arange(128)
    B = np.zeros((16, 8), np.float32)
    B[3, 0] = 1.0
    B[9, 5] = 1.0
    got = capi.debug_umma(image_mn_major_sw128(At), image_b_kmajor_noswizzle(B, 128, 256),
                          smem_desc(1024, 512, SW_128B_BASE32B), smem_desc(128, 256, SW_NONE),
                          idesc_tf32(128, 16, True, False), 1, 0, 0)
    want = At @ B.T
    ok = np.array_equal(got, want)
    _report("spmm_layout identity", ok, float(np.abs(got - want).max()))
    if not ok:
        np.save("gpurun_out/umma_probe_identity_got.npy", got)
    assert ok


def test_sddmm_layout_kmajor_both_with_k_advance(capi):
    """The SDDMM configuration: both operands K-major, 128B swizzle, SBO=1024; four K=8 steps that
    advance the descriptor start address by 32 bytes inside the swizzled 128-byte rows."""
    rng = np.random.default_rng(1)
    A = rng.integers(-8, 9, size=(128, 32)).astype(np.float32)
    B = rng.integers(-4, 5, size=(16, 32)).astype(np.float32)
    want = A @ B.T
    got = capi.debug_umma(image_k_major_sw128(A, 128), image_k_major_sw128(B, 16),
                          smem_desc(16, 1024, SW_128B), smem_desc(16, 1024, SW_128B),
                          idesc_tf32(128, 16, False, False), 4, 32, 32)
    err = float(np.abs(got - want).max())
    _report("sddmm_layout(K-major SW128 both, 4 k-steps of 32B)", err == 0.0, err)
    if err != 0.0:
        np.save("gpurun_out/umma_probe_sddmm_got.npy", got)
        np.save("gpurun_out/umma_probe_sddmm_want.npy", want)
    assert err == 0.0


def test_tf32_operand_handling_is_exact_after_rna(capi):
    """Operands that are already RNA-rounded to TF32 must go through the tensor core unchanged
    (this is what makes the kernels bit-compatible with the reference's cvt.rna + wmma path);
    also records what the hardware does with unrounded fp32 bits (truncate vs round), for DESIGN.md."""
    import tcgnn_oracle as orc
    rng = np.random.default_rng(2)
    At = orc.tf32_rna(rng.standard_normal((128, 8)).astype(np.float32))
    B = np.zeros((16, 8), np.float32)
    for n in range(8):
        B[n, n] = 1.0
    got = capi.debug_umma(image_mn_major_sw128(At), image_b_kmajor_noswizzle(B, 128, 256),
                          smem_desc(1024, 512, SW_128B_BASE32B), smem_desc(128, 256, SW_NONE),
                          idesc_tf32(128, 16, True, False), 1, 0, 0)
    assert np.array_equal(got[:, :8], At)
    raw = rng.standard_normal((128, 8)).astype(np.float32)
    got = capi.debug_umma(image_mn_major_sw128(raw), image_b_kmajor_noswizzle(B, 128, 256),
                          smem_desc(1024, 512, SW_128B_BASE32B), smem_desc(128, 256, SW_NONE),
                          idesc_tf32(128, 16, True, False), 1, 0, 0)
    trunc = (raw.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    mode = "truncate" if np.array_equal(got[:, :8], trunc) else ("rna" if np.array_equal(got[:, :8], orc.tf32_rna(raw)) else "other")
    _report(f"raw fp32 operand handling by kind::tf32 = {mode}", True, 0.0)


@pytest.mark.parametrize("ksteps", [1, 3])
def test_a_operand_from_tensor_memory(capi, ksteps):
    """The register-gather SpMM: A (gathered rows, lane = feature, one column per neighbour k) written to TMEM with
    tcgen05.st.32x32b.x8 and consumed from there (K-major), B = the 16x8 tile from shared memory as before."""
    rng = np.random.default_rng(3 + ksteps)
    A = rng.integers(-8, 9, size=(128, 8 * ksteps)).astype(np.float32)
    Bs = rng.integers(-4, 5, size=(ksteps, 16, 8)).astype(np.float32)
    want = sum(A[:, 8 * s:8 * s + 8] @ Bs[s].T for s in range(ksteps))
    b_img = np.concatenate([image_b_kmajor_noswizzle(Bs[s], 128, 256) for s in range(ksteps)])
    got = capi.debug_umma(np.ascontiguousarray(A), b_img, 0, smem_desc(128, 256, SW_NONE),
                          idesc_tf32(128, 16, False, False), ksteps, 0, 512)
    err = float(np.abs(got - want).max())
    _report(f"A from TMEM (tcgen05.st x8, K-major), ksteps={ksteps}", err == 0.0, err)
    if err != 0.0:
        np.save(f"gpurun_out/umma_probe_ts_got_{ksteps}.npy", got)
        np.save(f"gpurun_out/umma_probe_ts_want_{ksteps}.npy", want)
    assert err == 0.0
