"""API constants of the TC-GNN aggregation path (reference: config.py:1-8, TCGNN_conv/config.h:4-6).

BLK_H x BLK_W is the shape of one condensed TC block; it defines the SGT arrays bit for bit, so it
is part of the operator contract and not a tuning knob (the tcgen05 kernels use their own internal
tile shape)."""

BLK_H = 16
BLK_W = 8
WARP_SIZE = 32


def func(x):
    """Degree clamp used for the (unused) degree normalisation vector, reference config.py:5-9."""
    return x if x > 0 else 1
