"""gkr_b200 -- B200-native GKR prover hot path (sumcheck rounds, wiring-predicate sums, MLE evaluation)
behind the reference's prover seam.  Compute lives in gkr_b200/libgkr_b200.so (CUDA, sm_100a);
this package is the thin host mirror of the reference interface.  No CPU fallback."""
from .field import P, fr_to_ints, ints_to_fr  # noqa: F401
from .prover import (DenseLayer, DenseProof, GKRCircuit, Input, Layer, Proof, Prover, prove)  # noqa: F401

__all__ = ["P", "fr_to_ints", "ints_to_fr", "DenseLayer", "DenseProof", "GKRCircuit", "Input", "Layer", "Proof",
           "Prover", "prove"]
