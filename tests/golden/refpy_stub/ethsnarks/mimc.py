"""`mimc_hash(list of ints)` = MiMC7-91 multi_hash(message, key 0), the transcript hash of the Rust prover (`mimc-rs`,
rust/src/gkr/sumcheck.rs:45,84,129,152), taken over the coefficient list without its leading zeros; the algorithm
itself is the oracle's restatement (oracle/l0_reference.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))))
from oracle import l0_reference as _l0  # noqa: E402


def mimc_hash(x, k=0):
    msg = [int(v) for v in x]
    while len(msg) > 1 and msg[0] == 0:      # the prototype always lists degree + 1 coefficients (see make_refpy_vectors.py)
        msg = msg[1:]
    return _l0.multi_hash(msg, int(k))
