"""`mimc_hash(list of ints)` = MiMC7-91 multi_hash(message, key 0), the transcript hash of the Rust prover (`mimc-rs`,
rust/src/gkr/sumcheck.rs:45,84,129,152), taken over the coefficient list without its leading zeros; the algorithm
itself is the oracle's restatement (oracle/l0_reference.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))))
from oracle import l0_reference as _l0  # noqa: E402


# True: hash the message without its leading zero coefficients (lines the prototype up with the Rust prover's own
# transcript on full-degree messages).  False: hash the prototype's list exactly as it is -- the comparison then runs
# OUR provers with a transcript callback that pads a message to the prototype's four coefficients, which also covers
# circuits with messages of lower degree (tests/golden/make_refpy_vectors.py, "gkr_native_transcript").
STRIP_LEADING_ZEROS = True


def mimc_hash(x, k=0):
    msg = [int(v) for v in x]
    while STRIP_LEADING_ZEROS and len(msg) > 1 and msg[0] == 0:
        msg = msg[1:]
    return _l0.multi_hash(msg, int(k))
