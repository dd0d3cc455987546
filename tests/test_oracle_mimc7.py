"""MiMC7 / keccak known answers against BOTH oracle implementations
(pure-Python L0 and C L1).  These public vectors are what pins the transcript; the reference's
own crate (mimc-rs, unpinned git dependency) is not available offline."""
from oracle import l0_reference as l0
from oracle import oracle as orc

KECCAK_EMPTY = "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
C1 = 20888961410941983456478427210666206549300505294776164667214940546594746570981
C2 = 15265126113435022738560151911929040668591755459209400716467504685752745317193
H_1_2 = 0x176c6eefc3fdf8d6136002d8e6f7a885bbd1c4e3957b93ddc1ec3ae7859f1a08
MH_123 = 0x25f5a6429a9764564be3955e6f56b0b9143c571528fd30a80ae6c27dc8b4a40c
MH_4 = 0x284bc1f34f335933a23a433b6ff3ee179d682cd5e5e2fcdd2d964afa85104beb


def test_keccak_empty():
    assert l0.keccak256(b"").hex() == KECCAK_EMPTY
    assert orc.keccak256(b"").hex() == KECCAK_EMPTY


def test_keccak_multiblock_agree():
    data = bytes(range(256)) * 3
    for n in (1, 31, 135, 136, 137, 272, 500):
        assert l0.keccak256(data[:n]) == orc.keccak256(data[:n])


def test_constants():
    c = l0.mimc7_constants()
    assert len(c) == 91 and c[0] == 0 and c[1] == C1 and c[2] == C2
    for i in range(91):
        assert orc.mimc7_constant(i) == c[i]


def test_hash_kats():
    for mod in (l0, orc):
        assert mod.mimc7_hash(1, 2) == H_1_2
        assert mod.multi_hash([1, 2, 3], 0) == MH_123
        assert mod.multi_hash([12, 45, 78, 41], 0) == MH_4


def test_multi_hash_empty_and_agreement():
    assert l0.multi_hash([], 0) == 0 and orc.multi_hash([], 0) == 0
    import random
    rng = random.Random(7)
    for _ in range(20):
        arr = [rng.randrange(l0.P) for _ in range(rng.randrange(1, 5))]
        assert l0.multi_hash(arr, 0) == orc.multi_hash(arr, 0)


def test_lane_hash_matches_oracle():
    """the library's many-at-once multi_hash (AVX-512 IFMA lanes where the CPU has them, csrc/mimc7_lanes.cpp; the
    scalar chain otherwise) against the oracle, on ragged batches: empty messages, 0, p-1, 1..40 messages"""
    import random

    from gkr_b200.batch import multi_hash_many
    from gkr_b200.field import P
    rng = random.Random(77)
    assert multi_hash_many([[1, 2, 3], [12, 45, 78, 41]]) == [MH_123, MH_4]
    for count in (1, 2, 7, 8, 9, 15, 16, 17, 33, 40):
        msgs = [[rng.randrange(P) for _ in range(rng.randrange(0, 5))] for _ in range(count)]
        msgs[0] = [0, P - 1, 0][: 1 + count % 3]
        assert multi_hash_many(msgs) == [orc.multi_hash(m, 0) for m in msgs]
