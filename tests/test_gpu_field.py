"""GPU: the PTX field library (gkr_b200/csrc/fr.cuh) against Python big integers, through the C ABI."""
import random

import pytest

from gkr_b200.field import P, fr_to_ints, ints_to_fr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pv():
    from gkr_b200 import Prover
    return Prover(0)


def _edge():
    vals = [0, 1, 2, P - 1, P - 2, (P - 1) // 2, (P + 1) // 2, 2**32 - 1, 2**32, 2**64 - 1, 2**128, 2**253, 2**253 + 12345]
    vals += [(1 << (32 * i)) - 1 for i in range(1, 8)] + [P - (1 << (32 * i)) for i in range(1, 8)]
    return [v % P for v in vals]


def test_add_sub_mul(pv):
    rng = random.Random(1)
    e = _edge()
    a = [x for x in e for _ in e] + [rng.randrange(P) for _ in range(20000)]
    b = [y for _ in e for y in e] + [rng.randrange(P) for _ in range(20000)]
    A, B = ints_to_fr(a), ints_to_fr(b)
    for op, fn in ((0, lambda x, y: (x + y) % P), (1, lambda x, y: (x - y) % P), (2, lambda x, y: x * y % P)):
        got = fr_to_ints(pv.fr_binop(op, A, B))
        assert got == [fn(x, y) for x, y in zip(a, b)], f"op {op}"


def test_range_check(pv):
    from gkr_b200._lib import GkrError
    with pytest.raises(GkrError) as ei:
        pv.fr_binop(0, ints_to_fr([P]), ints_to_fr([1]))
    assert ei.value.code == -4
    # the context stays usable
    assert fr_to_ints(pv.fr_binop(0, ints_to_fr([P - 1]), ints_to_fr([2]))) == [1]
