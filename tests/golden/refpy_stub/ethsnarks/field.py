"""Prime field element over the BN254 scalar field with the operations python/poly.py, sumcheck.py and gkr.py use
(`ethsnarks.field.FQ`: construction from any int, + - * ** ==, int(), repr(), hashing, zero/one/random)."""
SNARK_SCALAR_FIELD = 21888242871839275222246405745257275088548364400416034343698204186575808495617


class FQ(object):
    __slots__ = ("n", "m")

    def __init__(self, n, field_modulus=SNARK_SCALAR_FIELD):
        self.m = field_modulus
        self.n = (n.n if isinstance(n, FQ) else int(n)) % self.m

    def _other(self, other):
        return other.n if isinstance(other, FQ) else int(other) % self.m

    def __add__(self, other):
        return FQ(self.n + self._other(other), self.m)

    __radd__ = __add__

    def __sub__(self, other):
        return FQ(self.n - self._other(other), self.m)

    def __rsub__(self, other):
        return FQ(self._other(other) - self.n, self.m)

    def __mul__(self, other):
        if not isinstance(other, (FQ, int)):
            return NotImplemented          # lets `FQ * term` style products fall through to the other operand
        return FQ(self.n * self._other(other), self.m)

    __rmul__ = __mul__

    def __neg__(self):
        return FQ(-self.n, self.m)

    def __pow__(self, e):
        return FQ(pow(self.n, self._other(e), self.m), self.m)

    def inv(self):
        return FQ(pow(self.n, self.m - 2, self.m), self.m)

    def __truediv__(self, other):
        return self * FQ(self._other(other), self.m).inv()

    def __eq__(self, other):
        if other == 0.0 and not isinstance(other, FQ):
            return self.n == 0
        return isinstance(other, (FQ, int)) and self.n == self._other(other)

    def __ne__(self, other):
        return not self == other

    def __hash__(self):
        return hash(self.n)

    def __int__(self):
        return self.n

    def __repr__(self):
        return repr(self.n)

    @classmethod
    def zero(cls, modulus=SNARK_SCALAR_FIELD):
        return cls(0, modulus)

    @classmethod
    def one(cls, modulus=SNARK_SCALAR_FIELD):
        return cls(1, modulus)

    @classmethod
    def random(cls, modulus=SNARK_SCALAR_FIELD):
        return cls(0, modulus)             # see the package docstring: z_0 = 0 as in the Rust prover
