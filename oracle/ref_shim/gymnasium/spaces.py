"""Stand-in spaces (oracle shim): only n / nvec / sample() / contains() are provided."""
import numpy


class Discrete:
    def __init__(self, n, start=0):
        self.n = int(n)
        self.start = int(start)
        self.shape = ()
        self.dtype = numpy.int64

    def sample(self):
        return int(numpy.random.randint(self.start, self.start + self.n))

    def contains(self, x):
        return self.start <= int(x) < self.start + self.n

    def __repr__(self):
        return f"Discrete({self.n})"


class MultiDiscrete:
    def __init__(self, nvec):
        self.nvec = numpy.asarray(nvec, dtype=numpy.int64)
        self.shape = self.nvec.shape
        self.dtype = numpy.int64

    def sample(self):
        return numpy.array([numpy.random.randint(0, n) for n in self.nvec], dtype=numpy.int64)

    def contains(self, x):
        x = numpy.asarray(x)
        return x.shape == self.nvec.shape and bool(((x >= 0) & (x < self.nvec)).all())

    def __repr__(self):
        return f"MultiDiscrete({self.nvec.tolist()})"
