"""The slice of petsc4py's Mat / Vec / IS API that thermalporous/preconditioners.py uses in its decoupling algebra
(create_decoup_* : preconditioners.py:680-873, 1442-1543), on scipy.sparse / numpy, one process.

Used only by tests/golden/make_pc_golden.py, which hands the reference's own create_decoup_* methods the
sub-blocks of a golden Jacobian and records what they produce.  Nothing here solves anything."""
import numpy as np
import scipy.sparse as sp


class _Vec:
    def __init__(self, a=None):
        self.a = None if a is None else np.array(a, dtype=float)

    def create(self, comm=None):
        return self

    def setSizes(self, size, bsize=None):
        n = size[1] if isinstance(size, (tuple, list)) else int(size)
        self.a = np.zeros(int(n))

    def setUp(self):
        pass

    def setValue(self, i, v, addv=None):
        self.a[int(i)] = v

    def getValue(self, i):
        return float(self.a[int(i)])

    def assemblyBegin(self):
        pass

    def assemblyEnd(self):
        pass

    def reciprocal(self):
        self.a = 1.0 / self.a

    def getSize(self):
        return self.a.size

    def __getitem__(self, i):
        return self.a[i]

    @property
    def array(self):
        return self.a


class _Mat:
    def __init__(self, m=None):
        self.m = None if m is None else sp.csr_matrix(m)
        self._bs = (1, 1)

    # ---- construction
    def create(self, comm=None):
        return self

    def setSizes(self, sizes, bsize=None):
        (_, M), (_, N) = sizes
        self.m = sp.lil_matrix((int(M), int(N)))

    def setBlockSizes(self, r, c):
        self._bs = (r, c)

    def getBlockSizes(self):
        return self._bs

    def setUp(self):
        pass

    def setDiagonal(self, v, addv=None):
        m = sp.lil_matrix(self.m)
        m.setdiag(v.a)
        self.m = m

    def setValue(self, i, j, v, addv=None):
        self.m[int(i), int(j)] = v

    def assemblyBegin(self):
        pass

    def assemblyEnd(self):
        self.m = sp.csr_matrix(self.m)

    # ---- queries
    def getSizes(self):
        M, N = self.m.shape
        return ((M, M), (N, N))

    def getSize(self):
        return self.m.shape

    def getOwnershipRange(self):
        return (0, self.m.shape[0])

    def getDiagonal(self):
        return _Vec(sp.csr_matrix(self.m).diagonal())

    def getRowSum(self):
        # PETSc MatGetRowSum: plain sums of the rows' entries
        return _Vec(np.asarray(sp.csr_matrix(self.m).sum(axis=1)).ravel())

    def getValue(self, i, j):
        return float(sp.csr_matrix(self.m)[int(i), int(j)])

    # ---- algebra
    def transpose(self, out=None):
        t = sp.csr_matrix(self.m).T.tocsr()
        if out is None:
            return _Mat(t)
        out.m = t
        return out

    def matMult(self, other):
        return _Mat(sp.csr_matrix(self.m) @ sp.csr_matrix(other.m))

    def mult(self, x, y):
        y.a = sp.csr_matrix(self.m) @ x.a

    def multTranspose(self, x, y):
        y.a = sp.csr_matrix(self.m).T @ x.a

    def axpy(self, alpha, X, structure=None):
        self.m = sp.csr_matrix(self.m) + alpha * sp.csr_matrix(X.m)

    def __sub__(self, other):
        return _Mat(sp.csr_matrix(self.m) - sp.csr_matrix(other.m))

    def copy(self):
        return _Mat(sp.csr_matrix(self.m).copy())


class _IS:
    def __init__(self, idx=None):
        self.indices = None if idx is None else np.asarray(idx, dtype=np.int64)

    def createGeneral(self, idx, comm=None):
        self.indices = np.asarray(idx, dtype=np.int64)
        return self

    @property
    def array(self):
        return self.indices

    def getLocalSize(self):
        return int(self.indices.size)

    def getSize(self):
        return int(self.indices.size)


class _Options:
    def getString(self, name, default=None):
        return default


class PETSc:
    Mat = _Mat
    Vec = _Vec
    IS = _IS
    Options = _Options
