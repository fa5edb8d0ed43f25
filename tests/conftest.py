"""Shared test helpers.  GPU tests are marked ``@pytest.mark.gpu`` and call through the C-ABI;
everything else runs on CPU (oracle vs golden fixtures, host logic, ABI surface)."""
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")

FIXTURES = ("simpleOFE3", "AMOREBracket", "patch16", "patch16_m3", "patch16_tri", "patch16_m3s", "single_lower", "single_icm", "single_quad", "zoo_s1", "zoo_s2", "zoo_s3")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    from mofem_b200.flatten import FlatProblem
    d = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    return FlatProblem.from_npz_dict(d), d


def row_scaled_error(K, Kref):
    """P2 metric of SURVEY.md section 8: max_ij |K_ij - Kref_ij| / max_j |Kref_ij| (sparse matrices, same shape)."""
    import scipy.sparse as sp
    K = sp.csr_matrix(K); Kref = sp.csr_matrix(Kref)
    rm = abs(Kref).max(axis=1).toarray().ravel()
    rm[rm == 0] = 1.0
    d = abs(K - Kref).max(axis=1).toarray().ravel()
    return float((d / rm).max()) if d.size else 0.0


def coo_to_csr(n, I, J, V):
    import scipy.sparse as sp
    return sp.coo_matrix((V, (I, J)), shape=(n, n)).tocsr()


def fixed_from_constraints(constraints, n_dof):
    """fix mask / values exactly as AMORE.py:213-226 builds fixIndexList / fixList."""
    fixed = np.zeros(n_dof, np.uint8)
    val = np.zeros(n_dof)
    for c in constraints:
        nd = int(c[0])
        if c[1]:
            fixed[2 * nd] = 1; val[2 * nd] = c[3]
        if c[2]:
            fixed[2 * nd + 1] = 1; val[2 * nd + 1] = c[4]
    return fixed, val


@pytest.fixture(scope="session")
def gpu_ctx():
    from mofem_b200.device import Context
    ctx = Context(0)
    yield ctx
    ctx.close()
