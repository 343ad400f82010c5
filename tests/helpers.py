"""Shared helpers for the parity tests (load golden cases, tolerances)."""
import glob
import os

import numpy as np
import scipy.sparse as sp
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star tolerance: embeddings / affinities within 1e-4 relative in fp32.
RTOL = 1e-4
ATOL = 2e-5   # absolute floor for entries that are ~0 after cancellation


def golden_cases(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["params"] = {k[2:]: torch.from_numpy(z[k].copy()) for k in z.files if k.startswith("p/")}
    d["grads"] = {k[2:]: torch.from_numpy(z[k].copy()) for k in z.files if k.startswith("g/")}
    d["out"] = {k[2:]: z[k] for k in z.files if k.startswith("o/")}
    return d


def case_adjacency(c):
    n = int(c["n"])
    return sp.csr_matrix((c["a_data"], c["a_indices"], c["a_indptr"]), shape=(n, n))


def case_adj_lists(c):
    rp, col = c["adj_rowptr"], c["adj_col"]
    return {i: set(int(t) for t in col[rp[i]:rp[i + 1]]) for i in range(int(c["n"]))}


def assert_close(a, b, rtol=RTOL, atol=ATOL, what=""):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, err_msg=what, equal_nan=True)
