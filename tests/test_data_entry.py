"""Host-side data entry (ggad_b200.data: load_mat, preprocess_features, normalize_adj) against goldens produced by the
reference's own utils.py on synthetic .mat files (tests/golden/make_golden.py::data_case).  CPU only."""
import random

import numpy as np
import pytest
import scipy.io as sio
import scipy.sparse as sp

import oracle
from helpers import golden_cases, load_case


def _write_mat(c, root):
    alt = bool(int(c["i/alt_keys"]))
    mat = {("gnd" if alt else "Label"): c["i/label"], ("X" if alt else "Attributes"): sp.csc_matrix(c["i/attrs"]),
           ("A" if alt else "Network"): sp.csc_matrix(c["i/network"])}
    if "i/str" in c:
        mat["str_anomaly_label"], mat["attr_anomaly_label"] = c["i/str"], c["i/attr"]
    (root / "dataset").mkdir()
    sio.savemat(str(root / "dataset" / (str(c["i/dataset"]) + ".mat")), mat)


@pytest.mark.parametrize("name", golden_cases("data_"))
def test_load_mat_is_the_reference_split(name, tmp_path, monkeypatch, capsys):
    """Same .mat, same seed of Python's ``random`` -> the same 12-tuple as utils.py:66-141: adjacency, attributes, labels
    and every index list element for element (so the two shuffles consume the generator in the reference's order)."""
    from ggad_b200 import data
    c = load_case(name)
    _write_mat(c, tmp_path)
    monkeypatch.chdir(tmp_path)
    random.seed(int(c["i/seed"]))
    out = data.load_mat(str(c["i/dataset"]))
    assert len(out) == 12
    adj, feat, ano, all_idx, tr, va, te, ano2, str_a, attr_a, normal, abnormal = out
    o = c["out"]
    assert sp.isspmatrix_csr(adj) and sp.isspmatrix_lil(feat)
    assert np.array_equal(adj.toarray(), o["adj"]) and np.array_equal(feat.toarray(), o["feat"])
    assert np.array_equal(ano, o["ano_labels"]) and ano2 is ano
    for got, key in ((all_idx, "all_idx"), (tr, "idx_train"), (va, "idx_val"), (te, "idx_test"),
                     (normal, "normal_label_idx"), (abnormal, "abnormal_label_idx")):
        assert isinstance(got, list) and got == o[key].tolist(), key
    if "str" in o:
        assert np.array_equal(str_a, o["str"]) and np.array_equal(attr_a, o["attr"])
    else:
        assert str_a is None and attr_a is None
    printed = capsys.readouterr().out
    assert printed.startswith("Training Counter(") and "Training rate 0.5" in printed
    # the oracle's restatement of the split (what the GPU tests and smoke() use) agrees with both
    tr2, va2, te2, normal2, abnormal2 = oracle.load_mat_split(np.asarray(ano), str(c["i/dataset"]), int(c["i/seed"]))
    assert (tr2, va2, te2, normal2, abnormal2) == (tr, va, te, normal, abnormal)
    # outlier seeds: 5 % of the labelled normals on Amazon, 15 % elsewhere
    frac = 0.05 if str(c["i/dataset"]) == "Amazon" else 0.15
    assert len(abnormal) == int(len(normal) * frac)


@pytest.mark.parametrize("name", golden_cases("data_"))
def test_preprocess_features_and_normalize_adj_bit_exact(name):
    from ggad_b200 import data
    c = load_case(name)
    o = c["out"]
    dense, (coords, values, shape) = data.preprocess_features(sp.lil_matrix(o["feat"]))
    assert np.array_equal(np.asarray(dense), o["pre_dense"])                    # same numpy / scipy calls: same bits
    assert np.array_equal(coords, o["pre_coords"]) and np.array_equal(values, o["pre_values"])
    assert tuple(shape) == tuple(o["pre_shape"])
    assert np.all(np.asarray(dense)[5] == 0)                                    # the attribute-less node: inf -> 0
    a_hat = data.normalize_adj(sp.csr_matrix(o["adj"]))
    assert np.array_equal(a_hat.toarray(), o["normalize_adj"])


def test_load_mat_errors_are_loud(tmp_path, monkeypatch):
    from ggad_b200 import data
    monkeypatch.chdir(tmp_path)
    with pytest.raises(FileNotFoundError):
        data.load_mat("nope")
    (tmp_path / "dataset").mkdir()
    sio.savemat(str(tmp_path / "dataset" / "bad.mat"), {"Label": np.zeros((3, 1)), "Network": sp.eye(3, format="csc")})
    with pytest.raises(KeyError):
        data.load_mat("bad")


def test_split_does_not_touch_global_random_when_given_its_own_generator():
    from ggad_b200 import data
    labels = (np.arange(100) % 7 == 0).astype(int)
    random.seed(5)
    before = random.getstate()
    a = data.semi_supervised_split(labels, "photo", rng=random.Random(1))
    assert random.getstate() == before
    b = data.semi_supervised_split(labels, "photo", rng=random.Random(1))
    assert a == b and len(a[1]) == 30 and len(a[2]) == 10 and len(a[3]) == 60
    assert sorted(a[0]) == list(range(100)) and set(a[5]) <= set(a[4]) <= set(a[1])


@pytest.mark.parametrize("name", golden_cases("mbdata_"))
def test_minibatch_data_helpers_match_reference(name, tmp_path):
    """normalize / sparse_to_adj_lists / pos_neg_split against the reference's src/utils.py run on the same inputs:
    feature scaling bit for bit (dense and sparse input), the adjacency lists as the same dict of sets (and the same
    pickle round trip), the positive / negative lists element for element, duplicated id included."""
    import pickle

    from ggad_b200 import data, graph
    c = load_case(name)
    o = c["out"]
    n = int(c["i/n"])
    a = sp.csr_matrix((c["i/a_data"], c["i/a_indices"], c["i/a_indptr"]), shape=(n, n))
    x = c["i/x"]
    assert np.array_equal(np.asarray(data.normalize(x)), o["normalize_dense"])
    assert np.array_equal(data.normalize(sp.csr_matrix(x)).toarray(), o["normalize_sparse"])
    assert np.all(np.asarray(data.normalize(x))[4] == 0)                         # row sum -0.01: 1 / 0 -> 0
    fn = tmp_path / "adj_list"
    adj_lists = data.sparse_to_adj_lists(a, str(fn))
    keys = sorted(adj_lists)
    assert keys == o["adj_keys"].tolist()
    ref_sets, pos = {}, 0
    for k, ln in zip(o["adj_keys"].tolist(), o["adj_lens"].tolist()):
        ref_sets[k] = set(o["adj_flat"][pos:pos + ln].tolist())
        pos += ln
    assert all(adj_lists[k] == ref_sets[k] for k in keys)
    assert all(type(v) is set and all(type(t) is int for t in v) for v in adj_lists.values())
    with open(fn, "rb") as f:
        assert pickle.load(f) == adj_lists
    # and it is what the aggregators ingest: host CSR with sorted neighbour lists
    csr = graph.AdjListCSR(adj_lists, n)
    assert int(csr.rowptr[-1]) == sum(len(v) for v in adj_lists.values())
    p, ng = data.pos_neg_split(c["i/nodes"].tolist(), c["i/labels"].tolist())
    assert p == o["pos"].tolist() and ng == o["neg"].tolist()


def test_data_helpers_properties():
    """Size-independent properties of the host entry: the adjacency lists are symmetric and cover exactly the stored
    entries; the split partitions the node ids; feature scaling makes every non-empty row sum to 1 (program A) or to
    s / (s + 0.01) (program B)."""
    from hypothesis import given, settings, strategies as st

    from ggad_b200 import data

    @settings(max_examples=25, deadline=None)
    @given(st.integers(2, 60), st.integers(0, 2 ** 31 - 1))
    def check(n, seed):
        rng = np.random.default_rng(seed)
        a = sp.random(n, n, density=min(1.0, 3.0 / n), random_state=rng, data_rvs=lambda k: np.ones(k)).tocsr()
        adj = data.sparse_to_adj_lists(a)
        r, c = a.nonzero()
        pairs = set(zip(r.tolist(), c.tolist())) | set(zip(c.tolist(), r.tolist()))
        assert {(u, v) for u, vs in adj.items() for v in vs} == pairs
        assert all(u in adj[v] for u, vs in adj.items() for v in vs)
        labels = (rng.random(n) < 0.2).astype(int)
        all_idx, tr, va, te, normal, abnormal = data.semi_supervised_split(labels, "x", rng=random.Random(seed))
        assert sorted(tr + va + te) == list(range(n)) == sorted(all_idx)
        assert all(labels[i] == 0 for i in normal) and set(abnormal) <= set(normal) <= set(tr)
        x = rng.random((n, 5))
        x[0] = 0
        dense = np.asarray(data.preprocess_features(sp.csr_matrix(x))[0])
        assert np.allclose(dense[1:].sum(1), 1.0) and np.all(dense[0] == 0)
        s = x.sum(1)
        assert np.allclose(np.asarray(data.normalize(x)).sum(1), s / (s + 0.01))

    check()
