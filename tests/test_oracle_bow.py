"""CPU: the bag-of-words oracle (oracle/bow_oracle.c) against golden vectors made from the reference's own
DBoW2 (tests/golden/make_golden_bow.py) and, in the development container, against that DBoW2 live."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden

VARIANTS = [(0, 0), (1, 2), (5, 1), (5, 3), (2, 0)]
TREES = {"reg": (4, 2, 1, 0), "irr": (4, 2)}


def retitled(path, scoring, weighting, tmp_path):
    """The committed vocabulary with another 'scoring weighting' pair in its header line."""
    lines = open(path).read().split("\n")
    k, L = lines[0].split()[:2]
    lines[0] = "%s %s  %d %d" % (k, L, scoring, weighting)
    p = os.path.join(str(tmp_path), "v_%d_%d.txt" % (scoring, weighting))
    open(p, "w").write("\n".join(lines))
    return p


def check(res, g, prefix, exact_vals=True):
    ids, vals, nodes, feats = res
    assert np.array_equal(ids, g[prefix + "bow_ids"])
    assert vals.tobytes() == g[prefix + "bow_vals"].tobytes()   # doubles, bit for bit
    assert np.array_equal(nodes, g[prefix + "fv_nodes"])
    assert np.array_equal(np.array([len(f) for f in feats], np.int32), g[prefix + "fv_sizes"])
    flat = np.concatenate(feats) if feats else np.zeros(0, np.uint32)
    assert np.array_equal(flat, g[prefix + "fv_feats"])


@pytest.mark.parametrize("tree", sorted(TREES))
def test_oracle_bow_matches_reference_golden(oracle, tree, tmp_path):
    g = load_golden("bow_golden")
    q = g[tree + "_queries"]
    path = os.path.join(GOLDEN, "bow_vocab_%s.txt" % tree)
    for (sc, we) in VARIANTS:
        v = oracle.Vocabulary(path=retitled(path, sc, we, tmp_path))
        assert v.n_nodes == int(g[tree + "_n_nodes"]) and v.n_words == int(g[tree + "_n_words"])
        for lu in TREES[tree]:
            check(v.transform(q, lu), g, "%s_s%dw%d_lu%d_" % (tree, sc, we, lu))
    v = oracle.Vocabulary(path=path)
    assert np.array_equal(v.features(q)[0], g[tree + "_words"])
    check(v.transform(q[:0], 4), g, tree + "_empty_")
    check(v.transform(q[:1], 4), g, tree + "_one_")


def test_text_round_trip_and_synthetic_vocabulary(oracle, tmp_path):
    k, L = 4, 3
    parent, leaf, desc, weight = oracle.synth_vocab(k, L, seed=11)
    assert len(parent) == (k ** (L + 1) - 1) // (k - 1) and int(leaf.sum()) == k ** L
    p = os.path.join(str(tmp_path), "s.txt")
    oracle.save_vocab_text(p, k, L, parent, leaf, desc, weight)
    assert not open(p).read().endswith("\n")
    a = oracle.Vocabulary(k, L, parent, leaf, desc, weight)
    b = oracle.Vocabulary(path=p)
    ka, La, sa, wa, pa, la, da, wta = a.arrays()
    kb, Lb, sb, wb, pb, lb, db, wtb = b.arrays()
    assert (ka, La, sa, wa) == (kb, Lb, sb, wb) == (k, L, 0, 0)
    assert np.array_equal(pa[1:], pb[1:]) and np.array_equal(la[1:], lb[1:]) and np.array_equal(da[1:], db[1:])
    assert wta.tobytes() == wtb.tobytes()
    q = oracle.synth_descriptors(0, 300, 5)
    ra, rb = a.transform(q, 2), b.transform(q, 2)
    assert all(np.array_equal(x, y) for x, y in zip(ra[:3], rb[:3]))
    # a trailing newline must not change what is loaded (the reference would grow a phantom node)
    open(p, "a").write("\n")
    c = oracle.Vocabulary(path=p)
    assert c.n_nodes == a.n_nodes
    # the L1-normalised vector sums to one, word ids increase, every unstopped feature is in one node list
    ids, vals, nodes, feats = ra
    assert abs(vals.sum() - 1.0) < 1e-12 and np.all(np.diff(ids.astype(np.int64)) > 0)
    wid, w, nid = a.features(q, 2)
    assert sorted(np.concatenate(feats).tolist()) == np.nonzero(w > 0)[0].tolist()
    for node, fl in zip(nodes, feats):
        assert np.all(nid[fl] == node) and np.all(np.diff(fl.astype(np.int64)) > 0)


def test_oracle_bow_matches_reference_live(oracle, tmp_path):
    from oracle import ref as R
    if not R.bow_available():
        pytest.skip("oracle/_ref/libbow_ref.so needs /root/reference (development container)")
    rng = np.random.default_rng(5)
    for (k, L, seed) in [(3, 2, 1), (10, 3, 2), (7, 4, 3)]:
        parent, leaf, desc, weight = oracle.synth_vocab(k, L, seed=seed)
        p = os.path.join(str(tmp_path), "v%d.txt" % seed)
        oracle.save_vocab_text(p, k, L, parent, leaf, desc, weight)
        vo, vr = oracle.Vocabulary(path=p), R.Vocabulary(p)
        leaves = np.nonzero(leaf)[0]
        q = desc[rng.choice(leaves, 600)].copy()
        for i in range(len(q)):
            for b in rng.integers(0, 256, int(rng.integers(0, 20))):
                q[i, b >> 3] ^= np.uint8(1 << (b & 7))
        q = np.concatenate([q, oracle.synth_descriptors(0, 200, seed)])
        for lu in (L + 1, L, L - 1, 1, 0):
            a, b = vo.transform(q, lu), vr.transform(q, lu)
            assert np.array_equal(a[0], b[0]) and a[1].tobytes() == b[1].tobytes()
            assert np.array_equal(a[2], b[2]) and len(a[3]) == len(b[3])
            assert all(np.array_equal(x, y) for x, y in zip(a[3], b[3]))
        assert np.array_equal(vo.features(q)[0], vr.words(q))
