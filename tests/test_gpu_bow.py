"""GPU parity of the bag-of-words transform (orbv_* of the C ABI) against golden vectors made from the
reference's own DBoW2 and against the CPU oracle; full-size properties on a k=10, L=6 vocabulary."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from test_oracle_bow import TREES, VARIANTS, check, retitled

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    import orb_slam_fusion_b200 as P
    return P


def same(a, b):
    return (np.array_equal(a[0], b[0]) and a[1].tobytes() == b[1].tobytes() and np.array_equal(a[2], b[2]) and
            len(a[3]) == len(b[3]) and all(np.array_equal(x, y) for x, y in zip(a[3], b[3])))


@pytest.mark.parametrize("tree", sorted(TREES))
def test_bow_matches_reference_golden(P, tree, tmp_path):
    g = load_golden("bow_golden")
    q = g[tree + "_queries"]
    path = os.path.join(GOLDEN, "bow_vocab_%s.txt" % tree)
    for (sc, we) in VARIANTS:
        v = P.ORBVocabulary.loadFromTextFile(retitled(path, sc, we, tmp_path))
        assert (v.n_nodes, v.n_words, v.scoring, v.weighting) == (int(g[tree + "_n_nodes"]), int(g[tree + "_n_words"]), sc, we)
        for lu in TREES[tree]:
            check(v.transform(q, lu), g, "%s_s%dw%d_lu%d_" % (tree, sc, we, lu))
    v = P.ORBVocabulary(path=path)
    assert np.array_equal(v.features(q)[0], g[tree + "_words"])
    check(v.transform(q[:0], 4), g, tree + "_empty_")
    check(v.transform(q[:1], 4), g, tree + "_one_")
    assert v.size() == int(g[tree + "_n_words"]) and not v.empty()


def test_bow_vs_oracle_shapes_and_edge_cases(P, oracle, tmp_path):
    rng = np.random.default_rng(3)
    for (k, L, seed) in [(2, 1, 1), (3, 4, 2), (10, 3, 3), (20, 2, 4)]:
        parent, leaf, desc, weight = oracle.synth_vocab(k, L, seed=seed)
        vo = oracle.Vocabulary(k, L, parent, leaf, desc, weight)
        vg = P.ORBVocabulary(k, L, parent, leaf, desc, weight)
        leaves = np.nonzero(leaf)[0]
        q = desc[rng.choice(leaves, 700)].copy()
        for i in range(len(q)):
            for b in rng.integers(0, 256, int(rng.integers(0, 24))):
                q[i, b >> 3] ^= np.uint8(1 << (b & 7))
        q = np.concatenate([q, oracle.synth_descriptors(0, 333, seed), q[:50]])   # duplicates: repeated words
        for lu in (L + 2, L, 1, 0):
            assert same(vg.transform(q, lu), vo.transform(q, lu))
        a, b = vg.features(q, 1), vo.features(q, 1)
        assert np.array_equal(a[0], b[0]) and a[1].tobytes() == b[1].tobytes() and np.array_equal(a[2], b[2])
    # ties: identical children -> the first (lowest node id) wins; all weights zero -> empty vectors
    parent = np.array([0, 0, 0, 0], np.int32)
    leaf = np.array([0, 1, 1, 1], np.uint8)
    desc = np.zeros((4, 32), np.uint8)
    vt = P.ORBVocabulary(3, 1, parent, leaf, desc, np.array([0, 2.0, 3.0, 4.0]))
    ids, vals, nodes, feats = vt.transform(np.full((5, 32), 255, np.uint8), 0)
    assert ids.tolist() == [0] and vals.tolist() == [1.0] and nodes.tolist() == [1] and feats[0].tolist() == [0, 1, 2, 3, 4]
    vz = P.ORBVocabulary(3, 1, parent, leaf, desc, np.zeros(4))
    ids, vals, nodes, feats = vz.transform(np.zeros((5, 32), np.uint8), 0)
    assert len(ids) == 0 and len(nodes) == 0
    # a vocabulary without words: transform() returns empty vectors (TemplatedVocabulary.h:1063)
    ve = P.ORBVocabulary(10, 5, np.zeros(1, np.int32), np.zeros(1, np.uint8), np.zeros((1, 32), np.uint8), np.zeros(1))
    assert ve.empty() and len(ve.transform(q[:10], 4)[0]) == 0
    # errors: unreadable file, bad header, parent out of range, too many features per frame
    with pytest.raises(P.OrbxError):
        P.ORBVocabulary(path=os.path.join(str(tmp_path), "missing.txt"))
    bad = os.path.join(str(tmp_path), "bad.txt")
    open(bad, "w").write("30 6  0 0\n0 1 " + "0 " * 32 + "1.0")
    with pytest.raises(P.OrbxError):
        P.ORBVocabulary(path=bad)
    with pytest.raises(P.OrbxError):
        P.ORBVocabulary(3, 1, np.array([0, 7], np.int32), np.ones(2, np.uint8), np.zeros((2, 32), np.uint8), np.ones(2))
    with pytest.raises(P.OrbxError) as e:
        vt.transform_batch(np.zeros((1, 20000, 32), np.uint8))
    assert e.value.code == -6


def test_bow_full_size_batch_on_device(P, oracle):
    """ORBvoc-shaped tree (k=10, L=6: 1 111 111 nodes, 35 MB of descriptors) over a batch of extracted
    frames, device-resident end to end: extractor outputs feed orbv_transform without a host round trip."""
    import torch
    k, L = 10, 6
    parent, leaf, desc, weight = oracle.synth_vocab(k, L, seed=7)
    assert len(parent) == 1111111
    vg = P.ORBVocabulary(k, L, parent, leaf, desc, weight)
    vo = oracle.Vocabulary(k, L, parent, leaf, desc, weight)
    F = 48
    frames = P.synth_frames("blocks", F, 752, 480, seed=1)
    ex = P.OrbExtractor(1000, 1.2, 8, 20, 7, max_batch=F)
    n, nm, kps, d = ex.extract_batch(frames)
    r = vg.transform_batch(d, n, 4)
    r2 = vg.transform_batch(d, n, 4)
    torch.cuda.synchronize()
    n = n.cpu().numpy()
    dn = d.cpu().numpy()
    R = {key: val.cpu().numpy() for key, val in r.items()}
    for key in r:
        if key.endswith("_n") or key == "fv_total":
            assert torch.equal(r[key], r2[key])
    for f in range(F):
        nb, nf, tot = int(R["bow_n"][f]), int(R["fv_n"][f]), int(R["fv_total"][f])
        assert 0 < nb <= tot <= n[f] and 0 < nf <= tot
        ids, vals = R["bow_ids"][f, :nb].astype(np.int64), R["bow_vals"][f, :nb]
        assert np.all(np.diff(ids) > 0) and abs(vals.sum() - 1.0) < 1e-9 and np.all(vals > 0)
        assert torch.equal(r["bow_vals"][f, :nb], r2["bow_vals"][f, :nb])       # deterministic doubles
        feats = R["fv_feats"][f, :tot].astype(np.int64)
        assert len(np.unique(feats)) == tot and feats.max() < n[f]
    for f in (0, 17, F - 1):   # spot frames against the oracle, bit for bit
        want = vo.transform(dn[f, :n[f]], 4)
        nb, nf, tot = int(R["bow_n"][f]), int(R["fv_n"][f]), int(R["fv_total"][f])
        begin = R["fv_begin"][f]
        ends = list(begin[1:nf]) + [tot]
        got = (R["bow_ids"][f, :nb].astype(np.uint32), R["bow_vals"][f, :nb], R["fv_nodes"][f, :nf].astype(np.uint32),
               [R["fv_feats"][f, begin[j]:ends[j]].astype(np.uint32) for j in range(nf)])
        assert same(got, want)
