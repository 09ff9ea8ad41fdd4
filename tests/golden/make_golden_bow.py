#!/usr/bin/env python3
"""Generate the bag-of-words golden vectors under tests/golden/ (run in the development container).

The outputs come from the REFERENCE's own vendored DBoW2 (3rdparty/DBoW2: TemplatedVocabulary.h
loadFromTextFile + transform, FORB.cpp, BowVector.cpp, FeatureVector.cpp, ScoringObject.cpp), compiled
unmodified on the mini-cv shim into oracle/_ref/libbow_ref.so -- not from the C oracle.  The vocabularies
and queries are made here with numpy (seeded), written in the reference's text format:

    python tests/golden/make_golden_bow.py     # rewrites tests/golden/bow_*.txt and bow_golden.npz
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref as R  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
# (scoring, weighting) pairs of DBoW2/BowVector.h:31-44 exercised besides ORBvoc's (L1_NORM, TF_IDF)
VARIANTS = [(0, 0), (1, 2), (5, 1), (5, 3), (2, 0)]


def flip(d, bits):
    d = d.copy()
    for b in bits:
        d[b >> 3] ^= np.uint8(1 << (b & 7))
    return d


def make_tree(rng, k, L, irregular):
    """Nodes in file order (parents before children).  Returns parent, leaf, desc, weight (index 0 = root)."""
    parent, leaf, desc, weight, depth = [0], [0], [np.zeros(32, np.uint8)], [0.0], [0]
    frontier = [0]
    for lev in range(1, L + 1):
        nxt = []
        for p in frontier:
            nchild = k if not irregular else int(rng.integers(2, k + 1))
            for _ in range(nchild):
                if lev == 1:
                    d = rng.integers(0, 256, 32, dtype=np.uint8)
                else:
                    d = flip(desc[p], rng.integers(0, 256, 256 >> lev))
                is_leaf = lev == L or (irregular and lev >= 2 and rng.random() < 0.3)
                parent.append(p); leaf.append(int(is_leaf)); desc.append(d); depth.append(lev)
                # idf-like weights with a few digits, some stopped words (weight 0)
                weight.append((0.0 if rng.random() < 0.05 else float(np.round(rng.uniform(0.2, 9.0), 5))) if is_leaf else 0.0)
                if not is_leaf:
                    nxt.append(len(parent) - 1)
        frontier = nxt
    return (np.array(parent, np.int32), np.array(leaf, np.uint8), np.stack(desc), np.array(weight, np.float64))


def write_text(path, k, L, scoring, weighting, parent, leaf, desc, weight):
    """saveToTextFile's format (TemplatedVocabulary.h:1335-1356) without the final newline: the reference's
    loader appends a phantom root child with an unset descriptor for an empty last line (:1288-1296)."""
    lines = ["%d %d  %d %d" % (k, L, scoring, weighting)]
    for i in range(1, len(parent)):
        lines.append("%d %d %s %s" % (parent[i], leaf[i], " ".join(str(int(b)) for b in desc[i]), repr(float(weight[i]))))
    with open(path, "w") as f:
        f.write("\n".join(lines))


def queries(rng, leaf, desc, n_near, n_rand):
    leaves = np.nonzero(leaf)[0]
    q = [flip(desc[rng.choice(leaves)], rng.integers(0, 256, int(rng.integers(0, 13)))) for _ in range(n_near)]
    q += [rng.integers(0, 256, 32, dtype=np.uint8) for _ in range(n_rand)]
    q = np.stack(q)
    return q[rng.permutation(len(q))]


def pack(prefix, res, out):
    ids, vals, nodes, feats = res
    out[prefix + "bow_ids"] = ids
    out[prefix + "bow_vals"] = vals
    out[prefix + "fv_nodes"] = nodes
    out[prefix + "fv_sizes"] = np.array([len(f) for f in feats], np.int32)
    out[prefix + "fv_feats"] = np.concatenate(feats) if feats else np.zeros(0, np.uint32)


def main():
    rng = np.random.default_rng(20261018)
    out = {}
    trees = {"reg": (5, 3, False), "irr": (6, 3, True)}
    for name, (k, L, irregular) in trees.items():
        parent, leaf, desc, weight = make_tree(rng, k, L, irregular)
        path = os.path.join(HERE, "bow_vocab_%s.txt" % name)
        write_text(path, k, L, 0, 0, parent, leaf, desc, weight)
        q = queries(rng, leaf, desc, 420, 80)
        out[name + "_queries"] = q
        out[name + "_n_nodes"] = np.int32(len(parent))
        out[name + "_n_words"] = np.int32(int(leaf.sum()))
        for (sc, we) in VARIANTS:
            with tempfile.TemporaryDirectory() as d:
                p2 = os.path.join(d, "v.txt")
                write_text(p2, k, L, sc, we, parent, leaf, desc, weight)
                v = R.Vocabulary(p2)
                assert v.n_words == int(leaf.sum())
                # irregular tree: levels where every path is long enough for the node id to be set (>= depth 2 leaves)
                for lu in ((4, 2, 1, 0) if not irregular else (4, 2)):
                    pack("%s_s%dw%d_lu%d_" % (name, sc, we, lu), v.transform(q, lu), out)
                if (sc, we) == (0, 0):
                    out[name + "_words"] = v.words(q)
                    pack(name + "_empty_", v.transform(q[:0], 4), out)
                    pack(name + "_one_", v.transform(q[:1], 4), out)
    np.savez_compressed(os.path.join(HERE, "bow_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
