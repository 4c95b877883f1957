import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'tests')); sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import numpy as np
import corpus
import test_gpu_synthetic as t
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
sc = t._random_placement_scene(seed, 160, 400, 300)
ref, info = corpus.render_oracle(sc, want_debug=True)
r, stages = corpus.make_product(sc)
r.render(stages[0])
edges, epath = r.debug_edges(0)
oe, op = info["edges"], info["edge_path"]
print("n gpu", len(edges), "n oracle", len(oe))
n = min(len(edges), len(oe))
bad = np.nonzero((edges[:n] != oe[:n]).any(axis=1))[0]
print("mismatching edges", len(bad), "first", bad[:10])
badp = np.nonzero(epath[:n] != op[:n])[0]
print("mismatching paths", len(badp), badp[:5])
if len(bad):
    i = bad[0]
    for k in range(max(0, i - 2), min(n, i + 6)):
        print(k, "path", epath[k], op[k], "gpu", edges[k].tolist(), "oracle", oe[k].tolist())
    # group by path: edges per path
    paths = np.unique(op[bad])
    print("paths with mismatches:", paths[:20], "of", len(np.unique(op)))
    for p in paths[:3]:
        idx = np.nonzero(op == p)[0]
        print("path", p, "edges", len(idx), "range", idx[0], idx[-1], "mismatching", np.isin(idx, bad).sum())
        b = idx[np.isin(idx, bad)]
        print("  first bad offsets within path:", (b[:10] - idx[0]).tolist())
        # segment structure: consecutive edges share endpoints; find segment breaks in oracle
print(r.stats())
