import numpy as np, sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import vaporetto_b200 as vb
from vpt_testlib.oracle import OraclePredictor
d = "tests/golden/fuzz_cases/"
mb = np.load(d + "case1266_model.npy").tobytes()
text = np.load(d + "case1266_text.npy"); offs = np.load(d + "case1266_offs.npy")
p, o = vb.Predictor(vb.Model.read(mb)), OraclePredictor(mb)
print(p.info)
def run(t, of, label):
    r = p.predict_batch(t, of)
    sc, bd, boff, st = o.predict_batch(t, of, nthreads=4)
    print(label, "offsets equal", np.array_equal(r.bound_offsets, boff), "status equal", np.array_equal(r.status, st),
          "scores equal", np.array_equal(r.scores, sc))
    n = len(of) - 1
    bad = []
    for s in range(n):
        a, b = int(boff[s]), int(boff[s + 1])
        if not np.array_equal(r.scores[a:b], sc[a:b]):
            w = np.nonzero(r.scores[a:b] != sc[a:b])[0]
            bad.append((s, s // 64, b - a, w[:6].tolist(), (r.scores[a:b][w[:3]] - sc[a:b][w[:3]]).tolist()))
    print(label, "bad sentences", len(bad), bad[:12])
run(text, offs, "full")
# each group alone
n = len(offs) - 1
for g in range(0, n, 64):
    sub = offs[g:min(g + 64, n) + 1]
    t = text[int(sub[0]):int(sub[-1])]
    run(t, sub - sub[0], "group %d alone" % (g // 64))
