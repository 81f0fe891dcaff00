"""Development aid: lr_grad_hess against NumPy for a few shapes, printing where the two differ."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nums_b200 import cuda_compute
from nums_b200.cuda_system import CudaSystem

system = CudaSystem(); system.init()
for n, d in [(256, 28), (1000, 28), (100003, 28)]:
    rng = np.random.default_rng(71)
    X = rng.standard_normal((n, d))
    y = (rng.random(n) < 0.5).astype(np.float64)
    beta = rng.standard_normal(d) / np.sqrt(d)
    mu = 1.0 / (1.0 + np.exp(-(X @ beta)))
    g = X.T @ (mu - y)
    H = X.T @ ((mu * (1 - mu))[:, None] * X)
    try:
        out = system.get(cuda_compute.lr_grad_hess(system.put(X), system.put(y), system.put(beta)))
    except Exception as exc:
        print(n, d, "EXC", type(exc).__name__, exc)
        continue
    Hg = out[d:].reshape(d, d)
    eg = np.abs(out[:d] - g) / np.abs(g).max()
    eh = np.abs(Hg - H) / np.abs(H).max()
    print(n, d, "g err max %.3e  H err max %.3e  symmetric %s" % (eg.max(), eh.max(), np.array_equal(Hg, Hg.T)))
    if eg.max() > 1e-10:
        print(" bad g idx", np.nonzero(eg > 1e-10)[0][:28], "\n got", out[:6], "\n want", g[:6])
    if eh.max() > 1e-10:
        bad = np.argwhere(eh > 1e-10)
        print(" bad H entries", len(bad), "first", bad[:10].tolist())
        blocks = sorted({(int(r) // 4, int(c) // 4) for r, c in bad})
        print(" bad 4x4 blocks", blocks)
        print(" ratio sample", (Hg / H)[tuple(bad[0])])
