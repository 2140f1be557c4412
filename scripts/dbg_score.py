import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "df-vo_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from b200 import native
from oracle import cvreplay, synth
from util import dptr
lib = native.load()
rs = np.random.RandomState(6)
kp_ref, kp_cur, _ = synth.correspondences(seed=45, n=2048, outlier_frac=0.3)
cx, cy, fx, fy = synth.kitti_intrinsics()
x1 = np.ascontiguousarray((kp_cur - np.array([cx, cy])) / fx); x2 = np.ascontiguousarray((kp_ref - np.array([cx, cy])) / fx)
M = 10000
E = rs.standard_normal((M, 9)); E /= np.linalg.norm(E, axis=1, keepdims=True)
thr2 = (0.2 / fx) ** 2 * 400
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
counts = torch.zeros(M, dtype=torch.int32, device="cuda")
lib.check(lib.dfvo_score_hypotheses(dptr(cu(E)), M, dptr(cu(x1)), dptr(cu(x2)), 2048, thr2, dptr(counts), None))
torch.cuda.synchronize()
got = counts.cpu().numpy()
def sampson_plain(e, x1, x2):
    u1, v1, u2, v2 = x1[:, 0], x1[:, 1], x2[:, 0], x2[:, 1]
    a0 = (e[0] * u1 + e[1] * v1) + e[2]; a1 = (e[3] * u1 + e[4] * v1) + e[5]; a2 = (e[6] * u1 + e[7] * v1) + e[8]
    b0 = (e[0] * u2 + e[3] * v2) + e[6]; b1 = (e[1] * u2 + e[4] * v2) + e[7]
    m = (u2 * a0 + v2 * a1) + a2
    return m * m / (((a0 * a0 + a1 * a1) + b0 * b0) + b1 * b1)
want_np = np.array([int((cvreplay.sampson_errors(E[i].reshape(3, 3), x1, x2) <= thr2).sum()) for i in range(M)])
want_pl = np.array([int((sampson_plain(E[i], x1, x2) <= thr2).sum()) for i in range(M)])
print("gpu vs numpy-blas: n diff", int((got != want_np).sum()), "max", int(np.abs(got - want_np).max()))
print("gpu vs plain-order: n diff", int((got != want_pl).sum()), "max", int(np.abs(got - want_pl).max()))
print("numpy-blas vs plain: n diff", int((want_np != want_pl).sum()))
i = int(np.argmax(np.abs(got - want_pl)))
print("worst model", i, got[i], want_pl[i], want_np[i], "errs near thr:", np.sort(np.abs(sampson_plain(E[i], x1, x2) / thr2 - 1))[:4])
