import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle'); sys.path.insert(0, 'tests')
import numpy as np, li_b200 as li, li_oracle as O
from util import hostile_positions
for T in (np.float32, np.float64):
    rng = np.random.default_rng(42)
    for S in (4, 6):
        for ncols in (5, 300):
            x = hostile_positions(ncols, T, rng)
            J, W = O.maketable(O.kernel(O.LANCZOS, S), x, ncols, T)
            A = li.TabulatedInterpolator(T, li.LanczosKernel(T, S), x, ncols)
            bad = np.argwhere(A.W != W)
            print(T.__name__, S, ncols, "mismatches", len(bad), "of", W.size)
            for s, i in bad[:8]:
                f = np.floor(x[i]); t = T(x[i] - f)
                print("   x=%r t=%r s=%d dev=%r ora=%r ulps=%d" % (x[i], t, s, A.W[s, i], W[s, i],
                      int(A.W[s,i].view(np.int32 if T==np.float32 else np.int64)) - int(W[s,i].view(np.int32 if T==np.float32 else np.int64))))
