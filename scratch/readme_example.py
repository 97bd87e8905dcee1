import os,sys
sys.path.insert(0,'/root/repo')
import numpy as np, li_b200 as li          # test/benchmarks.jl:23-49, with device arrays
T = np.float32
ker = li.CatmullRomSpline(T)
A1 = li.SparseUnidimensionalInterpolator(ker, 1, np.arange(1, 301), li.Range(1, 300, length=43))
A2 = li.SparseUnidimensionalInterpolator(ker, 2, np.arange(1, 201), li.Range(1, 200, length=29))
H = A1 * A2                                 # one fused kernel per apply, forward and adjoint
x = li.to_device(np.asfortranarray(np.random.randn(43, 29).astype(T)))
y = H * x                                   # vcreate + apply!(1, Direct, H, x, false, 0, y)
g = H.adj * y                               # deterministic gather-form adjoint
li.apply_b(0.5, li.Adjoint, H, y, False, 2, g)          # apply!(α, P, H, x, scratch, β, y)
xfit = li.conjgrad(H.adj * H, H.adj * y, maxiter=50)    # test/benchmarks.jl:40, the loop runs inside the library
xfit = li.fit_separable([A1.table, A2.table], y, maxiter=50)   # the same solve in ONE library call (li_conjgrad_sep)

import numpy as np
print('ok', li.to_host(xfit).shape, float(np.abs(li.to_host(xfit)-li.to_host(x)).max()))
