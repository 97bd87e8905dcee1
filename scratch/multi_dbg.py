import sys, os, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1:
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
    import numpy as np
    import li_b200 as li
    import li_oracle as O
    from test_gpu_multi import _frames, _oracle_sum
    from util import assert_within_bound
    T = np.float32
    S, m1, m2, nf, sh, seed = [float(v) if "." in v else int(v) for v in sys.argv[1:7]]
    tabs, otabs, ys = _frames(li, T, "LanczosKernel", S, O.FLAT, m1, m2, m1, m2, nf, 1.0, sh, seed=seed)
    model0 = np.zeros((m1, m2), T, order="F")
    dys = [li.to_device(y) for y in ys]
    tl = [(a.table, b.table) for a, b in tabs]
    m = li.to_device(model0)
    li.adjoint_accumulate(tl, dys, m, alpha=1, beta=0)
    got = li.to_host(m)
    ref, bound = _oracle_sum(otabs, ys, m1, m2, 1, 0, model0, T)
    assert_within_bound(got, ref, bound, T, "dbg")
    print("OK")
else:
    for args in (["4", "1100", "64", "3", "6.0", "403"], ["4", "1100", "64", "3", "0.4", "403"], ["4", "1024", "64", "3", "6.0", "403"],
                 ["4", "1100", "64", "3", "6.0", "404"], ["6", "1100", "64", "3", "6.0", "403"], ["4", "600", "64", "3", "6.0", "403"]):
        r = subprocess.run([sys.executable, __file__] + args, capture_output=True, text=True, env=dict(os.environ, CUDA_LAUNCH_BLOCKING="1", LI_B200_TRACE="1"))
        tail = [l for l in (r.stdout + r.stderr).splitlines() if "adj stream" not in l][-4:]
        print(args, "rc", r.returncode, "|", " / ".join(tail)[-400:])
