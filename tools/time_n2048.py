import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import torus_fhe_b200 as T
# two parties at the 16-party parameter shape (n = 590): per-step time of the N = 2048 kernel, synthetic random keys
n, k, N = 590, 2, 2048
ctx = T._cabi.Context(n, N, k, 1, 26, 4, 3, device=0)
r = np.random.default_rng(1)
for p in range(k):
    ctx.load_bsk(p, r.integers(-2**63, 2**63-1, (n, 4, 1, N), dtype=np.int64))
    ctx.load_ksk(p, r.integers(-2**31, 2**31, (N, 4, 7, n + 1)).astype(np.int32))
ctx.finalize_keys()
for G in ([int(v) for v in sys.argv[1:]] or (1, 148, 296)):
    a = r.integers(-2**31, 2**31, (G, k, n)).astype(np.int32); b = r.integers(-2**31, 2**31, G).astype(np.int32)
    ctx.bootstrap_batch(1 << 61, a, b)
    ctx.bootstrap_batch(1 << 61, a, b)
    br, ksw = ctx.last_kernel_ms()
    print(f"N=2048 n={n} k={k}: G={G}: blind rotate {br:.2f} ms ({1e3*br/(k*n):.2f} us/step), key switch {ksw:.2f} ms -> {G/(br+ksw)*1e3:.0f} bootstraps/s")
