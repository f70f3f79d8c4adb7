"""Diagnostic (kernel built with -DMK_PHASE_TRACE=1, selected through MKTFHE_B200_LIB): SM clock at the start of every blind-rotate step of
the two gates of CTA 0 and CTA 1 -> how far apart in phase the two gates of a CTA run, and the spread of the step times."""
import sys
import numpy as np
sys.path.insert(0, '/root/repo')
import torus_fhe_b200 as T
n, k, N = 520, 2, 1024
ctx = T._cabi.Context(n, N, k, 2, 7, 3, 3, device=0)
r = np.random.default_rng(1)
for p in range(k):
    ctx.load_bsk(p, r.integers(-2**63, 2**63 - 1, (n, 4, 2, N), dtype=np.int64))
    ctx.load_ksk(p, r.integers(-2**31, 2**31, (N, 3, 7, n + 1)).astype(np.int32))
ctx.finalize_keys()
G = 2 * 296
a = r.integers(-2**31, 2**31, (G, k, n)).astype(np.int32); b = r.integers(-2**31, 2**31, G).astype(np.int32)
for rep in range(2):
    ext, acc = ctx.blind_rotate_batch(1 << 61, a, b, want_acc=True)
kn = k * n
for cta in (0, 1):
    t0, t1 = acc[2 * cta].reshape(-1)[:kn].astype(np.int64), acc[2 * cta + 1].reshape(-1)[:kn].astype(np.int64)
    ok = (t0 > 0) & (t1 > 0)                       # steps skipped by a == 0 leave no stamp
    d0 = np.diff(t0[ok]); step = np.median(d0)
    off = (t1 - t0)[ok]
    print(f"CTA {cta}: step median {step:.0f} clk (p5 {np.percentile(d0,5):.0f}, p95 {np.percentile(d0,95):.0f}); gate1 - gate0 start offset: first 5 steps {off[:5]}, "
          f"median {np.median(off):.0f} clk = {np.median(off)/step:+.2f} step, p5 {np.percentile(off,5):.0f}, p95 {np.percentile(off,95):.0f}, last {off[-3:]}")
