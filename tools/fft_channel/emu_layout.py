import numpy as np, sys
sys.path.insert(0, '/root/repo/tools/fft_channel')
import proto
tw = np.fromfile('/tmp/fftemu/tw.bin', np.float64).reshape(-1, 2)
tw = tw[:, 0] + 1j * tw[:, 1]
TF_A, TF_B = 0, 32
TI_A = TF_B + 480; T_WJ = TI_A + 240; T_UT = T_WJ + 256; T_UT2 = T_UT + 256
M = 512; N = 1024
C = 0.70710678118654752440
def ct(a, b, w): t = w * b; return a + t, a - t
def fwd_stage0_real(a0, a1, a2, a3, h):
    cs = -C if h else C
    return (a0 + cs * (a1 - a3)) + 1j * (a2 + cs * (a1 + a3))
def fwd_passA(v, h):   # v [16] for one thread
    w = tw[TF_A + h]
    for r in range(8): v[r], v[r + 8] = ct(v[r], v[r + 8], w)
    for g in range(2):
        w = tw[TF_A + 2 + 2 * h + g]
        for r in range(4): v[8 * g + r], v[8 * g + r + 4] = ct(v[8 * g + r], v[8 * g + r + 4], w)
    for g in range(4):
        w = tw[TF_A + 6 + 4 * h + g]
        for r in range(2): v[4 * g + r], v[4 * g + r + 2] = ct(v[4 * g + r], v[4 * g + r + 2], w)
    for g in range(8):
        w = tw[TF_A + 14 + 8 * h + g]
        v[2 * g], v[2 * g + 1] = ct(v[2 * g], v[2 * g + 1], w)
def fwd_passB(v, lane):
    w = tw[TF_B + lane]
    for c in range(8): v[c], v[c + 8] = ct(v[c], v[c + 8], w)
    for g in range(2):
        w = tw[TF_B + 32 + g * 32 + lane]
        for c in range(4): v[8 * g + c], v[8 * g + c + 4] = ct(v[8 * g + c], v[8 * g + c + 4], w)
    for g in range(4):
        w = tw[TF_B + 96 + g * 32 + lane]
        for c in range(2): v[4 * g + c], v[4 * g + c + 2] = ct(v[4 * g + c], v[4 * g + c + 2], w)
    for g in range(8):
        w = tw[TF_B + 224 + g * 32 + lane]
        v[2 * g], v[2 * g + 1] = ct(v[2 * g], v[2 * g + 1], w)
def inv_passB(v):
    C1, S1 = 0.92387953251128675613, 0.38268343236508977173
    for c in range(0, 16, 2): v[c], v[c + 1] = ct(v[c], v[c + 1], 1)
    for c in range(0, 16, 4):
        v[c], v[c + 2] = ct(v[c], v[c + 2], 1); v[c + 1], v[c + 3] = ct(v[c + 1], v[c + 3], -1j)
    for c in range(0, 16, 8):
        v[c], v[c + 4] = ct(v[c], v[c + 4], 1)
        v[c + 1], v[c + 5] = ct(v[c + 1], v[c + 5], C - 1j * C)
        v[c + 2], v[c + 6] = ct(v[c + 2], v[c + 6], -1j)
        v[c + 3], v[c + 7] = ct(v[c + 3], v[c + 7], -C - 1j * C)
    ws = [1, C1 - 1j * S1, C - 1j * C, S1 - 1j * C1, -1j, -S1 - 1j * C1, -C - 1j * C, -C1 - 1j * S1]
    for j in range(8): v[j], v[j + 8] = ct(v[j], v[j + 8], ws[j])
def inv_passA(v, l16):
    rs = 1
    while rs <= 8:
        for e in range(rs):
            w = tw[TI_A + 16 * (rs - 1) + e * 16 + l16]
            for r0 in range(0, 16, 2 * rs): v[r0 + e], v[r0 + e + rs] = ct(v[r0 + e], v[r0 + e + rs], w)
        rs *= 2
def rows_to_cols(V):   # V [32 lanes][16]
    buf = np.zeros(512, complex)
    for lane in range(32):
        h, l16 = lane >> 4, lane & 15
        for r in range(16): buf[h * 256 + r * 16 + (l16 ^ r)] = V[lane][r]
    for lane in range(32):
        h, l16 = lane >> 4, lane & 15
        for c in range(16): V[lane][c] = buf[h * 256 + l16 * 16 + (c ^ l16)]
def cols_to_rows(V):
    buf = np.zeros(512, complex)
    for lane in range(32):
        h, l16 = lane >> 4, lane & 15
        for c in range(16): buf[h * 256 + l16 * 16 + (c ^ l16)] = V[lane][c]
    for lane in range(32):
        h, l16 = lane >> 4, lane & 15
        for r in range(16): V[lane][r] = buf[h * 256 + r * 16 + (l16 ^ r)]
def warp_forward(a):   # a real [1024] -> spectrum layout [c*32+lane]
    V = [[0j] * 16 for _ in range(32)]
    for lane in range(32):
        h, l16 = lane >> 4, lane & 15
        for r in range(16):
            j = 16 * r + l16
            V[lane][r] = fwd_stage0_real(a[j], a[j + 256], a[j + 512], a[j + 768], h)
        fwd_passA(V[lane], h)
    rows_to_cols(V)
    out = np.zeros(512, complex)
    for lane in range(32):
        fwd_passB(V[lane], lane)
        for c in range(16): out[c * 32 + lane] = V[lane][c]
    return out
def warp_inverse(spec):  # spec layout [c*32+lane] -> Y natural pos (before last stage)
    V = [[spec[c * 32 + lane] for c in range(16)] for lane in range(32)]
    for lane in range(32): inv_passB(V[lane])
    cols_to_rows(V)
    Y = np.zeros(512, complex)
    for lane in range(32):
        h, l16 = lane >> 4, lane & 15
        inv_passA(V[lane], l16)
        for r in range(16): Y[h * 256 + r * 16 + l16] = V[lane][r]
    return Y
def p4(Y):
    c = np.zeros(N)
    for j in range(256):
        lo, hi = ct(Y[j], Y[j + 256], tw[T_WJ + j])
        e, f = lo * tw[T_UT + j], hi * tw[T_UT2 + j]
        c[j], c[j + 256], c[j + 512], c[j + 768] = e.real, f.real, e.imag, f.imag
    return c
rng = np.random.default_rng(3)
a = rng.integers(-64, 64, N).astype(float)
# 1. forward matches the prototype network up to the thread layout: position pos = lane*16 + c
A = warp_forward(a); P = proto.forward(proto.fold(a))
lay = np.array([P[lane * 16 + c] for c in range(16) for lane in range(32)])
print("forward max err", np.abs(A - lay).max())
# 2. full product
b = rng.integers(-2**21, 2**21, N).astype(float)
B = warp_forward(b) / 512
c = p4(warp_inverse(A * B))
want = proto.negacyclic_exact(a.astype(np.int64), b.astype(np.int64))
print("product max err", np.abs(c - want).max(), "exact after rounding:", np.array_equal(np.rint(c).astype(np.int64), want))
