"""TEST INFRASTRUCTURE -- CPU restatement (numpy, Torus64) of the reference's KMS multi-key bootstrapped-gate path
(Kwak-Min-Song; `mk_gate_nand_new` / `mk_bootstrap_new`).  Only tests/ may import this file; the product never does -- and the product
has NO engine for this scheme yet: this file is the "oracle first" step of SURVEY.md section 8(f) rank 4, third slice.

Follows 3-gen-mk-tfhe/src/:
  new_mk_internals.jl:1-36     BootstrapKeyPart_new: a fresh RLWE key R per party, TGSW encryptions of the LWE key bits under R
                               (tgsw.jl:62-105), the uni-encryption of R under the party's RLWE key z (mk_internals.jl:390-446)
  new_mk_internals.jl:85-127   UniProduct_new (one v for the whole sample, unlike the CCS product)
  new_mk_internals.jl:180-209  mk_mux_rotate_new (TLev accumulator, tgsw_intern_mul = one external product per TLev row, tlev.jl:88-95),
                               mk_lev_rlwe_mul (tlev_extern_mul of every polynomial, tlev.jl:75-79, then f - UniProduct_new(e))
  new_mk_internals.jl:212-246  mk_ith_blind_rotate, mk_blind_rotate_new (parties outer; zero rotations skipped)
  new_mk_internals.jl:268-318  mk_blind_rotate_and_extract_new, mk_rlwe_extract_sample_64 (t64tot32), mk_bootstrap_wo_keyswitch_new, mk_bootstrap_new
  mk_internals.jl:162-174, 220-262, 712-726   SharedKey, PublicKey (b = z a + e), mk_keyswitch;  keyswitch.jl:14-80
  tgsw.jl:1-34, 112-138        gadget values, offset, decompose (64-bit);  tlev.jl:37-64  tlev_trivial_int
  new_mk_gates.jl:1-7, mk_api.jl:501-517, 607-610   mk_gate_nand_new, mk_encrypt_new, decryption by the phase
Three gadget triples: gsw (the per-party blind rotation), lev (the TLev accumulator's rows), uni (the hybrid product).
Exact wrap-around integer products (mod 2^64) stand for the reference's Float64 FFT, which is only approximate for Torus64 operands.
Parity pin: the reference holds no test for this path (multikey_new.jl is a demo, measurements_KMS_3.jl a timing script): PARITY UNPINNED,
pinned functionally by the decrypted NAND truth table (tests/test_kms_oracle.py).  Pure numpy loops: toy ring sizes only.
"""
import numpy as np


def negacyclic_mul(a, b):
    """a * b mod (X^N + 1, 2^64): both int64, wrap-around arithmetic (np.convolve on int64 wraps)."""
    a, b = np.asarray(a, np.int64), np.asarray(b, np.int64)
    N = a.shape[-1]
    with np.errstate(over="ignore"):
        full = np.convolve(a, b)
        out = full[:N].copy()
        out[: N - 1] -= full[N:]
    return out


def mul_by_monomial(p, s):
    """p X^s mod X^N + 1 for any integer s (reduced mod 2N)."""
    p = np.asarray(p, np.int64)
    N = p.shape[-1]
    s %= 2 * N
    idx = (np.arange(N) - s) % (2 * N)
    with np.errstate(over="ignore"):
        v = p[..., idx % N]
        return np.where(idx >= N, -v, v)


def gadget(l, bgbit):
    return np.array([np.int64(1) << np.int64(64 - (q + 1) * bgbit) for q in range(l)], np.int64)          # tgsw.jl:26


def decompose(c, l, bgbit):
    """tgsw.jl:112-138 with bit = 64: signed digits in [-Bg/2, Bg/2), int64 [l][N]."""
    c = np.asarray(c, np.int64)
    with np.errstate(over="ignore"):
        off = np.int64(0)
        for q in range(1, l + 1):
            off = off + (np.int64(1) << np.int64(64 - q * bgbit + bgbit - 1))                                  # (Bg / 2) * gadget_q, wrapped
        t = c + off
        return np.stack([((t >> np.int64(64 - q * bgbit)) & np.int64((1 << bgbit) - 1)) - np.int64(1 << (bgbit - 1)) for q in range(1, l + 1)])


def dtot64(d):
    """numeric-functions.jl:105-107: trunc(Int64, d * 2^64)."""
    return np.trunc(np.asarray(d, np.float64) * 2.0 ** 64).astype(np.int64)


def t64tot32(v):
    """numeric-functions.jl:109-111."""
    return np.trunc(np.asarray(v, np.int64).astype(np.float64) / 2.0 ** 32).astype(np.int32)


def decode_message(x, space):
    """numeric-functions.jl:70-73 on Int32."""
    log2 = space.bit_length() - 1
    x = np.asarray(x, np.int32)
    with np.errstate(over="ignore"):
        return (x + np.int32(1 << (32 - log2 - 1))) >> np.int32(32 - log2)


def _inner(digits, rows):
    out = np.zeros(rows.shape[-1], np.int64)
    with np.errstate(over="ignore"):
        for q in range(digits.shape[0]):
            out = out + negacyclic_mul(digits[q], rows[q])
    return out


def keygen(rng, prm):
    """mk_api.jl:341-347 (SharedKey over the uni gadget), :413-436 (CloudKeyPart_new) per party.  prm: n, N, k, gsw = (l, bgbit),
    lev = (l, bgbit), uni = (l, bgbit), t, basebit, sigma_gsw, sigma_uni, sigma_ks."""
    n, N, k = prm["n"], prm["N"], prm["k"]
    (lg, bg), (ll, bl), (lu, bu) = prm["gsw"], prm["lev"], prm["uni"]
    g_gsw, g_uni = gadget(lg, bg), gadget(lu, bu)
    uni64 = lambda shape: rng.integers(-2 ** 63, 2 ** 63 - 1, shape, dtype=np.int64, endpoint=True)
    gauss = lambda sigma, shape: dtot64(rng.standard_normal(shape) * sigma)
    K = dict(prm)
    K["a"] = uni64((lu, N))                                                  # SharedKey.a
    K["s"] = rng.integers(0, 2, (k, n)).astype(np.int32)                     # LWE keys
    K["z"] = rng.integers(0, 2, (k, N)).astype(np.int64)                     # RLWE keys (binary, rlwe.jl:17-22)
    K["R"] = np.empty((k, N), np.int64)                                      # the parties' blind-rotation keys (kept for the tests only)
    K["pk"] = np.empty((k, lu, N), np.int64)
    K["gsw_rows"] = np.empty((k, n, lg, 2, 2, N), np.int64)                  # [party][j][q][row: gadget on mask / body][mask, body][N]
    K["d"], K["f0"], K["f1"] = (np.empty((k, lu, N), np.int64) for _ in range(3))
    B1 = (1 << prm["basebit"]) - 1
    K["ksk"] = np.empty((k, N, prm["t"], B1, n + 1), np.int32)
    with np.errstate(over="ignore"):
        for p in range(k):
            z = K["z"][p]
            for q in range(lu):
                K["pk"][p, q] = negacyclic_mul(z, K["a"][q]) + gauss(prm["sigma_uni"], N)                     # mk_internals.jl:225-262
            R = K["R"][p] = rng.integers(0, 2, N).astype(np.int64)            # rand_key (new_mk_internals.jl:21)
            for j in range(n):                                                # tgsw_encrypt(lwe_key.key[j], alpha_gsw, rand_key, tgsw_params)
                for q in range(lg):
                    for row in range(2):
                        mask = uni64(N)
                        body = negacyclic_mul(R, mask) + gauss(prm["sigma_gsw"], N)                           # rlwe_encrypt_zero, rlwe.jl:79-110
                        pair = [mask, body]
                        pair[row][0] += np.int64(int(K["s"][p, j])) * g_gsw[q]                                # tgsw.jl:65-86: message * gadget on component `row`
                        K["gsw_rows"][p, j, q, row, 0], K["gsw_rows"][p, j, q, row, 1] = pair
            r = rng.integers(0, 2, N).astype(np.int64)                        # uni-encryption of the POLYNOMIAL R under z (mk_internals.jl:390-446)
            for q in range(lu):
                K["d"][p, q] = negacyclic_mul(r, K["a"][q]) + gauss(prm["sigma_uni"], N) + R * g_uni[q]
                K["f1"][p, q] = uni64(N)
                K["f0"][p, q] = negacyclic_mul(z, K["f1"][p, q]) + gauss(prm["sigma_uni"], N) + r * g_uni[q]
            noise = rng.standard_normal((N, prm["t"], B1)) * prm["sigma_ks"]                                  # keyswitch.jl:14-41
            noise -= noise.mean()
            ka = rng.integers(-2 ** 31, 2 ** 31, (N, prm["t"], B1, n)).astype(np.int32)
            h = np.arange(1, B1 + 1, dtype=np.int64)[None, None, :]
            sh = (32 - np.arange(1, prm["t"] + 1) * prm["basebit"])[None, :, None]
            msg = (z[:, None, None] * h) << sh
            kb = (msg + np.trunc(noise * 2.0 ** 32).astype(np.int64) + (ka.astype(np.int64) * K["s"][p]).sum(-1)).astype(np.int32)
            K["ksk"][p] = np.concatenate([ka, kb[..., None]], axis=-1)
    return K


def tgsw_extern_mul(rl, G, l, bgbit):
    """tgsw.jl:143-147 on an RLWE pair rl = [mask, body]; G: [l][row 2][mask, body][N]."""
    dm, db = decompose(rl[0], l, bgbit), decompose(rl[1], l, bgbit)
    out = np.zeros((2, rl.shape[-1]), np.int64)
    with np.errstate(over="ignore"):
        for c in range(2):
            out[c] = _inner(dm, G[:, 0, c]) + _inner(db, G[:, 1, c])
    return out


def ith_blind_rotate(K, party, bara):
    """mk_ith_blind_rotate (new_mk_internals.jl:212-225): TLev encryption of X^(sum_j bara_j s_j) under the party's R; int64 [l_lev][2][N]."""
    (lg, bg), (ll, bl) = K["gsw"], K["lev"]
    N = K["N"]
    lev = np.zeros((ll, 2, N), np.int64)
    lev[:, 1, 0] = gadget(ll, bl)                                              # tlev_trivial_int(1): body += gadget (tlev.jl:37-64)
    with np.errstate(over="ignore"):
        for j in range(K["n"]):
            if bara[j] != 0:
                for q in range(ll):                                            # mk_mux_rotate_new + tgsw_intern_mul: one external product per row
                    temp = mul_by_monomial(lev[q], int(bara[j])) - lev[q]
                    lev[q] = lev[q] + tgsw_extern_mul(temp, K["gsw_rows"][party, j], lg, bg)
    return lev


def uni_product(acc_a, acc_b, K, party):
    """UniProduct_new (new_mk_internals.jl:85-127): acc_a int64 [k][N], acc_b [N]; `party` 0-based."""
    k = K["k"]
    lu, bu = K["uni"]
    dec_a = [decompose(acc_a[i], lu, bu) for i in range(k)]
    dec_b = decompose(acc_b, lu, bu)
    d, f0, f1 = K["d"][party], K["f0"][party], K["f1"][party]
    with np.errstate(over="ignore"):
        u = np.stack([_inner(dec_a[i], d) for i in range(k)])
        u0 = _inner(dec_b, d)
        v = np.zeros(K["N"], np.int64)
        for i in range(k):
            v = v + _inner(dec_a[i], K["pk"][i])
        v = v - _inner(dec_b, K["a"])
        dv = decompose(v, lu, bu)
        u[party] = u[party] + _inner(dv, f1)
        return u, u0 + _inner(dv, f0)


def lev_rlwe_mul(acc_a, acc_b, lev, K, party):
    """mk_lev_rlwe_mul (new_mk_internals.jl:186-209): only the masks of the parties before `party` are multiplied (the others are still zero)."""
    k, N = K["k"], K["N"]
    ll, bl = K["lev"]
    e_a, f_a = np.zeros((k, N), np.int64), np.zeros((k, N), np.int64)

    def tlev_extern_mul(c):                                                     # tlev.jl:75-79
        dc = decompose(c, ll, bl)
        return _inner(dc, lev[:, 0]), _inner(dc, lev[:, 1])
    for i in range(party):
        e_a[i], f_a[i] = tlev_extern_mul(acc_a[i])
    e_b, f_b = tlev_extern_mul(acc_b)
    ua, ub = uni_product(e_a, e_b, K, party)
    with np.errstate(over="ignore"):
        return f_a - ua, f_b - ub


def bootstrap_wo_keyswitch(K, mu, xa, xb):
    """mk_bootstrap_wo_keyswitch_new (:301-312) with fast_boot = false: xa int32 [k][n] -> (ext_a int32 [k][N], ext_b)."""
    k, N = K["k"], K["N"]
    barb, bara = decode_message(xb, 2 * N), decode_message(xa, 2 * N)
    acc_a = np.zeros((k, N), np.int64)
    acc_b = mul_by_monomial(np.full(N, mu, np.int64), -int(barb))
    for p in range(k):                                                          # mk_blind_rotate_new (:243-254)
        lev = ith_blind_rotate(K, p, bara[p])
        acc_a, acc_b = lev_rlwe_mul(acc_a, acc_b, lev, K, p)
    with np.errstate(over="ignore"):                                            # mk_rlwe_extract_sample_64 (:292-297): reverse_polynomial, t64tot32
        ext_a = np.stack([t64tot32(np.concatenate([acc_a[p][:1], -acc_a[p][:0:-1]])) for p in range(k)])
    return ext_a, t64tot32(acc_b[0])


def keyswitch(ksk, a, b, t, basebit):
    """keyswitch.jl:45-80 on one LWE sample of dimension N."""
    N = a.shape[0]
    with np.errstate(over="ignore"):
        abar = a.astype(np.int32) + np.int32(1 << (32 - (1 + basebit * t)))
        ra, rb = np.zeros(ksk.shape[-1] - 1, np.int32), np.int32(b)
        for i in range(N):
            for j in range(1, t + 1):
                d = (int(abar[i]) >> (32 - j * basebit)) & ((1 << basebit) - 1)
                if d:
                    ra = ra - ksk[i, j - 1, d - 1, :-1]
                    rb = np.int32(rb - ksk[i, j - 1, d - 1, -1])
    return ra, rb


def mk_keyswitch(K, ext_a, ext_b):
    """mk_internals.jl:712-726."""
    out, b = [], np.int32(ext_b)
    with np.errstate(over="ignore"):
        for p in range(K["k"]):
            ra, rb = keyswitch(K["ksk"][p], ext_a[p], 0, K["t"], K["basebit"])
            out.append(ra)
            b = np.int32(b + rb)
    return np.stack(out), b


def gate_nand(K, x, y):
    """mk_gate_nand_new (new_mk_gates.jl:1-7): bootstrap(encode(1, 8) - x - y) with the Torus64 output message encode_message64(1, 8)."""
    with np.errstate(over="ignore"):
        ta = (-x[0].astype(np.int64) - y[0]).astype(np.int32)
        tb = np.int64((1 << 29) - int(x[1]) - int(y[1])).astype(np.int32)
    return mk_keyswitch(K, *bootstrap_wo_keyswitch(K, np.int64(1) << np.int64(61), ta, tb))


def encrypt(rng, K, bit, sigma):
    """mk_encrypt_new (mk_api.jl:501-517)."""
    a = rng.integers(-2 ** 31, 2 ** 31, (K["k"], K["n"])).astype(np.int32)
    e = int(np.trunc(rng.standard_normal() * sigma * 2.0 ** 32))
    b = np.int64((1 << 29) * (1 if bit else -1) + e + int((a.astype(np.int64) * K["s"]).sum())).astype(np.int32)
    return a, b


def phase(K, x):
    return int(np.int64(int(x[1]) - int((x[0].astype(np.int64) * K["s"]).sum())).astype(np.int32))


def decrypt(K, x):
    return phase(K, x) > 0
