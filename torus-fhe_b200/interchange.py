"""Binary interchange format for 3gen keys and ciphertext batches.

Purpose: a host that CAN run the reference (Julia) dumps its integer keys / ciphertexts into these files, and this engine loads the
very same bytes -- the missing link for pinning parity against the real reference.  The Julia writer is `write_keys` /
`write_ciphertexts` in julia/TFHE_B200.jl.  Everything is little-endian, no padding.

keys file:
    char[8]  "MKTFHE3K"      uint32 version = 1
    int32    n, N, k, l, bgbit, t, basebit, has_secrets
    float64  lwe_noise_stddev, gsw_noise_stddev, ks_noise_stddev
    per party p = 0..k-1:  int64 bsk[n][4][l][N]    (BootstrapKeyPart_3gen.gsw_key[j].part_{1..4}[q].coeffs, 3gen_mk_internals.jl:10-43)
    per party p = 0..k-1:  int32 ksk[N][t][2^basebit - 1][n+1]   (KeyswitchKey.key[h, j, i] as (a, b) rows, keyswitch.jl:7-42)
    if has_secrets:        int32 lwe_key[k][n]      (SecretKey_3gen.key.key, api.jl:196-204; fixtures only, never production keys)
ciphertext file:
    char[8]  "MKTFHE3C"      uint32 version = 1
    int32    k, n;  uint64 count
    int32    a[count][k][n]  (MKLweSample.a, column-major (n, k) in Julia == this order);  int32 b[count]
"""
import struct

import numpy as np

KEY_MAGIC, CT_MAGIC, VERSION = b"MKTFHE3K", b"MKTFHE3C", 1


def write_keys(path, params, bsk_parts, ksk_parts, lwe_keys=None):
    k, n, N, l = params.max_parties, params.lwe_size, params.rlwe_polynomial_degree, params.gsw_decomp_length
    t, bb = params.ks_decomp_length, params.ks_log2_base
    with open(path, "wb") as f:
        f.write(KEY_MAGIC + struct.pack("<I", VERSION))
        f.write(struct.pack("<8i", n, N, k, l, params.gsw_log2_base, t, bb, 0 if lwe_keys is None else 1))
        f.write(struct.pack("<3d", params.lwe_noise_stddev, params.gsw_noise_stddev, params.ks_noise_stddev))
        for p in range(k):
            a = np.ascontiguousarray(bsk_parts[p], dtype="<i8")
            assert a.shape == (n, 4, l, N), a.shape
            f.write(a.tobytes())
        for p in range(k):
            a = np.ascontiguousarray(ksk_parts[p], dtype="<i4")
            assert a.shape == (N, t, (1 << bb) - 1, n + 1), a.shape
            f.write(a.tobytes())
        if lwe_keys is not None:
            f.write(np.ascontiguousarray(lwe_keys, dtype="<i4").reshape(k, n).tobytes())


def read_keys(path):
    """-> (SchemeParameters_3gen, bsk_parts, ksk_parts, lwe_keys or None)"""
    from .tfhe3gen import SchemeParameters_3gen
    with open(path, "rb") as f:
        if f.read(8) != KEY_MAGIC:
            raise ValueError(f"{path}: not a MKTFHE3K key file")
        (ver,) = struct.unpack("<I", f.read(4))
        if ver != VERSION:
            raise ValueError(f"{path}: unsupported version {ver}")
        n, N, k, l, bgbit, t, bb, has_sec = struct.unpack("<8i", f.read(32))
        s_lwe, s_gsw, s_ks = struct.unpack("<3d", f.read(24))
        params = SchemeParameters_3gen(n, s_lwe, N, 1, False, l, bgbit, s_gsw, t, bb, s_ks, k)
        B1 = (1 << bb) - 1
        bsk = [np.frombuffer(f.read(n * 4 * l * N * 8), dtype="<i8").reshape(n, 4, l, N) for _ in range(k)]
        ksk = [np.frombuffer(f.read(N * t * B1 * (n + 1) * 4), dtype="<i4").reshape(N, t, B1, n + 1) for _ in range(k)]
        lwe = np.frombuffer(f.read(k * n * 4), dtype="<i4").reshape(k, n) if has_sec else None
        if f.read(1):
            raise ValueError(f"{path}: trailing bytes")
    return params, bsk, ksk, lwe


def write_ciphertexts(path, a, b):
    a, b = np.ascontiguousarray(a, dtype="<i4"), np.ascontiguousarray(b, dtype="<i4").reshape(-1)
    count, k, n = b.size, a.shape[-2], a.shape[-1]
    with open(path, "wb") as f:
        f.write(CT_MAGIC + struct.pack("<I", VERSION) + struct.pack("<2i", k, n) + struct.pack("<Q", count))
        f.write(a.reshape(count, k, n).tobytes())
        f.write(b.tobytes())


def read_ciphertexts(path):
    with open(path, "rb") as f:
        if f.read(8) != CT_MAGIC:
            raise ValueError(f"{path}: not a MKTFHE3C ciphertext file")
        (ver,) = struct.unpack("<I", f.read(4))
        if ver != VERSION:
            raise ValueError(f"{path}: unsupported version {ver}")
        k, n = struct.unpack("<2i", f.read(8))
        (count,) = struct.unpack("<Q", f.read(8))
        a = np.frombuffer(f.read(count * k * n * 4), dtype="<i4").reshape(count, k, n)
        b = np.frombuffer(f.read(count * 4), dtype="<i4")
    return a, b
