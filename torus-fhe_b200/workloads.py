"""The reference's dark-pool volume-matching workload (3-gen-mk-tfhe/VolumeMatching.jl:31-190) on level-batched circuits.

The reference fans sub-circuits out over Julia `Distributed` workers (`addprocs(20)`, `@spawnat`), shipping the whole key set in
every closure.  Here the independent pieces become INSTANCES of one batched circuit:
  * the two prefix-sum chains (buy, sell) advance together as two instances of one adder per step       (:47-70)
  * `sellGRTbuy`, then the WIDTH muxes selecting the smaller side's total as one level of 2·WIDTH ANDs    (:94-103)
  * the m + n per-order sub / leq / mux blocks run as m + n instances of the same three circuits          (:105-177)
Semantics are the reference's: every worker subtracts its order's prefix from the SAME `total1` (closures capture a copy), the mux
outputs are not bootstrapped, and `res = ordLeq ? order : total - prefix`.
"""
import numpy as np

from .circuits import add_mod_3gen, gate_level, mk_add_3gen, mk_grt_3gen, mk_int_mul_3gen, mk_int_mul_lo_3gen, mk_leq_3gen, mk_sub_3gen
from .engine import shard_bounds
from .tfhe3gen import MKLweSample, MKLweSampleGPU, encode_message, mk_lwe_noiseless_trivial


def _stack(samples):
    return MKLweSample(samples[0].params, np.stack([s.a for s in samples]), np.stack([s.b for s in samples]), 0.0)


def _mux_bits(bk, ks, sel, ys, zs):
    """WIDTH muxes with one selector in ONE launch: mk_gate_mux_3gen (3gen_mk_gates.jl:133-150) = AND(x, y) and AND(-x, z) bootstrapped,
    then 1/8 + t1 + t2 not bootstrapped."""
    W = len(ys)
    outs = gate_level(bk, ks, [("and", sel, ys[j]) for j in range(W)] + [("and", -sel, zs[j]) for j in range(W)])
    k = len(bk)
    return [mk_lwe_noiseless_trivial(encode_message(1, 8), outs[j].params, k, outs[j].b.shape) + outs[j] + outs[W + j] for j in range(W)]


def VolumeMatch(bk, ks, ord_buy, ord_sell, accBuy, accSell, zero_arr, one, zero, WIDTH):
    """ord_buy / ord_sell: lists of encrypted WIDTH-bit orders (lists of MKLweSample bits, batch shape ()); returns the matched volumes
    (res_buy_orders, res_sell_orders) as lists of bit lists, like `resOrd.buy` / `resOrd.sell` in the reference."""
    m, n = len(ord_buy), len(ord_sell)
    # prefix sums: res_buy[i] = sum of the first i buy orders (res_buy[0] = zero_arr), same for sell; the two chains are 2 instances
    res_buy, res_sell = [zero_arr], [zero_arr]
    acc = [accBuy, accSell]
    for step in range(max(m, n)):
        live = [c for c, cnt in ((0, m), (1, n)) if step < cnt]
        orders = [ord_buy[step] if c == 0 else ord_sell[step] for c in live]
        a_bits = [_stack([acc[c][j] for c in live]) for j in range(WIDTH)]
        x_bits = [_stack([o[j] for o in orders]) for j in range(WIDTH)]
        s_bits = mk_add_3gen(bk, ks, a_bits, x_bits, zero, WIDTH)
        for idx, c in enumerate(live):
            acc[c] = [s_bits[j][idx] for j in range(WIDTH)]
            (res_buy if c == 0 else res_sell).append(acc[c])
    accBuy, accSell = acc
    sellGRTbuy = mk_grt_3gen(bk, ks, accSell, accBuy, one, WIDTH)
    total = _mux_bits(bk, ks, sellGRTbuy, accBuy, accSell)          # the smaller of the two totals
    if m + n == 0:
        return [], []
    # per order: diff = total - prefix_before_order; res = (order <= diff) ? order : diff      -- m + n instances at once
    orders = list(ord_buy) + list(ord_sell)
    prefixes = res_buy[:m] + res_sell[:n]
    o_bits = [_stack([o[j] for o in orders]) for j in range(WIDTH)]
    p_bits = [_stack([p[j] for p in prefixes]) for j in range(WIDTH)]
    diff = mk_sub_3gen(bk, ks, total, p_bits, one, WIDTH)
    ord_leq = mk_leq_3gen(bk, ks, o_bits, diff, one, WIDTH)
    W = WIDTH
    outs = gate_level(bk, ks, [("and", ord_leq, o_bits[j]) for j in range(W)] + [("and", -ord_leq, diff[j]) for j in range(W)])
    k = len(bk)
    res = [mk_lwe_noiseless_trivial(encode_message(1, 8), outs[j].params, k, outs[j].b.shape) + outs[j] + outs[W + j] for j in range(W)]
    per_order = [[res[j][i] for j in range(W)] for i in range(m + n)]
    return per_order[:m], per_order[m:]


def volume_match_plain(buy, sell):
    """Plaintext model of the same computation (what the decrypted outputs must equal)."""
    pb = np.concatenate([[0], np.cumsum(buy)])
    ps = np.concatenate([[0], np.cumsum(sell)])
    total = pb[-1] if ps[-1] > pb[-1] else ps[-1]          # sellGRTbuy ? accBuy : accSell
    rb = [int(o) if o <= total - pb[i] else int(total - pb[i]) for i, o in enumerate(buy)]
    rs = [int(o) if o <= total - ps[i] else int(total - ps[i]) for i, o in enumerate(sell)]
    return rb, rs


# ---------------------------------------------------------------------------------------------------------------------------
# Encrypted convolution layer (BASELINE configs[4]; the reference's enc_conv2d, 3gen_mk_gates.jl:364-397, does not run: it indexes
# `0:kernel_size` inclusive from 1-based `i * stride + m`, passes single samples where bit vectors are expected, calls the
# commented-out mk_int_add_3gen, accumulates into an undefined `sum` and never stores it).  Same signature and intent -- every output
# is the sum over the kernel window of encrypted-input x encrypted-weight products, computed with the integer circuits -- with the
# indexing defined as a plain valid cross-correlation.
# ---------------------------------------------------------------------------------------------------------------------------
def _is_gpu(x):
    return isinstance(x, MKLweSampleGPU)


def _take(x, idx):
    """x[idx] for a tuple of integer index arrays over the leading (batch) dimensions, host or device sample."""
    if _is_gpu(x):
        import torch
        idx = tuple(torch.as_tensor(np.array(i), device=x.a.device) for i in idx)
    return x[idx]


def _cat(xs, axis):
    if _is_gpu(xs[0]):
        import torch
        return MKLweSampleGPU(xs[0].params, torch.cat([x.a for x in xs], axis), torch.cat([x.b for x in xs], axis), 0.0)
    return MKLweSample(xs[0].params, np.concatenate([x.a for x in xs], axis), np.concatenate([x.b for x in xs], axis), 0.0)


def _pad(x, zero, p):
    """Zero-pad a (H, W) batch of bit samples with copies of the encrypted ZERO bit."""
    if p == 0:
        return x
    H, W = x.batch_shape
    if _is_gpu(x):
        a = zero.a.expand(H + 2 * p, W + 2 * p, *zero.a.shape[-2:]).clone()
        b = zero.b.expand(H + 2 * p, W + 2 * p).clone()
    else:
        a = np.broadcast_to(zero.a, (H + 2 * p, W + 2 * p) + zero.a.shape[-2:]).copy()
        b = np.broadcast_to(zero.b, (H + 2 * p, W + 2 * p)).copy()
    a[p:p + H, p:p + W] = x.a
    b[p:p + H, p:p + W] = x.b
    return type(x)(x.params, a, b, 0.0)


def conv2d_output_shape(in_shape, kernel_shape, stride, padding):
    (H, W), (C, K, K2) = in_shape, kernel_shape
    assert K == K2
    return C, (H + 2 * padding - K) // stride + 1, (W + 2 * padding - K) // stride + 1


def enc_conv2d(bk, ks, input, ZERO, kernels, stride, padding, WIDTH, shard=None, mul="exact"):
    """input: WIDTH bit samples (LSB first), each a batch of shape (H, W); kernels: WIDTH bit samples of batch shape (C, K, K);
    ZERO: an encrypted 0 bit (scalar sample).  Returns the WIDTH bit samples of the outputs, batch shape (C, OH, OW):
        out[c, i, j] = sum_{m, n < K} input[i * stride + m - padding, j * stride + n - padding] * kernels[c, m, n]   (mod 2^WIDTH)
    Every (output, tap) product is one instance of the multiplier circuit, so each dependency level of the whole layer is one
    launch; the window sum is a pairwise tree of `add_mod_3gen` instances.
    shard = (world_size, rank): compute only this rank's contiguous slice of the C * OH * OW outputs (flattened, row-major) and
    return (lo, hi, bits of batch shape (hi - lo,)) -- outputs are independent, so multi-GPU needs no exchange.
    mul = "exact": mk_int_mul_lo_3gen; "reference": the reference's mk_int_mul_3gen wiring, slip included."""
    C, OH, OW = conv2d_output_shape(input[0].batch_shape, kernels[0].batch_shape, stride, padding)
    K = kernels[0].batch_shape[1]
    total = C * OH * OW
    lo, hi = (0, total) if shard is None else shard_bounds(total, shard[0], shard[1])
    flat = np.arange(lo, hi)
    c, i, j = flat // (OH * OW), (flat // OW) % OH, flat % OW
    m, n = np.repeat(np.arange(K), K), np.tile(np.arange(K), K)
    rows, cols = i[:, None] * stride + m[None, :], j[:, None] * stride + n[None, :]
    cc = np.broadcast_to(c[:, None], rows.shape)
    mm, nn = np.broadcast_to(m[None, :], rows.shape), np.broadcast_to(n[None, :], rows.shape)
    x_bits = [_take(_pad(input[q], ZERO, padding), (rows, cols)) for q in range(WIDTH)]          # (S, K*K) instances
    w_bits = [_take(kernels[q], (cc, mm, nn)) for q in range(WIDTH)]
    if hi == lo:
        out = x_bits
    else:
        if mul == "exact":
            terms = mk_int_mul_lo_3gen(bk, ks, x_bits, w_bits, WIDTH)
        else:
            terms = mk_int_mul_3gen(bk, ks, x_bits, w_bits, ZERO, WIDTH)
        T = K * K
        while T > 1:                                                   # pairwise tree over the taps (last batch axis)
            h = T // 2
            left = [t[:, :h] for t in terms]
            right = [t[:, h:2 * h] for t in terms]
            summed = add_mod_3gen(bk, ks, left, right, WIDTH)
            terms = [_cat([summed[q], terms[q][:, 2 * h:]], 1) if T > 2 * h else summed[q] for q in range(WIDTH)]
            T = h + (T - 2 * h)
        out = [t[:, 0] for t in terms]
    if shard is not None:
        return lo, hi, out
    return [type(o)(o.params, o.a.reshape((C, OH, OW) + tuple(o.a.shape[-2:])), o.b.reshape((C, OH, OW)), 0.0) for o in out]


def conv2d_gate_count(in_shape, kernel_shape, stride, padding, WIDTH):
    """Bootstrapped gates of one enc_conv2d call with the exact multiplier."""
    C, OH, OW = conv2d_output_shape(in_shape, kernel_shape, stride, padding)
    T = kernel_shape[1] ** 2
    add = lambda w: 1 if w == 1 else 5 * w - 6
    mul = WIDTH * (WIDTH + 1) // 2 + sum(add(WIDTH - i) for i in range(1, WIDTH))
    return C * OH * OW * (T * mul + (T - 1) * add(WIDTH))


def conv2d_plain(inp, ker, stride, padding, WIDTH):
    """Plaintext model: valid cross-correlation of signed WIDTH-bit ints, wrapped to WIDTH bits."""
    inp, ker = np.asarray(inp, np.int64), np.asarray(ker, np.int64)
    x = np.pad(inp, padding)
    C, K = ker.shape[0], ker.shape[1]
    OH, OW = (x.shape[0] - K) // stride + 1, (x.shape[1] - K) // stride + 1
    out = np.zeros((C, OH, OW), np.int64)
    for m in range(K):
        for n in range(K):
            out += ker[:, m, n][:, None, None] * x[m:m + stride * OH:stride, n:n + stride * OW:stride][None]
    half = 1 << (WIDTH - 1)
    return ((out + half) % (1 << WIDTH)) - half
