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

from .circuits import gate_level, mk_add_3gen, mk_grt_3gen, mk_leq_3gen, mk_sub_3gen
from .tfhe3gen import MKLweSample, encode_message, mk_lwe_noiseless_trivial


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
