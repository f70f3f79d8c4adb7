"""Integer circuits of the reference (3-gen-mk-tfhe/src/3gen_mk_gates.jl:183-362) as level-batched launches.

The reference evaluates every gate as a scalar call (5·WIDTH sequential bootstraps for an adder).  Here each operand bit is
an `MKLweSample` that may carry a batch of independent circuit INSTANCES (leading dimensions), and every dependency level
of the circuit is ONE mixed-gate launch (`mktfhe_gate_batch_mixed`): level 0 of the ripple adder is all XOR(a_i, b_i) and
AND(a_i, b_i) (2·WIDTH·I gates), then two levels per bit.  Same names, argument order and results as the reference,
including its quirks (`mk_int_mul_3gen` reuses row WIDTH-1 in its last adder, 3gen_mk_gates.jl:336-350).
"""
import numpy as np

from . import _cabi
from .tfhe3gen import MKLweSample, MKLweSampleGPU, engine_for, mk_copy_3gen

_G = {"nand": _cabi.GATE_NAND, "or": _cabi.GATE_OR, "and": _cabi.GATE_AND, "xor": _cabi.GATE_XOR}


def gate_level(bk, ks, jobs):
    """Evaluate a list of independent gates [(kind, x, y), ...] (kind in nand/or/and/xor; x, y of equal batch shape per job)
    in ONE launch; returns the list of outputs in order."""
    eng = engine_for(bk, ks)
    k, n = eng.params.max_parties, eng.params.lwe_size
    if isinstance(jobs[0][1], MKLweSampleGPU):
        return _gate_level_gpu(eng, jobs, k, n)
    xa, xb, ya, yb, ids, shapes = [], [], [], [], [], []
    for kind, x, y in jobs:
        if x.b.shape != y.b.shape:
            x, y = _broadcast(x, y)
        shapes.append(x.b.shape)
        xa.append(x.a.reshape(-1, k, n)); xb.append(x.b.reshape(-1))
        ya.append(y.a.reshape(-1, k, n)); yb.append(y.b.reshape(-1))
        ids.append(np.full(xb[-1].size, _G[kind], np.int32))
    oa, ob = eng.ctx.gate_batch_mixed(np.concatenate(ids), (np.concatenate(xa), np.concatenate(xb)),
                                      (np.concatenate(ya), np.concatenate(yb)))
    outs, pos = [], 0
    params = jobs[0][1].params
    for shp in shapes:
        cnt = int(np.prod(shp, dtype=np.int64))
        outs.append(MKLweSample(params, oa[pos:pos + cnt].reshape(shp + (k, n)), ob[pos:pos + cnt].reshape(shp), 0.0))
        pos += cnt
    return outs


def _gate_level_gpu(eng, jobs, k, n):
    """Device-resident level: operands are gathered with torch.cat in HBM and the level is one mktfhe_gate_batch_mixed_dev call."""
    import torch
    xa, xb, ya, yb, ids, shapes = [], [], [], [], [], []
    dev = jobs[0][1].a.device
    if any(t.a.device != dev or t.b.device != dev for _, x, y in jobs for t in (x, y)):
        raise ValueError("the operands of a gate level live on different devices")
    ctx = eng.ctx_on(dev.index)                 # raises when the engine holds no key replica on that GPU
    for kind, x, y in jobs:
        shp = torch.broadcast_shapes(tuple(x.b.shape), tuple(y.b.shape))
        shapes.append(tuple(shp))
        xa.append(x.a.expand(shp + (k, n)).reshape(-1, k, n)); xb.append(x.b.expand(shp).reshape(-1))
        ya.append(y.a.expand(shp + (k, n)).reshape(-1, k, n)); yb.append(y.b.expand(shp).reshape(-1))
        ids.append(torch.full((xb[-1].numel(),), _G[kind], dtype=torch.int32, device=dev))
    xa, xb, ya, yb, ids = (torch.cat(t).contiguous() for t in (xa, xb, ya, yb, ids))
    G = xb.numel()
    oa = torch.empty((G, k, n), dtype=torch.int32, device=dev)
    ob = torch.empty(G, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    if not stream:                      # legacy default stream: the context's own stream does not wait for it -- operands must be complete
        torch.cuda.synchronize(dev)
    ctx.gate_batch_mixed_dev(G, ids.data_ptr(), xa.data_ptr(), xb.data_ptr(), ya.data_ptr(), yb.data_ptr(), 0, 0, oa.data_ptr(), ob.data_ptr(),
                             stream=stream)
    if not stream:
        torch.cuda.synchronize(dev)
    outs, pos = [], 0
    for shp in shapes:
        cnt = int(np.prod(shp, dtype=np.int64))
        outs.append(MKLweSampleGPU(jobs[0][1].params, oa[pos:pos + cnt].reshape(shp + (k, n)), ob[pos:pos + cnt].reshape(shp), 0.0))
        pos += cnt
    return outs


def mk_int_add_3gen_gpu(bk, ks, a, b, Cin, WIDTH):
    """3gen_mk_gates.jl:447-463: the adder on device-resident samples (the reference's version allocates a CuArray of host structs)."""
    return mk_add_3gen(bk, ks, a, b, Cin, WIDTH)


def _broadcast(x, y):
    shp = np.broadcast_shapes(x.b.shape, y.b.shape)
    kn = x.a.shape[-2:]
    bx = MKLweSample(x.params, np.broadcast_to(x.a, shp + kn), np.broadcast_to(x.b, shp), x.current_variance)
    by = MKLweSample(y.params, np.broadcast_to(y.a, shp + kn), np.broadcast_to(y.b, shp), y.current_variance)
    return bx, by


def mk_int_add_with_carry_3gen(bk, ks, a, b, Cin, WIDTH):
    """3gen_mk_gates.jl:291-310: WIDTH sum bits plus the final carry (WIDTH + 1 samples)."""
    lvl0 = gate_level(bk, ks, [("xor", a[i], b[i]) for i in range(WIDTH)] + [("and", a[i], b[i]) for i in range(WIDTH)])
    tmp1, tmp2 = lvl0[:WIDTH], lvl0[WIDTH:]
    result, cin = [], Cin
    for i in range(WIDTH):
        s, tmp3 = gate_level(bk, ks, [("xor", tmp1[i], cin), ("and", tmp1[i], cin)])
        (cin,) = gate_level(bk, ks, [("or", tmp2[i], tmp3)])
        result.append(s)
    return result + [cin]


def mk_add_3gen(bk, ks, a, b, Cin, WIDTH):
    """3gen_mk_gates.jl:183-200."""
    return mk_int_add_with_carry_3gen(bk, ks, a, b, Cin, WIDTH)[:WIDTH]


mk_add_3gen_v2 = mk_add_3gen   # identical in the reference (3gen_mk_gates.jl:203-220)


def mk_inv_3gen(bk, ks, a, one, WIDTH):
    """3gen_mk_gates.jl:223-233: bitwise NOT as XOR with an encryption of 1, one launch."""
    return gate_level(bk, ks, [("xor", a[i], one) for i in range(WIDTH)])


def mk_sub_3gen(bk, ks, a, b, one, WIDTH):
    """3gen_mk_gates.jl:236-244: a + ~b + 1."""
    return mk_add_3gen(bk, ks, a, mk_inv_3gen(bk, ks, b, one, WIDTH), one, WIDTH)


def mk_less_3gen(bk, ks, a, b, one, WIDTH):
    """3gen_mk_gates.jl:247-255: sign bit of a - b."""
    return mk_copy_3gen(mk_sub_3gen(bk, ks, a, b, one, WIDTH)[WIDTH - 1])


def mk_grt_3gen(bk, ks, a, b, one, WIDTH):
    """3gen_mk_gates.jl:258-266."""
    return mk_copy_3gen(mk_sub_3gen(bk, ks, b, a, one, WIDTH)[WIDTH - 1])


def mk_leq_3gen(bk, ks, a, b, one, WIDTH):
    """3gen_mk_gates.jl:269-277."""
    return gate_level(bk, ks, [("xor", mk_grt_3gen(bk, ks, a, b, one, WIDTH), one)])[0]


def mk_geq_3gen(bk, ks, a, b, one, WIDTH):
    """3gen_mk_gates.jl:280-288."""
    return gate_level(bk, ks, [("xor", mk_less_3gen(bk, ks, a, b, one, WIDTH), one)])[0]


def mk_int_mul_3gen(bk, ks, a, b, ZERO, WIDTH):
    """3gen_mk_gates.jl:312-362, transcribed with 0-based indices.  All WIDTH^2 partial products are one launch.  As in the
    reference, the final adder re-uses partial-product row `ctr` (the last row already added, or row 1 when WIDTH = 2)
    instead of row WIDTH, so the result is not a*b in general -- replicated, not fixed (tests compare against a plaintext
    model of the same wiring)."""
    prods = gate_level(bk, ks, [("and", a[j], b[i]) for i in range(WIDTH) for j in range(WIDTH)])
    BArr = [[prods[i * WIDTH + j] for j in range(WIDTH)] for i in range(WIDTH)]
    result = [None] * (2 * WIDTH + 1)
    result[0] = mk_copy_3gen(BArr[0][0])
    tmpIn = [mk_copy_3gen(BArr[0][i + 1]) for i in range(WIDTH - 1)] + [mk_copy_3gen(ZERO)]
    ctr = 1                                   # 1-based, as in the reference
    for i in range(2, WIDTH):                 # for i = 2:WIDTH-1
        tmpArr = mk_int_add_with_carry_3gen(bk, ks, tmpIn, BArr[i - 1], ZERO, WIDTH)
        result[i - 1] = mk_copy_3gen(tmpArr[0])
        tmpIn = [mk_copy_3gen(tmpArr[j + 1]) for j in range(WIDTH)]
        ctr = i
    tmpArr = mk_int_add_with_carry_3gen(bk, ks, tmpIn, BArr[ctr - 1], ZERO, WIDTH)
    for i in range(WIDTH + 1):
        result[i + ctr] = mk_copy_3gen(tmpArr[i])
    return [mk_copy_3gen(result[i]) for i in range(WIDTH)]


def mk_int_add_3gen(bk, ks, a, b, Cin, WIDTH):
    """The adder `enc_conv2d` calls (3gen_mk_gates.jl:389; its definition at :158-180 is commented out in the reference): same
    ripple-carry circuit as mk_add_3gen."""
    return mk_add_3gen(bk, ks, a, b, Cin, WIDTH)


def add_mod_3gen(bk, ks, a, b, w):
    """a + b mod 2^w on w-bit operands: the ripple adder of 3gen_mk_gates.jl:291-310 without the carry-in (bit 0 is a half adder)
    and without the gates that only feed the dropped carry-out: 5w - 6 gates (1 for w = 1) on 2w - 2 levels."""
    if w == 1:
        return gate_level(bk, ks, [("xor", a[0], b[0])])
    lvl0 = gate_level(bk, ks, [("xor", a[i], b[i]) for i in range(w)] + [("and", a[i], b[i]) for i in range(w - 1)])
    t1, t2 = lvl0[:w], lvl0[w:]
    result, cin = [t1[0]], t2[0]
    for i in range(1, w - 1):
        s, t3 = gate_level(bk, ks, [("xor", t1[i], cin), ("and", t1[i], cin)])
        (cin,) = gate_level(bk, ks, [("or", t2[i], t3)])
        result.append(s)
    result += gate_level(bk, ks, [("xor", t1[w - 1], cin)])
    return result


def mk_int_mul_lo_3gen(bk, ks, a, b, WIDTH):
    """Low WIDTH bits of a * b (two's-complement wrap, i.e. the product of WIDTH-bit ints mod 2^WIDTH): the shift-and-add array the
    reference's mk_int_mul_3gen (3gen_mk_gates.jl:312-362) aims at, with its last-row slip corrected and the partial products
    that cannot reach the kept bits left out.  WIDTH (WIDTH + 1) / 2 ANDs in one launch, then WIDTH - 1 shrinking adders."""
    prods = gate_level(bk, ks, [("and", a[j], b[i]) for i in range(WIDTH) for j in range(WIDTH - i)])
    rows, pos = [], 0
    for i in range(WIDTH):
        rows.append(prods[pos:pos + WIDTH - i]); pos += WIDTH - i
    acc = list(rows[0])
    for i in range(1, WIDTH):
        acc[i:] = add_mod_3gen(bk, ks, acc[i:], rows[i], WIDTH - i)
    return acc
