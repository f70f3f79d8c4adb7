# TFHE_B200.jl -- `ccall` shim that puts libmktfhe_b200.so behind the reference's 3gen gate API.
#
# NOT EXECUTED IN THIS REPO'S CI: the build image has no Julia.  The executable twin of this file is
# torus-fhe_b200/tfhe3gen.py (ctypes, same C ABI, same names), which the tests and benchmarks drive.
#
# Usage inside the reference tree (3-gen-mk-tfhe/): `include("src/TFHE.jl"); include("TFHE_B200.jl"); using .TFHE_B200`
# after key generation replace
#     bk_keys = [TransformedBootstrapKeyPart_3gen(bk_keys[i]) for i in 1:parties]      # multikey_3gen.jl:28
# by
#     bk_keys = [TFHE_B200.TransformedBootstrapKeyPart_3gen(bk_keys[i]) for i in 1:parties]
# and call TFHE_B200.mk_gate_nand_3gen(bk_keys, ks_keys, x, y) etc.  Signatures are those of
# 3-gen-mk-tfhe/src/3gen_mk_gates.jl:8-150 and src/3gen_mk_internals.jl:112-116; vector methods batch.
module TFHE_B200

using ..TFHE: TGswParams, RLweParams, BootstrapKeyPart_3gen, KeyswitchKey, MKLweSample, LweSample, LweParams, RLweSample,
              encode_message, encode_message64, decode_message, mk_lwe_noiseless_trivial, torus_polynomial, mul_by_monomial

const LIB = get(ENV, "MKTFHE_B200_LIB", "libmktfhe_b200")

struct CParams            # mktfhe_params (include/mktfhe_b200.h)
    n::Int32; N::Int32; k::Int32; l::Int32; bgbit::Int32; t::Int32; basebit::Int32; reserved::Int32
end

const GATE_NAND, GATE_OR, GATE_AND, GATE_XOR, GATE_AND3 = Cint(0), Cint(1), Cint(2), Cint(3), Cint(4)

check(ctx, rc) = rc == 0 ? nothing :
    error("libmktfhe_b200 error $rc: " * unsafe_string(ccall((:mktfhe_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx)))

"""Same name, constructor arity and fields as 3gen_mk_internals.jl:45-55, but it keeps the INTEGER polynomials
(the reference keeps only Complex{Float64} FFTs, from which the exact key cannot be recovered)."""
struct TransformedBootstrapKeyPart_3gen
    tgsw_params :: TGswParams
    rlwe_params :: RLweParams
    gsw_key :: Array{Int64, 4}        # (N, l, 4, n) column-major == C int64 [n][4][l][N]
    key_size :: Int
    function TransformedBootstrapKeyPart_3gen(bk::BootstrapKeyPart_3gen)
        n, l = bk.key_size, bk.tgsw_params.decomp_length
        N = bk.rlwe_params.polynomial_degree
        g = Array{Int64, 4}(undef, N, l, 4, n)
        for j in 1:n, q in 1:l
            s = bk.gsw_key[j]
            g[:, q, 1, j] = s.part_1[q].coeffs
            g[:, q, 2, j] = s.part_2[q].coeffs
            g[:, q, 3, j] = s.part_3[q].coeffs
            g[:, q, 4, j] = s.part_4[q].coeffs
        end
        new(bk.tgsw_params, bk.rlwe_params, g, n)
    end
end

mutable struct Engine
    ctx :: Ptr{Cvoid}
    prm :: CParams
end

const ENGINES = IdDict{Any, Engine}()          # keyed by the bk array
const PART_OWNER = IdDict{Any, Tuple{Engine, Int}}()   # key part -> (engine holding it, 0-based party): the per-element entry points
const ENGINES_BY_KS = IdDict{Any, Engine}()    # the same engines keyed by the ks array (mk_keyswitch_3gen receives ks only)

"""KeyswitchKey.key :: Array{LweSample,3} of dims (base-1, t, N) (keyswitch.jl:7-42) -> Int32 (n+1, base-1, t, N),
i.e. C int32 [N][t][base-1][n+1]."""
function flatten_ksk(ks::KeyswitchKey)
    B1, t, N = size(ks.key)
    n = ks.out_lwe_params.size
    rows = Array{Int32, 4}(undef, n + 1, B1, t, N)
    for i in 1:N, j in 1:t, h in 1:B1
        rows[1:n, h, j, i] = ks.key[h, j, i].a
        rows[n + 1, h, j, i] = ks.key[h, j, i].b
    end
    rows
end

"""GPUs an engine spans when the caller does not say: ENV["MKTFHE_B200_DEVICES"] = "all" | "0,1,2,3" | unset (GPU 0 only)."""
function default_devices()
    v = get(ENV, "MKTFHE_B200_DEVICES", "")
    v == "" && return Cint[0]
    v == "all" && return Cint[]                  # empty = every visible GPU (n_devices = 0)
    Cint[parse(Cint, x) for x in split(v, ",")]
end

"""The engine holding (bk, ks): created on first use.  `devices` lists the GPUs it spans (one context behind the whole gate
API: keys are loaded once and broadcast GPU to GPU inside mktfhe_finalize_keys, every batched gate call is sharded over
them by the library); an empty list means every visible GPU."""
function engine_for(bk::Array{TransformedBootstrapKeyPart_3gen, 1}, ks::Array{KeyswitchKey, 1}; devices::Vector{Cint} = default_devices())
    haskey(ENGINES, bk) && return ENGINES[bk]
    bk[1].rlwe_params.is32 && error("rlwe_is32 = true is not part of the 3gen path")
    prm = CParams(bk[1].key_size, bk[1].rlwe_params.polynomial_degree, length(bk), bk[1].tgsw_params.decomp_length,
                  bk[1].tgsw_params.log2_base, ks[1].params.decomp_length, ks[1].params.log2_base, 0)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:mktfhe_create_multi, LIB), Cint, (Ref{CParams}, Cint, Ptr{Cint}, Ref{Ptr{Cvoid}}), prm, length(devices),
               isempty(devices) ? Ptr{Cint}(C_NULL) : pointer(devices), out)
    rc == 0 || error("mktfhe_create_multi: " * unsafe_string(ccall((:mktfhe_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    ctx = out[]
    for p in 1:length(bk)
        check(ctx, ccall((:mktfhe_load_bsk, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Int64}), ctx, p - 1, bk[p].gsw_key))
        check(ctx, ccall((:mktfhe_load_ksk, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Int32}), ctx, p - 1, flatten_ksk(ks[p])))
    end
    check(ctx, ccall((:mktfhe_finalize_keys, LIB), Cint, (Ptr{Cvoid},), ctx))      # includes the key broadcast to the other GPUs
    e = Engine(ctx, prm)
    for p in 1:length(bk); PART_OWNER[bk[p].gsw_key] = (e, p - 1); end
    ENGINES_BY_KS[ks] = e
    ENGINES[bk] = e
end

"""GPUs the engine of (bk, ks) spans."""
device_count(bk, ks) = Int(ccall((:mktfhe_device_count, LIB), Cint, (Ptr{Cvoid},), engine_for(bk, ks).ctx))

"""Frees the device key replicas of (bk, ks) (about 200 MB per GPU at 2 parties, gigabytes for the N = 2048 sets).  The lookup
tables above hold every engine strongly -- the gate API must find it from `bk` alone, at any time -- so engines live until
released here."""
function release!(bk::Array{TransformedBootstrapKeyPart_3gen, 1}, ks::Array{KeyswitchKey, 1})
    haskey(ENGINES, bk) || return nothing
    e = ENGINES[bk]
    delete!(ENGINES, bk); delete!(ENGINES_BY_KS, ks)
    for p in 1:length(bk); delete!(PART_OWNER, bk[p].gsw_key); end
    e.ctx == C_NULL || ccall((:mktfhe_destroy, LIB), Cvoid, (Ptr{Cvoid},), e.ctx)
    e.ctx = C_NULL
    nothing
end

# entry points of the reference that receive only one of (bk, ks): the GPU context holds both, so the pair must have been seen before
cached_engine(d::IdDict, keys, what) = haskey(d, keys) ? d[keys] :
    error("no GPU engine holds these $what yet: call TFHE_B200.engine_for(bk, ks) (or any gate) with the key pair first")

# MKLweSample.a is (n, parties) column-major == C int32 [k][n]; a batch is the samples back to back
pack_a(xs::Vector{MKLweSample}) = reduce(hcat, [vec(x.a) for x in xs])     # (k*n, G)
pack_b(xs::Vector{MKLweSample}) = Int32[x.b for x in xs]

function unpack(params::LweParams, n, k, oa::Matrix{Int32}, ob::Vector{Int32})
    [MKLweSample(params, reshape(oa[:, g], n, k), ob[g], 0.0) for g in 1:length(ob)]   # variance 0.0: mk_internals.jl:738-743
end

"""mk_bootstrap_3gen(bk, ks, mu, x) (3gen_mk_internals.jl:112-116); `x` may be a vector (one launch for the batch)."""
function mk_bootstrap_3gen(bk::Array{TransformedBootstrapKeyPart_3gen, 1}, ks::Array{KeyswitchKey, 1},
                           mu::Union{Int32, Int64}, xs::Vector{MKLweSample})
    e = engine_for(bk, ks); n, k, G = Int(e.prm.n), Int(e.prm.k), length(xs)
    a, b = pack_a(xs), pack_b(xs)
    oa, ob = Matrix{Int32}(undef, n * k, G), Vector{Int32}(undef, G)
    check(e.ctx, ccall((:mktfhe_bootstrap_batch, LIB), Cint,
        (Ptr{Cvoid}, Int64, Csize_t, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}), e.ctx, Int64(mu), G, a, b, oa, ob))
    unpack(xs[1].params, n, k, oa, ob)
end
mk_bootstrap_3gen(bk::Array{TransformedBootstrapKeyPart_3gen, 1}, ks, mu, x::MKLweSample) = mk_bootstrap_3gen(bk, ks, mu, [x])[1]

function blind_rotate_batch(e::Engine, mu::Int64, a::Matrix{Int32}, b::Vector{Int32})
    N, G = Int(e.prm.N), length(b)
    ext = Matrix{Int32}(undef, N + 1, G)                      # C int32 [G][N+1]: a'[0..N-1], b'
    check(e.ctx, ccall((:mktfhe_blind_rotate_batch, LIB), Cint,
        (Ptr{Cvoid}, Int64, Csize_t, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int64}), e.ctx, mu, G, a, b, ext, Ptr{Int64}(C_NULL)))
    [LweSample(LweParams(N), ext[1:N, g], ext[N + 1, g], 0.0) for g in 1:G]
end

"""mk_bootstrap_wo_keyswitch_3gen(bk, mu, x) (3gen_mk_internals.jl:99-109): the extracted LweSample of dimension N."""
function mk_bootstrap_wo_keyswitch_3gen(bk::Array{TransformedBootstrapKeyPart_3gen, 1}, mu::Union{Int32, Int64}, xs::Vector{MKLweSample})
    blind_rotate_batch(cached_engine(ENGINES, bk, "bootstrapping keys"), Int64(mu), pack_a(xs), pack_b(xs))
end
mk_bootstrap_wo_keyswitch_3gen(bk::Array{TransformedBootstrapKeyPart_3gen, 1}, mu::Union{Int32, Int64}, x::MKLweSample) =
    mk_bootstrap_wo_keyswitch_3gen(bk, mu, [x])[1]

"""mk_blind_rotate_and_extract_3gen(v, bk, barb, bara) (3gen_mk_internals.jl:88-95).  `v` must be the constant test vector
repeat([mu], N), the only one the reference builds (:105-108).  The kernel mod-switches its input itself, so it is handed the
torus elements bar * 2^32 / 2N, which decode_message(., 2N) maps back to bar."""
function mk_blind_rotate_and_extract_3gen(v, bk::Array{TransformedBootstrapKeyPart_3gen, 1}, barb::Int32, bara::Array{Int32, 2})
    e = cached_engine(ENGINES, bk, "bootstrapping keys")
    N = Int(e.prm.N)
    c = v.coeffs
    (length(c) == N && all(==(c[1]), c)) || error("the engine rotates the constant test vector repeat([mu], N) only")
    size(bara) == (Int(e.prm.n), Int(e.prm.k)) || error("bara must be (n, parties)")
    sh = 32 - (trailing_zeros(N) + 1)
    blind_rotate_batch(e, Int64(c[1]), reshape(bara .<< sh, :, 1), Int32[barb << sh])[1]
end

"""mk_keyswitch_3gen(ks, sample) (mk_internals.jl:730-744); `sample` may be a vector of extracted LweSamples."""
function mk_keyswitch_3gen(ks::Array{KeyswitchKey, 1}, us::Vector{LweSample})
    e = cached_engine(ENGINES_BY_KS, ks, "key-switching keys"); n, k, N, G = Int(e.prm.n), Int(e.prm.k), Int(e.prm.N), length(us)
    ext = Matrix{Int32}(undef, N + 1, G)
    for g in 1:G
        length(us[g].a) == N || error("mk_keyswitch_3gen expects LweSamples of dimension N = $N")
        ext[1:N, g] = us[g].a; ext[N + 1, g] = us[g].b
    end
    oa, ob = Matrix{Int32}(undef, n * k, G), Vector{Int32}(undef, G)
    check(e.ctx, ccall((:mktfhe_keyswitch_batch, LIB), Cint, (Ptr{Cvoid}, Csize_t, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}), e.ctx, G, ext, oa, ob))
    unpack(ks[1].out_lwe_params, n, k, oa, ob)
end
mk_keyswitch_3gen(ks::Array{KeyswitchKey, 1}, u::LweSample) = mk_keyswitch_3gen(ks, [u])[1]

# ---- the path's internal entry points, stage by stage (3gen_mk_internals.jl:59-84, tgsw_3gen.jl:102-113); the tested twin is
# torus-fhe_b200/tfhe3gen.py.  Gates and bootstraps never take this route: they run the whole loop inside one kernel.
"""Element j (1-based) of a party's bootstrapping key, resident on the GPU (the reference's type holds its FFTs, tgsw_3gen.jl:23-39)."""
struct TransformedTGswSample_3gen
    part :: TransformedBootstrapKeyPart_3gen
    j :: Int
end
tgsw_samples(bk::TransformedBootstrapKeyPart_3gen) = [TransformedTGswSample_3gen(bk, j) for j in 1:bk.key_size]

"""tgsw_extern_mul_3gen(accum, sample) (tgsw_3gen.jl:102-113) through mktfhe_extprod_batch: accum.a = [mask, body]."""
function tgsw_extern_mul_3gen(accum::RLweSample, sample::TransformedTGswSample_3gen)
    haskey(PART_OWNER, sample.part.gsw_key) || error("no GPU engine holds this key part yet: call TFHE_B200.engine_for(bk, ks) first")
    e, party = PART_OWNER[sample.part.gsw_key]
    N = Int(e.prm.N)
    acc = hcat(Vector{Int64}(accum.a[1].coeffs), Vector{Int64}(accum.a[2].coeffs))      # (N, 2) column-major == C int64 [2][N]
    out = Matrix{Int64}(undef, N, 2)
    elem = Int32[party * Int(e.prm.n) + sample.j - 1]
    check(e.ctx, ccall((:mktfhe_extprod_batch, LIB), Cint, (Ptr{Cvoid}, Csize_t, Ptr{Int32}, Ptr{Int64}, Ptr{Int64}), e.ctx, 1, elem, acc, out))
    RLweSample(accum.params, [torus_polynomial(out[:, 1]), torus_polynomial(out[:, 2])], 0.0)
end

mk_mux_rotate_3gen(accum::RLweSample, bki::TransformedTGswSample_3gen, barai::Int32) =                  # :59-62
    accum + tgsw_extern_mul_3gen(mul_by_monomial(accum, barai) - accum, bki)

function mk_ith_blind_rotate_3gen(acc::RLweSample, gsw_key::Array{TransformedTGswSample_3gen, 1}, bara::Array{Int32, 1})   # :66-75
    for i in eachindex(bara)
        if bara[i] != 0
            acc = mk_mux_rotate_3gen(acc, gsw_key[i], bara[i])
        end
    end
    acc
end

function mk_blind_rotate_3gen(accum::RLweSample, bk::Array{TransformedBootstrapKeyPart_3gen, 1}, bara::Array{Int32, 2})     # :78-84
    for i in 1:length(bk)
        accum = mk_ith_blind_rotate_3gen(accum, tgsw_samples(bk[i]), bara[:, i])
    end
    accum
end

function gate_batch(bk, ks, gate::Cint, xs::Vector{MKLweSample}, ys::Vector{MKLweSample}, zs::Union{Nothing, Vector{MKLweSample}} = nothing)
    e = engine_for(bk, ks); n, k, G = Int(e.prm.n), Int(e.prm.k), length(xs)
    length(ys) == G || error("gate operands have different batch sizes")
    xa, xb, ya, yb = pack_a(xs), pack_b(xs), pack_a(ys), pack_b(ys)
    za = zs === nothing ? Ptr{Int32}(C_NULL) : pack_a(zs)
    zb = zs === nothing ? Ptr{Int32}(C_NULL) : pack_b(zs)
    oa, ob = Matrix{Int32}(undef, n * k, G), Vector{Int32}(undef, G)
    check(e.ctx, ccall((:mktfhe_gate_batch, LIB), Cint,
        (Ptr{Cvoid}, Cint, Csize_t, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
        e.ctx, gate, G, xa, xb, ya, yb, za, zb, oa, ob))
    unpack(xs[1].params, n, k, oa, ob)
end

const BK = Array{TransformedBootstrapKeyPart_3gen, 1}
const KS = Array{KeyswitchKey, 1}
const V = Vector{MKLweSample}

# scalar and batched gates, 3gen_mk_gates.jl:8-150
for (fname, gid) in ((:mk_gate_nand_3gen, GATE_NAND), (:mk_gate_or_3gen, GATE_OR), (:mk_gate_and_3gen, GATE_AND), (:mk_gate_xor_3gen, GATE_XOR))
    @eval $fname(bk::BK, ks::KS, x::V, y::V) = gate_batch(bk, ks, $gid, x, y)
    @eval $fname(bk::BK, ks::KS, x::MKLweSample, y::MKLweSample) = gate_batch(bk, ks, $gid, [x], [y])[1]
end
mk_gate_3and_3gen(bk::BK, ks::KS, x::V, y::V, z::V) = gate_batch(bk, ks, GATE_AND3, x, y, z)
mk_gate_3and_3gen(bk::BK, ks::KS, x::MKLweSample, y::MKLweSample, z::MKLweSample) = gate_batch(bk, ks, GATE_AND3, [x], [y], [z])[1]
mk_gate_not_3gen(x::MKLweSample) = -x
# the `_wb` variants (3gen_mk_gates.jl:16-21, 32-37, 48-53, 76-81): the linear prologue alone, not bootstrapped -- plain Julia
mk_gate_nand_3gen_wb(bk::BK, ks::KS, x::MKLweSample, y::MKLweSample) = mk_lwe_noiseless_trivial(encode_message(1, 8), x.params, length(bk)) - x - y
mk_gate_or_3gen_wb(bk::BK, ks::KS, x::MKLweSample, y::MKLweSample) = mk_lwe_noiseless_trivial(encode_message(1, 8), x.params, length(bk)) + x + y
mk_gate_and_3gen_wb(bk::BK, ks::KS, x::MKLweSample, y::MKLweSample) = mk_lwe_noiseless_trivial(encode_message(-1, 8), x.params, length(bk)) + x + y
mk_gate_xor_3gen_wb(bk::BK, ks::KS, x::MKLweSample, y::MKLweSample) =
    mk_lwe_noiseless_trivial(encode_message(1, 4), x.params, length(bk)) + Int32(2) * x + Int32(2) * y

"""3gen_mk_gates.jl:133-150: the two ANDs in one launch, then 1/8 + t1 + t2 NOT bootstrapped (as the reference)."""
function mk_gate_mux_3gen(bk::BK, ks::KS, x::MKLweSample, y::MKLweSample, z::MKLweSample)
    t = gate_batch(bk, ks, GATE_AND, [x, -x], [y, z])
    mk_lwe_noiseless_trivial(encode_message(1, 8), t[1].params, length(bk)) + t[1] + t[2]
end

# ---- level-batched circuits, 3gen_mk_gates.jl:183-310 (torus-fhe_b200/circuits.py is the tested twin): every dependency level of a
# circuit is ONE mixed-gate launch (mktfhe_gate_batch_mixed) instead of one launch per gate.  `jobs` = [(gate_id, x, y), ...]
function gate_level(bk::BK, ks::KS, jobs::Vector{Tuple{Cint, MKLweSample, MKLweSample}})
    e = engine_for(bk, ks); n, k, G = Int(e.prm.n), Int(e.prm.k), length(jobs)
    ids = Int32[j[1] for j in jobs]
    xs, ys = MKLweSample[j[2] for j in jobs], MKLweSample[j[3] for j in jobs]
    oa, ob = Matrix{Int32}(undef, n * k, G), Vector{Int32}(undef, G)
    check(e.ctx, ccall((:mktfhe_gate_batch_mixed, LIB), Cint,
        (Ptr{Cvoid}, Csize_t, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
        e.ctx, G, ids, pack_a(xs), pack_b(xs), pack_a(ys), pack_b(ys), Ptr{Int32}(C_NULL), Ptr{Int32}(C_NULL), oa, ob))
    unpack(xs[1].params, n, k, oa, ob)
end

mk_copy_3gen(x::MKLweSample) = MKLweSample(x.params, copy(x.a), x.b, x.current_variance)

"""3gen_mk_gates.jl:291-310: WIDTH sum bits plus the final carry; 1 + 2 WIDTH launches instead of 5 WIDTH."""
function mk_int_add_with_carry_3gen(bk::BK, ks::KS, a::Vector{MKLweSample}, b::Vector{MKLweSample}, Cin::MKLweSample, WIDTH)
    lvl0 = gate_level(bk, ks, vcat([(GATE_XOR, a[i], b[i]) for i in 1:WIDTH], [(GATE_AND, a[i], b[i]) for i in 1:WIDTH]))
    tmp1, tmp2 = lvl0[1:WIDTH], lvl0[WIDTH+1:2*WIDTH]
    result = Vector{MKLweSample}(undef, WIDTH + 1)
    cin = Cin
    for i in 1:WIDTH
        st = gate_level(bk, ks, [(GATE_XOR, tmp1[i], cin), (GATE_AND, tmp1[i], cin)])
        result[i] = st[1]
        cin = gate_level(bk, ks, [(GATE_OR, tmp2[i], st[2])])[1]
    end
    result[WIDTH + 1] = cin
    result
end
mk_add_3gen(bk::BK, ks::KS, a, b, Cin::MKLweSample, WIDTH) = mk_int_add_with_carry_3gen(bk, ks, a, b, Cin, WIDTH)[1:WIDTH]      # :183-200
mk_add_3gen_v2(bk::BK, ks::KS, a, b, Cin::MKLweSample, WIDTH) = mk_add_3gen(bk, ks, a, b, Cin, WIDTH)                            # :203-220
mk_inv_3gen(bk::BK, ks::KS, a, one::MKLweSample, WIDTH) = gate_level(bk, ks, [(GATE_XOR, a[i], one) for i in 1:WIDTH])           # :223-233
mk_sub_3gen(bk::BK, ks::KS, a, b, one::MKLweSample, WIDTH) = mk_add_3gen(bk, ks, a, mk_inv_3gen(bk, ks, b, one, WIDTH), one, WIDTH)   # :236-244
mk_less_3gen(bk::BK, ks::KS, a, b, one::MKLweSample, WIDTH) = mk_copy_3gen(mk_sub_3gen(bk, ks, a, b, one, WIDTH)[WIDTH])          # :247-255
mk_grt_3gen(bk::BK, ks::KS, a, b, one::MKLweSample, WIDTH) = mk_copy_3gen(mk_sub_3gen(bk, ks, b, a, one, WIDTH)[WIDTH])           # :258-266
mk_leq_3gen(bk::BK, ks::KS, a, b, one::MKLweSample, WIDTH) = mk_gate_xor_3gen(bk, ks, mk_grt_3gen(bk, ks, a, b, one, WIDTH), one)   # :269-277
mk_geq_3gen(bk::BK, ks::KS, a, b, one::MKLweSample, WIDTH) = mk_gate_xor_3gen(bk, ks, mk_less_3gen(bk, ks, a, b, one, WIDTH), one)  # :280-288

"""3gen_mk_gates.jl:312-362, statement for statement on top of the batched levels: all WIDTH^2 partial products are one launch.
As in the reference the final adder re-uses row `ctr` (the last row already added) instead of row WIDTH -- replicated, not fixed."""
function mk_int_mul_3gen(bk::BK, ks::KS, a::Vector{MKLweSample}, b::Vector{MKLweSample}, ZERO::MKLweSample, WIDTH)
    prods = gate_level(bk, ks, [(GATE_AND, a[j], b[i]) for i in 1:WIDTH for j in 1:WIDTH])
    BArr = [prods[(i - 1) * WIDTH + j] for i in 1:WIDTH, j in 1:WIDTH]
    result = Vector{MKLweSample}(undef, 2 * WIDTH + 1)
    result[1] = mk_copy_3gen(BArr[1, 1])
    tmpIn = MKLweSample[[mk_copy_3gen(BArr[1, i + 1]) for i in 1:WIDTH-1]; mk_copy_3gen(ZERO)]
    ctr = 1
    for i in 2:WIDTH-1
        tmpArr = mk_int_add_with_carry_3gen(bk, ks, tmpIn, BArr[i, :], ZERO, WIDTH)
        result[i] = mk_copy_3gen(tmpArr[1])
        tmpIn = MKLweSample[mk_copy_3gen(tmpArr[j + 1]) for j in 1:WIDTH]
        ctr = i
    end
    tmpArr = mk_int_add_with_carry_3gen(bk, ks, tmpIn, BArr[ctr, :], ZERO, WIDTH)
    for i in 1:WIDTH+1
        result[i + ctr] = mk_copy_3gen(tmpArr[i])
    end
    MKLweSample[mk_copy_3gen(result[i]) for i in 1:WIDTH]
end

# ---- interchange files (torus-fhe_b200/interchange.py documents the layout): dump the reference's own keys / ciphertexts so that
# the B200 engine and its CPU oracle can be checked on the very bytes the Julia code produced
function write_keys(path::String, params, bk::Array{TransformedBootstrapKeyPart_3gen, 1}, ks::Array{KeyswitchKey, 1}, secret_keys = nothing)
    open(path, "w") do f
        write(f, "MKTFHE3K"); write(f, UInt32(1))
        for v in (bk[1].key_size, bk[1].rlwe_params.polynomial_degree, length(bk), bk[1].tgsw_params.decomp_length, bk[1].tgsw_params.log2_base,
                  ks[1].params.decomp_length, ks[1].params.log2_base, secret_keys === nothing ? 0 : 1)
            write(f, Int32(v))
        end
        write(f, Float64(params.lwe_noise_stddev)); write(f, Float64(params.gsw_noise_stddev)); write(f, Float64(params.ks_noise_stddev))
        for p in 1:length(bk); write(f, bk[p].gsw_key); end            # (N, l, 4, n) column-major == int64 [n][4][l][N]
        for p in 1:length(ks); write(f, flatten_ksk(ks[p])); end       # (n+1, B-1, t, N) column-major == int32 [N][t][B-1][n+1]
        if secret_keys !== nothing
            for sk in secret_keys; write(f, Int32.(sk.key.key)); end
        end
    end
end

function write_ciphertexts(path::String, xs::Vector{MKLweSample})
    open(path, "w") do f
        write(f, "MKTFHE3C"); write(f, UInt32(1))
        write(f, Int32(size(xs[1].a, 2))); write(f, Int32(size(xs[1].a, 1))); write(f, UInt64(length(xs)))
        for x in xs; write(f, x.a); end
        for x in xs; write(f, Int32(x.b)); end
    end
end

export TransformedBootstrapKeyPart_3gen, mk_bootstrap_3gen, mk_bootstrap_wo_keyswitch_3gen, mk_blind_rotate_and_extract_3gen,
       mk_keyswitch_3gen, mk_int_mul_3gen, engine_for, mk_gate_nand_3gen, mk_gate_or_3gen, mk_gate_and_3gen,
       mk_gate_xor_3gen, mk_gate_3and_3gen, mk_gate_not_3gen, mk_gate_mux_3gen, mk_gate_nand_3gen_wb, mk_gate_or_3gen_wb, mk_gate_and_3gen_wb,
       mk_gate_xor_3gen_wb, gate_level, mk_copy_3gen, mk_int_add_with_carry_3gen,
       mk_add_3gen, mk_add_3gen_v2, mk_inv_3gen, mk_sub_3gen, mk_less_3gen, mk_grt_3gen, mk_leq_3gen, mk_geq_3gen, write_keys, write_ciphertexts,
       TransformedTGswSample_3gen, tgsw_samples, tgsw_extern_mul_3gen, mk_mux_rotate_3gen, mk_ith_blind_rotate_3gen, mk_blind_rotate_3gen,
       release!, device_count, default_devices

end # module
