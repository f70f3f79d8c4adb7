# TFHE1_B200.jl -- `ccall` shim that puts libmktfhe_b200.so behind the reference's SINGLE-KEY gate API
# (3-gen-mk-tfhe/src/api.jl:212-266, gates.jl:16-177, bootstrap.jl:73-100, keyswitch.jl:45-80).
#
# NOT EXECUTED IN THIS REPO'S CI: the build image has no Julia.  The executable twin of this file is
# torus-fhe_b200/tfhe1.py (ctypes, same C ABI, same names), which the tests and benchmarks drive; DESIGN.md section 4b
# explains the mapping (one party, Torus32 values carried as v << 32, a standard TGSW sample as the four 3gen parts).
#
# Usage inside the reference tree: `include("src/TFHE.jl"); include("TFHE1_B200.jl"); using .TFHE1_B200`, then
#     secret_key, cloud_key = TFHE1_B200.make_key_pair(rng, params)     # params: tfhe_parameters_80 / tfhe_parameters_128 (N = 1024)
#     TFHE1_B200.gate_nand(cloud_key, x, y)                             # x, y :: LweSample or Vector{LweSample} (one launch per vector)
module TFHE1_B200

using Random
using ..TFHE: SchemeParameters, SecretKey, LweSample, LweParams, RLweKey, KeyswitchKey, TGswParams,
              lwe_parameters, rlwe_parameters, tgsw_parameters, keyswitch_parameters, tgsw_encrypt, encode_message,
              lwe_noiseless_trivial

const LIB = get(ENV, "MKTFHE_B200_LIB", "libmktfhe_b200")

struct CParams            # mktfhe_params (include/mktfhe_b200.h)
    n::Int32; N::Int32; k::Int32; l::Int32; bgbit::Int32; t::Int32; basebit::Int32; reserved::Int32
end

check(ctx, rc) = rc == 0 ? nothing :
    error("libmktfhe_b200 error $rc: " * unsafe_string(ccall((:mktfhe_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx)))

"""api.jl:212-228 with the INTEGER bootstrapping key kept (the reference's BootstrapKey holds only the Float64 FFTs of its TGSW
samples, bootstrap.jl:12-15) and the GPU context that holds both keys."""
mutable struct CloudKey
    params :: SchemeParameters
    parts :: Array{Int64, 4}          # (N, l, 4, n) column-major == C int64 [n][4][l][N]: the four 3gen parts, Torus32 values << 32
    keyswitch_key :: KeyswitchKey
    ctx :: Ptr{Cvoid}
end

"""GPUs a cloud key's engine spans: ENV["MKTFHE_B200_DEVICES"] = "all" | "0,1,2,3" | unset (GPU 0 only)."""
function default_devices()
    v = get(ENV, "MKTFHE_B200_DEVICES", "")
    v == "" && return Cint[0]
    v == "all" && return Cint[]
    Cint[parse(Cint, x) for x in split(v, ",")]
end

function flatten_ksk(ks::KeyswitchKey)        # Array{LweSample,3} (base-1, t, N) -> Int32 (n+1, base-1, t, N) == C [N][t][base-1][n+1]
    B1, t, N = size(ks.key)
    n = ks.out_lwe_params.size
    rows = Array{Int32, 4}(undef, n + 1, B1, t, N)
    for i in 1:N, j in 1:t, h in 1:B1
        rows[1:n, h, j, i] = ks.key[h, j, i].a
        rows[n + 1, h, j, i] = ks.key[h, j, i].b
    end
    rows
end

function CloudKey(rng::AbstractRNG, secret_key::SecretKey; devices::Vector{Cint} = default_devices())
    params = secret_key.params
    rlwe_key = RLweKey(rng, rlwe_parameters(params))
    tp = tgsw_parameters(params)
    n, l, N = params.lwe_size, tp.decomp_length, params.rlwe_polynomial_degree
    parts = Array{Int64, 4}(undef, N, l, 4, n)
    # gadget digits of up to 8 bits ride the default kernels (Torus32 values as v << 32); wider ones (tfhe_parameters_80: Bg = 2^10)
    # use the library's Torus32 mode (include/mktfhe_b200.h, MKTFHE_FLAG_TORUS32): keys unshifted
    t32 = tp.log2_base > 8
    sh = t32 ? 0 : 32
    for j in 1:n
        s = tgsw_encrypt(rng, secret_key.key.key[j], params.bs_noise_stddev, rlwe_key, tp)     # TGswSample: samples[q, row], tgsw.jl:36-46
        for q in 1:l
            # part_1 body <- body digits, part_2 body <- mask digits, part_3 mask <- mask digits, part_4 mask <- body digits
            parts[:, q, 1, j] = Int64.(s.samples[q, 2].a[2].coeffs) .<< sh
            parts[:, q, 2, j] = Int64.(s.samples[q, 1].a[2].coeffs) .<< sh
            parts[:, q, 3, j] = Int64.(s.samples[q, 1].a[1].coeffs) .<< sh
            parts[:, q, 4, j] = Int64.(s.samples[q, 2].a[1].coeffs) .<< sh
        end
    end
    ks = KeyswitchKey(rng, params.ks_noise_stddev, keyswitch_parameters(params), secret_key.key, rlwe_key)
    prm = CParams(n, N, 1, l, tp.log2_base, ks.params.decomp_length, ks.params.log2_base, t32 ? 1 : 0)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:mktfhe_create_multi, LIB), Cint, (Ref{CParams}, Cint, Ptr{Cint}, Ref{Ptr{Cvoid}}), prm, length(devices),
               isempty(devices) ? Ptr{Cint}(C_NULL) : pointer(devices), out)
    rc == 0 || error("mktfhe_create_multi: " * unsafe_string(ccall((:mktfhe_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    ctx = out[]
    check(ctx, ccall((:mktfhe_load_bsk, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Int64}), ctx, 0, parts))
    check(ctx, ccall((:mktfhe_load_ksk, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Int32}), ctx, 0, flatten_ksk(ks)))
    check(ctx, ccall((:mktfhe_finalize_keys, LIB), Cint, (Ptr{Cvoid},), ctx))
    ck = CloudKey(params, parts, ks, ctx)
    finalizer(x -> (x.ctx == C_NULL || ccall((:mktfhe_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.ctx); x.ctx = C_NULL), ck)
    ck
end

"""api.jl:237-245."""
function make_key_pair(rng::AbstractRNG, params::SchemeParameters)
    secret_key = SecretKey(rng, params)
    secret_key, CloudKey(rng, secret_key)
end

pack_a(xs::Vector{LweSample}) = reduce(hcat, [x.a for x in xs])       # (n, G) column-major == C int32 [G][1][n]
pack_b(xs::Vector{LweSample}) = Int32[x.b for x in xs]
unpack(params::LweParams, oa::Matrix{Int32}, ob::Vector{Int32}) = [LweSample(params, oa[:, g], ob[g], 0.0) for g in 1:length(ob)]

const MU = Int64(encode_message(1, 8)) << 32          # the test-vector message, carried as v << 32

"""bootstrap(mu0 + cx x + cy y) with output message 1/8: every bootstrapped two-input gate of gates.jl:16-142."""
function affine_gate(ck::CloudKey, mu0::Int32, cx::Integer, cy::Integer, xs::Vector{LweSample}, ys::Vector{LweSample})
    n, G = Int(ck.params.lwe_size), length(xs)
    oa, ob = Matrix{Int32}(undef, n, G), Vector{Int32}(undef, G)
    check(ck.ctx, ccall((:mktfhe_affine_bootstrap_batch, LIB), Cint,
        (Ptr{Cvoid}, Int32, Int32, Int32, Int32, Int64, Csize_t, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
        ck.ctx, mu0, Int32(cx), Int32(cy), Int32(0), MU, G, pack_a(xs), pack_b(xs), pack_a(ys), pack_b(ys), Ptr{Int32}(C_NULL), Ptr{Int32}(C_NULL), oa, ob))
    unpack(xs[1].params, oa, ob)
end

for (name, m, space, cx, cy) in ((:gate_nand, 1, 8, -1, -1), (:gate_or, 1, 8, 1, 1), (:gate_and, -1, 8, 1, 1), (:gate_xor, 1, 4, 2, 2),
                                 (:gate_xnor, -1, 4, -2, -2), (:gate_nor, -1, 8, -1, -1), (:gate_andny, -1, 8, -1, 1), (:gate_andyn, -1, 8, 1, -1),
                                 (:gate_orny, 1, 8, -1, 1), (:gate_oryn, 1, 8, 1, -1))
    @eval begin
        $name(ck::CloudKey, x::Vector{LweSample}, y::Vector{LweSample}) = affine_gate(ck, encode_message($m, $space), $cx, $cy, x, y)
        $name(ck::CloudKey, x::LweSample, y::LweSample) = $name(ck, [x], [y])[1]
    end
end

gate_not(ck::CloudKey, x::LweSample) = -x                                                  # gates.jl:80-83
gate_constant(ck::CloudKey, value::Bool) = lwe_noiseless_trivial(encode_message(value ? 1 : -1, 8), lwe_parameters(ck.params))

"""bootstrap_wo_keyswitch(bk, mu, x) (bootstrap.jl:73-86) on a vector: the extracted samples of dimension N."""
function bootstrap_wo_keyswitch(ck::CloudKey, mu::Int32, xs::Vector{LweSample})
    N, G = Int(ck.params.rlwe_polynomial_degree), length(xs)
    ext = Matrix{Int32}(undef, N + 1, G)
    check(ck.ctx, ccall((:mktfhe_blind_rotate_batch, LIB), Cint,
        (Ptr{Cvoid}, Int64, Csize_t, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int64}), ck.ctx, Int64(mu) << 32, G, pack_a(xs), pack_b(xs), ext, Ptr{Int64}(C_NULL)))
    [LweSample(LweParams(N), ext[1:N, g], ext[N + 1, g], 0.0) for g in 1:G]
end

"""keyswitch(ks, sample) (keyswitch.jl:45-80) on a vector of extracted samples."""
function keyswitch(ck::CloudKey, us::Vector{LweSample})
    n, N, G = Int(ck.params.lwe_size), Int(ck.params.rlwe_polynomial_degree), length(us)
    ext = Matrix{Int32}(undef, N + 1, G)
    for g in 1:G
        ext[1:N, g] = us[g].a
        ext[N + 1, g] = us[g].b
    end
    oa, ob = Matrix{Int32}(undef, n, G), Vector{Int32}(undef, G)
    check(ck.ctx, ccall((:mktfhe_keyswitch_batch, LIB), Cint, (Ptr{Cvoid}, Csize_t, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}), ck.ctx, G, ext, oa, ob))
    unpack(LweParams(n), oa, ob)
end

"""bootstrap(bk, ks, mu, x) (bootstrap.jl:97-100)."""
function bootstrap(ck::CloudKey, mu::Int32, xs::Vector{LweSample})
    n, G = Int(ck.params.lwe_size), length(xs)
    oa, ob = Matrix{Int32}(undef, n, G), Vector{Int32}(undef, G)
    check(ck.ctx, ccall((:mktfhe_bootstrap_batch, LIB), Cint,
        (Ptr{Cvoid}, Int64, Csize_t, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}), ck.ctx, Int64(mu) << 32, G, pack_a(xs), pack_b(xs), oa, ob))
    unpack(xs[1].params, oa, ob)
end

"""gate_mux (gates.jl:166-177): two bootstraps without key switch, the OR in the extracted domain, one key switch."""
function gate_mux(ck::CloudKey, xs::Vector{LweSample}, ys::Vector{LweSample}, zs::Vector{LweSample})
    mu = encode_message(1, 8)
    G = length(xs)
    t1 = [lwe_noiseless_trivial(encode_message(-1, 8), xs[g].params) + xs[g] + ys[g] for g in 1:G]
    t2 = [lwe_noiseless_trivial(encode_message(-1, 8), xs[g].params) - xs[g] + zs[g] for g in 1:G]
    u = bootstrap_wo_keyswitch(ck, mu, vcat(t1, t2))
    t3 = [lwe_noiseless_trivial(mu, u[g].params) + u[g] + u[G + g] for g in 1:G]
    keyswitch(ck, t3)
end
gate_mux(ck::CloudKey, x::LweSample, y::LweSample, z::LweSample) = gate_mux(ck, [x], [y], [z])[1]

"""Frees the device key replicas now (otherwise the finalizer does)."""
function release!(ck::CloudKey)
    ck.ctx == C_NULL || ccall((:mktfhe_destroy, LIB), Cvoid, (Ptr{Cvoid},), ck.ctx)
    ck.ctx = C_NULL
    nothing
end

export CloudKey, make_key_pair, gate_nand, gate_or, gate_and, gate_xor, gate_xnor, gate_not, gate_constant, gate_nor, gate_andny, gate_andyn,
       gate_orny, gate_oryn, gate_mux, bootstrap, bootstrap_wo_keyswitch, keyswitch, release!, default_devices

end # module
