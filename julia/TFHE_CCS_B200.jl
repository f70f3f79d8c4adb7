# TFHE_CCS_B200.jl -- `ccall` shim that puts libmktfhe_b200.so behind the reference's CCS multi-key gate API
# (Chen-Chillotti-Song: 3-gen-mk-tfhe/src/mk_api.jl:333-459, mk_internals.jl:703-719, 793-858, mk_gates.jl).
#
# NOT EXECUTED IN THIS REPO'S CI: the build image has no Julia.  The executable twin of this file is
# torus-fhe_b200/tfhe_ccs.py (ctypes, same C ABI, same names), which the tests and benchmarks drive; DESIGN.md section 4c
# explains the composition: the hybrid product (UniProduct, mk_internals.jl:471-535) is two rounds of body-only Torus32 external
# products of the engine, run by the library for the whole blind rotation (mktfhe_ccs_blind_rotate_batch).
#
# The reference's own key objects are used as they are: SharedKey, SecretKey, CloudKeyPart (its MKTGswUESample rows are integer
# polynomials, mk_internals.jl:351-363) -- only MKCloudKey is replaced, by the type below that loads the GPU.
#
# Usage inside the reference tree: `include("src/TFHE.jl"); include("TFHE_CCS_B200.jl"); using .TFHE_CCS_B200`, then
#     shared_key = SharedKey(rng, params); ck_parts = [CloudKeyPart(rng, sk, shared_key) for sk in secret_keys]
#     cloud_key = TFHE_CCS_B200.MKCloudKey(ck_parts, shared_key)
#     TFHE_CCS_B200.mk_gate_nand(cloud_key, x, y)            # x, y :: MKLweSample or Vector{MKLweSample} (one library call per vector)
module TFHE_CCS_B200

using ..TFHE: SchemeParameters, SharedKey, CloudKeyPart, KeyswitchKey, MKLweSample, LweParams, TGswParams,
              lwe_parameters, tgsw_parameters, encode_message, mk_lwe_noiseless_trivial

const LIB = get(ENV, "MKTFHE_B200_LIB", "libmktfhe_b200")
const FLAG_TORUS32 = Int32(1)                 # mktfhe_params.reserved: MKTFHE_FLAG_TORUS32

struct CParams            # mktfhe_params (include/mktfhe_b200.h)
    n::Int32; N::Int32; k::Int32; l::Int32; bgbit::Int32; t::Int32; basebit::Int32; reserved::Int32
end

check(ctx, rc) = rc == 0 ? nothing :
    error("libmktfhe_b200 error $rc: " * unsafe_string(ccall((:mktfhe_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx)))

function create(prm::CParams, device::Integer)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:mktfhe_create, LIB), Cint, (Ref{CParams}, Cint, Ref{Ptr{Cvoid}}), prm, device, out)
    rc == 0 || error("mktfhe_create: " * unsafe_string(ccall((:mktfhe_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    out[]
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

"""mk_api.jl:390-405.  Building it loads the GPU: one Torus32-mode context holding the k (k + 2) n key elements of the hybrid product
(as k (k + 2) pseudo-parties of n elements: pseudo-party party (k + 1) + i serves round 1 of polynomial i -- part_1 = d1 of that key bit,
part_4 = -a for the b polynomial or b_i for a_i --, pseudo-party k (k + 1) + party round 2 -- part_1 = f0, part_4 = f1), and one
context holding the parties' key-switching keys."""
mutable struct MKCloudKey
    parties :: Int
    params :: SchemeParameters
    products :: Ptr{Cvoid}
    switch :: Ptr{Cvoid}
end

function MKCloudKey(ck_parts::Array{CloudKeyPart, 1}, shared_key::SharedKey; device::Integer = 0)
    params = ck_parts[1].params
    k = length(ck_parts)
    @assert k <= params.max_parties
    tp = tgsw_parameters(params)
    n, l, N = Int(params.lwe_size), Int(tp.decomp_length), Int(params.rlwe_polynomial_degree)
    t, bb = Int(params.ks_decomp_length), Int(params.ks_log2_base)
    products = create(CParams(n, N, k * (k + 2), l, tp.log2_base, t, bb, FLAG_TORUS32), device)
    elems = zeros(Int64, N, l, 4, n)              # (N, l, 4, n) column-major == C int64 [n][4][l][N], values unshifted (Torus32 mode)
    second(i) = i == 0 ? [-Int64.(shared_key.a[q].coeffs) for q in 1:l] : [Int64.(ck_parts[i].bk_part.public_key.b[q].coeffs) for q in 1:l]
    for party in 1:k
        ue = ck_parts[party].bk_part.key_uni_enc                       # MKTGswUESample per key bit (mk_internals.jl:351-363)
        for i in 0:k
            rows = second(i)
            for j in 1:n, q in 1:l
                elems[:, q, 1, j] = Int64.(ue[j].d1[q].coeffs)
                elems[:, q, 4, j] = rows[q]
            end
            check(products, ccall((:mktfhe_load_bsk, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Int64}), products, (party - 1) * (k + 1) + i, elems))
        end
        for j in 1:n, q in 1:l
            elems[:, q, 1, j] = Int64.(ue[j].f0[q].coeffs)
            elems[:, q, 4, j] = Int64.(ue[j].f1[q].coeffs)
        end
        check(products, ccall((:mktfhe_load_bsk, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Int64}), products, k * (k + 1) + party - 1, elems))
    end
    check(products, ccall((:mktfhe_mark_keys_received, LIB), Cint, (Ptr{Cvoid},), products))      # this context multiplies only
    check(products, ccall((:mktfhe_finalize_keys, LIB), Cint, (Ptr{Cvoid},), products))
    switch = create(CParams(n, N, k, 2, 7, t, bb, Int32(0)), device)
    for party in 1:k
        check(switch, ccall((:mktfhe_load_ksk, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Int32}), switch, party - 1, flatten_ksk(ck_parts[party].ks)))
    end
    check(switch, ccall((:mktfhe_mark_keys_received, LIB), Cint, (Ptr{Cvoid},), switch))          # ... and this one switches only
    check(switch, ccall((:mktfhe_finalize_keys, LIB), Cint, (Ptr{Cvoid},), switch))
    ck = MKCloudKey(k, params, products, switch)
    finalizer(release!, ck)
    ck
end

"""Frees the device key replicas now (otherwise the finalizer does)."""
function release!(ck::MKCloudKey)
    ck.products == C_NULL || ccall((:mktfhe_destroy, LIB), Cvoid, (Ptr{Cvoid},), ck.products)
    ck.switch == C_NULL || ccall((:mktfhe_destroy, LIB), Cvoid, (Ptr{Cvoid},), ck.switch)
    ck.products = ck.switch = C_NULL
    nothing
end

pack_a(xs::Vector{MKLweSample}) = reduce(hcat, [vec(x.a) for x in xs])     # x.a is (n, k) column-major: (n k, G) == C int32 [G][k][n]
pack_b(xs::Vector{MKLweSample}) = Int32[x.b for x in xs]

"""mk_bootstrap_wo_keyswitch (mk_internals.jl:839-850) on a vector: the extracted samples, one mask of dimension N per party."""
function mk_bootstrap_wo_keyswitch(ck::MKCloudKey, mu::Int32, xs::Vector{MKLweSample})
    k, N, G = ck.parties, Int(ck.params.rlwe_polynomial_degree), length(xs)
    ext_a, ext_b = Array{Int32, 3}(undef, N, k, G), Vector{Int32}(undef, G)
    check(ck.products, ccall((:mktfhe_ccs_blind_rotate_batch, LIB), Cint,
        (Ptr{Cvoid}, Cint, Int32, Csize_t, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}), ck.products, k, mu, G, pack_a(xs), pack_b(xs), ext_a, ext_b))
    [MKLweSample(LweParams(N), ext_a[:, :, g], ext_b[g], 0.0) for g in 1:G]
end

"""mk_keyswitch (mk_internals.jl:703-719) on a vector of extracted samples."""
function mk_keyswitch(ck::MKCloudKey, us::Vector{MKLweSample})
    k, n, G = ck.parties, Int(ck.params.lwe_size), length(us)
    oa, ob = Array{Int32, 3}(undef, n, k, G), Vector{Int32}(undef, G)
    check(ck.switch, ccall((:mktfhe_mk_keyswitch_batch, LIB), Cint, (Ptr{Cvoid}, Csize_t, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
        ck.switch, G, pack_a(us), pack_b(us), oa, ob))
    [MKLweSample(lwe_parameters(ck.params), oa[:, :, g], ob[g], 0.0) for g in 1:G]
end

"""mk_bootstrap (mk_internals.jl:853-856)."""
mk_bootstrap(ck::MKCloudKey, mu::Int32, xs::Vector{MKLweSample}) = mk_keyswitch(ck, mk_bootstrap_wo_keyswitch(ck, mu, xs))
mk_bootstrap(ck::MKCloudKey, mu::Int32, x::MKLweSample) = mk_bootstrap(ck, mu, [x])[1]

"""mk_gate_nand (mk_gates.jl:7-13): bootstrap(1/8 - x - y)."""
function mk_gate_nand(ck::MKCloudKey, xs::Vector{MKLweSample}, ys::Vector{MKLweSample})
    mu = encode_message(1, 8)
    temp = [mk_lwe_noiseless_trivial(mu, xs[g].params, ck.parties) - xs[g] - ys[g] for g in 1:length(xs)]
    mk_bootstrap(ck, mu, temp)
end
mk_gate_nand(ck::MKCloudKey, x::MKLweSample, y::MKLweSample) = mk_gate_nand(ck, [x], [y])[1]

export MKCloudKey, mk_gate_nand, mk_bootstrap, mk_bootstrap_wo_keyswitch, mk_keyswitch, release!

end # module
