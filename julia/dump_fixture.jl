# dump_fixture.jl -- run the REAL reference once and write what it computed, so that this repo's oracle and GPU engine can be pinned
# on the reference's own bytes (DESIGN.md section 1c: today "parity unpinned" -- no Julia in the build image).
#
# NOT EXECUTED IN THIS REPO'S CI.  On any machine with Julia and the reference's packages:
#     cd Torus-FHE/3-gen-mk-tfhe
#     julia --project=. /path/to/repo/julia/dump_fixture.jl /path/to/repo/tests/golden/julia [seed]
# It follows multikey_3gen.jl:15-30 for key generation, keeps the INTEGER bootstrapping key next to the reference's FFT form, runs the
# reference's own mk_gate_{nand,and,or,xor}_3gen and mk_bootstrap_3gen on a truth table, and writes (little-endian, layouts of
# torus-fhe_b200/interchange.py):
#     keys.bin          MKTFHE3K: parameters, integer bsk, ksk, LWE secret keys
#     x.bin, y.bin      MKTFHE3C: the input ciphertexts
#     nand.bin, and.bin, or.bin, xor.bin, boot.bin     MKTFHE3C: the REFERENCE's outputs (FFT path)
#     plain.txt         the plaintext bits, one "x y" pair per line
# tests/test_julia_fixture.py picks the directory up when it exists (it is skipped otherwise).
include(joinpath(pwd(), "src", "TFHE.jl"))
using Random
using .TFHE
include(joinpath(@__DIR__, "TFHE_B200.jl"))       # write_keys / write_ciphertexts only: nothing here calls into libmktfhe_b200

function main()
    outdir = ARGS[1]
    seed = length(ARGS) > 1 ? parse(Int, ARGS[2]) : 20261018
    mkpath(outdir)
    parties = 2
    params = mktfhe_parameters_2party_3gen
    rng = MersenneTwister(seed)

    secret_keys = [SecretKey_3gen(rng, params) for _ in 1:parties]
    rlwe_keys = [RLweKey(rng, rlwe_parameters(params), true) for _ in 1:parties]
    crp_a = CRP_3gen(rng, tgsw_parameters(params), rlwe_parameters(params), true)
    pubkeys = [PublicKey(rng, rlwe_keys[i], params.gsw_noise_stddev, crp_a, tgsw_parameters(params), 1) for i in 1:parties]
    common_pubkey = CommonPubKey_3gen(pubkeys, params, parties)
    bk_int = [BootstrapKeyPart_3gen(rng, secret_keys[i].key, params.gsw_noise_stddev, crp_a, common_pubkey,
                                    tgsw_parameters(params), rlwe_parameters(params), 1) for i in 1:parties]
    bk_ref = [TransformedBootstrapKeyPart_3gen(bk_int[i]) for i in 1:parties]                  # the reference's FFT form
    bk_b200 = [TFHE_B200.TransformedBootstrapKeyPart_3gen(bk_int[i]) for i in 1:parties]       # the same key as integers
    ks_keys = [KeyswitchKey(rng, params.ks_noise_stddev, keyswitch_parameters(params), secret_keys[i].key, rlwe_keys[i]) for i in 1:parties]
    TFHE_B200.write_keys(joinpath(outdir, "keys.bin"), params, bk_b200, ks_keys, secret_keys)

    bits = [(false, false), (false, true), (true, false), (true, true)]
    bits = vcat(bits, bits, bits, bits)                                                         # 16 gates of each kind
    xs = [mk_encrypt_3gen(rng, secret_keys, b[1]) for b in bits]
    ys = [mk_encrypt_3gen(rng, secret_keys, b[2]) for b in bits]
    TFHE_B200.write_ciphertexts(joinpath(outdir, "x.bin"), xs)
    TFHE_B200.write_ciphertexts(joinpath(outdir, "y.bin"), ys)
    open(joinpath(outdir, "plain.txt"), "w") do f
        for b in bits; println(f, Int(b[1]), " ", Int(b[2])); end
    end
    for (name, gate) in (("nand", mk_gate_nand_3gen), ("and", mk_gate_and_3gen), ("or", mk_gate_or_3gen), ("xor", mk_gate_xor_3gen))
        outs = [gate(bk_ref, ks_keys, xs[i], ys[i]) for i in 1:length(bits)]
        for i in 1:length(bits)
            mk_decrypt_3gen(secret_keys, outs[i]) == (name == "nand" ? !(bits[i][1] && bits[i][2]) : name == "and" ? (bits[i][1] && bits[i][2]) :
                                                      name == "or" ? (bits[i][1] || bits[i][2]) : xor(bits[i][1], bits[i][2])) ||
                println("note: the reference itself decrypts gate ", name, " #", i, " wrongly (scheme noise)")
        end
        TFHE_B200.write_ciphertexts(joinpath(outdir, name * ".bin"), outs)
    end
    boots = [mk_bootstrap_3gen(bk_ref, ks_keys, encode_message64(1, 8), xs[i]) for i in 1:length(bits)]
    TFHE_B200.write_ciphertexts(joinpath(outdir, "boot.bin"), boots)
    println("wrote fixture to ", outdir)
end

main()
