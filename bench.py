#!/usr/bin/env python
"""bench.py -- bootstrapped 2-party MK NAND gates/sec (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--gates G] [--impl ours|reference]

A step = one pass of the hot path (mk_gate_nand_3gen: linear prologue -> blind rotate -> sample extract -> MK key
switch) over one batch of G independent gates (default 16384 = BASELINE configs[1]) per GPU.  Multi-GPU is weak
scaling: every rank processes its own G gates, keys are generated on rank 0 and broadcast once over NCCL.

  value  gates/s with the ciphertexts already resident in HBM (device-pointer C-ABI entry, CUDA events)
  e2e    gates/s through the host-pointer C-ABI entry (pinned host buffers, H2D + D2H inside the timed region)
  roofline      dominant kernel = blind_rotate_kernel; algorithmic bytes = streamed bootstrapping-key bytes per gate
                 x gates per launch, against the measured HBM copy peak
  cpu_baseline  the oracle's Float64-FFT restatement of the reference algorithm on the host cores (bounded sample)

`--impl reference` times that CPU restatement alone (the reference itself is Julia and cannot run here).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bootstrapped 2-party MK NAND gates/sec"
UNIT = "gates/s"
KEY_SEED = 0xB20000A1
DATA_SEED = 0xB2000001


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(int(r[0])); mx = max(mx, int(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def oracle_keyset():
    from oracle import mk_oracle as O
    O.build()
    return O, O.KeySet(O.PARAMS_2PARTY, seed=KEY_SEED, nthreads=host_cores())


def cpu_sample(O, ks, count, threads):
    """`count` NAND gates through the Float64-FFT restatement (polynomials.jl:208-242, tgsw_3gen.jl:102-113, keyswitch.jl:45-80)."""
    bits = np.random.default_rng(DATA_SEED).integers(0, 2, (2, count)).astype(np.uint8)
    x, y = ks.encrypt(bits[0], DATA_SEED), ks.encrypt(bits[1], DATA_SEED + 1)
    O.lib().mko_prepare_fft_key(ks.h)
    t0 = time.perf_counter()
    oa, ob = ks.gate_batch(O.FFT, O.GATE_NAND, x, y, nthreads=threads)
    dt = time.perf_counter() - t0
    ok = bool(np.array_equal(ks.decrypt(oa, ob), ~(bits[0].astype(bool) & bits[1].astype(bool))))
    return dt, ok


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    O, ks = oracle_keyset()
    cores = host_cores()
    per_step = max(cores, 8) * 2            # ~0.17 s per gate per core -> a step is well under a minute
    for _ in range(args.warmup):
        cpu_sample(O, ks, cores, cores)
    t, ok_all = 0.0, True
    for _ in range(args.steps):
        dt, ok = cpu_sample(O, ks, per_step, cores)
        t += dt; ok_all &= ok
    v = per_step * args.steps / t
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": "2-party NAND x16384 (mktfhe_parameters_2party_3gen), bounded CPU sample",
                                            "gates_per_step": per_step},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{per_step} gates/step x {args.steps} steps, Float64-FFT restatement of the reference algorithm "
                                       "(C, pthreads, one gate per thread); the Julia reference cannot run in this image"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "decryptions_correct": ok_all}
    print(json.dumps(line))
    return 0


def generate_keys(T, params, rng, ksk_on_device=True):
    """multikey_3gen.jl:15-30 through the product's own host mirror (exact key products run on the GPU)."""
    k = params.max_parties
    secret_keys = [T.SecretKey_3gen(rng, params) for _ in range(k)]
    rlwe_keys = [T.RLweKey(rng, T.rlwe_parameters(params), True) for _ in range(k)]
    crp = T.CRP_3gen(rng, T.tgsw_parameters(params), T.rlwe_parameters(params), True)
    pubkeys = [T.PublicKey(rng, rlwe_keys[i], params.gsw_noise_stddev, crp, T.tgsw_parameters(params), 1) for i in range(k)]
    common = T.CommonPubKey_3gen(pubkeys, params, k)
    bk = [T.BootstrapKeyPart_3gen(rng, secret_keys[i].key, params.gsw_noise_stddev, crp, common, T.tgsw_parameters(params),
                                  T.rlwe_parameters(params), 1) for i in range(k)]
    bk = [T.TransformedBootstrapKeyPart_3gen(b) for b in bk]
    mk_ks = T.KeyswitchKey.on_device if ksk_on_device else T.KeyswitchKey       # rows generated by the GPU (mktfhe_generate_ksk) or by numpy
    ks = [mk_ks(rng, params.ks_noise_stddev, T.keyswitch_parameters(params), secret_keys[i].key, rlwe_keys[i]) for i in range(k)]
    return secret_keys, bk, ks


def profile_evidence(T, parties, lib_path, N=1024, l=2, engine="ntt_rns"):
    """Measured constants the roofline quotes, read from files under profiles/ (never literals): the IMAD-pipe peak and the HBM
    key-stream rate from the micro-benchmarks, and -- only when the capture manifest belongs to THIS library's machine code and
    parameter set -- the DRAM traffic and pipe-busy figures of the dominant kernel's `ncu --set full` capture."""
    import hashlib
    import re
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from kernel_id import hot_kernels, kernel_id
    prof = os.path.join(ROOT, "profiles")
    ev = {"kernel_id": kernel_id(lib_path, hot_kernels(N, l, engine)), "imad_peak": None, "fp64_peak": None, "key_stream": None, "capture": None,
          "capture_note": None}

    def sha(path):
        return hashlib.sha256(open(path, "rb").read()).hexdigest()[:16]
    try:
        f = os.path.join(prof, "pipe_ubench_r1.txt")
        m = re.search(r"^IMAD \(lo\)\s+[\d.]+ ms\s+([\d.]+) T thread-ops/s\s+\(([\d.]+) lanes/clk/SM", open(f).read(), flags=re.M)
        ev["imad_peak"] = {"value": float(m.group(1)) * 1e12, "lanes_per_clk_sm": float(m.group(2)), "source": "profiles/pipe_ubench_r1.txt", "sha256": sha(f)}
    except Exception as e:
        ev["imad_peak_note"] = f"profiles/pipe_ubench_r1.txt unreadable: {e}"
    try:
        f = os.path.join(prof, "fp64_pipe_ubench_r2.txt")
        m = re.search(r"^DFMA, all warps\s+[\d.]+ ms\s+int\s+[\d.]+\s+f64\s+([\d.]+)\s+lanes/clk/SM", open(f).read(), flags=re.M)
        ev["fp64_peak"] = {"lanes_per_clk_sm": float(m.group(1)), "source": "profiles/fp64_pipe_ubench_r2.txt", "sha256": sha(f)}
    except Exception as e:
        ev["fp64_peak_note"] = f"profiles/fp64_pipe_ubench_r2.txt unreadable: {e}"
    try:
        f = os.path.join(prof, "key_stream_ubench_r1.txt")
        m = re.search(r"x 1 CTAs x 12 warps.*?:\s+([\d.]+) GB/s", open(f).read())
        ev["key_stream"] = {"value": float(m.group(1)), "source": "profiles/key_stream_ubench_r1.txt (the kernel's key access pattern over a 4 GiB buffer at "
                            "the kernel's occupancy; in the product L2 serves the stream)", "sha256": sha(f)}
    except Exception:
        pass
    f = os.path.join(prof, f"ncu_capture_{parties}party.json" if engine == "ntt_rns" else f"ncu_capture_{parties}party_{engine}.json")
    if not os.path.exists(f):
        ev["capture_note"] = f"no ncu capture manifest for the {parties}-party set ({os.path.relpath(f, ROOT)})"
    else:
        man = json.load(open(f))
        summ = os.path.join(ROOT, man["summary_file"])
        if ev["kernel_id"] is None:
            ev["capture_note"] = "cuobjdump unavailable: the library's kernel identity could not be checked against the capture"
        elif man["kernel_id"] != ev["kernel_id"]:
            ev["capture_note"] = f"capture {man['summary_file']} was taken from kernel id {man['kernel_id']}, this library is {ev['kernel_id']}: not quoted"
        elif not os.path.exists(summ) or hashlib.sha256(open(summ, "rb").read()).hexdigest() != man["summary_sha256"]:
            ev["capture_note"] = f"{man['summary_file']} is missing or does not match the manifest's sha256: not quoted"
        else:
            ev["capture"] = man
    return ev


def run_ours(args):
    import torch
    import torch.distributed as dist
    import torus_fhe_b200 as T

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU baseline)")
    # --abi-multi: ONE process drives args.gpus GPUs through one mktfhe_create_multi context (what a Julia host gets);
    # default: one process per GPU (torchrun), NCCL key broadcast by the host side
    ndev = args.gpus if args.abi_multi else 1
    if args.abi_multi and world > 1:
        raise SystemExit("bench.py: --abi-multi is a single-process mode (do not launch it under torchrun)")
    if ndev > torch.cuda.device_count():
        raise SystemExit(f"bench.py: --abi-multi --gpus {ndev} but {torch.cuda.device_count()} GPUs are visible")
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "NONE"        # NCCL prints its version banner on stdout at VERSION/WARN: keep stdout to the JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    params = {2: T.mktfhe_parameters_2party_3gen, 3: T.mktfhe_parameters_3party_3gen, 4: T.mktfhe_parameters_4party_3gen,
              5: T.mktfhe_parameters_5party_3gen, 8: T.mktfhe_parameters_8party_3gen, 16: T.mktfhe_parameters_16party_3gen}[args.parties]
    G, k, n = args.gates, params.max_parties, params.lwe_size
    devices = list(range(ndev)) if args.abi_multi else None
    eng = T.Engine(params, device=local, devices=devices)
    rng = np.random.default_rng(KEY_SEED)
    secret_keys = None
    t_keys = time.perf_counter()
    if rank == 0:
        secret_keys, bk, ks = generate_keys(T, params, rng)
        eng.load_keys([b.gsw_key for b in bk], [q.engine_part() for q in ks])      # multi-device context: includes the in-library broadcast
    if world > 1:
        eng.broadcast_keys(src=0)
    t_keys = time.perf_counter() - t_keys
    ctx = eng.ctx
    GT = G * ndev                                     # gates this process handles per step (weak scaling: G per GPU)

    # synthetic random ciphertexts (timing is data-independent); the first 64 gates of every batch are valid encryptions
    # so rank 0 can check the decrypted truth table of exactly the batches it timed
    drng = np.random.default_rng(DATA_SEED + rank)
    host = [torch.empty((GT, k, n), dtype=torch.int32).pin_memory(), torch.empty(GT, dtype=torch.int32).pin_memory(),
            torch.empty((GT, k, n), dtype=torch.int32).pin_memory(), torch.empty(GT, dtype=torch.int32).pin_memory()]
    for h in host:
        h.numpy()[...] = drng.integers(-2 ** 31, 2 ** 31, size=tuple(h.shape), dtype=np.int64).astype(np.int32)
    V = min(64, G)
    bits = drng.integers(0, 2, (2, V)).astype(bool)
    if secret_keys is not None:
        ex, ey = T.mk_encrypt_3gen(drng, secret_keys, bits[0]), T.mk_encrypt_3gen(drng, secret_keys, bits[1])
        host[0].numpy()[:V], host[1].numpy()[:V], host[2].numpy()[:V], host[3].numpy()[:V] = ex.a, ex.b, ey.a, ey.b
    host_oa, host_ob = torch.empty((GT, k, n), dtype=torch.int32).pin_memory(), torch.empty(GT, dtype=torch.int32).pin_memory()
    # per GPU of this process: the replica context, its slice of the inputs resident in HBM, outputs, a stream
    reps = []
    for i in range(ndev):
        d = eng.devices[i] if args.abi_multi else local
        lo, hi = ctx.shard_bounds(GT, i) if args.abi_multi else (0, G)
        with torch.cuda.device(d):
            st = torch.cuda.Stream(device=d)      # kernels are launched on THIS stream; the CUDA events below are recorded on it
            reps.append({"dev": d, "ctx": ctx.replica(i) if ndev > 1 else ctx, "n": hi - lo, "stream": st,
                         "in": [h[lo:hi].to(f"cuda:{d}") for h in host],
                         "oa": torch.empty((hi - lo, k, n), dtype=torch.int32, device=f"cuda:{d}"),
                         "ob": torch.empty(hi - lo, dtype=torch.int32, device=f"cuda:{d}")})

    def step_dev():
        for r in reps:
            i = r["in"]
            r["ctx"].gate_batch_dev(T._cabi.GATE_NAND, r["n"], i[0].data_ptr(), i[1].data_ptr(), i[2].data_ptr(), i[3].data_ptr(), 0, 0,
                                    r["oa"].data_ptr(), r["ob"].data_ptr(), stream=r["stream"].cuda_stream)

    def step_host():
        ctx.gate_batch(T._cabi.GATE_NAND, (host[0].numpy(), host[1].numpy()), (host[2].numpy(), host[3].numpy()),
                       out=(host_oa.numpy(), host_ob.numpy()))

    def sync_all():
        for r in reps:
            torch.cuda.synchronize(r["dev"])

    def barrier():
        sync_all()
        if world > 1:
            dist.barrier()
        sync_all()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        step_dev()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- timed region: device-resident inputs -------------------------------------------------------------------
    launches0 = ctx.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in reps]
    barrier()
    for r, (e0, _) in zip(reps, ev):
        e0.record(r["stream"])
    for _ in range(args.steps):
        step_dev()
    for r, (_, e1) in zip(reps, ev):
        e1.record(r["stream"])
    barrier()
    ms = max_over_ranks(max(e0.elapsed_time(e1) for e0, e1 in ev))      # slowest GPU of this process, then of all ranks
    launches = ctx.launch_count() - launches0
    br_ms, ks_ms = ctx.last_kernel_ms()           # the last step's two kernels (multi-device: slowest GPU), CUDA events on the launching stream
    fused_ks = bool(reps[0]["ctx"].describe()["keyswitch_fused"])       # reported by the library, not inferred from a timing
    # ---- e2e: host-pointer C-ABI call, pinned host buffers, H2D and D2H inside the timed region --------------------
    step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else None

    # latency of ONE bootstrap alone on a GPU: the reference's own protocol times single calls, min / median of 100
    # (measurements/test_suites/performance_comparison_test/perf_comp.jl:20,126-142)
    lat = None
    if rank == 0:
        r = reps[0]
        one = [d[:1].contiguous() for d in r["in"]]
        samples = []
        for it in range(args.latency_trials + 3):
            r["ctx"].gate_batch_dev(T._cabi.GATE_NAND, 1, one[0].data_ptr(), one[1].data_ptr(), one[2].data_ptr(), one[3].data_ptr(), 0, 0,
                                    r["oa"].data_ptr(), r["ob"].data_ptr(), stream=r["stream"].cuda_stream)
            torch.cuda.synchronize(r["dev"])
            if it >= 3:
                samples.append(sum(r["ctx"].last_kernel_ms()))
        lat = {"min": float(np.min(samples)), "median": float(np.median(samples)), "trials": len(samples)}
    ok = None
    if secret_keys is not None:
        got = T.mk_decrypt_3gen(secret_keys, T.MKLweSample(None, host_oa.numpy()[:V], host_ob.numpy()[:V]))
        got_dev = T.mk_decrypt_3gen(secret_keys, T.MKLweSample(None, reps[0]["oa"][:V].cpu().numpy(), reps[0]["ob"][:V].cpu().numpy()))
        ok = bool(np.array_equal(got, ~(bits[0] & bits[1])) and np.array_equal(got_dev, got))

    if rank == 0:
        hbm_peak, peak_src = measured_peaks()
        N, l = params.rlwe_polynomial_degree, params.gsw_decomp_length
        big = N != 1024            # N = 2048 sets: no slot model / ncu capture manifest
        engine = reps[0]["ctx"].describe().get("external_product", "ntt_rns") if not big else "ntt_rns"   # reported by the library
        evd = profile_evidence(T, k, T._cabi.LIB_PATH, N, l, engine)
        bsk_1limb = k * n * 4 * l * N * 8                         # 68.2 MB per 2-party bootstrap (the reference's transformed key size)
        bsk_stream, ksk_gather = ctx.algorithmic_bytes()          # what this build streams (u32 residues per coefficient) / gathers per gate
        ct_io = 2 * (k * n + 1) * 4 + (N + 1) * 4
        per_gate = bsk_1limb + ct_io + (ksk_gather if fused_ks else 0)   # key + gathered ksk rows + ciphertext I/O
        Gk = max(r["n"] for r in reps)                            # gates in the launch that br_ms timed (the slowest GPU's slice)
        hbm_achieved = Gk * per_gate / (br_ms * 1e-3) / 1e9
        # algorithmic IMAD-pipe slots per gate of the three-prime RNS formulation, counting only work the
        # formulation cannot avoid: per blind-rotate step 6l forward NTTs of 4608 multiplying butterflies (the first stage of a digit
        # transform is a table lookup) and 6 inverse NTTs of 5120, 4 slots each (IMAD.HI is half rate); 12l x 1024 pointwise products
        # accumulated in 64 bits (IMAD.WIDE = 2 slots) + 6 x 1024 Montgomery reductions (3 slots); 2048 CRT lifts (17 slots)
        imad_slots = k * n * ((6 * l * 4608 + 6 * 5120) * 4 + 12 * l * 1024 * 2 + 6 * 1024 * 3 + 2 * 1024 * 17)
        cap, peak = evd["capture"], evd["imad_peak"]
        traffic = cap["dram_bytes"] * Gk / cap["gates_in_launch"] if cap else None
        # algorithmic FP64-pipe slots per gate of the exact three-limb FFT formulation: per step 2l forward transforms
        # (512 points x 6 for the first stage formed from the digit bytes + 8 stages x 256 butterflies x 6), 6 x 512 x 2l complex
        # multiply-adds (4), 6 inverse transforms (32 x 148 for the constant-twiddle pass + 4 x 256 x 6), and 512 x 3 last-stage / untwist /
        # rounding tasks (18)
        fp64_slots = k * n * (2 * l * (512 * 6 + 8 * 256 * 6) + 6 * 512 * 2 * l * 4 + 6 * (32 * 148 + 4 * 256 * 6) + 512 * 3 * 18)
        fp64 = None
        if engine == "fft64" and evd["fp64_peak"]:
            sms = int(reps[0]["ctx"].describe()["sms"])
            mhz = float((clocks or {}).get("sm_max_mhz") or 1965)
            fpk = evd["fp64_peak"]["lanes_per_clk_sm"] * sms * mhz * 1e6
            fp64 = {"achieved": Gk * fp64_slots / (br_ms * 1e-3) / 1e12, "peak": fpk / 1e12, "unit": "T DFMA-slots/s", "frac": Gk * fp64_slots / (br_ms * 1e-3) / fpk}
        imad = None
        if not big and peak:
            imad = {"achieved": Gk * imad_slots / (br_ms * 1e-3) / 1e12, "peak": peak["value"] / 1e12, "unit": "T IMAD-slots/s",
                    "frac": Gk * imad_slots / (br_ms * 1e-3) / peak["value"]}
        value = world * GT * args.steps / (ms * 1e-3)
        hbm = {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak, "peak_source": peak_src,
               "algorithmic_bytes_per_gate": per_gate,
               "algorithmic_bytes_breakdown": {"bsk_reference_transformed_size": bsk_1limb, "ksk_rows_gathered": ksk_gather if fused_ks else 0, "ciphertext_io": ct_io},
               "streamed_bytes_per_gate_this_build": bsk_stream,
               "note": "the contract's HBM figure over ALGORITHMIC bytes (SURVEY 8d); these bytes are served by L2 (see traffic), so HBM does not bind",
               "key_stream_from_hbm_GBps": dict(evd["key_stream"], frac_of_peak=evd["key_stream"]["value"] / hbm_peak) if evd["key_stream"] else None}
        roofline = {"kernel": "blind_rotate2k_kernel" if big else "blind_rotate_fft_kernel" if engine == "fft64" else "blind_rotate_kernel", "kernel_id": evd["kernel_id"], "kernel_ms": br_ms,
                    "gates_in_timed_launch": Gk, "keyswitch_ms": ks_ms, "keyswitch_fused_into_blind_rotate": fused_ks,
                    "keyswitch_gather_GBps": None if fused_ks or ks_ms <= 0 else Gk * ksk_gather / (ks_ms * 1e-3) / 1e9,
                    "kernel_share_of_step": br_ms / (ms / args.steps),
                    "traffic": traffic, "traffic_source": ({"file": cap["summary_file"], "sha256": cap["summary_sha256"][:16], "kernel_id": cap["kernel_id"],
                                                            "dram_bytes_in_captured_launch": cap["dram_bytes"], "gates_in_captured_launch": cap["gates_in_launch"],
                                                            "l2_hit_rate": cap["l2_hit_rate"], "scaled": "x gates in this launch / gates in the captured launch"}
                                                           if cap else evd["capture_note"]),
                    "hbm": hbm}
        if fp64:          # FFT channel: the arithmetic runs on the FP64 pipe; the capture names the load/store data path as the busiest unit
            roofline.update({"bound": "fp64_pipe", "achieved": fp64["achieved"], "peak": fp64["peak"], "unit": fp64["unit"], "frac": fp64["frac"],
                             "algorithmic_slots_per_gate": fp64_slots,
                             "peak_source": dict(evd["fp64_peak"], sms=sms, sm_mhz=mhz),
                             "ncu_fp64_pipe_busy": cap["fp64_pipe_busy"] if cap else None,
                             "ncu_lsu_data_path_busy": cap["lsu_data_path_busy"] if cap else None,
                             "ncu_issue_active": cap["issue_active"] if cap else None,
                             "ncu_l2_to_sm_bytes_per_gate": cap["l2_to_sm_bytes"] / cap["gates_in_launch"] if cap else None,
                             "ncu_source": cap["summary_file"] if cap else evd["capture_note"],
                             "ntt_formulation_equivalent": ({"imad_slots_per_gate": imad_slots, "frac_of_imad_peak": imad["frac"],
                                                             "note": "the roofline of the three-prime NTT kernels this kernel replaces (round-1 review: 0.61): "
                                                                     "their unavoidable IMAD-pipe slots over this kernel's time; no IMAD of that count is executed"}
                                                            if imad else None)})
        elif imad:        # primary bound: the integer multiply pipe; HBM figure kept beside it as the contract asks
            roofline.update({"bound": "imad_pipe", "achieved": imad["achieved"], "peak": imad["peak"], "unit": imad["unit"], "frac": imad["frac"],
                             "algorithmic_slots_per_gate": imad_slots,
                             "peak_source": {k2: peak[k2] for k2 in ("source", "sha256", "lanes_per_clk_sm")},
                             "ncu_fmaheavy_pipe_busy": cap["fmaheavy_pipe_busy"] if cap else None, "ncu_issue_active": cap["issue_active"] if cap else None,
                             "ncu_source": cap["summary_file"] if cap else evd["capture_note"]})
        else:
            roofline.update({k2: hbm[k2] for k2 in ("bound", "achieved", "peak", "unit", "frac")})
        line = {"metric": METRIC.replace("2-party", f"{k}-party"), "value": value, "unit": UNIT, "n_gpus": world * ndev, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "ms_per_bootstrap_amortized": ms / args.steps / G,
                "ms_single_bootstrap_latency": lat, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": ("u32 RNS (four 28-bit-prime NTTs + CRT, exact mod 2^64) / int32 LWE" if big else
                                                "f64 (three-limb folded FFT, rounded to the exact integers mod 2^64) / int32 LWE" if engine == "fft64" else
                                                "u32 RNS (three 28-bit-prime NTTs + CRT, exact mod 2^64) / int32 LWE"), "data": "synthetic",
                "config": {"workload": f"{k}-party NAND x{G} per GPU (mktfhe_parameters_{k}party_3gen: n={n} N={N} l={l} Bg=2^{params.gsw_log2_base} "
                                       f"t={params.ks_decomp_length} Bks=2^{params.ks_log2_base})",
                           "gates_per_step_per_gpu": G,
                           "parallelism": (f"one process, one C-ABI context spanning {ndev} GPUs (mktfhe_create_multi): gate-sharded replicas, key broadcast "
                                           f"{ctx.describe()['key_broadcast']}" if args.abi_multi else f"gate-sharded replicas x{world} (one process per GPU)"),
                           "l2_policy": f"inputs ({2 * G * (k * n + 1) * 4 / 1e6:.0f} MB/step) + keys ({(bsk_stream + ctx.key_buffers()[1][1]) / 1e6:.0f} MB) exceed the 126 MB L2; no flush needed",
                           "keys": "generated by the product host mirror (GPU exact products)" + (", broadcast over NCCL" if world > 1 else
                                   ", broadcast inside mktfhe_finalize_keys" if ndev > 1 else ""), "key_setup_s": round(t_keys, 2)},
                "e2e": {"value": world * GT * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(2 * (GT * k * n + GT) * 4),
                        "d2h_bytes_per_step": int((GT * k * n + GT) * 4)},
                "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks, "decryptions_correct": ok}
        if world * ndev == 1 and not args.no_cpu_baseline and args.parties == 2:
            O, oks = oracle_keyset()
            cores = host_cores()
            cnt = max(cores, 8) * 4
            cpu_sample(O, oks, cores, cores)
            dt, cok = cpu_sample(O, oks, cnt, cores)
            dt1, _ = cpu_sample(O, oks, 2, 1)
            line["cpu_baseline"] = {"value": cnt / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{cnt} NAND gates of the same workload, Float64-FFT restatement of the reference algorithm "
                                              f"(oracle/mk_oracle.c), one gate per thread; single-thread {1e3 * dt1 / 2:.0f} ms/bootstrap",
                                    "decryptions_correct": cok}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()
    return 0


def run_single_key(args):
    """SURVEY 8(f) rank 4, first slice: single-key TFHE (gates.jl:16-22 gate_nand on tfhe_parameters_128, api.jl:100-113) through the
    same engine with one party.  One step = G NAND gates; value with inputs resident in HBM, e2e through the host mirror's gate_nand."""
    import torch
    import torus_fhe_b200 as T
    T1 = T.tfhe1
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device")
    torch.cuda.set_device(0)
    rng = np.random.default_rng(KEY_SEED)
    t0 = time.perf_counter()
    pname = "tfhe_parameters_80" if args.parties == 80 else "tfhe_parameters_128"      # --parties 80 selects the 80-bit set (Torus32 mode)
    sk, ck = T1.make_key_pair(rng, getattr(T1, pname)())
    eng = T1.engine_for(ck, device=0)
    engine = eng.ctx.describe().get("external_product", "ntt_rns")
    t_keys = time.perf_counter() - t0
    G, n = args.gates, sk.params.lwe_size
    bits = rng.integers(0, 2, (2, G)).astype(bool)
    x, y = T1.encrypt(rng, sk, bits[0]), T1.encrypt(rng, sk, bits[1])
    dev = [torch.from_numpy(np.ascontiguousarray(v)).cuda() for v in (x.a.reshape(G, 1, n), x.b, y.a.reshape(G, 1, n), y.b)]
    oa, ob = torch.empty((G, 1, n), dtype=torch.int32, device="cuda"), torch.empty(G, dtype=torch.int32, device="cuda")
    stream = torch.cuda.Stream()
    mu0, mu = int(T1.encode_message(1, 8)), int(T1.encode_message(1, 8)) << 32

    def step_dev():
        eng.ctx.affine_bootstrap_batch_dev(mu0, -1, -1, 0, mu, G, dev[0].data_ptr(), dev[1].data_ptr(), dev[2].data_ptr(), dev[3].data_ptr(), 0, 0,
                                           oa.data_ptr(), ob.data_ptr(), stream=stream.cuda_stream)
    for _ in range(args.warmup):
        step_dev()
    torch.cuda.synchronize()
    sampler = ClockSampler(0); sampler.start()
    l0 = eng.ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_dev()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = eng.ctx.launch_count() - l0
    br_ms, ks_ms = eng.ctx.last_kernel_ms()
    out = T1.gate_nand(ck, x, y)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = T1.gate_nand(ck, x, y)
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()
    ok = bool(np.array_equal(T1.decrypt(sk, out), ~(bits[0] & bits[1])) and np.array_equal(out.b, ob.cpu().numpy()))
    p = sk.params
    print(json.dumps({"metric": f"bootstrapped single-key TFHE NAND gates/sec ({pname})", "value": G * args.steps / (ms * 1e-3), "unit": UNIT,
                      "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": (("f64 (two-limb folded FFT, Torus32 mode)" if args.parties == 80 else "f64 (three-limb folded FFT, Torus32 carried as v << 32)")
                                                    if engine == "fft64" else "u32 RNS (three 28-bit-prime NTTs + CRT)") + " / int32 LWE", "data": "synthetic",
                      "config": {"workload": f"single-key NAND x{G} (api.jl:76-113: n={p.lwe_size} N={p.rlwe_polynomial_degree} l={p.bs_decomp_length} "
                                             f"Bg=2^{p.bs_log2_base} t={p.ks_decomp_length} Bks=2^{p.ks_log2_base}), the 3gen engine with one party",
                                 "key_setup_s": round(t_keys, 2)},
                      "e2e": {"value": G * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(2 * G * (n + 1) * 4), "d2h_bytes_per_step": int(G * (n + 1) * 4)},
                      "gpu_launches": int(launches), "kernel_ms": {"blind_rotate": br_ms, "keyswitch": ks_ms,
                                                                   "keyswitch_fused": bool(eng.ctx.describe()["keyswitch_fused"])},
                      "clocks": clocks, "decryptions_correct": ok}))
    eng.close()
    return 0


def run_ccs(args):
    """SURVEY 8(f) rank 4, second slice: the CCS multi-key scheme (mk_gate_nand, mk_gates.jl:7-13, on mktfhe_parameters_2party, mk_api.jl:4-10)
    composed from batched external products of the same kernels (torus-fhe_b200/tfhe_ccs.py).  One step = G NAND gates from host buffers
    to host buffers (the accumulators live in HBM during the blind rotation)."""
    import torch
    import torus_fhe_b200 as T
    TC = T.tfhe_ccs
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device")
    torch.cuda.set_device(0)
    rng = np.random.default_rng(KEY_SEED)
    params = TC.mktfhe_parameters_2party
    t0 = time.perf_counter()
    secret_keys = [TC.SecretKey(rng, params) for _ in range(2)]
    shared_key = TC.SharedKey(rng, params)
    ck = TC.MKCloudKey([TC.CloudKeyPart(rng, sk, shared_key) for sk in secret_keys], shared_key)
    t_keys = time.perf_counter() - t0
    G = args.gates
    bits = rng.integers(0, 2, (2, G)).astype(bool)
    x, y = TC.mk_encrypt(rng, secret_keys, bits[0]), TC.mk_encrypt(rng, secret_keys, bits[1])
    TC.mk_gate_nand(ck, x[:64], y[:64])                          # warm-up
    sampler = ClockSampler(0); sampler.start()
    l0 = ck.products.ctx.launch_count() + ck.switch.ctx.launch_count()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = TC.mk_gate_nand(ck, x, y)
    dt = (time.perf_counter() - t0) / args.steps
    launches = (ck.products.ctx.launch_count() + ck.switch.ctx.launch_count() - l0) // args.steps
    clocks = sampler.stop()
    ok = bool(np.array_equal(TC.mk_decrypt(secret_keys, out), ~(bits[0] & bits[1])))
    print(json.dumps({"metric": "bootstrapped CCS 2-party MK NAND gates/sec (mktfhe_parameters_2party)", "value": G / dt, "unit": UNIT, "n_gpus": 1,
                      "steps": args.steps, "ms_per_step": 1e3 * dt, "higher_is_better": True, "vs_baseline": None,
                      "dtype": ("f64 (two-limb folded FFT, rounded to the exact integers mod 2^32)" if ck.products.ctx.describe().get("external_product") == "fft64"
                                else "u32 RNS (three 28-bit-prime NTTs + CRT)") + ", Torus32 mode (9-bit gadget digits) / int32 LWE", "data": "synthetic",
                      "config": {"workload": f"CCS NAND x{G} (mk_api.jl:4-10: n={params.lwe_size} N=1024 l=3 Bg=2^9 t=8 Bks=2^2, 2 parties): per blind-rotate "
                                             "step two launches of G x 3 external products + elementwise device ops; host buffers in and out",
                                 "key_setup_s": round(t_keys, 2)},
                      "gpu_launches": int(launches), "clocks": clocks, "decryptions_correct": ok}))
    ck.close()
    return 0


def run_perf_comp(args):
    """The reference's only published measurement protocol (measurements/test_suites/performance_comparison_test/perf_comp.jl:13-20, 107-142;
    its plot is docs/speedup.png): k = 2, 4, 8, 16 parties on the 16-PARTY parameter set, 100 single calls of
    mk_bootstrap_3gen(bk, ks, encode(1, 8), enc(true)), minimum and median of the per-call time.  Here: wall clock of the same call through
    the host mirror (host buffers, H2D + kernel + D2H: what @elapsed sees) and the device time of its kernels beside it."""
    import dataclasses
    import torch
    import torus_fhe_b200 as T
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device")
    torch.cuda.set_device(0)
    ref_s = {2: 0.23, 4: 0.53, 8: 1.06, 16: 2.1}      # read off docs/speedup.png: Julia, one CPU thread, hardware not stated
    rows = []
    for k in (2, 4, 8, 16):
        params = dataclasses.replace(T.mktfhe_parameters_16party_3gen, max_parties=k)
        rng = np.random.default_rng(KEY_SEED + k)
        t0 = time.perf_counter()
        secret_keys, bk, ks = generate_keys(T, params, rng)
        eng = T.engine_for(bk, ks, device=0)
        t_keys = time.perf_counter() - t0
        mu = T.encode_message64(1, 8)
        wall, devt, ok = [], [], True
        for trial in range(args.latency_trials + 3):
            x = T.mk_encrypt_3gen(rng, secret_keys, True)
            t0 = time.perf_counter()
            y = T.mk_bootstrap_3gen(bk, ks, mu, x)
            dt = time.perf_counter() - t0
            ok &= bool(T.mk_decrypt_3gen(secret_keys, y))
            if trial >= 3:
                wall.append(dt * 1e3); devt.append(sum(eng.ctx.last_kernel_ms()))
        rows.append({"parties": k, "blind_rotate_steps": k * params.lwe_size, "wall_ms_min": float(np.min(wall)), "wall_ms_median": float(np.median(wall)),
                     "device_ms_min": float(np.min(devt)), "device_ms_median": float(np.median(devt)), "trials": len(wall),
                     "reference_julia_s": ref_s[k], "speedup_vs_reference_median": ref_s[k] * 1e3 / float(np.median(wall)),
                     "key_setup_s": round(t_keys, 2), "decryptions_correct": ok})
        T.release_engine(bk, ks)
    print(json.dumps({"metric": "single mk_bootstrap_3gen latency on the 16-party parameter set, k = 2/4/8/16 parties (perf_comp.jl protocol)", "unit": "ms",
                      "higher_is_better": False, "n_gpus": 1, "rows": rows,
                      "config": {"workload": "mktfhe_parameters_16party_3gen (n=590 N=2048 l=1 Bg=2^26 t=4 Bks=2^3) with k parties, one bootstrap per call",
                                 "reference_numbers": "docs/speedup.png: 0.23 / 0.53 / 1.06 / 2.1 s (Julia, one CPU thread; hardware not stated)"}}))
    return 0


def run_circuit(args):
    """BASELINE configs[3]: WIDTH-bit ripple-carry adder (mk_add_3gen_v2, 3gen_mk_gates.jl:203-220) on I independent instances,
    one mixed-gate launch per dependency level (1 + 2*WIDTH levels, 5*WIDTH gates per instance)."""
    import torch
    import torus_fhe_b200 as T
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device")
    params = T.mktfhe_parameters_2party_3gen
    rng = np.random.default_rng(KEY_SEED)
    secret_keys, bk, ks = generate_keys(T, params, rng)
    W, I = args.width, args.instances
    a, b = rng.integers(0, 1 << (W - 1), I), rng.integers(0, 1 << (W - 1), I)
    ca, cb = T.mk_int_encrypt_3gen(rng, secret_keys, a, W), T.mk_int_encrypt_3gen(rng, secret_keys, b, W)
    zero = T.mk_encrypt_3gen(rng, secret_keys, np.zeros(I, bool))
    eng = T.engine_for(bk, ks)
    where = "host-resident"
    if args.resident:          # operands uploaded once, every level's gather and launch in HBM, only the result comes back
        ca, cb = [T.MKLweSampleGPU.from_host(c) for c in ca], [T.MKLweSampleGPU.from_host(c) for c in cb]
        zero = T.MKLweSampleGPU.from_host(zero)
        where = "HBM-resident"
    host = (lambda r: r.cpu()) if args.resident else (lambda r: r)
    if args.workload == "less":
        return run_comparator(args, T, rng, secret_keys, bk, ks, eng, a, b, ca, cb, host, where)
    T.mk_add_3gen_v2(bk, ks, [c[:8] for c in ca], [c[:8] for c in cb], zero[:8], W)      # warm-up (small)
    l0 = eng.ctx.launch_count()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = [host(r) for r in T.mk_add_3gen_v2(bk, ks, ca, cb, zero, W)]
    dt = (time.perf_counter() - t0) / args.steps
    got = T.mk_int_decrypt_3gen(secret_keys, res, W)
    exp = ((a + b + (1 << (W - 1))) % (1 << W)) - (1 << (W - 1))
    print(json.dumps({"metric": f"{W}-bit MK adder circuits/sec (2-party, {I} instances batched per level)", "value": I / dt, "unit": "circuits/s",
                      "gates_per_s": 5 * W * I / dt, "levels": 1 + 2 * W, "ms_per_level": 1e3 * dt / (1 + 2 * W), "launches_per_circuit_batch":
                      (eng.ctx.launch_count() - l0) // args.steps, "instances_correct_frac": float(np.mean(got == exp)), "n_gpus": 1,
                      "note": "bootstrapped outputs carry phase noise sigma ~0.026 at the reference's default parameters (same in the oracle's "
                              "Float64-FFT restatement), i.e. ~3e-4 failures per gate fed by bootstrapped inputs: a few % of 16-bit sums differ; the "
                              "GPU path is bit-exact with the exact oracle gate by gate (tests/test_gpu_parity.py)",
                      "config": {"workload": f"mk_add_3gen_v2 WIDTH={W} x {I} instances, {where} ciphertexts between levels"}}))
    return 0


def run_comparator(args, T, rng, secret_keys, bk, ks, eng, a, b, ca, cb, host, where):
    """BASELINE configs[3], the comparator half: mk_less_3gen (3gen_mk_gates.jl:247-255) = sign bit of a - b through mk_sub_3gen:
    WIDTH XOR gates (the inversion of b) in one launch, then the ripple-carry adder's 1 + 2*WIDTH levels; 6*WIDTH gates per instance."""
    W, I = args.width, args.instances
    one = T.mk_encrypt_3gen(rng, secret_keys, np.ones(I, bool))
    if args.resident:
        one = T.MKLweSampleGPU.from_host(one)
    T.mk_less_3gen(bk, ks, [c[:8] for c in ca], [c[:8] for c in cb], one[:8], W)          # warm-up (small)
    l0 = eng.ctx.launch_count()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = host(T.mk_less_3gen(bk, ks, ca, cb, one, W))
    dt = (time.perf_counter() - t0) / args.steps
    got = np.asarray(T.mk_decrypt_3gen(secret_keys, res))
    print(json.dumps({"metric": f"{W}-bit MK less-than circuits/sec (2-party, {I} instances batched per level)", "value": I / dt, "unit": "circuits/s",
                      "gates_per_s": 6 * W * I / dt, "levels": 2 + 2 * W, "launches_per_circuit_batch": (eng.ctx.launch_count() - l0) // args.steps,
                      "instances_correct_frac": float(np.mean(got == (a < b))), "n_gpus": 1,
                      "note": "same per-gate failure rate of the scheme's default parameters as the adder line",
                      "config": {"workload": f"mk_less_3gen WIDTH={W} x {I} instances, {where} ciphertexts between levels"}}))
    return 0


def run_conv(args):
    """BASELINE configs[4]: encrypted convolution layer on a synthetic 28x28 input of WIDTH-bit encrypted ints, one encrypted 3x3
    kernel, stride 1, outputs sharded across ranks (no exchange), ciphertexts resident in HBM between the
    dependency levels.  The timed region covers the H2D copy of the encrypted image/kernel and the D2H copy of this rank's outputs."""
    import torch
    import torch.distributed as dist
    import torus_fhe_b200 as T
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device")
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "NONE"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    params = T.mktfhe_parameters_2party_3gen
    W, H, K = args.width, args.image, 3
    half = 1 << (W - 1)
    eng = T.Engine(params, device=local)
    payload = [None]
    if rank == 0:                                      # keys and encrypted data are made on rank 0 only
        rng = np.random.default_rng(KEY_SEED)
        secret_keys, bk, ks = generate_keys(T, params, rng)
        eng.load_keys([b.gsw_key for b in bk], [q.engine_part() for q in ks])
        T.attach_engine(bk, ks, eng)
        inp, ker = rng.integers(-half, half, (H, H)), rng.integers(-half, half, (1, K, K))
        cin, cker = T.mk_int_encrypt_3gen(rng, secret_keys, inp, W), T.mk_int_encrypt_3gen(rng, secret_keys, ker, W)
        zero = T.mk_encrypt_3gen(rng, secret_keys, False)
        payload = [(secret_keys, inp, ker, cin, cker, zero)]
    if world > 1:
        eng.broadcast_keys(src=0)                      # transformed bsk + ksk over NCCL, once
        dist.broadcast_object_list(payload, src=0)     # the encrypted image / kernel (and the secret keys, for the final check)
        if rank != 0:
            bk, ks = T.RemoteKeys.pair(eng)
    secret_keys, inp, ker, cin, cker, zero = payload[0]
    up = lambda bits: [T.MKLweSampleGPU.from_host(b, local) for b in bits]

    def layer():
        lo, hi, bits = T.enc_conv2d(bk, ks, up(cin), T.MKLweSampleGPU.from_host(zero, local), up(cker), 1, 0, W, shard=(world, rank))
        return lo, hi, [b.cpu() for b in bits]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    small = [c[:8, :8] for c in cin]
    T.enc_conv2d(bk, ks, up(small), T.MKLweSampleGPU.from_host(zero, local), up(cker), 1, 0, W)         # warm-up on an 8x8 corner
    l0 = eng.ctx.launch_count()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        lo, hi, bits = layer()
    barrier()
    dt = (time.perf_counter() - t0) / args.steps
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    got = T.mk_int_decrypt_3gen(secret_keys, bits, W)
    exp = T.conv2d_plain(inp, ker, 1, 0, W).reshape(-1)
    ok = torch.tensor([float(np.sum(got == exp[lo:hi])), float(hi - lo)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ok)
    gates = T.conv2d_gate_count((H, H), (1, K, K), 1, 0, W)
    if rank == 0:
        print(json.dumps({"metric": f"encrypted 3x3 conv layer, {H}x{H} input of {W}-bit ints (2-party): bootstrapped gates/sec", "value": gates / dt,
                          "unit": "gates/s", "n_gpus": world, "steps": args.steps, "layer_s": dt, "outputs": int(exp.size), "outputs_per_s": exp.size / dt,
                          "gates_per_layer": gates, "launches_per_layer": (eng.ctx.launch_count() - l0) // args.steps, "scaling": "strong",
                          "outputs_equal_to_plaintext_frac": float(ok[0].item() / ok[1].item()),
                          "note": "gates fed by bootstrapped inputs fail with p ~ 3e-4 .. 1e-3 at the reference's default parameters (same in the oracle's "
                                  "Float64-FFT restatement), so only a fraction of the outputs (328 gates each at WIDTH 4) decrypts to the plaintext sum; the GPU path "
                                  "is bit-exact with the exact oracle gate by gate, and the layer is exact on a low-noise parameter set "
                                  "(tests/test_gpu_api.py::test_encrypted_conv_layer)",
                          "config": {"workload": f"enc_conv2d {H}x{H} x one 3x3 kernel, WIDTH={W}, outputs sharded over {world} GPU(s), "
                                                 "device-resident ciphertexts between levels, H2D of inputs and D2H of outputs timed"}}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--gates", type=int, default=16384, help="gates per step per GPU")
    ap.add_argument("--parties", type=int, default=2, choices=[2, 3, 4, 5, 8, 16, 80],
                    help="parameter set (BASELINE configs[2]: 4 and 8; 16 = the first N = 2048 set, one gate per SM: use --gates 148 or a multiple)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--abi-multi", action="store_true", help="one process, ONE C-ABI context spanning --gpus GPUs (mktfhe_create_multi) instead of one "
                                                             "process per GPU under torchrun")
    ap.add_argument("--latency-trials", type=int, default=100, help="single-bootstrap latency: min / median over this many calls")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="nand", choices=["nand", "adder", "less", "conv", "single", "perf_comp", "ccs"],
                    help="nand = the headline metric; adder / less = BASELINE configs[3] (adder, comparator); conv = BASELINE configs[4]; "
                         "single = single-key TFHE NAND (tfhe_parameters_128) through the same engine")
    ap.add_argument("--width", type=int, default=None, help="bits per encrypted integer (adder: 16, conv: 4)")
    ap.add_argument("--image", type=int, default=28, help="conv: input height = width")
    ap.add_argument("--instances", type=int, default=1024)
    ap.add_argument("--resident", action="store_true", help="adder / less: keep the ciphertexts in HBM between dependency levels (MKLweSampleGPU)")
    args = ap.parse_args()
    if args.width is None:
        args.width = 4 if args.workload == "conv" else 16
    if args.workload in ("adder", "less"):
        sys.exit(run_circuit(args))
    if args.workload == "conv":
        sys.exit(run_conv(args))
    if args.workload == "single":
        sys.exit(run_single_key(args))
    if args.workload == "perf_comp":
        sys.exit(run_perf_comp(args))
    if args.workload == "ccs":
        sys.exit(run_ccs(args))
    sys.exit(run_reference(args) if args.impl == "reference" else run_ours(args))


if __name__ == "__main__":
    main()
