#!/usr/bin/env python
"""bench.py -- tokens/s of the BioGPT-base `biogpt_eval` hot path on B200 (BASELINE.json metric).

Workload (config.workload): BASELINE.json configs[1] -- BioGPT-base Q4_0, one stream, batch 1,
decode over the whole context (n_past 0 -> seq-1), synthetic weights at the true shapes
(seed 1234), greedy sampling.  One "step" = one such sequence (seq tokens).

  value     device-resident: ids fed back on the GPU (bgpt_cuda_decode_greedy), CUDA events
  e2e       the reference-facing call per token with HOST buffers: bgpt_cuda_eval_topk (token id H2D, the
            top-40 (logit, id) pairs D2H -- what biogpt_sample_top_k_top_p needs) + the host's pick,
            wall clock around the loop
  roofline  the dominant kernel's algorithmic bytes / its CUDA-event duration vs the measured
            HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  the UNMODIFIED reference (oracle/_ref, built from /root/reference by
            oracle/Makefile) timed on the host cores on a bounded sample of the same workload

N > 1 (torchrun): the path does not shard a single sequence -- replicas only.  Every rank
decodes its own independent sequence on its own GPU (no collective on the data path);
value = all ranks' tokens / max-over-ranks time; "scaling": "weak".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

from _bootstrap import load_pkg  # noqa: E402

gf = load_pkg().ggml_file

MODEL_DIR = os.environ.get("BGPT_MODEL_DIR", "/tmp/biogpt_b200_models")
METRIC = "tokens/sec BioGPT-base Q4_0 decode"
UNIT = "tokens/s"


# stdout carries exactly ONE json line: libraries that print to file descriptor 1 (NCCL's version banner, the reference's
# loader) are sent to stderr for the whole run, and the line is written to the saved descriptor at the end
_REAL_STDOUT = None


def capture_stdout() -> None:
    """called first thing in main() (not at import: tools/ import this module for model_path)"""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def model_path(ftype: str) -> str:
    """synthetic BioGPT-base `.bin` (written once per machine; ~7 s f32 synth + ~7 s quantise)"""
    os.makedirs(MODEL_DIR, exist_ok=True)
    p = os.path.join(MODEL_DIR, f"base-{ftype}.bin")
    if not os.path.exists(p):
        tensors = gf.synth_tensors(gf.BASE, seed=1234)
        tmp = p + f".tmp{os.getpid()}"
        gf.write_model(tmp, gf.BASE, tensors, gf.FTYPE_BY_NAME[ftype])
        os.replace(tmp, p)
    return p


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def bytes_per_token(ftype: str, p: int) -> int:
    """algorithmic HBM bytes of one decoded token at position p"""
    W = 345_391_104 * {"f32": 128, "f16": 64, "q4_0": 18, "q4_1": 20, "q5_0": 22, "q5_1": 24, "q8_0": 34}[ftype] // 32
    row = 1024 * {"f32": 128, "f16": 64, "q4_0": 18, "q4_1": 20, "q5_0": 22, "q5_1": 24, "q8_0": 34}[ftype] // 32
    return W + 1_286_144 + 2 * row + 196_608 * (p + 1) + 196_608 + 169_536


class ClockSampler:
    """nvidia-smi SM clock + throttle reasons while the timed region runs"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t.start()
        return self

    def stop(self):
        self._stop.set()
        self._t.join(timeout=6)
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if s[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def workload_config(ftype: str, seq: int, world: int) -> dict:
    """the SAME dict for both arms (the driver compares them)"""
    return {"workload": f"BioGPT-base {ftype} decode, batch 1, seq 1->{seq} (BASELINE.json configs[1])",
            "ftype": ftype, "seq": seq, "streams_per_gpu": 1, "parallelism": f"replicas x{world}",
            "l2": "inputs larger than L2: every token streams the full weight set (+KV) through the cache",
            "parity": "logits bit-identical to the reference CPU path (tests/test_gpu_eval.py)"}


_REF_CACHE = {}


def _ref_handle(ftype: str, threads: int):
    """the reference (oracle/_ref: the UNMODIFIED sources, oracle/Makefile) or, where it was not built, the C port"""
    import ref
    if ftype not in _REF_CACHE:
        path = model_path(ftype)
        _REF_CACHE[ftype] = (ref.Ref(path, n_batch=8, n_threads=threads), "reference") if ref.have_ref() else (ref.Oracle(path), "port")
    R, kind = _REF_CACHE[ftype]
    if kind == "reference":
        R.n_threads = threads
    return R, kind


def _ref_time(R, kind, tok, p, reps) -> float:
    if kind == "reference":
        return R.time_eval_us(tok, p, reps) / 1e6
    t0 = time.perf_counter()
    for _ in range(reps):
        R.eval(tok, p)
    return time.perf_counter() - t0


def best_reference_threads(ftype: str, seq: int):
    """ggml's spin-barrier pool is sensitive to oversubscription: try 4 (the reference's default -t), 8, 16 and nproc threads on a
    short sample at mid-context and keep the fastest (`-t nproc` and `-t 4`)."""
    nproc = os.cpu_count() or 1
    cands = sorted({t for t in (4, 8, 16, nproc) if 1 <= t <= nproc})
    tok = np.array([1234], dtype=np.int32)
    res = {}
    for t in cands:
        R, kind = _ref_handle(ftype, t)
        if kind != "reference":
            return 1, {1: None}
        _ref_time(R, kind, tok, seq // 2, 2)
        res[t] = 8 / _ref_time(R, kind, tok, seq // 2, 8)
    best = max(res, key=res.get)
    return best, {k: round(v, 1) for k, v in res.items()}


def cpu_reference_tokens_per_s(ftype: str, seq: int, budget_s: float, threads: int):
    """the reference's own biogpt_eval on the host cores, sampled at evenly spaced positions of the same decode workload;
    returns (tokens/s, description, kind, seconds actually spent)."""
    positions = [int(round(x)) for x in np.linspace(0, seq - 1, 9)]
    tok = np.array([1234], dtype=np.int32)
    R, kind = _ref_handle(ftype, threads)
    _ref_time(R, kind, tok, 0, 2)  # warm-up (page in the weights, spin up threads)
    first = _ref_time(R, kind, tok, positions[len(positions) // 2], 1)
    reps = min(256, max(1, int(budget_s / max(first, 1e-4) / len(positions))))
    per_tok = []
    t_spent = 0.0
    for p in positions:
        dt = _ref_time(R, kind, tok, p, reps)
        per_tok.append(dt / reps)
        t_spent += dt
    tps = 1.0 / float(np.mean(per_tok))
    sample = (f"{reps} evals of N=1 at each n_past in {positions} ({len(positions) * reps} tokens, "
              f"{t_spent:.1f} s), tokens/s = 1/mean(time per token)")
    return tps, sample, kind, t_spent


def side_configs(capi, device: int, seq: int):
    """BASELINE.json configs[2] (Q8_0 prompt, the reference's n_batch = 8 chunking) and configs[3] (Q5_1, 8 lock-step streams
    per GPU), measured on rank 0 after the headline run; both run on the fused skinny-batch schedule (csrc/bgpt_skinny.cuh).
    Device time = CUDA events inside each call (bgpt_cuda_last_eval_ms)."""
    pk, _ = peaks()
    out = {}
    # ---- configs[2]: `seq` prompt tokens in evals of 8 rows, un-masked, logits of the last row of every eval
    M = capi.Model.load(model_path("q8_0"), device=device, max_batch=8)
    toks = gf.synth_tokens(seq, gf.BASE.n_vocab, seed=5)
    for rep in range(3):
        ms = 0.0; flops = 0.0; nbytes = 0.0
        t0 = time.perf_counter()
        for p in range(0, seq, 8):
            M.eval(toks[p:p + 8], p)
            ms += M.last_eval_ms
            flops += 2.0 * 8 * 301989888 + 2 * 43401216 + 98304.0 * 8 * (p + 8)          # SURVEY 8(d), L_rows = 1
            nbytes += bytes_per_token("q8_0", p + 7) + 7 * (196_608 + 2 * 1088)            # weights once, K/V rows of p+8 positions, 8 rows appended
        wall = time.perf_counter() - t0
    out["prompt_q8_0_n_batch_8"] = {
        "workload": f"BioGPT-base Q8_0, {seq} prompt tokens in {seq // 8} un-masked evals of 8 rows (BASELINE.json configs[2])",
        "tokens_per_s": seq / (ms / 1e3), "tokens_per_s_host_buffers_wall": seq / wall, "ms_total": ms, "tflops": flops / (ms / 1e3) / 1e12,
        "roofline": {"bound": "hbm", "achieved": nbytes / (ms / 1e3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                     "frac": nbytes / (ms / 1e3) / 1e9 / pk["hbm_gbs"], "note": "AI ~ 15 FLOP/B at 8 rows: HBM-bound by SURVEY 8(d)",
                     "traffic": _traffic("prompt_q8_0_n_batch_8")},
        "schedule": "fused skinny-batch" if M.batch_path(8) else "per-operator", "launches_per_eval": None}
    l0 = M.launch_count; M.eval(toks[:8], 0); out["prompt_q8_0_n_batch_8"]["launches_per_eval"] = M.launch_count - l0
    M.close()
    # ---- configs[3]: 8 independent sequences per GPU in lock step, each with its own F32 KV cache.  Device-resident: the greedy loop
    #      runs on the GPU (bgpt_cuda_decode_greedy_streams); host buffers: one bgpt_cuda_eval_streams_topk call per step (40 pairs per stream D2H)
    S = 8
    M = capi.Model.load(model_path("q5_1"), device=device, max_batch=S)
    M.set_streams(S)
    first = gf.synth_tokens(S, gf.BASE.n_vocab, seed=9).astype(np.int32)
    M.decode_greedy_streams(first, 0, seq)
    ids_dev, ms = M.decode_greedy_streams(first, 0, seq)
    nbytes = sum(bytes_per_token("q5_1", p) + (S - 1) * (196_608 * (p + 1) + 196_608 + 169_536 + 2 * 768) for p in range(seq))
    cur = first.copy()
    t0 = time.perf_counter()
    ids_host = []
    for p in range(seq):
        vals, tids, n_out, exact, full = M.eval_streams_topk(cur, p, 40)     # host ids in, 40 (logit, id) pairs per stream out
        cur = tids[:, 0].copy()
        for s_ in np.flatnonzero(exact == 0):
            cur[s_] = int(np.argmax(full[s_]))
        ids_host.append(cur.copy())
    wall = time.perf_counter() - t0
    out["streams_q5_1_x8"] = {
        "workload": f"BioGPT-base Q5_1, {S} lock-step streams on one GPU, seq 1->{seq} (BASELINE.json configs[3], per GPU)",
        "tokens_per_s": S * seq / (ms / 1e3), "tokens_per_s_host_buffers_wall": S * seq / wall, "us_per_step": ms * 1e3 / seq,
        "ids_equal_host_loop": bool(np.array_equal(np.stack(ids_host), ids_dev)),
        "roofline": {"bound": "hbm", "achieved": nbytes / (ms / 1e3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                     "frac": nbytes / (ms / 1e3) / 1e9 / pk["hbm_gbs"], "traffic": _traffic("streams_q5_1_x8")},
        "schedule": "fused skinny-batch" if M.batch_path(S) else "per-operator"}
    M.close()
    # ---- configs[0]: the reference's own CPU-runnable case -- F16, a 4-token prompt as one N = 4 batch, then 15 single-token evals
    #      (-n 16); reported like main.cpp:160 (predict ms and ms per token), the reference on the host cores beside it
    toks4 = gf.synth_tokens(4, gf.BASE.n_vocab, seed=1)
    M = capi.Model.load(model_path("f16"), device=device, max_batch=8)

    def run_ours():
        t0 = time.perf_counter()
        l = M.eval(toks4, 0); ids = [int(np.argmax(l))]
        for i in range(15):
            l = M.eval(np.array([ids[-1]], np.int32), 4 + i); ids.append(int(np.argmax(l)))
        return (time.perf_counter() - t0) * 1e3, ids
    run_ours()
    ours_ms, ours_ids = run_ours()
    gen_f16 = M.decode_generation
    M.close()
    c0 = {"workload": "BioGPT-base F16, 4-token prompt (one N = 4 batch) + 15 single-token evals, n_predict = 16 (BASELINE.json configs[0])",
          "b200_predict_ms": ours_ms, "b200_ms_per_token": ours_ms / 19, "b200_decode_kernel_generation": gen_f16}
    try:
        import ref
        if ref.have_ref():
            threads, _ = best_reference_threads("f16", seq)
            R, kind = _ref_handle("f16", threads)
            best = None
            for rep in range(3):
                t0 = time.perf_counter()
                l = R.eval(toks4, 0); rids = [int(np.argmax(l))]
                for i in range(15):
                    l = R.eval(np.array([rids[-1]], np.int32), 4 + i); rids.append(int(np.argmax(l)))
                dt = (time.perf_counter() - t0) * 1e3
                best = dt if best is None else min(best, dt)
            c0.update({"reference_cpu_predict_ms": best, "reference_cpu_ms_per_token": best / 19, "reference_threads": threads,
                       "ids_identical": rids == ours_ids})
    except Exception as e:
        c0["reference_error"] = repr(e)
    out["f16_cpu_case"] = c0
    # ---- configs[4]: format sweep at n_past 511 (single-token decode): us per token, GB/s of algorithmic bytes, fraction of peak
    sweep = {}
    for ft in ("f16", "q4_0", "q4_1", "q5_0", "q5_1", "q8_0"):
        M = capi.Model.load(model_path(ft), device=device, max_batch=8)
        M.decode_greedy(2, 496, 16)
        _, ms = M.decode_greedy(2, 496, 32)                  # n_past 496..527, mean position 511.5
        us = ms * 1e3 / 32
        b = np.mean([bytes_per_token(ft, p) for p in range(496, 528)])
        sweep[ft] = {"us_per_token": us, "gb_per_s": b / us / 1e3, "frac_of_hbm_peak": b / us / 1e3 / pk["hbm_gbs"],
                     "decode_kernel_generation": M.decode_generation}
        M.close()
    out["format_sweep_n_past_511"] = {"workload": "single-token decode at n_past ~511, all six formats (BASELINE.json configs[4])", "formats": sweep}
    # ---- the reference's UNMODIFIED front end loop (examples/main/main.cpp:93-151): biogpt_eval -- the whole logit row returns to the
    #      host -- then biogpt_sample_top_k_top_p on it (top_k 40, top_p 0.9, temp 0.8 ~ the reference's defaults), one token at a
    #      time, through the C++ API of libbiogpt_b200.so; wall clock over the whole context
    try:
        import ctypes as C
        H = C.CDLL(os.path.join(ROOT, "biogpt.cpp_b200", "host", "libbiogpt_b200.so"))
        H.bgpt_host_open.restype = C.c_void_p
        H.bgpt_host_open.argtypes = [C.c_char_p, C.c_int]
        H.bgpt_host_close.argtypes = [C.c_void_p]
        H.bgpt_host_main_loop.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_uint32, C.c_void_p, C.POINTER(C.c_double)]
        os.environ["BIOGPT_CUDA_DEVICE"] = str(device)
        hs = H.bgpt_host_open(model_path("q4_0").encode(), 8)
        ids_m = np.zeros(seq, np.int32); ws = C.c_double(0)
        H.bgpt_host_main_loop(hs, 2, 0, min(seq, 64), 40, 0.9, 0.8, 1, ids_m.ctypes.data, C.byref(ws))
        rc = H.bgpt_host_main_loop(hs, 2, 0, seq, 40, 0.9, 0.8, 1, ids_m.ctypes.data, C.byref(ws))
        H.bgpt_host_close(hs)
        out["unmodified_main_loop_q4_0"] = {
            "workload": f"BioGPT-base Q4_0, the loop of examples/main/main.cpp:93-151 as written (biogpt_eval with the full logit row to the host + "
                        f"biogpt_sample_top_k_top_p, top_k 40 / top_p 0.9 / temp 0.8), seq 1->{seq}",
            "rc": rc, "tokens_per_s": seq / ws.value, "us_per_token": ws.value / seq * 1e6, "d2h_bytes_per_token": 42384 * 4}
    except Exception as e:
        out["unmodified_main_loop_q4_0"] = {"error": repr(e)}
    # ---- the tcgen05 prompt path: the whole context as ONE un-masked eval (n_batch = seq).  Quantised: bit-exact k_tcw_exact
    #      (warp-specialised, TMA-fed; csrc/bgpt_tcw.cuh); F16: the exact-order SIMT kernels by default, k_tcw_f16 when opted in.
    big = {}
    toks = gf.synth_tokens(seq, gf.BASE.n_vocab, seed=5)
    flops = 2.0 * seq * 301989888 + 2 * 43401216 + 98304.0 * seq * seq
    for ft, opt in (("q8_0", None), ("f16", 0), ("f16", 32)):
        try:
            M = capi.Model.load(model_path(ft), device=device, max_batch=seq)
            if opt is not None:
                M.set_f16_tc_min_rows(opt)
            M.eval(toks, 0); M.eval(toks, 0)
            ms = M.last_eval_ms
            key = ft if opt is None else f"{ft}_tc_{'on' if opt else 'off'}"
            big[key] = {"ms": ms, "tokens_per_s": seq / (ms / 1e3), "tflops": flops / (ms / 1e3) / 1e12,
                        "frac_of_bf16_tensor_peak": flops / (ms / 1e3) / 1e12 / pk.get("bf16_tflops_sustained", 1391.1), "eval_path": M.eval_path(seq),
                        "parity": "bit-identical to the reference" if M.eval_path(seq) != 7 else "tolerance-close (opt-in), see include/bgpt_cuda.h"}
            M.close()
        except Exception as e:                                # never lose the headline line to a side measurement
            big[f"{ft}_{opt}"] = {"error": repr(e)}
    out["prompt_one_eval"] = {"workload": f"BioGPT-base, {seq} prompt tokens in one un-masked eval (tcgen05 matmuls; un-masked f32 attention in exact lane order)",
                              "formats": big}
    return out


def _traffic(key: str):
    """dram__bytes_read + dram__bytes_write per launch of the skinny-batch schedule's matmul kernels, from the committed
    `ncu --set full` capture (profiles/r1_skinny_mm_ncu_summary.json: 8 Q5_1 streams at n_past 511; the matmul kernels read their
    weight slice exactly once, e.g. fc1 = 4096 x 1024 Q5_1 rows = 3,145,728 B algorithmic against 3,211,520 B measured)"""
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "r1_skinny_mm_ncu_summary.json")))
        return {"source": prof["source"], "config": key,
                "launches": [{"kernel": l["kernel"], "grid": l["grid"], "dram_bytes": l["dram_bytes_read"] + l["dram_bytes_write"],
                              "duration_us": l["duration_us"]} for l in prof["launches"]]}
    except Exception:
        return None


def run_streams_workload(args, capi, dist, barrier, rank, local_rank, world):
    """BASELINE.json configs[3]: S independent Q5_1 sequences per GPU decoded in lock step (every weight read serves the S rows; each
    stream has its own F32 KV cache), the model replicated on every rank, no collective on the data path.  One step = one pass over
    the context (seq tokens per stream).  value = all ranks' tokens / max-over-ranks device time (CUDA events inside each call)."""
    ftype = "q5_1" if args.ftype == "q4_0" else args.ftype          # the config names Q5_1; --ftype overrides
    S, seq = args.streams, args.seq
    if rank == 0:
        model_path(ftype)
    if dist is not None:
        dist.barrier()
    M = capi.Model.load(model_path(ftype), device=local_rank, max_batch=S)
    M.set_streams(S)
    first = gf.synth_tokens(S, gf.BASE.n_vocab, seed=9 + rank).astype(np.int32)

    def one_pass():
        cur = first.copy(); ms = 0.0
        for p in range(seq):
            # host token ids in, per stream the 40 best (logit, id) pairs out (selected on the device), next token picked on the host
            vals, tids, n_out, exact, full = M.eval_streams_topk(cur, p, 40)
            ms += M.last_eval_ms
            cur = tids[:, 0].copy()
            for s_ in np.flatnonzero(exact == 0):
                cur[s_] = int(np.argmax(full[s_]))
        return ms
    for _ in range(max(1, args.warmup)):
        one_pass()
    sampler = ClockSampler(local_rank).start()
    barrier()
    l0 = M.launch_count
    t0 = time.perf_counter()
    ms_total = sum(one_pass() for _ in range(args.steps))
    barrier()
    wall = time.perf_counter() - t0
    launches = (M.launch_count - l0) / args.steps
    clocks = sampler.stop()
    if dist is not None:
        import torch
        t = torch.tensor([ms_total, wall], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, wall = float(t[0].item()), float(t[1].item())
    M.close()
    if dist is not None:
        dist.destroy_process_group()
    if rank != 0:
        return
    pk, pk_src = peaks()
    nbytes = sum(bytes_per_token(ftype, p) + (S - 1) * (196_608 * (p + 1) + 196_608 + 169_536 + 2 * 768) for p in range(seq))
    step_ms = ms_total / args.steps
    achieved = nbytes / (step_ms / 1e3) / 1e9
    emit({"metric": "tokens/sec BioGPT-base Q5_1 decode, lock-step streams", "value": world * S * seq * args.steps / (ms_total / 1e3), "unit": UNIT,
          "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
          "vs_baseline": None, "dtype": "int8*int8->f32 (Q8_1 activations), f32 KV", "data": "synthetic",
          "config": {"workload": f"BioGPT-base {ftype}, {S} lock-step streams per GPU x {world} GPUs, seq 1->{seq} (BASELINE.json configs[3])",
                     "ftype": ftype, "seq": seq, "streams_per_gpu": S, "parallelism": f"replicas x{world}",
                     "l2": "inputs larger than L2: every step streams the full weight set and S KV caches"},
          "clocks": clocks,
          "e2e": {"value": world * S * seq * args.steps / wall, "unit": UNIT, "h2d_bytes_per_step": seq * (4 * S + 16), "d2h_bytes_per_step": seq * S * (8 * 40 + 16),
                  "call": "bgpt_cuda_eval_streams_topk per lock-step token (host ids in, top-40 pairs per stream out)"},
          "gpu_launches": launches,
          "roofline": {"bound": "hbm", "kernel": "fused skinny-batch schedule (k_sk_mm / k_sk_attn / k_sk_ln / k_sk_gq), per GPU", "achieved": achieved,
                       "peak": pk["hbm_gbs"], "peak_source": pk_src, "unit": "GB/s", "frac": achieved / pk["hbm_gbs"], "traffic": None},
          "cpu_baseline": None})


def run_streams_cpp_driver(args):
    """BASELINE.json configs[3] through the C++ replica driver (include/bgpt_replicas.h, host/replicas.cpp): ONE process, one host
    thread + one engine per GPU, stream s on device s % G, `--streams` lock-step streams per GPU, every device running its whole
    greedy loop on the GPU.  value = all streams' tokens / max over devices of the CUDA-event time of the device's loop; e2e = the
    same streams through bgpt_replicas_eval (host tokens in, host logits out, host argmax), wall clock."""
    import ctypes as C
    ftype = "q5_1" if args.ftype == "q4_0" else args.ftype
    G, S, seq = args.gpus, args.streams, args.seq
    L = C.CDLL(os.path.join(ROOT, "biogpt.cpp_b200", "host", "libbiogpt_b200.so"))
    L.bgpt_replicas_open.restype = C.c_void_p
    L.bgpt_replicas_open.argtypes = [C.c_char_p, C.c_int, C.c_int]
    L.bgpt_replicas_close.argtypes = [C.c_void_p]
    L.bgpt_replicas_devices.argtypes = [C.c_void_p]
    L.bgpt_replicas_n_vocab.argtypes = [C.c_void_p]
    i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
    f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
    L.bgpt_replicas_eval.argtypes = [C.c_void_p, i32p, C.c_int, f32p]
    L.bgpt_replicas_decode_greedy.argtypes = [C.c_void_p, i32p, C.c_int, C.c_int, i32p, f32p]
    L.bgpt_replicas_eval_topk.argtypes = [C.c_void_p, i32p, C.c_int, C.c_int, f32p, i32p, i32p, i32p, f32p]
    L.bgpt_replicas_last_error.restype = C.c_char_p
    r = L.bgpt_replicas_open(model_path(ftype).encode(), G, G * S)
    if not r:
        raise SystemExit("bench.py: bgpt_replicas_open: " + L.bgpt_replicas_last_error().decode())
    G = L.bgpt_replicas_devices(r)
    n_streams, n_vocab = G * S, L.bgpt_replicas_n_vocab(r)
    first = gf.synth_tokens(n_streams, gf.BASE.n_vocab, seed=9).astype(np.int32)
    ids = np.zeros(seq * n_streams, np.int32)
    dev_ms = np.zeros(G, np.float32)

    def one_pass():
        if L.bgpt_replicas_decode_greedy(r, first, 0, seq, ids, dev_ms) != 0:
            raise SystemExit("bench.py: " + L.bgpt_replicas_last_error().decode())
        return float(dev_ms.max())
    for _ in range(max(1, args.warmup)):
        one_pass()
    sampler = ClockSampler(0).start()
    ms_total = sum(one_pass() for _ in range(args.steps))
    clocks = sampler.stop()
    # host buffers: one bgpt_replicas_eval_topk per lock-step token -- host token ids in, per stream the 40 best (logit, id) pairs out
    # (selected on the device; a stream whose pairs tie also gets its full row), next token picked on the host (a bounded 64-step sample)
    n_e2e = min(seq, 64)
    TOPK = 40
    logits = np.zeros((n_streams, n_vocab), np.float32)
    tv = np.zeros(n_streams * TOPK, np.float32); ti = np.zeros(n_streams * TOPK, np.int32)
    tn = np.zeros(n_streams, np.int32); tex = np.zeros(n_streams, np.int32)
    cur = first.copy()
    L.bgpt_replicas_eval_topk(r, cur, 0, TOPK, tv, ti, tn, tex, logits.reshape(-1))      # untimed: the call's one-time buffers (mapped packets, scratch)
    t0 = time.perf_counter()
    for p in range(n_e2e):
        if L.bgpt_replicas_eval_topk(r, cur, p, TOPK, tv, ti, tn, tex, logits.reshape(-1)) != 0:
            raise SystemExit("bench.py: " + L.bgpt_replicas_last_error().decode())
        cur = ti.reshape(n_streams, TOPK)[:, 0].copy()
        for s_ in np.flatnonzero(tex == 0):
            cur[s_] = int(np.argmax(logits[s_]))
    wall = time.perf_counter() - t0
    same = bool(np.array_equal(cur, ids.reshape(seq, n_streams)[n_e2e - 1]))
    L.bgpt_replicas_close(r)
    pk, pk_src = peaks()
    nbytes = sum(bytes_per_token(ftype, p) + (S - 1) * (196_608 * (p + 1) + 196_608 + 169_536 + 2 * 768) for p in range(seq))
    step_ms = ms_total / args.steps
    achieved = nbytes / (step_ms / 1e3) / 1e9
    emit({"metric": "tokens/sec BioGPT-base Q5_1 decode, lock-step streams", "value": n_streams * seq * args.steps / (ms_total / 1e3), "unit": UNIT,
          "n_gpus": G, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
          "vs_baseline": None, "dtype": "int8*int8->f32 (Q8_1 activations), f32 KV", "data": "synthetic",
          "config": {"workload": f"BioGPT-base {ftype}, {S} lock-step streams per GPU x {G} GPUs, seq 1->{seq} (BASELINE.json configs[3])",
                     "ftype": ftype, "seq": seq, "streams_per_gpu": S, "parallelism": f"replicas x{G}", "driver": "C++ replica driver, one host thread per GPU (no torch, no NCCL)",
                     "l2": "inputs larger than L2: every step streams the full weight set and S KV caches"},
          "clocks": clocks,
          "e2e": {"value": n_streams * n_e2e / wall, "unit": UNIT, "h2d_bytes_per_step": seq * 4 * n_streams, "d2h_bytes_per_step": seq * n_streams * (8 * TOPK + 16),
                  "sample": f"first {n_e2e} lock-step tokens through bgpt_replicas_eval_topk (top-40 pairs per stream)", "ids_equal_device_loop": same},
          "gpu_launches": None, "per_device_ms": [float(x) for x in dev_ms],
          "roofline": {"bound": "hbm", "kernel": "fused skinny-batch schedule (k_sk_mm / k_sk_attn / k_sk_ln / k_sk_gq), per GPU", "achieved": achieved,
                       "peak": pk["hbm_gbs"], "peak_source": pk_src, "unit": "GB/s", "frac": achieved / pk["hbm_gbs"], "traffic": _traffic("streams_q5_1_x8")},
          "cpu_baseline": None})


def run_reference_arm(args):
    """`--impl reference`: the reference's own CPU implementation of the path on this box's host cores.  A "step" is a BOUNDED sample
    of the workload (9 evenly spaced positions of the 1024-token decode, `reps` evals each); ms_per_step is the time that sample
    really took, value = tokens/s over the sample.  Threads: best of {4, 8, 16, nproc}."""
    rank, local_rank, world = dist_env()
    if rank != 0:
        return
    threads, sweep = best_reference_threads(args.ftype, args.seq)
    vals, secs = [], []
    sample = kind = ""
    per_step = args.cpu_budget / max(1, args.steps + args.warmup)
    for i in range(args.warmup + args.steps):
        tps, sample, kind, spent = cpu_reference_tokens_per_s(args.ftype, args.seq, per_step, threads)
        if i >= args.warmup:
            vals.append(tps); secs.append(spent)
    v = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int8*int8->f32 (Q8_0 activations), f32 KV", "data": "synthetic",
            "config": workload_config(args.ftype, args.seq, world),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": kind, "sample": "per step: " + sample,
                             "threads_tried_tokens_per_s": sweep, "host_cores": os.cpu_count()},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def main():
    capture_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ftype", default="q4_0", choices=list(gf.FTYPE_BY_NAME))
    ap.add_argument("--seq", type=int, default=1024)
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--workload", default="decode", choices=["decode", "streams"],
                    help="decode: BASELINE configs[1] (the headline, default); streams: configs[3] -- `--streams` lock-step Q5_1 sequences per GPU, "
                         "replicated over the ranks (64 streams on 8 GPUs)")
    ap.add_argument("--streams", type=int, default=8)
    ap.add_argument("--driver", default="ranks", choices=["ranks", "cpp"],
                    help="streams workload: ranks = one engine per torchrun rank (default); cpp = ONE process drives --gpus devices through the "
                         "C++ replica driver of libbiogpt_b200.so (include/bgpt_replicas.h: one host thread per GPU, no torch)")
    ap.add_argument("--no-extras", action="store_true", help="skip the BASELINE.json configs[2] / configs[3] side measurements")
    args = ap.parse_args()
    rank, local_rank, world = dist_env()

    if args.impl == "reference":
        if rank == 0:
            model_path(args.ftype)
        run_reference_arm(args)
        return

    import importlib
    capi = importlib.import_module("biogpt_cpp_b200.capi")
    if capi.device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device: " + capi.last_error())

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod
        if rank == 0:
            model_path(args.ftype)
        dist.barrier()
    def barrier():
        if dist is not None:
            import torch
            torch.cuda.synchronize()
            dist.barrier()

    if args.workload == "streams" and args.driver == "cpp":
        if rank == 0:
            run_streams_cpp_driver(args)
        if dist is not None:
            dist.barrier(); dist.destroy_process_group()
        return
    if args.workload == "streams":
        run_streams_workload(args, capi, dist, barrier, rank, local_rank, world)
        return

    path = model_path(args.ftype)
    M = capi.Model.load(path, device=local_rank, max_batch=8)
    seq = args.seq
    first_token = 2

    # ---- warm-up
    for _ in range(args.warmup):
        M.decode_greedy(first_token, 0, seq)

    # ---- timed: device-resident decode, CUDA events inside decode_greedy
    sampler = ClockSampler(local_rank).start()
    barrier()
    l0 = M.launch_count
    ms_steps = []
    ids = None
    for _ in range(args.steps):
        ids, ms = M.decode_greedy(first_token, 0, seq)
        ms_steps.append(ms)
    barrier()
    launches = (M.launch_count - l0) / args.steps
    clocks = sampler.stop()
    ms_total = float(np.sum(ms_steps))
    if dist is not None:
        import torch
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    value = world * seq * args.steps / (ms_total / 1e3)

    # ---- e2e: the reference-facing call per token with host buffers
    e2e = None
    if not args.no_e2e:
        TOPK = 40                                    # the reference's default top_k (biogpt.h:113)

        import ctypes as C
        H = C.CDLL(os.path.join(ROOT, "biogpt.cpp_b200", "host", "libbiogpt_b200.so"))
        H.bgpt_host_sampling_loop.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_double)]
        e2e_buf = np.zeros(seq, np.int32)

        def e2e_pass():
            # the token loop of examples/main/main.cpp:93-151 in C++ (host/host_capi.cpp: bgpt_host_sampling_loop): per token ONE
            # bgpt_cuda_eval_topk call with host buffers -- token id in, 40 (logit, id) pairs out -- and the host picks the next token
            rc = H.bgpt_host_sampling_loop(M.h, first_token, 0, seq, TOPK, e2e_buf.ctypes.data, None)
            assert rc == 0, f"bgpt_host_sampling_loop: {rc}"
            return e2e_buf.tolist()
        e2e_pass()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_ids = e2e_pass()
        barrier()
        dt = time.perf_counter() - t0
        if dist is not None:
            import torch
            t = torch.tensor([dt], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        assert ids is None or e2e_ids == ids.tolist(), "device-side greedy ids differ from the host-sampled ids"
        e2e = {"value": world * seq * args.steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": seq * 4, "d2h_bytes_per_step": seq * (8 * TOPK + 8),
               "call": "bgpt_cuda_eval_topk (C ABI; what biogpt_eval_sample calls) once per token from a C++ loop (host_capi: bgpt_host_sampling_loop): host token in, top-40 (logit, id) pairs out, next token picked on the host"}

    if rank != 0:
        M.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel: k_mega, the persistent kernel that IS the decode step
    # (one launch per token).  achieved = algorithmic bytes per launch / CUDA-event duration per
    # launch, both averaged over the seq launches of a step (n_past 0..seq-1).
    pk, pk_src = peaks()
    total_bytes = sum(bytes_per_token(args.ftype, p) for p in range(seq))
    step_ms = ms_total / args.steps
    achieved = total_bytes / (step_ms / 1e3) / 1e9
    traffic = None
    try:   # dram__bytes_read+write of one k_mega launch from the committed ncu capture (profiles/)
        prof = json.load(open(os.path.join(ROOT, "profiles", "r2_mega5_ncu_summary.json" if M.decode_generation == 5 else "r1_mega4_ncu_summary.json")))
        if prof.get("ftype") == args.ftype:
            traffic = {"bytes_per_launch": prof["dram_bytes_per_launch"], "at_n_past": prof["n_past"],
                       "algorithmic_bytes_at_that_n_past": bytes_per_token(args.ftype, prof["n_past"]), "source": prof["source"]}
    except Exception:
        pass
    gen = M.decode_generation
    kname = {5: "k_mega5", 4: "k_mega4", 3: "k_mega"}.get(gen, "per-operator kernels")
    roofline = {"bound": "hbm", "kernel": f"{kname}<{args.ftype}> (persistent decode kernel, generation {gen}, 1 launch per token; mean over n_past 0..{seq - 1})",
                "achieved": achieved, "peak": pk["hbm_gbs"], "peak_source": pk_src, "unit": "GB/s",
                "frac": achieved / pk["hbm_gbs"], "traffic": traffic,
                "algorithmic_bytes_per_launch_mean": total_bytes / seq,
                "us_per_launch_mean": step_ms * 1e3 / seq,
                "note": "latency-bound: a batch-1 step is a chain of 4 all-to-all exchanges through L2 and 2 cluster exchanges per layer "
                        "(~0.8 us each) plus LayerNorm / dot / softmax latencies between them (profiles/README.md, DESIGN.md 4.2); bytes are not the limit"}

    cpu = None
    if not args.no_cpu_baseline:
        threads, sweep = best_reference_threads(args.ftype, seq)
        tps, sample, kind, _ = cpu_reference_tokens_per_s(args.ftype, seq, args.cpu_budget, threads)
        cpu = {"value": tps, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample,
               "threads_tried_tokens_per_s": sweep, "host_cores": os.cpu_count()}

    extras = None
    if not args.no_extras:
        M.close()
        try:
            extras = side_configs(capi, local_rank, seq)
        except Exception as e:                       # the headline line must survive a side measurement
            extras = {"error": repr(e)}
        M = capi.Model.load(path, device=local_rank, max_batch=8)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int8*int8->f32 (Q8_0 activations), f32 KV", "data": "synthetic",
            "config": workload_config(args.ftype, seq, world),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "other_configs": extras}
    emit(line)
    M.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
