#!/usr/bin/env python
"""Headline benchmark: Hamming comparisons/s of the matching hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2]

Workload at every N = BASELINE.json configs[1]: one synthetic pair of 10 000 x 10 000 512-bit descriptors per GPU,
top-2 search + ratio test + cross-check. One "step" = one pass of the hot path over that pair through
ocb_match_top2_device: the n1 x n2 row comparisons are computed once and give both the per-query top-2 (ratio test)
and the per-candidate best query (cross-check). The library picks the engine by size: for this pair the tensor-core
engine (K1T: expansion, search and finish kernels - `value`, `roofline`); the same step on the integer-pipe engine that
north_star specifies (K1: one fused kernel) is timed in the same run (`integer_pipe_engine`, with its own roofline).
Ranks are independent (pairs shard embarrassingly; no data-path collective): weak scaling, value = comparisons of all
ranks / max-over-ranks device time. `secondary` carries configs[2] (MSAC scoring), the dense-stage guided matcher and,
at every N, the sharded configs[3] survey (`c4_survey`: Hilbert partition of the overlap graph, host gather of the match
lists inside its timed region, strong scaling against the single-GPU rate measured in the same job).

  value  device-resident: descriptors already in HBM, CUDA events on the launching stream around each step, the L2
         flushed (256 MiB write) between timed steps;
  e2e    the same step through the reference-facing C++ entry point (match_features_subset on std::vector<feature_2d>
         held on the host + cross-check flags): row packing, H2D, kernels, D2H, double ratio test, std::sort - wall clock,
         `--callers` concurrent OpenMP callers (the reference's run_parallel executes one closure per pair on all its
         workers), at least eight calls each; the single caller is reported next to it;
  cpu_baseline / --impl reference: the reference's own match_features.cpp object code (oracle/_ref, -mpopcnt build;
         the as-shipped build is reported next to it) under the reference's OpenMP-over-pairs driver on the box's host cores.
"""
import argparse
import json
import os
import sys
import threading
import time

# The host side of the survey path runs OpenMP teams next to threads that wait for the GPU, with as few as four cores
# per rank on an 8-GPU box: idle OpenMP workers must sleep, not spin, or they take the cores the working threads need.
# (libgomp reads this when it is loaded, i.e. before the first import below that pulls it in.)
os.environ.setdefault("OMP_WAIT_POLICY", "passive")

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N1 = N2 = 10000
METRIC = "hamming_comparisons_per_s"
UNIT = "Gcmp/s"
WORKLOAD = "configs[1]: synthetic pair 10000x10000 512-bit descriptors, top-2 + ratio test + cross-check"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.mhz, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.mhz.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.004)

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.mhz)) if self.mhz else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.mhz)}


def k1_profile(running_variant):
    """The instruction mix, DRAM traffic and pipe utilisation of the headline kernel are NOT typed in here: they come
    from profiles/k1_roofline.json, written by `tools/ncu_summary.py roofline` from an `ncu --set full` capture of this
    same command plus the SASS of the built library, and carrying the kernel name, the variant and the hash of the
    kernel source they belong to. A record that does not belong to the kernel that is running now is not used."""
    import hashlib
    path = os.path.join(ROOT, "profiles", "k1_roofline.json")
    if not os.path.exists(path):
        return None, "profiles/k1_roofline.json is missing"
    rec = json.load(open(path))
    if rec.get("k1_variant") != running_variant:
        return None, f"profiles/k1_roofline.json describes k1_variant {rec.get('k1_variant')}, running {running_variant}"
    for f, want in rec.get("source_sha256", {}).items():
        got = hashlib.sha256(open(os.path.join(ROOT, "opencalibration_b200", "csrc", f), "rb").read()).hexdigest()
        if got != want:
            return None, f"profiles/k1_roofline.json was taken with another {f}"
    if not rec.get("cross_check"):
        return None, "profiles/k1_roofline.json describes the kernel without the fused cross-check"
    return rec, None


def roofline(achieved, sm_max, held, peak_src, ms_per_step, peaks, running_variant):
    """Bound = the integer pipes (north_star: XOR + POPC, not HBM, not tensor cores). `peak` is the PLAIN form's
    POPC-pipe roofline (16 XOR + 16 POPC per comparison) so that numbers stay comparable between
    variants; `mix_peak` is the tighter pipe bound of the instruction mix that actually runs."""
    popc_peak = 148 * 16 * sm_max * 1e6 / 16 / 1e9  # G comparisons/s: 16 POPC/clk/SM, 16 POPC per comparison
    algo_bytes = (N1 + N2) * 64 + N1 * 8 + N2 * 4
    prof, why_not = k1_profile(running_variant)
    out = {"bound": "int-pipe (popc/alu)", "achieved": achieved, "peak": popc_peak, "unit": UNIT,
           "frac": achieved / popc_peak, "traffic": prof["ncu"]["dram_bytes"] if prof else None,
           "peak_source": f"148 SMs x 16 POPC/clk/SM (probed on this pool: 16.0) x {sm_max:.0f} MHz "
                          f"({peak_src} sm_max_mhz) / 16 POPC per comparison (plain XOR+POPC form)",
           "frac_at_held_clock": achieved / (148 * 16 * held * 1e6 / 16 / 1e9),
           "hbm": {"achieved_gbs": algo_bytes / (ms_per_step * 1e-3) / 1e9, "peak_gbs": peaks.get("hbm_gbs"),
                   "algorithmic_bytes_per_step": algo_bytes,
                   "measured_dram_bytes_per_step": prof["ncu"]["dram_bytes"] if prof else None}}
    if prof:
        alu, xu = prof["sass"]["alu_ops_per_cmp"], prof["sass"]["xu_ops_per_cmp"]
        mix_peak = 148 * sm_max * 1e6 / max(alu / 64.0, xu / 16.0) / 1e9  # SM clocks per comparison of the mix that runs
        out.update({"mix_peak": mix_peak, "mix_frac": achieved / mix_peak,
                    "mix": {"alu_ops_per_cmp": alu, "xu_ops_per_cmp": xu, "fma_ops_per_cmp": prof["sass"]["fma_ops_per_cmp"],
                            "counted_in": f"SASS inner loop {prof['sass']['inner_loop']} of {prof['kernel']}",
                            "note": "prefix carry-save form trades POPCs for LOP3s, so it exceeds the plain-POPC "
                                    "roofline; mix_peak = 148 SMs x clock / max(alu/64, xu/16)"},
                    "ncu_pipes": {k: prof["ncu"][k] for k in ("xu_pct_of_peak", "alu_pct_of_peak", "fma_pct_of_peak",
                                                              "issue_slots_pct", "time_us")},
                    "profile": {"file": "profiles/k1_roofline.json", "kernel": prof["kernel"],
                                "k1_variant": prof["k1_variant"], "report": prof["report"],
                                "git_rev": prof["git_rev_of_capture_summary"]}})
    else:
        out.update({"mix_peak": None, "mix_frac": None, "profile": None, "profile_unusable": why_not})
    return out


def tensor_roofline(achieved_gcmp, ms_per_step, peaks, peak_src, sm_max, int_roofline, mma_cycles=None):
    """K1T: the search as an exact s8 contraction on the tensor cores. Algorithmic work per comparison = 512 multiply-
    accumulates = 1024 integer operations (SURVEY 8d counts the same comparison as 16 XOR + 16 POPC on the integer
    pipes). Peak = dense s8 rate = 2 x the measured bf16 GEMM rate of MEASURED_PEAKS.json (the tensor pipe runs 8-bit
    operands at twice the 16-bit rate; nominal 4.5 vs 2.25 P). `achieved` uses the whole step (expansion, search and
    finish kernels), so it is a lower bound for the search kernel alone."""
    ops = N1 * N2 * 1024.0
    achieved = ops / (ms_per_step * 1e-3) / 1e12
    bf16 = float(peaks.get("bf16_tflops", 1590.0))
    peak = 2.0 * bf16
    import hashlib
    prof, why_not = None, "profiles/k1t_roofline.json is missing"
    path = os.path.join(ROOT, "profiles", "k1t_roofline.json")
    if os.path.exists(path):
        prof, why_not = json.load(open(path)), None
        src = os.path.join(ROOT, "opencalibration_b200", "csrc", "hamming_tensor.cu")
        if hashlib.sha256(open(src, "rb").read()).hexdigest() != prof.get("source_sha256", {}).get("hamming_tensor.cu"):
            prof, why_not = None, "profiles/k1t_roofline.json was taken with another hamming_tensor.cu"
    algo_bytes = (N1 + N2) * 64 + N1 * 8 + N2 * 4
    return {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "traffic": prof["ncu"]["dram_bytes_per_step"] if prof else None,
            "op": "s8 multiply-accumulate = 2 integer operations, accumulated exactly in s32 (tcgen05.mma.kind::i8)",
            "peak_source": f"2 x bf16_tflops ({bf16:.1f}, {peak_src} MEASURED_PEAKS.json, burst): dense 8-bit rate of the "
                           f"tensor pipe",
            "limiter": "the latency of dependent tcgen05.mma on one accumulator (~150 cycles whatever N is: "
                       "ocb_probe_umma, `mma_cycles` below) with 256 of the 512 tensor-memory columns holding the query "
                       "operand: two accumulators of 64 columns per buffer retire an M128 x N64 x K32 MMA every ~64 cycles "
                       "against 32 of arithmetic; next the L2 -> SM stream (every group of 256 query rows streams all "
                       f"candidate tiles: {(N1 + 255) // 256} x {N2} x 512 B = {(N1 + 255) // 256 * N2 * 512 / 1e6:.0f} MB per "
                       "step) and the epilogue's issue slots (DESIGN.md section 3)",
            "l2_stream_TBps": (N1 + 255) // 256 * N2 * 512 / (ms_per_step * 1e-3) / 1e12,
            "mma_cycles": mma_cycles,
            "popc_roofline_equivalent": {"achieved": achieved_gcmp, "peak": int_roofline["peak"], "unit": UNIT,
                                         "frac": achieved_gcmp / int_roofline["peak"],
                                         "note": "the same step against the plain XOR+POPC roofline of SURVEY 8d (the "
                                                 "target of north_star); above 1 because the work left the POPC pipe"},
            "hbm": {"achieved_gbs": algo_bytes / (ms_per_step * 1e-3) / 1e9, "peak_gbs": peaks.get("hbm_gbs"),
                    "algorithmic_bytes_per_step": algo_bytes,
                    "measured_dram_bytes_per_step": prof["ncu"]["dram_bytes_per_step"] if prof else None},
            "profile": prof, "profile_unusable": why_not}


def _timed(torch, fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def secondary_measurements(torch, capi, synthetic, stream):
    """The other configurations of BASELINE.json, reported next to the headline (not bench lines of their own):
    configs[2] RANSAC scoring 4096 hypotheses x 20000 correspondences (K2/K3, device resident), and a slice of
    configs[3]: batched pairs of a synthetic aerial grid through ocb_match_pairs (descriptors resident, result
    lists copied back)."""
    out = {}
    # ---- configs[2]
    H, N = 4096, 20000
    from opencalibration_b200 import host
    corr, _H = synthetic.homography_scene(14000, 6000, seed=42)
    d_corr7 = torch.from_numpy(corr).cuda()
    d_corr4 = torch.zeros(N * 4, dtype=torch.float64, device="cuda")
    order = np.random.default_rng(0).permutation(N).astype(np.uint32)
    d_order = torch.from_numpy(order.view(np.int32)).cuda()
    d_pos = torch.zeros(N, dtype=torch.int32, device="cuda")
    capi.prepare_correspondences_device(d_corr7.data_ptr(), d_order.data_ptr(), N, d_corr4.data_ptr(), d_pos.data_ptr(),
                                        stream)
    d_score = torch.zeros(H, dtype=torch.float64, device="cuda")
    d_count = torch.zeros(H, dtype=torch.int32, device="cuda")
    for kind, name, thr in ((0, "homography", 0.005), (1, "epipolar", 0.01)):
        if kind == 0:  # hypotheses fitted to random 4-samples, as the RANSAC driver produces them (~24% all-inlier)
            rng = np.random.default_rng(5)
            models = np.stack([host.fit(0, corr, rng.choice(N, 4, replace=False).astype(np.uintp)) for _ in range(H)])
        else:
            models = synthetic.random_models(kind, H, seed=3)
        d_models = torch.from_numpy(np.ascontiguousarray(models)).cuda()
        ms = _timed(torch, lambda: capi.score_models_device(kind, d_models.data_ptr(), H, d_corr4.data_ptr(), None, N,
                                                            thr, d_score.data_ptr(), d_count.data_ptr(), None, stream), 5)
        out[f"ransac_scoring_{name}"] = {"workload": "configs[2]: 4096 hypotheses x 20000 correspondences, MSAC in "
                                                     "evaluation order, fp64 bit-exact", "ms": ms,
                                         "M_residuals_per_s": H * N / (ms * 1e-3) / 1e6}
    # ---- configs[3] slice
    rows, cols, n_desc = 6, 8, 8192
    images, _pos, pairs = synthetic.grid_survey(rows, cols, n_desc, seed=7)
    for i, im in enumerate(images):
        capi.register_descriptors(1000 + i, im)
    plist = [(1000 + a, 1000 + b) for a, b in pairs]
    nq = [n_desc] * len(plist)
    res = np.zeros(len(plist) * n_desc, capi.TOP2_DTYPE)
    capi.match_pairs(plist, nq, out=res)  # untimed: sizes the thread's device / page-locked staging areas
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        capi.match_pairs(plist, nq, out=res)
    secs = (time.perf_counter() - t0) / reps
    for i in range(len(images)):
        capi.unregister_descriptors(1000 + i)
    out["grid_pairs_batched"] = {"workload": f"configs[3] slice: {rows}x{cols} image grid, {len(plist)} directed pairs "
                                             f"x {n_desc}x{n_desc} rows, one ocb_match_pairs submission, results to host",
                                 "pairs_per_s": len(plist) / secs, "Gcmp_per_s": len(plist) * n_desc * n_desc / secs / 1e9,
                                 "ms": secs * 1e3}
    # ---- dense-stage guided matcher: K4 on candidate lists, device resident
    n_q, n_c = 60000, 20000
    w = synthetic.guided_visits(n_q, n_c, seed=3)
    total = int(w["begin"][-1])
    d_q4, d_c4 = torch.from_numpy(w["q"].view(np.int64)).cuda(), torch.from_numpy(w["c"].view(np.int64)).cuda()
    d_lq = torch.arange(n_q, dtype=torch.int32, device="cuda")
    d_lb = torch.from_numpy(w["begin"].view(np.int64)).cuda()
    d_lc = torch.from_numpy(w["nearby"].view(np.int32)).cuda()
    d_o = torch.zeros(n_q, dtype=torch.int64, device="cuda")
    ms = _timed(torch, lambda: capi.match_lists_device(d_q4.data_ptr(), d_c4.data_ptr(), d_lq.data_ptr(), d_lb.data_ptr(),
                                                       d_lc.data_ptr(), n_q, d_o.data_ptr(), stream), 20)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oc_oracle as O
    orc = O.Oracle()
    lq = np.arange(n_q, dtype=np.uint32)
    bp, bd, sd, _good = orc.match_lists(w["q"], w["c"], lq, w["begin"], w["nearby"])
    got = d_o.cpu().numpy().view(capi.TOP2_DTYPE)
    to_int = lambda x: np.where(np.isinf(x), 0xFFFF, np.rint(x * 486)).astype(np.uint16)
    parity = bool(np.array_equal(got["best_k"], bp) and np.array_equal(got["best_d"], to_int(bd)) and
                  np.array_equal(got["second_d"], to_int(sd)))
    cpu_s = min(orc.bench_match_lists(w["q"], w["c"], lq, w["begin"], w["nearby"]) for _ in range(3))
    algo_bytes = total * 68 + n_q * (64 + 4 + 8 + 8)
    out["dense_guided_lists"] = {"workload": f"dense stage (src/dense/dense_stereo.cpp:244-281): {n_q} source features x "
                                             f"candidate lists within 150 px among {n_c} features of one image, "
                                             f"{total} comparisons", "ms": ms, "Gcmp_per_s": total / (ms * 1e-3) / 1e9,
                                 "gathered_GB_per_s": algo_bytes / (ms * 1e-3) / 1e9, "parity_vs_oracle": parity,
                                 "cpu_port_Gcmp_per_s": total / cpu_s / 1e9, "cpu_cores": orc.num_procs()}
    return out



def c2_config(pairs_per_s, l2, k1_variant, parity_full, survivors, cross_check):
    """The SAME keys in both arms (ours / --impl reference); values that only one arm has are null in the other."""
    return {"workload": WORKLOAD, "n1": N1, "n2": N2, "comparisons_per_pair": N1 * N2, "cross_check": cross_check,
            "pairs_per_s": pairs_per_s, "l2": l2, "k1_variant": k1_variant, "parity_full": parity_full,
            "ratio_test_survivors": survivors}


def cpu_reference_run(steps, warmup, threads=0):
    """The reference's CPU path on the SAME configuration as the product arm: every step matches `threads` pairs (one
    pair per OpenMP worker, src/pipeline/pipeline.cpp:42-49), each the full N1 x N2 rows of configs[1] (different
    seeds per worker), through the reference's own match_features.cpp object code (oracle/_ref, -mpopcnt build)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oc_oracle as O
    from opencalibration_b200 import synthetic
    kind = "reference"
    try:
        ref, ref_shipped = O.Reference(popcnt=True), O.Reference(popcnt=False)
    except (FileNotFoundError, OSError):
        ref, ref_shipped, kind = O.Oracle(), None, "port"
    cores = ref.num_procs() if threads <= 0 else threads
    qs, cs = [], []
    for t in range(cores):
        a, b = synthetic.config2_pair(N1, N2, seed=1 + t)  # worker 0 matches the product arm's rank-0 pair
        qs.append(a)
        cs.append(b)
    q, c = np.concatenate(qs), np.concatenate(cs)
    cmp_per_step = cores * N1 * N2
    for _ in range(max(1, warmup)):
        ref.bench_match_pairs(q[:cores * 500], c[:cores * 500], cores, 500, 500, cores)
    runs = [ref.bench_match_pairs(q, c, cores, N1, N2, cores) for _ in range(steps)]
    secs = [r[0] for r in runs]
    value = cmp_per_step / (sum(secs) / len(secs)) / 1e9
    shipped = None
    if ref_shipped is not None:  # the flags the reference ships with (no -mpopcnt): a quarter of the queries is enough
        s = ref_shipped.bench_match_pairs(q[:cores * N1 // 4], c, cores, N1 // 4, N2, cores)[0]
        shipped = cores * (N1 // 4) * N2 / s / 1e9
    return dict(value=value, unit=UNIT, cores=cores, kind=kind,
                sample=f"{cores} pairs per step (one per OpenMP thread), each the full {N1}x{N2} rows of the workload, "
                       f"{steps} steps, -mpopcnt build of the reference's own match_features.cpp",
                as_shipped_flags_value=shipped, ms_per_step=1e3 * sum(secs) / len(secs),
                pairs_per_s=cores / (sum(secs) / len(secs)), survivors_per_pair=runs[-1][1] / cores)


def run_reference_survey(args):
    """CPU arm of --workload c4 / c5: the reference's own match_features.cpp object code (oracle/_ref, -mpopcnt) on
    pairs of the same synthetic survey, one pair per OpenMP worker like run_parallel; a bounded sample of the pair
    list (every pair costs 8192 x 8192 comparisons), same unit: pairs matched per second."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oc_oracle as O
    from opencalibration_b200 import synthetic
    kind = "reference"
    try:
        ref = O.Reference(popcnt=True)
    except (FileNotFoundError, OSError):
        ref, kind = O.Oracle(), "port"
    cores = ref.num_procs()
    rows, cols = (25, 40) if args.workload == "c4" else (50, 100)
    survey = synthetic.PlanarSurvey(rows, cols, 8192, seed=7)
    sample = survey.pairs[:: max(1, len(survey.pairs) // cores)][:cores]
    cache = {}

    def desc(i):
        if i not in cache:
            cache[i] = survey.image(i)[0]
        return cache[i]

    q = np.concatenate([desc(a) for a, _ in sample])
    c = np.concatenate([desc(b) for _, b in sample])
    ref.bench_match_pairs(q[:cores * 256], c[:cores * 256], cores, 256, 256, cores)
    steps = max(1, min(args.steps, 3))
    secs = [ref.bench_match_pairs(q, c, len(sample), 8192, 8192, cores)[0] for _ in range(steps)]
    s = sum(secs) / len(secs)
    value = len(sample) / s
    line = {"impl": "reference", "metric": "image_pairs_matched_per_s", "value": value, "unit": "pairs/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": 1, "ms_per_step": s * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"configs[{3 if args.workload == 'c4' else 4}]: {rows}x{cols} image survey, "
                                   f"8192 features per image, match_features_subset per pair",
                       "Gcmp_per_s": len(sample) * 8192 * 8192 / s / 1e9},
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": kind,
                             "sample": f"{len(sample)} pairs of the survey's pair list (one per OpenMP thread), "
                                       f"all 8192 x 8192 rows each, -mpopcnt build"},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload != "c2":
        return run_reference_survey(args)
    steps = max(1, min(args.steps, 40))  # a step is `cores` full pairs (~0.5 s): bounded so the arm ends in a minute
    r = cpu_reference_run(steps, max(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": c2_config(r["pairs_per_s"], "n/a (CPU arm)", None, None, int(r["survivors_per_pair"]),
                                "not computed: src/match/match_features.cpp has no cross-check"),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "as_shipped_flags_value")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def run_ours(args):
    import torch
    from opencalibration_b200 import capi, host, synthetic

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    capi.init(local_rank)
    peaks, peak_src = measured_peaks()

    # ---- inputs: this rank's pair, resident in HBM
    a, b = synthetic.config2_pair(N1, N2, seed=1 + rank)
    dq = torch.from_numpy(a.view(np.int64)).cuda()
    dc = torch.from_numpy(b.view(np.int64)).cuda()
    dout = torch.zeros(N1, dtype=torch.int64, device="cuda")
    dcol = torch.zeros(N2, dtype=torch.int32, device="cuda")
    wsb = capi.match_top2_workspace_bytes(N1, N2, True)
    ws = torch.zeros(wsb + 512, dtype=torch.uint8, device="cuda")
    wsp = (ws.data_ptr() + 255) // 256 * 256
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    cmp_per_step = N1 * N2  # every (query, candidate) row pair is compared once; cross-check is fused in the sweep

    def step():
        capi.match_top2_device(dq.data_ptr(), N1, dc.data_ptr(), N2, dout.data_ptr(), dcol.data_ptr(), wsp, wsb, stream)

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()

    def timed(engine):
        """W warm-up steps, then K steps with the L2 flushed before each, CUDA events on the launching stream, max over
        ranks. engine: 0 = the library's choice (tensor cores for this pair), 1 = integer pipes, 2 = tensor cores."""
        capi.set_option("k1_engine", engine)
        for _ in range(max(args.warmup, 3)):
            step()
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        launches0 = capi.kernel_launches()
        barrier()
        for e0, e1 in ev:
            flush.fill_(1)  # evict L2 between timed steps (outside the event bracket)
            e0.record()
            step()
            e1.record()
        barrier()
        n_launch = capi.kernel_launches() - launches0
        dev_ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
        t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
        if dist:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        capi.set_option("k1_engine", 0)
        return float(t.item()) / args.steps, n_launch, (dout.cpu().numpy().view(capi.TOP2_DTYPE).copy(),
                                                        dcol.cpu().numpy().copy())

    # the integer-pipe engine (K1: what north_star specifies) first, then the headline: the library's own choice for a
    # pair of this size, the tensor-core engine (K1T); both are checked against the oracle below
    # (the clock sampler covers both timed regions and the end-to-end measurement below: a single timed region of the
    # tensor engine lasts two milliseconds, less than one NVML query)
    sampler = ClockSampler(local_rank)
    sampler.start()
    int_ms, int_launches, int_result = timed(1)
    ms_per_step, launches, (r_timed, col_timed) = timed(0)
    headline_engine = "tensor cores (K1T)" if launches >= 3 * args.steps else "integer pipes (K1)"
    value = world * cmp_per_step / (ms_per_step * 1e-3) / 1e9

    # ---- e2e: the reference-facing C++ entry point on host vectors (gather, H2D, kernel, D2H, ratio test, sort).
    # (a) one caller, one pair at a time; (b) the reference's calling convention: run_parallel executes one
    # match_features_subset closure per pair on OpenMP workers (src/pipeline/pipeline.cpp:42-49), so several calls
    # are in flight and one call's host work overlaps another's kernel. Every call copies its inputs and results.
    fa, fb = host.FeatureSet(a), host.FeatureSet(b)
    idx1, idx2 = np.arange(N1, dtype=np.uintp), np.arange(N2, dtype=np.uintp)
    matcher = host.Matcher(N1)
    for _ in range(3):
        n_matches = matcher(fa, fb, idx1, idx2, cross_check=True)
    e2e_steps = max(3, min(args.steps, 50))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        n_matches = matcher(fa, fb, idx1, idx2, cross_check=True)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_single_ms = float(e2e_s.item()) * 1e3 / e2e_steps
    callers = max(1, min(args.callers, (os.cpu_count() or 1) // max(world, 1)))
    host.run_parallel_handles([fa] * callers, [fb] * callers, threads=callers, cross_check=True, reps=2)
    # every caller makes at least eight calls (with two the start-up of the workers is a third of the region)
    conc_steps = callers * max(8, min(args.steps, 200) // callers)
    barrier()
    conc_secs, _ = host.run_parallel_handles([fa] * callers, [fb] * callers, threads=callers, cross_check=True,
                                             reps=conc_steps // callers)
    barrier()
    e2e_s = torch.tensor([conc_secs], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_s.item()) * 1e3 / conc_steps
    e2e_value = world * cmp_per_step / (e2e_ms * 1e-3) / 1e9
    clocks = sampler.result()
    int_clocks = clocks
    secondary = None
    if rank == 0 and world == 1 and not args.no_secondary:
        secondary = secondary_measurements(torch, capi, synthetic, stream)
    # ---- the sharded survey (BASELINE.json configs[3]) at THIS N: the overlap-graph partition north_star names, with the
    # host gather of the match lists inside its timed region; every rank takes part
    del dq, dc, dout, dcol, ws, flush
    torch.cuda.empty_cache()
    survey_block = None
    if not args.no_survey:
        survey_block = measure_survey(torch, dist, rank, local_rank, world, "c4", steps=3, warmup=1, with_ransac=False,
                                      verify=not args.no_verify)

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    # ---- parity of what was timed: EVERY query and EVERY candidate column of the timed pair against the checker (all
    # host cores, about a second) + CPU baseline on this box
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oc_oracle as O
    orc = O.Oracle()
    bk, bd, sd = orc.match_top2(a, b)
    col_want = orc.match_col_best(a, b)

    def equal(rec, col):
        return bool(np.array_equal(rec["best_k"], bk) and np.array_equal(rec["best_d"], bd) and
                    np.array_equal(rec["second_d"], sd) and np.array_equal(col, col_want))
    parity = equal(r_timed, col_timed)
    int_parity = equal(*int_result)
    assert parity and int_parity, "a timed kernel's results differ from the oracle's"
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_reference_run(steps=4, warmup=1)
        cpu = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "as_shipped_flags_value")}

    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    achieved = cmp_per_step / (ms_per_step * 1e-3) / 1e9 if world == 1 else value / world
    held = clocks.get("sm_mhz") or sm_max
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic",
        "config": c2_config(world / (ms_per_step * 1e-3), "flushed between timed steps (256 MiB write)",
                            capi.get_option("k1_variant"), parity, int(n_matches),
                            "fused column-wise best query in the same sweep (no second pass)"),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": (N1 + N2) * 64,
                "d2h_bytes_per_step": N1 * 8 + N2 * 4, "ms_per_step": e2e_ms, "steps": conc_steps,
                "api": "match_features_subset(std::vector<feature_2d>...) + cross-check flags via libocb_host.so; "
                       "one step = one call on one pair, inputs in pageable host vectors, copies inside the call",
                "concurrent_callers": callers,
                "calling_convention": "OpenMP workers, one closure per pair, like the reference's run_parallel "
                                      "(src/pipeline/pipeline.cpp:42-49)",
                "pairs_per_s": world / (e2e_ms * 1e-3),
                "single_caller": {"value": world * cmp_per_step / (e2e_single_ms * 1e-3) / 1e9,
                                  "ms_per_step": e2e_single_ms, "steps": e2e_steps}},
        "gpu_launches": int(launches),
        "engine": headline_engine,
    }
    int_achieved = cmp_per_step / (int_ms * 1e-3) / 1e9
    int_roofline = roofline(int_achieved, sm_max, int_clocks.get("sm_mhz") or sm_max, peak_src, int_ms, peaks,
                            capi.get_option("k1_variant"))
    if headline_engine.startswith("tensor"):
        # the tensor pipe measured by itself on this GPU: cycles per M128 x N x K32 s8 MMA (arithmetic: N / 2 cycles)
        probes = {"a_in_tmem_n64_2acc": (1, 64, 2), "a_in_tmem_n64_4acc": (1, 64, 4), "a_in_tmem_n128_2acc": (1, 128, 2),
                  "a_in_tmem_n256_1acc": (1, 256, 1), "a_in_smem_n128_2acc": (0, 128, 2), "a_in_smem_n256_1acc": (0, 256, 1)}
        mma_cycles = {k: round(capi.probe_umma(*v), 1) for k, v in probes.items()}
        line["roofline"] = tensor_roofline(achieved, ms_per_step, peaks, peak_src, sm_max, int_roofline, mma_cycles)
        line["dtype"] = "s8 x s8 -> s32 (exact)"
    else:
        line["roofline"] = int_roofline
    line["integer_pipe_engine"] = {
        "what": "the same step on the engine north_star specifies (K1: XOR + POPC on the integer pipes, one fused kernel)",
        "value": world * int_achieved, "unit": UNIT, "ms_per_step": int_ms, "gpu_launches": int(int_launches),
        "parity_full": int_parity, "clocks": int_clocks, "roofline": int_roofline}
    if cpu:
        line["cpu_baseline"] = cpu
    if survey_block:
        secondary = dict(secondary or {})
        secondary["c4_survey"] = survey_block
    if secondary:
        line["secondary"] = secondary
    emit(line)
    if dist:
        dist.destroy_process_group()


def measure_survey(torch, dist, rank, local_rank, world, workload, steps, warmup, with_ransac, spacing=0.5,
                   pairs_per_submission=256, scale=1.0, verify=True, drop_in=False):
    """BASELINE.json configs[3] / configs[4]: a synthetic aerial survey (PlanarSurvey: grid of nadir cameras over a
    textured plane, 8192 features per image, directed pairs = 10 nearest cameras minus self) through the batched LinkStage
    runner (host/link_batch.hpp = src/pipeline/link_stage.cpp:75-112 for a pair list): subsample, descriptor upload, K1
    matching in large submissions, ratio test + compaction on the device (K5), the reference's std::sort [, rays on the
    device (K6), RANSAC with GPU scoring, decomposition, inlier assembly]. Pairs are sharded over the ranks by the
    Hilbert-curve partition of the camera positions (host/partition.cpp); no data-path collective. The ONE cross-rank
    step is inside the timed region: the host gather of every pair's match list to rank 0, restored to the serial pair
    order (link_stage.cpp:119-131), through shared memory (opencalibration_b200/sharding.py: MatchGather).
    One step = the whole survey once, host vectors in, gathered lists on rank 0 out. value = pairs of all ranks /
    max-over-ranks wall time of the step (host code is part of this path, so it is wall time, bracketed by device
    synchronisation + barrier). Returns the block for the JSON line on rank 0, None elsewhere."""
    from concurrent.futures import ThreadPoolExecutor
    from opencalibration_b200 import capi, host, sharding, synthetic

    rows, cols = (25, 40) if workload == "c4" else (50, 100)
    if scale != 1.0:
        rows, cols = max(2, int(rows * scale)), max(2, int(cols * scale))
    survey = synthetic.PlanarSurvey(rows, cols, 8192, seed=7)
    pairs = survey.pairs
    if workload == "c5":
        pairs = pairs[:40000]
    shards = sharding.partition(survey.positions, pairs, world)
    shard = shards[rank]
    resident = shard.resident_images.tolist()
    local_index = {g: i for i, g in enumerate(resident)}
    cores = os.cpu_count() or 1
    with ThreadPoolExecutor(max(1, min(8, cores // world))) as ex:
        images = list(ex.map(survey.image, resident))
    sets = [host.FeatureSet(d, xy, st) for d, xy, st in images]
    cams = [survey.camera8()] * len(sets)
    local_pairs = [(local_index[a], local_index[b]) for a, b in shard.pairs]
    threads = max(1, cores // world)  # the submission threads sleep while the GPU works (blocking waits)
    most = max(len(sh.pairs) for sh in shards)
    gather = sharding.MatchGather(capacity_records=max(1, most) * 8192, max_pairs=most, dist=dist)
    packed = gather.buffers()  # this rank's region of the segment: the runner's tail workers write into it directly

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()

    phase = np.zeros(3)  # this rank's wall time in link_pairs / pack / gather, summed over the timed steps

    def step():
        t_a = time.perf_counter()
        res = host.link_pairs(sets, cams, local_pairs, threads=threads, pairs_per_submission=pairs_per_submission,
                              run_ransac=with_ransac, spacing=spacing, packed=packed)
        t_b = time.perf_counter()
        out = gather.gather([sh.pair_ids for sh in shards], len(pairs))
        phase[:] += (t_b - t_a, 0.0, time.perf_counter() - t_b)
        return res, out

    # warm-up: whole untimed steps (they size the page-locked result buffers, the per-thread staging areas and the
    # device memory pool the descriptor sets live in)
    n_warm = max(1, min(warmup, 2))
    for _ in range(n_warm):
        step()[0].close()
    steps = max(1, min(steps, 3))
    launches0 = capi.kernel_launches()
    sampler = ClockSampler(local_rank)
    sampler.start()
    phase[:] = 0
    barrier()
    t0 = time.perf_counter()
    res = gathered = None
    for _ in range(steps):
        if res is not None:
            res.close()
        res, gathered = step()
    barrier()
    secs = (time.perf_counter() - t0) / steps
    stats = res.stats
    clocks = sampler.result()
    launches = capi.kernel_launches() - launches0
    kept = sum(1 for p in range(res.n_pairs) if res.sizes(p)[0] > 0)
    res.close()
    t = torch.tensor([secs], dtype=torch.float64, device="cuda")
    agg = torch.tensor([len(local_pairs), stats["comparisons"], stats["matches"], stats["ransac_inliers"], kept,
                        len(resident)], dtype=torch.float64, device="cuda")
    per_rank = torch.zeros(world, 5, dtype=torch.float64, device="cuda")
    per_rank[rank] = torch.tensor([len(local_pairs), len(resident), *(phase / steps)], dtype=torch.float64)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
        dist.all_reduce(per_rank, op=dist.ReduceOp.SUM)
    block = None
    if rank == 0:
        secs = float(t.item())
        n_pairs, cmps, matches, inliers, kept_all, resident_all = [float(x) for x in agg.tolist()]
        assert gathered is not None and len(gathered) == len(pairs) and gathered.total() == int(matches)
        block = {
            "workload": f"configs[{3 if workload == 'c4' else 4}]: {rows}x{cols} image survey, {int(n_pairs)} directed "
                        f"pairs x 8192 features, batched LinkStage runner (match + ratio test + sort"
                        f"{' + rays + RANSAC + decompose' if with_ransac else ''}), Hilbert partition over {world} "
                        f"rank(s), match lists gathered to rank 0 inside the step",
            "pairs_per_s": n_pairs / secs, "ms_per_step": secs * 1e3, "steps": steps, "warmup": n_warm,
            "scaling": "strong", "images": survey.n_images, "pairs": int(n_pairs), "subsample_spacing_px": spacing,
            "comparisons_per_step": cmps, "Gcmp_per_s": cmps / secs / 1e9, "matches": int(matches),
            "ransac_inliers": int(inliers), "pairs_with_matches": int(kept_all),
            "resident_images_all_ranks": int(resident_all), "host_threads_per_rank": threads, "host_cores": cores,
            "rank0_breakdown_s": {k: v for k, v in stats.items() if k.startswith("seconds")},
            "per_rank": [{"pairs": int(r[0]), "resident_images": int(r[1]), "link_pairs_s": r[2],
                          "gather_wait_s": r[4]} for r in per_rank.cpu().tolist()],
            "gather": {"what": "every pair's match list (12-byte records), all ranks -> rank 0, serial pair order",
                       "transport": "POSIX shared memory on the box (records and index written by the tail workers while "
                                    "the rank is still matching) + torch.distributed barrier",
                       "records": gathered.total(), "bytes": gathered.total() * 12},
            "h2d_bytes_per_step": int(resident_all) * 8192 * (64 + (16 if with_ransac else 0)),
            "d2h_bytes_per_step": int(matches) * 12, "gpu_launches": int(launches), "clocks": clocks,
        }
    # ---- after the clock: what was gathered against (a) rank 0 running the WHOLE survey alone in this same job and
    # (b) the reference's own object code on a sample of pairs
    if verify and world > 1:
        if rank == 0:
            missing = [g for g in range(survey.n_images) if g not in local_index]
            with ThreadPoolExecutor(max(1, min(16, cores))) as ex:
                extra = list(ex.map(survey.image, missing))
            every = {g: sets[local_index[g]] for g in resident}
            every.update({g: host.FeatureSet(d, xy, st) for g, (d, xy, st) in zip(missing, extra)})
            all_sets = [every[g] for g in range(survey.n_images)]
            all_cams = [survey.camera8()] * len(all_sets)
            alone_threads = max(1, cores - 1)
            host.link_pairs(all_sets, all_cams, pairs, threads=alone_threads, pairs_per_submission=pairs_per_submission,
                            run_ransac=with_ransac, spacing=spacing).close()
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            alone = host.link_pairs(all_sets, all_cams, pairs, threads=alone_threads,
                                    pairs_per_submission=pairs_per_submission, run_ransac=with_ransac, spacing=spacing)
            alone_secs = time.perf_counter() - t1
            counts1, rec1 = alone.pack_matches(threads=alone_threads)
            alone.close()
            same = bool(np.array_equal(counts1.astype(np.int64), gathered.counts))
            o = 0
            for p in range(len(pairs)):
                if not same:
                    break
                n = int(counts1[p])
                same = np.array_equal(rec1[o:o + n], gathered.records(p))
                o += n
            assert same, "the gathered match lists differ from the single-GPU result"
            block["gathered_equals_single_gpu_result"] = True
            block["single_gpu_in_this_job"] = {"pairs_per_s": len(pairs) / alone_secs, "host_threads": alone_threads,
                                               "note": "rank 0 alone on the whole survey while the other ranks wait"}
            block["strong_scaling_efficiency_vs_single_gpu_in_this_job"] = \
                block["pairs_per_s"] / (world * len(pairs) / alone_secs)
        if dist:
            dist.barrier()
    if rank == 0 and verify:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oc_oracle as O
        try:
            ref, kind = O.Reference(popcnt=True), "reference object code"
        except (FileNotFoundError, OSError):
            ref, kind = O.Oracle(), "oracle port"
        if not with_ransac:  # with RANSAC only the pairs that keep a relation keep their matches (link_stage.cpp:103-107)
            sample = [p for p in shard.pair_ids.tolist()][:: max(1, len(shard.pair_ids) // 16)][:16]

            def want(p):
                a, b = pairs[p]
                (da, xa, sa), (db, xb, sb) = images[local_index[a]], images[local_index[b]]
                return ref.match_features_subset(da, db, ref.subsample(xa, sa, spacing), ref.subsample(xb, sb, spacing))

            with ThreadPoolExecutor(max(1, min(16, cores))) as ex:
                wanted = list(ex.map(want, sample))
            ok = all(all(np.array_equal(x, y) for x, y in zip(gathered.pair(p), w)) for p, w in zip(sample, wanted))
            assert ok, "a gathered match list differs from the reference's"
            block["parity_sample"] = {"pairs": len(sample), "against": kind, "equal": True}
    if rank == 0 and drop_in and local_pairs:
        # the same pairs through the DROP-IN entry point, the way the unmodified reference calls it: one
        # match_features_subset(std::vector<feature_2d>...) closure per pair on OpenMP workers (pipeline.cpp:42-49),
        # every call packing, uploading and matching its own two images (rank 0, a bounded sample of its pairs)
        sample = local_pairs[:min(len(local_pairs), 768)]
        sq, sc = [sets[a] for a, _ in sample], [sets[b] for _, b in sample]
        host.run_parallel_handles(sq[:64], sc[:64], threads=threads)  # sizes the workers' staging areas
        d_secs, d_matches = host.run_parallel_handles(sq, sc, threads=threads)
        block["drop_in_per_pair_calls"] = {
            "pairs_per_s": len(sample) / d_secs, "pairs": len(sample), "host_threads": threads, "matches": d_matches,
            "api": "match_features_subset(std::vector<feature_2d>...) per pair, all features of both images, one "
                   "closure per OpenMP worker like run_parallel (src/pipeline/pipeline.cpp:42-49)"}
    del gathered, packed
    gather.close()
    for fs in sets:
        fs.close()
    return block


def run_survey(args):
    """--workload c4 / c5: the survey of measure_survey as the bench line of its own."""
    import torch
    from opencalibration_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    capi.init(local_rank)
    if os.environ.get("OCB_K1_ITEMS"):  # experiments: work items per SM of the K1 planner
        capi.set_option("k1_items_per_sm", int(os.environ["OCB_K1_ITEMS"]))
    b = measure_survey(torch, dist, rank, local_rank, world, args.workload, args.steps if args.steps < 50 else 1,
                       args.warmup, args.with_ransac, args.spacing, args.pairs_per_submission, args.survey_scale,
                       verify=not args.no_verify, drop_in=not args.no_secondary)
    if rank == 0:
        line = {"metric": "image_pairs_matched_per_s", "value": b["pairs_per_s"], "unit": "pairs/s", "n_gpus": world,
                "steps": b["steps"], "warmup": b["warmup"], "ms_per_step": b["ms_per_step"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                "config": {k: v for k, v in b.items() if k not in ("clocks", "gpu_launches")},
                "clocks": b["clocks"], "gpu_launches": b["gpu_launches"],
                "e2e": {"value": b["pairs_per_s"], "unit": "pairs/s", "h2d_bytes_per_step": b["h2d_bytes_per_step"],
                        "d2h_bytes_per_step": b["d2h_bytes_per_step"],
                        "note": "the step is end to end by construction: features start in host vectors, the gathered "
                                "match lists end in rank 0's host memory"}}
        emit(line)
    if dist:
        dist.destroy_process_group()


_REAL_STDOUT = None


def quiet_stdout():
    """Libraries (NCCL's version banner, for one) write to fd 1; the contract is ONE JSON line on stdout. Everything
    else goes to stderr: fd 1 is pointed at fd 2 for the duration of the run and emit() writes to the saved fd."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.buffer.write(data)
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c4", "c5"],
                    help="c2 = headline (default); c4 / c5 = survey through the batched LinkStage runner")
    ap.add_argument("--survey-scale", type=float, default=1.0, help="shrink the c4/c5 grid (for quick runs)")
    ap.add_argument("--spacing", type=float, default=0.5,
                    help="c4/c5: subsample spacing in px (0.5 keeps all 8192 features; LinkStage uses 40)")
    ap.add_argument("--pairs-per-submission", type=int, default=256)
    ap.add_argument("--with-ransac", action="store_true",
                    help="c4/c5: continue every pair through rays + RANSAC + decomposition + inlier assembly "
                         "(default: the metric's unit, pairs MATCHED: match lists after ratio test and sort)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-survey", action="store_true", help="c2: skip the sharded configs[3] survey block")
    ap.add_argument("--no-verify", action="store_true", help="skip the after-the-clock checks of the survey block")
    ap.add_argument("--callers", type=int, default=12,
                    help="concurrent host callers of the e2e measurement (capped by the host cores of this rank)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload != "c2":
        run_survey(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
