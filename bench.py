"""Benchmark of the GNS rollout hot path (BASELINE.json metric: particle-steps/s).

    python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl reference]

A "step" is one rollout step of the whole particle cloud: neighbor search -> features ->
GNS forward (10 message-passing steps) -> integrate.  Prints ONE JSON line (rank 0).
See DESIGN.md "Measurement" for what each key means and how it is computed.
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "particle-steps/s GNS rollout (LDC-3D)"
UNIT = "particle-steps/s"
MP_STEPS = 10
EDGE_BYTES = 1032  # read e (512) + write e' (512) + 2 int32 indices, per edge per MP step
NODE_BYTES = 1024  # read the node row once for the gather side (512) + write the aggregate (512)


def metric_name(workload):
    """BASELINE.json's metric; other workloads (parity-test cases run by hand) are named in it."""
    return METRIC if workload.startswith("ldc3d") else METRIC.replace("LDC-3D", workload)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML polled from a thread every
    ~2 ms (the default timed region lasts tens of milliseconds); nvidia-smi -lms as the fallback."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASON_BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
                   0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, device_index, period_s=0.002):
        self.device_index = device_index
        self.period_s = period_s
        self.proc = None
        self.thread = None
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = False
        self._nvml = None
        try:
            import pynvml
            import torch

            pynvml.nvmlInit()
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(device_index).uuid)
                self._handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self._handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM))
            self._nvml = pynvml
        except Exception:
            self._nvml = None

    def _poll(self):
        nv = self._nvml
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        # the first poll waits until the timed call has handed its launches to the driver (a fraction of a
        # millisecond): NVML queries hold a driver lock, and a query right at the start delayed the first graph
        # launch by up to a millisecond of device idle time inside the timed region
        time.sleep(float(os.environ.get("BENCH_SAMPLER_DELAY_MS", "3")) * 1e-3)
        while not self._stop:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._handle, nv.NVML_CLOCK_SM)))
                bits = int(reasons_fn(self._handle))
                for bit, name in self.REASON_BITS.items():
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period_s)

    def start(self):
        if self._nvml is not None:
            import threading

            self._stop = False
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.device_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self._stop = True
            self.thread.join(timeout=2)
            sm = self.samples
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(self.reasons), "samples": len(sm), "source": "nvml, %g ms period" % (self.period_s * 1e3)}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


def ncu_traffic(workload, kernel="edge_mp_tc2_kernel"):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    of this workload (profiles/ncu_traffic.json), or None when no capture exists for it."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            table = json.load(f)
        ent = table.get(workload, {}).get(kernel)
        return None if ent is None else ent["dram_bytes_per_launch"]
    except (OSError, ValueError, KeyError):
        return None


def build_workload(name, n_future, seed, dtype_name):
    from lagrangebench_b200 import synthetic

    npd = np.float64 if dtype_name == "float64" else np.float32
    # quiet dynamics: a random-init network must not blow the cloud up over a long horizon (synthetic.py)
    return synthetic.make_case(name, 6, n_future, seed, npd, quiet=True)


def node_in_of(case_spec):
    d = case_spec["metadata"]["dim"]
    n = 5 * d
    if not any(case_spec["metadata"]["periodic_boundary_conditions"]):
        n += 2 * d
    if case_spec["force"] is not None:
        n += d
    return n


# --------------------------------------------------------------------------------------
# oracle legs (cpu_baseline / --impl reference).  The ONLY place bench.py executes oracle/.
def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU legs should use every host core."""
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=os.cpu_count())
    except Exception:  # noqa: BLE001
        pass


def oracle_steps(spec, n_steps, warmup, seed, dtype_name):
    """Time the oracle port of the path on all host cores: per step the NumPy neighbor update + features
    (float64 by default, like the reference), the float32 GNS forward as torch-CPU ops following the
    reference's data flow literally (oracle/gns_torch.py), the integrate + kinematic override.
    Returns (seconds for n_steps, N, E)."""
    use_all_host_threads()
    import torch

    from oracle import case as ocase
    from oracle import gns as ogns
    from oracle import gns_torch
    from oracle import rollout as orollout

    torch.set_num_threads(os.cpu_count() or 1)
    npd = np.float64 if dtype_name == "float64" else np.float32
    force = spec["force"]
    ofn = None
    if force is not None:
        lo, hi = np.array(force.lo), np.array(force.hi)
        ofn = lambda r: hi if r[force.axis] > force.threshold else lo  # noqa: E731
    case = ocase.case_builder(spec["box"], spec["metadata"], 6, cfg_neighbors={"multiplier": spec["multiplier"]},
                              external_force_fn=ofn, dtype=npd, noise_std=0.0)
    d = spec["metadata"]["dim"]
    params = gns_torch.pack(ogns.init_params(node_in_of(spec), d + 1, d, num_mp_steps=MP_STEPS, seed=seed,
                                             perturb=False))
    current = spec["positions"][:, :6].astype(npd)
    ptype = spec["particle_type"]
    _, nbrs = case.allocate_eval((current, ptype))

    def apply(p, state, sample):
        feats, pt = sample
        return gns_torch.forward(p, feats, pt, MP_STEPS), state

    elapsed = 0.0
    for step in range(warmup + n_steps):
        t0 = time.perf_counter()
        feats, nbrs = case.preprocess_eval((current, ptype), nbrs)
        if nbrs.did_buffer_overflow:
            _, nbrs = case.allocate_eval((current, ptype))
            feats, nbrs = case.preprocess_eval((current, ptype), nbrs)
        tgt = current[:, -1]
        current, _ = orollout.forward_eval(apply, case.integrate, params, {}, (feats, ptype), current, tgt)
        if step >= warmup:
            elapsed += time.perf_counter() - t0
    return elapsed, current.shape[0], nbrs.n_edges


def jax_reference_steps(spec, n_steps, warmup, seed, dtype_name):
    """The reference itself on JAX-CPU -- ``lagrangebench.case_builder`` + ``models.GNS`` under haiku +
    ``evaluate.rollout._forward_eval``, its own per-step loop (``rollout.py:125-169``) -- when the JAX stack
    is importable (site-packages or ``baseline/_ref``).  Returns (seconds, N, E) or raises ImportError."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(ref_dir) and ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    os.environ.setdefault("JAX_PLATFORMS", "cpu")
    import haiku as hk  # noqa: F401
    import jax
    import jax.numpy as jnp
    import jax_sph  # noqa: F401
    import jraph  # noqa: F401
    import lagrangebench
    from lagrangebench.evaluate.rollout import _forward_eval

    if dtype_name == "float64":
        jax.config.update("jax_enable_x64", True)
    force = spec["force"]
    ffn = None
    if force is not None:
        lo, hi = jnp.array(force.lo), jnp.array(force.hi)
        ffn = lambda r: jnp.where(r[force.axis] > force.threshold, hi, lo)  # noqa: E731
    case = lagrangebench.case_builder(box=np.asarray(spec["box"]), metadata=spec["metadata"], input_seq_length=6,
                                      cfg_neighbors={"backend": "jaxmd_vmap", "multiplier": spec["multiplier"]},
                                      noise_std=0.0, external_force_fn=ffn, dtype=dtype_name)
    d = spec["metadata"]["dim"]

    def model_fn(x):
        return lagrangebench.models.GNS(particle_dimension=d, latent_size=128, blocks_per_step=2,
                                        num_mp_steps=MP_STEPS, particle_type_embedding_size=16)(x)

    model = hk.without_apply_rng(hk.transform_with_state(model_fn))
    current = jnp.asarray(spec["positions"][:, :6])
    ptype = jnp.asarray(spec["particle_type"])
    feats, nbrs = case.allocate_eval((current, ptype))
    params, state = model.init(jax.random.PRNGKey(seed), (feats, ptype))
    elapsed = 0.0
    for step in range(warmup + n_steps + 1):  # one extra leading step: jit compilation is not step time
        t0 = time.perf_counter()
        feats, nbrs = case.preprocess_eval((current, ptype), nbrs)
        if bool(nbrs.did_buffer_overflow):
            feats, nbrs = case.allocate_eval((current, ptype))
        current, state = _forward_eval(params, state, (feats, ptype), current, current[:, -1], model.apply,
                                       case.integrate)
        jax.block_until_ready(current)
        if step >= warmup + 1:
            elapsed += time.perf_counter() - t0
    n_edges = int((np.asarray(nbrs.idx[0]) < current.shape[0]).sum())
    return elapsed, int(current.shape[0]), n_edges


def sub_lattice_dims(dims, n_target):
    """Shrink a lattice to about n_target sites keeping its aspect ratio (min 8 per side)."""
    dims = np.array(dims, dtype=np.float64)
    scale = (n_target / dims.prod()) ** (1.0 / len(dims))
    return tuple(int(max(8, round(v * scale))) for v in dims)


def oracle_sample_spec(workload, n_target, seed, dtype_name):
    from lagrangebench_b200 import synthetic

    npd = np.float64 if dtype_name == "float64" else np.float32
    full = synthetic.CASES[workload]["dims"]
    dims = full if n_target >= int(np.prod(full)) else sub_lattice_dims(full, n_target)
    if workload == "dam2d":
        dims = full  # masked lattice: always the full tank
    return synthetic.make_case(workload, 6, 0, seed, npd, dims=dims, quiet=True), dims


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  The JAX stack is
    probed at run time (it is not installable in this image: no network, not in /opt/wheelhouse); when it is
    there the reference itself is timed (kind "reference"), otherwise the oracle port (kind "port").  The
    workload is the full configuration whenever K + W steps of it fit the time budget."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from lagrangebench_b200 import synthetic

    total_steps = args.steps + args.warmup
    budget_s = float(os.environ.get("BENCH_REFERENCE_BUDGET_S", "300"))
    kind, stepper, note = "port", oracle_steps, None
    try:
        probe, _ = oracle_sample_spec(args.workload, 2000, args.seed, args.dtype)
        jax_reference_steps(probe, 1, 0, args.seed, args.dtype)
        kind, stepper = "reference", jax_reference_steps
        note = "tumaer/lagrangebench on JAX-CPU (its own case_builder, GNS and _forward_eval), all host cores"
    except Exception as exc:  # noqa: BLE001
        note = ("oracle port of the reference path on host cores: NumPy neighbor search / features in %s, float32 "
                "torch-CPU GNS forward with the reference's data flow (the reference's JAX stack is not importable "
                "here: %s: %s)" % (args.dtype, type(exc).__name__, str(exc)[:80]))
    probe, _ = oracle_sample_spec(args.workload, 4000, args.seed, args.dtype)
    stepper(probe, 1, 1, args.seed, args.dtype)  # warm the thread pools
    t_probe, n_probe, _ = stepper(probe, 1, 0, args.seed, args.dtype)
    rate = n_probe / max(t_probe, 1e-9)
    n_full = int(np.prod(synthetic.CASES[args.workload]["dims"]))
    n_target = int(min(n_full, max(2000, rate * budget_s / total_steps)))
    if n_target >= 0.8 * n_full:
        n_target = n_full
    spec, dims = oracle_sample_spec(args.workload, n_target, args.seed, args.dtype)
    elapsed, n, e = stepper(spec, args.steps, args.warmup, args.seed, args.dtype)
    value = n * args.steps / elapsed
    cores = os.cpu_count()
    same = n == n_full
    sample = (f"{args.workload} full configuration" if same else f"{args.workload} sub-lattice {dims}") + \
        f" = {n} particles, {e} edges, {args.steps} full rollout steps"
    line = {
        "impl": "reference", "metric": metric_name(args.workload), "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "particles": n, "edges": e, "mp_steps": MP_STEPS, "latent": 128,
                   "positions": args.dtype, "same_config": same},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": note,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from lagrangebench_b200 import GNS, RolloutEngine, _cabi, case_builder
    from lagrangebench_b200 import models as lbmodels

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    _cabi.require_cuda()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _cabi.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    K, W = args.steps, args.warmup
    # replicas: every rank rolls out its own trajectory of the same shape (clouds of
    # this size do not shard; the reference's batch vmap maps to one trajectory per GPU)
    spec = build_workload(args.workload, K + W, args.seed + rank, args.dtype)
    d = spec["metadata"]["dim"]
    n = spec["positions"].shape[0]
    case = case_builder(spec["box"], spec["metadata"], 6, cfg_neighbors={"multiplier": spec["multiplier"]},
                        external_force_fn=spec["force"], dtype=args.dtype, noise_std=0.0)
    params = lbmodels.init_params(node_in_of(spec), d, 128, MP_STEPS, 16, seed=args.seed)
    model = GNS(d, 128, 2, MP_STEPS, 16)
    engine = RolloutEngine(case, model, params, steps_per_sync=max(K, W, 1))
    tdt = torch.float64 if args.dtype == "float64" else torch.float32
    window = torch.as_tensor(spec["positions"][:, :6]).to(dev, tdt).contiguous()
    targets_all = torch.as_tensor(spec["positions"][:, 6:6 + K + W]).permute(1, 0, 2).to(dev, tdt).contiguous()
    ptype = torch.as_tensor(spec["particle_type"]).to(dev)

    # warm-up: W steps (sizes the neighbor capacities and scratch), then an untimed rehearsal of the timed
    # call on the very same buffers, so that the library has its step graph for exactly these
    # arguments (one-time capture + instantiation is set-up, not step time); the state is restored
    _, nbrs = engine.run(window, ptype, targets_all[:W], W)
    window_start = window.clone()
    preds_buf = torch.empty((K, n, d), dtype=tdt, device=dev)
    _, nbrs = engine.run(window, ptype, targets_all[W:W + K], K, nbrs, out=preds_buf)
    window.copy_(window_start)
    barrier()
    launches0 = lib.lb200_launch_count()
    realloc0 = engine.n_reallocations
    # NVML queries share a driver lock with the graph launches of the timed region: a 2 ms poll added up to 4 % of
    # run-to-run spread (and ~1.5 % on average) to ms_per_step, 8 ms leaves 4-5 samples in the ~37 ms region
    sampler = ClockSampler(local_rank, period_s=float(os.environ.get("BENCH_SAMPLER_MS", "8")) * 1e-3)
    if rank == 0 and not os.environ.get("BENCH_NO_SAMPLER"):
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    preds, nbrs = engine.run(window, ptype, targets_all[W:W + K], K, nbrs, out=preds_buf)  # the API path: graph replay
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    launches = lib.lb200_launch_count() - launches0
    n_edges = nbrs.n_edges
    assert torch.isfinite(preds).all(), "rollout diverged"
    # kernel leg (roofline): the same K steps once more with a CUDA event pair around every message /
    # node kernel launch on its launch stream (event records switch the graph replay off, so this
    # pass runs the launches eagerly; its own wall time is reported in config)
    import ctypes as C

    window_p = torch.as_tensor(spec["positions"][:, :6]).to(dev, tdt).contiguous()
    _, nb_p = engine.run(window_p, ptype, targets_all[:W], W)
    barrier()
    lib.lb200_profile(1)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    engine.run(window_p, ptype, targets_all[W:W + K], K, nb_p)
    g1.record()
    barrier()
    ms_prof = g0.elapsed_time(g1)
    kms = (C.c_double * 2)()
    kl = (C.c_int64 * 2)()
    _cabi.check(lib.lb200_profile_read(kms, kl))
    lib.lb200_profile(0)

    # end-to-end leg: the same steps through the public per-step API with HOST buffers -- every step is one
    # RolloutEngine.run(...) call that uploads that step's kinematic-target frame from pinned memory, advances
    # one step and reads the predicted positions back into pinned memory: one host synchronisation per step
    # (as the reference loop has), which covers upload, step, read-back and status
    esz = 8 if args.dtype == "float64" else 4
    h_targets = torch.as_tensor(spec["positions"][:, 6:6 + K + W]).permute(1, 0, 2).to(tdt).contiguous().pin_memory()
    h_out = torch.empty((1, n, d), dtype=tdt).pin_memory()
    d_out = torch.empty((1, n, d), dtype=tdt, device=dev)
    window_e = torch.as_tensor(spec["positions"][:, :6]).to(dev, tdt).contiguous()
    engine_e = RolloutEngine(case, model, params, steps_per_sync=1)
    nb_e = None
    for t in range(min(W, 3) + 4):  # untimed: also lets the per-step graph be captured
        tt = min(t, K + W - 1)
        _, nb_e = engine_e.run(window_e, ptype, h_targets[tt:tt + 1], 1, nb_e, out=d_out, host_out=h_out)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(K):
        _, nb_e = engine_e.run(window_e, ptype, h_targets[W + t:W + t + 1], 1, nb_e, out=d_out, host_out=h_out)
    torch.cuda.current_stream().synchronize()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    assert torch.isfinite(h_out).all()

    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
        cnt = torch.tensor([float(launches)], device=dev, dtype=torch.float64)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        launches = int(cnt.item())

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        edge_ms_avg = kms[0] / max(kl[0], 1)
        alg_bytes = EDGE_BYTES * n_edges + NODE_BYTES * n
        achieved = alg_bytes / (edge_ms_avg * 1e-3) / 1e9 if edge_ms_avg > 0 else 0.0
        flops_launch = n_edges * 65536.0  # restructured count: W1 node terms hoisted
        value = world * n * K / (ms * 1e-3)
        line = {
            "metric": metric_name(args.workload), "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "particles": n, "edges": n_edges, "mp_steps": MP_STEPS,
                       "latent": 128, "positions": args.dtype, "multi_gpu": "replicas" if world > 1 else "single",
                       "l2": "inputs larger than L2 (edge latents %.0f MB)" % (n_edges * 512 / 1e6)
                       if n_edges * 512 > 126e6 else "working set below L2 (edge latents %.0f MB): latency-bound"
                       % (n_edges * 512 / 1e6),
                       "reallocations_in_timed_region": engine.n_reallocations - realloc0,
                       "step_loop": "device-resident, CUDA-graph replay (lb200_rollout_steps)",
                       "dynamics": "quiet synthetic statistics (vel_std 2e-4 dx, acc_std 1e-7 dx, noise_std 0): a "
                                   "random-init model keeps the cloud on its lattice for any K",
                       "ms_per_step_kernel_leg_eager_with_events": ms_prof / K},
            "roofline": {"bound": "hbm", "kernel": "edge_mp_tc2_kernel (tcgen05, weights in TMEM: fused gather + edge MLP + "
                         "LayerNorm + residual + segmented sum)", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": ncu_traffic(args.workload),
                         "traffic_source": "profiles/ncu_traffic.json (dram bytes per launch of the committed ncu --set full "
                                           "capture of this kernel and workload; not measured in this run)",
                         "algorithmic_bytes": alg_bytes, "peak_source": peak_kind,
                         "avg_launch_ms": edge_ms_avg, "launches": int(kl[0]),
                         "share_of_step": kms[0] / ms_prof, "node_kernel_share_of_step": kms[1] / ms_prof,
                         "fp32_tflops": flops_launch / (edge_ms_avg * 1e-3) / 1e12 if edge_ms_avg > 0 else 0.0},
            "e2e": {"value": world * n * K / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": n * d * esz,
                    "d2h_bytes_per_step": n * d * esz},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            sample_spec, dims = oracle_sample_spec(args.workload, args.cpu_sample, args.seed, args.dtype)
            t_cpu, n_cpu, e_cpu = oracle_steps(sample_spec, 1, 0, args.seed, args.dtype)
            line["cpu_baseline"] = {
                "value": n_cpu / t_cpu, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                "sample": f"1 full rollout step of the oracle port (NumPy neighbor search + features, float32 torch-CPU "
                          f"GNS forward on all cores) on {args.workload} {dims} = {n_cpu} particles / {e_cpu} edges "
                          f"({t_cpu:.1f} s)"}
    # strong scaling of ONE sharded 1 M-particle cloud over the same ranks (every N, N = 1 included)
    ss = None
    if not args.no_strong_scaling:
        del engine, engine_e, preds, preds_buf, targets_all, window, window_p, window_e
        torch.cuda.empty_cache()
        ss = strong_scaling_leg(args, world, rank, dev, max(K, 20), max(W, 3))
    if rank == 0:
        line["strong_scaling"] = ss
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def strong_scaling_leg(args, world, rank, dev, K, W):
    """ONE periodic cloud (BASELINE.json configs[4]: RPF-3D, 1 M particles) sharded over all ranks: slab
    decomposition, peer-memory halo, device-resident step loop (domain.py).  Run at every N, N = 1 included,
    so that the driver's per-N lines carry the strong-scaling series.  Returns the dict of the
    ``strong_scaling`` key (rank 0) or None.

    Correctness rides along: the first ``check_steps`` steps are compared with the single-GPU engine
    (rank 0 rolls the whole cloud out alone) -- max |dpos| in units of dx -- plus position checksums."""
    import ctypes as C

    import torch
    import torch.distributed as dist

    from lagrangebench_b200 import GNS, RolloutEngine, _cabi, case_builder
    from lagrangebench_b200 import models as lbmodels
    from lagrangebench_b200.domain import DistributedRollout

    lib = _cabi.load()
    name = args.scaling_workload

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    spec = build_workload(name, 0, args.seed, "float32")
    d = spec["metadata"]["dim"]
    dx = spec["metadata"]["dx"]
    n_total = spec["positions"].shape[0]
    params = lbmodels.init_params(node_in_of(spec), d, 128, MP_STEPS, 16, seed=args.seed)
    tdt = torch.float64 if args.dtype == "float64" else torch.float32
    check_steps = args.check_steps
    dr = DistributedRollout(spec["box"], spec["metadata"], params, MP_STEPS, force=spec["force"], dtype=tdt,
                            multiplier=spec["multiplier"], noise_std=0.0, steps_per_sync=max(K, W, check_steps, 1))
    dr.scatter(spec["positions"], spec["particle_type"])
    # ---- correctness against the single-GPU trajectory
    dr.run(check_steps)
    pos_sharded = dr.gather_positions(n_total)
    check = None
    if rank == 0:
        case = case_builder(spec["box"], spec["metadata"], 6, cfg_neighbors={"multiplier": spec["multiplier"]},
                            external_force_fn=spec["force"], dtype=args.dtype, noise_std=0.0)
        eng = RolloutEngine(case, GNS(d, 128, 2, MP_STEPS, 16), params, steps_per_sync=check_steps)
        w1 = torch.as_tensor(spec["positions"][:, :6]).to(dev, tdt).contiguous()
        ref, _ = eng.run(w1, spec["particle_type"], None, check_steps)
        diff = case.displacement(pos_sharded, ref[-1]).abs().max().item()
        check = {"steps": check_steps, "against": "single-GPU engine on rank 0, same initial cloud",
                 "max_abs_dpos": diff, "max_dpos_over_dx": diff / dx,
                 "checksum_sharded": float(pos_sharded.double().sum().item()),
                 "checksum_single": float(ref[-1].double().sum().item())}
        del eng, w1, ref
    del pos_sharded
    torch.cuda.empty_cache()
    barrier()  # rank 0 rolled the whole cloud out alone: nobody spins on its signals meanwhile
    # ---- timing: W warm-up steps, then K steps = one call, one host synchronisation
    dr.run(W)
    barrier()
    launches0 = lib.lb200_launch_count()
    sel0, real0 = dr.n_selections, dr.n_reallocations
    sampler = ClockSampler(dev.index, period_s=0.025)  # NVML queries share a driver lock with kernel launches
    if rank == 0 and not os.environ.get("BENCH_NO_SAMPLER"):
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    dr.run(K)
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    launches = lib.lb200_launch_count() - launches0
    # ---- kernel leg: a few eager steps with an event pair around every message / node kernel launch
    kp = min(K, 5)
    lib.lb200_profile(1)
    dr.run(kp)
    torch.cuda.synchronize()
    kms, kl = (C.c_double * 2)(), (C.c_int64 * 2)()
    _cabi.check(lib.lb200_profile_read(kms, kl))
    lib.lb200_profile(0)
    n_own, n_edges = dr.window.shape[0], dr.edges_last
    per_rank = {"particles": n_own, "ghosts": dr.n_ghost_left + dr.n_ghost_right, "edges": n_edges,
                "halo_rows_sent": dr.halo_rows, "message_kernel_ms": kms[0] / max(kl[0], 1),
                "node_kernel_ms": kms[1] / max(kl[1], 1)}
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        cnt = torch.tensor([float(launches), float(n_own)], device=dev, dtype=torch.float64)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        launches, owned_total = int(cnt[0].item()), int(cnt[1].item())
        tables = [None] * world
        dist.all_gather_object(tables, per_rank)
    else:
        owned_total, tables = n_own, [per_rank]
    out = None
    if rank == 0:
        assert owned_total == n_total, "particles lost in migration"
        peaks, peak_kind = measured_peaks()
        edge_ms = per_rank["message_kernel_ms"]
        alg_bytes = EDGE_BYTES * n_edges + NODE_BYTES * n_own
        achieved = alg_bytes / (edge_ms * 1e-3) / 1e9 if edge_ms > 0 else 0.0
        out = {
            "workload": name, "particles": n_total, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "value": n_total * K / (ms * 1e-3), "unit": UNIT, "scaling": "strong",
            "positions": args.dtype, "mp_steps": MP_STEPS,
            "multi_gpu": "slab decomposition along axis %d; ghost positions and the sender projections of boundary "
                         "rows stored into the neighbours' peer-mapped heaps (node-kernel epilogue), one signal/wait "
                         "kernel per exchange, status bits OR-ed over the ranks on the device; steps replayed from a "
                         "CUDA graph, one host synchronisation per %d steps" % (dr.axis, K),
            "halo_bytes_per_step_rank0": dr.halo_bytes + dr.halo_rows * d * (8 if args.dtype == "float64" else 4),
            "halo_margin": dr.halo_margin, "ghost_selections_in_timed_region": dr.n_selections - sel0,
            "reallocations_in_timed_region": dr.n_reallocations - real0,
            "per_rank": tables, "gpu_launches": launches,
            "roofline_rank0": {"kernel": "edge_mp_tc2_kernel", "bound": "hbm", "achieved": achieved,
                               "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                               "peak_source": peak_kind, "avg_launch_ms": edge_ms, "algorithmic_bytes": alg_bytes},
            "check": check, "clocks": clocks,
        }
    dr.close()
    del dr
    torch.cuda.empty_cache()
    return out


def run_sharded(args):
    """``--sharded``: only the strong-scaling leg, printed as the line's main value (manual runs)."""
    import torch
    import torch.distributed as dist

    from lagrangebench_b200 import _cabi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    _cabi.require_cuda()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    args.scaling_workload = args.workload if args.workload.startswith("rpf3d") else args.scaling_workload
    ss = strong_scaling_leg(args, world, rank, dev, args.steps, args.warmup)
    if rank == 0:
        line = {"metric": METRIC.replace("LDC-3D", ss["workload"]), "value": ss["value"], "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ss["ms_per_step"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": ss["workload"], "particles": ss["particles"]}, "strong_scaling": ss,
                "gpu_launches": ss["gpu_launches"], "clocks": ss["clocks"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ldc3d_28k")
    ap.add_argument("--dtype", default="float64", choices=["float32", "float64"],
                    help="position / preprocessing dtype (reference default float64); the network is float32")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=28000, help="particles in the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong-scaling", action="store_true",
                    help="skip the sharded 1 M-particle leg (the strong_scaling key of the line)")
    ap.add_argument("--scaling-workload", default="rpf3d_1m")
    ap.add_argument("--check-steps", type=int, default=2,
                    help="steps of the sharded cloud compared with the single-GPU engine before timing")
    ap.add_argument("--sharded", action="store_true",
                    help="shard ONE periodic cloud over the GPUs (slab decomposition, strong scaling) "
                         "instead of one replica per GPU")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.sharded:
        run_sharded(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
