"""Rollout evaluation: host-side mirror of ``lagrangebench/evaluate/rollout.py``.

``infer`` / ``eval_rollout`` keep the reference signatures.  When the model is this
package's :class:`~lagrangebench_b200.models.GNS`, the step loop of
``_eval_batched_rollout`` (``rollout.py:125-169``) runs device-resident through
``lb200_rollout_steps`` -- one host synchronisation per chunk of steps instead of one
per step -- with the reference's overflow contract (re-allocate, retry the same step).
Any other ``model_apply`` callable goes through the generic per-step loop, whose stages
(neighbor update, features, integrate) are still the CUDA kernels.
"""

import ctypes as C
import os
import pickle
import time

import numpy as np
import torch

from . import _cabi
from .defaults import merged
from .case_setup import NeighborList, num_real_particles
from .models import GNS, gns_cfg
from .utils import broadcast_from_batch, get_kinematic_mask, load_haiku, set_seed


# ----------------------------------------------------------------------------- metrics
class MetricsComputer:
    """Position metrics of ``lagrangebench/evaluate/metrics.py:17-147`` that live on the
    rollout path: ``mse``, ``mae`` (+ horizon slices) and ``e_kin``.  ``sinkhorn`` is an
    O(N^2) optimal-transport solve outside the hot path and is not provided."""

    METRICS = ["mse", "mae", "e_kin"]

    def __init__(self, active_metrics, dist_fn, metadata, input_seq_length, stride=10, loss_ranges=None,
                 ot_backend=None):
        active_metrics = list(active_metrics or [])
        unsupported = [m for m in active_metrics if m not in self.METRICS]
        if unsupported:
            raise NotImplementedError(f"metrics {unsupported} are outside the rollout hot path")
        self._active_metrics = active_metrics
        self._dist_fn = dist_fn
        self._loss_ranges = loss_ranges or [1, 5, 10, 20, 50, 100]
        self._input_seq_length = input_seq_length
        self._stride = stride
        self._metadata = metadata

    def __call__(self, pred_rollout, target_rollout):
        pred = torch.as_tensor(pred_rollout)
        target = torch.as_tensor(target_rollout).to(device=pred.device, dtype=pred.dtype)
        metrics = {}
        for name in self._active_metrics:
            if name in ("mse", "mae"):
                d = self._dist_fn(pred, target)
                per_step = (d**2).mean(dim=(1, 2)) if name == "mse" else d.abs().mean(dim=(1, 2))
                metrics[name] = per_step
                for i in self._loss_ranges:
                    if i < per_step.shape[0]:
                        metrics[f"{name}{i}"] = per_step[:i]
            elif name == "e_kin":
                dt = self._metadata["dt"] * self._metadata["write_every"]
                dx, dim = self._metadata["dx"], self._metadata["dim"]

                def e_kin(roll):
                    vel = self._dist_fn(roll[1::self._stride], roll[0:-1:self._stride]) / dt
                    return (vel**2).sum(dim=(1, 2)) * dx**dim  # metrics.py:98-125,157-160

                ep, et = e_kin(pred), e_kin(target)
                metrics[name] = {"predicted": ep, "target": et, "mse": ((ep - et) ** 2).mean()}
        return metrics


def averaged_metrics(eval_metrics):
    """Rollout-averaged scalars, ``metrics.py:233-252``: ``mse`` and ``mae`` both feed ``val/loss``
    (the key the trainer checkpoints on), ``e_kin`` contributes its ``mse``; every key also gets a
    ``val/std<key>`` across rollouts."""
    per_key = {}
    for rollout in eval_metrics.values():
        for k, v in rollout.items():
            if k == "e_kin":
                v = v["mse"]
            if k in ("mse", "mae"):
                k = "loss"
            per_key.setdefault(k, []).append(float(torch.as_tensor(v, dtype=torch.float64).mean()))
    out = {f"val/{k}": float(np.mean(v)) for k, v in per_key.items()}
    out.update({f"val/std{k}": float(np.std(v)) for k, v in per_key.items()})
    return out


# ----------------------------------------------------------------------------- engine
class RolloutEngine:
    """Owns the device state of one device-resident rollout (weights blob, scratch, status)."""

    def __init__(self, case, model, params, steps_per_sync=32):
        if not isinstance(model, GNS):
            raise TypeError("RolloutEngine drives lagrangebench_b200.models.GNS")
        self.case, self.model = case, model
        self.h = case._lb200
        if self.h["force_mode"] == 2:
            raise NotImplementedError("a generic Python force callable needs the per-step loop")
        self.packed = model.packed_params(params)
        self._params = params
        self._siblings = []  # engines for the other trajectories of a batch (evaluate._eval_batched_rollout)
        self.steps_per_sync = int(steps_per_sync)
        self._cfg = None
        self._cfg_key = None
        self._scratch = None
        self._stream = None  # non-default stream: lets lb200_rollout_steps replay steps from a CUDA graph
        # persistent small buffers: stable device pointers let the library reuse its captured graph
        self._status = None
        self._status_host = None
        self._target_buf = None
        self._ptype = (None, None, 0)  # (source, int32 device copy, number of real particles)
        self.n_reallocations = 0
        self.n_launch_calls = 0

    def _configure(self, neighbors, n_valid):
        lib = _cabi.load()
        h = self.h
        n = neighbors._grid.n
        key = (id(neighbors._grid), n, n_valid, neighbors.max_occupancy, neighbors.cell_list_capacity)
        if self._cfg is not None and self._cfg_key == key:
            return
        cfg = _cabi.RolloutCfg()
        cfg.grid = neighbors._grid
        cfg.grid.n_valid = n_valid
        cfg.feat = h["feature_cfg"](n)
        cfg.gns = gns_cfg(self.packed, n, neighbors.max_occupancy, cfg.feat.node_stride, cfg.feat.node_stride,
                          self.model.edge_impl)
        cfg.integ = h["integrate_cfg"](n, h["isl"], 0)
        cfg.cell_capacity = neighbors.cell_list_capacity or 0
        cfg.e_cap = neighbors.max_occupancy
        nbytes = lib.lb200_rollout_scratch_bytes(C.byref(cfg))
        if self._scratch is None or self._scratch.numel() < nbytes:
            self._scratch = torch.empty(nbytes, dtype=torch.uint8, device=neighbors.reference_position.device)
        self._cfg, self._cfg_key, self._grid_ref = cfg, key, neighbors._grid

    def run(self, window, particle_type, targets, n_steps, neighbors=None, out=None, host_out=None):
        """Advance ``window`` (N, isl, d) in place by ``n_steps``.

        ``targets`` (n_steps, N, d) or None supplies the positions of kinematic particles; a HOST tensor
        (pinned for an asynchronous copy) is uploaded on the engine's stream in front of the steps.
        ``out``: optional preallocated ``(n_steps, N, d)`` tensor for the predictions (a caller that
        repeats a call with the very same buffers lets the library replay its captured step graph
        from the first step on).  ``host_out``: optional pinned host tensor of that shape; every chunk's
        predictions are copied into it behind the steps, so that the chunk's one synchronisation covers
        upload, steps, read-back and status.  Returns ``(predictions (n_steps, N, d), neighbors)``; the returned
        list carries the capacities and builds its ``idx`` array on first access (the step loop
        itself works on the receiver-major view of the graph)."""
        job = self.start(window, particle_type, targets, n_steps, neighbors, out, host_out)
        while not job.finished:
            job.enqueue()
            job.collect()
        return job.result()

    def start(self, window, particle_type, targets, n_steps, neighbors=None, out=None, host_out=None):
        """A rollout as a job: ``enqueue()`` puts the next chunk of steps on this engine's stream without
        waiting, ``collect()`` reads the chunk's status (the host synchronisation).  Several engines can
        have their chunks in flight at once (``run_batched``)."""
        return _RolloutJob(self, window, particle_type, targets, n_steps, neighbors, out, host_out)


class _RolloutJob:
    def __init__(self, engine, window, particle_type, targets, n_steps, neighbors, out, host_out=None):
        self.engine = eng = engine
        h = eng.h
        assert window.is_cuda and window.is_contiguous() and window.dtype == h["dtype"]
        n, isl, dim = window.shape
        dev = window.device
        if eng._ptype[0] is particle_type and eng._ptype[1].device == dev:
            ptype, n_valid = eng._ptype[1], eng._ptype[2]
        else:
            ptype = torch.as_tensor(particle_type).to(dev, torch.int32).contiguous()
            n_valid = num_real_particles(particle_type)
            eng._ptype = (particle_type, ptype, n_valid)
        if ptype.shape[0] != n:
            raise ValueError(f"particle_type has {ptype.shape[0]} rows, the position window {n}")
        nfn = h["neighbor_fn"]
        if neighbors is None or neighbors._grid is None or neighbors._grid.n != n:
            # also a trajectory with another particle count than the list was allocated for (rollout.py:383)
            neighbors = nfn.allocate(window[:, -1].contiguous(), num_particles=n_valid)
        eng._configure(neighbors, n_valid)
        if out is not None:
            assert out.shape == (n_steps, n, dim) and out.dtype == window.dtype and out.device == dev \
                and out.is_contiguous()
            preds = out
        else:
            preds = torch.empty((n_steps, n, dim), dtype=window.dtype, device=dev)
        if eng._status is None or eng._status.device != dev:
            eng._status = torch.zeros(4, dtype=torch.int32, device=dev)
            eng._status_host = torch.zeros(4, dtype=torch.int32).pin_memory()
        if eng._stream is None:
            eng._stream = torch.cuda.Stream(device=dev)
        self.upload = None
        if targets is not None:
            targets = torch.as_tensor(targets)
            assert tuple(targets.shape) == (n_steps, n, dim)
            if not targets.is_cuda:  # host frames: uploaded on the engine's stream, into a buffer the step graph keeps seeing
                buf = eng._target_buf
                if buf is None or buf.shape != targets.shape or buf.dtype != window.dtype or buf.device != dev:
                    buf = eng._target_buf = torch.empty(targets.shape, dtype=window.dtype, device=dev)
                self.upload = targets if targets.dtype == window.dtype else targets.to(window.dtype)
                targets = buf
            else:
                targets = targets.to(dev, window.dtype).contiguous()
        if host_out is not None:
            assert not host_out.is_cuda and tuple(host_out.shape) == (n_steps, n, dim) and host_out.dtype == window.dtype
        self.host_out = host_out
        self.window, self.ptype, self.n_valid, self.targets, self.preds = window, ptype, n_valid, targets, preds
        self.neighbors, self.n_steps, self.done, self.n_edges = neighbors, n_steps, 0, 0
        self.dev = dev

    @property
    def finished(self):
        return self.done >= self.n_steps

    def enqueue(self):
        eng, lib = self.engine, _cabi.load()
        chunk = min(eng.steps_per_sync, self.n_steps - self.done)
        eng._stream.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(eng._stream):
            if self.upload is not None:
                self.targets.copy_(self.upload, non_blocking=True)
                self.upload = None
            # base pointers + first frame: every chunk of a long rollout replays the same cached step graph;
            # status is reset on the device at the start of every lb200_rollout_steps call
            _cabi.check(lib.lb200_rollout_steps(
                C.byref(eng._cfg), chunk, _cabi.ptr(eng.packed.blob), _cabi.ptr(self.window), _cabi.ptr(self.ptype), None,
                _cabi.ptr(self.targets), _cabi.ptr(self.preds), self.done, None, _cabi.ptr(eng._status),
                _cabi.ptr(eng._scratch), eng._scratch.numel(), _cabi.stream()))
            if self.host_out is not None:  # read-back behind the steps (rows of a step that did not complete are rewritten)
                sl = slice(self.done, self.done + chunk)
                self.host_out[sl].copy_(self.preds[sl], non_blocking=True)
            eng._status_host.copy_(eng._status, non_blocking=True)
        eng.n_launch_calls += 1

    def collect(self):
        eng = self.engine
        eng._stream.synchronize()  # the one host sync per chunk: upload, steps, read-back, status
        torch.cuda.current_stream(self.dev).wait_stream(eng._stream)
        completed, overflow, n_edges, _ = eng._status_host.tolist()
        self.done += completed
        self.n_edges = n_edges
        if overflow & _cabi.ERR_NONFINITE:
            raise FloatingPointError(
                "rollout produced NaN / Inf accelerations: an activation left the range of the fp16 split of the "
                "tensor-core kernels (|x| > 65504).  Set model.edge_impl = 'simt' for the float32 CUDA-core kernels.")
        if overflow:  # rollout.py:135-151: re-allocate from the current state, retry the step
            eng.n_reallocations += 1
            self.neighbors = eng.h["neighbor_fn"].allocate(self.window[:, -1].contiguous(), num_particles=self.n_valid)
            eng._configure(self.neighbors, self.n_valid)

    def result(self):
        # the list a per-step caller would hold now: built (lazily) on the positions the last step started from
        eng, window, nbrs = self.engine, self.window, self.neighbors
        ref = window[:, -2] if (self.n_steps > 0 and window.shape[1] > 1) else window[:, -1]
        if self.n_steps > 0:  # edges of the last step are known on the host; the overflow bits are clear (retried away)
            out_nl = NeighborList(eng.h["neighbor_fn"], None, None, ref.clone(), nbrs.cell_list_capacity,
                                  nbrs.max_occupancy, nbrs._scratch, nbrs._grid, n_edges=self.n_edges)
        else:
            out_nl = NeighborList(eng.h["neighbor_fn"], None, nbrs._stats, ref.clone(), nbrs.cell_list_capacity,
                                  nbrs.max_occupancy, nbrs._scratch, nbrs._grid)
        return self.preds, out_nl


def run_batched(engines, jobs):
    """The reference ``vmap``s a batch of trajectories (``rollout.py:221-228``); here every trajectory of
    the batch gets its own engine and stream, and the chunks of all of them are in flight together --
    small clouds are launch / latency bound, so their step graphs overlap on the device.
    ``jobs``: one ``(window, particle_type, targets, n_steps, neighbors)`` per engine.
    Returns ``[(predictions, neighbors), ...]``."""
    live = [eng.start(*args) for eng, args in zip(engines, jobs)]
    while any(not j.finished for j in live):
        todo = [j for j in live if not j.finished]
        for j in todo:
            j.enqueue()
        for j in todo:
            j.collect()
    return [j.result() for j in live]


# ----------------------------------------------------------------------------- reference loop
def _forward_eval(params, state, sample, current_positions, target_positions, model_apply, case_integrate):
    """``rollout.py:31-75`` (generic path: any ``model_apply``)."""
    _, particle_type = sample
    pred, state = model_apply(params, state, sample)
    next_position = torch.as_tensor(case_integrate(pred, current_positions))
    mask = get_kinematic_mask(torch.as_tensor(particle_type).to(next_position.device))
    target_positions = torch.as_tensor(target_positions).to(next_position.device, next_position.dtype)
    next_position = torch.where(mask[:, None], target_positions, next_position)
    current_positions = torch.cat([current_positions[:, 1:], next_position[:, None, :]], dim=1)
    return current_positions, state


def _eval_batched_rollout(model_apply, case, params, state, traj_batch_i, neighbors, metrics_computer,
                          n_rollout_steps, t_window, n_extrap_steps=0, engine=None):
    """``rollout.py:78-178``.  The reference ``vmap``s over the batch; here the trajectories of a batch
    run concurrently, one engine and stream each (``run_batched``).  Returns
    ``(predictions (B, T, N, d), [metrics per trajectory], neighbors)``."""
    _cabi.require_cuda()
    h = case._lb200
    pos_input_batch = torch.as_tensor(traj_batch_i[0])
    particle_type_batch = torch.as_tensor(traj_batch_i[1])
    bsz, n_nodes, t_total, dim = pos_input_batch.shape
    if n_rollout_steps == -1:
        n_rollout_steps = t_total - t_window
    traj_len = n_rollout_steps + n_extrap_steps
    dev = torch.device("cuda")
    predictions, metrics = [], []
    currents, ptypes, targets_b, gts = [], [], [], []
    for b in range(bsz):
        pos_b = pos_input_batch[b].to(dev, h["dtype"])
        ptypes.append(particle_type_batch[b].to(dev, torch.int32))
        currents.append(pos_b[:, :t_window].contiguous())
        targets = pos_b[:, t_window:t_window + traj_len].permute(1, 0, 2).contiguous()  # (T, N, d)
        if targets.shape[0] < traj_len:  # extrapolation: JAX clamps the step index (rollout.py:158)
            pad = targets[-1:].expand(traj_len - targets.shape[0], -1, -1)
            targets = torch.cat([targets, pad], dim=0).contiguous()
        targets_b.append(targets)
        gts.append(pos_b[:, t_window:t_window + n_rollout_steps].permute(1, 0, 2))
    if engine is not None:
        # the batch axis of the reference's vmap: one engine (stream, scratch, step graph) per trajectory
        while len(engine._siblings) < bsz - 1:
            engine._siblings.append(RolloutEngine(engine.case, engine.model, engine._params, engine.steps_per_sync))
        engines = [engine] + engine._siblings[:bsz - 1]
        results = run_batched(engines, [(currents[b], ptypes[b], targets_b[b], traj_len, neighbors if b == 0 else None)
                                        for b in range(bsz)])
        predictions = [r[0] for r in results]
        neighbors = results[0][1]
    else:
        for b in range(bsz):
            current, ptype, targets = currents[b], ptypes[b], targets_b[b]
            preds = torch.empty((traj_len, n_nodes, dim), dtype=h["dtype"], device=dev)
            st, step = state, 0
            while step < traj_len:
                features, neighbors = case.preprocess_eval((current, ptype), neighbors)
                if bool(neighbors.did_buffer_overflow):  # rollout.py:135: blocking read
                    _, neighbors = case.allocate_eval((current, ptype))
                    continue
                current, st = _forward_eval(params, st, (features, ptype), current, targets[step], model_apply,
                                            case.integrate)
                preds[step] = current[:, -1]
                step += 1
            predictions.append(preds)
    for b in range(bsz):
        metrics.append(metrics_computer(predictions[b][:n_rollout_steps], gts[b]) if metrics_computer is not None else {})
    return torch.stack(predictions), metrics, neighbors


class SimpleLoader:
    """Minimal stand-in for the reference's torch ``DataLoader`` + ``numpy_collate``
    (``rollout.py:363-369``): yields ``(pos (B, N, T, d), particle_type (B, N))``."""

    def __init__(self, dataset, batch_size):
        self.dataset, self.batch_size = dataset, int(batch_size)

    def __iter__(self):
        n = len(self.dataset)
        for a in range(0, n, self.batch_size):
            items = [self.dataset[i] for i in range(a, min(n, a + self.batch_size))]
            yield (np.stack([np.asarray(p) for p, _ in items]), np.stack([np.asarray(t) for _, t in items]))


def _to_numpy(tree):
    if isinstance(tree, dict):
        return {k: _to_numpy(v) for k, v in tree.items()}
    if isinstance(tree, torch.Tensor):
        return tree.detach().cpu().numpy()
    return tree


def eval_rollout(model_apply, case, params, state, loader_eval, neighbors, metrics_computer, n_rollout_steps,
                 n_trajs, rollout_dir, out_type="none", n_extrap_steps=0):
    """``rollout.py:181-308``: roll out ``n_trajs`` trajectories, compute metrics, optionally
    write ``rollout_{i}.pkl`` with keys ``predicted_rollout`` / ``ground_truth_rollout`` /
    ``particle_type`` (``rollout.py:271-275``)."""
    batch_size = loader_eval.batch_size
    t_window = loader_eval.dataset.input_seq_length
    eval_metrics = {}
    if rollout_dir is not None:
        os.makedirs(rollout_dir, exist_ok=True)
    model = getattr(model_apply, "__self__", None)
    engine = None
    if isinstance(model, GNS) and case._lb200["force_mode"] != 2:
        engine = RolloutEngine(case, model, params)
    ind = -1
    for i, traj_batch_i in enumerate(loader_eval):
        n_traj_left = n_trajs - i * batch_size
        if n_traj_left <= 0:
            break
        if n_traj_left < batch_size:
            traj_batch_i = tuple(x[:n_traj_left] for x in traj_batch_i)
        rollout_batch, metrics_batch, neighbors = _eval_batched_rollout(
            model_apply, case, params, state, traj_batch_i, neighbors, metrics_computer, n_rollout_steps, t_window,
            n_extrap_steps, engine=engine)
        for j in range(rollout_batch.shape[0]):
            ind = i * batch_size + j
            eval_metrics[f"rollout_{ind}"] = metrics_batch[j]
            if rollout_dir is not None and out_type == "pkl":
                pos_input = np.asarray(traj_batch_i[0][j]).transpose(1, 0, 2)  # (t, nodes, dim)
                example_full = np.concatenate([pos_input[:t_window], rollout_batch[j].cpu().numpy()])
                payload = {"predicted_rollout": example_full, "ground_truth_rollout": pos_input,
                           "particle_type": np.asarray(traj_batch_i[1][j])}
                with open(os.path.join(rollout_dir, f"rollout_{ind}.pkl"), "wb") as f:
                    pickle.dump(payload, f)
            elif rollout_dir is not None and out_type == "vtk":
                raise NotImplementedError("vtk output (evaluate/utils.py) is outside the rollout hot path")
        if ind + 1 >= n_trajs:
            break
    if rollout_dir is not None:
        t = time.strftime("%Y_%m_%d_%H_%M_%S", time.localtime())
        with open(f"{rollout_dir}/metrics{t}.pkl", "wb") as f:
            pickle.dump(_to_numpy(eval_metrics), f)
    return eval_metrics


def infer(model, case, data_test, params=None, state=None, load_ckp=None, cfg_eval_infer=None, rollout_dir=None,
          n_rollout_steps=20, seed=0):
    """``rollout.py:311-399``."""
    assert params is not None or load_ckp is not None, \
        "Either params or a load_ckp directory must be provided for inference."
    cfg = merged("eval.infer", cfg_eval_infer)
    n_trajs = cfg["n_trajs"]
    if n_trajs == -1:
        n_trajs = data_test.num_samples if hasattr(data_test, "num_samples") else len(data_test)
    if params is not None:
        state = {} if state is None else state
    else:
        params, state, _, _ = load_haiku(load_ckp)
    set_seed(seed)
    loader_test = SimpleLoader(data_test, cfg["batch_size"])
    metrics_computer = MetricsComputer(cfg["metrics"], case.displacement, data_test.metadata,
                                       data_test.input_seq_length, cfg["metrics_stride"])
    pos0, ptype0 = data_test[0]
    _, _, _, neighbors = case.allocate(seed, (pos0, ptype0))
    return eval_rollout(model.apply, case, params, state, loader_test, neighbors, metrics_computer, n_rollout_steps,
                        n_trajs, rollout_dir, cfg["out_type"], cfg["n_extrap_steps"])
