"""Oracle: rollout metrics (test infrastructure only).

Restates ``lagrangebench/evaluate/metrics.py`` in NumPy for the metrics that live on the rollout
path: ``MetricsComputer.__call__`` for ``mse`` / ``mae`` with the horizon slices
(``metrics.py:88-103,139-147``), ``e_kin`` (``metrics.py:105-131,157-160``) and
``averaged_metrics`` (``metrics.py:233-252``).  ``sinkhorn`` (an optimal-transport solve in
``ott`` / ``pot``, third-party) is not on the path.  Parity unpinned by the reference's own tests
(it ships none for the metrics); the arithmetic is elementwise and is restated line by line.
"""

import numpy as np

LOSS_RANGES = [1, 5, 10, 20, 50, 100]


def compute(active_metrics, dist_fn, metadata, pred_rollout, target_rollout, stride=10, loss_ranges=None):
    """``MetricsComputer(active_metrics, dist_fn, metadata, ...)(pred, target)``; rollouts ``(T, N, d)``."""
    loss_ranges = LOSS_RANGES if loss_ranges is None else loss_ranges
    pred = np.asarray(pred_rollout)
    target = np.asarray(target_rollout, dtype=pred.dtype)
    out = {}
    for name in active_metrics:
        if name in ("mse", "mae"):
            d = dist_fn(pred, target)  # vmap(vmap(dist_fn)) over (step, particle)
            per_step = (d ** 2).mean(axis=(1, 2)) if name == "mse" else np.abs(d).mean(axis=(1, 2))
            out[name] = per_step
            for i in loss_ranges:
                if i < per_step.shape[0]:
                    out[f"{name}{i}"] = per_step[:i]
        elif name == "e_kin":
            dt = metadata["dt"] * metadata["write_every"]
            dx, dim = metadata["dx"], metadata["dim"]

            def e_kin(roll):
                vel = dist_fn(roll[1::stride], roll[0:-1:stride]) / dt
                return (vel ** 2).sum(axis=2).sum(axis=1) * dx ** dim

            ep, et = e_kin(pred), e_kin(target)
            out[name] = {"predicted": ep, "target": et, "mse": ((ep - et) ** 2).mean()}
        else:
            raise NotImplementedError(name)
    return out


def averaged_metrics(eval_metrics):
    """``metrics.py:233-252``."""
    avg = {}
    for rollout in eval_metrics.values():
        for k, v in rollout.items():
            if k == "e_kin":
                v = v["mse"]
            if k in ("mse", "mae"):
                k = "loss"
            avg.setdefault(k, []).append(float(np.mean(v)))
    small = {f"val/{k}": float(np.mean(v)) for k, v in avg.items()}
    small.update({f"val/std{k}": float(np.std(v)) for k, v in avg.items()})
    return small
