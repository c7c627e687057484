"""Oracle: autoregressive rollout loop (test infrastructure only).

Restates ``lagrangebench/evaluate/rollout.py:31-178`` (``_forward_eval``,
``_eval_batched_rollout``), ``lagrangebench/utils.py:17-35`` (``NodeType``,
``get_kinematic_mask``) and the position-MSE of
``lagrangebench/evaluate/metrics.py:139-147`` in NumPy.  The reference ``vmap``s over the
batch; the oracle loops over it.  Pinned by ``tests/rollout_test.py:74-195``.
"""

import numpy as np

PAD_VALUE, FLUID, SOLID_WALL, MOVING_WALL, RIGID_BODY, SIZE = -1, 0, 1, 2, 3, 9


def get_kinematic_mask(particle_type):  # utils.py:28-35
    particle_type = np.asarray(particle_type)
    return (particle_type == SOLID_WALL) | (particle_type == MOVING_WALL) | (particle_type == PAD_VALUE)


def forward_eval(model_apply, case_integrate, params, state, sample, current_positions, target_positions):
    """``_forward_eval`` (``rollout.py:31-75``)."""
    _, particle_type = sample
    pred, state = model_apply(params, state, sample)
    next_position = case_integrate(pred, current_positions)
    mask = get_kinematic_mask(particle_type)
    next_position = np.where(mask[:, None], target_positions, next_position)
    current_positions = np.concatenate([current_positions[:, 1:], next_position[:, None, :]], axis=1)
    return current_positions, state


def eval_batched_rollout(model_apply, case, params, state, traj_batch_i, neighbors,
                         n_rollout_steps, t_window, n_extrap_steps=0):
    """``_eval_batched_rollout`` (``rollout.py:78-178``) without the metrics call.

    Returns ``(predictions (B, T, N, d), neighbors)``."""
    pos_input_batch, particle_type_batch = traj_batch_i
    bsz, n_nodes, _, dim = pos_input_batch.shape
    if n_rollout_steps == -1:
        n_rollout_steps = pos_input_batch.shape[2] - t_window
    traj_len = n_rollout_steps + n_extrap_steps
    predictions = np.zeros((bsz, traj_len, n_nodes, dim), dtype=pos_input_batch.dtype)
    for b in range(bsz):
        current = pos_input_batch[b, :, 0:t_window]
        targets = pos_input_batch[b, :, t_window:t_window + traj_len]
        ptype = particle_type_batch[b]
        st = state
        step = 0
        while step < traj_len:
            feats, neighbors = case.preprocess_eval((current, ptype), neighbors)
            if neighbors.did_buffer_overflow:  # rollout.py:135-151: re-allocate, retry the step
                _, neighbors = case.allocate_eval((current, ptype))
                continue
            # beyond the ground truth (extrapolation) the target slice is empty in JAX
            # (dynamic clamp); kinematic particles then keep the last available target
            tgt = targets[:, min(step, targets.shape[1] - 1)]
            current, st = forward_eval(model_apply, case.integrate, params, st, (feats, ptype), current, tgt)
            predictions[b, step] = current[:, -1]
            step += 1
    return predictions, neighbors


def mse(displacement_fn, pred, target):
    """Per-step position MSE with the case's displacement (``metrics.py:139-147``)."""
    d = displacement_fn(pred, target)
    return (d**2).mean(axis=(-1, -2))
