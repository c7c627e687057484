"""Forward-only piece of the training strategies that reuses the rollout path.

``push_forward_build`` mirrors ``lagrangebench/train/strats.py:110-159``: the pushforward trick unrolls
the solver for some steps WITHOUT gradients before the step that is trained on -- each unrolled step
is exactly one step of the rollout path (model forward -> integrate -> window shift -> neighbor /
feature update), so it runs on the same kernels.  The sampling of the unroll length
(``push_forward_sample_steps``) and the loss itself are training code and stay with the reference.
"""

import torch


def push_forward_build(model_apply, case):
    """-> ``push_forward_fn(features, current_pos, particle_type, neighbors, params, state)`` returning
    ``(current_pos, neighbors, features)`` after one unrolled step (``strats.py:137-159``; no buffer
    overflow check, as there)."""

    def push_forward_fn(features, current_pos, particle_type, neighbors, params, state):
        pred, _ = model_apply(params, state, (features, particle_type))
        current_pos = torch.as_tensor(current_pos)
        next_pos = torch.as_tensor(case.integrate(pred, current_pos))
        current_pos = torch.cat([current_pos.to(next_pos.device, next_pos.dtype)[:, 1:], next_pos[:, None, :]], dim=1)
        features, neighbors = case.preprocess_eval((current_pos, particle_type), neighbors)
        return current_pos, neighbors, features

    return push_forward_fn
