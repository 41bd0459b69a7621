"""Reference-API edge cases on the CUDA path (see tests/api_cases.py): differentiable ``evaluate_actions``,
``forward(gp=True)`` handles, expert loader shorter than the policy generator / ``compute_loss(batch_size=...)`` with a
dropped remainder, the +inf tail of ``predict_reward``.  TF32 tolerances as in tests/test_grads_gpu.py."""
import pytest

import api_cases as AC

pytestmark = pytest.mark.gpu


def test_evaluate_actions_backward_gpu():
    AC.evaluate_actions_is_differentiable("cuda", 48, 0.998, 8e-2)   # TF32 + mask flips, see tests/test_grads_gpu.py


def test_forward_gp_handles_gpu():
    AC.forward_gp_returns_first_order_handles("cuda", 16, 8e-2)


def test_expert_loader_edge_cases_gpu():
    AC.expert_loader_shorter_and_dropped_remainder("cuda", 2e-2)


def test_reward_inf_tail_gpu():
    AC.reward_saturates_to_inf("cuda")
