"""GPU gradient-level parity: `engine.flat.grad` after ONE backward of the CUDA path (TF32 tcgen05 contractions, fp32
accumulation) vs torch autograd on the fp32 oracle restatement, per parameter tensor, before clipping / Adam.

Tolerance: cosine >= 0.999 and ||got - ref||_F <= 3e-3 ||ref||_F for every tensor.  B = 64 keeps the
CPU autograd reference (double backward included) at a few seconds; B = 200 adds ragged M tiles and several split-K
partials per weight gradient."""
import pytest

import grad_cases as GC

pytestmark = pytest.mark.gpu

MIN_COS, MAX_REL = 0.999, 3e-3


@pytest.mark.parametrize("B,Be,clipped", [(64, 0, True), (48, 16, True), (200, 0, True), (32, 0, False)])
def test_policy_gradients_match_autograd_gpu(B, Be, clipped):
    got, ref = GC.policy_grads("cuda", B, Be, clipped=clipped)
    cos, rel = GC.compare(got, ref, MIN_COS, MAX_REL, "policy")
    print(f"policy B={B} Be={Be}: worst cosine {cos:.6f}, worst rel-Frobenius {rel:.2e}")


@pytest.mark.parametrize("B", [32, 100])
def test_critic_gradients_match_autograd_gpu(B):
    got, ref, gs, rs = GC.critic_grads("cuda", B)
    cos, rel = GC.compare(got, ref, MIN_COS, MAX_REL, "critic")
    print(f"critic B={B}: worst cosine {cos:.6f}, worst rel-Frobenius {rel:.2e}; wd {gs['wd']:.6f}/{rs['wd']:.6f} gp {gs['gp']:.6f}/{rs['gp']:.6f}")
    assert abs(gs["wd"] - rs["wd"]) <= 1e-4 + 3e-3 * abs(rs["wd"])
    assert abs(gs["gp"] - rs["gp"]) <= 1e-4 + 3e-3 * abs(rs["gp"])
