"""GPU gradient-level parity: `engine.flat.grad` after ONE backward of the CUDA path (TF32 tcgen05 contractions, fp32
accumulation) vs torch autograd on the fp32 oracle restatement, per parameter tensor, before clipping / Adam.

What was measured on B200 (profiles/r02_parity_probe*.txt) and why the tolerances are what they are:

* **Linearised networks** (LeakyReLU slope 1 on both sides, `grad_cases.linearised`): no activation mask exists, the
  deviation is the arithmetic of the contractions alone - 8.6e-4 rel-Frobenius, cosine 0.9999997.  Asserted at the
  SURVEY.md section 8d level: cosine >= 0.9999, rel-Frobenius <= 3e-3.
* **The real networks** (slope 0.2): a TF32-level perturbation (2^-11 relative) of a pre-activation that is itself within
  2^-11 of zero flips its LeakyReLU' mask, and a flipped element is off by a factor 5.  ~1e-4 of all elements flip; that
  alone is ~1-4e-2 rel-Frobenius on the conv gradients - for ANY TF32 implementation: stock PyTorch (cuDNN / cuBLAS with
  TF32 allowed, the reference's own GPU numerics) deviates from the same fp32 CPU result by 2.4e-2 (B=64) / 4.3e-2
  (B=200), this repo by 3.1e-2 / 5.6e-2.  Asserted: cosine >= 0.998, rel-Frobenius <= 8e-2, and per tensor no more than
  1.6x the deviation of stock TF32 PyTorch on the same minibatch (+2e-3 absolute slack for the tiny FC tensors).
"""
import pytest

import grad_cases as GC

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,Be", [(64, 0), (48, 16), (200, 0)])
def test_linearised_policy_gradients_are_tf32_exact(B, Be):
    with GC.linearised():
        got, ref = GC.policy_grads("cuda", B, Be)
    cos, rel = GC.compare(got, ref, 0.9999, 3e-3, "policy (slope 1)")
    print(f"linearised policy B={B} Be={Be}: worst cosine {cos:.7f}, worst rel-Frobenius {rel:.2e}")


@pytest.mark.parametrize("B", [32, 100])
def test_linearised_critic_weight_gradients_are_tf32_exact(B):
    """With slope 1 the critic is linear, dD/dx does not depend on x and the expert / policy bias terms cancel almost
    exactly (their reference norms are ~1e-4 of the weight gradients'), so the bias gradients are compared on the scale of
    their layer's weight gradient; the weight gradients (Wasserstein term + hand-derived penalty term) at 3e-3."""
    with GC.linearised():
        got, ref, gs, rs = GC.critic_grads("cuda", B)
    e = GC.rel_errors(got, ref)
    for k, (cos, rel) in e.items():
        if k.endswith(".weight"):
            assert cos >= 0.9999 and rel <= 3e-3, f"critic (slope 1) {k}: cosine {cos:.6f}, rel-Frobenius {rel:.3e}"
        else:
            wk = k[:-len("bias")] + "weight"
            scale = max(float(ref[k].norm()), 1e-2 * float(ref[wk].norm()))
            err = float((got[k].double() - ref[k].double()).norm())
            assert err <= 3e-3 * scale, f"critic (slope 1) {k}: |diff| {err:.3e} vs scale {scale:.3e}"
    assert abs(gs["gp"] - rs["gp"]) <= 1e-4 + 1e-3 * abs(rs["gp"])


@pytest.mark.parametrize("B,Be,clipped", [(64, 0, True), (48, 16, True), (200, 0, True), (32, 0, False)])
def test_policy_gradients_match_autograd_gpu(B, Be, clipped):
    got, ref = GC.policy_grads("cuda", B, Be, clipped=clipped)
    cos, rel = GC.compare(got, ref, 0.998, 8e-2, "policy")
    print(f"policy B={B} Be={Be}: worst cosine {cos:.6f}, worst rel-Frobenius {rel:.2e}")


@pytest.mark.parametrize("B", [64, 200])
def test_policy_gradients_no_worse_than_stock_tf32(B):
    tf32, fp32 = GC.stock_tf32_policy_grads(B)
    got, ref = GC.policy_grads("cuda", B, 0)
    stock, mine = GC.rel_errors(tf32, fp32), GC.rel_errors(got, ref)
    for k in mine:
        assert mine[k][1] <= 1.6 * stock[k][1] + 2e-3, \
            f"{k}: rel-Frobenius {mine[k][1]:.3e} vs stock TF32 PyTorch {stock[k][1]:.3e} (both against fp32 CPU autograd)"


@pytest.mark.parametrize("B", [32, 100])
def test_critic_gradients_match_autograd_gpu(B):
    got, ref, gs, rs = GC.critic_grads("cuda", B)
    cos, rel = GC.compare(got, ref, 0.998, 8e-2, "critic")
    print(f"critic B={B}: worst cosine {cos:.6f}, worst rel-Frobenius {rel:.2e}; wd {gs['wd']:.6f}/{rs['wd']:.6f} gp {gs['gp']:.6f}/{rs['gp']:.6f}")
    assert abs(gs["wd"] - rs["wd"]) <= 1e-4 + 1e-2 * abs(rs["wd"])
    assert abs(gs["gp"] - rs["gp"]) <= 1e-4 + 3e-3 * abs(rs["gp"])
