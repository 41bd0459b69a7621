"""GPU parity of the tcgen05 implicit-GEMM engine, op by op, through the C ABI.

Integer-valued inputs are exact in TF32, so those cases must match the CPU statement (oracle/abi_emu.py) bit for bit
in value: any mismatch is a descriptor / layout bug.  Random fp32 inputs check the TF32 rounding level:
relative Frobenius error <= 2e-3 (TF32 has a 10-bit mantissa; TMA converts fp32->tf32 round-to-nearest)."""
import pytest
import torch

import gpu_probe_umma as P

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", P.CASES)
def test_contraction_exact_on_integers(case):
    assert P.run_case(case)


# (2176 = 17 row tiles: CTA-pair mode with a phantom 18th tile; 4096 x 512: plain pair mode; the others: single-CTA paths)
@pytest.mark.parametrize("M,N,K", [(256, 512, 512), (384, 100, 25632), (130, 3 * 32, 100), (2176, 512, 544), (4096, 256, 512),
                                   (2100, 160, 96)])
def test_linear_tf32_rounding_level(M, N, K):
    from gail_carla_b200 import _abi as A
    g = torch.Generator().manual_seed(M + N + K)
    ld = (K + 3) // 4 * 4
    x = torch.zeros(M, ld); x[:, :K] = torch.randn(M, K, generator=g)
    w = torch.zeros(N, ld); w[:, :K] = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    ldy = (N + 3) // 4 * 4
    y = torch.zeros(M, ldy, device="cuda")
    A.linear_fwd(x.cuda(), ld, w.cuda(), ld, b.cuda(), y, ldy, M, N, K, A.EPI_BIAS, 0.2, 1)
    ref = x[:, :K].double() @ w[:, :K].double().t() + b.double()
    rel = ((y[:, :N].cpu().double() - ref).norm() / ref.norm()).item()
    assert rel < 2e-3, rel
    # every row individually (a mis-paired or clipped tile would leave whole rows wrong while the norm barely moves)
    row_rel = (y[:, :N].cpu().double() - ref).norm(dim=1) / ref.norm(dim=1)
    assert row_rel.max().item() < 1e-2, row_rel.argmax().item()


def test_conv_stack_tf32_vs_fp64():
    """conv1..conv4 forward through the engine vs torch fp64 conv on the same normalised input."""
    import torch.nn.functional as F
    from gail_carla_b200 import _abi as A, engine as E
    g = torch.Generator().manual_seed(3)
    B = 3
    obs = torch.rand(B, 3, 192, 192, generator=g)
    x0 = torch.zeros(B, 96 * 96 * 16, device="cuda")
    A.gather_obs_s2d(obs.cuda(), None, x0, B)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1); std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    ref = ((obs - mean) / std).double()
    acts = [x0, torch.zeros(B, 96 * 96 * 32, device="cuda"), torch.zeros(B, 46 * 46 * 64, device="cuda"),
            torch.zeros(B, 22 * 22 * 128, device="cuda"), torch.zeros(B, A.LDF, device="cuda")]
    for i in range(1, 5):
        cin, cout = E.CONV_CH[i - 1], E.CONV_CH[i]
        w = torch.randn(cout, cin, 4, 4, generator=g) / (cin * 16) ** 0.5; b = torch.randn(cout, generator=g) * 0.1
        n = 2048 if i == 1 else cout * cin * 16
        wf = torch.zeros(n, device="cuda")
        A.prep_conv_weight(w.cuda(), wf, None, cout, cin, i == 1)
        A.conv_fprop(E.conv_geom(i, B), acts[i - 1], wf, b.cuda(), acts[i], A.EPI_BIAS_LRELU, 0.2)
        ref = F.leaky_relu(F.conv2d(ref, w.double(), b.double(), stride=2), 0.2)
        if i == 1:
            got = acts[1].view(B, 96, 96, 32)[:, :95, :95].permute(0, 3, 1, 2)
        elif i == 4:
            got = acts[4][:, :25600].view(B, 10, 10, 256).permute(0, 3, 1, 2)
        else:
            s = ref.shape[-1]
            got = acts[i].view(B, s, s, cout).permute(0, 3, 1, 2)
        rel = ((got.cpu().double() - ref).norm() / ref.norm()).item()
        assert rel < 3e-3, (i, rel)


@pytest.mark.parametrize("layer", [2, 3, 4])
def test_bitmask_path_equals_fp32_mask_path(layer):
    """The LeakyReLU' bit masks written by the fprop epilogue give the same dgrad / masked-fprop results as TMA-loading the
    fp32 activation (integer inputs -> bit-exact)."""
    from gail_carla_b200 import _abi as A
    from oracle import abi_emu as E_
    B = 3
    g = P.geom(layer, B)
    nin, nout = B * g.in_batch_stride, B * g.out_batch_stride
    Kw = g.KH * g.KW * g.Cin
    x = P.ints((nin,), seed=1); w = P.ints((g.Cout * Kw,), -1, 2, seed=3); bias = P.ints((g.Cout,), seed=4)
    y = torch.zeros(nout, device="cuda"); bits = torch.zeros(nout // 32, dtype=torch.int32, device="cuda")
    A.conv_fprop(g, x.cuda(), w.cuda(), bias.cuda(), y, A.EPI_BIAS_LRELU, 0.5, mask_bits=bits)
    y_ref = torch.zeros(nout); E_.conv_fprop(g, x, w, bias, y_ref, 1, 0.5)
    assert torch.equal(y.cpu(), y_ref)
    # masked fprop (second-order chain): bits vs fp32 mask source
    x2 = P.ints((nin,), seed=7)
    o_bits = torch.zeros(nout, device="cuda"); o_ref = torch.zeros(nout)
    A.conv_fprop(g, x2.cuda(), w.cuda(), None, o_bits, A.EPI_MASK, 0.5, mask_src=y, mask_bits=bits)
    E_.conv_fprop(g, x2, w, None, o_ref, 3, 0.5, mask_src=y_ref)
    assert torch.equal(o_bits.cpu(), o_ref)
    # dgrad of the NEXT layer would land on y; emulate with this layer's geometry: mask = sign of x-shaped activation
    act = P.ints((nin,), seed=9)
    if g.Cin % 32 == 0 and g.in_batch_stride % 32 == 0:
        abits = torch.zeros(nin // 32, dtype=torch.int32)
        pos = (act > 0).view(-1, 32).to(torch.int64)
        abits = (pos << torch.arange(32)).sum(1)
        abits = torch.where(abits >= 2 ** 31, abits - 2 ** 32, abits).to(torch.int32)
        dy = P.ints((nout,), seed=11); wd = P.ints((g.Cout * Kw,), -1, 2, seed=13)
        dx = torch.zeros(nin, device="cuda"); dx_ref = torch.zeros(nin)
        A.conv_dgrad(g, dy.cuda(), wd.cuda(), dx, act.cuda(), 0.5, mask_bits=abits.cuda())
        E_.conv_dgrad(g, dy, wd, dx_ref, act, 0.5)
        v = lambda t: E_._in_view(g, t)[:, :g.H, :g.W]
        assert torch.equal(v(dx.cpu()), v(dx_ref))


def _bits_of(act):
    pos = (act > 0).view(-1, 32).to(torch.int64)
    b = (pos << torch.arange(32)).sum(1)
    return torch.where(b >= 2 ** 31, b - 2 ** 32, b).to(torch.int32)


@pytest.mark.parametrize("layer,B,nb", [(2, 3, 2), (3, 5, 3), (4, 7, 7), (4, 131, 100), (3, 40, 0)])
def test_dgrad_epilogue_bias_gradient_is_exact(layer, B, nb):
    """gc_conv_dgrad(dbias_in=...): per-channel sums of the masked dx over the first `nb` samples, taken from the staged
    output tiles (integer inputs -> bit-exact against the CPU statement; samples >= nb are stored but not summed; clipped
    rows of partial tiles and the phantom tile of CTA-pair mode do not count)."""
    from gail_carla_b200 import _abi as A
    from oracle import abi_emu as E_
    g = P.geom(layer, B)
    nin, nout = B * g.in_batch_stride, B * g.out_batch_stride
    Kw = g.KH * g.KW * g.Cin
    act = P.ints((nin,), seed=9)
    dy = P.ints((nout,), -1, 2, seed=11); wd = P.ints((g.Cout * Kw,), -1, 2, seed=13)
    dy = dy * (torch.rand(nout, generator=torch.Generator().manual_seed(5)) < 0.25)      # sparse: keeps every sum exact in fp32
    dx = torch.zeros(nin, device="cuda"); dx_ref = torch.zeros(nin)
    db = torch.full((g.Cin,), 3.0, device="cuda"); db_ref = torch.full((g.Cin,), 3.0)       # accumulates (+=)
    A.conv_dgrad(g, dy.cuda(), wd.cuda(), dx, act.cuda(), 0.5, mask_bits=_bits_of(act).cuda(), dbias_in=db, dbias_samples=nb)
    E_.conv_dgrad(g, dy, wd, dx_ref, act, 0.5, dbias_in=db_ref, dbias_samples=nb)
    v = lambda t: E_._in_view(g, t)[:, :g.H, :g.W]
    assert torch.equal(v(dx.cpu()), v(dx_ref))
    assert db_ref.abs().max() < 2 ** 22 and torch.equal(db.cpu(), db_ref), (db.cpu() - db_ref).abs().max()


@pytest.mark.parametrize("M,rows", [(300, 200), (4096, 4096), (517, 0)])
def test_linear_dgrad_epilogue_column_sums_are_exact(M, rows):
    from gail_carla_b200 import _abi as A
    from oracle import abi_emu as E_
    N, K, mod = 1024, 96, 256
    dy = P.ints((M, K), -1, 2, seed=1); w = P.ints((K, N), -1, 2, seed=2); act = P.ints((M, N), seed=3)
    dx = torch.zeros(M, N, device="cuda"); dx_ref = torch.zeros(M, N)
    cs = torch.zeros(mod, device="cuda"); cs_ref = torch.zeros(mod)
    A.linear_dgrad(dy.cuda(), K, w.cuda(), N, dx, N, M, N, K, mask_src=act.cuda(), ldm=N, slope=0.5, mask_bits=_bits_of(act).cuda(),
                   colsum=cs, colsum_mod=mod, colsum_rows=rows)
    E_.linear_dgrad(dy, K, w, N, dx_ref, N, M, N, K, mask_src=act, ldm=N, slope=0.5, colsum=cs_ref, colsum_mod=mod, colsum_rows=rows)
    assert torch.equal(dx.cpu(), dx_ref)
    assert cs_ref.abs().max() < 2 ** 22 and torch.equal(cs.cpu(), cs_ref), (cs.cpu() - cs_ref).abs().max()
