"""Size-independent properties of the tensor-core contractions at the full BASELINE batch (B = 4096 rows per launch,
configs[2]/[3]), where the CPU statements would take minutes: for every conv layer the three implicit GEMMs are adjoint
views of one trilinear form,

    <y, fprop(x; w)>  =  <dgrad(y; w), x>  =  <wgrad(y, x), w>,

evaluated on small-integer tensors (exact in TF32; products and the fp32 accumulations stay below 2^24, so the three
numbers must agree to fp64 round-off of the final dot products).  This exercises every tile / patch / slab / split-K path
of gc_conv_fprop / gc_conv_dgrad / gc_conv_wgrad, including the partial tiles at the image borders and the ragged last
batch tile, at sizes the oracle cannot reach; it also pins the conv1 bias-gradient column (pad channel = 1)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

CH = (3, 32, 64, 128, 256)


def _ints(shape, density, g, dev):
    v = torch.randint(-2, 3, shape, generator=g, device=dev, dtype=torch.int8)
    keep = torch.rand(shape, generator=g, device=dev) < density
    return (v * keep).float()


@pytest.mark.parametrize("layer,B", [(1, 4096), (2, 4096), (3, 4096), (4, 4096), (2, 4099), (3, 517)])
def test_conv_contractions_are_adjoint_at_full_batch(layer, B):
    from gail_carla_b200 import _abi as A, engine as E
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(100 + layer)
    geom = E.conv_geom4_compact(B) if layer == 4 else E.conv_geom(layer, B)
    cin, cout = CH[layer - 1], CH[layer]
    w = _ints((cout, cin, 4, 4), 1.0, g, dev)
    n_op = 2048 if layer == 1 else cout * cin * 16
    wf = torch.zeros(n_op, device=dev); wd = torch.zeros(n_op, device=dev)
    A.prep_conv_weight(w, wf, wd, cout, cin, layer == 1)
    x = _ints((B, geom.in_batch_stride), 0.5, g, dev)
    yshape = (B, geom.OHp, geom.OWp, geom.Cout)
    y = _ints(yshape, 1.0 / 64, g, dev)                     # sparse: keeps the wgrad sums (B*OH*OW terms) exact in fp32
    valid = torch.zeros(yshape, device=dev)
    valid[:, :geom.OH, :geom.OW] = 1                        # only valid output pixels take part
    y = (y * valid).reshape(B, -1)
    if y.shape[1] != geom.out_batch_stride:
        y = torch.nn.functional.pad(y, (0, geom.out_batch_stride - y.shape[1]))
    y = y.contiguous()
    # fprop
    fx = torch.zeros(B, geom.out_batch_stride, device=dev)
    A.conv_fprop(geom, x, wf, None, fx, A.EPI_STORE, 0.2)
    t_f = (y.double() * fx.double()).sum().item()
    # dgrad
    dx = torch.zeros(B, geom.in_batch_stride, device=dev)
    A.conv_dgrad(geom, y, wd, dx, None, 0.2)
    t_d = (dx.double() * x.double()).sum().item()
    # wgrad
    splits = A.conv_wgrad_splits(geom)
    part = torch.zeros(splits * n_op, device=dev)
    dw = torch.zeros(cout, cin, 4, 4, device=dev)
    db = torch.zeros(cout, device=dev)
    A.conv_wgrad(geom, y, x, part, splits)
    A.unprep_conv_wgrad(part, splits, dw, cout, cin, layer == 1, db if layer == 1 else None)
    t_w = (dw.double() * w.double()).sum().item()
    torch.cuda.synchronize()
    scale = max(1.0, abs(t_f))
    assert abs(t_f - t_d) <= 1e-9 * scale, (t_f, t_d)
    assert abs(t_f - t_w) <= 1e-9 * scale, (t_f, t_w)
    assert abs(t_f) > 0
    if layer == 1:
        # pad channel (index 3 of every (dy,dx) group) of tap (0,0): with x_pad = 1 the column is sum_pixels y[n]
        x1 = x.view(B, 96, 96, 4, 4).clone(); x1[..., 3] = 1.0
        A.conv_wgrad(geom, y, x1.view(B, -1), part, splits)
        A.unprep_conv_wgrad(part, splits, dw, cout, cin, True, db)
        ref = y.view(B, geom.OHp, geom.OWp, cout).double().sum((0, 1, 2))
        assert torch.equal(db.double(), ref), (db - ref.float()).abs().max().item()
