"""Host-side orchestration of the trunks: parameter flattening, operand-layout preparation, workspaces, and the
forward / hand-derived backward passes of the policy (tools/model.py:56-128) and of the WDGAIL critic including the
gradient-penalty second-order pass (algo/wdgail.py:40-98).  All arithmetic happens in the C-ABI kernels (``_abi``);
this module only sequences launches and owns device buffers (allocated through torch, i.e. plumbing).

Layouts: images are space-to-depth NHWC ``[B,96,96,16]`` (normalised, pad channel 0), conv
activations NHWC (conv1 output with row/column pitch 96), conv4 output lands directly in the feature matrix
``F[B, LDF]`` (LDF = 25600 + 32) whose last 32 columns hold the 13 metric features (+2 action columns for the critic).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _abi as A
from ._abi import ConvGeom, LDF, S2D_PER_SAMPLE, EPI_BIAS, EPI_BIAS_LRELU, EPI_MASK, EPI_STORE

SLOPE = 0.2                      # nn.LeakyReLU(0.2) everywhere (tools/model.py:138-144,95-99,112)
FEAT = 25600
CONV_CH = (3, 32, 64, 128, 256)
INV_STD = (1.0 / 0.229, 1.0 / 0.224, 1.0 / 0.225)   # tools/model.py:155
ALIGN = 64                       # every parameter starts on a 256-byte boundary inside the flat buffers


def conv_geom(layer: int, B: int) -> ConvGeom:
    """Implicit-GEMM geometry of conv `layer` (1-based) of ProcessObsFeatures (tools/model.py:137-143)."""
    if layer == 1:   # k4,s2 on [192,192,3] == k2,s1 on the space-to-depth image [96,96,16]
        return ConvGeom(B, 96, 96, 96, 96, 16, 2, 2, 1, 95, 95, 96, 96, 32, 96 * 96 * 16, 96 * 96 * 32)
    if layer == 2:
        return ConvGeom(B, 95, 95, 96, 96, 32, 4, 4, 2, 46, 46, 46, 46, 64, 96 * 96 * 32, 46 * 46 * 64)
    if layer == 3:
        return ConvGeom(B, 46, 46, 46, 46, 64, 4, 4, 2, 22, 22, 22, 22, 128, 46 * 46 * 64, 22 * 22 * 128)
    if layer == 4:
        return ConvGeom(B, 22, 22, 22, 22, 128, 4, 4, 2, 10, 10, 10, 10, 256, 22 * 22 * 128, LDF)
    raise ValueError(layer)


def conv_geom4_compact(B: int) -> ConvGeom:
    """conv4 with a compact [B,10,10,256] output side (used for gradients w.r.t. the conv features)."""
    return ConvGeom(B, 22, 22, 22, 22, 128, 4, 4, 2, 10, 10, 10, 10, 256, 22 * 22 * 128, FEAT)


ACT_ELEMS = (S2D_PER_SAMPLE, 96 * 96 * 32, 46 * 46 * 64, 22 * 22 * 128)   # per-sample floats of X0, A1, A2, A3


class FlatParams:
    """All parameters of a module as views into one flat fp32 buffer (same for grads and Adam moments), so that
    clip_grad_norm_ + Adam (algo/ppo.py:115-119) is two launches and the NCCL gradient all-reduce is one bucket."""

    def __init__(self, module: nn.Module):
        self.module = module
        self.names: List[str] = []
        self.offsets: Dict[str, int] = {}
        self.flat = self.grad = None
        self.rebuild()

    def rebuild(self) -> None:
        params = list(self.module.named_parameters())
        dev = params[0][1].device
        off = 0
        self.names, self.offsets = [], {}
        for n, p in params:
            self.names.append(n)
            self.offsets[n] = off
            off += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        self.numel = off
        flat = torch.zeros(off, dtype=torch.float32, device=dev)
        grad = torch.zeros(off, dtype=torch.float32, device=dev)
        for n, p in params:
            o = self.offsets[n]
            flat[o:o + p.numel()].copy_(p.data.reshape(-1).float())
            p.data = flat[o:o + p.numel()].view(p.shape)
            p.grad = grad[o:o + p.numel()].view(p.shape)
        self.flat, self.grad = flat, grad
        self.grad_clean = True     # the gradient buffer is all zeros (fresh, or zeroed by the fused clip+Adam kernel)

    def begin_backward(self) -> None:
        """optimizer.zero_grad() (algo/ppo.py:115, algo/wdgail.py:140): a no-op when the last optimiser step already
        left zeros behind (gc_clip_adam zero_grad=1), else one memset."""
        if not self.grad_clean:
            self.grad.zero_()
        self.grad_clean = False

    def span(self, first: str, last: str = None):
        """[lo, hi) element range of the flat buffers covering parameters `first`..`last` (inclusive, declaration order)."""
        names = self.names
        i0 = names.index(first)
        i1 = names.index(last) if last is not None else len(names) - 1
        hi = self.offsets[names[i1 + 1]] if i1 + 1 < len(names) else self.numel
        return self.offsets[names[i0]], hi

    def version(self) -> int:
        """Sum of the parameters' autograd version counters: changes whenever torch code (an external optimiser,
        load_state_dict, a manual ``p.data.copy_``) writes a parameter in place, so stale operand copies are rebuilt."""
        return sum(p._version for p in self.module.parameters())

    def ok(self) -> bool:
        """True while every parameter is still a view of the flat buffer (``.to()`` / ``.cpu()`` break that)."""
        base = self.flat.data_ptr()
        for n, p in self.module.named_parameters():
            if p.data_ptr() != base + 4 * self.offsets[n] or p.device != self.flat.device:
                return False
        return True

    def restore_grads(self) -> None:
        """Re-attach ``p.grad`` to the flat gradient buffer.  A stock ``torch.optim`` optimiser's ``zero_grad()`` sets the
        gradients to ``None`` (set_to_none=True is the default); None means zero, a foreign tensor is copied in."""
        base = self.grad.data_ptr()
        for n, p in self.module.named_parameters():
            o = self.offsets[n]
            if p.grad is None or p.grad.data_ptr() != base + 4 * o:
                view = self.grad[o:o + p.numel()].view(p.shape)
                if p.grad is None:
                    view.zero_()
                else:
                    view.copy_(p.grad)
                    self.grad_clean = False
                p.grad = view

    def ensure(self) -> bool:
        if not self.ok():
            self.rebuild()
            return True
        self.restore_grads()
        return False

    def g(self, name: str) -> torch.Tensor:
        return self.module.get_parameter(name).grad

    def p(self, name: str) -> torch.Tensor:
        return self.module.get_parameter(name).data


class Workspace:
    """Device buffers for one trunk at a given row capacity.  Gradient buffers are created on first backward."""

    def __init__(self, device, rows: int, with_input_grad: bool):
        self.device, self.rows = device, rows
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=device)
        self.X0 = z(rows, S2D_PER_SAMPLE)
        self.A = [self.X0, z(rows, ACT_ELEMS[1]), z(rows, ACT_ELEMS[2]), z(rows, ACT_ELEMS[3])]
        self.F = z(rows, LDF)
        # LeakyReLU' bit masks of A1, A2, A3 and F (1 bit per fp32 element, written by the conv fprop epilogues)
        zi = lambda n: torch.zeros(rows, n, dtype=torch.int32, device=device)
        self.mbits = [None, zi(ACT_ELEMS[1] // 32), zi(ACT_ELEMS[2] // 32), zi(ACT_ELEMS[3] // 32), zi(LDF // 32)]
        self.dA: Optional[List[torch.Tensor]] = None
        self.with_input_grad = with_input_grad
        self.part: Dict[str, torch.Tensor] = {}
        self.small: Dict[str, torch.Tensor] = {}

    def release(self) -> None:
        """Drop every buffer (engines may still hold a reference to a workspace that has been replaced by a larger one)."""
        self.X0 = self.F = self.dA = None
        self.A, self.mbits, self.part, self.small, self.rows = [], [], {}, {}, 0

    def grads(self):
        if self.dA is None:
            z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=self.device)
            self.dA = [z(self.rows, S2D_PER_SAMPLE) if self.with_input_grad else None,
                       z(self.rows, ACT_ELEMS[1]), z(self.rows, ACT_ELEMS[2]), z(self.rows, ACT_ELEMS[3]),
                       z(self.rows, FEAT)]
            self.dFt = z(self.rows, 32)
        return self.dA

    def buf(self, key: str, *shape) -> torch.Tensor:
        t = self.small.get(key)
        n = math.prod(shape)
        if t is None or t.numel() < n:
            t = torch.zeros(n, dtype=torch.float32, device=self.device)
            self.small[key] = t
        return t[:n].view(*shape)

    def partial(self, key: str, numel: int) -> torch.Tensor:
        t = self.part.get(key)
        if t is None or t.numel() < numel:
            t = torch.empty(numel, dtype=torch.float32, device=self.device)
            self.part[key] = t
        return t


_SHARED: Dict[str, Workspace] = {}


def shared_workspace(device, rows: int, with_input_grad: bool) -> Workspace:
    """One activation / gradient workspace per device, shared by the policy and the critic (their updates never
    overlap in time), grown on demand.  At B=4096 the critic's 3B-row workspace is ~70 GB, so sharing matters."""
    key = str(device)
    ws = _SHARED.get(key)
    if ws is None or ws.rows < rows:
        if ws is not None:
            _SHARED.pop(key)
            ws.release()           # PolicyEngine.ws / CriticEngine.ws may still point at it: free its buffers now
            del ws
            if torch.cuda.is_available():
                torch.cuda.empty_cache()
        ws = Workspace(device, rows, with_input_grad)
        _SHARED[key] = ws
    elif with_input_grad and not ws.with_input_grad:
        ws.with_input_grad = True
        if ws.dA is not None:
            ws.dA[0] = torch.zeros(ws.rows, S2D_PER_SAMPLE, dtype=torch.float32, device=device)
    return ws


def release_workspaces() -> None:
    """Free every shared workspace (they are re-created on demand)."""
    for ws in _SHARED.values():
        ws.release()
    _SHARED.clear()
    if torch.cuda.is_available():
        torch.cuda.empty_cache()


def _splits_for(m_tiles: int, n_tiles: int, k_iters: int, cap: int = 8) -> int:
    tiles = max(1, m_tiles * n_tiles)
    s = max(1, min(cap, (2 * 148) // tiles))
    return max(1, min(s, k_iters // 8 if k_iters >= 16 else 1))


class ConvStack:
    """The four Conv2d(k4,s2)+LeakyReLU layers + flatten of ProcessObsFeatures (tools/model.py:131-164)."""

    def __init__(self, flat: FlatParams, prefix: str, need_input_grad: bool):
        self.flat, self.prefix, self.need_input_grad = flat, prefix, need_input_grad
        self.wf: List[torch.Tensor] = []
        self.wd: List[Optional[torch.Tensor]] = []
        self._dev = None
        self.fused_dbias = True      # conv2-4 bias gradients from the dgrad epilogues (False: gc_colsum passes)

    def wname(self, i: int) -> str:
        return f"{self.prefix}main.{2 * (i - 1)}.weight"

    def bname(self, i: int) -> str:
        return f"{self.prefix}main.{2 * (i - 1)}.bias"

    def prepare(self) -> None:
        """(Re)build the GEMM operand copies of the conv weights after the parameters changed."""
        dev = self.flat.flat.device
        if self._dev != dev:
            self.wf, self.wd = [], []
            for i in range(1, 5):
                n = 2048 if i == 1 else CONV_CH[i] * CONV_CH[i - 1] * 16
                self.wf.append(torch.zeros(n, dtype=torch.float32, device=dev))
                need_wd = i > 1 or self.need_input_grad
                self.wd.append(torch.zeros(n, dtype=torch.float32, device=dev) if need_wd else None)
            self._dev = dev
        for i in range(1, 5):
            A.prep_conv_weight(self.flat.p(self.wname(i)), self.wf[i - 1], self.wd[i - 1], CONV_CH[i], CONV_CH[i - 1], i == 1)

    @staticmethod
    def bits(ws: Workspace, i: int, row0: int):
        """LeakyReLU' bit mask of activation i (A1, A2, A3 or F), written by the forward epilogues when training."""
        return None if i < 1 else ws.mbits[i][row0:]

    def forward(self, ws: Workspace, B: int, row0: int = 0, training: bool = True) -> None:
        """X0[row0:row0+B] -> A1, A2, A3 -> F[:, :25600] (bias + LeakyReLU fused in the GEMM epilogue)."""
        for i in range(1, 5):
            x = ws.A[i - 1][row0:]
            y = ws.F[row0:] if i == 4 else ws.A[i][row0:]
            A.conv_fprop(conv_geom(i, B), x, self.wf[i - 1], self.flat.p(self.bname(i)), y, EPI_BIAS_LRELU, SLOPE,
                         mask_bits=self.bits(ws, i, row0) if training else None)

    def forward_masked(self, ws: Workspace, B: int, row0: int) -> None:
        """Second-order chain of the gradient penalty: v_k = LeakyReLU'(a_k) * conv_k(v_{k-1}) written in place of
        a_k (no bias) for rows [row0,row0+B).  ws.X0 rows must already hold u = d gp / d g."""
        for i in range(1, 5):
            x = ws.A[i - 1][row0:]
            y = ws.F[row0:] if i == 4 else ws.A[i][row0:]
            A.conv_fprop(conv_geom(i, B), x, self.wf[i - 1], None, y, EPI_MASK, SLOPE, mask_src=y, mask_bits=self.bits(ws, i, row0))

    def backward_data(self, ws: Workspace, B: int, row0: int = 0, B_bias: int = 0) -> None:
        """delta_4 (= ws.dA[4], already multiplied by LeakyReLU'(a_4)) -> delta_3, delta_2, delta_1 for rows
        [row0,row0+B); each dgrad epilogue applies LeakyReLU' of the layer it lands on.  With B_bias > 0 the dgrads that
        produce delta_3 and delta_2 also leave conv3's / conv2's bias gradient (sum of delta over the pixels of the first
        B_bias samples) in the flat gradient buffer - taken from the staged output tiles, not from another pass."""
        dA = ws.grads()
        for i in (4, 3, 2):
            g = conv_geom4_compact(B) if i == 4 else conv_geom(i, B)
            fused = B_bias > 0 and i > 2          # delta_1's bias gradient is a column of conv1's wgrad (see backward_params)
            A.conv_dgrad(g, dA[i][row0:], self.wd[i - 1], dA[i - 1][row0:], ws.A[i - 1][row0:], SLOPE,
                         mask_bits=self.bits(ws, i - 1, row0),
                         dbias_in=self.flat.g(self.bname(i - 1)) if fused else None, dbias_samples=B_bias if fused else 0)

    def input_grad(self, ws: Workspace, B: int, row0: int) -> None:
        """dX0 = conv1^T(delta_1) for rows [row0,row0+B): the dD/dx of algo/wdgail.py:85-91 (normalised-input space)."""
        dA = ws.grads()
        A.conv_dgrad(conv_geom(1, B), dA[1][row0:], self.wd[0], dA[0][row0:], None, SLOPE)

    def backward_params(self, ws: Workspace, B: int, B_bias: int) -> None:
        """Weight gradients from (delta_k, a_{k-1}) over rows [0,B); bias gradients over rows [0,B_bias).
        conv1's bias gradient is a column of its wgrad: the pad channel of the space-to-depth image is 1.0 for image
        rows and 0.0 for the gradient-penalty rows (u = d gp / d g written by gc_grad_penalty), which are exactly the
        rows [B_bias, B) that must not contribute - so dY1, the largest gradient tensor, is not re-read."""
        dA = ws.grads()
        for i in range(1, 5):
            g = conv_geom4_compact(B) if i == 4 else conv_geom(i, B)
            splits = A.conv_wgrad_splits(g)
            n = 2048 if i == 1 else CONV_CH[i] * CONV_CH[i - 1] * 16
            part = ws.partial(f"cw{i}", splits * n)
            A.conv_wgrad(g, dA[i], ws.A[i - 1], part, splits)
            fused_bias = i == 1 and B_bias > 0
            A.unprep_conv_wgrad(part, splits, self.flat.g(self.wname(i)), CONV_CH[i], CONV_CH[i - 1], i == 1,
                                self.flat.g(self.bname(i)) if fused_bias else None)
            # conv2-4: the bias gradients were summed by the dgrad epilogues that produced delta_2..delta_4
            # (backward_data / the FC1 dgrad of the owning engine) when fused_dbias is on; else one column-sum pass each
            if B_bias > 0 and not fused_bias and not self.fused_dbias:
                rows = B_bias * (96 * 96, 46 * 46, 22 * 22, 100)[i - 1]
                A.colsum(dA[i], CONV_CH[i], rows, CONV_CH[i], self.flat.g(self.bname(i)))


def linear_fwd(ws: Workspace, key: str, x, ldx, w, ldw, bias, y, ldy, M, N, K, epilogue) -> None:
    """nn.Linear forward with split-K when the tile grid would leave most SMs idle."""
    bn = min(256, (N + 15) // 16 * 16)
    splits = _splits_for((M + 127) // 128, (N + bn - 1) // bn, (K + 31) // 32)
    if splits == 1:
        A.linear_fwd(x, ldx, w, ldw, bias, y, ldy, M, N, K, epilogue, SLOPE, 1)
        return
    ldp = (N + 3) // 4 * 4
    part = ws.partial(key, splits * M * ldp)
    A.linear_fwd(x, ldx, w, ldw, None, part, ldp, M, N, K, EPI_STORE, SLOPE, splits)
    A.splitk_reduce(part, splits, M, N, ldp, bias, None, 0, y, ldy, epilogue, SLOPE)


def linear_wgrad(ws: Workspace, key: str, dy, lddy, x, ldx, dw, lddw, M, N, K) -> None:
    bn = min(256, (N + 31) // 32 * 32)
    splits = _splits_for((M + 127) // 128, (N + bn - 1) // bn, (K + 31) // 32)
    if splits == 1:
        A.linear_wgrad(dy, lddy, x, ldx, dw, lddw, M, N, K, 1)
        return
    part = ws.partial(key, splits * M * lddw)
    A.linear_wgrad(dy, lddy, x, ldx, part, lddw, M, N, K, splits)
    A.splitk_reduce(part, splits, M, N, lddw, None, None, 0, dw, lddw, EPI_STORE, SLOPE)


