"""Drop-in for tools/model.py ``Policy`` (act / get_value / evaluate_actions) on B200.

The module tree only exists to own parameters under the reference's ``state_dict`` names
(``base.obs_processor.main.{0,2,4,6}``, ``base.metrics_processor.road_option_embedding``, ``base.body.body.{0,2,4}``,
``base.head.head.{0,2}``, tools/model.py:56-128) and to draw the same default initialisation in the same order; the
arithmetic runs in the C-ABI kernels through :class:`PolicyEngine`.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn as nn

from . import _abi as A
from . import engine as E
from ._abi import LDF, EPI_BIAS_LRELU
from .graphs import StepGraph

N_METRIC_FEAT = 13


def _conv_seq() -> nn.Sequential:
    layers = []
    for i in range(4):
        layers += [nn.Conv2d(E.CONV_CH[i], E.CONV_CH[i + 1], 4, stride=2), nn.LeakyReLU(E.SLOPE)]
    return nn.Sequential(*layers)


class _Holder(nn.Module):
    """Parameter container; attribute names give the reference's state_dict keys."""

    def __init__(self, **children):
        super().__init__()
        for k, v in children.items():
            setattr(self, k, v)


def _policy_tree() -> nn.Module:
    obs = _Holder(main=_conv_seq())
    met = _Holder(road_option_embedding=nn.Embedding(10, 8))
    body = _Holder(body=nn.Sequential(nn.Linear(E.FEAT + N_METRIC_FEAT, 512), nn.LeakyReLU(E.SLOPE), nn.Linear(512, 512),
                                      nn.LeakyReLU(E.SLOPE), nn.Linear(512, 512), nn.LeakyReLU(E.SLOPE)))
    head = _Holder(head=nn.Sequential(nn.Linear(512, 256), nn.LeakyReLU(E.SLOPE), nn.Linear(256, 3)))
    return _Holder(obs_processor=obs, metrics_processor=met, body=body, head=head)


HIDDEN = (("base.body.body.2", 512, 512), ("base.body.body.4", 512, 512), ("base.head.head.0", 512, 256))


class PolicyEngine:
    """Forward / backward of CNNBase (tools/model.py:56-86) over a batch resident in the workspace."""

    def __init__(self, module: nn.Module):
        self.flat = E.FlatParams(module)
        self.conv = E.ConvStack(self.flat, "base.obs_processor.", need_input_grad=False)
        self.ws: Optional[E.Workspace] = None
        self.w1: Optional[torch.Tensor] = None
        self.dirty = True

    # ---- bookkeeping -------------------------------------------------------------------------------------
    def sync_params(self) -> None:
        """Refresh flat buffers / operand copies if parameters moved or changed."""
        moved = self.flat.ensure()
        dev = self.flat.flat.device
        if dev.type != "cuda" and not getattr(A, "EMULATED", False):
            raise RuntimeError("gail_carla_b200.Policy runs on CUDA only: call .to('cuda') first (no CPU fallback)")
        ver = self.flat.version()
        if ver != getattr(self, "_version", None):     # torch code wrote a parameter in place since the last preparation
            self.dirty, self._version = True, ver
        if moved or self.dirty or self.w1 is None or self.w1.device != dev:
            if self.w1 is None or self.w1.device != dev:
                self.w1 = torch.zeros(512, LDF, dtype=torch.float32, device=dev)
                self.dw1 = None
                self.ws = None
            self.conv.prepare()
            A.prep_fc1_weight(self.flat.p("base.body.body.0.weight"), self.w1, 512, N_METRIC_FEAT, LDF)
            self.dirty = False

    def workspace(self, rows: int) -> E.Workspace:
        self.ws = E.shared_workspace(self.flat.flat.device, rows, with_input_grad=False)
        return self.ws

    # ---- forward -----------------------------------------------------------------------------------------
    def load_inputs(self, obs_rows, metrics_rows, idx, B: int, row0: int = 0) -> None:
        """Gather + normalise images into the space-to-depth layout and stage the raw metrics.
        obs_rows [R,3,192,192], metrics_rows [R,4]; idx int64 [B] (None: rows 0..B-1)."""
        ws = self.ws
        A.gather_obs_s2d(obs_rows, idx, ws.X0[row0:], B)
        m = ws.buf("metrics", ws.rows, 4)
        A.gather_rows(metrics_rows, idx, m[row0:], B, 4, 4)

    def forward(self, B: int, training: bool = False) -> torch.Tensor:
        """Rows [0,B) of the workspace -> head output [B,4] = {value, mu0_raw, mu1_raw, 0}."""
        ws, P = self.ws, self.flat.p
        self.conv.forward(ws, B, training=training)
        A.metrics_features(ws.buf("metrics", ws.rows, 4), P("base.metrics_processor.road_option_embedding.weight"),
                           ws.F[:, E.FEAT:], LDF, 32, B)
        h1 = ws.buf("h1", ws.rows, 512)
        E.linear_fwd(ws, "fc1p", ws.F, LDF, self.w1, LDF, P("base.body.body.0.bias"), h1, 512, B, 512, LDF, EPI_BIAS_LRELU)
        x = h1
        for i, (name, fin, fout) in enumerate(HIDDEN):
            y = ws.buf(f"h{i + 2}", ws.rows, fout)
            E.linear_fwd(ws, f"fc{i + 2}p", x, fin, P(name + ".weight"), fin, P(name + ".bias"), y, fout, B, fout, fin,
                         EPI_BIAS_LRELU)
            x = y
        out = ws.buf("head", ws.rows, 4)
        A.small_linear_fwd(x, 256, P("base.head.head.2.weight"), P("base.head.head.2.bias"), out, 4, B, 3, 256)
        return out

    # ---- backward ----------------------------------------------------------------------------------------
    EARLY_BUCKET = "base.metrics_processor.road_option_embedding.weight"   # first parameter of the early all-reduce bucket

    def zero_contribution(self, reducer=None) -> None:
        """This rank holds no row of the current global minibatch (exact sharding): its gradient is zero, but it must issue
        the SAME sequence of collectives as the ranks that ran a backward pass - same buckets, same order."""
        self.flat.begin_backward()
        if reducer is not None:
            reducer.ready(self.flat, *self.flat.span(self.EARLY_BUCKET))

    def backward(self, B: int, d_head: torch.Tensor, reducer=None) -> None:
        """d loss / d head_out [B,4] -> gradients of every parameter (written into the flat grad buffer).
        `reducer` (optim.GradReducer): told when a slice of the gradient buffer is final, so the multi-GPU all-reduce of
        the fully connected layers (90 % of the bytes) overlaps the convolution gradients."""
        ws, P, G = self.ws, self.flat.p, self.flat.g
        self.flat.begin_backward()
        dA = ws.grads()
        hs = [ws.buf("h1", ws.rows, 512)] + [ws.buf(f"h{i + 2}", ws.rows, f[2]) for i, f in enumerate(HIDDEN)]
        # head.2 (256 -> 3): SIMT; dx masked by LeakyReLU'(h4)
        d = ws.buf("dh4", ws.rows, 256)
        A.small_linear_bwd(hs[3], 256, P("base.head.head.2.weight"), d_head, 4, d, 256, G("base.head.head.2.weight"),
                           G("base.head.head.2.bias"), B, B, 3, 256, E.SLOPE)
        # hidden layers, last to first: dW = d^T x, db = colsum(d), dx = LeakyReLU'(x) * d W
        for i in (2, 1, 0):
            name, fin, fout = HIDDEN[i]
            x = hs[i]
            E.linear_wgrad(ws, f"w{i}", d, fout, x, fin, G(name + ".weight"), fin, fout, fin, B)
            A.colsum(d, fout, B, fout, G(name + ".bias"))
            dx = ws.buf(f"dh{i + 1}", ws.rows, fin)
            A.linear_dgrad(d, fout, P(name + ".weight"), fin, dx, fin, B, fin, fout, mask_src=x, ldm=fin, slope=E.SLOPE)
            d = dx
        # body.0 (25613 -> 512) on the permuted / padded operand copy
        if getattr(self, "dw1", None) is None or self.dw1.device != self.w1.device:
            self.dw1 = torch.zeros(512, LDF, dtype=torch.float32, device=self.w1.device)
        E.linear_wgrad(ws, "w1", d, 512, ws.F, LDF, self.dw1, LDF, 512, LDF, B)
        A.unprep_fc1_wgrad(self.dw1, 1, G("base.body.body.0.weight"), 512, N_METRIC_FEAT, LDF)
        A.colsum(d, 512, B, 512, G("base.body.body.0.bias"))
        # features: conv part masked by LeakyReLU'(a4) (-> delta_4); metric part feeds the embedding
        fb = self.conv.fused_dbias           # conv4's bias gradient = per-channel sums of delta_4, from this dgrad's epilogue
        A.linear_dgrad(d, 512, self.w1, LDF, dA[4], E.FEAT, B, E.FEAT, 512, mask_src=ws.F, ldm=LDF, slope=E.SLOPE,
                       mask_bits=ws.mbits[4], colsum=G(self.conv.bname(4)) if fb else None, colsum_mod=256, colsum_rows=B)
        A.linear_dgrad(d, 512, self.w1[:, E.FEAT:], LDF, ws.dFt, 32, B, 32, 512)
        A.metrics_features_bwd(ws.buf("metrics", ws.rows, 4), ws.dFt, 32, G("base.metrics_processor.road_option_embedding.weight"), B)
        if reducer is not None:     # embedding + every Linear are final; the convolutions come first in the flat order
            reducer.ready(self.flat, *self.flat.span(self.EARLY_BUCKET))
        self.conv.backward_data(ws, B, B_bias=B if fb else 0)
        self.conv.backward_params(ws, B, B)


class CNNBase(nn.Module):
    """Holds the ``base.*`` parameters (tools/model.py:56-69)."""

    def __init__(self, activation: bool, logstd: Sequence[float]):
        super().__init__()
        tree = _policy_tree()
        self.obs_processor = tree.obs_processor
        self.metrics_processor = tree.metrics_processor
        self.body = tree.body
        self.head = tree.head
        self.logstd = torch.tensor([float(v) for v in logstd])   # plain tensor, not in state_dict (tools/model.py:67)
        self.activation = bool(activation)


class Policy(nn.Module):
    """tools/model.py:15-53 - same constructor, same methods, same return shapes; CUDA only."""

    def __init__(self, obs_shape, metrics_space, action_space, activation, logstd, multi_head=False):
        super().__init__()
        if tuple(obs_shape) != (3, 192, 192) or metrics_space.shape[0] != 4 or action_space.shape[0] != 2:
            raise ValueError("Policy supports obs (3,192,192), metrics (4,), action (2,) - the CARLA shapes of the reference")
        self.base = CNNBase(activation, logstd)
        self.max = torch.Tensor([1, 1])     # tools/model.py:22-23 (unused by the reference as well)
        self.min = torch.Tensor([-1, 0])
        self._engine: Optional[PolicyEngine] = None
        self._act_graph = StepGraph("act")
        self._act_state = None

    # the reference's drivers move the module around; parameters are re-flattened lazily
    @property
    def engine(self) -> PolicyEngine:
        if self._engine is None:
            self._engine = PolicyEngine(self)
        return self._engine

    def mark_params_changed(self) -> None:
        if self._engine is not None:
            self._engine.dirty = True

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self.mark_params_changed()
        return r

    def _run(self, obs, metrics, training: bool = False) -> tuple:
        eng = self.engine
        eng.sync_params()
        dev = eng.flat.flat.device
        if obs.dtype != torch.uint8:
            obs = obs.to(dev, torch.float32)
        obs = obs.to(dev).contiguous()
        metrics = metrics.to(dev, torch.float32).contiguous()
        B = obs.shape[0]
        eng.workspace(B)
        eng.load_inputs(obs, metrics, None, B)
        return eng, eng.forward(B, training=training), B

    def act(self, obs, metrics, deterministic=False):
        """tools/model.py:25-36 -> (value [B,1], action [B,2], action_log_probs [B,1]).

        This is the per-env-step call of the rollout loop (tools/learn.py:111-133): a batch of N envs, ~25 launches that
        are each shorter than their launch overhead.  The inputs are copied into static device buffers and the whole
        step (gather + normalise, trunk forward, sampling head) is replayed as one CUDA graph (graphs.StepGraph)."""
        with torch.no_grad():
            eng = self.engine
            eng.sync_params()
            dev = eng.flat.flat.device
            B = int(obs.shape[0])
            u8 = obs.dtype == torch.uint8
            # static staging buffers, one set per (batch size, dtype, device) and kept for the module's lifetime: a captured
            # graph holds their addresses, so they must never be re-created behind it
            if self._act_state is None:
                self._act_state = {}
            st = self._act_state.get((B, u8, str(dev)))
            if st is None:
                st = dict(obs=torch.zeros(B, 3, 192, 192, dtype=torch.uint8 if u8 else torch.float32, device=dev),
                          met=torch.zeros(B, 4, device=dev), noise=torch.zeros(B, 2, device=dev), value=torch.empty(B, 1, device=dev),
                          action=torch.empty(B, 2, device=dev), logp=torch.empty(B, 1, device=dev))
                self._act_state[(B, u8, str(dev))] = st
            src = obs.as_subclass(torch.Tensor) if u8 else obs
            st["obs"].copy_(src, non_blocking=True)
            st["met"].copy_(metrics, non_blocking=True)
            # Normal.sample() of the reference = mean + std * N(0,1) drawn from the default generator; the draw stays on
            # the host generator (B x 2 floats) so a GPU run consumes the same random stream as the CPU oracle
            if not deterministic:
                st["noise"].copy_(torch.randn(B, 2), non_blocking=True)
            logstd, activation = self.base.logstd.tolist(), self.base.activation

            def device_step():
                eng.workspace(B)
                eng.load_inputs(st["obs"], st["met"], None, B)
                head = eng.forward(B)
                A.policy_act(head, None if deterministic else st["noise"], st["value"], st["action"], st["logp"], B, logstd, activation)

            ws = eng.workspace(B)
            key = (ws.X0.data_ptr(), ws.rows, B, u8, bool(deterministic), eng.flat.flat.data_ptr(), tuple(logstd), activation) + \
                tuple(st[k].data_ptr() for k in ("obs", "met", "noise", "value", "action", "logp"))
            self._act_graph.run(key, device_step, dev)
            return st["value"].clone(), st["action"].clone(), st["logp"].clone()

    def get_value(self, obs, metrics):
        """tools/model.py:41-43."""
        with torch.no_grad():
            eng, head, B = self._run(obs, metrics)
            return head[:B, 0:1].clone()

    def evaluate_actions(self, obs, metrics, action):
        """tools/model.py:45-53 -> (value, log-probs, entropy, steer log-std, throttle log-std).

        ``value`` and ``log-probs`` are differentiable w.r.t. the parameters (the reference's learn_bc.py:37-45 calls
        ``.backward()`` on a loss built from them): under ``torch.enable_grad()`` they carry a grad_fn whose backward
        runs the hand-derived trunk backward of the engine and accumulates into ``p.grad``.  The activations live in
        the shared workspace, so ``backward()`` must run before the next forward of this policy (it raises
        otherwise).  PPO.update does not come through here - it uses the fused loss forward+backward kernel."""
        ls = self.base.logstd
        entropy = (0.5 + 0.5 * torch.log(torch.tensor(2 * torch.pi)) + ls).sum()
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            value, logp = _EvaluateActions.apply(self, obs, metrics, action, *self.parameters())
        else:
            with torch.no_grad():
                value, logp = _evaluate_forward(self, obs, metrics, action, training=False)
        return value, logp, entropy.to(value.device), ls[0].clone(), ls[1].clone()


def _evaluate_forward(policy: Policy, obs, metrics, action, training: bool):
    eng, head, B = policy._run(obs, metrics, training=training)
    dev = head.device
    value = torch.empty(B, 1, device=dev); logp = torch.empty(B, 1, device=dev)
    A.ppo_loss(head, action.to(dev, torch.float32).contiguous(), None, None, None, None, None, None, value, logp, None, B,
               policy.base.logstd.tolist(), policy.base.activation, 0.0, 0.0, 0.0, 2)
    return value, logp


class _EvaluateActions(torch.autograd.Function):
    """Autograd bridge of ``Policy.evaluate_actions``: forward = engine forward (training mode: LeakyReLU' bit masks are
    kept), backward = fused head-tail derivative (gc_ppo_loss_fwd_bwd, BC mode with unit weight gives d logp / d head)
    + the engine's hand-derived trunk backward.  The parameters are passed as inputs only so that autograd routes their
    gradients; the gradients are accumulated like autograd would (``p.grad += ...``)."""

    @staticmethod
    def forward(ctx, policy, obs, metrics, action, *params):
        value, logp = _evaluate_forward(policy, obs, metrics, action, training=True)
        eng = policy.engine
        eng.forward_serial = getattr(eng, "forward_serial", 0) + 1
        ctx.policy, ctx.serial, ctx.B = policy, eng.forward_serial, value.shape[0]
        ctx.action = action.to(value.device, torch.float32).contiguous()
        return value, logp

    @staticmethod
    def backward(ctx, g_value, g_logp):
        policy, B = ctx.policy, ctx.B
        eng = policy.engine
        if eng.forward_serial != ctx.serial or eng.ws is None or eng.ws.rows < B:
            raise RuntimeError("evaluate_actions: backward() must run before the next forward pass of this Policy "
                               "(activations live in the shared workspace)")
        ws = eng.ws
        with torch.no_grad():
            head = ws.buf("head", ws.rows, 4)
            d_head = ws.buf("dhead", ws.rows, 4)
            # BC mode with weight -1 and norm 1: d_head[b] = d logp_b / d head_b (column 0 = 0)
            A.ppo_loss(head, ctx.action, None, None, None, None, None, d_head, None, None, None, B,
                       policy.base.logstd.tolist(), policy.base.activation, 0.0, 0.0, -1.0, 1, norm=1)
            gl = torch.zeros(B, 1, device=head.device) if g_logp is None else g_logp.reshape(B, 1).to(head.device, torch.float32)
            d_head[:B].mul_(gl)
            if g_value is not None:
                d_head[:B, 0:1].copy_(g_value.reshape(B, 1))
            flat = eng.flat
            flat.restore_grads()          # (a torch optimiser's zero_grad() sets p.grad to None)
            keep = None if flat.grad_clean else flat.grad.clone()
            eng.backward(B, d_head)
            grads = tuple(p.grad.clone() if p.requires_grad else None for p in policy.parameters())
            if keep is None:
                flat.grad.zero_(); flat.grad_clean = True
            else:
                flat.grad.copy_(keep)
        for p, g in zip(policy.parameters(), grads):      # accumulate like autograd (p.grad is a view of the flat buffer)
            if g is not None:
                p.grad.add_(g)
        flat.grad_clean = False
        eng.forward_serial += 1
        return (None, None, None, None) + tuple(None for _ in grads)
