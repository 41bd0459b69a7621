"""Drop-in for algo/wdgail.py ``Discriminator`` (tanh-Wasserstein critic with input-gradient penalty) on B200.

Same constructor, ``update`` (7-tuple), ``compute_loss`` (3-tuple), ``predict_reward`` (CPU ``[N,1]`` tensor),
``forward`` and ``compute_grad_pen``; state_dict keys ``obs_processor.main.{0,2,4,6}``,
``metrics_processor.road_option_embedding``, ``trunk.{0,2}`` (algo/wdgail.py:19-38).

``update`` runs expert, policy and mixed-up samples as ONE batch of 3B rows through the critic and never builds an
autograd graph.  The penalty's double backward (algo/wdgail.py:85-97, ``create_graph=True``) is hand-derived: every
non-linearity is piecewise linear, so with the LeakyReLU slope masks D_k frozen
``g = dD/dx = C1^T D1 C2^T D2 C3^T D3 C4^T D4 W1a^T D5 w2^T`` is linear in each weight, and with ``u = d gp / d g``
  d gp / d C_k = wgrad(delta_k, v_{k-1}),  d gp / d W1a = delta_5 (x) v_4,  d gp / d w2 = sum_b D5 W1a v_4,
where delta_k is the ordinary dgrad chain seeded with 1 and ``v_k = D_k C_k v_{k-1}`` (v_0 = u) is a forward pass
without biases.  Biases, the embedding and the metric/action columns of trunk.0 get exactly zero penalty gradient.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from . import _abi as A
from . import engine as E
from ._abi import LDF, S2D_PER_SAMPLE, EPI_BIAS_LRELU, EPI_MASK, EPI_STORE
from .model import _Holder, _conv_seq, N_METRIC_FEAT
from .expert import DeviceBatch, expert_rows
from .graphs import StepGraph
from .optim import FusedClipAdam, world_size
from .running_mean_std import RunningMeanStd

TAIL = N_METRIC_FEAT + 2      # metric features + action columns of trunk.0 (algo/wdgail.py:26-29)
LDH = 128                     # row pitch of the hidden activations [rows, hidden_dim<=128]


class CriticEngine:
    def __init__(self, module: nn.Module, hidden: int):
        if hidden > LDH or hidden % 4:
            raise ValueError("hidden_dim must be a multiple of 4 and <= 128")
        self.hidden = hidden
        self.flat = E.FlatParams(module)
        self.conv = E.ConvStack(self.flat, "obs_processor.", need_input_grad=True)
        self.ws: Optional[E.Workspace] = None
        self.w1 = None
        self.dirty = True

    def sync_params(self) -> None:
        moved = self.flat.ensure()
        dev = self.flat.flat.device
        if dev.type != "cuda" and not getattr(A, "EMULATED", False):
            raise RuntimeError("gail_carla_b200.Discriminator runs on CUDA only: move it to a CUDA device (no CPU fallback)")
        ver = self.flat.version()
        if ver != getattr(self, "_version", None):     # torch code wrote a parameter in place since the last preparation
            self.dirty, self._version = True, ver
        if moved or self.dirty or self.w1 is None or self.w1.device != dev:
            if self.w1 is None or self.w1.device != dev:
                self.w1 = torch.zeros(self.hidden, LDF, dtype=torch.float32, device=dev)
                self.dw1 = torch.zeros(self.hidden, LDF, dtype=torch.float32, device=dev)
                self.ws = None
            self.conv.prepare()
            A.prep_fc1_weight(self.flat.p("trunk.0.weight"), self.w1, self.hidden, TAIL, LDF)
            self.dirty = False

    def workspace(self, rows: int) -> E.Workspace:
        self.ws = E.shared_workspace(self.flat.flat.device, rows, with_input_grad=True)
        return self.ws

    # rows [row0,row0+B) <- images / metrics / actions given as dense rows or gathered by idx
    def load_inputs(self, obs_rows, met_rows, act_rows, idx, B: int, row0: int) -> None:
        ws = self.ws
        A.gather_obs_s2d(obs_rows, idx, ws.X0[row0:], B)
        A.gather_rows(met_rows, idx, ws.buf("metrics", ws.rows, 4)[row0:], B, 4, 4)
        A.gather_rows(act_rows, idx, ws.buf("actions", ws.rows, 2)[row0:], B, 2, 2)

    def load_pair_mix(self, e_obs, e_met, e_act, e_idx, p_obs, p_met, p_act, p_idx, alpha, B: int) -> bool:
        """Expert rows -> [0,B), policy rows -> [B,2B) and their mix-up -> [2B,3B) of the workspace.  With both image
        sources uint8 this is ONE pass (gc_gather_pair_mix_u8_s2d) and returns True: update_step must then skip its own
        mix-up launch.  Otherwise the two plain gathers run and it returns False."""
        ws = self.ws
        fused = e_obs.dtype == torch.uint8 and p_obs.dtype == torch.uint8
        if fused:
            A.gather_pair_mix(e_obs, e_idx, p_obs, p_idx, alpha, ws.X0, B)
        else:
            A.gather_obs_s2d(e_obs, e_idx, ws.X0, B)
            A.gather_obs_s2d(p_obs, p_idx, ws.X0[B:], B)
        m, a = ws.buf("metrics", ws.rows, 4), ws.buf("actions", ws.rows, 2)
        A.gather_rows(e_met, e_idx, m, B, 4, 4); A.gather_rows(e_act, e_idx, a, B, 2, 2)
        A.gather_rows(p_met, p_idx, m[B:], B, 4, 4); A.gather_rows(p_act, p_idx, a[B:], B, 2, 2)
        return fused

    def tail_features(self, B: int, row0: int, mix_from: Optional[tuple] = None) -> None:
        """ProcessMetrics + action passthrough into F[:, 25600:]; ``mix_from=(row_e,row_p,alpha)`` mixes the raw inputs."""
        ws = self.ws
        m, a = ws.buf("metrics", ws.rows, 4), ws.buf("actions", ws.rows, 2)
        emb = self.flat.p("metrics_processor.road_option_embedding.weight")
        out = ws.F[row0:, E.FEAT:]
        if mix_from is None:
            A.metrics_features(m[row0:], emb, out, LDF, 32, B, action=a[row0:])
        else:
            re, rp, alpha = mix_from
            A.metrics_features(m[re:], emb, out, LDF, 32, B, action=a[re:], metrics2=m[rp:], action2=a[rp:], alpha=alpha)

    def trunk_forward(self, rows: int) -> torch.Tensor:
        """F[0:rows] -> hidden H (LeakyReLU) -> critic output d [rows]."""
        ws, P = self.ws, self.flat.p
        H = ws.buf("H", ws.rows, LDH)
        E.linear_fwd(ws, "fc1d", ws.F, LDF, self.w1, LDF, P("trunk.0.bias"), H, LDH, rows, self.hidden, LDF, EPI_BIAS_LRELU)
        d = ws.buf("d", ws.rows)
        A.small_linear_fwd(H, LDH, P("trunk.2.weight"), P("trunk.2.bias"), d, 1, rows, 1, self.hidden)
        return d

    def forward(self, rows: int, training: bool = False) -> torch.Tensor:
        self.conv.forward(self.ws, rows, training=training)
        return self.trunk_forward(rows)

    EARLY_BUCKET = "metrics_processor.road_option_embedding.weight"   # first parameter of the early all-reduce bucket

    def zero_contribution(self, reducer=None) -> None:
        """No row of the current global minibatch on this rank (exact sharding): zero gradient, but the same sequence of
        collectives as the ranks that ran update_step (same buckets, same order)."""
        self.flat.begin_backward()
        if reducer is not None:
            reducer.ready(self.flat, *self.flat.span(self.EARLY_BUCKET))

    def update_step(self, B: int, alpha: torch.Tensor, acc: torch.Tensor, lambda_: float = 10.0, norm=None, reducer=None,
                    premixed: bool = False) -> None:
        """Rows [0,B) expert, [B,2B) policy already loaded.  Builds the mix-up rows, runs forward + the full
        backward (Wasserstein part + gradient penalty) and leaves the gradients in the flat buffer.
        `norm`: rows the batch means run over (default B; the global minibatch size when B is one rank's share of it).
        `reducer` (optim.GradReducer) is told when slices of the gradient buffer are final (multi-GPU overlap)."""
        ws, P, G, H_ = self.ws, self.flat.p, self.flat.g, self.hidden
        R = 3 * B
        # ---- forward over expert | policy | mix-up (algo/wdgail.py:116,121,66-82)
        if not premixed:       # (load_pair_mix already wrote the mix-up rows)
            A.mixup(ws.X0, ws.X0[B:], alpha, ws.X0[2 * B:], B, S2D_PER_SAMPLE)
        self.tail_features(B, 0); self.tail_features(B, B); self.tail_features(B, 2 * B, mix_from=(0, B, alpha))
        d = self.forward(R, training=True)
        dd = ws.buf("dd", ws.rows)
        A.disc_loss_seed(d, dd, acc, B, norm)                             # acc[0:4]; seeds -/+tanh'/B and 1
        # ---- backward
        self.flat.begin_backward()
        dA = ws.grads()
        H = ws.buf("H", ws.rows, LDH)
        dH = ws.buf("dH", ws.rows, LDH)
        A.small_linear_bwd(H, LDH, P("trunk.2.weight"), dd, 1, dH, LDH, G("trunk.2.weight"), G("trunk.2.bias"), R, 2 * B, 1,
                           H_, E.SLOPE)
        # delta_4 for all rows; metric/action columns only for the rows that carry loss (expert, policy)
        fb = self.conv.fused_dbias     # conv2-4 bias gradients (expert + policy rows only) from the dgrad epilogues
        A.linear_dgrad(dH, LDH, self.w1, LDF, dA[4], E.FEAT, R, E.FEAT, H_, mask_src=ws.F, ldm=LDF, slope=E.SLOPE,
                       mask_bits=ws.mbits[4], colsum=G(self.conv.bname(4)) if fb else None, colsum_mod=256, colsum_rows=2 * B)
        A.linear_dgrad(dH, LDH, self.w1[:, E.FEAT:], LDF, ws.dFt, 32, 2 * B, 32, H_)
        m = ws.buf("metrics", ws.rows, 4)
        emb_g = G("metrics_processor.road_option_embedding.weight")
        A.metrics_features_bwd(m, ws.dFt, 32, emb_g, 2 * B)
        self.conv.backward_data(ws, R, B_bias=2 * B if fb else 0)
        # ---- gradient penalty: g = dD/dx on the mix-up rows, u = d gp/d g, second-order forward chain in place
        self.conv.input_grad(ws, B, 2 * B)
        A.grad_penalty(dA[0][2 * B:], ws.X0[2 * B:], acc[4:], B, S2D_PER_SAMPLE, lambda_, E.INV_STD, norm)
        self.conv.forward_masked(ws, B, 2 * B)
        A.zero_block(ws.F[2 * B:, E.FEAT:], rows=B, width=LDF - E.FEAT, pitch=LDF)   # penalty gradient of the tail columns is 0
        t = ws.buf("v5", ws.rows, LDH)
        E.linear_fwd(ws, "fc1v", ws.F[2 * B:], LDF, self.w1, LDF, None, t, LDH, B, H_, LDF, EPI_STORE)
        A.splitk_reduce(t, 1, B, H_, LDH, None, H[2 * B:], LDH, t, LDH, EPI_MASK, E.SLOPE)
        A.colsum(t, LDH, B, H_, G("trunk.2.weight"))
        # ---- weight gradients over all 3B rows ((delta, a) pairs for expert/policy, (delta_hat, v) for the penalty)
        E.linear_wgrad(ws, "w1d", dH, LDH, ws.F, LDF, self.dw1, LDF, H_, LDF, R)
        A.unprep_fc1_wgrad(self.dw1, 1, G("trunk.0.weight"), H_, TAIL, LDF)
        A.colsum(dH, LDH, 2 * B, H_, G("trunk.0.bias"))
        if reducer is not None:      # embedding + trunk are final; only the convolution weight gradients remain
            reducer.ready(self.flat, *self.flat.span(self.EARLY_BUCKET))
        self.conv.backward_params(ws, R, 2 * B)


class Discriminator(nn.Module):
    # device staging sets of the expert-batch prefetch ring (prefetch depth N_STAGING - 1).  Three sets were measured
    # against two (profiles/r02_bench_c4_n1_round1_binary.json: e2e 1581 vs 1600 ms/step, i.e. nothing) and cost
    # another 1.8 GB of HBM per set at B=4096 fp32, so two it is.
    N_STAGING = 2
    def __init__(self, state_shape, metrics_space, action_space, hidden_dim, device, lr, eps, betas, max_grad_norm=None):
        super(Discriminator, self).__init__()
        if tuple(state_shape) != (3, 192, 192) or metrics_space.shape[0] != 4 or action_space.shape[0] != 2:
            raise ValueError("Discriminator supports state (3,192,192), metrics (4,), action (2,)")
        self.device = device
        self.obs_processor = _Holder(main=_conv_seq())
        self.metrics_processor = _Holder(road_option_embedding=nn.Embedding(10, 8))
        self.trunk = nn.Sequential(nn.Linear(E.FEAT + TAIL, hidden_dim), nn.LeakyReLU(E.SLOPE), nn.Linear(hidden_dim, 1))
        self.hidden_dim = hidden_dim
        self.max_grad_norm = max_grad_norm
        self._engine: Optional[CriticEngine] = None
        self.optimizer = FusedClipAdam(lambda: self.engine.flat, self.parameters(), lr, eps, betas, max_grad_norm)
        self.returns = None
        self.ret_rms = RunningMeanStd(shape=())     # constructed and never used, as in algo/wdgail.py:37-38
        self.exact_sharding = False                 # multi-GPU exact mode, see PPO.exact_sharding
        self._graph = StepGraph("critic")
        self._alpha_buf = None

    @property
    def engine(self) -> CriticEngine:
        if self._engine is None:
            self._engine = CriticEngine(self, self.hidden_dim)
        return self._engine

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        if self._engine is not None:
            self._engine.dirty = True
        return r

    def _dev(self):
        return self.engine.flat.flat.device

    def _to_dev(self, *ts):
        """fp32 on the device; uint8 observations (bytes standing for b/255, see storage.ByteObs) stay uint8."""
        dev = self._dev()
        out = []
        for t in ts:
            if t.dtype == torch.uint8:
                out.append(t.as_subclass(torch.Tensor).to(dev, non_blocking=True).contiguous())
            else:
                out.append(t.to(dev, torch.float32, non_blocking=True).contiguous())
        return out

    # ---- algo/wdgail.py:40-54
    def forward(self, state, metrics, action, gp=False):
        """gp=False: critic output [B,1] (no autograd graph).  gp=True (algo/wdgail.py:51-52): ``(output,
        state_transformed, metrics_transformed, action_transformed)`` where ``state_transformed`` is a leaf copy of the
        raw input image with ``requires_grad`` and ``output`` is differentiable w.r.t. it to first order
        (``autograd.grad(output, state_transformed, ones)`` = the dD/dx of algo/wdgail.py:85-91, computed by the engine's
        dgrad chain).  The second-order graph the reference builds with ``create_graph=True`` does not exist here -
        ``compute_grad_pen`` / ``update`` use the hand-derived second-order pass instead."""
        if gp:
            eng = self.engine
            eng.sync_params()
            state_d, metrics_d, action_d = self._to_dev(state, metrics, action)
            if state_d.dtype == torch.uint8:
                state_d = state_d.float().div_(255.0)
            state_t = state_d.clone().requires_grad_(True)
            out = _CriticOfImage.apply(self, state_t, metrics_d, action_d)
            B = state_t.shape[0]
            with torch.no_grad():
                feats = eng.ws.F[:B, E.FEAT:E.FEAT + N_METRIC_FEAT].clone()
            return out, state_t, feats, action_d.clone()
        with torch.no_grad():
            eng = self.engine
            eng.sync_params()
            state, metrics, action = self._to_dev(state, metrics, action)
            B = state.shape[0]
            eng.workspace(B)
            eng.load_inputs(state, metrics, action, None, B, 0)
            eng.tail_features(B, 0)
            return eng.forward(B)[:B].clone().view(B, 1)

    # ---- algo/wdgail.py:56-98 (value only; alpha drawn from the CPU default generator like the reference)
    def compute_grad_pen(self, expert_state, expert_metrics, expert_action, policy_state, policy_metrics, policy_action,
                         lambda_=10):
        with torch.no_grad():
            eng = self.engine
            eng.sync_params()
            B = expert_state.shape[0]
            alpha = torch.rand(B, 1, 1, 1).view(B).to(self._dev())
            eng.workspace(3 * B)
            eng.load_inputs(*self._to_dev(expert_state, expert_metrics, expert_action), None, B, 0)
            eng.load_inputs(*self._to_dev(policy_state, policy_metrics, policy_action), None, B, B)
            acc = torch.zeros(8, dtype=torch.float64, device=self._dev())
            eng.update_step(B, alpha, acc, float(lambda_))
            return (float(lambda_) * acc[4] / B).float()

    def _upload_alpha(self, alpha: torch.Tensor) -> torch.Tensor:
        """Mix-up coefficients drawn on the host (RNG parity) -> device through a small ring of pinned buffers."""
        dev = self._dev()
        if dev.type != "cuda":
            return alpha
        ring = getattr(self, "_alpha_ring", None)
        if ring is None or ring[0][0].numel() != alpha.numel():
            ring = [[torch.empty(alpha.numel(), pin_memory=True), None] for _ in range(4)]
            self._alpha_ring, self._alpha_i = ring, 0
        slot = ring[self._alpha_i]
        self._alpha_i = (self._alpha_i + 1) % len(ring)
        if slot[1] is not None:
            slot[1].synchronize()          # the upload that last used this pinned slot has completed (normally long ago)
        slot[0].copy_(alpha)
        out = slot[0].to(dev, non_blocking=True)
        slot[1] = torch.cuda.Event()
        slot[1].record()
        return out

    def _prefetched(self, pairs):
        """Yield ``(expert tensors on the device, idx, release)`` for every (expert batch, policy index batch) pair.

        The expert loader hands out host tensors (algo/wdgail.py:112,119; 1.8 GB per batch at B=4096).  They are uploaded
        on a side stream into a ring of three preallocated device staging sets, so the copies of batches i+1 and i+2 run
        while batch i is being processed (two batches of slack: other host->device traffic - a rollout streaming in for
        the next iteration - shares the DMA queue and would otherwise stall a minibatch every time it gets in front);
        ``release()`` (called once the batch has been gathered into the workspace) lets the copy stream reuse that set.
        The copies only overlap compute when the host tensors are PINNED (``DataLoader(pin_memory=True)``,
        ``SyntheticExpertLoader(pin=True)``); CUDA copies pageable tensors synchronously with the host, which is correct but
        serialises the upload with the enqueueing of the step.  A device-resident table (``DeviceExpertLoader``) needs no copy."""
        dev = self._dev()
        if dev.type != "cuda":
            for batch, idx in pairs:
                yield (batch if isinstance(batch, DeviceBatch) else self._to_dev(*batch)), idx, (lambda: None)
            return
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._stage_bufs = [None] * self.N_STAGING
        nsets = self.N_STAGING
        main = torch.cuda.current_stream(dev)
        trace = getattr(self, "_trace", None)      # diagnostics: list collecting (tag, start_event, end_event)
        ready = [torch.cuda.Event() for _ in range(nsets)]
        consumed = [None] * nsets

        def stage(pair, k):
            batch, idx = pair
            if isinstance(batch, DeviceBatch):               # device-resident expert data: nothing to copy
                ready[k].record(main)
                return batch, idx
            bufs = self._stage_bufs[k]
            if bufs is None or any(b.shape != t.shape or b.dtype != _stage_dtype(t) for b, t in zip(bufs, batch)):
                bufs = [torch.empty(t.shape, dtype=_stage_dtype(t), device=dev) for t in batch]
                self._stage_bufs[k] = bufs
                self._copy_stream.wait_stream(main)          # fresh buffers: order after whatever main was doing
            if consumed[k] is not None:
                self._copy_stream.wait_event(consumed[k])    # the previous tenant of this staging set has been gathered
            with torch.cuda.stream(self._copy_stream):
                if trace is not None:
                    c0 = torch.cuda.Event(enable_timing=True); c0.record(self._copy_stream)
                for b, t in zip(bufs, batch):
                    b.copy_(t, non_blocking=True)
                ready[k].record(self._copy_stream)
                if trace is not None:
                    c1 = torch.cuda.Event(enable_timing=True); c1.record(self._copy_stream)
                    if len(trace) < 4096:          # diagnostics only: never grow without bound
                        trace.append(("copy", c0, c1))
            return bufs, idx

        def releaser(k):
            def release():
                ev = torch.cuda.Event()
                ev.record(main)
                consumed[k] = ev
            return release

        it = iter(pairs)
        pending = []                  # staged batches not yet handed out: [((tensors, idx), set index)]
        state = {"next": 0, "done": False}

        def stage_one():
            try:
                pair = next(it)
            except StopIteration:
                state["done"] = True
                return
            k = state["next"]
            pending.append((stage(pair, k), k))
            state["next"] = (k + 1) % nsets

        while not state["done"] and len(pending) < nsets - 1:
            stage_one()
        while pending:
            cur, k = pending.pop(0)
            # the set handed out now is in use until release(); the other nsets-1 sets are staged ahead.  The set being
            # refilled here belonged to the batch handed out (and released) in the previous iteration.
            while not state["done"] and len(pending) < nsets - 1:
                stage_one()
            main.wait_event(ready[k])
            marker = consumed[k]
            yield cur[0], cur[1], releaser(k)
            if consumed[k] is marker:                        # caller forgot to release: be safe
                releaser(k)()

    # ---- algo/wdgail.py:100-147
    def update(self, expert_loader, rollouts):
        eng = self.engine
        eng.sync_params()
        dev = self._dev()
        B = expert_loader.batch_size
        world = world_size()
        rank = dist.get_rank() if world > 1 else 0
        # exact multi-GPU mode (see PPO.exact_sharding): `expert_loader` yields the same GLOBAL batches of B rows on every
        # rank, the policy minibatch is a global one, and this rank processes the positions whose env it owns - expert row
        # i, policy row i and alpha[i] stay paired exactly as algo/wdgail.py:66-80 pairs them
        exact = bool(self.exact_sharding) and world > 1
        if exact and rollouts.shard != (rank, world):
            raise RuntimeError("exact_sharding needs rollouts.set_shard(rank, world) on every rank")
        self.optimizer.grad_scale = 1.0 if exact else None
        obs_rows, met_rows, act_rows = rollouts.flat("obs"), rollouts.flat("metrics"), rollouts.flat("actions")
        n = 0
        # Mix-up coefficients (algo/wdgail.py:66: torch.rand(B,1,1,1) per batch on the CPU default generator).  All
        # host->device transfers share one DMA queue, so a small per-batch upload on the compute stream would queue
        # behind the expert prefetch and stall compute; when the number of batches is known the draws are made in
        # the reference's order up front and uploaded once.
        if exact:
            pairs = ((_rows_of(e, pos), (pos, idx)) for e, (pos, idx) in zip(expert_loader, rollouts.sharded_minibatches(B)))
        else:
            pairs = zip(expert_loader, ((None, idx) for idx in rollouts.minibatch_indices(B)))
        alphas = None
        n_batches = None
        if hasattr(expert_loader, "__len__"):
            first = next(pairs, None)          # advances both iterators exactly like the first zip step (draws randperm)
            n_rollout = rollouts.num_steps * rollouts.num_processes * (world if exact else 1)
            n_batches = min(len(expert_loader), n_rollout // B) if first is not None else 0
            if n_batches and dev.type == "cuda":
                host = torch.empty(n_batches, B, pin_memory=True)
                for i in range(n_batches):
                    host[i] = torch.rand(B, 1, 1, 1).view(B)
                alphas = host.to(dev, non_blocking=True)
            import itertools
            pairs = itertools.chain([first], pairs) if first is not None else iter(())
        opt = self.optimizer
        if n_batches:
            opt.begin_schedule(n_batches)
        # static device state of the replayable step, kept per (batch size, device) for the module's lifetime
        if self._alpha_buf is None:
            self._alpha_buf = {}
        if (B, str(dev)) not in self._alpha_buf:
            self._alpha_buf[(B, str(dev))] = (torch.zeros(8, dtype=torch.float64, device=dev), torch.zeros(B, device=dev))
        acc, alpha_buf = self._alpha_buf[(B, str(dev))]
        A.zero_block(acc)
        if not eng.flat.grad_clean:        # a replayed step assumes the zeroed gradient buffer the previous step left behind
            eng.flat.grad.zero_(); eng.flat.grad_clean = True

        state = {"premixed": False}

        def device_step():
            """Forward + full backward + optimiser step on the 2B rows already gathered into the workspace; reads only
            device-resident state (alpha_buf, Adam's scalars), so it is captured once and replayed (graphs.StepGraph)."""
            eng.update_step(B, alpha_buf, acc, reducer=opt.reducer, premixed=state["premixed"])
            opt.step(from_device_hyper=True)
            eng.dirty = True
            eng.sync_params()

        with torch.no_grad():
            for i_batch, (e_batch, (pos, idx), release) in enumerate(self._prefetched(pairs)):
                e_obs, e_met, e_act, e_idx, e_rows = expert_rows(e_batch, dev)
                Bl = int(idx.shape[0])
                if e_rows != Bl or (not exact and Bl != B):
                    raise ValueError("expert batches must all have expert_loader.batch_size rows (drop_last=True)")
                if alphas is not None and i_batch < alphas.shape[0]:
                    alpha = alphas[i_batch]
                else:
                    alpha = self._upload_alpha(torch.rand(B, 1, 1, 1).view(B))
                if n_batches is None or i_batch >= n_batches:
                    opt.begin_schedule(1)          # loader of unknown length: stage the scalars step by step
                opt.advance()
                if not exact:                      # fixed-shape step: inputs gathered eagerly, the rest replayed as a graph
                    ws = eng.workspace(3 * B)
                    alpha_buf.copy_(alpha, non_blocking=True)
                    state["premixed"] = eng.load_pair_mix(e_obs, e_met, e_act, e_idx, obs_rows, met_rows, act_rows, idx, alpha_buf, B)
                    release()
                    key = (ws.X0.data_ptr(), ws.rows, B, eng.flat.flat.data_ptr(), eng.flat.grad.data_ptr(), world, self.max_grad_norm,
                           state["premixed"], acc.data_ptr(), alpha_buf.data_ptr(), opt.device_hyper(eng.flat).data_ptr())
                    self._graph.run(key, device_step, dev)
                    n += B * world
                    continue
                alpha = alpha[pos.to(alpha.device)].contiguous()
                if Bl:
                    eng.workspace(3 * Bl)
                    eng.load_inputs(e_obs, e_met, e_act, e_idx, Bl, 0)
                    release()
                    eng.load_inputs(obs_rows, met_rows, act_rows, idx, Bl, Bl)
                    eng.update_step(Bl, alpha, acc, norm=B, reducer=opt.reducer)
                else:              # this rank owns no member of the global minibatch: zero gradient, still all-reduces
                    release()
                    eng.zero_contribution(opt.reducer)
                opt.step(from_device_hyper=True)
                eng.dirty = True
                eng.sync_params()
                n += B
        if world > 1:           # the tuple reports global-batch means (algo/wdgail.py:147)
            dist.all_reduce(acc, op=dist.ReduceOp.SUM)
        s_de, s_dp, s_te, s_tp, s_gp = acc[:5].cpu().tolist()   # single read-back per update
        wd = s_te - s_tp
        gp = 10.0 * s_gp
        return (-wd + gp) / n, s_dp / n, s_de / n, wd / n, gp / n, s_te / n, s_tp / n

    # ---- algo/wdgail.py:149-179
    def compute_loss(self, expert_loader, rollouts, batch_size=None):
        eng = self.engine
        eng.sync_params()
        dev = self._dev()
        B = expert_loader.batch_size
        world = world_size()
        exact = bool(self.exact_sharding) and world > 1
        obs_rows, met_rows, act_rows = rollouts.flat("obs"), rollouts.flat("metrics"), rollouts.flat("actions")
        acc = torch.zeros(8, dtype=torch.float64, device=dev)
        n = 0
        if exact:     # same global permutation / expert batches on every rank, each evaluates the positions it owns
            pairs = ((_rows_of(e, pos), idx) for e, (pos, idx) in zip(expert_loader, rollouts.sharded_minibatches(B, batch_size)))
        else:
            pairs = zip(expert_loader, rollouts.minibatch_indices(B, batch_size))
        with torch.no_grad():
            for expert_batch, idx in pairs:
                Bl = int(idx.shape[0])
                if Bl:
                    eng.workspace(3 * Bl)
                    e_obs, e_met, e_act, e_idx, _ = expert_rows(expert_batch, dev)
                    eng.load_inputs(e_obs, e_met, e_act, e_idx, Bl, 0)
                    eng.load_inputs(obs_rows, met_rows, act_rows, idx, Bl, Bl)
                    eng.tail_features(Bl, 0); eng.tail_features(Bl, Bl)
                    d = eng.forward(2 * Bl)
                    A.disc_loss_seed(d, eng.ws.buf("dd", eng.ws.rows), acc, Bl)     # only the tanh sums are used here
                n += B if exact else Bl * world
        if world > 1:           # every rank evaluated its own rows: report the mean over all of them
            dist.all_reduce(acc, op=dist.ReduceOp.SUM)
        _, _, s_te, s_tp = acc[:4].cpu().tolist()
        if n == 0:
            return 0, 0, 0
        return (s_te - s_tp) / n, s_te / n, s_tp / n

    # ---- algo/wdgail.py:181-189 (gamma, masks, update_rms are ignored by the reference, so also here)
    def predict_reward(self, state, metrics, action, gamma, masks, update_rms=True):
        with torch.no_grad():
            d = self.forward(state, metrics, action)
            r = torch.empty_like(d)
            A.reward_epilogue(d, r, d.numel())
            return r.cpu()

    def predict_rewards_rollout(self, rollouts, chunk: int = 4096) -> None:
        """tools/learn.py:196-202 as one batched pass: gail_rewards[t, n] for every stored step, written in HBM."""
        eng = self.engine
        eng.sync_params()
        T, N = rollouts.num_steps, rollouts.num_processes
        obs_rows, met_rows, act_rows = rollouts.flat("obs"), rollouts.flat("metrics"), rollouts.flat("actions")
        out = rollouts.gail_rewards.view(-1)
        with torch.no_grad():
            for s in range(0, T * N, chunk):
                B = min(chunk, T * N - s)
                eng.workspace(B)
                eng.load_inputs(obs_rows[s:], met_rows[s:], act_rows[s:], None, B, 0)     # consecutive rows: no index vector
                eng.tail_features(B, 0)
                d = eng.forward(B)
                A.reward_epilogue(d, out[s:], B)


def _stage_dtype(t: torch.Tensor):
    """Staging dtype of an expert-batch tensor: uint8 observations stay bytes, everything else is fp32."""
    return torch.uint8 if t.dtype == torch.uint8 else torch.float32


def _rows_of(batch, pos: torch.Tensor):
    """Rows `pos` (CPU int64) of an expert batch in either form (exact multi-GPU mode)."""
    if isinstance(batch, DeviceBatch):
        return DeviceBatch(batch.obs_table, batch.metrics_table, batch.actions_table, batch.idx[pos.to(batch.idx.device)].contiguous())
    return tuple(t[pos] for t in batch)


class _CriticOfImage(torch.autograd.Function):
    """Autograd bridge of ``Discriminator.forward(gp=True)``: D(x) with a first-order backward dD/dx through the engine's
    dgrad chain (algo/wdgail.py:85-91 with create_graph=False).  The workspace holds the activations, so backward must
    run before the next forward of this critic."""

    @staticmethod
    def forward(ctx, disc, state, metrics, action):
        eng = disc.engine
        B = state.shape[0]
        eng.workspace(B)
        eng.load_inputs(state.detach().contiguous(), metrics, action, None, B, 0)
        eng.tail_features(B, 0)
        out = eng.forward(B, training=True)[:B].clone().view(B, 1)
        eng.forward_serial = getattr(eng, "forward_serial", 0) + 1
        ctx.disc, ctx.B, ctx.serial = disc, B, eng.forward_serial
        return out

    @staticmethod
    def backward(ctx, g_out):
        disc, B = ctx.disc, ctx.B
        eng = disc.engine
        if eng.forward_serial != ctx.serial:
            raise RuntimeError("forward(gp=True): backward must run before the next forward pass of this Discriminator")
        ws, P, H_ = eng.ws, eng.flat.p, eng.hidden
        with torch.no_grad():
            dA = ws.grads()
            dd = ws.buf("dd", ws.rows)
            dd[:B].copy_(g_out.reshape(B).to(dd.device, torch.float32))
            H = ws.buf("H", ws.rows, LDH); dH = ws.buf("dH", ws.rows, LDH)
            A.small_linear_bwd(H, LDH, P("trunk.2.weight"), dd, 1, dH, LDH, None, None, B, 0, 1, H_, E.SLOPE)
            A.linear_dgrad(dH, LDH, eng.w1, LDF, dA[4], E.FEAT, B, E.FEAT, H_, mask_src=ws.F, ldm=LDF, slope=E.SLOPE,
                           mask_bits=ws.mbits[4])
            eng.conv.backward_data(ws, B)
            eng.conv.input_grad(ws, B, 0)
            # dX0 [B,96,96,(dy,dx,c4)] in normalised space -> d/d raw image [B,3,192,192] (x 1/std, drop the pad channel)
            g = dA[0][:B].view(B, 96, 96, 2, 2, 4)[..., :3].permute(0, 5, 1, 3, 2, 4).reshape(B, 3, 192, 192)
            g = g * torch.tensor(E.INV_STD, device=g.device).view(1, 3, 1, 1)
        return None, g, None, None
