"""Batched, device-resident form of the reference's per-image optimisation loop.

Reference path (one image at a time, two host syncs per step):
    optimize_images -> optimization           src/baselines/optimize_image.py:14-97
    objective_function_parametric             src/optimize_image_param.py:237-259
    apply_params (8 default filters)          src/baselines/image_transformations/image_transformations.py:7-66
    ValenceArousalLoss.forward                src/baselines/losses/ValenceArousalLoss.py:59-73
Here B images are B independent problems (own x, Adam state, target, crop draws, best-x) advanced together; one step is a
fixed sequence of librgie.so launches on one stream, captured once into a CUDA graph and replayed.  Step-dependent
scalars (learning-rate ramp, Adam bias corrections) and the per-step crop offsets are device tables indexed by a device
step counter, so there is no host round trip inside the loop.  PyTorch only owns the buffers.
"""
from __future__ import annotations

import math
import os
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib, ops
from ._lib import check, ptr

DEFAULT_FILTERS = ['exposure', 'saturation', 'tone', 'color', 'contrast', 'sharp', 'blur', 'scale']
_X0 = [0.0, 1.0] + [1.0] * 8 + [1.0] * 24 + [1.0, 0.0, 1e-4, 1.0, 1.0, 0.0, 0.0]     # optimize_image_param.py:121-209


def lr_schedule(step: int, num_steps: int, learning_rate: float, lr_rampdown_length: float = 0.25,
                lr_rampup_length: float = 0.05) -> float:
    """optimize_image.py:69-75 (host float64, identical expression order)."""
    t = step / num_steps
    lr_ramp = min(1.0, (1.0 - t) / lr_rampdown_length)
    lr_ramp = 0.5 - 0.5 * np.cos(lr_ramp * np.pi)
    lr_ramp = lr_ramp * min(1.0, t / lr_rampup_length)
    return float(learning_rate * lr_ramp)


def _on_own_device(fn):
    """Run a method with the engine's device current: handles, launches and streams all belong to self.dev."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *a, **kw):
        with torch.cuda.device(self.dev):
            return fn(self, *a, **kw)
    return wrapped


class ParametricEditEngine:
    def __init__(self, state_dict: Dict[str, torch.Tensor], batch: int, height: int, width: int, num_steps: int,
                 precision: str = "bf16", micro_batch: Optional[int] = None, input_size: int = 480,
                 crop_size: int = 448, reps: int = 10, device=None, use_graph: bool = True, folded=None):
        if not torch.cuda.is_available():
            raise _lib.RgieError("ParametricEditEngine needs a CUDA device: there is no CPU path")
        self.lib = _lib.load()
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.dev.index is None:
            self.dev = torch.device("cuda", torch.cuda.current_device())
        with torch.cuda.device(self.dev):
            self._build(state_dict, batch, height, width, num_steps, precision, micro_batch, input_size, crop_size, reps,
                        use_graph, folded)

    def _build(self, state_dict, batch, height, width, num_steps, precision, micro_batch, input_size, crop_size, reps,
               use_graph, folded):
        self.B, self.H, self.W, self.steps = batch, height, width, num_steps
        self.reps, self.input_size, self.crop = reps, input_size, crop_size
        self.mb = batch if micro_batch is None else micro_batch
        if batch % self.mb != 0:
            raise _lib.RgieError("batch must be a multiple of micro_batch")
        self.use_graph = use_graph
        self.filters = list(DEFAULT_FILTERS)
        self.kinds = [_lib.FILTER_KINDS[f] for f in self.filters]
        # exposure -> saturation -> tone -> colour run as one pixel pass each way (rgie_filter_prefix_*): stages 1..3 of
        # the chain are never materialised.  RGIE_FUSED_PREFIX=0 keeps the four separate passes (bit-identical images).
        self.n_prefix = 4 if (self.filters[:4] == DEFAULT_FILTERS[:4] and os.environ.get("RGIE_FUSED_PREFIX", "1") != "0") else 0
        self.poff, o = [], 0
        for f in self.filters:
            self.poff.append(o)
            o += _lib.FILTER_NPARAM[f]
        self.NP = o
        self.Hr, self.Wr = ops.resize_output_size(height, width, input_size)
        self.resize = ops.Resize(height, width, self.Hr, self.Wr)
        self.reg = ops.Regressor(state_dict, max_crops=self.mb * reps, crop_size=crop_size, precision=precision,
                                 device=self.dev, folded=folded)
        self.nc = self.reg.num_classes
        f32 = dict(dtype=torch.float32, device=self.dev)
        B, NP = batch, self.NP
        self.x = torch.empty(B, NP, **f32); self.p = torch.empty(B, NP, **f32); self.gp = torch.zeros(B, NP, **f32)
        self.m = torch.zeros(B, NP, **f32); self.v = torch.zeros(B, NP, **f32)
        self.best_x = torch.empty(B, NP, **f32); self.best_loss = torch.empty(B, **f32)
        self.best_step = torch.zeros(B, dtype=torch.int32, device=self.dev)
        self.loss = torch.zeros(B, **f32); self.preds = torch.zeros(B, self.nc, **f32)
        self.target = torch.zeros(B, 2, **f32)
        self.logits = torch.zeros(B * reps, self.nc, **f32); self.dlogits = torch.zeros(B * reps, self.nc, **f32)
        self.stage = [torch.empty(B, 3, height, width, **f32) if (k == 0 or k >= self.n_prefix) else None
                      for k in range(len(self.filters) + 1)]
        self.gA = torch.empty(B, 3, height, width, **f32); self.gB = torch.empty(B, 3, height, width, **f32)
        if self.resize.identity:
            self.resized, self.dresized = None, None
        else:
            self.resized = torch.empty(B, 3, self.Hr, self.Wr, **f32)
            self.dresized = torch.empty(B, 3, self.Hr, self.Wr, **f32)
        self.rs_tmp = torch.empty(B * 3 * height * self.Wr, **f32)
        self.ws = torch.empty(self.lib.rgie_filter_ws_floats(B, height, width), **f32)
        self.loss_log = torch.zeros(num_steps, B, **f32); self.pred_log = torch.zeros(num_steps, B, self.nc, **f32)
        self.x_log = torch.zeros(num_steps, B, NP, **f32)        # x BEFORE each step (what the objective was evaluated at)
        self.sched = torch.zeros(num_steps, 2, **f32)
        self.counter = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.offsets = torch.zeros(num_steps, B, reps, 2, dtype=torch.int32, device=self.dev)
        self.scale = 0.15
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.launches_per_step = 0
        self._warm = False
        self.steps_done = 0
        # optional (start, end) torch.cuda.Event pair around resize -> crops -> resnet50 fwd -> VA head -> resnet50 bwd ->
        # crop gather -> resize^T of an EAGER step (SURVEY.md 8d "regressor fwd+bwd ms"); never set while capturing
        self.span_events = None

    # ------------------------------------------------------------------------------------------------------------
    def _st(self):
        return _lib.stream_ptr(self.dev)

    def _filters_fwd(self, p: torch.Tensor):
        lib, st = self.lib, self._st()
        if self.n_prefix:
            check(lib.rgie_filter_prefix_fwd(ptr(self.stage[0]), ptr(self.stage[self.n_prefix]), ptr(p), self.NP, self.B,
                                             self.H, self.W, st), "filter_prefix_fwd")
        for k, kind in enumerate(self.kinds):
            if k < self.n_prefix:
                continue
            check(lib.rgie_filter_fwd(kind, ptr(self.stage[k]), ptr(self.stage[k + 1]), p.data_ptr() + 4 * self.poff[k],
                                      self.NP, self.B, self.H, self.W, ptr(self.ws), st), "filter_fwd")

    def _regressor_fwd(self, img: torch.Tensor, offsets: torch.Tensor, step_ptr, stride: int):
        """img [B,3,H,W] -> self.logits, through resize + crops, micro-batch by micro-batch (forward only)."""
        lib, st = self.lib, self._st()
        src = img
        if not self.resize.identity:
            check(lib.rgie_resize_fwd(self.resize._h, ptr(img), ptr(self.resized), self.B * 3, ptr(self.rs_tmp), st),
                  "resize_fwd")
            src = self.resized
        return src

    def _step(self):
        """One optimisation step for all B problems (graph-capturable: no allocation, no sync)."""
        lib, st = self.lib, self._st()
        B, mb, reps, nc = self.B, self.mb, self.reps, self.nc
        n_launch = 0
        check(lib.rgie_params_default_fwd(ptr(self.x), ptr(self.p), B, float(self.H), st), "params_fwd")
        check(lib.rgie_record(ptr(self.x), ptr(self.x_log), ptr(self.counter), B * self.NP, st), "record")
        self._filters_fwd(self.p)
        if self.span_events is not None:           # eager timing of the regressor span (bench.py: regressor_fwd_bwd_ms)
            self.span_events[0].record()
        src = self._regressor_fwd(self.stage[-1], self.offsets, self.counter, 0)
        dsrc = self.gA if self.resize.identity else self.dresized
        stride = B * reps * 2
        for c in range(B // mb):
            img_c = src[c * mb:(c + 1) * mb]
            off_c = self.offsets[0, c * mb:(c + 1) * mb]
            lg, dlg = self.logits[c * mb * reps:(c + 1) * mb * reps], self.dlogits[c * mb * reps:(c + 1) * mb * reps]
            check(lib.rgie_regressor_forward_ex(self.reg._h, ptr(img_c), mb, self.Hr, self.Wr, ptr(off_c),
                                                ptr(self.counter), stride, reps, 1, ptr(lg), st), "regressor_forward")
            check(lib.rgie_va_head(ptr(lg), mb, reps, nc, 1, ptr(self.target[c * mb:(c + 1) * mb]), 0.5, 0.0, 3, self.scale,
                                   ptr(self.preds[c * mb:(c + 1) * mb]), ptr(self.loss[c * mb:(c + 1) * mb]), ptr(dlg), st),
                  "va_head")
            check(lib.rgie_regressor_backward(self.reg._h, ptr(dlg), ptr(dsrc[c * mb:(c + 1) * mb]), st),
                  "regressor_backward")
        check(lib.rgie_record(ptr(self.loss), ptr(self.loss_log), ptr(self.counter), B, st), "record")
        check(lib.rgie_record(ptr(self.preds), ptr(self.pred_log), ptr(self.counter), B * nc, st), "record")
        g_cur, g_nxt = self.gA, self.gB
        if not self.resize.identity:
            check(lib.rgie_resize_bwd(self.resize._h, ptr(self.dresized), ptr(g_cur), B * 3, ptr(self.rs_tmp), st),
                  "resize_bwd")
        if self.span_events is not None:
            self.span_events[1].record()
        for k in reversed(range(self.n_prefix, len(self.kinds))):
            check(lib.rgie_filter_bwd(self.kinds[k], ptr(self.stage[k]), ptr(g_cur), ptr(g_nxt),
                                      self.p.data_ptr() + 4 * self.poff[k], self.NP, self.gp.data_ptr() + 4 * self.poff[k],
                                      self.NP, B, self.H, self.W, ptr(self.ws), st), "filter_bwd")
            g_cur, g_nxt = g_nxt, g_cur
        if self.n_prefix:
            check(lib.rgie_filter_prefix_bwd(ptr(self.stage[0]), ptr(g_cur), ptr(self.p), self.NP, ptr(self.gp), self.NP, B,
                                             self.H, self.W, ptr(self.ws), st), "filter_prefix_bwd")
        check(lib.rgie_params_default_bwd(ptr(self.x), ptr(self.gp), B, float(self.H), st), "params_bwd")
        check(lib.rgie_adam_step_sched(ptr(self.x), ptr(self.gp), ptr(self.m), ptr(self.v), B, self.NP, ptr(self.sched),
                                       ptr(self.counter), 1.0 - 0.9, 0.999, 1.0 - 0.999, 1e-8, ptr(self.loss),
                                       ptr(self.best_loss), ptr(self.best_x), ptr(self.best_step), st), "adam")
        check(lib.rgie_counter_add(ptr(self.counter), 1, st), "counter")
        return n_launch

    @_on_own_device
    def predict(self, images: torch.Tensor, offsets: torch.Tensor) -> torch.Tensor:
        """No-grad regressor prediction (ValenceArousalLoss.predict_loss_metric, :131-138): [B,nc] after sigmoid."""
        lib, st = self.lib, self._st()
        B, mb, reps, nc = self.B, self.mb, self.reps, self.nc
        src = self._regressor_fwd(images, offsets, None, 0)
        preds = torch.empty(B, nc, dtype=torch.float32, device=self.dev)
        offsets = offsets.contiguous()
        for c in range(B // mb):
            lg = self.logits[c * mb * reps:(c + 1) * mb * reps]
            check(lib.rgie_regressor_forward_ex(self.reg._h, ptr(src[c * mb:(c + 1) * mb]), mb, self.Hr, self.Wr,
                                                ptr(offsets[c * mb:(c + 1) * mb]), None, 0, reps, 1, ptr(lg), st),
                  "regressor_forward")
            check(lib.rgie_va_head(ptr(lg), mb, reps, nc, 1, None, 0.5, 0.0, 0, 1.0, ptr(preds[c * mb:(c + 1) * mb]), None,
                                   None, st), "va_head")
        return preds

    # ------------------------------------------------------------------------------------------------------------
    @_on_own_device
    def load_problem(self, images: torch.Tensor, offsets: torch.Tensor, alpha: Optional[float] = 0.1,
                     target: Optional[torch.Tensor] = None, learning_rate: float = 0.05, weight_clf: float = 0.15,
                     clf_weight: float = 1.0, x0: Optional[torch.Tensor] = None) -> None:
        """images [B,3,H,W] (device), offsets int32 [1+steps, B, reps, 2] (device): draw 0 feeds
        get_condition_from_alpha (optimize_image.py:34-36,119-123), draw 1+s feeds step s."""
        if tuple(images.shape) != (self.B, 3, self.H, self.W) or not images.is_cuda:
            raise _lib.RgieError(f"load_problem: images must be a CUDA tensor of shape {(self.B, 3, self.H, self.W)}")
        if offsets.dtype != torch.int32 or tuple(offsets.shape[1:]) != (self.B, self.reps, 2):
            raise _lib.RgieError(f"load_problem: offsets must be int32 [1+steps or steps, {self.B}, {self.reps}, 2]")
        if offsets.shape[0] not in (self.steps, 1 + self.steps):
            raise _lib.RgieError(f"load_problem: offsets hold {offsets.shape[0]} draws, expected {self.steps} or {1 + self.steps}")
        # the crop kernels index the resized image with these: (top, left) must keep the crop inside it
        lo, hi_t, hi_l = int(offsets.min()), int(offsets[..., 0].max()), int(offsets[..., 1].max())
        if lo < 0 or hi_t > self.Hr - self.crop or hi_l > self.Wr - self.crop:
            raise _lib.RgieError(f"load_problem: crop offsets outside [0, {self.Hr - self.crop}] x [0, {self.Wr - self.crop}]")
        self.steps_done = 0
        self.stage[0].copy_(images)
        new_scale = float(weight_clf) * float(clf_weight)
        if new_scale != self.scale:
            self.graph = None                 # the loss scale is a baked kernel argument of the captured step
        self.scale = new_scale
        if offsets.shape[0] == 1 + self.steps:
            self.offsets.copy_(offsets[1:])
            self.pred0 = self.predict(self.stage[0], offsets[0])
        else:
            if target is None:
                raise _lib.RgieError("load_problem: offsets must be [1+steps, ...] unless a target is given")
            self.offsets.copy_(offsets)
            self.pred0 = torch.zeros(self.B, self.nc, dtype=torch.float32, device=self.dev)
        if target is None:
            target = torch.clamp(self.pred0[:, :2] + float(alpha), 0.0, 1.0)
        self.target.copy_(target.to(self.dev).reshape(-1, 2).expand(self.B, 2))
        x0 = torch.tensor(_X0, dtype=torch.float32) if x0 is None else x0
        self.x.copy_(x0.to(self.dev).expand(self.B, self.NP))
        self.best_x.copy_(self.x)
        self.m.zero_(); self.v.zero_(); self.gp.zero_()
        self.best_loss.fill_(float("inf")); self.best_step.zero_(); self.counter.zero_()
        sched = np.zeros((self.steps, 2), np.float64)
        for s in range(self.steps):
            lr = lr_schedule(s, self.steps, learning_rate)
            sched[s] = ops.adam_scalars(lr, s + 1)
        self.sched.copy_(torch.from_numpy(sched.astype(np.float32)))

    @_on_own_device
    def ensure_graph(self) -> None:
        """Capture one step into a CUDA graph (capturing does not execute; needs one prior eager step so that every
        kernel is instantiated and its function attributes are set outside the capture)."""
        if self.graph is not None or not self.use_graph:
            return
        if not self._warm:
            raise _lib.RgieError("ensure_graph: run one eager step first (advance(1))")
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._step()
        self.graph = g

    @_on_own_device
    def advance(self, n_steps: int) -> None:
        """Run n_steps optimisation steps: the very first step of an engine eagerly, the rest as CUDA-graph replays.
        The device step counter indexes the per-step tables (offsets, schedule, logs): never step past num_steps."""
        if self.steps_done + n_steps > self.steps:
            raise _lib.RgieError(f"advance({n_steps}): {self.steps_done} of {self.steps} steps already done "
                                 f"(load_problem() starts a new run)")
        self.steps_done += n_steps
        done = 0
        if self.use_graph and self.graph is None and n_steps > 0:
            self._step()
            self._warm = True
            done = 1
            if n_steps > 1:
                self.ensure_graph()
        for _ in range(done, n_steps):
            if self.graph is not None:
                self.graph.replay()
            else:
                self._step()

    @_on_own_device
    def probe_gradient(self, x: torch.Tensor, step: int) -> Dict[str, torch.Tensor]:
        """Diagnostic (teacher forcing): evaluate the objective and d(loss)/d(x) of all B problems AT the given raw
        parameter vectors x [B,NP] with the crop draws of optimisation step `step`, without touching the run state that
        matters afterwards (x, Adam moments, best-x and the counter are restored).  Returns loss [B], preds, grad [B,NP]."""
        if not 0 <= step < self.steps:
            raise _lib.RgieError(f"probe_gradient: step {step} outside [0, {self.steps})")
        keep = [t.clone() for t in (self.x, self.m, self.v, self.best_x, self.best_loss, self.best_step, self.counter)]
        self.x.copy_(x.to(self.dev).reshape(-1, self.NP).expand(self.B, self.NP))
        self.counter.fill_(step)
        self._step()
        out = dict(loss=self.loss.clone(), preds=self.preds.clone(), grad=self.gp.clone())
        for dst, src in zip((self.x, self.m, self.v, self.best_x, self.best_loss, self.best_step, self.counter), keep):
            dst.copy_(src)
        return out

    @_on_own_device
    def results(self) -> Dict[str, torch.Tensor]:
        """Edited images from best_x (optimize_image.py:97 returns best_x; output_transform re-applies the filters)."""
        lib, st = self.lib, self._st()
        pbest = torch.empty_like(self.p)
        check(lib.rgie_params_default_fwd(ptr(self.best_x), ptr(pbest), self.B, float(self.H), st), "params_fwd")
        self._filters_fwd(pbest)
        return dict(best_x=self.best_x.clone(), x_last=self.x.clone(), losses=self.loss_log.clone(),
                    preds=self.pred_log.clone(), xs=self.x_log.clone(), target=self.target.clone(), pred0=self.pred0.clone(),
                    best_loss=self.best_loss.clone(), best_step=self.best_step.clone(),
                    edited=self.stage[-1].clone())

    def run(self, images: torch.Tensor, offsets: torch.Tensor, alpha: Optional[float] = 0.1,
            target: Optional[torch.Tensor] = None, learning_rate: float = 0.05, weight_clf: float = 0.15,
            clf_weight: float = 1.0, x0: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        self.load_problem(images, offsets, alpha, target, learning_rate, weight_clf, clf_weight, x0)
        self.advance(self.steps)
        return self.results()
