"""Training step of the B200 MultiTaskNet: flat parameter storage, train-mode forward / backward through
the C ABI, and the data-parallel trainer of BASELINE.json configs[4].

Reference behaviour reproduced (train.py:24-108, libs/loss.py):
  * `model.train()` forward uses batch statistics in every BatchNorm2d and updates `running_mean` /
    `running_var` (momentum 0.1, unbiased variance) and `num_batches_tracked`;
  * loss = 0.001 * CrossEntropy(logits, label) + JointsMSELoss(heatmap, target, target_weight);
  * optimiser = torch.optim.AdamW(model.parameters(), lr) (weight decay 0.01).

Storage: the module's parameters become views of ONE flat fp32 device block whose layout the library owns
(state_dict order), the gradients live in a second block of the same layout, the BatchNorm running statistics
in a third.  `torch.optim.AdamW` keeps working on the views (the drop-in path: `MultiTaskNet.forward` in train
mode is a `torch.autograd.Function`), and `DataParallelTrainer` runs the fused path: forward -> loss kernel ->
backward -> ONE all-reduce over the flat gradient block (NCCL over NVLink; the reference itself is
single-GPU, train.py:228-229) -> fused AdamW.  There is no CPU or PyTorch-operator fallback.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib, packing

BN_MOMENTUM = 0.1  # nn.BatchNorm2d default, reference model/gelan.py:46


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


class TrainPlan:
    """A bound hgr_train_plan (one per batch size) plus the workspace tensor that owns its memory."""

    def __init__(self, state: "TrainState", batch: int):
        lib = _lib.load()
        m = state.model
        s, j, c = m.image_size[0], m.num_joints, m.num_classes
        nbytes = lib.hgr_train_workspace_bytes(s, j, c, batch)
        if nbytes == 0:
            _lib.check(-1, "hgr_train_workspace_bytes")
        self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=state.device)
        self.state = state
        self.batch = batch
        handle = C.c_void_p()
        _lib.check(lib.hgr_train_plan_create(C.byref(handle), s, j, c, batch, state.params.data_ptr(),
                                             state.grads.data_ptr(), state.bnstats.data_ptr(),
                                             state.pos_embedding.data_ptr(), self.workspace.data_ptr(), nbytes),
                   "hgr_train_plan_create")
        self.handle = handle
        self.generation = 0  # bumped by every train-mode forward: the activations of the previous one are gone

    def buffer(self, name: str) -> torch.Tensor:
        """bf16 view of a named workspace buffer (activations, gradients, probabilities) for tests / attnmap."""
        ptr, nb = C.c_void_p(), C.c_size_t()
        dims = (C.c_int64 * 4)()
        _lib.check(_lib.load().hgr_train_buffer(self.handle, name.encode(), C.byref(ptr), C.byref(nb), dims),
                   "hgr_train_buffer")
        off = ptr.value - self.workspace.data_ptr()
        t = self.workspace[off: off + nb.value].view(torch.bfloat16)
        return t.view(*dims) if dims[0] > 0 else t

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.load().hgr_train_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class TrainState:
    """Flat fp32 parameter / gradient / running-statistics blocks of one MultiTaskNet on one device."""

    def __init__(self, model, device):
        self.model = model
        self.device = device
        j, c = model.num_joints, model.num_classes
        lib = _lib.load()
        self.layout = _lib.train_param_layout(j, c)
        self.bn_layout = _lib.train_bnstat_layout()
        self.numel = lib.hgr_train_param_floats(j, c)
        self.params = torch.zeros(self.numel, dtype=torch.float32, device=device)
        self.grads = torch.zeros(self.numel, dtype=torch.float32, device=device)
        self.bnstats = torch.zeros(lib.hgr_train_bnstat_floats(), dtype=torch.float32, device=device)
        named_p = dict(model.named_parameters())
        named_b = dict(model.named_buffers())
        if len(named_p) != len(self.layout):
            raise RuntimeError(f"parameter layout mismatch: module has {len(named_p)}, library {len(self.layout)}")
        self._views = []
        with torch.no_grad():
            for name, off, n in self.layout:
                p = named_p[name]
                if p.numel() != n:
                    raise RuntimeError(f"parameter '{name}': {p.numel()} elements, library expects {n}")
                view = self.params[off: off + n].view(p.shape)
                view.copy_(p.detach().to(device=device, dtype=torch.float32))
                p.data = view
                self._views.append((p, off, n))
            for name, off, n in self.bn_layout:
                b = named_b[name]
                view = self.bnstats[off: off + n].view(b.shape)
                view.copy_(b.detach().to(device=device, dtype=torch.float32))
                b.data = view
            nbt = [(k, b) for k, b in named_b.items() if k.endswith("num_batches_tracked")]
            self.num_batches_tracked = torch.zeros(len(nbt), dtype=torch.int64, device=device)
            for i, (_, b) in enumerate(nbt):
                self.num_batches_tracked[i] = b.to(device)
                b.data = self.num_batches_tracked[i]
        f = model.image_size[0] // 16
        self.pos_embedding = packing.sincos_table(f, f).to(device).to(torch.bfloat16).contiguous()
        self._plans = {}

    def attached(self) -> bool:
        """True while every parameter still is a view of the flat block (`.to()` / `load_state_dict` keep it)."""
        base = self.params.data_ptr()
        return all(p.data_ptr() == base + 4 * off and p.device == self.params.device for p, off, _ in self._views)

    def plan_for(self, batch: int) -> TrainPlan:
        if batch not in self._plans:
            self._plans[batch] = TrainPlan(self, batch)
        return self._plans[batch]

    def grad_views(self):
        return [self.grads[off: off + n].view(p.shape) for p, off, n in self._views]


def train_state(model, device) -> TrainState:
    device = torch.device(device)
    if device.type == "cuda" and device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    st = getattr(model, "_train_state_obj", None)
    if st is None or st.device != device or not st.attached():
        st = TrainState(model, device)
        model._train_state_obj = st
        model._packed = None  # the inference pack is keyed on data pointers; force a re-pack after re-homing
    return st


def _dt(t):
    return _lib.F32 if t.dtype == torch.float32 else _lib.BF16


def forward_train(state: TrainState, x: torch.Tensor, update_running: bool = True):
    """Raw train-mode forward: (logits fp32 (B, C), heatmaps fp32 (B, J, S/4, S/4), plan)."""
    m = state.model
    b, s = x.shape[0], m.image_size[0]
    plan = state.plan_for(b)
    logits = torch.empty(b, m.num_classes, dtype=torch.float32, device=x.device)
    heat = torch.empty(b, m.num_joints, s // 4, s // 4, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().hgr_train_forward(plan.handle, x.data_ptr(), _dt(x), logits.data_ptr(), heat.data_ptr(),
                                                 BN_MOMENTUM if update_running else -1.0, _stream(x.device)),
                   "hgr_train_forward")
        plan.generation += 1
        if update_running:
            state.num_batches_tracked += 1
    return logits, heat, plan


def backward_train(state: TrainState, plan: TrainPlan, x, dlogits, dheat, part=None):
    """Fills state.grads from the loss gradients (fp32, contiguous).  part = 0, 1, 2 (in this order) runs one third of
    the backward: each completes one contiguous range of the gradient block (grad_buckets)."""
    with torch.cuda.device(x.device):
        if part is None:
            _lib.check(_lib.load().hgr_train_backward(plan.handle, x.data_ptr(), _dt(x), dlogits.data_ptr(),
                                                      dheat.data_ptr(), _stream(x.device)), "hgr_train_backward")
        else:
            _lib.check(_lib.load().hgr_train_backward_part(plan.handle, x.data_ptr(), _dt(x), dlogits.data_ptr(),
                                                           dheat.data_ptr(), part, _stream(x.device)),
                       "hgr_train_backward_part")


def grad_buckets(num_joints: int, num_classes: int):
    """[start, stop) of the gradient block's three ranges in the order the backward completes them
    (hgr_train_backward_part 0, 1, 2): proj + decoder, encoder.down2 + cspelan3, the rest of the backbone."""
    layout = _lib.train_param_layout(num_joints, num_classes)
    off = {name: o for name, o, _ in layout}
    total = layout[-1][1] + layout[-1][2]
    a, b = off["encoder.down2.conv.weight"], off["proj.weight"]
    return [(b, total), (a, b), (0, a)]


class _TrainFunction(torch.autograd.Function):
    """MultiTaskNet.forward under .train(): the drop-in path for the reference's train.py (Lightning backward +
    torch.optim.AdamW).  Parameters are passed so that autograd routes their gradients."""

    @staticmethod
    def forward(ctx, model, x, *params):
        if ctx.needs_input_grad[1]:
            raise RuntimeError("the B200 MultiTaskNet does not compute a gradient for its input "
                               "(x.requires_grad is set); detach the input")
        state = train_state(model, x.device)
        logits, heat, plan = forward_train(state, x)
        ctx.state, ctx.plan, ctx.x, ctx.generation = state, plan, x, plan.generation
        attn = None
        if model.return_attention:
            t = model.image_size[0] // 16 * (model.image_size[0] // 16) + 1
            # rows are stored with a padded pitch; ALWAYS a copy (a bf16 `.to(bf16)` would be a live view of the
            # workspace that the next forward overwrites)
            attn = plan.buffer(f"l{3}.probs")[..., :t].to(x.dtype, copy=True)
            ctx.mark_non_differentiable(attn)
        if x.dtype != torch.float32:
            logits, heat = logits.to(x.dtype), heat.to(x.dtype)
        return logits, heat, attn

    @staticmethod
    def backward(ctx, dlogits, dheat, _dattn):
        state, plan, x = ctx.state, ctx.plan, ctx.x
        if plan.generation != ctx.generation:
            raise RuntimeError("backward of a train-mode forward whose activations are gone: another train-mode forward "
                               "of the same batch size ran in between (the activations live in one workspace per batch "
                               "size); run backward before the next forward, or use .eval() for the extra call")
        m = state.model
        if dlogits is None:
            dlogits = torch.zeros(x.shape[0], m.num_classes, device=x.device)
        if dheat is None:
            dheat = torch.zeros(x.shape[0], m.num_joints, m.image_size[0] // 4, m.image_size[0] // 4, device=x.device)
        backward_train(state, plan, x, dlogits.float().contiguous(), dheat.float().contiguous())
        # clones: autograd may keep or accumulate into what it is given, the flat block is rewritten every step
        grads = [g.clone() for g in state.grad_views()]
        return (None, None, *grads)


def forward_autograd(model, x):
    state = train_state(model, x.device)
    params = [p for p, _, _ in state._views]
    return _TrainFunction.apply(model, x, *params)


def loss_and_grads(logits, heat, labels, target, target_weight, cls_weight=0.001, want_grads=True):
    """train.py:63-64 on the device: returns (loss3 = [total, class, joints], dlogits, dheat)."""
    b, c = logits.shape
    _, j, h, w = heat.shape
    dev = logits.device
    loss3 = torch.empty(3, dtype=torch.float32, device=dev)
    scratch = torch.empty(1024, dtype=torch.float32, device=dev)
    dlogits = torch.empty_like(logits) if want_grads else None
    dheat = torch.empty_like(heat) if want_grads else None
    # keep every converted operand alive until the launch is enqueued (a temporary's block could be re-used)
    tw = target_weight.reshape(b, j).float().contiguous()
    lab = labels.to(torch.int64).contiguous()
    tgt = target.float().contiguous()
    logits, heat = logits.contiguous(), heat.contiguous()
    with torch.cuda.device(dev):
        _lib.check(_lib.load().hgr_loss(logits.data_ptr(), heat.data_ptr(), lab.data_ptr(),
                                        tgt.data_ptr(), tw.data_ptr(), b, j, c, h * w,
                                        cls_weight, dlogits.data_ptr() if want_grads else None,
                                        dheat.data_ptr() if want_grads else None, scratch.data_ptr(),
                                        loss3.data_ptr(), _stream(dev)), "hgr_loss")
    return loss3, dlogits, dheat


def broadcast_from_rank0_(tensors, group=None) -> int:
    """Make every rank start from rank 0's values (parameters, BatchNorm statistics, optimiser moments), like
    DistributedDataParallel does at construction: identical updates only keep replicas identical if they START
    identical, and a rank-dependent seed or a checkpoint loaded on rank 0 only would otherwise diverge silently.
    Returns the world size."""
    if not (dist.is_available() and dist.is_initialized()):
        return 1
    world = dist.get_world_size(group)
    if world > 1:
        src = dist.get_global_rank(group, 0) if group is not None else 0
        for t in tensors:
            dist.broadcast(t, src=src, group=group)
    return world


def replicas_in_sync(flat: torch.Tensor, group=None) -> bool:
    """True when every rank holds the same values in `flat` (checked with MIN / MAX all-reduces of a checksum)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return True
    v = flat.double()
    chk = torch.stack([v.sum(), (v * v).sum()])
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    return bool(torch.equal(lo, hi))


def allreduce_sum_(flat: torch.Tensor, group=None) -> int:
    """SUM all-reduce of the flat gradient block over the data-parallel group; returns the world size.
    (One call, one bucket: 7.4 M fp32 values = 29.6 MB cross NVSwitch in tens of microseconds, so the cost is
    launch latency, not bandwidth - SURVEY.md 8e.)"""
    if not (dist.is_available() and dist.is_initialized()):
        return 1
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return world


class DataParallelTrainer:
    """One rank of the data-parallel training step (BASELINE.json configs[4]).

    step(x, labels, target, target_weight): local forward with LOCAL BatchNorm batch statistics (the reference
    has plain nn.BatchNorm2d, no SyncBN) -> loss kernel -> backward -> all-reduce(SUM) of the flat gradient
    block -> fused AdamW with grad_scale = 1 / world (identical update on every rank).  Running statistics stay
    rank-local, like DDP with broadcast_buffers=False.
    """

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, cls_weight=0.001, group=None,
                 cuda_graph=False, overlap_allreduce=False):
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("DataParallelTrainer needs the module on a CUDA device (no CPU path)")
        self.model = model
        self.state = train_state(model, dev)
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.cls_weight = cls_weight
        self.group = group
        # cuda_graph=True: forward + loss + backward (~360 launches of a few microseconds each at batch 32) are
        # captured once per input shape and replayed; the all-reduce and the AdamW update stay ordinary launches
        self.cuda_graph = bool(cuda_graph)
        self._graphs = {}
        # overlap_allreduce (world > 1): the backward runs in three parts and the all-reduce of the range each part
        # completes (proj + decoder, then down2 + cspelan3 = 58 % of the block) rides on a side stream under the
        # next part; only the last range (conv1 .. cspelan2, 18 %) is exchanged after the backward.  Bitwise the
        # same step, but OFF by default: measured on 2 B200s (profiles/r2e_overlap_2gpu.txt) it is slower, 4.88 ms
        # against 4.75 ms - the whole exchange costs under 0.1 ms of the step, less than the three graph replays,
        # the two extra collective launches and the SMs NCCL takes from the persistent backward kernels
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.overlap = bool(overlap_allreduce) and world > 1
        self._buckets = grad_buckets(model.num_joints, model.num_classes)
        self._side = torch.cuda.Stream(dev) if self.overlap else None
        self.exp_avg = torch.zeros_like(self.state.params)
        self.exp_avg_sq = torch.zeros_like(self.state.params)
        self.steps = 0
        # rank 0's weights, running statistics and moments are the replica every rank starts from
        broadcast_from_rank0_([self.state.params, self.state.bnstats, self.state.num_batches_tracked,
                               self.exp_avg, self.exp_avg_sq], group)

    def _forward_loss(self, x, labels, target, target_weight, update_running=True):
        st = self.state
        logits, heat, plan = forward_train(st, x, update_running)
        loss3, dlogits, dheat = loss_and_grads(logits, heat, labels, target, target_weight, self.cls_weight)
        return loss3, plan, dlogits, dheat

    def _forward_backward(self, x, labels, target, target_weight, update_running=True):
        loss3, plan, dlogits, dheat = self._forward_loss(x, labels, target, target_weight, update_running)
        backward_train(self.state, plan, x, dlogits, dheat)
        return loss3

    def _stages(self, x, labels, target, target_weight):
        """The step's device work as three callables (forward + loss + backward part 0, part 1, part 2) and the loss
        tensor: plain launches, or three CUDA graphs captured once per input shape."""
        if not self.cuda_graph:
            box = {}

            def s0():
                box["v"] = self._forward_loss(x, labels, target, target_weight)
                backward_train(self.state, box["v"][1], x, box["v"][2], box["v"][3], 0)

            def part(k):
                return lambda: backward_train(self.state, box["v"][1], x, box["v"][2], box["v"][3], k)

            return [s0, part(1), part(2)], lambda: box["v"][0]
        key = ("parts", tuple(x.shape), x.dtype, tuple(target.shape))
        entry = self._graphs.get(key)
        if entry is None:
            dev = x.device
            static = [torch.empty_like(t) for t in (x, labels, target, target_weight)]
            for s_, t in zip(static, (x, labels, target, target_weight)):
                s_.copy_(t)
            self._forward_backward(*static, update_running=False)  # warm-up outside the capture, as in _forward_backward_graph
            torch.cuda.synchronize(dev)
            graphs = [torch.cuda.CUDAGraph() for _ in range(3)]
            pool = torch.cuda.graph_pool_handle()
            with torch.cuda.graph(graphs[0], pool=pool):
                loss3, plan, dlogits, dheat = self._forward_loss(*static)
                backward_train(self.state, plan, static[0], dlogits, dheat, 0)
            for k in (1, 2):
                with torch.cuda.graph(graphs[k], pool=pool):
                    backward_train(self.state, plan, static[0], dlogits, dheat, k)
            # the graphs hold raw pointers into the plan's workspace and into the loss gradients: keep their owners
            entry = self._graphs[key] = (graphs, static, loss3, (plan, dlogits, dheat))
        graphs, static, loss3, _ = entry
        for s_, t in zip(static, (x, labels, target, target_weight)):
            if s_.data_ptr() != t.data_ptr():
                s_.copy_(t, non_blocking=True)
        return [g.replay for g in graphs], lambda: loss3

    def _forward_backward_graph(self, x, labels, target, target_weight):
        key = (tuple(x.shape), x.dtype, tuple(target.shape))
        entry = self._graphs.get(key)
        if entry is None:
            dev = x.device
            static = [torch.empty_like(t) for t in (x, labels, target, target_weight)]
            for s_, t in zip(static, (x, labels, target, target_weight)):
                s_.copy_(t)
            # plan creation, kernel attributes and the first-touch allocations happen outside the capture; the
            # warm-up leaves the running statistics alone (the captured forward updates them exactly once per step)
            self._forward_backward(*static, update_running=False)
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                loss3 = self._forward_backward(*static)
            entry = self._graphs[key] = (graph, static, loss3)
            # the capture itself does not execute: replay below performs the step
        graph, static, loss3 = entry
        for s_, t in zip(static, (x, labels, target, target_weight)):
            if s_.data_ptr() != t.data_ptr():
                s_.copy_(t, non_blocking=True)
        graph.replay()
        return loss3

    def _step_overlapped(self, x, labels, target, target_weight):
        """forward + backward in three parts with the all-reduce of each finished gradient range on a side stream."""
        st = self.state
        main = torch.cuda.current_stream(x.device)
        stages, loss = self._stages(x, labels, target, target_weight)
        for k, run in enumerate(stages):
            run()
            lo, hi = self._buckets[k]
            if k < 2:
                self._side.wait_stream(main)  # the range is final once this part has run
                with torch.cuda.stream(self._side):
                    dist.all_reduce(st.grads[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
            else:
                dist.all_reduce(st.grads[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
        main.wait_stream(self._side)
        return loss(), dist.get_world_size(self.group)

    def step(self, x, labels, target, target_weight):
        st = self.state
        if self.overlap:
            loss3, world = self._step_overlapped(x, labels, target, target_weight)
        else:
            if self.cuda_graph:
                loss3 = self._forward_backward_graph(x, labels, target, target_weight)
            else:
                loss3 = self._forward_backward(x, labels, target, target_weight)
            world = allreduce_sum_(st.grads, self.group)
        self.steps += 1
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().hgr_adamw_step(st.params.data_ptr(), st.grads.data_ptr(), self.exp_avg.data_ptr(),
                                                  self.exp_avg_sq.data_ptr(), st.numel, self.lr, self.betas[0],
                                                  self.betas[1], self.eps, self.weight_decay, self.steps, 1.0 / world,
                                                  _stream(x.device)), "hgr_adamw_step")
        self.model._packed = None  # the eval-mode weight pack is stale now (the kernel does not bump tensor versions)
        return loss3

    # ---- checkpoint / resume (train.py:50-51 + Lightning's optimizer_states) ----
    def state_dict(self):
        """The optimiser state in torch.optim.AdamW's own state_dict layout (one group, parameters in
        model.parameters() order, per-parameter step / exp_avg / exp_avg_sq), so that a checkpoint written here
        resumes under `torch.optim.AdamW(model.parameters())` and the other way round.  The weights and BatchNorm
        statistics are the module's own state_dict."""
        group = dict(torch.optim.AdamW([torch.nn.Parameter(torch.zeros(1))]).defaults)
        group.update(lr=self.lr, betas=tuple(self.betas), eps=self.eps, weight_decay=self.weight_decay,
                     params=list(range(len(self.state.layout))))
        state = {}
        if self.steps > 0:
            for i, (_, off, n, shape) in enumerate(self._param_views()):
                state[i] = {"step": torch.tensor(float(self.steps)),
                            "exp_avg": self.exp_avg[off: off + n].view(shape).clone(),
                            "exp_avg_sq": self.exp_avg_sq[off: off + n].view(shape).clone()}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        """Inverse of state_dict(); also accepts torch.optim.AdamW(model.parameters()).state_dict().  Every parameter
        must carry the same step count (the fused update keeps one)."""
        groups = sd["param_groups"]
        views = self._param_views()
        if len(groups) != 1 or len(groups[0]["params"]) != len(views):
            raise ValueError("expected ONE parameter group over all of model.parameters()")
        g = groups[0]
        self.lr, self.betas, self.eps, self.weight_decay = g["lr"], tuple(g["betas"]), g["eps"], g["weight_decay"]
        if g.get("amsgrad") or g.get("maximize"):
            raise ValueError("amsgrad / maximize are not part of the fused AdamW update")
        state = sd["state"]
        if not state:
            self.steps = 0
            self.exp_avg.zero_()
            self.exp_avg_sq.zero_()
            return
        if sorted(state.keys()) != list(range(len(views))):
            raise ValueError("optimizer state must cover every parameter")
        steps = {int(float(state[i]["step"])) for i in state}
        if len(steps) != 1:
            raise ValueError(f"parameters carry different step counts: {sorted(steps)}")
        self.steps = steps.pop()
        for i, (name, off, n, shape) in enumerate(views):
            for key, flat in (("exp_avg", self.exp_avg), ("exp_avg_sq", self.exp_avg_sq)):
                t = state[i][key]
                if tuple(t.shape) != tuple(shape):
                    raise ValueError(f"{key} of {name}: shape {tuple(t.shape)}, expected {tuple(shape)}")
                flat[off: off + n].copy_(t.reshape(-1).to(flat.device, torch.float32))

    def _param_views(self):
        shapes = {name: tuple(p.shape) for name, p in self.model.named_parameters()}
        return [(name, off, n, shapes[name]) for name, off, n in self.state.layout]
