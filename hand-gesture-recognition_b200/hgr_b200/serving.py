"""Serving shim with the I/O contract of the reference's exported classifier.

`detect.py:69-77,143-145` talks to the exported ONNX model through onnxruntime:

    inname = [i.name for i in self.classifier.get_inputs()]
    label_pred, heatmap_pred = self.classifier.run(None, {inname[0]: hand})      # hand: (1, 3, 192, 192) fp32

`ClassifierSession` offers the same two calls on top of the B200 MultiTaskNet, so `detect.py` can swap
`ort.InferenceSession(path, providers=...)` for `ClassifierSession(model)` without touching the call site
(`export.py:73-74` does not name the graph's inputs / outputs and `detect.py` reads the input name from
`get_inputs()` and unpacks the outputs by position, so the names here are free).  Any batch size is accepted
(one plan per batch size), so several hands / frames can be classified in one call.

Latency mode: `detect.py` classifies one or two hand crops per camera frame, where the forward is bound by its 38
kernel launches and the Python in front of them, not by the GPU.  Batches up to `graph_max_batch` are therefore
captured once into a CUDA graph over static device buffers and replayed (`cuda_graph=True`, the default): a `run()` is
then one H2D copy, one graph launch and two D2H copies (tools/serving_latency.py measures both modes).  A captured
graph holds raw pointers into the plan's workspace and the packed weights of the moment it was captured; the session
keeps both alive next to the graph and re-captures by itself when the model's parameters have changed since
(`load_state_dict`, an optimiser step, `.to()`), so a replay can neither read freed memory nor stale weights.
"""
from __future__ import annotations

from collections import namedtuple

import numpy as np
import torch

from .model import MultiTaskNet

_Arg = namedtuple("NodeArg", ["name", "shape", "type"])


class ClassifierSession:
    def __init__(self, model: MultiTaskNet, input_name: str = "input", output_names=("label_pred", "heatmap_pred"),
                 cuda_graph: bool = True, graph_max_batch: int = 64):
        p = next(model.parameters())
        if not p.is_cuda:
            raise RuntimeError("ClassifierSession needs the model on a CUDA device; there is no CPU path")
        self.model = model.eval()
        self.device = p.device
        self.cuda_graph, self.graph_max_batch = bool(cuda_graph), int(graph_max_batch)
        # batch -> (graph, static input, static logits, static heatmaps, pinned host copies, plan, packed weights):
        # the last two are what the graph's kernels point into, kept here so that they outlive the model's own cache
        self._graphs = {}
        s = model.image_size[0]
        self._inputs = [_Arg(input_name, ["batch", 3, s, s], "tensor(float)")]
        self._outputs = [_Arg(output_names[0], ["batch", model.num_classes], "tensor(float)"),
                         _Arg(output_names[1], ["batch", model.num_joints, s // 4, s // 4], "tensor(float)")]

    def get_inputs(self):
        return list(self._inputs)

    def get_outputs(self):
        return list(self._outputs)

    def run(self, output_names, input_feed):
        """onnxruntime's InferenceSession.run: numpy in, list of numpy arrays out (label_pred, heatmap_pred)."""
        name = self._inputs[0].name
        if name not in input_feed:
            raise ValueError(f"missing input '{name}'")
        x = np.ascontiguousarray(input_feed[name], dtype=np.float32)
        s = self.model.image_size[0]
        if x.ndim != 4 or tuple(x.shape[1:]) != (3, s, s):
            raise ValueError(f"input {x.shape} does not match (batch, 3, {s}, {s})")
        if self.cuda_graph and 0 < x.shape[0] <= self.graph_max_batch:
            cls, hm = self._run_graph(x)
        else:
            cls, hm = self._run_eager(x)
        outs = {self._outputs[0].name: cls, self._outputs[1].name: hm}
        wanted = [o.name for o in self._outputs] if output_names is None else list(output_names)
        return [outs[n] for n in wanted]

    def invalidate(self):
        """Drop the captured graphs (done automatically when the model's parameters change)."""
        self._graphs.clear()

    def _forward(self, x):
        keep = self.model.return_attention
        self.model.return_attention = False
        try:
            with torch.no_grad():
                cls, hm, _ = self.model(x)
        finally:
            self.model.return_attention = keep
        return cls, hm

    def _run_eager(self, x):
        cls, hm = self._forward(torch.from_numpy(x).to(self.device, non_blocking=True))
        return cls.cpu().numpy(), hm.cpu().numpy()

    def _capture(self, b):
        s = self.model.image_size[0]
        with torch.cuda.device(self.device):
            xs = torch.zeros(b, 3, s, s, dtype=torch.float32, device=self.device)
            self._forward(xs)  # builds the plan and packs the weights outside the capture
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                cls, hm = self._forward(xs)
            h_x = torch.empty(xs.shape, dtype=torch.float32).pin_memory()
            h_cls = torch.empty(cls.shape, dtype=torch.float32).pin_memory()
            h_hm = torch.empty(hm.shape, dtype=torch.float32).pin_memory()
            plan = self.model.plan_for(b, self.device)
        return graph, xs, cls, hm, h_x, h_cls, h_hm, plan, plan.params

    def _run_graph(self, x):
        b = x.shape[0]
        # the model re-packs (and drops its plans) when a parameter's storage or version changed: a graph captured
        # against the previous pack is stale then, whatever batch size it was captured for
        params = self.model._packed_params(self.device)
        entry = self._graphs.get(b)
        if entry is None or entry[-1] is not params:
            if entry is not None:
                self._graphs.clear()
            entry = self._graphs[b] = self._capture(b)
        graph, xs, cls, hm, h_x, h_cls, h_hm = entry[:7]
        with torch.cuda.device(self.device):
            h_x.copy_(torch.from_numpy(x))
            xs.copy_(h_x, non_blocking=True)
            graph.replay()
            h_cls.copy_(cls, non_blocking=True)
            h_hm.copy_(hm, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
        return h_cls.numpy().copy(), h_hm.numpy().copy()
