"""Serving shim with the I/O contract of the reference's exported classifier.

`detect.py:69-77,143-145` talks to the exported ONNX model through onnxruntime:

    inname = [i.name for i in self.classifier.get_inputs()]
    label_pred, heatmap_pred = self.classifier.run(None, {inname[0]: hand})      # hand: (1, 3, 192, 192) fp32

`ClassifierSession` offers the same two calls on top of the B200 MultiTaskNet, so `detect.py` can swap
`ort.InferenceSession(path, providers=...)` for `ClassifierSession(model)` without touching the call site
(`export.py:73-74` does not name the graph's inputs / outputs and `detect.py` reads the input name from
`get_inputs()` and unpacks the outputs by position, so the names here are free).  Any batch size is accepted
(one plan per batch size), so several hands / frames can be classified in one call.
"""
from __future__ import annotations

from collections import namedtuple

import numpy as np
import torch

from .model import MultiTaskNet

_Arg = namedtuple("NodeArg", ["name", "shape", "type"])


class ClassifierSession:
    def __init__(self, model: MultiTaskNet, input_name: str = "input", output_names=("label_pred", "heatmap_pred")):
        p = next(model.parameters())
        if not p.is_cuda:
            raise RuntimeError("ClassifierSession needs the model on a CUDA device; there is no CPU path")
        self.model = model.eval()
        self.device = p.device
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
        keep = self.model.return_attention
        self.model.return_attention = False
        try:
            with torch.no_grad():
                cls, hm, _ = self.model(torch.from_numpy(x).to(self.device, non_blocking=True))
        finally:
            self.model.return_attention = keep
        outs = {self._outputs[0].name: cls.cpu().numpy(), self._outputs[1].name: hm.cpu().numpy()}
        wanted = [o.name for o in self._outputs] if output_names is None else list(output_names)
        return [outs[n] for n in wanted]
