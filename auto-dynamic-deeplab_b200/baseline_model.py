"""The non-dense siblings of ADD — B200 drop-ins for `modeling/baseline_model.py::Baselin_Model` (:93-254) and
`modeling/autodeeplab.py::AutoDeepLab` (:94-204).  Same cells, stems, ASPP and decoder as ADD; every cell reads only
the two previous outputs (no dense links, no `dense_process`).  Everything — plans, CUDA graphs, kernels — is ADD's."""
from __future__ import annotations

from typing import List

import torch
import torch.nn as nn

from . import runtime as rt
from .ADD import ADD, Cell


class Cell_baseline(Cell):
    """baseline_model.py:14-90 — ADD's Cell without dense input / dense output; forward returns (prev_input, concat)."""

    def __init__(self, BatchNorm, B, prev_prev_C, prev_C, cell_arch, network_arch, C_out, downup_sample):
        super().__init__(BatchNorm, B, prev_prev_C, prev_C, cell_arch, network_arch, C_out, downup_sample, False, False)

    def forward(self, prev_prev_input, prev_input):
        return prev_input, super().forward(prev_prev_input, prev_input)


Cell_AutoDeepLab = Cell_baseline      # autodeeplab.py:15-91 is the same cell


class Baselin_Model(ADD):
    """baseline_model.py:93-254: multi-exit like ADD.forward (same aspp_size rule, conv_aspp adapters, decoder)."""
    DENSE = False

    def __init__(self, network_arch, C_index, cell_arch, num_classes, args, low_level_layer):
        super().__init__(network_arch, C_index, cell_arch, num_classes, args, low_level_layer)
        self.pooling = nn.MaxPool2d(3, stride=2)          # parameter-free members of the reference (baseline_model.py:189-191)
        self.gap = nn.AdaptiveAvgPool2d(1)
        self.relu = nn.ReLU()

    def get_feature(self, x):
        raise NotImplementedError("Baselin_Model has no get_feature / dynamic_inference in the reference")

    def dynamic_inference(self, *a, **k):
        raise NotImplementedError("Baselin_Model has no get_feature / dynamic_inference in the reference")

    dynamic_inference_batch = dynamic_evaluate = dynamic_inference


class AutoDeepLab(ADD):
    """autodeeplab.py:94-204: single exit after the last cell; forward returns (None, logits) like the reference.
    The reference feeds the last feature to ASPP as is (no resize-to-aspp_size step, autodeeplab.py:198-200)."""
    DENSE = False

    def __init__(self, network_arch, cell_arch, num_classes, args, low_level_layer):
        super().__init__(network_arch, [], cell_arch, num_classes, args, low_level_layer)
        self.num_model_layers = self.num_net
        self.model_network = self.network_arch

    def _emit_exit_lowres(self, b, y, st, i, aspp_size, conv_aspp_iter, resize=True, relu_feature=False):
        return super()._emit_exit_lowres(b, y, st, i, aspp_size, conv_aspp_iter, resize=False, relu_feature=relu_feature)

    def forward(self, x: torch.Tensor, iter_rate=1.0):
        return None, super().forward(x)[-1]

    def get_feature(self, x):
        raise NotImplementedError("AutoDeepLab has no get_feature / dynamic_inference in the reference")

    def dynamic_inference(self, *a, **k):
        raise NotImplementedError("AutoDeepLab has no get_feature / dynamic_inference in the reference")

    dynamic_inference_batch = dynamic_evaluate = dynamic_inference
