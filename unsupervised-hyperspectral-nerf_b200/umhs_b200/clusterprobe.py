"""Cluster probe: cosine similarity of spectra against the endmembers -> class probabilities.

Mirrors ``ClusterLookup`` of ref umhsnerf/utils/clusterprobe.py:5-38 (same signature and return values).
UMHSModel gets the same quantities from the fused compositing epilogue; this module is the standalone form.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import Tensor, nn


class ClusterLookup(nn.Module):
    def __init__(self, dim: int, n_classes: int):
        super().__init__()
        self.dim, self.n_classes = dim, n_classes

    def forward(self, x: Tensor, alpha, log_probs: bool = False, clusters: Tensor = None):
        inner = F.normalize(x, dim=1) @ F.normalize(clusters, dim=1).t()
        if log_probs:
            return F.log_softmax(inner * alpha, dim=1)
        if alpha is None:
            probs = F.one_hot(inner.argmax(dim=1), clusters.shape[0]).to(torch.float32)
        else:
            probs = F.softmax(inner * alpha, dim=1)
        return inner, probs
