"""Layer-shape tables of the BASELINE.json configurations (SURVEY.md section 8 /
Appendix B), expressed as bare ``nn.Conv2d`` / ``nn.Linear`` stacks.

``preconditioner.step()`` only sees, per preconditioned layer, its input
``a``, its output gradient ``g`` and the parameter gradients -- so a workload
is the ordered list of layers plus the input spatial size / sequence length
of each; the rest of the network (BatchNorm, ReLU, pooling, attention) never
enters the hot path.  The tables are generated from the architectures'
definitions (ResNet-50 v1.5 bottlenecks [3,4,6,3]; VGG-16 CIFAR; BERT-base QA)
and cross-checked against the reference's own record
(scripts/dp_block_partition.py:92-93) in tests/test_workloads.py.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
import torch.nn as nn


@dataclass
class LayerShape:
    module: nn.Module
    hw: Optional[Tuple[int, int]] = None      # conv input H, W
    seq: Optional[int] = None                 # linear: sequence length of (B,T,D) inputs


class LayerStack(nn.Module):
    """Container whose ``modules()`` order is the layer order."""
    def __init__(self, layers: List[nn.Module]):
        super().__init__()
        self.layers = nn.ModuleList(layers)


def resnet50_layers(image: int = 224, num_classes: int = 1000) -> List[LayerShape]:
    """53 bias-free convs + fc of ResNet-50 v1.5 (stride on the 3x3 conv), in
    ``model.modules()`` order: conv1; per bottleneck conv1, conv2, conv3 and, in the
    first block of each stage, the downsample conv (registered after conv3)."""
    out: List[LayerShape] = []
    h = image
    out.append(LayerShape(nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False), (h, h)))
    h = (h + 6 - 7) // 2 + 1          # 112
    h = (h + 2 - 3) // 2 + 1          # maxpool -> 56
    inplanes = 64
    for planes, blocks, stride in ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)):
        for b in range(blocks):
            s = stride if b == 0 else 1
            out.append(LayerShape(nn.Conv2d(inplanes, planes, 1, bias=False), (h, h)))
            out.append(LayerShape(nn.Conv2d(planes, planes, 3, stride=s, padding=1, bias=False), (h, h)))
            h2 = (h + 2 - 3) // s + 1
            out.append(LayerShape(nn.Conv2d(planes, planes * 4, 1, bias=False), (h2, h2)))
            if b == 0:
                out.append(LayerShape(nn.Conv2d(inplanes, planes * 4, 1, stride=s, bias=False), (h, h)))
            inplanes = planes * 4
            h = h2
    out.append(LayerShape(nn.Linear(512 * 4, num_classes)))
    return out


def vgg16_cifar_layers(image: int = 32, num_classes: int = 10) -> List[LayerShape]:
    cfg = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512, 'M']
    out, c, h = [], 3, image
    for x in cfg:
        if x == 'M':
            h //= 2
        else:
            out.append(LayerShape(nn.Conv2d(c, x, 3, padding=1), (h, h)))
            c = x
    out.append(LayerShape(nn.Linear(512, num_classes)))
    return out


def bert_base_qa_layers(seq: int = 384, hidden: int = 768, inter: int = 3072,
                        blocks: int = 12) -> List[LayerShape]:
    out = []
    for _ in range(blocks):
        for _ in range(4):                                    # q, k, v, attention output dense
            out.append(LayerShape(nn.Linear(hidden, hidden), seq=seq))
        out.append(LayerShape(nn.Linear(hidden, inter), seq=seq))
        out.append(LayerShape(nn.Linear(inter, hidden), seq=seq))
    out.append(LayerShape(nn.Linear(hidden, 2), seq=seq))     # qa_outputs
    return out


def mlp_layers() -> List[LayerShape]:
    return [LayerShape(nn.Linear(784, 512)), LayerShape(nn.Linear(512, 10))]


WORKLOADS = {
    # name: (layer table, variant, per-GPU batch, hyper-parameters from the reference's train_*.sh)
    'resnet50': (resnet50_layers, 'eigen_dp', 32, dict(lr=0.0125, damping=0.002, kl_clip=0.001)),
    'vgg16': (vgg16_cifar_layers, 'inverse_dp', 128, dict(lr=0.1, damping=0.03, kl_clip=0.001)),
    'bert_base': (bert_base_qa_layers, 'eigen_dp', 8, dict(lr=3e-5, damping=0.03, kl_clip=None)),
    'mlp': (mlp_layers, 'eigen_dp', 32, dict(lr=0.1, damping=0.003, kl_clip=0.001)),
}


def layer_dims(ls: LayerShape) -> Tuple[int, int]:
    m = ls.module
    if isinstance(m, nn.Conv2d):
        dA = m.in_channels * m.kernel_size[0] * m.kernel_size[1]
        dG = m.out_channels
    else:
        dA, dG = m.in_features, m.out_features
    return dA + (1 if m.bias is not None else 0), dG


def io_shapes(ls: LayerShape, batch: int):
    """Shapes of the layer input ``a`` and of the output gradient ``g``."""
    m = ls.module
    if isinstance(m, nn.Conv2d):
        h, w = ls.hw
        oh = (h + 2 * m.padding[0] - m.kernel_size[0]) // m.stride[0] + 1
        ow = (w + 2 * m.padding[1] - m.kernel_size[1]) // m.stride[1] + 1
        return (batch, m.in_channels, h, w), (batch, m.out_channels, oh, ow)
    if ls.seq is not None:
        return (batch, ls.seq, m.in_features), (batch, ls.seq, m.out_features)
    return (batch, m.in_features), (batch, m.out_features)


def rows_of(ls: LayerShape, batch: int) -> int:
    m = ls.module
    if isinstance(m, nn.Conv2d):
        _, gshape = io_shapes(ls, batch)
        return batch * gshape[2] * gshape[3]
    return batch


def algorithmic_work(layers: List[LayerShape], batch: int, variant: str):
    """SURVEY.md section 8(d) conventions, per step, summed over ``layers``."""
    F_factor = B_factor = F_inv = B_inv = F_apply = B_apply = B_final = 0
    for ls in layers:
        dA, dG = layer_dims(ls)
        rows = rows_of(ls, batch)
        ashape, gshape = io_shapes(ls, batch)
        na = 1
        for s in ashape:
            na *= s
        ng = 1
        for s in gshape:
            ng *= s
        F_factor += 2 * rows * (dA * dA + dG * dG)
        B_factor += 4 * (na + ng) + 2 * 4 * (dA * dA + dG * dG)
        if variant == 'eigen_dp':
            F_inv += 9 * (dA ** 3 + dG ** 3)
            F_apply += 4 * (dG * dG * dA + dG * dA * dA)
            B_apply += 4 * 2 * (dA * dA + dG * dG) + 3 * 4 * dG * dA
        else:
            F_inv += dA ** 3 + dG ** 3
            F_apply += 2 * (dG * dG * dA + dG * dA * dA)
            B_apply += 4 * (dA * dA + dG * dG) + 3 * 4 * dG * dA
        B_inv += 2 * 4 * (dA * dA + dG * dG)
        B_final += 2 * 4 * dG * dA
    return dict(F_factor=F_factor, B_factor=B_factor, F_inverse=F_inv, B_inverse=B_inv,
                F_apply=F_apply, B_apply=B_apply, B_finalize=B_final)
