"""1-D ResNet backbone: same constructor names / kwargs / state_dict keys as the reference
(src/models/backbones/resnet.py:135-204, 366-428), re-hosted on the libsemiseg_b200 kernels.

The torch modules created here (`nn.Conv1d`, `nn.BatchNorm1d`) are PARAMETER CONTAINERS: they
give the reference's parameter names, shapes, registration order and initial-value stream
(same RNG consumption as the reference constructors, so `torch.manual_seed(s)` reproduces the
reference's random init bit-for-bit).  Their own `forward` is never used: compute is issued by
`semiseg_b200.net.NetPlan` through the C ABI, and there is no CPU fallback.
"""
import math
from typing import Optional, Sequence

import torch.nn as nn

__all__ = ["ResNet", "resnet18", "resnet34", "resnet50", "resnet101", "resnet152"]

_BASIC_LAYOUTS = {"resnet18": (2, 2, 2, 2), "resnet34": (3, 4, 6, 3)}
_BOTTLENECK_LAYOUTS = {"resnet50": (3, 4, 6, 3), "resnet101": (3, 4, 23, 3), "resnet152": (3, 8, 36, 3)}


class BasicBlock(nn.Module):
    """conv3(s)-BN-ReLU-conv3-BN (+ 1x1(s)-BN shortcut) -add-ReLU  (reference resnet.py:19-72)."""
    expansion = 1

    def __init__(self, inplanes: int, planes: int, stride: int, with_shortcut_conv: bool):
        super().__init__()
        self.conv1 = nn.Conv1d(inplanes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm1d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv1d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm1d(planes)
        self.downsample = None
        if with_shortcut_conv:
            self.downsample = nn.Sequential(nn.Conv1d(inplanes, planes, 1, stride=stride, bias=False),
                                            nn.BatchNorm1d(planes))
        self.stride = stride

    def forward(self, x):
        raise RuntimeError("BasicBlock is a parameter container; run the model through EncoderDecoder")


class Bottleneck(nn.Module):
    """1x1-BN-ReLU - conv3(s)-BN-ReLU - 1x1(x4)-BN (+ 1x1(s)-BN shortcut) -add-ReLU  (reference resnet.py:75-132)."""
    expansion = 4

    def __init__(self, inplanes: int, planes: int, stride: int, with_shortcut_conv: bool):
        super().__init__()
        self.conv1 = nn.Conv1d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm1d(planes)
        self.conv2 = nn.Conv1d(planes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm1d(planes)
        self.conv3 = nn.Conv1d(planes, planes * self.expansion, 1, bias=False)
        self.bn3 = nn.BatchNorm1d(planes * self.expansion)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = None
        if with_shortcut_conv:
            self.downsample = nn.Sequential(nn.Conv1d(inplanes, planes * self.expansion, 1, stride=stride, bias=False),
                                            nn.BatchNorm1d(planes * self.expansion))
        self.stride = stride

    def forward(self, x):
        raise RuntimeError("Bottleneck is a parameter container; run the model through EncoderDecoder")


class ResNet(nn.Module):
    def __init__(self, num_leads: int, stem_channels: int = 64, base_channels: int = 64, num_stages: int = 4,
                 strides: Sequence[int] = (1, 2, 2, 2), dilations: Sequence[int] = (1, 1, 1, 1),
                 deep_stem: bool = False, avg_down: bool = False, frozen_stages: int = -1, norm_layer=nn.BatchNorm1d,
                 multi_grid: Optional[Sequence[int]] = None, contract_dilation: bool = False, block=BasicBlock,
                 stage_blocks: Sequence[int] = (2, 2, 2, 2), zero_init_residual: bool = False,
                 out_indices: Sequence[int] = (0, 1, 2, 3)):
        super().__init__()
        assert 1 <= num_stages <= 4, "num_stages should be in [1, 4]"
        assert len(strides) == len(dilations) == num_stages, \
            f"strides and dilations should have num_stages={num_stages} entries, got {len(strides)}, {len(dilations)}"
        unsupported = []
        if deep_stem:
            unsupported.append("deep_stem=True")
        if avg_down:
            unsupported.append("avg_down=True")
        if any(d != 1 for d in dilations) or multi_grid is not None:
            unsupported.append("dilation != 1 / multi_grid")
        if block not in (BasicBlock, Bottleneck):
            unsupported.append(f"block {block!r}")
        if norm_layer is not nn.BatchNorm1d:
            unsupported.append("norm_layer other than BatchNorm1d")
        if frozen_stages >= 0:
            unsupported.append("frozen_stages")
        if any(s not in (1, 2) for s in strides):
            unsupported.append("stride other than 1 or 2")
        if unsupported:
            raise NotImplementedError(
                "ResNet variant outside the accelerated hot path: " + ", ".join(unsupported) +
                " (kernels cover k in {1,3,7}, stride in {1,2}, dilation 1, BasicBlock / Bottleneck; SURVEY.md section 2)")
        self.num_leads = num_leads
        self.stem_channels = stem_channels
        self.base_channels = base_channels
        self.num_stages = num_stages
        self.strides = tuple(strides)
        self.dilations = tuple(dilations)
        self.stage_blocks = tuple(stage_blocks[:num_stages])
        self.out_indices = tuple(out_indices)
        self.zero_init_residual = zero_init_residual
        self.block = block

        self.stem = nn.Sequential(nn.Conv1d(num_leads, stem_channels, 7, stride=2, padding=3, bias=False),
                                  nn.BatchNorm1d(stem_channels), nn.ReLU(inplace=True))
        self.maxpool = nn.MaxPool1d(kernel_size=3, stride=2, padding=1)
        self.res_layers = []
        width_in = stem_channels
        for i, depth in enumerate(self.stage_blocks):
            width = base_channels * 2 ** i
            blocks = []
            for j in range(depth):
                s = self.strides[i] if j == 0 else 1
                cin = width_in if j == 0 else width * block.expansion
                # (the shortcut module is created before the blocks of the stage, like the reference's _make_res_layer,
                #  resnet.py:267-293 -- but registered under the first block, so the random-init stream is consumed in
                #  module-traversal order either way: _init_weights walks self.modules())
                blocks.append(block(cin, width, s, with_shortcut_conv=(j == 0 and (s != 1 or cin != width * block.expansion))))
            name = f"layer{i + 1}"
            self.add_module(name, nn.Sequential(*blocks))
            self.res_layers.append(name)
            width_in = width * block.expansion
        self.feat_dim = block.expansion * base_channels * 2 ** (len(self.stage_blocks) - 1)
        self._init_weights()

    def _init_weights(self):
        # He-normal with n = k * C_out; BN gamma=1, beta=0 (reference resnet.py:326-339)
        for m in self.modules():
            if isinstance(m, nn.Conv1d):
                m.weight.data.normal_(0, math.sqrt(2.0 / (m.kernel_size[0] * m.out_channels)))
            elif isinstance(m, nn.BatchNorm1d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
        if self.zero_init_residual:
            for m in self.modules():
                if isinstance(m, Bottleneck):
                    nn.init.constant_(m.bn3.weight, 0)
                elif isinstance(m, BasicBlock):
                    nn.init.constant_(m.bn2.weight, 0)

    def no_weight_decay(self):
        return set()

    def forward(self, x):
        raise RuntimeError("ResNet is a parameter container here; wrap it in models.encoder_decoder.EncoderDecoder "
                           "(compute runs through libsemiseg_b200, there is no stand-alone torch forward)")


def _make(name, num_leads, **kwargs):
    if name in _BOTTLENECK_LAYOUTS:
        return ResNet(num_leads=num_leads, block=Bottleneck, stage_blocks=list(_BOTTLENECK_LAYOUTS[name]), **kwargs)
    return ResNet(num_leads=num_leads, block=BasicBlock, stage_blocks=list(_BASIC_LAYOUTS[name]), **kwargs)


def resnet18(num_leads: int, **kwargs):
    return _make("resnet18", num_leads, **kwargs)


def resnet34(num_leads: int, **kwargs):
    return _make("resnet34", num_leads, **kwargs)


def resnet50(num_leads: int, **kwargs):
    return _make("resnet50", num_leads, **kwargs)


def resnet101(num_leads: int, **kwargs):
    return _make("resnet101", num_leads, **kwargs)


def resnet152(num_leads: int, **kwargs):
    return _make("resnet152", num_leads, **kwargs)
