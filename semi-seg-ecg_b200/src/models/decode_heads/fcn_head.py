"""FCN decode head: same constructor / state_dict keys as the reference
(src/models/decode_heads/fcn_head.py:9-97); compute runs through libsemiseg_b200.

Covered configuration (every shipped YAML): num_convs=1, kernel_size=3, dilation=1,
concat_input=False, BatchNorm1d + ReLU.  Default torch init, like the reference (no custom reset).
"""
import torch.nn as nn


class FCNHead(nn.Module):
    def __init__(self, in_channels: int, channels: int, num_classes: int, num_convs: int, kernel_size: int = 3,
                 concat_input: bool = True, dilation: int = 1, in_index: int = -1, dropout_ratio: float = 0.1,
                 align_corners: bool = False, norm_layer=nn.BatchNorm1d, act_layer=nn.ReLU):
        super().__init__()
        assert num_convs >= 0 and dilation > 0
        bad = []
        if num_convs != 1:
            bad.append(f"num_convs={num_convs}")
        if kernel_size != 3:
            bad.append(f"kernel_size={kernel_size}")
        if concat_input:
            bad.append("concat_input=True")
        if dilation != 1:
            bad.append(f"dilation={dilation}")
        if norm_layer is not nn.BatchNorm1d or act_layer is not nn.ReLU:
            bad.append("norm/act other than BatchNorm1d/ReLU")
        if bad:
            raise NotImplementedError("FCNHead variant outside the accelerated hot path: " + ", ".join(bad))
        self.in_channels, self.channels = in_channels, channels
        self.num_classes, self.in_index = num_classes, in_index
        self.align_corners, self.num_convs = align_corners, num_convs
        self.concat_input, self.kernel_size = concat_input, kernel_size
        self.dropout_ratio = float(dropout_ratio)
        self.convs = nn.Sequential(nn.Sequential(
            nn.Conv1d(in_channels, channels, kernel_size, padding=kernel_size // 2, bias=False),
            nn.BatchNorm1d(channels), nn.ReLU(inplace=True)))
        self.cls_seg = nn.Conv1d(channels, num_classes, 1)
        self.dropout = nn.Dropout(dropout_ratio) if dropout_ratio > 0 else None

    def forward(self, inputs):
        raise RuntimeError("FCNHead is a parameter container here; run the model through EncoderDecoder")
