"""The two CNNs that produce the sampler's inputs, on PyTorch/cuDNN (not the optimisation target: BASELINE.json north_star), so
that the maps can be handed to the sampler in HBM without the reference's detour through `.npy` files on disk
(models/mpp/data_loaders.py:30-71 reads what models/position_net/pos_net_model.py:338-349 and
models/shape_net/shape_net_model.py:130-182 wrote).

Architectures and parameter names follow the reference so that its checkpoints load with `load_state_dict`:
  * U-Net backbone (model_parts/unet/unet.py:24-61, unet_parts.py): hidden_dims [32, 64, 128, 256]
    (model_configs/posnet/config_pos.json, model_configs/shapenet/config_shape.json), 3x3 reflect-padded conv + BatchNorm + ReLU
    twice per level, 2x2 max-pool down, 2x2 transposed conv up, skip concatenation;
  * position net (models/position_net/pos_net.py:9-31): backbone + 1x1 conv to 3 channels (pointing vector x, y, mask logit);
    detection map = sigmoid(1x1 conv(divergence(vector field) * mask logit))  (pos_net_model.py:75-79,338-349, torch_div.py);
  * shape net (models/shape_net/shape_net.py:12-54): backbone + three 1x1 heads of 32 classes (size, ratio, angle,
    shape_net_model.py:80-85), softmax over classes.

`MapProducer.produce(images)` returns the maps in the layout the sampler reads in place: det (B, H, W) float32 and marks
(B, 3, H, W, 32) float32, class index fastest."""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn.functional as F
from torch import nn

HIDDEN_DIMS = (32, 64, 128, 256)
N_CLASSES = 32


class _DoubleConv(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.double_conv = nn.Sequential(
            nn.Conv2d(cin, cout, 3, padding=1, padding_mode="reflect"), nn.BatchNorm2d(cout), nn.ReLU(inplace=True),
            nn.Conv2d(cout, cout, 3, padding=1, padding_mode="reflect"), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))

    def forward(self, x):
        return self.double_conv(x)


class _Down(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), _DoubleConv(cin, cout))

    def forward(self, x):
        return self.maxpool_conv(x)


class _Up(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.up = nn.ConvTranspose2d(cin, cin // 2, 2, stride=2)
        self.conv = _DoubleConv(cin, cout)

    def forward(self, x, skip):
        return self.conv(torch.cat([skip, self.up(x)], dim=1))


class UNetBackbone(nn.Module):
    def __init__(self, in_channels: int = 3, hidden_dims: Sequence[int] = HIDDEN_DIMS):
        super().__init__()
        self.descending_path = nn.ModuleList()
        c = in_channels
        for i, h in enumerate(hidden_dims):
            self.descending_path.append(_DoubleConv(c, h) if i == 0 else _Down(c, h))
            c = h
        self.ascending_path = nn.ModuleList()
        for h in list(hidden_dims)[::-1][1:]:
            self.ascending_path.append(_Up(c, h))
            c = h
        self.out_channels = c
        self.depth = len(hidden_dims) - 1

    def forward(self, x):
        skips = []
        for d in self.descending_path:
            x = d(x)
            skips.append(x)
        for u, s in zip(self.ascending_path, skips[::-1][1:]):
            x = u(x, s)
        return x


class PositionNet(nn.Module):
    def __init__(self, hidden_dims: Sequence[int] = HIDDEN_DIMS):
        super().__init__()
        self.backbone = UNetBackbone(3, hidden_dims)
        self.final_layer = nn.Conv2d(self.backbone.out_channels, 3, 1)

    def forward(self, x):
        return self.final_layer(self.backbone(x))


class ShapeNet(nn.Module):
    def __init__(self, hidden_dims: Sequence[int] = HIDDEN_DIMS, n_marks: int = 3, n_classes: int = N_CLASSES):
        super().__init__()
        self.backbone = UNetBackbone(3, hidden_dims)
        self.final_layers = nn.ModuleList([nn.Sequential(nn.Conv2d(self.backbone.out_channels, n_classes, 1)) for _ in range(n_marks)])

    def forward(self, x) -> List[torch.Tensor]:
        f = self.backbone(x)
        return [h(f) for h in self.final_layers]


def _divergence(vec: torch.Tensor) -> torch.Tensor:
    """torch_div.py:9-31 with indexing 'ij': d vec[:, 0] / d row + d vec[:, 1] / d col (central differences, one-sided at edges)."""
    return torch.gradient(vec[:, 0], dim=1)[0] + torch.gradient(vec[:, 1], dim=2)[0]


class MapProducer(nn.Module):
    """Position net + divergence classifier + shape net -> (det, marks) of a batch of images, everything on the device."""

    def __init__(self, hidden_dims: Sequence[int] = HIDDEN_DIMS):
        super().__init__()
        self.posnet = PositionNet(hidden_dims)
        self.div_clf = nn.Conv2d(1, 1, 1)   # pos_net_model.py:75-79: Sequential(Divergence, Conv2d(1, 1, 1)); index 1 of that Sequential
        self.shapenet = ShapeNet(hidden_dims)

    @torch.no_grad()
    def produce(self, images: torch.Tensor):
        """images (B, 3, H, W) float -> det (B, H, W) f32, marks (B, 3, H, W, 32) f32 (softmax over the last axis)."""
        self.eval()
        b, _, h, w = images.shape
        div = 2 ** self.posnet.backbone.depth
        ph, pw = (-h) % div, (-w) % div  # unet.py:9-21 pad_before_infer
        x = F.pad(images.float(), (0, pw, 0, ph)) if (ph or pw) else images.float()
        x = x.contiguous(memory_format=torch.channels_last)
        out = self.posnet(x)[:, :, :h, :w]
        score = self.div_clf((_divergence(out[:, :2]) * out[:, 2]).unsqueeze(1))
        det = torch.sigmoid(score).squeeze(1).contiguous()
        marks = torch.empty((b, 3, h, w, N_CLASSES), dtype=torch.float32, device=images.device)
        for i, logits in enumerate(self.shapenet(x)):
            # softmax over the class channel, written class-fastest: (B, 32, H, W) -> (B, H, W, 32)
            marks[:, i] = torch.softmax(logits[:, :, :h, :w].float(), dim=1).permute(0, 2, 3, 1)
        return det, marks
