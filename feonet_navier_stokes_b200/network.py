"""Model zoo with the reference's public surface (class names, constructor arguments, output
shape `[B, 1, d_out]`, state_dict keys) -- `*/network.py:20-267` of the reference.

The networks are out of scope for the CUDA work (BASELINE.json: "the network's own layers stay
in PyTorch"); they are re-stated here so a training script can `from
feonet_navier_stokes_b200.network import *` without the reference tree.  The only addition is
`dof_major_head=True`: the final `nn.Linear` is evaluated as `W @ x^T` so the coefficient tensor
is born in the dof-major layout the residual kernels consume (same GEMM, same parameters, same
values -- it is a memory-format choice like channels_last; the gradient flows back through the
same cuBLAS calls).
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

__all__ = ["NetA", "Net2D", "Net3D", "FCNN", "ConvBNAct", "DoubleConv", "UNetFeatureExtractor", "UNetHead",
           "UNetWithHead", "VectorToSequenceRNN", "LinearT", "conv1d", "conv2d", "conv3d"]


class LinearT(nn.Linear):
    """nn.Linear whose [B, out] result is laid out dof-major (strides (1, ceil4(B)))."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:  # x: [B, in]
        if x.dim() != 2:
            return super().forward(x)
        B = x.shape[0]
        pad = (-B) % 4
        xt = (F.pad(x, (0, 0, 0, pad)) if pad else x).t()  # [in, ceil4(B)]
        yt = torch.addmm(self.bias.unsqueeze(1), self.weight, xt) if self.bias is not None else self.weight @ xt
        return yt[:, :B].t()


def _head(n_in: int, n_out: int, dof_major: bool) -> nn.Linear:
    return LinearT(n_in, n_out, bias=True) if dof_major else nn.Linear(n_in, n_out, bias=True)


def conv1d(in_planes, out_planes, stride=1, bias=True, kernel_size=5, padding=2, dialation=1):
    return nn.Conv1d(in_planes, out_planes, kernel_size=kernel_size, stride=stride, padding=padding, bias=bias)


def conv2d(in_planes, out_planes, stride=1, bias=True, kernel_size=5, padding=2, dialation=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=kernel_size, stride=stride, padding=padding, bias=bias)


def conv3d(in_planes, out_planes, stride=1, bias=True, kernel_size=5, padding=2, dialation=1):
    return nn.Conv3d(in_planes, out_planes, kernel_size=kernel_size, stride=stride, padding=padding, bias=bias)


class _ConvStack(nn.Module):
    """conv -> SiLU -> [conv, SiLU] x blocks -> conv -> flatten -> Linear -> [B,1,d_out]
    (shared body of NetA / Net2D / Net3D; attribute names follow the reference's state_dict)."""

    def __init__(self, make_conv, d_in, filters, d_out, fc_in, kernel_size, padding, blocks, dof_major_head):
        super().__init__()
        self.d_in, self.filters, self.d_out, self.blocks = d_in, filters, d_out, blocks
        self.kern, self.pad = kernel_size, padding
        self.swish = nn.SiLU()
        self.conv1 = make_conv(d_in, filters, kernel_size=kernel_size, padding=padding)
        mids: List[nn.Module] = []
        for _ in range(blocks):
            mids += [make_conv(filters, filters, kernel_size=kernel_size, padding=padding), self.swish]
        self.conv_list = nn.Sequential(*mids)
        self.convH = make_conv(filters, filters, kernel_size=kernel_size, padding=padding)
        self.fcH = _head(fc_in, d_out, dof_major_head)

    def forward(self, x):
        h = self.swish(self.conv1(x))
        if self.blocks != 0:
            h = self.conv_list(h)
        h = self.fcH(self.convH(h).flatten(start_dim=1))
        return h.unsqueeze(1) if h.stride(-1) != 1 else h.view(h.shape[0], 1, self.d_out)


class NetA(_ConvStack):
    def __init__(self, d_in, filters, d_out, kernel_size=7, padding=3, blocks=0, is_bdrylyaer=False, dof_major_head=False):
        fc_in = filters * (d_out - 1) if is_bdrylyaer else filters * d_out
        super().__init__(conv1d, d_in, filters, d_out, fc_in, kernel_size, padding, blocks, dof_major_head)


class Net2D(_ConvStack):
    def __init__(self, resol_in, d_in, filters, d_out, kernel_size=7, padding=3, blocks=0, dof_major_head=False):
        super().__init__(conv2d, d_in, filters, d_out, filters * resol_in ** 2, kernel_size, padding, blocks, dof_major_head)
        self.resol_in = resol_in


class Net3D(_ConvStack):
    def __init__(self, resol_in, d_in, filters, d_out, kernel_size=7, padding=3, blocks=0, dof_major_head=False):
        super().__init__(conv3d, d_in, filters, d_out, filters * resol_in ** 3, kernel_size, padding, blocks, dof_major_head)
        self.resol_in = resol_in


class FCNN(nn.Module):
    """Tanh MLP with dropout; output [B, output_dim] (the caller unsqueezes, steady NS train_FEONet.py:350)."""

    def __init__(self, resol_in, output_dim, hidden_dims: Sequence[int] = (2048, 1024, 512, 1024, 2048, 4096, 8192),
                 dropout_prob=0.2, dof_major_head=False):
        super().__init__()
        dims = [resol_in] + list(hidden_dims)
        layers: List[nn.Module] = []
        for a, b in zip(dims[:-1], dims[1:]):
            layers += [nn.Linear(a, b), nn.Tanh(), nn.Dropout(p=dropout_prob)]
        layers.append(_head(dims[-1], output_dim, dof_major_head))
        self.model = nn.Sequential(*layers)

    def forward(self, x):
        return self.model(x)


class ConvBNAct(nn.Module):
    def __init__(self, in_ch, out_ch, k=3, p=1):
        super().__init__()
        self.conv = nn.Conv2d(in_ch, out_ch, kernel_size=k, padding=p)
        self.bn = nn.BatchNorm2d(out_ch)
        self.act = nn.SiLU(inplace=True)

    def forward(self, x):
        return self.act(self.bn(self.conv(x)))


class DoubleConv(nn.Module):
    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.block = nn.Sequential(ConvBNAct(in_ch, out_ch, 3, 1), ConvBNAct(out_ch, out_ch, 3, 1))

    def forward(self, x):
        return self.block(x)


class UNetFeatureExtractor(nn.Module):
    """(B, in_ch, H, W) -> (B, latent_ch, H, W): two-level encoder/decoder with skip connections."""

    def __init__(self, in_ch=2, base_ch=32, latent_ch=16):
        super().__init__()
        c = base_ch
        self.enc1, self.pool1 = DoubleConv(in_ch, c), nn.MaxPool2d(2)
        self.enc2, self.pool2 = DoubleConv(c, 2 * c), nn.MaxPool2d(2)
        self.bottleneck = DoubleConv(2 * c, 4 * c)
        self.up2 = nn.ConvTranspose2d(4 * c, 2 * c, kernel_size=2, stride=2)
        self.dec2 = DoubleConv(4 * c, 2 * c)
        self.up1 = nn.ConvTranspose2d(2 * c, c, kernel_size=2, stride=2)
        self.dec1 = DoubleConv(2 * c, c)
        self.proj = nn.Conv2d(c, latent_ch, kernel_size=1)

    def forward(self, x):
        s1 = self.enc1(x)
        s2 = self.enc2(self.pool1(s1))
        h = self.bottleneck(self.pool2(s2))
        h = self.dec2(torch.cat([self.up2(h), s2], dim=1))
        h = self.dec1(torch.cat([self.up1(h), s1], dim=1))
        return self.proj(h)


class UNetHead(nn.Module):
    """(B, d_in, H, W) -> (B, 1, d_out)."""

    def __init__(self, resol_in: int, d_in: int, d_out: int, filters: int = 64, kernel_size: int = 7, padding: int = 3,
                 blocks: int = 1, dof_major_head: bool = False):
        super().__init__()
        self.act = nn.SiLU(inplace=True)
        self.conv1 = nn.Conv2d(d_in, filters, kernel_size=kernel_size, padding=padding)
        mids: List[nn.Module] = []
        for _ in range(blocks):
            mids += [nn.Conv2d(filters, filters, kernel_size=kernel_size, padding=padding), nn.SiLU(inplace=True)]
        self.mid = nn.Sequential(*mids)
        self.convH = nn.Conv2d(filters, filters, kernel_size=kernel_size, padding=padding)
        self.fc = _head(filters * resol_in ** 2, d_out, dof_major_head)

    def forward(self, x):
        h = self.act(self.conv1(x))
        if len(self.mid) > 0:
            h = self.mid(h)
        return self.fc(self.convH(h).flatten(start_dim=1)).unsqueeze(1)


class UNetWithHead(nn.Module):
    def __init__(self, resol_in: int, in_ch: int = 2, base_ch: int = 32, latent_ch: int = 16, d_out: int = 10,
                 head_filters: int = 64, head_blocks: int = 1, head_kernel_size: int = 7, head_padding: int = 3,
                 dof_major_head: bool = False):
        super().__init__()
        self.feature = UNetFeatureExtractor(in_ch=in_ch, base_ch=base_ch, latent_ch=latent_ch)
        self.head = UNetHead(resol_in=resol_in, d_in=latent_ch, d_out=d_out, filters=head_filters, blocks=head_blocks,
                             kernel_size=head_kernel_size, padding=head_padding, dof_major_head=dof_major_head)

    @torch.no_grad()
    def extract_latent(self, x):
        return self.feature(x)

    def forward(self, x):
        return self.head(self.feature(x))


class VectorToSequenceRNN(nn.Module):
    """Autoregressive GRU / LSTM decoder of the time-dependent variant (`FEONet_time_dep_Stokes/network.py:342-399`; same
    constructor arguments and state_dict keys `fc_init`, `rnn`, `fc_out`): the initial coefficient vector [B, input_dim] seeds
    the first layer's hidden state, a zero start token goes in, every output [B, 1, output_dim] is fed back as the next input;
    returns [B, seq_len, output_dim]."""

    def __init__(self, input_dim=1003, hidden_dim=512, output_dim=1003, rnn_type="gru", num_layers=1):
        super().__init__()
        self.input_dim, self.hidden_dim, self.output_dim, self.num_layers = input_dim, hidden_dim, output_dim, num_layers
        self.fc_init = nn.Linear(input_dim, hidden_dim)
        kinds = {"gru": nn.GRU, "lstm": nn.LSTM}
        if rnn_type.lower() not in kinds:
            raise ValueError("rnn_type must be 'gru' or 'lstm'")
        self.rnn = kinds[rnn_type.lower()](output_dim, hidden_dim, num_layers=num_layers, batch_first=True)
        self.fc_out = nn.Linear(hidden_dim, output_dim)

    def forward(self, x, seq_len):
        B = x.size(0)
        state = torch.cat([torch.tanh(self.fc_init(x)).unsqueeze(0),
                           torch.zeros(self.num_layers - 1, B, self.hidden_dim, device=x.device)], dim=0)
        if isinstance(self.rnn, nn.LSTM):
            state = (state, torch.zeros(self.num_layers, B, self.hidden_dim, device=x.device))
        token = torch.zeros(B, 1, self.output_dim, device=x.device)
        steps = []
        for _ in range(seq_len):
            out, state = self.rnn(token, state)
            token = self.fc_out(out)
            steps.append(token)
        return torch.cat(steps, dim=1)
