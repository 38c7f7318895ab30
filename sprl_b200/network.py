"""The reference's policy/value network, for building random-init evaluators.

The worker of the reference never defines a network: it loads whatever TorchScript
file the controller traced (cpp/src/networks/GridNetwork.hpp:44).  To exercise that
path without the controller (benchmarks use synthetic random-init weights) this
module restates the architecture the controller traces,
`BasicGridNetwork(num_rows, num_cols, action_size, history_size, num_blocks,
num_channels)` of /root/reference/src/networks/grid_networks.py:30-80, with the
same submodule names, so state_dicts are interchangeable with the reference's, and
traces it the way scripts/othello_controller.py:237-239 does.
"""
import torch
from torch import nn
import torch.nn.functional as F


class ResidualBlock(nn.Module):                      # grid_networks.py:8-27
    def __init__(self, channels):
        super().__init__()
        self.conv1 = nn.Conv2d(channels, channels, 3, 1, 1)
        self.bn1 = nn.BatchNorm2d(channels)
        self.conv2 = nn.Conv2d(channels, channels, 3, 1, 1)
        self.bn2 = nn.BatchNorm2d(channels)

    def forward(self, x):
        residual = x
        x = F.relu(self.bn1(self.conv1(x)))
        return F.relu(self.bn2(self.conv2(x)) + residual)


class BasicGridNetwork(nn.Module):                   # grid_networks.py:30-80
    def __init__(self, num_rows, num_cols, action_size, history_size, num_blocks, num_channels,
                 num_policy_channels=2, num_value_channels=1):
        super().__init__()
        self.conv = nn.Conv2d(2 * history_size + 1, num_channels, 3, 1, 1)
        self.bn = nn.BatchNorm2d(num_channels)
        self.residual_blocks = nn.ModuleList([ResidualBlock(num_channels) for _ in range(num_blocks)])
        self.policy_conv = nn.Conv2d(num_channels, num_policy_channels, 1)
        self.policy_size = num_policy_channels * num_rows * num_cols
        self.policy_fc = nn.Linear(self.policy_size, action_size)
        self.value_conv = nn.Conv2d(num_channels, num_value_channels, 1)
        self.value_size = num_value_channels * num_rows * num_cols
        self.value_fc1 = nn.Linear(self.value_size, num_channels)
        self.value_fc2 = nn.Linear(num_channels, 1)

    def forward(self, x):
        x = F.relu(self.bn(self.conv(x)))
        for block in self.residual_blocks:
            x = block(x)
        policy = F.relu(self.policy_conv(x)).view(-1, self.policy_size)
        policy = self.policy_fc(policy)
        value = F.relu(self.value_conv(x)).view(-1, self.value_size)
        value = F.relu(self.value_fc1(value))
        value = torch.tanh(self.value_fc2(value))
        return policy, value


# (rows, cols, actions, history, blocks, channels): scripts/{othello,connect_four,go}_controller.py
REFERENCE_SHAPES = {
    "othello": (8, 8, 65, 1, 2, 64),
    "c4": (6, 7, 7, 1, 2, 64),
    "go7": (7, 7, 50, 8, 6, 64),
    "go9": (9, 9, 82, 8, 6, 64),
}


def make_network(kind="othello", seed=0):
    rows, cols, actions, history, blocks, channels = REFERENCE_SHAPES[kind]
    torch.manual_seed(seed)
    return BasicGridNetwork(rows, cols, actions, history, blocks, channels).eval()


def trace_network(net, device="cpu"):
    """torch.jit.trace on a randn(1, C, R, C') example in eval mode (src/interface/tracer.py:9-19)."""
    net = net.to(device).eval()
    c = net.conv.in_channels
    rows = net.value_size // net.value_conv.out_channels
    # rows*cols is all the heads know; recover the board from the policy head
    shape = {v[0] * v[1]: (v[0], v[1]) for v in REFERENCE_SHAPES.values()}.get(rows)
    example = torch.randn(1, c, *shape, device=device)
    with torch.no_grad():
        return torch.jit.trace(net, example)


def num_parameters(net):
    return sum(p.numel() for p in net.parameters())
