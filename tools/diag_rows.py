import sys, torch
sys.path.insert(0, '.')
from sprl_b200.evalnet import EvalNet
from sprl_b200.network import make_network
net = make_network("othello", 1)
ev = EvalNet(net, device=0)
g = torch.Generator().manual_seed(3)
x = (torch.rand(1501, 3, 8, 8, generator=g) > 0.5).float().cuda()
perm = torch.randperm(x.shape[0], generator=g).cuda()
l0, v0 = ev(x); l0b, v0b = ev(x)
print("same input twice: logits equal", torch.equal(l0, l0b), "values equal", torch.equal(v0, v0b))
l1, v1 = ev(x[perm])
dl = (l0[perm] - l1).abs(); dv = (v0[perm] - v1).abs()
print("perm: logit mismatching rows", int((dl.amax(1) > 0).sum()), "max", dl.max().item(), "| value rows", int((dv.reshape(-1) > 0).sum()), dv.max().item())
l2, v2 = ev(x[:77])
print("prefix 77: logits equal", torch.equal(l0[:77], l2), "values equal", torch.equal(v0[:77], v2), (l0[:77]-l2).abs().max().item())
# swap partners only: exchange rows 1 and 3 (tile (0,1)->(0,3), (2,3)->(2,1))
y = x.clone(); y[[1, 3]] = x[[3, 1]]
l3, v3 = ev(y)
print("row 0 after partner swap equal:", torch.equal(l3[0], l0[0]), "row 1<->3:", torch.equal(l3[1], l0[3]), torch.equal(l3[3], l0[1]))
# same board in slot 0 vs slot 1 of a tile
z = x.clone(); z[[0, 1]] = x[[1, 0]]
l4, v4 = ev(z)
print("board moved from half 0 to half 1 equal:", torch.equal(l4[1], l0[0]), (l4[1]-l0[0]).abs().max().item())
