import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import splade_b200
dev = torch.device("cuda:0"); V = 50000
g = torch.Generator().manual_seed(0)
x = torch.log1p(torch.relu(torch.randn(148, V, generator=g).to(torch.bfloat16).float())).to(dev)
for _ in range(3):
    splade_b200.topk(x, 64, None)
torch.cuda.synchronize()
print("done")
