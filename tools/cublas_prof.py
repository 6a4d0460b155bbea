import torch
n=4096
A=torch.rand((n,n),dtype=torch.float64,device='cuda'); B=torch.rand((n,n),dtype=torch.float64,device='cuda')
for _ in range(3): C=A@B
torch.cuda.synchronize()
