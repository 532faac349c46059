import sys, os
sys.path.insert(0, os.getcwd())
import torch
from deepsphere_weather_b200 import functional as F_, graphs as G
nside, B, F, K = [int(a) for a in sys.argv[1:5]]
dev = torch.device("cuda:0")
lap = G.healpix_laplacian(nside).to(dev)
plan = F_.plan_for(lap)
x = torch.randn(B, lap.shape[0], F, device=dev)
for _ in range(3):
    out = F_.cheb_terms(x, plan, K)
torch.cuda.synchronize()
