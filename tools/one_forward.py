"""One dense POPCORN forward on a 2048^2 tile (for ncu launch lists / captures).  Development tool."""
import os, sys, warnings
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import popcorn_b200 as pb
from oracle import popcorn_oracle as po
H = W = int(os.environ.get("KB_SIZE", 2048))
sd = po.random_state_dict()
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    m = pb.POPCORN(6, occupancymodel=True, pretrained=False, biasinit=0.9407, sentinelbuildings=True, device="cuda")
m.load_state_dict(sd); m.eval()
x = torch.randn(1, 6, H, W, device="cuda")
with torch.no_grad():
    for _ in range(int(os.environ.get("KB_ITERS", 2))):
        out = m({"input": x}, padding=False)
torch.cuda.synchronize()
print(float(out["popcount"][0]))
