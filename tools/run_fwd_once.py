import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kalman_vae_b200 import functional as F, capi
from kalman_vae_b200.functional import Problem
from kalman_vae_b200.synthetic import Shape, make_case
dev = torch.device("cuda:0")
B, T, lanes = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
shape = Shape(B, T, 4, 2, 4, 3)
case = make_case(shape, seed=10, mask_kind="bernoulli")
g = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in case.items()}
pb = Problem(g["Y"], None, g["mask"], g["alpha"], g["A"], g["B"], g["C"], g["Q"], g["R"], g["mu0"], g["Sigma0"], False, False, lanes=lanes)
for _ in range(3):
    st, *_ = F.smooth_fwd(pb)
torch.cuda.synchronize()
print("ok", float(st.mus_smooth.sum()))
