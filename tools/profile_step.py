"""Where does a training step's GPU time go?  (torch profiler, top CUDA kernels)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from pointnet_autoencoder_b200 import models, parallel, synthetic

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = models.AutoEncoderUpconv().to(dev)
bucket = parallel.GradBucket(model.parameters())
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
label, _ = synthetic.s_chair(32, 2048)
x = torch.from_numpy(label).to(dev)

def step():
    pred, _ = model(x, 0.5)
    loss, _ = models.chamfer_loss(pred, x)
    bucket.zero(); loss.backward(); opt.step()

for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
