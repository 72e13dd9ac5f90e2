#!/usr/bin/env python
"""How long does the host need to enqueue one training step (vs the GPU time of the step)?"""
import contextlib, io, os, sys, time, cProfile, pstats
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import geoguessr_ai_b200 as gg
from geoguessr_ai_b200 import synth
from geoguessr_ai_b200.geocells import load_packaged_centroids

dev = torch.device("cuda:0")
B, D = 4096, 1024
cent = load_packaged_centroids()
with contextlib.redirect_stdout(io.StringIO()):
    model = gg.SuperGuessr(None, panorama=True, should_smooth_labels=True, embed_dim=D, centroids=cent).to(dev)
model.train()
params = [model.cell_layer.weight, model.cell_layer.bias]
opt = torch.optim.AdamW(params, lr=1e-4, fused=True)
emb, _, _, labels = synth.head_inputs(B, D, 8, V=4, seed=1)
emb, labels = emb.to(dev), labels.to(dev)
clf = torch.zeros(B, dtype=torch.int64, device=dev)

def step():
    opt.zero_grad(set_to_none=True)
    out = model(embedding=emb, labels=labels, labels_clf=clf)
    out.loss.backward()
    opt.step()
    return out.loss

for _ in range(10):
    step()
torch.cuda.synchronize()
N = 200
t0 = time.perf_counter()
for _ in range(N):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"enqueue {1e6 * (t1 - t0) / N:.1f} us/step, total {1e6 * (t2 - t0) / N:.1f} us/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(100):
    step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
print(s.getvalue()[:6000])
