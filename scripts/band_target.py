"""Target for ncu: a few BatchedWorld steps of the BASELINE config-4 pile (one world, banded kernel).
usage: band_target.py [steps] [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lcp_physics_b200.scenes import make_ball_pile
from lcp_physics_b200.world import BatchedWorld

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ic = make_ball_pile(B, nballs=512, cols=32, seed=3000, gap=0.05)
w = BatchedWorld(ic["pos"], ic["rad"], vel=ic["vel"], mass=ic["mass"], restitution=ic["rest"], fric_coeff=ic["fric"],
                 gravity=100.0, static=[0], contact_capacity=2048)
for _ in range(steps):
    w.step()
torch.cuda.synchronize()
print("ok", w.counts.tolist()[:4])
