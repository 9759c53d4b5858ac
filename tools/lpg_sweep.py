"""prints the LPG roofline dict of bench.py (back-to-back C-ABI launches, CUDA events)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
print(json.dumps(bench.lpg_roofline(torch, torch.device("cuda:0"), bench.peaks())))
