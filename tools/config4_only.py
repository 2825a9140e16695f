"""Runs bench.py's config-4 block alone (per-op milliseconds of the sampling / grouping / interpolation chain).  Development tool."""
import json, os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

D = types.SimpleNamespace(torch=torch, dev=torch.device("cuda:0"))
for k, v in bench.config4_block(D, 1965).items():
    print(k, json.dumps(v))
