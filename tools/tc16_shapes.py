#!/usr/bin/env python
"""Throughput of the tc16 Delayed-Acceptance kernel on cfg2-shaped problems of other sizes
(zero-padded operands): transitions/s per (d, m_c, m_f), 65536 chains, fine-level
stats-only history, CUDA events around 5 launches of 50 fine iterations after 3 warm-up launches.
usage: python tools/tc16_shapes.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tinyda_b200 import lower_problem
from tinyda_b200.engine import Engine, STORE_NONE, STORE_STATS
from tinyda_b200.workloads import cfg2_da

C, iters, J = 65536, 50, 10
stream = torch.cuda.current_stream().cuda_stream
for d, m_c, m_f in [(64, 128, 1024), (48, 128, 1024), (32, 128, 1024), (16, 128, 1024), (64, 64, 512), (32, 100, 1000), (64, 128, 1920)]:
    w = cfg2_da(d=d, m_f=m_f, m_c=m_c)
    spec = lower_problem(w["posteriors"], w["proposal"], J)
    theta0 = w["prior"].rvs(C, random_state=np.random.default_rng(1)).reshape(C, d)
    eng = Engine(spec, C, dtype="float32", rng="philox", seed=1, store=[STORE_NONE, STORE_STATS],
                 capacity_iterations=iters, stream=stream)
    eng.init(theta0)
    assert eng.kernel() == "tc16", eng.kernel()
    for _ in range(3):
        eng.history_reset(); eng.run(iters)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        eng.history_reset(); eng.run(iters)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print("d=%2d m_c=%3d m_f=%4d: %.3f ms per %d fine iterations -> %.0f M transitions/s"
          % (d, m_c, m_f, ms, iters, C * iters / ms / 1e3), flush=True)
    eng.close()
