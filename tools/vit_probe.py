"""Timing of the Viterbi kernels alone (CUDA events) at the cfg3 shape: python tools/vit_probe.py [n_utts] [P]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beer_b200 import ops, synthetic  # noqa: E402

U = int(sys.argv[1]) if len(sys.argv) > 1 else 1250
P = int(sys.argv[2]) if len(sys.argv) > 2 else 250
T, S = 1000, 4
K = P * S
dev = torch.device('cuda', 0)
graph, _, _ = synthetic.phone_loop_graph(P, S)
plan = ops.GraphPlan(graph.init_log_probs.numpy(), graph.final_log_probs.numpy(), graph.trans_log_probs.numpy(),
                     graph.pdf_id_mapping, n_pdfs=K)
llh = torch.randn(U * T, K, device=dev) * 3
off = torch.arange(U + 1, device=dev, dtype=torch.int64) * T
ws = torch.empty(U * T * K // 2 + 16, device=dev, dtype=torch.float32)
for _ in range(2):
    path = ops.hmm_viterbi(plan, llh, off, workspace=ws)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    path = ops.hmm_viterbi(plan, llh, off, workspace=ws)
e1.record()
torch.cuda.synchronize()
print(f'viterbi {U} utts x {T} frames, {P} units x {S}: {e0.elapsed_time(e1) / 5:.3f} ms, env VIT_DENSE={os.environ.get("BEER_B200_VIT_DENSE")}')
