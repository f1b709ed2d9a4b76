"""Drive only the fused tcgen05 edge kernel (for ncu): 8 synthetic frames, car_auto_T3 layer-2 weights."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402
from pointgnn_b200 import _lib  # noqa: E402
if os.environ.get('PG_USE_LAB_LIB'):      # lab build (make -C point-gnn_b200/csrc lab): in-kernel tracing via PG_TC_TRACE
    _lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), 'libpointgnn_b200_lab.so')
if os.environ.get('PG_LIB_VARIANT'):      # experiment builds (tools/build_variant.sh NAME -DFLAG...): lab/<NAME>.so
    _lib.LIB_PATH = os.path.join(ROOT, 'lab', os.environ['PG_LIB_VARIANT'] + '.so')
from pointgnn_b200.models import graph_gen  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
prec = int(sys.argv[3]) if len(sys.argv) > 3 else 1
cfg = json.load(open(os.path.join(ROOT, 'tests/golden/config_car_auto_T3_train.json')))
w = dict(np.load(os.path.join(ROOT, 'tests/golden/weights_car_auto_T3_train.npz')))
pts = np.vstack([synth.lidar_frame(i, 20000)[0] for i in range(frames)])
fp = torch.arange(frames + 1, dtype=torch.int32, device='cuda') * 20000
coords, kp, edges = graph_gen.gen_multi_level_local_graph_v3(torch.from_numpy(pts).cuda(), frame_ptr=fp,
                                                           **cfg['runtime_graph_gen_kwargs'])
k = coords[1].shape[0]
feats = torch.rand((k, 300), device='cuda') * 0.5
s = 'layer2/extract_vertex_features/fully_connected'
ws = [torch.from_numpy(w[s + '/weights']).cuda(), torch.from_numpy(w[s + '_1/weights']).cuda()]
bs = [torch.from_numpy(w[s + '/biases']).cuda(), torch.from_numpy(w[s + '_1/biases']).cuda()]
src, dst = edges[1][:, 0].contiguous(), edges[1][:, 1].contiguous()
print('K', k, 'E1', src.numel())
layer = _lib.PreparedLayer(_lib.PG_LAYER_EDGE_GNN, ws, bs, [303, 300, 300], prec)
if prec == 1 and not os.environ.get('PG_TC_TRACE'):
    ref = _lib.PreparedLayer(_lib.PG_LAYER_EDGE_GNN, ws, bs, [303, 300, 300], 0).edge_mlp_max(
        feats, coords[1], coords[1], None, src, dst, k, trusted=True)
    got = layer.edge_mlp_max(feats, coords[1], coords[1], None, src, dst, k, trusted=True)
    print('max |tensor-core - fp32 FFMA| = %.3g (empty segments equal: %s)' % (
        float((got - ref).abs()[ref > -1e30].max()), bool(((got < -1e30) == (ref < -1e30)).all())))
for _ in range(2):
    layer.edge_mlp_max(feats, coords[1], coords[1], None, src, dst, k, trusted=True)
torch.cuda.synchronize()
a = torch.cuda.Event(enable_timing=True)
b = torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    layer.edge_mlp_max(feats, coords[1], coords[1], None, src, dst, k, trusted=True)
b.record()
b.synchronize()
ms = a.elapsed_time(b) / reps
print('prepared edge layer (P GEMM + fill + fused edge kernel) precision %d: %.3f ms per call, %.1f algorithmic TFLOP/s'
      % (prec, ms, src.numel() * 361800 / ms / 1e9))
