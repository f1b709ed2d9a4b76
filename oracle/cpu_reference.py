"""The reference's CPU path, as fast as its own dependencies would run it - the CPU baseline.

TEST / BENCH INFRASTRUCTURE (see oracle/__init__.py); used by bench.py's cpu_baseline and
``--impl reference`` legs and validated against oracle/graph.py + oracle/gnn.py in the CPU tests.

* graph stage  : the same scikit-learn calls the reference makes (graph_gen.py:84-87 kd_tree 1-NN,
  :207-209 ball_tree radius_neighbors, n_jobs=1 as the reference pins) around the oracle's
  voxel-centroid restatement of open3d.voxel_down_sample (graph_gen.py:41-45).
* GNN stage    : gnn.py:222-283 / 298-373 / 133-163 + models.py:113-168 restated with torch-CPU
  ops (multi-threaded gather / concat / matmul / segment-max), standing in for TF-1.15's
  Eigen/MKL CPU kernels, with torch.set_num_threads(all host cores).
"""
import os

import numpy as np
import torch

from . import graph as ograph


def gen_graph(points_xyz, base_voxel_size, level_configs, add_rnd3d=False, downsample_method='center'):
    from sklearn.neighbors import NearestNeighbors
    assert not add_rnd3d and downsample_method == 'center'
    scales = [c['graph_scale'] for c in level_configs]
    coords = [points_xyz]
    keypoints = []
    last = 0
    for level in scales:
        base = coords[-1]
        if np.isclose(level, last):
            coords.append(base)
            keypoints.append(np.arange(base.shape[0])[:, None])
        else:
            cent = ograph.voxel_down_sample(points_xyz, np.asarray(base_voxel_size) * level)
            nbrs = NearestNeighbors(n_neighbors=1, algorithm='kd_tree', n_jobs=1).fit(base)
            idx = nbrs.kneighbors(cent, return_distance=False)
            coords.append(base[idx[:, 0], :])
            keypoints.append(idx)
        last = level
    edges = []
    for cfg in level_configs:
        lvl = cfg['graph_level']
        kw = cfg['graph_gen_kwargs']
        nbrs = NearestNeighbors(radius=kw['radius'], algorithm='ball_tree', n_jobs=1).fit(coords[lvl])
        ind = nbrs.radius_neighbors(coords[lvl + 1], return_distance=False)
        v = np.concatenate(ind)
        i = np.concatenate([j * np.ones(n.size, dtype=np.int32) for j, n in enumerate(ind)])
        edges.append(np.array([v, i]).transpose())
    return coords, keypoints, edges


class _W(object):
    def __init__(self, weights):
        self.w = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in weights.items()
                  if k.endswith('/weights') or k.endswith('/biases')}
        self.count = {}

    def fc(self, scope):
        i = self.count.get(scope, 0)
        self.count[scope] = i + 1
        base = scope + '/' + ('fully_connected' if i == 0 else 'fully_connected_%d' % i)
        return self.w[base + '/weights'], self.w[base + '/biases']


def _mlp(x, w, scope, n, is_logits):
    for i in range(n):
        wt, b = w.fc(scope)
        x = torch.addmm(b, x, wt)
        if not (is_logits and i == n - 1):
            x.clamp_min_(0)
    return x


def _segment_max(x, dst, num):
    """unsorted_segment_max (gnn.py:106-109; empty segment -> float lowest).  The edge lists of this path are
    grouped by destination, so the reduction is a sorted ``torch.segment_reduce`` over per-destination run
    lengths (deterministic and well threaded; the round-1 ``index_reduce_('amax')`` made the CPU arm vary
    4.5x between boxes).  Unsorted ids are sorted first."""
    lowest = torch.finfo(x.dtype).min
    if dst.numel() and bool((dst[1:] < dst[:-1]).any()):
        dst, order = torch.sort(dst, stable=True)
        x = x[order]
    lengths = torch.bincount(dst, minlength=num)
    return torch.segment_reduce(x, 'max', lengths=lengths, axis=0, initial=lowest, unsafe=True)


def predict(weights, layer_configs, num_classes, box_encoding_len, features, coords, keypoints, edges,
            chunk=1 << 19, num_threads=None):
    torch.set_num_threads(num_threads or os.cpu_count() or 1)
    w = _W(weights)
    f = torch.from_numpy(np.ascontiguousarray(features, dtype=np.float32))
    coords = [torch.from_numpy(np.ascontiguousarray(c, dtype=np.float32)) for c in coords]
    for lc in layer_configs[:-1]:
        lvl, kw, s = lc['graph_level'], lc['kwargs'], lc['scope']
        ed = torch.from_numpy(np.ascontiguousarray(edges[lvl]).astype(np.int64))
        x = coords[lvl]
        if lc['type'] == 'scatter_max_point_set_pooling':
            kp = torch.from_numpy(np.ascontiguousarray(keypoints[lvl]).astype(np.int64))[:, 0]
            nk = kp.shape[0]
            agg = torch.full((nk, kw['point_MLP_depth_list'][-1]), torch.finfo(torch.float32).min)
            for a in range(0, ed.shape[0], chunk):
                e = ed[a:a + chunk]
                e0 = torch.cat([f[e[:, 0]], x[e[:, 0]] - x[kp[e[:, 1]]]], dim=-1)
                w.count.pop(s + '/extract_vertex_features', None)
                h = _mlp(e0, w, s + '/extract_vertex_features', len(kw['point_MLP_depth_list']), False)
                agg = torch.maximum(agg, _segment_max(h, e[:, 1], nk))
            f = _mlp(agg, w, s + '/combined_features', len(kw['output_MLP_depth_list']), False)
        else:
            nv = f.shape[0]
            xd = x
            if kw.get('auto_offset'):
                xd = x + _mlp(f, w, s, len(kw['auto_offset_MLP_depth_list']), True)
            agg = torch.full((nv, kw['edge_MLP_depth_list'][-1]), torch.finfo(torch.float32).min)
            for a in range(0, ed.shape[0], chunk):
                e = ed[a:a + chunk]
                e0 = torch.cat([f[e[:, 0]], x[e[:, 0]] - xd[e[:, 1]]], dim=-1)
                w.count.pop(s + '/extract_vertex_features', None)
                h = _mlp(e0, w, s + '/extract_vertex_features', len(kw['edge_MLP_depth_list']), False)
                agg = torch.maximum(agg, _segment_max(h, e[:, 1], nv))
            f = _mlp(agg, w, s + '/combined_features', len(kw['update_MLP_depth_list']), True) + f
    pc = layer_configs[-1]
    width = 128 if pc['type'] == 'classaware_predictor_128' else 64
    p = pc['scope'] + '/predictor'
    logits = _mlp(f, w, p + '/cls', 2, True)
    boxes = torch.stack([_mlp(f, w, p + '/loc/cls_%d' % c, 3, True) for c in range(num_classes)], dim=1)
    assert width in (64, 128)
    probs = torch.softmax(logits, dim=-1)
    return logits.numpy(), boxes.numpy(), probs.numpy()
