"""CPU oracle for the GNN half of the hot path (reference models/gnn.py, models/models.py).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Line-by-line NumPy restatement, in
fp32 (dtype=np.float64 gives the "exact" arithmetic used to size tolerances), of

* multi_layer_neural_network_fn      gnn.py:86-104
* multi_layer_fc_fn                  gnn.py:34-84
* graph_scatter_max_fn               gnn.py:106-109  (tf.math.unsorted_segment_max:
                                     empty segment -> numeric_limits<float>::lowest())
* ClassAwarePredictor.apply_regular  gnn.py:133-163
* PointSetPooling.apply_regular      gnn.py:222-283
* GraphNetAutoCenter.apply_regular   gnn.py:298-373
* MultiLayerFastLocalGraphModelV2.predict / postprocess   models.py:79-168

``slim.fully_connected`` = act(x @ W + b) with W stored [in, out]; all shipped
configs use normalization 'NONE' and activation 'ReLU' (SURVEY fact 3), the only
combination implemented.  Variable names follow TF-slim scoping so the
reference's checkpoints load unchanged: '<scope>/fully_connected[_i]/{weights,biases}'.
"""
import numpy as np


class _Scope(object):
    """Mimics slim's per-variable_scope 'fully_connected', 'fully_connected_1', ... naming."""

    def __init__(self, weights, prefix):
        self.weights = weights
        self.prefix = prefix
        self.count = 0

    def sub(self, name):
        return _Scope(self.weights, self.prefix + '/' + name if self.prefix else name)

    def next_fc(self):
        name = 'fully_connected' if self.count == 0 else 'fully_connected_%d' % self.count
        self.count += 1
        base = self.prefix + '/' + name
        return self.weights[base + '/weights'], self.weights[base + '/biases']


def _fc(x, scope, relu):
    w, b = scope.next_fc()
    assert x.shape[1] == w.shape[0], (x.shape, w.shape, scope.prefix)
    y = x @ w.astype(x.dtype) + b.astype(x.dtype)[None, :]
    if relu:
        np.maximum(y, 0, out=y)
    return y


def _check_types(normalization_type, activation_type):
    assert normalization_type == 'NONE', 'only normalization NONE is used by the shipped configs'
    assert activation_type == 'ReLU', 'only ReLU is used by the shipped configs'


def multi_layer_neural_network_fn(features, scope, Ks=(64, 32, 64), is_logits=False,
                                  normalization_type='NONE', activation_type='ReLU'):
    """gnn.py:86-104."""
    assert len(features.shape) == 2
    _check_types(normalization_type, activation_type)
    for i in range(len(Ks)):
        last = i == len(Ks) - 1
        features = _fc(features, scope, relu=not (is_logits and last))
        assert features.shape[1] == Ks[i]
    return features


def multi_layer_fc_fn(sv, scope, Ks=(64, 32, 64), num_classes=4, is_logits=False, num_layer=4,
                      normalization_type='NONE', activation_type='ReLU'):
    """gnn.py:34-84 (mask unused by the predictor)."""
    assert len(sv.shape) == 2
    assert len(Ks) == num_layer - 1
    _check_types(normalization_type, activation_type)
    features = sv
    for i in range(num_layer - 1):
        features = _fc(features, scope, relu=True)
    features = _fc(features, scope, relu=not is_logits)
    assert features.shape[1] == num_classes
    return features


def graph_scatter_max_fn(point_features, point_centers, num_centers):
    """gnn.py:106-109; tf.math.unsorted_segment_max semantics."""
    point_centers = np.asarray(point_centers).reshape(-1).astype(np.int64)
    lowest = np.finfo(point_features.dtype).min
    out = np.full((int(num_centers), point_features.shape[1]), lowest, dtype=point_features.dtype)
    if point_features.shape[0] == 0:
        return out
    if np.any(np.diff(point_centers) < 0):
        order = np.argsort(point_centers, kind='stable')
        point_centers = point_centers[order]
        point_features = point_features[order]
    starts = np.flatnonzero(np.concatenate([[True], point_centers[1:] != point_centers[:-1]]))
    out[point_centers[starts]] = np.maximum.reduceat(point_features, starts, axis=0)
    return out


def _edge_chunks(num_edges, dst, chunk):
    """Chunk boundaries over a dst-grouped edge list (bounds oracle memory)."""
    s = 0
    while s < num_edges:
        e = min(num_edges, s + chunk)
        yield s, e
        s = e


def point_set_pooling(weights, scope_name, point_features, point_coordinates, keypoint_indices,
                      set_indices, point_MLP_depth_list=None, point_MLP_normalization_type='NONE',
                      point_MLP_activation_type='ReLU', output_MLP_depth_list=None,
                      output_MLP_normalization_type='NONE', output_MLP_activation_type='ReLU',
                      chunk=1 << 18):
    """PointSetPooling.apply_regular, gnn.py:222-283."""
    dt = point_features.dtype
    set_indices = np.asarray(set_indices).astype(np.int64)
    keypoint_indices = np.asarray(keypoint_indices).astype(np.int64)
    num_k = keypoint_indices.shape[0]
    lowest = np.finfo(dt).min
    set_features = np.full((num_k, point_MLP_depth_list[-1]), lowest, dtype=dt)
    for s, e in _edge_chunks(set_indices.shape[0], set_indices[:, 1], chunk):
        si = set_indices[s:e]
        psf = point_features[si[:, 0]]                                   # :256
        psc = point_coordinates[si[:, 0]]                                # :257
        kidx = keypoint_indices[si[:, 1]]                                # :259-260
        kc = point_coordinates[kidx[:, 0]]                               # :261-262
        psc = psc - kc                                                   # :264-265
        x = np.concatenate([psf, psc], axis=-1)                          # :266-267
        sc = _Scope(weights, scope_name).sub('extract_vertex_features')
        x = multi_layer_neural_network_fn(x, sc, Ks=point_MLP_depth_list, is_logits=False,
                                          normalization_type=point_MLP_normalization_type,
                                          activation_type=point_MLP_activation_type)
        part = graph_scatter_max_fn(x, si[:, 1], num_k)                  # :275-277
        np.maximum(set_features, part, out=set_features)
    sc = _Scope(weights, scope_name).sub('combined_features')
    return multi_layer_neural_network_fn(set_features, sc, Ks=output_MLP_depth_list,
                                         is_logits=False,
                                         normalization_type=output_MLP_normalization_type,
                                         activation_type=output_MLP_activation_type)


def graph_net_auto_center(weights, scope_name, input_vertex_features, input_vertex_coordinates,
                          NOT_USED, edges, edge_MLP_depth_list=None,
                          edge_MLP_normalization_type='NONE', edge_MLP_activation_type='ReLU',
                          update_MLP_depth_list=None, update_MLP_normalization_type='NONE',
                          update_MLP_activation_type='ReLU', auto_offset=False,
                          auto_offset_MLP_depth_list=None,
                          auto_offset_MLP_normalization_type='NONE',
                          auto_offset_MLP_feature_activation_type='ReLU', chunk=1 << 18,
                          return_intermediates=False):
    """GraphNetAutoCenter.apply_regular, gnn.py:298-373."""
    dt = input_vertex_features.dtype
    edges = np.asarray(edges).astype(np.int64)
    num_v = input_vertex_features.shape[0]
    top = _Scope(weights, scope_name)
    coords = input_vertex_coordinates
    offset = None
    if auto_offset:                                                      # :341-346
        offset = multi_layer_neural_network_fn(
            input_vertex_features, top, Ks=auto_offset_MLP_depth_list, is_logits=True,
            normalization_type=auto_offset_MLP_normalization_type,
            activation_type=auto_offset_MLP_feature_activation_type)
        coords = input_vertex_coordinates + offset
    lowest = np.finfo(dt).min
    agg = np.full((num_v, edge_MLP_depth_list[-1]), lowest, dtype=dt)
    for s, e in _edge_chunks(edges.shape[0], edges[:, 1], chunk):
        ed = edges[s:e]
        s_feat = input_vertex_features[ed[:, 0]]                         # :338
        s_coord = input_vertex_coordinates[ed[:, 0]]                     # :339 (un-offset)
        d_coord = coords[ed[:, 1]]                                       # :348 (offset)
        x = np.concatenate([s_feat, s_coord - d_coord], axis=-1)         # :350-352
        sc = _Scope(weights, scope_name).sub('extract_vertex_features')
        x = multi_layer_neural_network_fn(x, sc, Ks=edge_MLP_depth_list, is_logits=False,
                                          normalization_type=edge_MLP_normalization_type,
                                          activation_type=edge_MLP_activation_type)
        part = graph_scatter_max_fn(x, ed[:, 1], num_v)                  # :362-365
        np.maximum(agg, part, out=agg)
    sc = _Scope(weights, scope_name).sub('combined_features')
    update = multi_layer_neural_network_fn(agg, sc, Ks=update_MLP_depth_list, is_logits=True,
                                           normalization_type=update_MLP_normalization_type,
                                           activation_type=update_MLP_activation_type)
    out = update + input_vertex_features                                 # :372
    if return_intermediates:
        return out, dict(offset=offset, aggregated=agg, update=update)
    return out


def class_aware_predictor(weights, scope_name, features, num_classes, box_encoding_len,
                          normalization_type='NONE', activation_type='ReLU',
                          cls_Ks=(64,), loc_Ks=(64, 64)):
    """ClassAwarePredictor.apply_regular, gnn.py:133-163 with the fns of models.py:60-64."""
    pred = _Scope(weights, scope_name).sub('predictor')
    logits = multi_layer_fc_fn(features, pred.sub('cls'), Ks=cls_Ks, num_layer=len(cls_Ks) + 1,
                               num_classes=num_classes, is_logits=True,
                               normalization_type=normalization_type,
                               activation_type=activation_type)
    boxes = []
    for class_idx in range(num_classes):
        b = multi_layer_fc_fn(features, pred.sub('loc').sub('cls_%d' % class_idx), Ks=loc_Ks,
                              num_layer=len(loc_Ks) + 1, num_classes=box_encoding_len,
                              is_logits=True, normalization_type=normalization_type,
                              activation_type=activation_type)
        boxes.append(b[:, None, :])
    return logits, np.concatenate(boxes, axis=1)


_PREDICTOR_KS = {
    'classaware_predictor': ((64,), (64, 64)),
    'classaware_predictor_128': ((128,), (128, 128)),
}


def predict(weights, layer_configs, num_classes, box_encoding_len, t_initial_vertex_features,
            t_vertex_coord_list, t_keypoint_indices_list, t_edges_list, dtype=np.float32,
            return_features=False):
    """MultiLayerFastLocalGraphModelV2.predict, models.py:79-163."""
    feats = np.asarray(t_initial_vertex_features, dtype=dtype)
    coords = [np.asarray(c, dtype=dtype) for c in t_vertex_coord_list]
    feature_list = [feats]
    for layer_config in layer_configs[:-1]:
        lvl = layer_config['graph_level']
        kw = layer_config['kwargs']
        if layer_config['type'] == 'scatter_max_point_set_pooling':
            feats = point_set_pooling(weights, layer_config['scope'], feats, coords[lvl],
                                      t_keypoint_indices_list[lvl], t_edges_list[lvl], **kw)
        elif layer_config['type'] == 'scatter_max_graph_auto_center_net':
            feats = graph_net_auto_center(weights, layer_config['scope'], feats, coords[lvl],
                                          t_keypoint_indices_list[lvl], t_edges_list[lvl], **kw)
        else:
            raise KeyError(layer_config['type'])
        feature_list.append(feats)
    pc = layer_configs[-1]
    cls_Ks, loc_Ks = _PREDICTOR_KS[pc['type']]
    logits, boxes = class_aware_predictor(weights, pc['scope'], feats, num_classes,
                                          box_encoding_len, cls_Ks=cls_Ks, loc_Ks=loc_Ks,
                                          **pc['kwargs'])
    if return_features:
        return logits, boxes, feature_list
    return logits, boxes


def postprocess(logits):
    """models.py:165-168: softmax over classes."""
    z = logits - logits.max(axis=-1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=-1, keepdims=True)
