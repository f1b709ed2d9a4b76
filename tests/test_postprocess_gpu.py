"""GPU box decoding + NMS (pg_postprocess / pg_nms_boxes_3d / pg_decode_boxes through the reference-shaped
models.nms / models.box_encoding / models.postprocess API) vs the reference-derived fixtures and the NumPy oracle.

Tolerance: kept sets and labels identical; boxes and scores within 1e-4 (float32 trigonometry / exp of CUDA vs NumPy
differ in the last place; the geometry is float64 on both sides)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import postprocess as pp

pytestmark = pytest.mark.gpu
CASES = [('car', 'Car', 4), ('ped', 'Pedestrian_and_Cyclist', 6)]


def _cuda(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


@pytest.mark.parametrize('name,method,c', CASES)
def test_decode_matches_reference(name, method, c):
    from pointgnn_b200.models import box_encoding
    g = dict(np.load(os.path.join(GOLDEN, 'post_%s.npz' % name)))
    k = g['probs'].shape[0]
    # the reference's call shape (run.py:272-280): one row per (vertex, class) pair, the label in cls_labels
    labels = np.tile(np.arange(c)[None, :], (k, 1)).reshape(-1, 1)
    xyz = np.tile(g['points_xyz'][:, None, :], (1, c, 1)).reshape(-1, 3)
    fn = box_encoding.get_box_decoding_fn('classaware_all_class_box_encoding')
    dec = fn(labels, xyz, g['box_encodings'].reshape(-1, 1, 7), pp.LABEL_MAPS[method])
    assert dec.shape == (k * c, 1, 7)
    assert np.abs(dec[:, 0] - g['decoded']).max() < 1e-5
    assert box_encoding.get_encoding_len('classaware_all_class_box_encoding') == 7


@pytest.mark.parametrize('name,method,c', CASES)
@pytest.mark.parametrize('variant', ['uncertainty', 'merge_only', 'score_only', 'plain'])
def test_nms_entry_points_match_reference(name, method, c, variant):
    from pointgnn_b200.models import nms
    g = dict(np.load(os.path.join(GOLDEN, 'post_%s.npz' % name)))
    dec = pp.decode_boxes(g['box_encodings'], g['points_xyz'], pp.LABEL_MAPS[method])
    lab, boxes, scores, idx = pp.select_candidates(g['probs'], dec, c)
    fn = {'uncertainty': nms.nms_boxes_3d_uncertainty, 'merge_only': nms.nms_boxes_3d_merge_only,
          'score_only': nms.nms_boxes_3d_score_only, 'plain': nms.nms_boxes_3d}[variant]
    out = fn(lab, boxes, scores, overlapped_fn=nms.overlapped_boxes_3d_fast_poly, overlapped_thres=float(g['thres']),
             appr_factor=100.0, top_k=-1, attributes=np.arange(len(lab)))
    assert np.array_equal(out[3], g[variant + '_nms_index'])
    assert np.array_equal(out[0], g[variant + '_label'])
    assert np.abs(out[1] - g[variant + '_box']).max() < 1e-4
    assert np.abs(out[2] - g[variant + '_score']).max() < 1e-4
    with pytest.raises(NotImplementedError):
        fn(lab, boxes, scores, overlapped_fn=nms.overlapped_boxes_3d)


def test_fused_batch_postprocess_vs_oracle():
    """pg_postprocess on a 3-frame batch (different sizes, one frame without any candidate) = the oracle frame by
    frame: candidate lists, kept sets, labels, boxes, scores."""
    from pointgnn_b200.models import postprocess
    frames = [pp.synthetic_outputs(21, 20, 30, 4), pp.synthetic_outputs(22, 3, 10, 4), pp.synthetic_outputs(23, 40, 40, 4, 0.9)]
    empty = (frames[1][0].copy(), frames[1][1].copy(), np.tile(np.array([[1, 0, 0, 0]], np.float32), (30, 1)))
    frames.insert(1, empty)
    pts = np.vstack([f[0] for f in frames])
    enc = np.vstack([f[1] for f in frames])
    probs = np.vstack([f[2] for f in frames])
    fp = np.cumsum([0] + [len(f[0]) for f in frames]).astype(np.int32)
    det = postprocess.detect(_cuda(probs), _cuda(enc), _cuda(pts), _cuda(fp), label_method='Car',
                             nms_overlapped_thres=0.01, want_candidates=True)
    dfp = det['frame_ptr'].cpu().numpy()
    cfp = det['cand_frame_ptr'].cpu().numpy()
    assert dfp[0] == 0 and len(dfp) == len(frames) + 1
    for f, (p, e, pr) in enumerate(frames):
        dec = pp.decode_boxes(e, p, pp.LABEL_MAPS['Car'])
        lab, boxes, scores, idx = pp.select_candidates(pr, dec, 4)
        cand = det['cand_index'][cfp[f]:cfp[f + 1]].cpu().numpy()
        assert np.array_equal(cand - fp[f] * 4, idx)
        sl = slice(dfp[f], dfp[f + 1])
        if len(lab) == 0:
            assert dfp[f] == dfp[f + 1]
            continue
        want_l, want_b, want_s, order = pp.nms_boxes_3d_uncertainty(lab, boxes, scores, 0.01)
        assert np.array_equal(det['index'][sl].cpu().numpy() - fp[f] * 4, idx[order])
        assert np.array_equal(det['label'][sl].cpu().numpy(), want_l)
        assert np.abs(det['box'][sl].cpu().numpy() - want_b).max() < 1e-4
        assert np.abs(det['score'][sl].cpu().numpy() - want_s).max() < 1e-4


def test_ties_and_identical_boxes():
    """Exactly identical boxes with identical scores (ties in the sort, IoU = 1, even-sized medians)."""
    from pointgnn_b200.models import nms
    box = np.array([[1.0, 1.5, 20.0, 3.9, 1.5, 1.6, 0.3]], np.float32)
    boxes = np.repeat(box, 6, axis=0)
    boxes[4:, 0] += 30.0                      # a second, disjoint pair
    scores = np.array([0.9, 0.9, 0.5, 0.9, 0.7, 0.7], np.float32)
    labels = np.ones(6, np.int64)
    out = nms.nms_boxes_3d_uncertainty(labels, boxes, scores, overlapped_fn=nms.overlapped_boxes_3d_fast_poly,
                                       overlapped_thres=0.01, appr_factor=100.0, top_k=-1, attributes=np.arange(6))
    assert len(out[0]) == 2
    assert np.allclose(out[1][0], box[0], atol=1e-6) and np.allclose(out[1][1, 0], 31.0, atol=1e-6)
    assert abs(out[2][0] - (0.9 + 0.9 + 0.9 + 0.5)) < 1e-5 and abs(out[2][1] - 1.4) < 1e-5
    assert set(out[3].tolist()) <= {0, 1, 3, 4, 5}
