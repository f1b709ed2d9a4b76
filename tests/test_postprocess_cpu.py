"""Post-processing oracle (oracle/postprocess.py) vs the reference's own box_encoding.py / nms.py.

tests/golden/post_<case>.npz hold run.py:265-325 executed by the REFERENCE modules (shapely replaced by a convex
polygon stand-in - the footprints are rectangles) on seeded synthetic network outputs, for all four NMS variants.
The NumPy restatement - the checker of the CUDA kernels on the GPU box - must reproduce them."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import postprocess as pp

CASES = [('car', 'Car', 4), ('ped', 'Pedestrian_and_Cyclist', 6)]


def test_polygon_stand_in_on_known_shapes():
    sq = pp.Polygon([[0, 0], [2, 0], [2, 2], [0, 2]])
    assert sq.area == 4.0
    assert pp.Polygon([[0, 0], [0, 2], [2, 2], [2, 0]]).area == 4.0            # clockwise: unsigned
    shifted = pp.Polygon([[1, 1], [3, 1], [3, 3], [1, 3]])
    assert abs(sq.intersection(shifted).area - 1.0) < 1e-12
    assert abs(shifted.intersection(sq).area - 1.0) < 1e-12
    assert sq.intersection(pp.Polygon([[5, 5], [6, 5], [6, 6], [5, 6]])).area == 0.0
    # a diamond inscribed in the square: intersection = the diamond (area 2)
    diamond = pp.Polygon([[1, 0], [2, 1], [1, 2], [0, 1]])
    assert abs(sq.intersection(diamond).area - 2.0) < 1e-12
    # rotated square about the same centre (45 degrees): regular octagon, area 8 (sqrt2 - 1) a^2 with a = 1 ... = 3.3137
    r = np.sqrt(2.0)
    rot = pp.Polygon([[1 + r, 1], [1, 1 + r], [1 - r, 1], [1, 1 - r]])
    assert abs(sq.intersection(rot).area - 8 * (np.sqrt(2) - 1)) < 1e-9


@pytest.mark.parametrize('name,method,c', CASES)
def test_decode_and_candidates_match_reference(name, method, c):
    g = dict(np.load(os.path.join(GOLDEN, 'post_%s.npz' % name)))
    dec = pp.decode_boxes(g['box_encodings'], g['points_xyz'], pp.LABEL_MAPS[method])
    assert np.abs(dec.reshape(-1, 7) - g['decoded']).max() < 1e-5
    lab, boxes, scores, idx = pp.select_candidates(g['probs'], dec, c)
    assert np.array_equal(idx, g['cand_index'])


@pytest.mark.parametrize('name,method,c', CASES)
@pytest.mark.parametrize('variant,merge,rescore', [('uncertainty', True, True), ('merge_only', True, False),
                                                   ('score_only', False, True)])
def test_nms_restatement_matches_reference(name, method, c, variant, merge, rescore):
    g = dict(np.load(os.path.join(GOLDEN, 'post_%s.npz' % name)))
    dec = pp.decode_boxes(g['box_encodings'], g['points_xyz'], pp.LABEL_MAPS[method])
    lab, boxes, scores, idx = pp.select_candidates(g['probs'], dec, c)
    out_l, out_b, out_s, order = pp.nms_boxes_3d_uncertainty(lab, boxes, scores, float(g['thres']), merge, rescore)
    assert np.array_equal(order, g[variant + '_nms_index'])
    assert np.array_equal(out_l, g[variant + '_label'])
    assert np.abs(out_b - g[variant + '_box']).max() < 1e-5
    assert np.abs(out_s - g[variant + '_score']).max() < 1e-5


def test_postprocess_frame_wrapper():
    g = dict(np.load(os.path.join(GOLDEN, 'post_car.npz')))
    lab, bx, sc = pp.postprocess_frame(g['probs'], g['box_encodings'], g['points_xyz'], 'Car', float(g['thres']))
    assert np.array_equal(lab, g['uncertainty_label']) and np.abs(bx - g['uncertainty_box']).max() < 1e-5
    # no candidate at all
    probs = np.zeros((5, 4), np.float32)
    probs[:, 0] = 1.0
    lab, bx, sc = pp.postprocess_frame(probs, g['box_encodings'][:5], g['points_xyz'][:5], 'Car', 0.01)
    assert len(lab) == 0 and bx.shape == (0, 7)
