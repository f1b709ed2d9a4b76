"""NMS / box merging - same entry points as the reference's ``models/nms.py``
(/root/reference/models/nms.py:243-301), executed on the GPU (``pg_nms_boxes_3d``).

run.py:297-325 calls one of the four ``nms_boxes_3d*`` functions with
``overlapped_fn=nms.overlapped_boxes_3d_fast_poly`` and ``top_k=-1``; that is what is built.  The
rasterising ``overlapped_boxes_3d`` (nms.py:29-62, OpenCV fillPoly) is not.  NumPy arrays in -> NumPy
arrays out (one frame); CUDA tensors stay on the device.  For whole batches use
``postprocess.detect`` which also fuses candidate selection and box decoding.
"""
import numpy as np
import torch

from .. import _lib


def boxes_3d_to_corners(boxes_3d):
    """nms.py:9-27 (host helper: the KITTI writer projects these corners, run.py:361-370)."""
    all_corners = []
    for box_3d in boxes_3d:
        x3d, y3d, z3d, l, h, w, yaw = box_3d
        r = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
        corners = np.array([[l / 2, 0.0, w / 2], [l / 2, 0.0, -w / 2], [-l / 2, 0.0, -w / 2], [-l / 2, 0.0, w / 2],
                            [l / 2, -h, w / 2], [l / 2, -h, -w / 2], [-l / 2, -h, -w / 2], [-l / 2, -h, w / 2]])
        all_corners.append(corners.dot(np.transpose(r)) + np.array([x3d, y3d, z3d]))
    return np.array(all_corners)


def overlapped_boxes_3d_fast_poly(single_box, box_list):
    """Selector for ``overlapped_fn`` (nms.py:64-88): the convex-polygon IoU is evaluated inside the GPU NMS."""
    raise NotImplementedError('overlapped_boxes_3d_fast_poly is evaluated inside the GPU NMS kernels; pass it as '
                              'overlapped_fn to nms_boxes_3d* instead of calling it')


def overlapped_boxes_3d(single_box, box_list):
    raise NotImplementedError('the rasterising IoU (nms.py:29-62) is not built; run.py uses overlapped_boxes_3d_fast_poly')


def _run(class_labels, detection_boxes_3d, detection_scores, overlapped_thres, overlapped_fn, appr_factor, top_k,
         attributes, merge, rescore, int_corners):
    if overlapped_fn is not overlapped_boxes_3d_fast_poly:
        raise NotImplementedError('only overlapped_fn=overlapped_boxes_3d_fast_poly is built (run.py:297-325)')
    if top_k > 0:
        raise NotImplementedError('top_k > 0 is not built (run.py passes top_k=-1)')
    numpy_io = not isinstance(detection_boxes_3d, torch.Tensor)
    dev = torch.device('cuda', torch.cuda.current_device())
    boxes = torch.as_tensor(np.ascontiguousarray(detection_boxes_3d, dtype=np.float32)).to(dev) if numpy_io \
        else detection_boxes_3d.to(torch.float32).contiguous()
    labels = torch.as_tensor(np.asarray(class_labels)).to(dev).to(torch.int32).contiguous()
    scores = torch.as_tensor(np.ascontiguousarray(detection_scores, dtype=np.float32)).to(dev) if numpy_io \
        else detection_scores.to(torch.float32).contiguous()
    n = boxes.shape[0]
    if n == 0:
        return class_labels, detection_boxes_3d, detection_scores, attributes
    fp = torch.tensor([0, n], dtype=torch.int32, device=dev)
    lab, box, sc, idx, _ = _lib.nms_boxes_3d(labels, boxes, scores, fp, overlapped_thres, merge, rescore,
                                             appr_factor=appr_factor, int_corners=int_corners)
    if attributes is not None:
        attr = torch.as_tensor(np.asarray(attributes)).to(dev)[idx.long()] if numpy_io else attributes[idx.long()]
    else:
        attr = None
    if numpy_io:
        return (lab.cpu().numpy().astype(np.asarray(class_labels).dtype), box.cpu().numpy(), sc.cpu().numpy(),
                None if attr is None else attr.cpu().numpy())
    return lab, box, sc, attr


def nms_boxes_3d(class_labels, detection_boxes_3d, detection_scores, overlapped_thres=0.5,
                 overlapped_fn=overlapped_boxes_3d, appr_factor=10.0, top_k=-1, attributes=None):
    """nms.py:243-254 (bboxes_nms, nms.py:109-131: corners are converted to integer pixels * appr_factor)."""
    return _run(class_labels, detection_boxes_3d, detection_scores, overlapped_thres, overlapped_fn, appr_factor, top_k,
                attributes, merge=False, rescore=False, int_corners=True)


def nms_boxes_3d_uncertainty(class_labels, detection_boxes_3d, detection_scores, overlapped_thres=0.5,
                             overlapped_fn=overlapped_boxes_3d, appr_factor=10.0, top_k=-1, attributes=None):
    """nms.py:256-270."""
    return _run(class_labels, detection_boxes_3d, detection_scores, overlapped_thres, overlapped_fn, appr_factor, top_k,
                attributes, merge=True, rescore=True, int_corners=False)


def nms_boxes_3d_merge_only(class_labels, detection_boxes_3d, detection_scores, overlapped_thres=0.5,
                            overlapped_fn=overlapped_boxes_3d, appr_factor=10.0, top_k=-1, attributes=None):
    """nms.py:272-285."""
    return _run(class_labels, detection_boxes_3d, detection_scores, overlapped_thres, overlapped_fn, appr_factor, top_k,
                attributes, merge=True, rescore=False, int_corners=False)


def nms_boxes_3d_score_only(class_labels, detection_boxes_3d, detection_scores, overlapped_thres=0.5,
                            overlapped_fn=overlapped_boxes_3d, appr_factor=10.0, top_k=-1, attributes=None):
    """nms.py:287-301."""
    return _run(class_labels, detection_boxes_3d, detection_scores, overlapped_thres, overlapped_fn, appr_factor, top_k,
                attributes, merge=False, rescore=True, int_corners=False)
