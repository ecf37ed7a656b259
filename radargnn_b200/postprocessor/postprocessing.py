"""``BoxSuppressor`` on box matrices (reference postprocessor/postprocessing.py:336-435).

The reference wraps every box in a ``BoundingBox`` object and converts the list to a matrix before it calls
``torchvision.ops.nms`` / detectron2's ``nms_rotated`` on the CPU; those object lists are dataset plumbing (out of
scope).  The mirror starts at the matrix: same arguments from there on, same shift of negative
coordinates, same order of the kept boxes (descending score), computed by ``rgnn_nms`` on the GPU."""
from __future__ import annotations

import numpy as np
import torch

from .. import ops


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("radargnn_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


class BoxSuppressor:
    """Non-maximum suppression with aligned and rotated bounding boxes (postprocessing.py:336)."""

    @classmethod
    def apply_nms(cls, bounding_box_matrix, box_scores, box_labels, iou_nms: float, is_rotated: bool):
        """bounding_box_matrix: [n, 4] two-point representation (x_min, y_min, x_max, y_max) or, rotated, [n, 5]
        absolute representation (x, y, l, w, theta in degrees); box_scores [n, 1]; box_labels [n] or [n, 1].
        Returns (matrix, scores [m, 1], labels [m, 1]) of the kept boxes, like the reference (:346-353)."""
        m = np.asarray(bounding_box_matrix)
        s = np.asarray(box_scores).reshape(-1)
        lab = np.asarray(box_labels).reshape(-1)
        if m.shape[0] == 0:
            return m, s.reshape(0, 1), lab.reshape(0, 1)
        keep = cls.keep_indices(m, s, iou_nms, is_rotated).cpu().numpy()
        return m[keep], s[keep].reshape(-1, 1), lab[keep].reshape(-1, 1)

    @staticmethod
    def keep_indices(bounding_box_matrix, box_scores, iou_nms: float, is_rotated: bool, box_frame=None) -> torch.Tensor:
        dev = _dev()
        dt = torch.float64 if is_rotated else torch.float32   # postprocessing.py:368 / :407
        boxes = torch.as_tensor(np.asarray(bounding_box_matrix), dtype=dt).to(dev)
        scores = torch.as_tensor(np.asarray(box_scores).reshape(-1), dtype=dt).to(dev)
        fr = None if box_frame is None else torch.as_tensor(np.asarray(box_frame), dtype=torch.int32).to(dev)
        return ops.nms(boxes, scores, iou_nms, rotated=is_rotated, box_frame=fr, shift_negative=True)


def nearest_neighbor_positions(pos, frame_ptr=None) -> np.ndarray:
    """``pos[np.where(kneighbors_graph(pos, 1).toarray() == 1)[1]]`` (postprocessing.py:233-237, :468-472)."""
    dev = _dev()
    p = torch.as_tensor(np.asarray(pos, dtype=np.float64)).to(dev)
    _, pts = ops.nearest_neighbor(p, frame_ptr)
    return pts.cpu().numpy()
