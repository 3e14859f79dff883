"""Mask refinement between Phase A and Phase B - everything AROUND the SAM network call.

The reference's models/sam.py does three things per box: (1) builds a prompt for SAM (the box in pixels for LMD+,
sam_refine_boxes :189-193; a smoothed / thresholded token-attention map turned into a box or its arg-max turned into a
point for LMD, sam_refine_attn :126-155), (2) runs facebook/sam-vit-base, (3) post-processes its three candidate masks:
bilinear resize to the latent resolution + `.bool()` (:50-53), IoU against the coarse mask (:63-65) and the
"largest_over_conf" selection rule (:67-111).  (2) is a third-party network and stays an environment hook here:

    predict(image, input_boxes=None, input_points=None) -> (masks, iou_scores)
        input_boxes / input_points: exactly what the reference hands to SamProcessor - [[[x0, y0, x1, y1]]] from
                     refine_box, [[x0, y0, x1, y1]] from refine_attn (box input), [[[x, y]]] (point input)
        masks      : [3, h, w] bool / float array or tensor at IMAGE resolution
                     (= sam_processor.image_processor.post_process_masks(...)[0][0])
        iou_scores : [3]   (= outputs.iou_scores[0, 0])

(1) and (3) are this module: small integer / boolean host logic on 64x64 maps (numpy + scipy.ndimage + cv2, the same
libraries the reference uses for it), bit-exact against the UNMODIFIED reference functions in
tests/test_oracle_vs_reference.py.  `ReferenceEnv(model_dict, sam_predict=...)` plugs it in as `refine_mask`.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import guidance as G

# module-level defaults of generation/lmd.py:39-49 and generation/lmd_plus.py:36-39
GAUSSIAN_SIGMA_POINT_INPUT = 1.5
GAUSSIAN_SIGMA_BOX_INPUT = 0.1
MASK_TH_FOR_BOX = 0.05
N_ERODE_DILATE_MASK_FOR_BOX = 1
MASK_TH_FOR_POINT = 0.25
DISCOURAGE_MASK_BELOW_CONFIDENCE = 0.85
DISCOURAGE_MASK_BELOW_COARSE_IOU_LMD = 0.25
DISCOURAGE_MASK_BELOW_COARSE_IOU_LMD_PLUS = 0.2


def resize_masks(masks, target_shape):
    """models/sam.py:50-53: F.interpolate(masks.float(), (H, W), mode='bilinear').bool() - any positive interpolated value
    is foreground.  masks [3, h, w] -> numpy bool [3, H, W]"""
    m = torch.as_tensor(np.asarray(masks)) if not isinstance(masks, torch.Tensor) else masks
    m = m.detach().to("cpu", torch.float32)[None]
    return F.interpolate(m, tuple(target_shape), mode="bilinear")[0].type(torch.bool).numpy()


def iou(mask, masks, eps=1e-6):
    """utils/utils.py:124-131: intersection over (union + eps) of one mask against n candidates"""
    ref = np.asarray(mask).astype(bool)[None]
    cand = np.asarray(masks).astype(bool)
    inter = np.logical_and(ref, cand).reshape(len(cand), -1).sum(axis=1)
    union = np.logical_or(ref, cand).reshape(len(cand), -1).sum(axis=1)
    return inter / (union + eps)


def iou_with_resize(mask, masks, masks_shape):
    """models/sam.py:63-65: every candidate is resized to the coarse mask's shape with cv2 (INTER_LINEAR on uint8 * 255,
    then non-zero) before the IoU"""
    import cv2
    masks = np.array([cv2.resize(m.astype(np.uint8) * 255, tuple(masks_shape[::-1]), cv2.INTER_LINEAR).astype(bool)
                      for m in masks])
    return iou(mask, masks)


def select_mask(masks, conf_scores, coarse_ious=None, discourage_mask_below_confidence=DISCOURAGE_MASK_BELOW_CONFIDENCE,
                discourage_mask_below_coarse_iou=DISCOURAGE_MASK_BELOW_COARSE_IOU_LMD_PLUS):
    """models/sam.py:67-111, rule "largest_over_conf": the largest candidate wins; a candidate whose predicted IoU
    (confidence) or whose IoU with the coarse mask is below its threshold is pushed back by the largest mask size
    (each).  Returns (mask [H, W] bool, its confidence)."""
    masks = np.asarray(masks)
    conf_scores = np.asarray(conf_scores)
    mask_sizes = masks.sum(axis=(1, 2))
    max_mask_size = np.max(mask_sizes)
    scores = mask_sizes - (conf_scores < discourage_mask_below_confidence) * max_mask_size
    if coarse_ious is not None:
        scores = scores - (np.asarray(coarse_ious) < discourage_mask_below_coarse_iou) * max_mask_size
    mask_id = int(np.argmax(scores))
    return masks[mask_id], conf_scores[mask_id]


def preprocess_mask(token_attn_smooth, mask_th, n_erode_dilate_mask=0):
    """models/sam.py:113-122: min-max normalise, threshold, optional binary opening"""
    from scipy import ndimage
    a = token_attn_smooth - token_attn_smooth.min()
    a = a / a.max()
    m = a > mask_th
    if n_erode_dilate_mask:
        m = ndimage.binary_erosion(m, iterations=n_erode_dilate_mask)
        m = ndimage.binary_dilation(m, iterations=n_erode_dilate_mask)
    return m


def binary_mask_to_box(mask, enlarge_box_by_one=True, w_scale=1, h_scale=1):
    """utils/utils.py:72-88: bounding box [xmin, ymin, xmax, ymax] of the set cells (inclusive max, grown by one cell and
    clipped when enlarge_box_by_one), scaled to pixels.  An empty mask raises ValueError (min() of an empty sequence in
    the reference)."""
    rows, cols = np.nonzero(mask)
    if rows.size == 0:
        raise ValueError("binary_mask_to_box: the mask is empty")
    n_rows, n_cols = mask.shape
    grow = 1 if enlarge_box_by_one else 0
    ymin, ymax = int(rows.min()) - grow, int(rows.max()) + grow
    xmin, xmax = int(cols.min()) - grow, int(cols.max()) + grow
    if enlarge_box_by_one:
        ymin, ymax = max(ymin, 0), min(ymax, n_rows)
        xmin, xmax = max(xmin, 0), min(xmax, n_cols)
    return [xmin * w_scale, ymin * h_scale, xmax * w_scale, ymax * h_scale]


def attn_prompt(token_attn, height, width, use_box_input=False, gaussian_sigma=None, mask_th_for_box=MASK_TH_FOR_BOX,
                n_erode_dilate_mask_for_box=N_ERODE_DILATE_MASK_FOR_BOX, mask_th_for_point=MASK_TH_FOR_POINT):
    """models/sam.py:126-155 up to the SAM call: returns (coarse binary mask at the attention resolution,
    dict(input_boxes=...) or dict(input_points=...) in image pixels).  Keeps the reference's scale quirk: the (w, h)
    scale pair is computed as (height // map_width, width // map_height) - identical for square maps."""
    from scipy import ndimage
    token_attn = np.asarray(token_attn)
    if gaussian_sigma is None:
        gaussian_sigma = GAUSSIAN_SIGMA_BOX_INPUT if use_box_input else GAUSSIAN_SIGMA_POINT_INPUT
    smooth = ndimage.gaussian_filter(token_attn.astype(float), sigma=gaussian_sigma)
    mask_size_scale = height // smooth.shape[1], width // smooth.shape[0]
    if use_box_input:
        mask_binary = preprocess_mask(smooth, mask_th_for_box, n_erode_dilate_mask=n_erode_dilate_mask_for_box)
        box = binary_mask_to_box(mask_binary, w_scale=mask_size_scale[0], h_scale=mask_size_scale[1])
        return mask_binary, dict(input_boxes=[box])      # one nesting level less than sam_refine_boxes, like the reference
    mask_binary = preprocess_mask(smooth, mask_th_for_point, n_erode_dilate_mask=0)
    max_coord = np.unravel_index(smooth.argmax(), smooth.shape)
    return mask_binary, dict(input_points=[[[max_coord[1] * mask_size_scale[1], max_coord[0] * mask_size_scale[0]]]])


def refine_attn(predict, image, token_attn, height, width, H, W, use_box_input=False, gaussian_sigma=None,
                mask_th_for_box=MASK_TH_FOR_BOX, n_erode_dilate_mask_for_box=N_ERODE_DILATE_MASK_FOR_BOX,
                mask_th_for_point=MASK_TH_FOR_POINT,
                discourage_mask_below_confidence=DISCOURAGE_MASK_BELOW_CONFIDENCE,
                discourage_mask_below_coarse_iou=DISCOURAGE_MASK_BELOW_COARSE_IOU_LMD):
    """models/sam.py:126-176 sam_refine_attn (LMD): prompt from the token attention, three candidates, selection"""
    mask_binary, prompt = attn_prompt(token_attn, height, width, use_box_input, gaussian_sigma, mask_th_for_box,
                                      n_erode_dilate_mask_for_box, mask_th_for_point)
    masks, conf = predict(image, **prompt)
    three = resize_masks(masks, (H, W))
    ious = iou_with_resize(mask_binary, three, masks_shape=mask_binary.shape)
    return select_mask(three, np.asarray(conf), ious, discourage_mask_below_confidence, discourage_mask_below_coarse_iou)


def refine_box(predict, image, box, height, width, H, W,
               discourage_mask_below_confidence=DISCOURAGE_MASK_BELOW_CONFIDENCE,
               discourage_mask_below_coarse_iou=DISCOURAGE_MASK_BELOW_COARSE_IOU_LMD_PLUS):
    """models/sam.py:178-213 sam_refine_box(es) (LMD+): the box in pixels prompts SAM, the box raster at the latent
    resolution is the coarse mask"""
    px = G.scale_proportion(box, height, width)
    masks, conf = predict(image, input_boxes=[[list(px)]])
    three = resize_masks(masks, (H, W))
    x0, y0, x1, y1 = G.scale_proportion(box, H, W)
    mask_binary = np.zeros((H, W))
    mask_binary[y0:y1, x0:x1] = 1.0
    ious = iou_with_resize(mask_binary, three, masks_shape=mask_binary.shape)
    return select_mask(three, np.asarray(conf), ious, discourage_mask_below_confidence, discourage_mask_below_coarse_iou)
