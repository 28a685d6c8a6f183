"""TEST INFRASTRUCTURE ONLY -- plain-PyTorch restatement of the reference's per-image post-processing, the checker
of polyphonicformer_b200/csrc/pf_postprocess.cu.  Pinned against outputs of the REAL reference
(tests/golden/panoptic_*.npz, made by oracle/make_golden.py) in tests/test_oracle_golden.py.  Semantics follow:

  get_panoptic                     polyphonic/kernel_update.py:421-469
  merge_stuff_thing_stuff_joint    polyphonic/kernel_update.py:471-535
  rescale_masks / rescale_depth    polyphonic/kernel_update_head.py:593-626
  segm2result                      polyphonic/kernel_update_head.py:637-657
  depth_act                        polyphonic/funcs/depth_utils.py:1-19
  tensor_mask2box                  polyphonic/funcs/utils.py:12-22
"""
import numpy as np
import torch
import torch.nn.functional as F


def depth_act(depth_out, mode='monodepth', min_depth=0.01, max_depth=80.):
    if mode == 'monodepth':
        disp = depth_out.sigmoid()
        min_disp, max_disp = 1. / max_depth, 1. / min_depth
        return 1 / (min_disp + (max_disp - min_disp) * disp)
    if mode == 'sigmoid':
        return depth_out.sigmoid() * (max_depth - min_depth) + min_depth
    raise NotImplementedError


def _resize_crop_resize(t, img_meta):
    h, w, _ = img_meta['img_shape']
    t = F.interpolate(t, size=img_meta['batch_input_shape'], mode='bilinear', align_corners=False)
    t = t[:, :, :h, :w]
    return F.interpolate(t, size=img_meta['ori_shape'][:2], mode='bilinear', align_corners=False).squeeze(0)


def rescale_masks(masks_per_img, img_meta):
    return _resize_crop_resize(masks_per_img.unsqueeze(0).sigmoid(), img_meta)


def rescale_depth(depth, img_meta, depth_act_mode):
    return _resize_crop_resize(depth_act(depth, depth_act_mode)[None], img_meta)


def tensor_mask2box(masks):
    """(top, left, bottom, right) of the non-zero region per mask; (-1, -1, 10, 10) for an empty mask."""
    boxes = []
    for mask in masks:
        m = mask.nonzero()
        if m.numel() > 0:
            boxes.append((m[:, 1].min().item(), m[:, 0].min().item(), m[:, 1].max().item(), m[:, 0].max().item()))
        else:
            boxes.append((-1, -1, 10, 10))
    return np.asarray(boxes)


def segm2result(num_classes, mask_preds, det_labels, cls_scores, depth_preds):
    segm_result = [[] for _ in range(num_classes)]
    depth_result = [[] for _ in range(num_classes)]
    det_labels = det_labels.cpu().numpy()
    cls_scores = cls_scores.cpu().numpy()
    depth_preds = depth_preds.cpu().numpy()
    num_ins = mask_preds.shape[0]
    bboxes = np.zeros((num_ins, 5), dtype=np.float32)
    bboxes[:, -1] = cls_scores
    if num_ins:
        bboxes[:, :4] = np.array(tensor_mask2box(mask_preds).clip(min=0))
    mask_np = mask_preds.cpu().numpy()
    for idx in range(num_ins):
        segm_result[det_labels[idx]].append(mask_np[idx])
        depth_result[det_labels[idx]].append(depth_preds[idx])
    return bboxes, segm_result, depth_result


def merge_stuff_thing_stuff_joint(num_thing_classes, thing_masks, thing_labels, thing_scores, stuff_masks,
                                  stuff_labels, stuff_scores, merge_cfg, depth_all=None, depth_things=None,
                                  depth_stuff=None):
    H, W = thing_masks.shape[-2:]
    panoptic_seg = thing_masks.new_zeros((H, W), dtype=torch.int32)
    total_masks = torch.cat([thing_masks, stuff_masks], dim=0)
    total_scores = torch.cat([thing_scores, stuff_scores], dim=0)
    total_labels = torch.cat([thing_labels, stuff_labels], dim=0)
    total_depth = torch.cat([depth_things, depth_stuff], dim=0)
    cur_mask_ids = (total_scores.view(-1, 1, 1) * total_masks).argmax(0)
    segments_info = []
    sorted_inds = torch.argsort(-total_scores)
    # one pass of counting instead of two reductions per kernel inside the loop (same values)
    K = total_masks.shape[0]
    mask_areas = torch.bincount(cur_mask_ids.flatten(), minlength=K).cpu().tolist()
    orig_areas = (total_masks >= 0.5).flatten(1).sum(1).cpu().tolist()
    labels = total_labels.cpu().tolist()
    scores = total_scores.cpu().tolist()
    current_segment_id = 0
    for k in sorted_inds.cpu().tolist():
        pred_class = labels[k]
        isthing = pred_class < num_thing_classes
        if isthing and scores[k] < merge_cfg.instance_score_thr:
            continue
        mask_area, original_area = mask_areas[k], orig_areas[k]
        if mask_area > 0 and original_area > 0:
            if mask_area / original_area < merge_cfg.overlap_thr:
                continue
            current_segment_id += 1
            mask = cur_mask_ids == k
            panoptic_seg[mask] = current_segment_id
            if depth_all is not None:
                depth_all[mask] = total_depth[k][mask]
            if isthing:
                segments_info.append({'id': current_segment_id, 'isthing': isthing, 'score': scores[k],
                                      'category_id': pred_class, 'instance_id': k})
            else:
                segments_info.append({'id': current_segment_id, 'isthing': isthing, 'category_id': pred_class,
                                      'area': mask_area})
    return panoptic_seg.cpu().numpy(), segments_info


def get_panoptic(roi_head, last_head, cls_scores, mask_preds, test_cfg, img_meta, depth_preds, depth_init,
                 aspp_semantic=None):
    P, T = roi_head.num_proposals, roi_head.num_thing_classes
    mode = last_head.depth_act_mode
    depth_pred = rescale_depth(depth_preds, img_meta, mode)
    depth_init = rescale_depth(depth_init, img_meta, mode)
    thing_scores = cls_scores[:P][:, :T]
    thing_mask_preds = mask_preds[:P]
    thing_scores, topk_indices = thing_scores.flatten(0, 1).topk(test_cfg.max_per_img, sorted=True)
    mask_indices = topk_indices // T
    thing_labels = topk_indices % T
    thing_masks = rescale_masks(thing_mask_preds[mask_indices], img_meta)
    if not roi_head.merge_joint:
        thing_masks = thing_masks > test_cfg.mask_thr
    depth_pred_things = depth_pred[:P][mask_indices]
    depth_pred_stuff = depth_pred[P:]
    depth_final = depth_init.squeeze(0)
    depth_basic = depth_final.clone()
    stuff_scores = cls_scores[P:][:, T:].diag()
    stuff_scores, stuff_inds = torch.sort(stuff_scores, descending=True)
    stuff_masks = rescale_masks(mask_preds[P:][stuff_inds], img_meta)
    if not roi_head.merge_joint:
        raise NotImplementedError
    depth_pred_stuff = depth_pred_stuff[stuff_inds]
    stuff_labels = stuff_inds + T
    panoptic_result = merge_stuff_thing_stuff_joint(T, thing_masks, thing_labels, thing_scores, stuff_masks,
                                                    stuff_labels, stuff_scores, test_cfg.merge_stuff_thing,
                                                    depth_final, depth_pred_things, depth_pred_stuff)
    return None, None, panoptic_result, depth_basic.cpu().numpy(), depth_final.cpu().numpy()
