"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's data-format code either side of the hot path
(SURVEY 8f ranks 2-3); the checker of bts_b200/csrc/io.cu.  Only tests/ imports this.

  input_prep    pytorch/bts_dataloader.py:128-140 (np.asarray(image, float32)/255, depth /1000 | /256, random_crop),
                :202-214 (train_preprocess: flip), :216-235 (augment_image), :244-249 (ToTensor + Normalize)
  eval_errors   pytorch/bts_main.py:144-165 (compute_errors) after the clamps / masks of :275-296
  depth_to_u16  pytorch/bts_test.py:179-183
Pinned against the live reference functions where /root/reference (or oracle/_ref) is available: tests/test_io_oracle.py.
"""
import numpy as np

MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32)
STD = np.array([0.229, 0.224, 0.225], dtype=np.float32)


def augment_image(image, gamma, brightness, colors):
    """bts_dataloader.py:216-235 with the random draws passed in"""
    image_aug = image ** gamma
    image_aug = image_aug * brightness
    white = np.ones((image.shape[0], image.shape[1]))
    color_image = np.stack([white * colors[i] for i in range(3)], axis=2)
    image_aug *= color_image
    return np.clip(image_aug, 0, 1)


def input_prep(img_u8, depth_u16, depth_div, y0, x0, H, W, flip, augment, gamma, brightness, colors):
    """one sample: (H,W,3) uint8 frame (+ (H,W) uint16 depth) -> normalised CHW fp32 image, (1,H,W) fp32 depth"""
    image = np.asarray(img_u8, dtype=np.float32) / 255.0                     # :128
    depth = None
    if depth_u16 is not None:
        depth = np.expand_dims(np.asarray(depth_u16, dtype=np.float32), 2) / depth_div      # :129-135
    image = image[y0:y0 + H, x0:x0 + W, :]                                   # random_crop :196-199
    if depth is not None:
        depth = depth[y0:y0 + H, x0:x0 + W, :]
    if flip:                                                                 # :204-207
        image = (image[:, ::-1, :]).copy()
        if depth is not None:
            depth = (depth[:, ::-1, :]).copy()
    if augment:                                                              # :210-212
        image = augment_image(image, gamma, brightness, colors).astype(np.float32)
    image = image.transpose((2, 0, 1))                                       # ToTensor :262
    image = (image - MEAN[:, None, None]) / STD[:, None, None]               # Normalize :244
    return image.astype(np.float32), None if depth is None else depth.transpose((2, 0, 1)).astype(np.float32)


def compute_errors(gt, pred):
    """bts_main.py:144-165"""
    thresh = np.maximum((gt / pred), (pred / gt))
    d1 = (thresh < 1.25).mean()
    d2 = (thresh < 1.25 ** 2).mean()
    d3 = (thresh < 1.25 ** 3).mean()
    rms = np.sqrt(((gt - pred) ** 2).mean())
    log_rms = np.sqrt(((np.log(gt) - np.log(pred)) ** 2).mean())
    abs_rel = np.mean(np.abs(gt - pred) / gt)
    sq_rel = np.mean(((gt - pred) ** 2) / gt)
    err = np.log(pred) - np.log(gt)
    silog = np.sqrt(np.mean(err ** 2) - np.mean(err) ** 2) * 100
    log10 = np.mean(np.abs(np.log10(pred) - np.log10(gt)))
    return [silog, abs_rel, log10, rms, sq_rel, log_rms, d1, d2, d3]


def eval_errors(pred, gt, min_depth, max_depth, crop=None):
    """bts_main.py:275-298: clamp the prediction, mask, metrics of the valid pixels; crop = (y0,y1,x0,x1)"""
    pred = pred.copy()
    pred[pred < min_depth] = min_depth
    pred[pred > max_depth] = max_depth
    pred[np.isinf(pred)] = max_depth
    pred[np.isnan(pred)] = min_depth
    valid = np.logical_and(gt > min_depth, gt < max_depth)
    if crop is not None:
        m = np.zeros(valid.shape)
        m[crop[0]:crop[1], crop[2]:crop[3]] = 1
        valid = np.logical_and(valid, m)
    return compute_errors(gt[valid], pred[valid]), int(valid.sum())


def depth_to_u16(depth, scale):
    """bts_test.py:179-183"""
    return (depth * scale).astype(np.uint16)
