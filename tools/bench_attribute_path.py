"""Time attribute_space.path_attributes on one traversal path (33 frames of 1024 x 1024, all six predictors, seeded weights) on
ONE GPU: python tools/bench_attribute_path.py.  The face detector runs in full (network, decode, NMS); its boxes are then
replaced by a fixed plausible box, because seeded weights put them far outside the frame."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from warpedganspace_b200 import attribute_space as A
from warpedganspace_b200.eval_resnet import fairface_resnet34, hopenet_resnet50, celeba_attr_resnet50
from warpedganspace_b200.eval_sfd import SFDDetector
from warpedganspace_b200.eval_arcface import IDComparator
from warpedganspace_b200.eval_au import AUdetector

torch.manual_seed(0)
det, idc, au = SFDDetector(), IDComparator().cuda(), AUdetector()


def detector(x):
    found = det(x)
    return [[np.array([60.0, 70.0, 190.0, 200.0, 0.9], dtype=np.float32)] for _ in found]


preds = {'face_detector': detector, 'id_comparator': idc, 'au_detector': au, 'fairface': fairface_resnet34().cuda(),
         'hopenet': hopenet_resnet50().cuda(), 'celeba': celeba_attr_resnet50().cuda()}
frames = 255.0 * torch.rand(33, 3, 1024, 1024, device='cuda')
for _ in range(2):
    A.path_attributes(frames, preds, 'StyleGAN2')
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 5
for _ in range(n):
    out = A.path_attributes(frames, preds, 'StyleGAN2')
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / n * 1e3
print('path_attributes, 33 frames of 1024 x 1024, six predictors: %.1f ms per path (%.0f frames/s), keys %s'
      % (ms, 33 / ms * 1e3, sorted(out)))

# ---- where the time goes ---------------------------------------------------------------------------------------------
def timed(name, fn, n=3):
    fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        r = fn()
    torch.cuda.synchronize()
    print('  %-34s %.1f ms' % (name, (time.perf_counter() - t) / n * 1e3))
    return r


from warpedganspace_b200 import eval_sfd
small = timed('resize 1024 -> 256', lambda: A.resize_center_crop(frames, 256))
olist = timed('S3FD network', lambda: det.face_detector(small))
cand = timed('S3FD candidates (33 images, batched)', lambda: eval_sfd.candidates_batch(olist))
print('  candidates per image:', [len(c) for c in cand][:6])
timed('S3FD nms (33 images, device)', lambda: [eval_sfd.nms_device(c, 0.3) for c in cand])
timed('S3FD nms (33 images, host loop of the reference)', lambda: [eval_sfd.nms(c.cpu().numpy(), 0.3) for c in cand], n=1)
bb = [60.0, 70.0, 190.0, 200.0]
crops = timed('33 crops -> 224', lambda: torch.cat([A.resize_center_crop(A.crop_face(small, t, bb, 0.25) / 255.0, 224) for t in range(33)]))
timed('FairFace', lambda: preds['fairface'](A.normalize(crops)))
timed('Hopenet', lambda: preds['hopenet'](A.normalize(crops)))
timed('CelebA (incl. resize 1024 -> 224)', lambda: preds['celeba'](A.normalize(A.resize_center_crop(frames / 255.0 * 2 - 1, 224))))
timed('ArcFace', lambda: idc.extract_feats(small / 255.0 * 2 - 1))
timed('AU detector', lambda: au(torch.cat([A.resize_center_crop(A.crop_face(small, t, bb, 0.0), 256) for t in range(33)])))
