"""Attribute-space traversal (reference: traverse_attribute_space.py:140-613): walk the image sequences written by the
latent-space traversal and record how each attribute predictor responds along every path.

Built here (the consumer of the traversal frames):
  * the driver - directory conventions, per-path batches, the reference's score arithmetic, the ``eval_json`` / ``eval_np``
    files with the reference's names and layouts;
  * the three predictors that are ImageNet-style ResNets, as inference-only kernel chains on libwgs_b200
    (eval_resnet.fairface_resnet34 / hopenet_resnet50 / celeba_attr_resnet50: age / race / gender, yaw / pitch / roll and the
    five CelebA attributes);
  * the S3FD face detector (eval_sfd.SFDDetector: VGG trunk + heads on the same convs, soft-max / anchor decode / NMS as the
    reference's batch_detect);
  * face cropping, resize + centre crop + normalisation on the device.
  * the ArcFace identity comparator (eval_arcface.IDComparator, IR-SE-50) and the action-unit detector
    (eval_au.AUdetector, face-alignment network + lightweight hourglass).
Every predictor is optional (``predictors[...]`` holds callables with the reference's call signatures, so a reference module
can stand in for any of them); files of absent predictors are not written.  Without a face detector every frame uses the
reference's own no-detection fallback (the full 256 x 256 frame, traverse_attribute_space.py:396-399).
"""
import glob
import json
import os
import os.path as osp

import numpy as np
import torch
import torch.nn.functional as F

AUs = {"au_1": "Inner_Brow_Raiser", "au_2": "Outer_Brow_Raiser", "au_4": "Brow_Lowerer", "au_5": "Upper_Lid_Raiser",
       "au_6": "Cheek_Raiser", "au_9": "Nose_Wrinkler", "au_12": "Lip_Corner_Puller", "au_15": "Lip_Corner_Depressor",
       "au_17": "Chin_Raiser", "au_20": "Lip_stretcher", "au_25": "Lips_part", "au_26": "Jaw_Drop"}
_MEAN, _STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def crop_face(images, idx, bbox, padding=0.0):
    """traverse_attribute_space.py:36-58, including its axis convention (bbox x indexes dim 2)."""
    x_min = int((1.0 - padding) * bbox[0]) - 50
    y_min = int((1.0 - padding) * bbox[1]) - 50
    x_max = int((1.0 + padding) * bbox[2]) + 50
    y_max = int((1.0 + padding) * bbox[3]) + 30
    x_min, y_min = max(x_min, 0), max(y_min, 0)
    x_max, y_max = min(images.shape[2], x_max), min(images.shape[3], y_max)
    return images[idx, :, int(x_min):int(x_max), int(y_min):int(y_max)].unsqueeze(0)


def resize_center_crop(x, size):
    """transforms.Compose([Resize(size), CenterCrop(size)]) on a float [N, C, H, W] tensor (bilinear, antialias as torchvision
    does for tensors: off for up-sampling, on for down-sampling)."""
    h, w = x.shape[-2:]
    if h <= w:
        nh, nw = size, max(size, int(size * w / h))
    else:
        nh, nw = max(size, int(size * h / w)), size
    if (nh, nw) != (h, w):
        x = F.interpolate(x, size=(nh, nw), mode='bilinear', align_corners=False, antialias=True)
    top, left = int(round((nh - size) / 2.0)), int(round((nw - size) / 2.0))
    return x[..., top: top + size, left: left + size]


def normalize(x):
    mean = torch.tensor(_MEAN, device=x.device).view(1, 3, 1, 1)
    std = torch.tensor(_STD, device=x.device).view(1, 3, 1, 1)
    return (x - mean) / std


def load_path_images(path_dir, device):
    """lib/data.py:9-25 (PathImages): the sorted *.jpg of one path as a float [T, 3, H, W] tensor in 0..255, RGB."""
    from PIL import Image
    files = sorted(glob.glob(osp.join(path_dir, '*.jpg')))
    if not files:
        raise FileNotFoundError('no frames under %s' % path_dir)

    def decode(f):                      # PIL's decoder (the reference's pixels exactly); it releases the GIL while decoding
        with Image.open(f) as im:
            return torch.from_numpy(np.asarray(im.convert('RGB'), dtype=np.uint8).copy())

    if len(files) > 2:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(files), os.cpu_count() or 1, 16)) as pool:
            frames = list(pool.map(decode, files))
    else:
        frames = [decode(f) for f in files]
    return torch.stack(frames).to(device).permute(0, 3, 1, 2).float()


def fairface_scores(outputs):
    """traverse_attribute_space.py:412-433 -> (femaleness, age, race) numpy vectors."""
    o = outputs.double()
    gender = torch.softmax(o[:, 7:9], dim=1)[:, 1]
    age_s, race_s = torch.softmax(o[:, 9:18], dim=1), torch.softmax(o[:, :7], dim=1)
    age = (age_s.argmax(dim=1) + age_s.max(dim=1).values) / 9.0
    race = (race_s.argmax(dim=1) + race_s.max(dim=1).values) / 7.0
    return gender.cpu().numpy(), age.cpu().numpy(), race.cpu().numpy()


def hopenet_pose(yaw, pitch, roll):
    """traverse_attribute_space.py:448-456: soft-argmax over the 66 bins, degrees."""
    idx = torch.arange(66, dtype=torch.float32, device=yaw.device)
    return tuple((torch.sum(torch.softmax(t.float(), dim=1) * idx, 1) * 3 - 99) for t in (yaw, pitch, roll))


def path_attributes(frames, predictors, gan_type='StyleGAN2'):
    """One path: frames float [T, 3, H, W] in 0..255 -> dict of per-frame attribute lists (the rows the reference writes)."""
    T = frames.shape[0]
    out = {}
    small = resize_center_crop(frames, 256)                                        # face_detector_trans, :168
    det = predictors.get('face_detector')
    with torch.no_grad():
        detected = det(small) if det is not None else [[] for _ in range(T)]       # :318-319
    boxes, fw, fh = [], [], []
    for t in range(T):
        if len(detected[t]) > 0:
            bb = [float(v) for v in detected[t][0]]
            boxes.append(bb)
            fw.append((bb[2] - bb[0]) / 256.0)
            fh.append((bb[3] - bb[1]) / 256.0)
        else:
            fw.append(256.0)                                                       # (sic) :335-336
            fh.append(256.0)
    out['face_bbox'], out['face_width'], out['face_height'] = boxes, fw, fh

    def crops(size, padding, scale):
        faces = []
        for t in range(T):
            bb = detected[t][0][:-1] if len(detected[t]) > 0 else [0, 0, 256, 256]  # :396-399
            faces.append(resize_center_crop(crop_face(small, t, bb, padding) * scale, size))
        return torch.cat(faces, dim=0)

    celeba = predictors.get('celeba')
    if celeba is not None:                                                         # :341-372
        if gan_type == 'StyleGAN2':
            inp = frames / 255.0 * 2.0 - 1.0
        else:
            inp = (frames - frames.min()) / (frames.max() - frames.min())
        with torch.no_grad():
            preds = celeba(normalize(resize_center_crop(inp, 224)))
        for attr, key in (('Bangs', 'celeba_bangs'), ('Eyeglasses', 'celeba_eyeglasses'), ('No_Beard', 'celeba_beard'),
                          ('Smiling', 'celeba_smiling'), ('Young', 'celeba_age')):
            s = torch.softmax(preds[attr].float(), dim=1)
            out[key] = ((s.argmax(dim=1) + s.max(dim=1).values) / 6.0).cpu().numpy().tolist()
    idc = predictors.get('id_comparator')
    if idc is not None:                                                            # :374-392: similarity to the centre frame
        with torch.no_grad():
            if hasattr(idc, 'extract_feats'):                                      # one batched pass instead of T pairs
                feats = idc.extract_feats(small / 255.0 * 2.0 - 1.0)
                out['identity'] = F.cosine_similarity(feats[T // 2: T // 2 + 1], feats, dim=1, eps=1e-6).cpu().numpy().tolist()
            else:
                ref_img = small[T // 2: T // 2 + 1] / 255.0 * 2.0 - 1.0
                out['identity'] = [float(idc(ref_img, small[t: t + 1] / 255.0 * 2.0 - 1.0)) for t in range(T)]
    ff = predictors.get('fairface')
    if ff is not None:                                                             # :394-433
        with torch.no_grad():
            logits = ff(normalize(crops(224, 0.25, 1.0 / 255.0)))
        g, a, r = fairface_scores(logits)
        out['gender'], out['age'], out['race'] = g.tolist(), a.tolist(), r.tolist()
    hp = predictors.get('hopenet')
    if hp is not None:                                                             # :435-463
        with torch.no_grad():
            yaw, pitch, roll = hopenet_pose(*hp(normalize(crops(224, 0.0, 1.0 / 255.0))))
        out['pose'] = [yaw.cpu().numpy().tolist(), pitch.cpu().numpy().tolist(), roll.cpu().numpy().tolist()]
    au = predictors.get('au_detector')
    if au is not None:                                                             # :465-527
        with torch.no_grad():
            inten = au(crops(256, 0.0, 1.0)).detach().float().cpu().numpy().transpose()
        out['au'] = [inten[t].tolist() for t in range(len(AUs))]
    return out


def traverse_attribute_space(exp, pool, shift_steps=16, eps=0.2, predictors=None, gan_type=None, device='cuda', verbose=False):
    """The reference's main loop (traverse_attribute_space.py:226-603) over
    <exp>/results/<pool>/<2*steps>_<eps>_<len>/<hash>/paths_images/path_XXX/*.jpg.  Returns the list of hash directories
    processed; per hash it writes eval_json/*.json and eval_np/*.npy for the attributes of the predictors given."""
    predictors = predictors or {}
    if gan_type is None:
        with open(osp.join(exp, 'args.json')) as f:
            gan_type = json.load(f)['gan_type']
    cfg = '{}_{}_{}'.format(2 * shift_steps, eps, round(2 * shift_steps * eps, 3))
    hashes_dir = osp.join(exp, 'results', pool, cfg)
    if not osp.isdir(hashes_dir):
        raise NotADirectoryError('Error: traversal directory {} not found'.format(hashes_dir))
    done = []
    for h in sorted(d for d in os.listdir(hashes_dir)
                    if osp.isdir(osp.join(hashes_dir, d)) and d not in ('paths_gifs', 'validation_results')):
        h_dir = osp.join(hashes_dir, h)
        codes = torch.load(osp.join(h_dir, 'paths_latent_codes.pt'), map_location='cpu')
        n_paths, n_img = codes.shape[0], codes.shape[1]
        rows = {}
        for d in range(n_paths):
            frames = load_path_images(osp.join(h_dir, 'paths_images', 'path_{:03d}'.format(d)), device)
            if frames.shape[0] != n_img:
                raise RuntimeError('path %d of %s: %d frames on disk, %d latent codes' % (d, h, frames.shape[0], n_img))
            for k, v in path_attributes(frames, predictors, gan_type).items():
                rows.setdefault(k, {})[d] = v
            if verbose:
                print('  {}: path {:03d}/{:03d}'.format(h, d + 1, n_paths))
        jd, nd = osp.join(h_dir, 'eval_json'), osp.join(h_dir, 'eval_np')
        os.makedirs(jd, exist_ok=True)
        os.makedirs(nd, exist_ok=True)

        def table(key):
            return np.array([rows[key][d] for d in range(n_paths)], dtype=np.float64)

        with open(osp.join(jd, 'face_bbox.json'), 'w') as f:
            json.dump(rows['face_bbox'], f)
        np.save(osp.join(nd, 'face_width.npy'), table('face_width'))
        np.save(osp.join(nd, 'face_height.npy'), table('face_height'))
        for key in ('identity', 'age', 'race', 'gender', 'celeba_bangs', 'celeba_eyeglasses', 'celeba_beard', 'celeba_smiling',
                    'celeba_age'):
            if key in rows:
                with open(osp.join(jd, key + '.json'), 'w') as f:
                    json.dump(rows[key], f)
                np.save(osp.join(nd, key + '.npy'), table(key))
        if 'pose' in rows:
            with open(osp.join(jd, 'pose.json'), 'w') as f:
                json.dump(rows['pose'], f)
            for i, name in enumerate(('yaw', 'pitch', 'roll')):                    # radians in the .npy files, :461-463
                np.save(osp.join(nd, name + '.npy'), np.array([rows['pose'][d][i] for d in range(n_paths)]) * np.pi / 180)
        if 'au' in rows:
            with open(osp.join(jd, 'au.json'), 'w') as f:
                json.dump(rows['au'], f)
            for t, k in enumerate(AUs):
                np.save(osp.join(nd, '{}_{}.npy'.format(k, AUs[k])), np.array([rows['au'][d][t] for d in range(n_paths)]))
        done.append(h_dir)
    return done
