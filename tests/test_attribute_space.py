"""Attribute-space traversal: oracle of the two ResNet predictors against the reference-pinned fixture
(CPU), the driver's score arithmetic / file layout with stub predictors (CPU), and the B200 kernel chains against the same
fixture (GPU)."""
import json
import os

import numpy as np
import pytest
import torch

import oracle.eval_nets as o_en

HEADS = {'fc_yaw': 66, 'fc_pitch': 66, 'fc_roll': 66}


def gen(seed):
    return torch.Generator().manual_seed(seed)


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def test_oracle_resnets_match_the_reference_fixture(golden):
    fx = golden('eval_nets.pt')
    x = torch.randn(2, 3, 224, 224, generator=gen(fx['seed_x']))
    sd = o_en.init_state('basic', {'fc': 18}, gen(fx['fairface']['seed']))
    assert rel(o_en.resnet_forward(sd, x, 'basic', ('fc',))[0], fx['fairface']['out']) < 1e-5
    sd = o_en.init_state('bottleneck', HEADS, gen(fx['hopenet']['seed']))
    for got, want in zip(o_en.resnet_forward(sd, x, 'bottleneck', tuple(HEADS)), fx['hopenet']['out']):
        assert rel(got, want) < 1e-5
    sd = o_en.init_celeba_state(gen(fx['celeba']['seed']))
    got = o_en.celeba_forward(sd, x)
    assert list(got) == ['Bangs', 'Eyeglasses', 'No_Beard', 'Smiling', 'Young']
    for name, want in fx['celeba']['out'].items():
        assert rel(got[name], want) < 1e-5


def test_oracle_s3fd_matches_the_reference_fixture(golden):
    fx = golden('eval_nets.pt')['s3fd']
    sd = o_en.init_s3fd_state(gen(fx['seed']))
    x = 255.0 * torch.rand(2, 3, 128, 128, generator=gen(fx['seed_x']))
    outs = o_en.s3fd_forward(sd, x)
    for got, want in zip(outs, fx['out']):
        assert rel(got[:, :, ::2, ::2], want) < 1e-5
    dets = o_en.sfd_detect_from_batch(outs)
    for d, want in zip(dets, fx['detections']):
        assert d.shape == tuple(want.shape)
        a, b = d[np.lexsort(d.T)], want.numpy()[np.lexsort(want.numpy().T)]
        assert np.allclose(a, b, rtol=1e-4, atol=1e-3)


def test_oracle_arcface_and_au_match_the_reference_fixture(golden):
    fx = golden('eval_nets.pt')
    sd = o_en.init_arcface_state(gen(fx['arcface']['seed']))
    xa = torch.rand(3, 3, 256, 256, generator=gen(fx['arcface']['seed_x'])) * 2 - 1
    xa[2] = 0.7 * xa[0] + 0.3 * xa[1]
    assert rel(o_en.arcface_extract_feats(sd, xa), fx['arcface']['feats']) < 1e-5
    sims = torch.stack([o_en.id_similarity(sd, xa[0:1], xa[t: t + 1]) for t in range(3)])
    assert rel(sims, fx['arcface']['sim']) < 1e-5
    sd = o_en.init_au_state(gen(fx['au']['seed']))
    xu = 255.0 * torch.rand(2, 3, 256, 256, generator=gen(fx['au']['seed_x']))
    assert rel(o_en.detect_au(sd, xu), fx['au']['intensities']) < 1e-5
    assert rel(o_en.au_heatmaps(sd, (xu - xu.min()) / (xu.max() - xu.min()))[:, :, ::4, ::4], fx['au']['heat']) < 1e-5


def test_predictor_modules_carry_the_reference_keys_and_refuse_the_cpu():
    """Host-side contract of the six predictors (no GPU needed): the parameter trees are the reference's (so the published
    checkpoints load with load_state_dict) and a CPU call fails loudly instead of falling back."""
    from warpedganspace_b200.eval_resnet import fairface_resnet34, hopenet_resnet50, celeba_attr_resnet50
    from warpedganspace_b200.eval_sfd import S3FD
    from warpedganspace_b200.eval_arcface import IDComparator
    from warpedganspace_b200.eval_au import FANAU
    g = gen(5)
    cases = [(fairface_resnet34(), o_en.init_state('basic', {'fc': 18}, g), torch.zeros(1, 3, 224, 224)),
             (celeba_attr_resnet50(), o_en.init_celeba_state(g), torch.zeros(1, 3, 224, 224)),
             (S3FD(), o_en.init_s3fd_state(g), torch.zeros(1, 3, 64, 64)),
             (IDComparator().backbone, o_en.init_arcface_state(g), torch.zeros(1, 3, 112, 112)),
             (FANAU(), o_en.init_au_state(g), torch.zeros(1, 3, 256, 256))]
    for net, sd, x in cases:
        assert set(net.state_dict()) == set(sd), type(net).__name__
        net.load_state_dict(sd, strict=True)
        with pytest.raises(RuntimeError):
            net(x)
    hop = hopenet_resnet50()
    want = set(o_en.init_state('bottleneck', HEADS, g)) | {'fc_finetune.weight', 'fc_finetune.bias'}
    assert set(hop.state_dict()) == want


def _fake_traversal(tmp_path, n_paths=2, n_img=5, size=64):
    from PIL import Image
    exp = tmp_path / 'exp'
    h_dir = exp / 'results' / 'pool' / '4_0.2_0.8' / 'abc123'
    rng = np.random.RandomState(3)
    for d in range(n_paths):
        pdir = h_dir / 'paths_images' / ('path_%03d' % d)
        pdir.mkdir(parents=True)
        for t in range(n_img):
            Image.fromarray(rng.randint(0, 256, (size, size, 3), dtype=np.uint8)).save(str(pdir / ('%06d.jpg' % t)), quality=95)
    torch.save(torch.zeros(n_paths, n_img, 8), str(h_dir / 'paths_latent_codes.pt'))
    (exp / 'args.json').write_text(json.dumps({'gan_type': 'StyleGAN2'}))
    return str(exp), str(h_dir)


def test_driver_files_and_score_arithmetic_with_stub_predictors(tmp_path):
    """Directory walk, per-path batches, the reference's score formulas and output files - with stub predictors that
    return fixed logits, so this runs without a GPU."""
    from warpedganspace_b200 import attribute_space as A
    exp, h_dir = _fake_traversal(tmp_path)
    g = gen(9)
    ff_logits = torch.randn(5, 18, generator=g)
    pose_logits = [torch.randn(5, 66, generator=g) for _ in range(3)]
    seen = {}

    def fairface(x):
        seen['fairface'] = tuple(x.shape)
        return ff_logits

    def hopenet(x):
        seen['hopenet'] = tuple(x.shape)
        return tuple(pose_logits)

    cel_logits = {n: torch.randn(5, 6, generator=g) for n in ('Bangs', 'Eyeglasses', 'No_Beard', 'Smiling', 'Young')}

    def celeba(x):
        seen['celeba'] = (tuple(x.shape), float(x.min()) < -1.0)          # normalised [-1, 1] frames: values below -1 exist
        return cel_logits

    def detector(x):          # one box on even frames, none on odd ones
        return [[[60.0, 70.0, 200.0, 220.0, 0.99]] if t % 2 == 0 else [] for t in range(x.shape[0])]

    done = A.traverse_attribute_space(exp, 'pool', shift_steps=2, eps=0.2, device='cpu',
                                      predictors={'fairface': fairface, 'hopenet': hopenet, 'face_detector': detector,
                                                  'celeba': celeba})
    assert done == [h_dir]
    assert seen == {'fairface': (5, 3, 224, 224), 'hopenet': (5, 3, 224, 224), 'celeba': ((5, 3, 224, 224), True)}
    sm = torch.softmax(cel_logits['Smiling'], dim=1)           # traverse_attribute_space.py:352-356
    want_sm = ((sm.argmax(1) + sm.max(1).values) / 6.0).numpy()
    assert np.allclose(np.load(os.path.join(h_dir, 'eval_np', 'celeba_smiling.npy'))[0], want_sm, atol=1e-6)
    gender, age, race = o_en.fairface_scores(ff_logits)
    yaw, pitch, roll = o_en.hopenet_pose(*pose_logits)
    nd, jd = os.path.join(h_dir, 'eval_np'), os.path.join(h_dir, 'eval_json')
    for name, want in (('gender', gender), ('age', age), ('race', race)):
        got = np.load(os.path.join(nd, name + '.npy'))
        assert got.shape == (2, 5) and np.allclose(got[0], want.numpy(), atol=1e-6) and np.allclose(got[1], want.numpy(), atol=1e-6)
    for name, want in (('yaw', yaw), ('pitch', pitch), ('roll', roll)):
        assert np.allclose(np.load(os.path.join(nd, name + '.npy'))[1], want.numpy() * np.pi / 180, atol=1e-5)
    pose = json.load(open(os.path.join(jd, 'pose.json')))
    assert sorted(pose) == ['0', '1'] and np.allclose(pose['0'][0], yaw.numpy(), atol=1e-4)
    fw = np.load(os.path.join(nd, 'face_width.npy'))
    assert np.allclose(fw[0], [140 / 256.0, 256.0, 140 / 256.0, 256.0, 140 / 256.0])       # (sic: 256.0 when nothing is detected)
    bbox = json.load(open(os.path.join(jd, 'face_bbox.json')))
    assert len(bbox['0']) == 3 and bbox['0'][0][:4] == [60.0, 70.0, 200.0, 220.0]
    assert not os.path.exists(os.path.join(nd, 'identity.npy'))            # predictors that were not given write nothing


def test_crop_face_follows_the_reference():
    from warpedganspace_b200 import attribute_space as A
    imgs = torch.arange(2 * 3 * 256 * 256, dtype=torch.float32).reshape(2, 3, 256, 256)
    c = A.crop_face(imgs, 1, [60, 70, 200, 220], padding=0.25)
    # x: int(0.75*60) - 50 < 0 -> 0 .. int(1.25*200) + 50 > 256 -> 256;  y: int(0.75*70) - 50 = 2 .. min(256, int(1.25*220) + 30)
    assert c.shape == (1, 3, 256, 254) and torch.equal(c[0], imgs[1, :, 0:256, 2:256])


@pytest.mark.gpu
def test_fairface_and_hopenet_kernel_chains_match_the_reference_fixture(golden):
    from warpedganspace_b200.eval_resnet import fairface_resnet34, hopenet_resnet50
    torch.backends.cudnn.allow_tf32 = False
    fx = golden('eval_nets.pt')
    x = torch.randn(2, 3, 224, 224, generator=gen(fx['seed_x'])).cuda()
    sd = o_en.init_state('basic', {'fc': 18}, gen(fx['fairface']['seed']))
    net = fairface_resnet34()
    assert set(net.state_dict()) == set(sd)                    # torchvision's key names: published checkpoints load as they are
    net.load_state_dict(sd, strict=True)
    net.cuda()
    out = net(x)
    assert rel(out, fx['fairface']['out']) < 1e-3              # 36 convs, bf16x3 operands; measured ~1e-5
    assert torch.equal(out.argmax(1).cpu(), fx['fairface']['out'].argmax(1))
    sd = o_en.init_state('bottleneck', HEADS, gen(fx['hopenet']['seed']))
    net = hopenet_resnet50()
    assert set(net.state_dict()) - set(sd) == {'fc_finetune.weight', 'fc_finetune.bias'}
    net.load_state_dict(sd, strict=False)
    net.cuda()
    for got, want in zip(net(x), fx['hopenet']['out']):
        assert rel(got, want) < 1e-3
    with pytest.raises(RuntimeError):
        net(x.cpu())                                           # no CPU fallback
    from warpedganspace_b200.eval_resnet import celeba_attr_resnet50
    sd = o_en.init_celeba_state(gen(fx['celeba']['seed']))
    net = celeba_attr_resnet50()
    assert set(net.state_dict()) == set(sd)                    # the reference's keys (stem.fc.weight, classifier06Bangs.0.bn.running_var ...)
    net.load_state_dict(sd, strict=True)
    net.cuda()
    got = net(x)
    for name, want in fx['celeba']['out'].items():
        assert rel(got[name], want) < 1e-3 and torch.equal(got[name].argmax(1).cpu(), want.argmax(1))


@pytest.mark.gpu
def test_attribute_traversal_end_to_end_on_the_gpu(tmp_path):
    """Frames on disk -> crops -> the two kernel-chain predictors -> the reference's files, against the oracle run on the
    very same crops."""
    from warpedganspace_b200 import attribute_space as A
    from warpedganspace_b200.eval_resnet import fairface_resnet34, hopenet_resnet50
    exp, h_dir = _fake_traversal(tmp_path, n_paths=1, n_img=3, size=96)
    sd_f = o_en.init_state('basic', {'fc': 18}, gen(21))
    sd_h = o_en.init_state('bottleneck', HEADS, gen(22))
    ff, hp = fairface_resnet34(), hopenet_resnet50()
    ff.load_state_dict(sd_f)
    hp.load_state_dict(sd_h, strict=False)
    ff.cuda(); hp.cuda()
    crops = {}

    def tap(name, net):
        def run(x):
            crops[name] = x.detach().cpu()
            return net(x)
        return run

    A.traverse_attribute_space(exp, 'pool', shift_steps=2, eps=0.2, predictors={'fairface': tap('f', ff), 'hopenet': tap('h', hp)})
    gender, age, race = o_en.fairface_scores(o_en.resnet_forward(sd_f, crops['f'], 'basic', ('fc',))[0])
    nd = os.path.join(h_dir, 'eval_np')
    assert np.allclose(np.load(os.path.join(nd, 'gender.npy'))[0], gender.numpy(), atol=2e-3)
    assert np.allclose(np.load(os.path.join(nd, 'race.npy'))[0], race.numpy(), atol=2e-3)
    yaw, _, _ = o_en.hopenet_pose(*o_en.resnet_forward(sd_h, crops['h'], 'bottleneck', tuple(HEADS)))
    assert np.allclose(np.load(os.path.join(nd, 'yaw.npy'))[0], yaw.numpy() * np.pi / 180, atol=2e-3)


@pytest.mark.gpu
def test_s3fd_detector_kernel_chain_matches_the_reference_fixture(golden):
    """The face detector on the tensor-core convs: the twelve head outputs against the reference network, the detections
    (soft-max, anchor decode, NMS 0.3, score > 0.5) against the reference's batch_detect + nms."""
    from warpedganspace_b200.eval_sfd import S3FD, SFDDetector
    torch.backends.cudnn.allow_tf32 = False
    fx = golden('eval_nets.pt')['s3fd']
    sd = o_en.init_s3fd_state(gen(fx['seed']))
    det = SFDDetector()
    assert set(det.face_detector.state_dict()) == set(sd)          # the reference's key names: s3fd-619a316812.pth loads as it is
    det.face_detector.load_state_dict(sd, strict=True)
    det.face_detector.cuda()
    x = (255.0 * torch.rand(2, 3, 128, 128, generator=gen(fx['seed_x']))).cuda()
    outs = det.face_detector(x)
    for got, want in zip(outs, fx['out']):
        assert rel(got[:, :, ::2, ::2], want) < 1e-3
    found, error, _ = det.detect_from_batch(x)
    assert not error
    for mine, want in zip(found, fx['detections']):
        mine, want = np.array(mine, dtype=np.float64).reshape(-1, 5), want.numpy()
        assert abs(len(mine) - len(want)) <= 2                       # (a score within 1e-5 of 0.5 / an IoU within 1e-5 of 0.3 may flip)
        # random weights put exp(0.2 * loc) anywhere from 1e-3 to 1e9 pixels: compare relative to the box's own scale
        dist = (np.abs(mine[:, None, :] - want[None, :, :]) / (1.0 + np.abs(want[None, :, :]))).max(axis=2)
        assert (dist.min(axis=1) < 2e-2).mean() > 0.9 and (dist.min(axis=0) < 2e-2).mean() > 0.9
    with pytest.raises(RuntimeError):
        S3FD()(x.cpu())                                              # no CPU fallback


@pytest.mark.gpu
def test_arcface_identity_comparator_kernel_chain_matches_the_reference_fixture(golden):
    from warpedganspace_b200.eval_arcface import IDComparator
    fx = golden('eval_nets.pt')['arcface']
    sd = o_en.init_arcface_state(gen(fx['seed']))
    idc = IDComparator()
    assert set(idc.backbone.state_dict()) == set(sd)               # model_ir_se50.pth loads as it is
    idc.backbone.load_state_dict(sd, strict=True)
    idc.cuda()
    xa = torch.rand(3, 3, 256, 256, generator=gen(fx['seed_x'])) * 2 - 1
    xa[2] = 0.7 * xa[0] + 0.3 * xa[1]
    xa = xa.cuda()
    with torch.no_grad():
        assert rel(idc.extract_feats(xa), fx['feats']) < 1e-3
        sims = torch.stack([idc(xa[0:1], xa[t: t + 1]) for t in range(3)])
    assert (sims.cpu() - fx['sim']).abs().max() < 1e-4
    with pytest.raises(RuntimeError):
        idc.backbone(xa[:, :, :112, :112].cpu())


@pytest.mark.gpu
def test_au_detector_kernel_chain_matches_the_reference_fixture(golden):
    from warpedganspace_b200.eval_au import AUdetector
    fx = golden('eval_nets.pt')['au']
    sd = o_en.init_au_state(gen(fx['seed']))
    det = AUdetector()
    assert set(det.FAN.state_dict()) == set(sd)                    # disfa_adaptation_f0.pth['state_dict'] loads as it is
    det.FAN.load_state_dict(sd, strict=True)
    det.FAN.cuda()
    xu = (255.0 * torch.rand(2, 3, 256, 256, generator=gen(fx['seed_x']))).cuda()
    with torch.no_grad():
        heat = det.FAN((xu - xu.min()) / (xu.max() - xu.min()))
    assert rel(heat[:, :, ::4, ::4], fx['heat']) < 1e-3
    assert rel(det.detect_AU(xu), fx['intensities']) < 1e-3
    with pytest.raises(RuntimeError):
        det.FAN(xu.cpu())


@pytest.mark.gpu
def test_attribute_traversal_with_every_predictor_on_the_gpu(tmp_path):
    """All six predictors as kernel chains (seeded weights): every file of the reference's eval_json / eval_np layout is written,
    the identity score is the batched cosine similarity to the centre frame, the AU table has one row per action unit."""
    from warpedganspace_b200 import attribute_space as A
    from warpedganspace_b200.eval_resnet import fairface_resnet34, hopenet_resnet50, celeba_attr_resnet50
    from warpedganspace_b200.eval_sfd import SFDDetector
    from warpedganspace_b200.eval_arcface import IDComparator
    from warpedganspace_b200.eval_au import AUdetector
    exp, h_dir = _fake_traversal(tmp_path, n_paths=2, n_img=3, size=96)
    det, idc, au = SFDDetector(), IDComparator().cuda(), AUdetector()
    det.face_detector.load_state_dict(o_en.init_s3fd_state(gen(31)))
    det.face_detector.cuda()
    idc.backbone.load_state_dict(o_en.init_arcface_state(gen(32)))
    au.FAN.load_state_dict(o_en.init_au_state(gen(33)))
    au.FAN.cuda()
    def detector(x):        # seeded weights give boxes far outside the frame: run the detector, hand the driver a plausible box
        found = det(x)
        assert len(found) == x.shape[0]
        return [[np.array([60.0, 70.0, 190.0, 200.0, 0.9], dtype=np.float32)] for _ in found]

    preds = {'face_detector': detector, 'id_comparator': idc, 'au_detector': au, 'fairface': fairface_resnet34().cuda(),
             'hopenet': hopenet_resnet50().cuda(), 'celeba': celeba_attr_resnet50().cuda()}
    A.traverse_attribute_space(exp, 'pool', shift_steps=2, eps=0.2, predictors=preds, gan_type='StyleGAN2')
    nd, jd = os.path.join(h_dir, 'eval_np'), os.path.join(h_dir, 'eval_json')
    want = ['face_width', 'face_height', 'identity', 'age', 'race', 'gender', 'yaw', 'pitch', 'roll', 'celeba_bangs',
            'celeba_eyeglasses', 'celeba_beard', 'celeba_smiling', 'celeba_age'] + ['%s_%s' % kv for kv in A.AUs.items()]
    for name in want:
        t = np.load(os.path.join(nd, name + '.npy'))
        assert t.shape == (2, 3) and np.isfinite(t).all(), name
    ident = np.load(os.path.join(nd, 'identity.npy'))
    assert np.allclose(ident[:, 1], 1.0, atol=1e-5) and (ident <= 1.0 + 1e-5).all()
    import json
    with open(os.path.join(jd, 'au.json')) as f:
        au_rows = json.load(f)
    assert len(au_rows['0']) == 12 and len(au_rows['0'][0]) == 3
    frames = A.load_path_images(os.path.join(h_dir, 'paths_images', 'path_000'), 'cuda')
    small = A.resize_center_crop(frames, 256)
    with torch.no_grad():                                          # the per-pair call of the reference gives the same scores
        pair = [float(idc(small[1:2] / 255.0 * 2.0 - 1.0, small[t: t + 1] / 255.0 * 2.0 - 1.0)) for t in range(3)]
    assert np.allclose(ident[0], pair, atol=1e-5)

