"""Output stage of sampling / traversal: device-side ``tensor2image`` and JPEG files
(traverse_latent_space.py:26-41,466-483; sample_gan.py:10-25,172-176).

The reference pulls every fp32 image to the host, normalises it there and lets ToPILImage transpose CHW -> HWC.  Here
the per-image min/max and the uint8 conversion run in two launches over the NHWC image the generator produced
(csrc/imgio.cu, bit-identical pixels) and the frames are JPEG-encoded ON THE GPU by nvJPEG with the reference's encoder
settings (quality 95, optimised Huffman tables, progressive; csrc/jpeg.cu), so only the compressed bitstream crosses PCIe
and the host just writes files.  When a resize is requested (``img_size``, done by PIL in the reference) the PIL path of
the reference is used: one byte per value crosses PCIe and a small thread pool encodes.
"""
import ctypes
from concurrent.futures import ThreadPoolExecutor

import torch

from . import _lib


def images_to_uint8(img, adaptive=True):
    """img: logical NCHW fp32 CUDA tensor as returned by the generator wrappers (channels-last memory) or an NHWC
    tensor -> uint8 [N, H, W, C] on the device."""
    if not img.is_cuda:
        raise RuntimeError('images_to_uint8 runs on CUDA tensors only (no CPU fallback); got %s' % img.device)
    if img.dim() != 4:
        raise ValueError('expected a batch of images')
    if img.shape[1] <= 4 and img.shape[-1] > 4:                       # logical NCHW -> NHWC view (free for channels-last)
        img = img.permute(0, 2, 3, 1)
    x = img.float().contiguous()
    n = x.shape[0]
    out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    if n == 0:
        return out
    count = x[0].numel()
    pairs = n * 1024
    ws = torch.empty(2 * pairs, dtype=torch.float32, device=x.device)
    _lib.call('wgs_image_to_u8', _lib.ptr(x), n, count, 1 if adaptive else 0, _lib.ptr(ws), pairs, _lib.ptr(out),
              _lib.stream())
    return out


def save_jpeg(pixels_hwc, path, quality=95, img_size=None):
    """pixels_hwc: uint8 [H, W, C] (CPU tensor or array).  Same PIL call as the reference (progressive, optimised)."""
    from PIL import Image
    arr = pixels_hwc.cpu().numpy() if torch.is_tensor(pixels_hwc) else pixels_hwc
    im = Image.fromarray(arr[:, :, 0] if arr.shape[2] == 1 else arr)
    if img_size:
        im = im.resize((img_size, img_size))
    im.save(path, 'JPEG', quality=quality, optimize=True, progressive=True)


def nvjpeg_available():
    return bool(_lib.load().wgs_jpeg_available())


_HOST_BUF = {}


def encode_jpegs(pixels_nhwc, quality=95, progressive=True):
    """uint8 [N, H, W, C] CUDA tensor (C = 3 or 1) -> list of N JPEG bitstreams (bytes), encoded by nvJPEG on the device."""
    if not pixels_nhwc.is_cuda or pixels_nhwc.dtype != torch.uint8:
        raise RuntimeError('encode_jpegs needs a uint8 CUDA tensor (no CPU fallback)')
    x = pixels_nhwc.contiguous()
    n, h, w, c = x.shape
    cap = h * w * c + 65536
    key = (n, cap)
    if key not in _HOST_BUF:
        _HOST_BUF.clear()
        _HOST_BUF[key] = (torch.empty(n * cap, dtype=torch.uint8).pin_memory(), (ctypes.c_longlong * n)())
    buf, sizes = _HOST_BUF[key]
    _lib.call('wgs_jpeg_encode', _lib.ptr(x), n, h, w, c, int(quality), int(bool(progressive)),
              ctypes.c_void_p(buf.data_ptr()), cap, sizes, _lib.stream())
    arr = buf.numpy()
    return [arr[i * cap: i * cap + int(sizes[i])].tobytes() for i in range(n)]


def save_jpegs(pixels_nhwc, paths, quality=95, img_size=None, workers=8, backend=None):
    """Writes one JPEG per image.  backend 'nvjpeg' (default when available and no resize is requested): GPU encode, the host
    only writes the bitstreams; 'pil': one D2H copy of the uint8 batch, then parallel PIL encodes (the reference's encoder)."""
    if backend is None:
        backend = 'nvjpeg' if (img_size is None and pixels_nhwc.is_cuda and nvjpeg_available()) else 'pil'
    if backend == 'nvjpeg':
        if img_size is not None:
            raise ValueError('the nvJPEG path does not resize; use backend="pil" with img_size')
        for data, path in zip(encode_jpegs(pixels_nhwc, quality, True), paths):
            with open(path, 'wb') as f:
                f.write(data)
        return
    host = pixels_nhwc.cpu()
    with ThreadPoolExecutor(max_workers=workers) as pool:
        list(pool.map(lambda a: save_jpeg(a[0], a[1], quality, img_size), zip(host, paths)))
