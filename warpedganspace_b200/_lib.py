"""ctypes binding of libwgs_b200.so (the C ABI declared in include/wgs_b200.h).

There is no CPU fallback: if the library is missing or the tensors are not CUDA tensors the call
raises.  PyTorch is used only for device memory and streams.
"""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libwgs_b200.so')
HEADER = os.path.join(os.path.dirname(_HERE), 'include', 'wgs_b200.h')

_lib = None

_CTYPES = {
    'int': ctypes.c_int, 'float': ctypes.c_float, 'unsigned long long': ctypes.c_ulonglong,
    'long long': ctypes.c_longlong, 'void': None,
}


class LibraryMissing(RuntimeError):
    pass


def header_prototypes(path=HEADER):
    """Parse `ret name(args);` prototypes out of the public header -> {name: (ret, [arg types])}."""
    text = open(path).read()
    text = re.sub(r'/\*.*?\*/', ' ', text, flags=re.S)
    text = re.sub(r'^\s*#.*$', ' ', text, flags=re.M)            # preprocessor lines
    text = text.replace('extern "C" {', ' ')
    protos = {}
    for m in re.finditer(r'([A-Za-z_][\w \t\*]*?)\b(wgs_\w+)\s*\(([^)]*)\)\s*;', text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        arg_types = []
        if args and args != 'void':
            for a in args.split(','):
                a = a.strip()
                if '*' in a:
                    arg_types.append('ptr')
                else:
                    arg_types.append(' '.join(a.split()[:-1]).replace('const ', '').strip())
        protos[name] = (ret, arg_types)
    return protos


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            'libwgs_b200.so not found at %s — build it with `python -m warpedganspace_b200.csrc.build` '
            '(or __graft_entry__.build()); there is no CPU / PyTorch fallback.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (ret, args) in header_prototypes().items():
        fn = getattr(lib, name)            # AttributeError here = header/library mismatch
        fn.restype = ctypes.c_char_p if ret == 'const char*' else _CTYPES[ret.replace('const ', '')]
        fn.argtypes = [ctypes.c_void_p if a == 'ptr' else _CTYPES[a] for a in args]
    _lib = lib
    return lib


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('warpedganspace_b200 kernels need CUDA tensors (got %s); no CPU fallback exists'
                           % t.device)
    if not t.is_contiguous():
        raise RuntimeError('tensor must be contiguous')
    return ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def check(rc):
    if rc != 0:
        raise RuntimeError('libwgs_b200: ' + load().wgs_last_error().decode())


def call(name, *args):
    check(getattr(load(), name)(*args))


def launch_count():
    return int(load().wgs_launch_count())


def reset_launch_count():
    load().wgs_reset_launch_count()
