"""ctypes binding of libeat_b200.so.  The prototypes are parsed from include/eat_b200.h, so the
header is the single source of truth for the C ABI.  There is no CPU fallback: if the library is
missing, importing a product module that needs it raises."""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libeat_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "eat_b200.h")

_CTYPES = {"int": ctypes.c_int, "long long": ctypes.c_longlong, "float": ctypes.c_float,
           "double": ctypes.c_double, "cudaStream_t": ctypes.c_void_p}


def parse_header(path=HEADER_PATH):
    """-> {name: (restype, [argtypes])} for every function declared in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(const char\*|int)\s+(eat_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    ty = a.rsplit(" ", 1)[0]
                    argtypes.append(_CTYPES[ty])
        protos[name] = (ctypes.c_char_p if ret.startswith("const char") else ctypes.c_int, argtypes)
    return protos


class EatError(RuntimeError):
    pass


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -m efficientat_b200.build` "
                "(the product path has no CPU / PyTorch fallback)")
        self._dll = ctypes.CDLL(LIB_PATH)
        self.launches = 0          # number of C-ABI kernel launchers called (bench.py: gpu_launches)
        self.protos = parse_header()
        for name, (ret, args) in self.protos.items():
            fn = getattr(self._dll, name)
            fn.restype = ret
            fn.argtypes = args
            if ret is ctypes.c_int and name not in ("eat_abi_version",):
                setattr(self, name[4:], self._checked(fn, name))
            else:
                setattr(self, name[4:], fn)

    def _checked(self, fn, name):
        last_error = self._dll.eat_last_error
        last_error.restype = ctypes.c_char_p

        def call(*args):
            self.launches += 1
            rc = fn(*args)
            if rc != 0:
                raise EatError(f"{name} failed (code {rc}): {last_error().decode()}")
        call.__name__ = name
        return call


def check_module_tensors(module, device, what):
    """The kernels take raw `data_ptr()`s as `float*`: every parameter and buffer handed to them must be an fp32 (int64
    for BatchNorm's step counter), contiguous tensor on the input's CUDA device.  A forgotten `.to(device)`,
    `model.half()` or `.double()` raises here instead of dereferencing a host pointer / wrong-width data."""
    import itertools
    import torch
    for name, t in itertools.chain(module.named_parameters(), module.named_buffers()):
        ok_dtype = t.dtype == torch.float32 or (not t.is_floating_point())
        if t.device != device or not ok_dtype or not t.is_contiguous():
            raise RuntimeError(
                f"{what}: tensor '{name}' is {t.dtype} on {t.device}{'' if t.is_contiguous() else ' (non-contiguous)'}; "
                f"the sm_100a kernels need contiguous fp32 tensors on {device} (the input's device). "
                "Call module.to(device) / keep parameters in fp32 (bf16 activation storage is selected with "
                "get_model(precision='bf16'), not with .half()/.bfloat16()).")


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib
