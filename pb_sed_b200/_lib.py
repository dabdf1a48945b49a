"""ctypes binding of ``libpbsed_b200.so`` (the C ABI declared in ``include/pbsed_b200.h``).

The prototypes are parsed from the header itself, so the Python side cannot drift
from the declared ABI.  There is NO fallback: if the shared library is missing
``load()`` raises, and every op in this package goes through ``call()``.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), 'include', 'pbsed_b200.h')
LIB_PATH = os.path.join(_HERE, 'libpbsed_b200.so')
MAX_TAPS = 16


class TapGemmDesc(ctypes.Structure):
    """mirror of ``pbsed_tapgemm_desc``."""
    _fields_ = [
        ('B', ctypes.c_int), ('F_in', ctypes.c_int), ('F_out', ctypes.c_int), ('T', ctypes.c_int),
        ('Cin', ctypes.c_int), ('Cout', ctypes.c_int), ('ntaps', ctypes.c_int),
        ('df', ctypes.c_int * MAX_TAPS), ('dt', ctypes.c_int * MAX_TAPS),
        ('relu', ctypes.c_int), ('per_f', ctypes.c_int),
        ('w_tap_stride', ctypes.c_longlong), ('w_sn', ctypes.c_longlong), ('w_sc', ctypes.c_longlong),
        ('in_stride', ctypes.c_int), ('out_stride', ctypes.c_int), ('precision', ctypes.c_int),
        ('no_input_mask', ctypes.c_int), ('in_dtype', ctypes.c_int), ('out_dtype', ctypes.c_int),
    ]


_SCALARS = {'int': ctypes.c_int, 'long long': ctypes.c_longlong, 'float': ctypes.c_float,
            'double': ctypes.c_double}


def parse_header(path=HEADER):
    """-> {name: (restype, [argtypes], [argnames])} for every ``pbsed_*`` prototype."""
    src = open(path).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    protos = {}
    for m in re.finditer(r'\b(int|long long)\s+(pbsed_\w+)\s*\(([^)]*)\)\s*;', src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        argtypes, argnames = [], []
        if args and args != 'void':
            for a in args.split(','):
                a = ' '.join(a.split())
                if '*' in a:
                    base, an = a.rsplit('*', 1)
                    argtypes.append(ctypes.POINTER(TapGemmDesc) if 'pbsed_tapgemm_desc' in base
                                    else ctypes.c_void_p)
                    argnames.append(an.strip())
                else:
                    base, an = a.rsplit(' ', 1)
                    argtypes.append(_SCALARS[base.replace('const ', '').strip()])
                    argnames.append(an)
        protos[name] = (_SCALARS[ret], argtypes, argnames)
    return protos


_lib = None
_protos = None


def load():
    global _lib, _protos
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; '
            f'g.build()"` (nvcc, sm_100a).  pb_sed_b200 has no CPU / eager fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    _protos = parse_header()
    for name, (ret, argtypes, _) in _protos.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export it
        fn.restype = ret
        fn.argtypes = argtypes
    lib.pbsed_last_kernel.restype = ctypes.c_char_p      # the one non-integer entry point
    lib.pbsed_last_kernel.argtypes = []
    _lib = lib
    return lib


class PbsedError(RuntimeError):
    pass


# optional per-call device timing (bench.py's roofline pass): set to a list to record
# (name, start_event, end_event, args) for every C-ABI call made while it is set.
profile_sink = None


def call(name, *args):
    """call an int-returning entry point; raise on a non-zero status."""
    fn = getattr(load(), name)
    if profile_sink is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        profile_sink.append((name, e0, e1, args, load().pbsed_last_kernel().decode()))
    else:
        rc = fn(*args)
    if rc != 0:
        kind = {-1: 'PBSED_EINVAL (bad argument / unsupported shape)',
                -2: 'PBSED_EWORKSPACE'}.get(rc, f'cudaError {rc}' if rc > 0 else str(rc))
        raise PbsedError(f'{name} failed: {kind}')


def launch_count():
    return int(load().pbsed_launch_count())
