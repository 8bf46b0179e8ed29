"""ctypes binding of libptta_b200.so (the C ABI declared in include/ptta_b200.h).

There is no fallback: if the shared library is missing or a call fails, a RuntimeError is raised."""
import ctypes
import os
import re

from .build import LIB_PATH

c_void_p, c_int, c_float, c_ll, c_size_t, c_char_p = (ctypes.c_void_p, ctypes.c_int, ctypes.c_float,
                                                      ctypes.c_longlong, ctypes.c_size_t, ctypes.c_char_p)
HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'include', 'ptta_b200.h')

_CTYPE = {
    'int': c_int, 'float': c_float, 'double': ctypes.c_double, 'long long': c_ll, 'size_t': c_size_t, 'ptta_stream_t': c_void_p,
    'const char*': c_char_p, 'void': None,
}


def _parse_header():
    """Prototype table {name: (restype, [argtypes])} read from the header, so the Python side cannot drift
    from the declared ABI (and tests can check that every declared symbol is exported)."""
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    protos = {}
    for m in re.finditer(r'^([\w\s\*]+?)\b(ptta_\w+)\s*\(([^;{]*)\)\s*;', text, flags=re.M):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if ret.startswith('typedef'):
            continue
        def conv(t):
            t = re.sub(r'\s+', ' ', t.strip())
            if t in _CTYPE:
                return _CTYPE[t]
            if '*' in t:
                return c_void_p
            raise ValueError('unhandled C type %r in %s' % (t, name))
        argtypes = []
        if args and args != 'void':
            for a in args.split(','):
                a = a.strip()
                a = re.sub(r'\b\w+$', '', a).strip() if not a.endswith('*') else a   # drop the parameter name
                argtypes.append(conv(a))
        restype = conv(ret)
        protos[name] = (restype, argtypes)
    return protos


PROTOTYPES = _parse_header()
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError('%s is missing: run `python -c "import __graft_entry__ as g; g.build()"` '
                               '(there is no CPU or PyTorch fallback for the TTA step)' % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in PROTOTYPES.items():
            fn = getattr(L, name)      # AttributeError here == header / library mismatch
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = L
    return _lib


def last_error():
    return lib().ptta_last_error().decode()


def check(rc, what=''):
    if rc != 0:
        raise RuntimeError('libptta_b200 %s failed (%d): %s' % (what, rc, last_error()))


def ptr(t):
    """device pointer of a torch tensor (or None)"""
    if t is None:
        return None
    return c_void_p(t.data_ptr())
