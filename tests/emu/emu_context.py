"""tests/emu/emu_context.py -- TEST INFRASTRUCTURE: a gat_b200.device.Context on the SIMT-emulated build of the
kernels (build_emu.py), for tests/test_emu_parity.py and the --emu switch of the stress tools under tools/."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

_lib_handle = None


def library():
    global _lib_handle
    if _lib_handle is None:
        import build_emu
        from gat_b200 import _lib
        _lib_handle = _lib.bind(ctypes.CDLL(build_emu.build()))
    return _lib_handle


def context():
    from gat_b200 import device

    class EmuContext(device.Context):
        """device.Context on the emulated build (the wrappers only ever use ctx.lib)"""

        def __init__(self, lib):
            self.lib = lib
            h = ctypes.c_void_p()
            rc = lib.gatb_create(0, ctypes.byref(h))
            assert rc == 0, lib.gatb_last_error(None)
            self.handle = h
            self.device = 0

    return EmuContext(library())
