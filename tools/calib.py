"""Print device info and calibrated pipe rates (FFMA, FFMA2, MUFU.EX2, DFMA)."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kdsource_b200 import _lib  # noqa: E402


def calibrate(iters=4096):
    L = _lib.load()
    t = _lib.torch()
    sm, ma, mi, clk = (ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int())
    _lib.check(L.kdsb200_device_info(ctypes.byref(sm), ctypes.byref(ma), ctypes.byref(mi),
                                     ctypes.byref(clk)))
    scratch = _lib.empty(sm.value * 8 * 256, t.float32)
    out = {"sm": sm.value, "cc": "{}.{}".format(ma.value, mi.value), "clock_khz": clk.value}
    for mode, name in ((0, "ffma"), (1, "ffma2"), (2, "mufu_ex2"), (3, "ffma2_plus_ffma"),
                       (4, "dfma"), (5, "fadd2"), (6, "fmul2"), (7, "fadd2_ffma2_mix")):
        r = ctypes.c_double()
        _lib.check(L.kdsb200_calibrate_pipe(mode, _lib.ptr(scratch), iters, ctypes.byref(r),
                                            _lib.stream()))
        out[name + "_lane_ops_per_s"] = r.value
    return out


if __name__ == "__main__":
    print(json.dumps(calibrate()))
