"""Integer-pipe peak of the K1 instruction mix (IMAD.WIDE carry chains), GIMAD/s."""
import ctypes as C, sys
sys.path.insert(0, '.')
from relp_b200 import _lib
lib = _lib.load()
v = C.c_double()
for secs in (0.2, 1.0, 3.0):
    rc = lib.rg_measure_imad_peak(0, secs, C.byref(v))
    print(f"rc={rc} measured over ~{secs}s: {v.value/1e9:.1f} GIMAD.WIDE/s", flush=True)
