"""Random-access ceilings of the multi-deal table pattern: python profiles/rnd_probe.py [log2_lines ...]"""
import ctypes as C, sys
sys.path.insert(0, '.')
import torch
from scopa_b200 import _lib
torch.cuda.set_device(0)
lib = _lib.load()
for lg in [int(a) for a in sys.argv[1:]] or [19, 21, 23, 25, 27]:
    out = (C.c_double * 3)()
    _lib.check(lib.ms_debug_random_access_peaks(lg, out, None))
    print(f"2^{lg} lines ({(128 << lg) / 2**30:.3f} GiB): dependent {out[0]/1e9:.2f} G reads/s, independent x8 {out[1]/1e9:.2f} G reads/s "
          f"({out[1]*64/1e9:.0f} GB/s of 64-byte reads), RED x4 {out[2]/1e9:.2f} G lines/s", flush=True)
