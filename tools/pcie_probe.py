"""PCIe rates of this box through librest_b200's probe (GB/s)."""
import ctypes as C, json, sys
sys.path.insert(0, ".")
from rest_tensors_b200 import lib
from rest_tensors_b200._lib import check
from rest_tensors_b200.device import Context
ctx = Context(0)
out = {}
for name, mode, width in [("h2d", 0, 0), ("d2h", 1, 0), ("d2h_2d_1k", 2, 1024), ("d2h_2d_2k", 2, 2048), ("d2h_2d_4k", 2, 4096),
                          ("d2h_2d_8k", 2, 8192), ("d2h_2d_16k", 2, 16384), ("duplex_sum", 3, 0), ("d2h_zero_copy", 4, 0),
                          ("d2h_zero_copy_rows_2k", 5, 2048), ("d2h_zero_copy_rows_1k", 5, 1024)]:
    g = C.c_double()
    check(lib.rb_pcie_probe(ctx.h, mode, 1 << 30, width, 3, C.byref(g)), name)
    out[name] = round(g.value, 2)
    print(name, out[name], flush=True)
json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/pcie_probe.json", "w"), indent=1)
