"""One vu_patch_max_ws call on a BASELINE shape, for ncu (developer tool).   python bench/prof_k2.py [cfg2|cfg5] [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffuncertainty_b200 import _lib  # noqa: E402

SHAPES = {"cfg2": ((64, 64, 64), 128, (10, 10, 10)), "cfg5": ((1, 512, 1024), 16, (1, 10, 10)), "cfg3": ((1, 1024, 2048), 8, (1, 10, 10))}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    dims, B, box = SHAPES[name]
    lib = _lib.load()
    V = dims[0] * dims[1] * dims[2]
    maps = torch.rand((B, V), device="cuda") ** 3 * 0.69
    out_max = torch.empty(B, dtype=torch.float64, device="cuda")
    out_first = torch.empty(B, dtype=torch.int64, device="cuda")
    ws_bytes = int(lib.vu_patch_workspace_bytes(B, *dims, *box))
    ws = torch.empty(max(ws_bytes // 8, 1), dtype=torch.int64, device="cuda")
    for _ in range(iters):
        _lib.check(lib.vu_patch_max_ws(maps.data_ptr(), B, *dims, *box, 0, out_max.data_ptr(), out_first.data_ptr(), ws.data_ptr(),
                                       ws_bytes, _lib.current_stream_ptr()), "patch")
    torch.cuda.synchronize()
    print(name, float(out_max[0]), int(out_first[0]))


if __name__ == "__main__":
    main()
