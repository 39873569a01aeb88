"""Render demo images on the GPU through the public API and write them as PNG (no image library needed).
Usage: python scripts/render_image.py <out_dir>"""
import os, sys, zlib, struct
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cudatracerlib_b200 import Scene, PathTracer


def write_png(path, rgba):
    h, w, _ = rgba.shape
    raw = b"".join(b"\x00" + rgba[y, :, :3].tobytes() for y in range(h))
    def chunk(tag, data): return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)
    open(path, "wb").write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw, 9)) + chunk(b"IEND", b""))


out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
os.makedirs(out, exist_ok=True)
for kind, w, h, spp, depth in (("cornell", 256, 256, 256, 8), ("c2", 640, 360, 128, 8), ("c3", 640, 360, 128, 8), ("c4", 640, 360, 64, 8)):
    s = Scene(kind, w, h); t = PathTracer(w, h); t.InitializeScene(s); t.setParameter("MaxPathLength", depth)
    done = 0
    while done < spp:
        n = min(32, spp - done); t.DoPasses(n, new_trace=(done == 0)); done += n
    t.synchronize()
    img = t.resolveSRGB8()
    write_png(os.path.join(out, f"render_{kind}_{w}x{h}_{spp}spp.png"), img)
    print(kind, "mean sRGB", img[:, :, :3].mean(), "passes", t.getNumPassesDone(), "total rays", t.getTotalRays())
    t.close()
