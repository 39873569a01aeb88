"""Predicted traversal work per path ray from the CPU oracle's visit counters (no GPU): one pass of a 256x144 frame per workload, for the builder's raw
trees (CTL_SBVH_ROTATE=0), the post-optimised trees (default) and, for the instanced scenes, the re-braided scene level.  Algorithmic bytes per ray =
48 + 64 inner + 52 triangle tests + 108 instance entries.  Usage: python scripts/oracle_visit_counts.py > profiles/<name>.log"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys; sys.path.insert(0, sys.argv[3]); sys.path.insert(0, sys.argv[3] + "/tests")
import cudatracerlib_b200 as ctl
from cudatracerlib_b200 import api
import oracle_binding as ob
kind, depth = sys.argv[1], int(sys.argv[2])
s = ctl.Scene(kind, 256, 144)
img, rays, cnt = ob.render(s.view, 256, 144, 1, max_path_length=depth, counts=True)
print("nodes %d  rays %d  inner/ray %.2f  tris/ray %.2f  inst/ray %.2f  ->  %.0f B/ray" % (s.view.n_nodes, rays, cnt[0] / rays, cnt[1] / rays, cnt[2] / rays, api.traversal_bytes(cnt, rays) / rays))
'''
for kind, depth in (("c2", 8), ("c3", 8), ("c4", 8), ("c5", 32)):
    variants = [("raw trees", {"CTL_SBVH_ROTATE": "0"}), ("post-optimised (default)", {})]
    if kind in ("c4", "c5"):
        variants += [("post-optimised + re-braid 256", {"CTL_REBRAID": "256"}), ("post-optimised + re-braid 1024", {"CTL_REBRAID": "1024"}), ("post-optimised + re-braid 4096", {"CTL_REBRAID": "4096"})]
    for name, env in variants:
        e = dict(os.environ); e.update(env)
        r = subprocess.run([sys.executable, "-c", CODE, kind, str(depth), ROOT], env=e, capture_output=True, text=True)
        print("%-3s depth %-2d %-32s %s" % (kind, depth, name, (r.stdout.strip().splitlines() or [r.stderr[-300:]])[-1]), flush=True)
