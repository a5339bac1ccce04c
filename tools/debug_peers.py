#!/usr/bin/env python
"""Debug helper: the peer-memory resample with G handles on ONE device (in-process attach), step by step."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smarc_navigation_b200 import ParticleFilter, _lib  # noqa: E402


def main():
    G, M = int(sys.argv[1]) if len(sys.argv) > 1 else 2, int(sys.argv[2]) if len(sys.argv) > 2 else 5000
    L = _lib.load()
    N = G * M
    rng = np.random.default_rng(0)
    logw = (rng.normal(0, 4, N) - 150).astype(np.float32)
    poses = np.zeros((N, 6))
    poses[:, 0] = np.arange(N)
    pfs = [ParticleFilter(particle_count=M, global_offset=g * M, global_count=N, seed=1) for g in range(G)]
    handles = (C.c_void_p * G)(*[pf._h for pf in pfs])
    for g, pf in enumerate(pfs):
        _lib.check(L.auvpf_peer_attach(pf._h, handles, G, g))
        pf.set_poses(poses[g * M:(g + 1) * M])
        pf.set_log_weights(logw[g * M:(g + 1) * M])
    print("attached", flush=True)
    if os.environ.get("LONE"):
        for g in range(G):
            t = time.time()
            _lib.check(L.auvpf_shard_resample(pfs[g]._h, 0.37, None, None, None))
            print("enqueued rank", g, "%.3f s" % (time.time() - t), flush=True)
    else:
        _lib.check(L.auvpf_group_resample(handles, G, 0.37, None, None))
        print("enqueued group", flush=True)
    for g, pf in enumerate(pfs):
        t = time.time()
        pf.sync()
        print("synced rank", g, "%.3f s" % (time.time() - t), flush=True)
    got = np.concatenate([pf.poses() for pf in pfs])[:, 0].astype(np.int64)
    from oracle import c_oracle, pf_oracle as o
    want = c_oracle.systematic(o.weights_from_q(c_oracle.quantise_logw(logw)[0]), 0.37)
    print("ancestors equal:", np.array_equal(got, want), flush=True)


if __name__ == "__main__":
    main()
