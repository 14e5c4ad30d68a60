"""Generates the committed golden fixtures under tests/golden/ .  Run in the BUILD container (needs /root/reference
for oracle/_ref):   python tools/make_golden.py

extract_golden.npz : small synthetic scans and the outputs of the REAL reference extraction
                     (/root/reference/src/laserProcessingClass.cpp compiled in place, oracle/_ref) -- pins the oracle
                     restatement and, through it, the CUDA kernels.
odom_oracle_golden.npz : trajectory / map sizes of the oracle restatement of the odometry on a short sequence.  The
                     reference ships no fixture for this part and cannot be built here, so this file only guards the
                     oracle against regressions ("parity unpinned", see oracle/oracle_odom.cpp).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from pf_loader import pfb  # noqa: E402
import oracle as O  # noqa: E402


def main():
    assert O.have_ref(), "oracle/_ref/libpf_ref_extract.so missing: run `make -C oracle ref` where /root/reference exists"
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    cases = {}
    specs = [("hdl64_az400", dict(sensor_lines=64, azimuth_steps=400, seed=11), 3, 64),
             ("hdl64_az400_b", dict(sensor_lines=64, azimuth_steps=400, seed=12), 40, 64),
             ("vlp32_az600", dict(sensor_lines=32, azimuth_steps=600, seed=13), 5, 32),
             ("vlp16_az500", dict(sensor_lines=16, azimuth_steps=500, seed=14), 2, 16)]
    for name, kw, frame, lines in specs:
        p = pfb.synth.params(**kw)
        s = pfb.synth.scan(p, frame)
        e, u = O.ref_extract(s, num_lines=lines)
        cases[name + "_scan"] = s
        cases[name + "_edge"] = e
        cases[name + "_surf"] = u
        cases[name + "_lines"] = np.int32(lines)
        print(name, s.shape, len(e), len(u))
    np.savez_compressed(os.path.join(out, "extract_golden.npz"), **cases)

    p = pfb.synth.params(sensor_lines=64, azimuth_steps=600, seed=21)
    od = O.Odom(0.4, 0, 0.4, 75)
    poses, sizes = [], []
    for f in range(6):
        s = pfb.synth.scan(p, f)
        r = O.extract(s, order=1)
        e, u = s[r["edge_idx"]], s[r["surf_idx"]]
        if f == 0:
            od.init_map(e, u)
            continue
        poses.append(od.update(e, u))
        st = od.stats()
        sizes.append([st["n_edge_ds"], st["n_surf_ds"], st["n_edge_res"], st["n_surf_res"], st["map_edge"], st["map_surf"]])
    np.savez_compressed(os.path.join(out, "odom_oracle_golden.npz"), poses=np.array(poses), sizes=np.array(sizes, np.int32),
                        synth=np.array([64, 600, 21], np.int64))
    print("odom golden", np.array(poses)[-1], sizes[-1])


if __name__ == "__main__":
    main()
