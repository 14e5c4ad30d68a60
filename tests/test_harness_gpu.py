"""The compiled C++ harness (examples/harness.cpp): the reference's node call order through the header-only class wrappers
(LaserProcessingClass -> OdomEstimationClass -> LaserMappingClass) on the GPU; exit code 0 = tracked the ground truth."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_harness_runs():
    exe = os.path.join(ROOT, "examples", "harness")
    if not os.path.exists(exe):
        pytest.fail("examples/harness missing: run __graft_entry__.build()")
    r = subprocess.run([exe, "14"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    print(r.stdout)
    assert r.returncode == 0, r.stdout
    assert "14 frames" in r.stdout


def test_typed_adapter_sequence_matches_the_c_abi(pfb, capi):
    """tests/compat/node_sequence.cpp (reference-typed calls through compat_eigen_pcl.h) gives the poses and map sizes of the same frames
    through the plain C ABI."""
    import numpy as np
    exe = os.path.join(ROOT, "tests", "compat", "node_sequence")
    if not os.path.exists(exe):      # built by __graft_entry__.build() / the CPU compile test
        r = subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "oracle", "shim"), "-I", os.path.join(ROOT, "include"),
                            os.path.join(ROOT, "tests", "compat", "node_sequence.cpp"), "-o", exe,
                            os.path.join(ROOT, "pfilter-noetic_b200", "libpfilter_b200.so"), os.path.join(ROOT, "pfilter-noetic_b200", "libpf_synth.so"),
                            "-Wl,-rpath," + os.path.join(ROOT, "pfilter-noetic_b200")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout
    n = 6
    r = subprocess.run([exe, str(n)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    rows = [l.split() for l in r.stdout.splitlines() if l.startswith("frame")]
    assert len(rows) == n
    p = pfb.synth.config("cfg2")
    ex = capi.Extractor(num_lines=64, max_points=262144, surf_order=1)     # the adapter emits surf in the reference's order
    od = capi.Odometry(0.4, 0, 0.4, 75)
    for f in range(n):
        s = pfb.synth.scan(p, f)
        e, u, _ = ex.run(s, want_label=False)
        if f == 0:
            od.init_map(e, u)
            pose = np.array([0, 0, 0, 1, 0, 0, 0.0])
        else:
            pose = od.update(e, u)
        got = np.array([float(x) for x in rows[f][3:10]])
        # the adapter hands the pose out as an Isometry3d and the test reads it back through Quaterniond(rotation()): 1e-12
        assert np.abs(got - pose).max() < 1e-8, (f, got, pose)
        assert int(rows[f][11]) == len(e) and int(rows[f][13]) == len(u)
        assert int(rows[f][15]) == len(od.map_part(0)) and int(rows[f][17]) == len(od.map_part(1))
    ex.close(); od.close()


def test_packed_scans_through_the_frame_pipeline(pfb, capi, tmp_path):
    """F4 on the boundary's side: scan -> KITTI .bin -> pf_read_kitti_bin into a pinned buffer -> pf_frame_process, and scan ->
    PointCloud2 payload -> pf_pack_pointcloud2 -> pf_frame_process, give the poses of the direct call, bit for bit."""
    import numpy as np
    io = __import__("pfilter_noetic_b200.io", fromlist=["io"])
    p = pfb.synth.config("cfg2")
    scans = [pfb.synth.scan(p, f) for f in range(5)]

    def run(feed):
        ex = capi.Extractor(num_lines=64, max_points=131072)
        od = capi.Odometry(0.4, 0, 0.4, 75, max_map_points=1 << 19)
        out = np.array([capi.frame_process(ex, od, feed(k)) for k in range(len(scans))])
        ex.close(); od.close()
        return out
    direct = run(lambda k: scans[k])
    pin, ptr = capi.pinned_array((131072, 4), np.float32)
    for k, s in enumerate(scans):
        capi.write_kitti_bin(tmp_path / f"{k:06d}.bin", s)
    via_bin = run(lambda k: capi.read_kitti_bin(tmp_path / f"{k:06d}.bin", out=pin))
    assert via_bin.tobytes() == direct.tobytes()

    def via_msg(k):
        data, step, fields = io.xyzi_to_pointcloud2(scans[k])
        return capi.pack_pointcloud2(data, step, fields, len(scans[k]), out=pin)
    assert run(via_msg).tobytes() == direct.tobytes()
    capi.host_free(ptr)
